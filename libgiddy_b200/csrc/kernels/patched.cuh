// Patched decoders: a position-aligned decoder (Dictionary, Frame of Reference) and its a-posteriori
// patches applied in ONE pass.
//
// The reference patches a decompressed column with a second kernel (README.md:55-60): scatter
// (src/kernels/data_layout/scatter.cuh:26-46) or compressed-indices scatter
// (src/kernels/data_layout/compressed_indices_scatter.cuh:219-271).  Run on its own, a patch store
// dirties 4-8 bytes of a 32-byte sector somewhere in a multi-GB column: HBM has to read and rewrite
// the whole sector (measured: 32 + 32 bytes of DRAM traffic per patch), and both the reference's kernels
// and this library's stand-alone ones top out at the same ~23 G patches/s.  Here the patch is stored
// right after the vectors which contain it, by the same warp, while those lines are still dirty in L2 --
// the patch merges there and costs no HBM access of its own (measured: DRAM traffic = algorithmic bytes).
//
// Requirement (the only one beyond the two schemes' own): patch positions are in ASCENDING order, as a
// compressor which emits exceptions while scanning the column produces them.
//
// Decomposition: the output is cut into warp tiles of 32 * unroll vectors; every WARP owns a contiguous
// range of them and is an independent stream processor -- no CTA-wide barrier in the loop.  It seeks its
// first patch once, with a warp-wide 32-ary search (5 rounds for 5 M patches), and from then on walks the
// patch stream in step with its tiles:
//     request the NEXT tile's compressed bytes (software pipelining: two tiles of loads in flight per warp)
//     decode and store this tile's vectors
//     __syncwarp (orders those stores before the patch stores), ballot the candidates which fall inside
//     the tile, store them
// The candidate patches live in registers (index list: one candidate per lane, reloaded when all 32 are
// spent) or in a per-warp shared-memory window (compressed indices: 128 decoded positions).
#pragma once
#ifndef GIDDY_B200_KERNELS_PATCHED_CUH_
#define GIDDY_B200_KERNELS_PATCHED_CUH_

#include "elementwise.cuh"
#include "scatter.cuh"

namespace giddy {
namespace kernels {

namespace patched_detail {

constexpr unsigned unroll = GIDDY_ELEMENTWISE_UNROLL;
constexpr unsigned max_warps_per_block = GIDDY_PATCHED_BLOCK_THREADS / 32;
constexpr unsigned window_patches = GIDDY_PATCHED_WINDOW;       // decoded patch positions held in shared memory, per warp

// The smallest i in [0, count) with key(i) >= target, or count; key must be non-decreasing.
// Called by a whole warp with the same arguments.
template <typename KeyFunction>
__device__ __forceinline__ uint64_t warp_lower_bound(uint64_t count, uint64_t target, KeyFunction key)
{
	uint64_t lo = 0, hi = count; // the answer is in [lo, hi]
	while (lo < hi) {
		const uint64_t step = (hi - lo + warp_size - 1) / warp_size;
		const uint64_t pivot = lo + ((uint64_t) lane_index() + 1) * step - 1;
		const bool below = pivot < hi && key(pivot) < target;
		const uint64_t num_below = __popc(__ballot_sync(full_mask, below)); // keys are sorted: the pivots below the target are a prefix
		const uint64_t past_pivot = lo + (num_below + 1) * step - 1;
		lo += num_below * step;
		hi = minimum(hi, past_pivot);
	}
	return lo;
}

__device__ __forceinline__ uint64_t warp_sum(uint64_t x)
{
	#pragma unroll
	for (unsigned d = warp_size / 2; d > 0; d >>= 1) { x += __shfl_xor_sync(full_mask, (unsigned long long) x, d); }
	return x;
}

// ---- patch sources ---------------------------------------------------------------------------------

// Patches as a plain list: positions indices[i] (ascending), values data[i]; the arguments of scatter::scatter.
// Lane l holds candidate window_base + l; a second window is already on its way when the first runs out.
template <unsigned IndexSize, unsigned U>
struct index_list_patches {
	struct shared_t { };
	const uint_t<U>*          __restrict__ data;
	const uint_t<IndexSize>*  __restrict__ indices;
	uint64_t                               count;

	uint64_t           window_base;  // index of lane 0's candidate
	unsigned           spent;        // the window's candidates already applied: lanes [0, spent)
	uint_t<IndexSize>  position, next_position;
	uint_t<U>          value, next_value;

	__device__ __forceinline__ void load_next_window()
	{
		const uint64_t i = window_base + warp_size + lane_index();
		if (i < count) { next_position = __ldg(indices + i); next_value = __ldg(data + i); }
	}

	__device__ __forceinline__ void seek(uint64_t first_position, shared_t&)
	{
		window_base = 0;
		if (first_position != 0 && count != 0) {
			window_base = warp_lower_bound(count, first_position, [&](uint64_t i) { return (uint64_t) __ldg(indices + i); });
		}
		spent = 0;
		const uint64_t i = window_base + lane_index();
		if (i < count) { position = __ldg(indices + i); value = __ldg(data + i); }
		load_next_window();
	}

	__device__ __forceinline__ void prefetch() { }

	// Called by the whole warp after it has issued the stores of the tile ending at tile_end.
	__device__ __forceinline__ void apply(uint_t<U>* __restrict__ target, uint64_t tile_end, shared_t&)
	{
		__syncwarp(); // orders the tile's stores before the patch stores
		while (true) {
			const bool mine = lane_index() >= spent && window_base + lane_index() < count && position < tile_end;
			spent += __popc(__ballot_sync(full_mask, mine));
			if (mine) { target[position] = value; }
			if (spent < warp_size) { break; }
			window_base += warp_size; spent = 0;
			position = next_position; value = next_value;
			load_next_window();
			if (window_base >= count) { break; }
		}
	}
};

// Patches with run-compressed positions: the arguments of scatter::compressed_indices::scatter
// (SURVEY.md Appendix A.11).  Positions are decoded a window at a time into the warp's slice of shared
// memory -- 32 run descriptors (one per lane), a warp scan of the run lengths, then the patches, each lane
// locating its patches' runs by binary search -- and consumed from there as the tiles advance.
template <unsigned IndexSize, unsigned U, unsigned II, unsigned RL>
struct compressed_index_patches {
	struct shared_t {
		uint_t<IndexSize> positions[window_patches];
		uint64_t          run_baselines[warp_size];
		uint64_t          run_offset_starts[warp_size];  // byte offset of the run's first offset
		uint32_t          run_ends[warp_size];           // inclusive scan of the staged runs' (remaining, clamped) lengths
		uint8_t           run_widths[warp_size];         // bytes per offset
	};
	const uint_t<U>*          __restrict__ data;
	const uint_t<RL>*         __restrict__ run_lengths;
	const unsigned char*      __restrict__ run_offset_sizes;
	const uint_t<IndexSize>*  __restrict__ run_baselines;
	const uint_t<II>*         __restrict__ run_offsets_start_positions;
	const unsigned char*      __restrict__ offset_bytes;
	const uint_t<II>*         __restrict__ anchors;
	uint64_t                               anchoring_period;
	uint64_t                               num_runs;
	uint64_t                               count;          // number of patches

	// the stream's state; uniform over the warp
	uint64_t  cursor;        // the first patch not yet applied
	uint64_t  window_first;  // shared.positions[j] is the position of patch window_first + j, for window_first + j < window_end
	uint64_t  window_end;
	uint64_t  next_run;      // the first run with patches at or past window_end ...
	uint64_t  next_run_skip; // ... and how many of its patches are before window_end
	bool      exhausted;
	uint_t<U> my_value;      // data[cursor + lane] as of the last prefetch()

	__device__ __forceinline__ uint64_t first_position_of_run(uint64_t run) const
	{
		const unsigned width = (unsigned) __ldg(run_offset_sizes + run) / 8;
		const uint8_t* offsets = offset_bytes + (uint64_t) __ldg(run_offsets_start_positions + run) * IndexSize;
		return (uint64_t)(uint_t<IndexSize>)((uint64_t) __ldg(run_baselines + run) + ci_detail::load_offset(offsets, 0, width));
	}

	// Decode the positions of the next (up to) window_patches patches.  Whole warp.
	__device__ __forceinline__ void refill(shared_t& shared)
	{
		const unsigned lane = lane_index();
		const unsigned runs_here = (unsigned) minimum<uint64_t>(warp_size, num_runs - next_run);
		uint32_t length = 0;
		__syncwarp(); // every lane is done with the previous window
		if (lane < runs_here) {
			const uint64_t run = next_run + lane;
			uint64_t remaining = (uint64_t) __ldg(run_lengths + run);
			if (lane == 0) { remaining = remaining > next_run_skip ? remaining - next_run_skip : 0; }
			length = (uint32_t) minimum<uint64_t>(remaining, window_patches + 1);
			shared.run_baselines[lane] = (uint64_t) __ldg(run_baselines + run);
			shared.run_offset_starts[lane] = (uint64_t) __ldg(run_offsets_start_positions + run) * IndexSize;
			shared.run_widths[lane] = (uint8_t)(__ldg(run_offset_sizes + run) / 8);
		}
		const uint32_t end = warp_inclusive_sum(length);
		shared.run_ends[lane] = end;
		__syncwarp();
		const uint32_t staged_total = shuffle_from(end, warp_size - 1);
		const uint32_t fill = (uint32_t) minimum<uint64_t>(minimum<uint32_t>(staged_total, window_patches), count - window_end);
		for (uint32_t q = lane; q < fill; q += warp_size) {
			unsigned lo = 0, hi = runs_here - 1; // the first staged run whose end exceeds q
			while (lo < hi) {
				const unsigned mid = (lo + hi) / 2;
				if (shared.run_ends[mid] > q) { hi = mid; } else { lo = mid + 1; }
			}
			const uint32_t preceding = lo == 0 ? 0 : shared.run_ends[lo - 1];
			const uint64_t j = (uint64_t)(q - preceding) + (lo == 0 ? next_run_skip : 0);
			shared.positions[q] = (uint_t<IndexSize>)(shared.run_baselines[lo] +
				ci_detail::load_offset(offset_bytes + shared.run_offset_starts[lo], j, shared.run_widths[lo]));
		}
		const unsigned finished_runs = __popc(__ballot_sync(full_mask, lane < runs_here && end <= fill));
		if (finished_runs == 0) { next_run_skip += fill; }
		else { next_run_skip = finished_runs < runs_here ? fill - shuffle_from(end, finished_runs - 1) : 0; }
		next_run += finished_runs;
		window_first = window_end;
		window_end += fill;
		exhausted = fill == 0;
		__syncwarp(); // publishes shared.positions
	}

	// Apply (or, with Apply == false, skip) the patches with positions below `limit`.  Whole warp.
	template <bool Apply>
	__device__ __forceinline__ void consume(uint_t<U>* __restrict__ target, uint64_t limit, shared_t& shared, bool value_prefetched)
	{
		while (true) {
			if (cursor == window_end) {
				if (exhausted || window_end >= count) { break; }
				refill(shared);
				if (exhausted) { break; }
			}
			const uint64_t available = window_end - cursor;
			bool mine = lane_index() < available;
			uint_t<IndexSize> position = 0;
			if (mine) {
				position = shared.positions[cursor - window_first + lane_index()];
				mine = position < limit;
			}
			const unsigned matched = __popc(__ballot_sync(full_mask, mine));
			if (Apply && mine) { target[position] = value_prefetched ? my_value : __ldg(data + cursor + lane_index()); }
			cursor += matched;
			value_prefetched = false;
			if (matched < warp_size && matched < available) { break; } // the next patch is at or beyond the limit
		}
	}

	__device__ __forceinline__ void seek(uint64_t first_position, shared_t& shared)
	{
		cursor = window_first = window_end = next_run = next_run_skip = 0;
		exhausted = count == 0 || num_runs == 0;
		if (first_position == 0 || exhausted) { return; }
		const uint64_t num_periods = (count + anchoring_period - 1) / anchoring_period;
		const uint64_t period = warp_lower_bound(num_periods, first_position,
			[&](uint64_t a) { return first_position_of_run((uint64_t) __ldg(anchors + a)); });
		// the period before may hold patches at or beyond first_position in its later part; within it, so may
		// the last run which begins below first_position, while the runs before that one are skipped whole
		const uint64_t start = period == 0 ? 0 : period - 1;
		const uint64_t first_run = (uint64_t) __ldg(anchors + start);
		const uint64_t past_last_run = start + 1 < num_periods ? (uint64_t) __ldg(anchors + start + 1) : num_runs;
		const uint64_t runs_below = warp_lower_bound(past_last_run - first_run, first_position,
			[&](uint64_t i) { return first_position_of_run(first_run + i); });
		const uint64_t whole_runs = runs_below == 0 ? 0 : runs_below - 1;
		uint64_t skipped = 0;
		for (uint64_t i = lane_index(); i < whole_runs; i += warp_size) { skipped += (uint64_t) __ldg(run_lengths + first_run + i); }
		skipped = warp_sum(skipped);
		next_run = first_run + whole_runs;
		cursor = window_first = window_end = start * anchoring_period + skipped;
		consume<false>(nullptr, first_position, shared, false);
	}

	// Request this lane's candidate value for the coming tile (an L1 hit most of the time: the cursor advances
	// by a few patches per tile).
	__device__ __forceinline__ void prefetch()
	{
		const uint64_t i = cursor + lane_index();
		if (i < count) { my_value = __ldg(data + i); }
	}

	__device__ __forceinline__ void apply(uint_t<U>* __restrict__ target, uint64_t tile_end, shared_t& shared)
	{
		__syncwarp(); // orders the tile's stores before the patch stores
		consume<true>(target, tile_end, shared, true);
	}
};

// ---- decoders ----------------------------------------------------------------------------------------
// load(v): the compressed bytes of output vector v; decode(pack, v): its 16 decoded bytes;
// decode_one(i): element i on its own.  Same arithmetic as the stand-alone kernels in elementwise.cuh.

template <unsigned IndexSize, unsigned U, unsigned DI>
struct dictionary_decoder {
	static constexpr unsigned E = 16 / U;
	static constexpr unsigned in_bytes = E * DI;
	uint_t<U>*         __restrict__ decompressed;
	const uint_t<DI>*  __restrict__ compressed_input;
	const uint_t<U>*   __restrict__ table;     // global or shared
	bool                            table_in_shared_memory;

	__device__ __forceinline__ void setup(const uint_t<U>* __restrict__ dictionary_entries, uint64_t num_entries, bool cache)
	{
		table = dictionary_entries;
		table_in_shared_memory = cache;
		if (cache) {
			auto cached = reinterpret_cast<uint_t<U>*>(dynamic_shared_memory);
			for (uint64_t i = threadIdx.x; i < num_entries; i += blockDim.x) { cached[i] = __ldg(dictionary_entries + i); }
			__syncthreads();
			table = cached;
		}
	}
	__device__ __forceinline__ bool vectorizable() const { return is_aligned_to(decompressed, 16) && is_aligned_to(compressed_input, in_bytes); }
	__device__ __forceinline__ uint_t<U> lookup(uint_t<DI> index) const { return table_in_shared_memory ? table[index] : __ldg(table + index); }
	__device__ __forceinline__ byte_pack<in_bytes> load(uint64_t v) const { return load_pack<in_bytes>(compressed_input + v * E); }
	__device__ __forceinline__ byte_pack<16> decode(const byte_pack<in_bytes>& in, uint64_t) const
	{
		byte_pack<16> out;
		#pragma unroll
		for (unsigned j = 0; j < E; j++) { out.template set<U>(j, lookup(in.template get<DI>(j))); }
		return out;
	}
	__device__ __forceinline__ uint_t<U> decode_one(uint64_t i) const { return lookup(compressed_input[i]); }
};

template <unsigned IndexSize, unsigned U, unsigned C, int Model, unsigned ModelIndexSize>
struct frame_of_reference_decoder {
	using size_t_ = size_type<IndexSize>;
	static constexpr unsigned E = 16 / U;
	static constexpr unsigned in_bytes = E * C;
	static constexpr unsigned dim = (Model == model_linear) ? 2 : 1;
	uint_t<U>*       __restrict__ decompressed;
	const int_t<C>*  __restrict__ compressed_input;
	const int32_t*   __restrict__ coefficients;
	detail::segment_locator<size_t_> locator;

	__device__ __forceinline__ frame_of_reference_decoder(uint_t<U>* out, const int_t<C>* in, const int32_t* c, size_t_ segment_length) :
		decompressed(out), compressed_input(in), coefficients(c), locator(segment_length) { }
	__device__ __forceinline__ bool vectorizable() const { return is_aligned_to(decompressed, 16) && is_aligned_to(compressed_input, in_bytes); }
	__device__ __forceinline__ byte_pack<in_bytes> load(uint64_t v) const { return load_pack<in_bytes>(compressed_input + v * E); }
	__device__ __forceinline__ void load_coefficients(int32_t (&c)[dim], size_t_ s) const
	{
		#pragma unroll
		for (unsigned d = 0; d < dim; d++) { c[d] = __ldg(coefficients + (size_t) s * dim + d); }
	}
	__device__ __forceinline__ byte_pack<16> decode(const byte_pack<in_bytes>& in, uint64_t v) const
	{
		size_t_ s, x;
		locator.locate((size_t_)(v * E), s, x);
		int32_t c[dim];
		load_coefficients(c, s);
		byte_pack<16> out;
		#pragma unroll
		for (unsigned j = 0; j < E; j++) {
			const uint_t<U> predicted = detail::evaluate_model<Model, ModelIndexSize, int32_t, U>(c, x);
			const uint_t<U> offset = (uint_t<U>)(int_t<U>)(int_t<C>) in.template get<C>(j);
			out.template set<U>(j, (uint_t<U>)(predicted + offset));
			x++;
			if (x == locator.segment_length && j + 1 < E) { x = 0; s++; load_coefficients(c, s); } // L % E != 0
		}
		return out;
	}
	__device__ __forceinline__ uint_t<U> decode_one(uint64_t i) const
	{
		size_t_ s, x;
		locator.locate((size_t_) i, s, x);
		int32_t c[dim];
		load_coefficients(c, s);
		return (uint_t<U>)(detail::evaluate_model<Model, ModelIndexSize, int32_t, U>(c, x) + (uint_t<U>)(int_t<U>) compressed_input[i]);
	}
};

// ---- the tile loop -------------------------------------------------------------------------------------

template <unsigned U, typename Decoder, typename Patches>
__device__ __forceinline__ void decode_and_patch(const Decoder& decoder, Patches& patches, typename Patches::shared_t& shared, uint64_t length)
{
	constexpr unsigned E = Decoder::E;
	uint_t<U>* __restrict__ const decompressed = decoder.decompressed;
	const unsigned lane = lane_index();
	const bool vectorizable = decoder.vectorizable();
	const uint64_t elements_per_tile = (uint64_t) warp_size * unroll * (vectorizable ? E : 1);
	const uint64_t num_tiles = (length + elements_per_tile - 1) / elements_per_tile;
	const uint64_t first_tile = num_tiles * global_warp_index() / num_warps_in_grid();
	const uint64_t past_last_tile = num_tiles * (global_warp_index() + 1) / num_warps_in_grid();
	if (first_tile == past_last_tile) { return; }

	patches.seek(first_tile * elements_per_tile, shared);

	if (!vectorizable) {
		#pragma unroll 1
		for (uint64_t tile = first_tile; tile < past_last_tile; tile++) {
			const uint64_t tile_start = tile * elements_per_tile;
			const uint64_t tile_end = minimum(tile_start + elements_per_tile, length);
			patches.prefetch();
			#pragma unroll
			for (unsigned k = 0; k < unroll; k++) {
				const uint64_t i = tile_start + k * warp_size + lane;
				if (i < tile_end) { decompressed[i] = decoder.decode_one(i); }
			}
			patches.apply(decompressed, tile_end, shared);
		}
		return;
	}

	constexpr unsigned vectors_per_tile = warp_size * unroll;
	auto vectors_in_tile = [&](uint64_t tile) { // all of them but for the column's last tile
		return (unsigned)(minimum((tile + 1) * elements_per_tile, length) / E - tile * vectors_per_tile);
	};
	byte_pack<Decoder::in_bytes> in[unroll], in_next[unroll];
	auto request = [&](byte_pack<Decoder::in_bytes> (&packs)[unroll], uint64_t tile) {
		const unsigned num_vectors = vectors_in_tile(tile);
		#pragma unroll
		for (unsigned k = 0; k < unroll; k++) {
			const unsigned v = k * warp_size + lane;
			if (v < num_vectors) { packs[k] = decoder.load(tile * vectors_per_tile + v); }
		}
	};
	request(in_next, first_tile);

	#pragma unroll 1
	for (uint64_t tile = first_tile; tile < past_last_tile; tile++) {
		const uint64_t tile_start = tile * elements_per_tile;
		const uint64_t tile_end = minimum(tile_start + elements_per_tile, length);
		const uint64_t first_vector = tile * vectors_per_tile;
		const unsigned num_vectors = vectors_in_tile(tile);
		#pragma unroll
		for (unsigned k = 0; k < unroll; k++) { in[k] = in_next[k]; }
		if (tile + 1 < past_last_tile) { request(in_next, tile + 1); }
		patches.prefetch();
		#pragma unroll
		for (unsigned k = 0; k < unroll; k++) {
			const unsigned v = k * warp_size + lane;
			if (v < num_vectors) { store_pack_16(decompressed + (first_vector + v) * E, decoder.decode(in[k], first_vector + v)); }
		}
		// the column's < E trailing elements (in its last tile only)
		const uint64_t i = (first_vector + num_vectors) * E + lane;
		if (i < tile_end) { decompressed[i] = decoder.decode_one(i); }
		patches.apply(decompressed, tile_end, shared);
	}
}

} // namespace patched_detail

// -------------------------------------------------------------------------------------------------
// dictionary::decompress (reference src/kernels/decompression/dictionary.cuh:61-75) followed by
// scatter::scatter (src/kernels/data_layout/scatter.cuh:26-46) with OutputIndexSize = InputIndexSize =
// IndexSize, in one pass.  Parameters: the former's, then the latter's minus the target.
//   decompressed[i] = dictionary_entries[compressed_input[i]];  decompressed[patch_indices[k]] = patch_data[k]
// patch_indices must be ascending.
// -------------------------------------------------------------------------------------------------
template <unsigned IndexSize, unsigned U, unsigned DictionaryIndexSize>
__global__ void patched_dictionary_decompress(
	uint_t<U>*                          __restrict__  decompressed,
	const uint_t<DictionaryIndexSize>*  __restrict__  compressed_input,
	const uint_t<U>*                    __restrict__  dictionary_entries,
	size_type<IndexSize>                              length,
	size_type<DictionaryIndexSize>                    num_dictionary_entries,
	bool                                              cache_dictionary_in_shared_memory,
	const uint_t<U>*                    __restrict__  patch_data,
	const uint_t<IndexSize>*            __restrict__  patch_indices,
	size_type<IndexSize>                              num_patches)
{
	using patches_t = patched_detail::index_list_patches<IndexSize, U>;
	__shared__ typename patches_t::shared_t shared[patched_detail::max_warps_per_block];
	patched_detail::dictionary_decoder<IndexSize, U, DictionaryIndexSize> decoder;
	decoder.decompressed = decompressed;
	decoder.compressed_input = compressed_input;
	decoder.setup(dictionary_entries, num_dictionary_entries, cache_dictionary_in_shared_memory);
	patches_t patches;
	patches.data = patch_data; patches.indices = patch_indices; patches.count = num_patches;
	patched_detail::decode_and_patch<U>(decoder, patches, shared[warp_index_in_block()], length);
}

// -------------------------------------------------------------------------------------------------
// frame_of_reference::decompress (reference src/kernels/decompression/frame_of_reference.cuh:35-80)
// followed by scatter::compressed_indices::scatter (src/kernels/data_layout/compressed_indices_scatter.cuh:
// 219-271) with OutputIndexSize = IndexSize and RunLengthSize = InputIndexSize, in one pass.  Parameters: the
// former's, then the latter's minus the target.  Patch positions must be ascending over the whole patch list.
// -------------------------------------------------------------------------------------------------
template <unsigned IndexSize, unsigned U, unsigned C, int Model, unsigned ModelIndexSize, unsigned InputIndexSize>
__global__ void patched_frame_of_reference_decompress(
	uint_t<U>*                     __restrict__  decompressed,
	const int_t<C>*                __restrict__  compressed_input,
	const int32_t*                 __restrict__  segment_model_coefficients,
	size_type<IndexSize>                         length,
	size_type<IndexSize>                         segment_length,
	const uint_t<U>*               __restrict__  data_to_scatter,
	const uint_t<InputIndexSize>*  __restrict__  scatter_position_run_lengths,
	const unsigned char*           __restrict__  scatter_position_run_individual_offset_sizes,
	const uint_t<IndexSize>*       __restrict__  scatter_position_run_baseline_values,
	const uint_t<InputIndexSize>*  __restrict__  scatter_position_run_offsets_start_positions,
	const unsigned char*           __restrict__  scatter_position_offset_bytes,
	const uint_t<InputIndexSize>*  __restrict__  scatter_position_anchors,
	uint_t<InputIndexSize>                       anchoring_period,
	uint_t<InputIndexSize>                       num_scatter_position_runs,
	size_type<InputIndexSize>                    data_length)
{
	using patches_t = patched_detail::compressed_index_patches<IndexSize, U, InputIndexSize, InputIndexSize>;
	__shared__ typename patches_t::shared_t shared[patched_detail::max_warps_per_block];
	const patched_detail::frame_of_reference_decoder<IndexSize, U, C, Model, ModelIndexSize> decoder(
		decompressed, compressed_input, segment_model_coefficients, segment_length);
	patches_t patches;
	patches.data = data_to_scatter;
	patches.run_lengths = scatter_position_run_lengths;
	patches.run_offset_sizes = scatter_position_run_individual_offset_sizes;
	patches.run_baselines = scatter_position_run_baseline_values;
	patches.run_offsets_start_positions = scatter_position_run_offsets_start_positions;
	patches.offset_bytes = scatter_position_offset_bytes;
	patches.anchors = scatter_position_anchors;
	patches.anchoring_period = anchoring_period;
	patches.num_runs = num_scatter_position_runs;
	patches.count = data_length;
	patched_detail::decode_and_patch<U>(decoder, patches, shared[warp_index_in_block()], length);
}

} // namespace kernels
} // namespace giddy

#endif // GIDDY_B200_KERNELS_PATCHED_CUH_
