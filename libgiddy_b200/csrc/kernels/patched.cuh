// Patched decoders: a position-aligned decoder (Dictionary, Frame of Reference) and its a-posteriori
// patches applied in ONE pass over the column.
//
// The reference patches a decompressed column with a second kernel (README.md:55-60): scatter
// (src/kernels/data_layout/scatter.cuh:26-46) or compressed-indices scatter
// (src/kernels/data_layout/compressed_indices_scatter.cuh:219-271).  Run on its own, a patch store
// dirties 4-8 bytes of a 32-byte sector somewhere in a multi-GB column: HBM has to read and rewrite
// the whole sector (measured: 32 + 32 bytes of DRAM traffic per patch), and random sectors come at ~25 G/s
// whatever the kernel does.  Storing the patch right after the vector which contains it, while the line is still
// dirty in L2, saves the HBM traffic (measured: DRAM bytes = algorithmic bytes) but not the time: a partial-sector
// write still costs L2 a read-modify-write, and three different kernel structures all ended at ~30 G patches/s.
// Here a vector is therefore patched BEFORE it is stored, in the registers of the lane which decoded it, and leaves
// the SM whole; shared memory only carries the routing information (which vectors are hit, by which patches).
//
// Requirement (the only one beyond the two schemes' own): patch positions are in ASCENDING order, as a
// compressor which emits exceptions while scanning the column produces them.
//
// Decomposition: the output is cut into warp tiles of 32 x GIDDY_PATCHED_UNROLL vectors, handed grid-stride to the
// warps of a grid of GIDDY_PATCHED_BLOCKS_PER_MULTIPROCESSOR CTAs per multiprocessor.  What a warp needs to know is
// which patches fall into its tile; that comes from a small index built just before, in stream order, into a
// workspace the launcher takes from the library's pool for the duration of the call:
//     (compressed positions only)  expand_compressed_indices: the run-compressed positions as a plain list
//     patch_tile_ranges: for every tile with patches, the range of the patch list which falls into it (one
//                        thread per patch, no loops, no search: a patch opens / closes its tile's range if its
//                        neighbour in the list lies in another tile; the array starts out zeroed)
// -- 8 bytes per tile (0.2-0.4 % of the output) plus, for compressed positions, 4 bytes per patch.  The decode
// kernel then, per tile: (has its range, requested an iteration ahead) requests its compressed vectors and (lane l)
// patch first + l; the lanes holding patches mark, in a 512-byte table of the warp's, the vectors they hit; every lane
// reads its 8 marks with one load, decodes its vectors, replaces the patched elements in registers and stores whole
// vectors.  (Round 1 parked all decoded vectors in 4 KiB of shared memory per warp and patched them there: 8 STS.128 +
// 8 LDS.128 per lane and tile, mio_throttle among the top stalls, 33 KiB of static shared memory per CTA; a CTA-tile
// variant with a 16 KiB staging buffer and two barriers per tile measured slower still: profiles/patched_r01_notes.md.)
#pragma once
#ifndef GIDDY_B200_KERNELS_PATCHED_CUH_
#define GIDDY_B200_KERNELS_PATCHED_CUH_

#include "elementwise.cuh"
#include "scatter.cuh"

namespace giddy {
namespace kernels {

namespace patched_detail {

constexpr unsigned unroll = GIDDY_PATCHED_UNROLL;
constexpr unsigned max_warps_per_block = GIDDY_ELEMENTWISE_BLOCK_THREADS / 32; // the launchers refuse larger blocks
// CTAs per multiprocessor the kernels are compiled for: 64-bit indices, and dictionaries of 1- and 2-byte elements (8-16
// lookups per vector), need more registers than the column shapes of 4- and 8-byte elements this is tuned for
constexpr unsigned blocks_per_multiprocessor(unsigned index_size, unsigned element_size)
{
	return (index_size > 4 || element_size < 4) ? (GIDDY_PATCHED_BLOCKS_PER_MULTIPROCESSOR > 3 ? 3 : GIDDY_PATCHED_BLOCKS_PER_MULTIPROCESSOR) : GIDDY_PATCHED_BLOCKS_PER_MULTIPROCESSOR;
}

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

// out[element] = value for an element number known only at run time, with masks rather than a branch or an indexed
// access per candidate (which the compiler turns into a round trip through local memory).
template <unsigned U>
__device__ __forceinline__ void replace_element(byte_pack<16>& out, unsigned element, uint_t<U> value)
{
	constexpr unsigned per_word = (U >= 4) ? 1 : 4 / U; // elements per 32-bit word
	#pragma unroll
	for (unsigned w = 0; w < 4; w++) {
		uint32_t mask, bits;
		if (U == 8) {
			mask = (element == w / 2) ? 0xffffffffu : 0u;
			bits = (w % 2) ? (uint32_t)((uint64_t) value >> 32) : (uint32_t) value;
		}
		else if (U == 4) {
			mask = (element == w) ? 0xffffffffu : 0u;
			bits = (uint32_t) value;
		}
		else {
			const unsigned shift = (element % per_word) * 8 * U;
			const uint32_t field = (U == 2) ? 0xffffu : 0xffu;
			mask = (element / per_word == w) ? (field << shift) : 0u;
			bits = (uint32_t) value << shift;
		}
		out.u32[w] = (out.u32[w] & ~mask) | (bits & mask);
	}
}

// ---- the tile index -----------------------------------------------------------------------------------

// ranges[t] = { first, past_last }: tile t's patches are [first, past_last) of the (ascending) patch list; tiles without
// patches keep the zeroes the launcher filled the array with.  One thread per patch, no loops: patch j opens its tile's
// range if its predecessor lies in an earlier tile and closes it if its successor lies in a later one.  Positions at or
// beyond `length` (a list which is not what it should be) open and close nothing and are never applied.
template <unsigned IndexSize>
struct alignas(2 * sizeof(size_type<IndexSize>)) patch_range_t { size_type<IndexSize> first, past_last; };

template <unsigned IndexSize>
__global__ void patch_tile_ranges(
	patch_range_t<IndexSize>*  __restrict__  ranges,
	const uint_t<IndexSize>*   __restrict__  positions,
	size_type<IndexSize>                     num_patches,
	unsigned                                 log2_tile_elements,
	uint64_t                                 length)
{
	for (uint64_t j = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; j < num_patches; j += (uint64_t) gridDim.x * blockDim.x) {
		const uint64_t position = __ldg(positions + j);
		if (position >= length) { continue; }
		const uint64_t tile = position >> log2_tile_elements;
		const bool opens = j == 0 || ((uint64_t) __ldg(positions + j - 1) >> log2_tile_elements) != tile;
		const bool closes = j + 1 == num_patches || ((uint64_t) __ldg(positions + j + 1) >> log2_tile_elements) != tile;
		if (opens) { ranges[tile].first = (size_type<IndexSize>) j; }
		if (closes) { ranges[tile].past_last = (size_type<IndexSize>)(j + 1); }
	}
}

// ---- the tile loop -------------------------------------------------------------------------------------

// What a warp keeps in shared memory to route a tile's patches to the lanes which own the patched vectors:
//   slots[l * unroll + k]   for vector k * 32 + l of the tile (lane l's k-th vector): 1 + the index, within the tile's patch
//                           range, of the first patch which falls into that vector, or 0 (16 bytes per lane: one LDS.128);
//   positions / values      of the tile's first 32 patches (the ones requested ahead), position relative to the tile.
// ~1 KiB per warp -- the decoded vectors themselves never leave the registers.
template <unsigned IndexSize, unsigned U>
struct alignas(16) patch_overlay_t {
	uint16_t             slots[warp_size * unroll];
	size_type<IndexSize> positions[warp_size];
	uint_t<U>            values[warp_size];
};

template <unsigned IndexSize, unsigned U, typename Decoder>
__device__ __forceinline__ void decode_and_patch(
	const Decoder&                                decoder,
	const uint_t<U>*             __restrict__     patch_values,
	const uint_t<IndexSize>*     __restrict__     patch_positions,
	const patch_range_t<IndexSize>*  __restrict__ tile_ranges,
	unsigned                                      log2_tile_elements,
	size_type<IndexSize>                          length)
{
	// (element, vector and tile numbers are kept in the column's own index width: 32-bit arithmetic -- and half the
	// registers -- for columns of up to 2^32 elements)
	using index_t = size_type<IndexSize>;
	constexpr unsigned E = Decoder::E;
	static_assert(unroll == 8, "a lane's slots are read as one 16-byte vector of 8 x 16 bits");
	uint_t<U>* __restrict__ const decompressed = decoder.decompressed;
	const unsigned lane = lane_index();
	const index_t elements_per_tile = (index_t) 1 << log2_tile_elements;
	const bool vectorizable = elements_per_tile == (index_t)(warp_size * unroll * E) && decoder.vectorizable(); // (the launcher sized the tiles by the same test)
	const index_t num_tiles = (index_t)(((uint64_t) length + elements_per_tile - 1) >> log2_tile_elements);
	constexpr unsigned vectors_per_tile = warp_size * unroll;
	__shared__ patch_overlay_t<IndexSize, U> overlay_of_warp[max_warps_per_block];
	patch_overlay_t<IndexSize, U>& overlay = overlay_of_warp[warp_index_in_block()];

	const uint64_t first_tile = global_warp_index(), tile_stride = num_warps_in_grid();
	if (first_tile >= num_tiles) { return; }

	if (!vectorizable) { // misaligned buffers: element by element, patches stored after their elements
		#pragma unroll 1
		for (uint64_t tile = first_tile; tile < num_tiles; tile += tile_stride) {
			const uint64_t tile_start = tile << log2_tile_elements;
			const uint64_t tile_end = minimum<uint64_t>(tile_start + elements_per_tile, length);
			const patch_range_t<IndexSize> range = tile_ranges[tile];
			#pragma unroll
			for (unsigned k = 0; k < unroll; k++) {
				const uint64_t i = tile_start + k * warp_size + lane;
				if (i < tile_end) { decompressed[i] = decoder.decode_one(i); }
			}
			if (range.past_last > range.first) {
				__syncwarp();
				for (uint64_t q = (uint64_t) range.first + lane; q < range.past_last; q += warp_size) {
					const uint64_t position = __ldg(patch_positions + q);
					// (a position outside the tile can only come from a list which is not ascending: ignored, not written wild)
					if (position - tile_start < tile_end - tile_start) { decompressed[position] = __ldg(patch_values + q); }
				}
			}
		}
		return;
	}

	// A warp takes tiles w, w + W, ...  Two of them are in flight at any time: while tile t is decoded, patched and stored,
	// the compressed vectors and the first 32 patches of tile t + W are on their way (and the patch range of tile t + 2W,
	// which those requests will need) -- the kernel has registers for few warps, so each must keep more bytes in flight.
	struct pending_t {
		index_t                       tile;
		patch_range_t<IndexSize>      range;
		byte_pack<Decoder::in_bytes>  in[unroll];
		uint_t<IndexSize>             position;
		uint_t<U>                     value;
	};
	auto request = [&](pending_t& p) {
		const index_t tile_start = p.tile << log2_tile_elements;
		const index_t tile_end = (index_t) minimum<uint64_t>((uint64_t) tile_start + elements_per_tile, length);
		const index_t first_vector = p.tile * vectors_per_tile;
		const unsigned num_vectors = (unsigned)(tile_end / E - first_vector);
		#pragma unroll
		for (unsigned k = 0; k < unroll; k++) {
			const unsigned v = k * warp_size + lane;
			if (v < num_vectors) { p.in[k] = decoder.load(first_vector + v); }
		}
		p.position = 0; p.value = 0;
		if (p.range.past_last > p.range.first && lane < p.range.past_last - p.range.first) {
			p.position = __ldg(patch_positions + p.range.first + lane); p.value = __ldg(patch_values + p.range.first + lane);
		}
	};
	auto finish = [&](const pending_t& p) {
		const index_t tile_start = p.tile << log2_tile_elements;
		const index_t tile_end = (index_t) minimum<uint64_t>((uint64_t) tile_start + elements_per_tile, length);
		const index_t first_vector = p.tile * vectors_per_tile;
		const unsigned num_vectors = (unsigned)(tile_end / E - first_vector); // all of them but for the column's last tile
		const index_t first_patch = p.range.first, past_last_patch = p.range.past_last;
		const index_t num_patches = past_last_patch > first_patch ? past_last_patch - first_patch : 0; // uniform over the warp
		if (num_vectors == vectors_per_tile) {
			// A patch stored to global memory after the vector which contains it is a partial write of a 32-byte sector:
			// it merges in L2 (no HBM traffic) but each one still costs L2 a read-modify-write.  So a vector is patched
			// BEFORE it is stored, in the registers of the lane which owns it.  What that lane needs to know -- which of its
			// vectors are hit, and by which patches -- goes through the warp's overlay: the lanes holding the patches mark
			// the slot of the vector they hit, every lane reads its 8 slots with one 16-byte load.  Patches are ascending,
			// so those of one vector are neighbours in the list: only the first of them marks the slot, and the owner walks
			// on from there.
			uint4 my_slots = make_uint4(0u, 0u, 0u, 0u);
			if (num_patches != 0) {
				reinterpret_cast<uint4*>(overlay.slots)[lane] = my_slots;
				__syncwarp();
				// (a position outside the tile can only come from a list which is not ascending: ignored, not written wild)
				index_t in_tile = (index_t)(p.position - tile_start);
				bool valid = lane < num_patches && in_tile < elements_per_tile;
				if (!valid) { in_tile = ~(index_t) 0; }
				unsigned vector = (unsigned)(in_tile / E);
				overlay.positions[lane] = in_tile;
				overlay.values[lane] = p.value;
				unsigned previous = __shfl_up_sync(full_mask, vector, 1);
				if (valid && (lane == 0 || previous != vector)) { overlay.slots[(vector % warp_size) * unroll + vector / warp_size] = (uint16_t)(lane + 1); }
				// tiles with more than 32 patches: the rest of the list, 32 at a time (their owners will fetch them from global memory)
				for (index_t base = warp_size; base < num_patches; base += warp_size) {
					const unsigned carried = __shfl_sync(full_mask, vector, warp_size - 1);
					const index_t q = base + lane;
					in_tile = ~(index_t) 0;
					if (q < num_patches) { in_tile = (index_t)(__ldg(patch_positions + first_patch + q) - tile_start); }
					valid = in_tile < elements_per_tile;
					vector = valid ? (unsigned)(in_tile / E) : ~0u;
					previous = __shfl_up_sync(full_mask, vector, 1);
					if (lane == 0) { previous = carried; }
					if (valid && previous != vector && q < 0xffffu) { overlay.slots[(vector % warp_size) * unroll + vector / warp_size] = (uint16_t)(q + 1); }
				}
				__syncwarp();
				my_slots = reinterpret_cast<const uint4*>(overlay.slots)[lane];
			}
			const uint32_t slot_words[4] = { my_slots.x, my_slots.y, my_slots.z, my_slots.w };
			#pragma unroll
			for (unsigned k = 0; k < unroll; k++) {
				byte_pack<16> out = decoder.decode(p.in[k], first_vector + k * warp_size + lane);
				const unsigned slot = (slot_words[k / 2] >> (16 * (k % 2))) & 0xffffu;
				if (slot != 0) {
					const unsigned my_vector = k * warp_size + lane;
					index_t q = slot - 1;
					do {
						const index_t in_tile = q < warp_size ? overlay.positions[q] : (index_t)(__ldg(patch_positions + first_patch + q) - tile_start);
						if (in_tile / E != my_vector) { break; }
						const uint_t<U> value = q < warp_size ? overlay.values[q] : __ldg(patch_values + first_patch + q);
						replace_element<U>(out, (unsigned)(in_tile % E), value);
						q++;
					} while (q < num_patches);
				}
				store_pack_16(decompressed + (uint64_t)(first_vector + k * warp_size + lane) * E, out);
			}
			if (num_patches != 0) { __syncwarp(); } // the next tile's patches go to the same overlay
		}
		else {
			#pragma unroll
			for (unsigned k = 0; k < unroll; k++) {
				const unsigned v = k * warp_size + lane;
				if (v < num_vectors) { store_pack_16(decompressed + (uint64_t)(first_vector + v) * E, decoder.decode(p.in[k], first_vector + v)); }
			}
			// the column's < E trailing elements (in its last tile only)
			const uint64_t i = (uint64_t)(first_vector + num_vectors) * E + lane;
			if (i < tile_end) { decompressed[i] = decoder.decode_one(i); }
			if (num_patches != 0) { // the column's last, partial tile: patches stored after their vectors
				__syncwarp(); // orders the tile's stores before the patch stores
				if (lane < num_patches && (index_t)(p.position - tile_start) < tile_end - tile_start) { decompressed[p.position] = p.value; }
				for (uint64_t q = (uint64_t) first_patch + warp_size + lane; q < past_last_patch; q += warp_size) {
					const index_t position = __ldg(patch_positions + q);
					if ((index_t)(position - tile_start) < tile_end - tile_start) { decompressed[position] = __ldg(patch_values + q); }
				}
			}
		}
	};

	pending_t current, next;
	current.tile = (index_t) first_tile;
	current.range = tile_ranges[first_tile];
	request(current);
	bool has_next = first_tile + tile_stride < num_tiles;
	patch_range_t<IndexSize> range_after = current.range;
	if (has_next) { range_after = tile_ranges[first_tile + tile_stride]; }
	#pragma unroll 1
	while (true) {
		bool has_after = false;
		if (has_next) {
			next.tile = current.tile + (index_t) tile_stride;
			next.range = range_after;
			request(next);
			has_after = (uint64_t) next.tile + tile_stride < num_tiles;
			if (has_after) { range_after = tile_ranges[next.tile + (index_t) tile_stride]; }
		}
		finish(current);
		if (!has_next) { break; }
		current = next;
		has_next = has_after;
	}
}

} // namespace patched_detail

// -------------------------------------------------------------------------------------------------
// dictionary::decompress (reference src/kernels/decompression/dictionary.cuh:61-75) followed by
// scatter::scatter (src/kernels/data_layout/scatter.cuh:26-46) with OutputIndexSize = InputIndexSize =
// IndexSize, in one pass.  Parameters: the former's, then the latter's minus the target, then the tile index.
//   decompressed[i] = dictionary_entries[compressed_input[i]];  decompressed[patch_indices[k]] = patch_data[k]
// patch_indices must be ascending.
// -------------------------------------------------------------------------------------------------
template <unsigned IndexSize, unsigned U, unsigned DictionaryIndexSize>
__global__ void __launch_bounds__(GIDDY_ELEMENTWISE_BLOCK_THREADS, patched_detail::blocks_per_multiprocessor(IndexSize, U)) patched_dictionary_decompress(
	uint_t<U>*                          __restrict__  decompressed,
	const uint_t<DictionaryIndexSize>*  __restrict__  compressed_input,
	const uint_t<U>*                    __restrict__  dictionary_entries,
	size_type<IndexSize>                              length,
	size_type<DictionaryIndexSize>                    num_dictionary_entries,
	bool                                              cache_dictionary_in_shared_memory,
	const uint_t<U>*                    __restrict__  patch_data,
	const uint_t<IndexSize>*            __restrict__  patch_indices,
	const patched_detail::patch_range_t<IndexSize>*  __restrict__  tile_ranges,
	unsigned                                          log2_tile_elements)
{
	patched_detail::dictionary_decoder<IndexSize, U, DictionaryIndexSize> decoder;
	decoder.decompressed = decompressed;
	decoder.compressed_input = compressed_input;
	decoder.setup(dictionary_entries, num_dictionary_entries, cache_dictionary_in_shared_memory);
	patched_detail::decode_and_patch<IndexSize, U>(decoder, patch_data, patch_indices, tile_ranges, log2_tile_elements, length);
}

// -------------------------------------------------------------------------------------------------
// frame_of_reference::decompress (reference src/kernels/decompression/frame_of_reference.cuh:35-80)
// followed by scatter::compressed_indices::scatter (src/kernels/data_layout/compressed_indices_scatter.cuh:
// 219-271) with OutputIndexSize = IndexSize and RunLengthSize = InputIndexSize, in one pass.  The latter's
// positions arrive decoded (expand_compressed_indices, kernels/scatter.cuh) and must be ascending over the whole
// patch list.
// -------------------------------------------------------------------------------------------------
template <unsigned IndexSize, unsigned U, unsigned C, int Model, unsigned ModelIndexSize>
__global__ void __launch_bounds__(GIDDY_ELEMENTWISE_BLOCK_THREADS, patched_detail::blocks_per_multiprocessor(IndexSize, U)) patched_frame_of_reference_decompress(
	uint_t<U>*                   __restrict__  decompressed,
	const int_t<C>*              __restrict__  compressed_input,
	const int32_t*               __restrict__  segment_model_coefficients,
	size_type<IndexSize>                       length,
	size_type<IndexSize>                       segment_length,
	const uint_t<U>*             __restrict__  data_to_scatter,
	const uint_t<IndexSize>*     __restrict__  scatter_positions,
	const patched_detail::patch_range_t<IndexSize>*  __restrict__  tile_ranges,
	unsigned                                   log2_tile_elements)
{
	const patched_detail::frame_of_reference_decoder<IndexSize, U, C, Model, ModelIndexSize> decoder(
		decompressed, compressed_input, segment_model_coefficients, segment_length);
	patched_detail::decode_and_patch<IndexSize, U>(decoder, data_to_scatter, scatter_positions, tile_ranges, log2_tile_elements, length);
}

} // namespace kernels
} // namespace giddy

#endif // GIDDY_B200_KERNELS_PATCHED_CUH_
