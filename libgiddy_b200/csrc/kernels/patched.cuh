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
// Decomposition: the output is cut into warp tiles of 32 x GIDDY_PATCHED_UNROLL 32-byte pieces, handed grid-stride to the
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
// A lane's unit of work is a PIECE: 32 consecutive output bytes (E = 32 / U elements), decoded from E consecutive
// compressed elements and stored with one 256-bit access -- the patched decoders are bound by instruction issue, not by
// bytes, so every load, store and address computation should move as many bytes as the hardware lets it.
// load(p): the compressed bytes of piece p; decode(pack, p): its 32 decoded bytes; decode_one(i): element i on its own.
// Same arithmetic as the stand-alone kernels in elementwise.cuh.

constexpr unsigned piece_bytes = 32;
constexpr unsigned early_tiles = GIDDY_PATCHED_EARLY_TILES; // tiles a warp decodes before it needs the tile index

__device__ __forceinline__ void store_piece(void* p, const byte_pack<piece_bytes>& v)
{
	store_streaming_32(p, v.u32);
}

template <unsigned IndexSize, unsigned U, unsigned DI, bool Cached>
struct dictionary_decoder {
	static constexpr unsigned E = piece_bytes / U;
	static constexpr unsigned in_bytes = E * DI;
	uint_t<U>*         __restrict__ decompressed;
	const uint_t<DI>*  __restrict__ compressed_input;
	const uint_t<U>*   __restrict__ table;     // shared memory if Cached, else global

	__device__ __forceinline__ bool vectorizable() const { return is_aligned_to(decompressed, piece_bytes) && is_aligned_to(compressed_input, in_bytes < 16 ? in_bytes : 16); }
	__device__ __forceinline__ uint_t<U> lookup(uint_t<DI> index) const { return Cached ? table[index] : __ldg(table + index); }
	__device__ __forceinline__ byte_pack<in_bytes> load(uint64_t p) const { return load_pack<in_bytes>(compressed_input + p * E); }
	__device__ __forceinline__ byte_pack<piece_bytes> decode(const byte_pack<in_bytes>& in, uint64_t) const
	{
		byte_pack<piece_bytes> out;
		#pragma unroll
		for (unsigned j = 0; j < E; j++) { out.template set<U>(j, lookup(in.template get<DI>(j))); }
		return out;
	}
	__device__ __forceinline__ uint_t<U> decode_one(uint64_t i) const { return lookup(compressed_input[i]); }
};

template <unsigned IndexSize, unsigned U, unsigned C, int Model, unsigned ModelIndexSize>
struct frame_of_reference_decoder {
	using size_t_ = size_type<IndexSize>;
	static constexpr unsigned E = piece_bytes / U;
	static constexpr unsigned in_bytes = E * C;
	static constexpr unsigned dim = (Model == model_linear) ? 2 : 1;
	uint_t<U>*       __restrict__ decompressed;
	const int_t<C>*  __restrict__ compressed_input;
	const int32_t*   __restrict__ coefficients;
	detail::segment_locator<size_t_> locator;

	__device__ __forceinline__ frame_of_reference_decoder(uint_t<U>* out, const int_t<C>* in, const int32_t* c, size_t_ segment_length) :
		decompressed(out), compressed_input(in), coefficients(c), locator(segment_length) { }
	__device__ __forceinline__ bool vectorizable() const { return is_aligned_to(decompressed, piece_bytes) && is_aligned_to(compressed_input, in_bytes < 16 ? in_bytes : 16); }
	__device__ __forceinline__ byte_pack<in_bytes> load(uint64_t p) const { return load_pack<in_bytes>(compressed_input + p * E); }
	__device__ __forceinline__ void load_coefficients(int32_t (&c)[dim], size_t_ s) const
	{
		#pragma unroll
		for (unsigned d = 0; d < dim; d++) { c[d] = __ldg(coefficients + (size_t) s * dim + d); }
	}
	__device__ __forceinline__ byte_pack<piece_bytes> decode(const byte_pack<in_bytes>& in, uint64_t p) const
	{
		size_t_ s, x;
		locator.locate((size_t_)(p * E), s, x);
		int32_t c[dim];
		load_coefficients(c, s);
		byte_pack<piece_bytes> out;
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

// ---- the tile index -----------------------------------------------------------------------------------

// ranges[t] = { first, past_last }: tile t's patches are [first, past_last) of the (ascending) patch list; tiles without
// patches keep what the launcher filled the array with (all ones: first == past_last).  One thread per patch, no loops: patch j opens its tile's
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
	launch_dependent_grids(); // (the decoder may be scheduled as this kernel drains; it waits for the index on the device)
	// a thread takes 4 consecutive patches (one 16-byte load of 32-bit positions) and their two neighbours
	constexpr unsigned per_thread = 4;
	const bool wide = IndexSize == 4 && is_aligned_to(positions, 16);
	for (uint64_t j0 = ((uint64_t) blockIdx.x * blockDim.x + threadIdx.x) * per_thread; j0 < num_patches; j0 += (uint64_t) gridDim.x * blockDim.x * per_thread) {
		uint64_t p[per_thread + 2]; // p[1 + u] = position of patch j0 + u; p[0], p[5]: the neighbours
		const unsigned count = (unsigned) minimum<uint64_t>(per_thread, num_patches - j0);
		if (wide && count == per_thread) {
			const uint4 v = __ldg(reinterpret_cast<const uint4*>(positions + j0));
			p[1] = v.x; p[2] = v.y; p[3] = v.z; p[4] = v.w;
		}
		else {
			#pragma unroll
			for (unsigned u = 0; u < per_thread; u++) { p[1 + u] = (u < count) ? (uint64_t) __ldg(positions + j0 + u) : ~0ull; }
		}
		p[0] = (j0 == 0) ? ~0ull : (uint64_t) __ldg(positions + j0 - 1);
		p[per_thread + 1] = (j0 + per_thread < num_patches) ? (uint64_t) __ldg(positions + j0 + per_thread) : ~0ull;
		#pragma unroll
		for (unsigned u = 0; u < per_thread; u++) {
			if (u >= count || p[1 + u] >= length) { continue; }
			const uint64_t tile = p[1 + u] >> log2_tile_elements;
			const uint64_t previous = (u == 0) ? p[0] : p[u];
			const uint64_t next = (u + 1 < count) ? p[2 + u] : ((u + 1 == per_thread) ? p[per_thread + 1] : ~0ull);
			const bool opens = (j0 + u == 0) || (previous >> log2_tile_elements) != tile;
			const bool closes = (j0 + u + 1 == num_patches) || (next >> log2_tile_elements) != tile;
			if (opens) { ranges[tile].first = (size_type<IndexSize>)(j0 + u); }
			if (closes) { ranges[tile].past_last = (size_type<IndexSize>)(j0 + u + 1); }
		}
	}
}

// Every tile's range set to "empty" (all ones), by a kernel rather than a memset node so that the kernel after it can be its
// programmatic dependent: the positions are decoded while this runs.
template <unsigned Unused = 0>
__global__ void fill_tile_index(uint4* __restrict__ index, uint64_t num_vectors)
{
	launch_dependent_grids();
	for (uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; i < num_vectors; i += (uint64_t) gridDim.x * blockDim.x) {
		index[i] = make_uint4(~0u, ~0u, ~0u, ~0u);
	}
}

// The run-compressed positions of a compressed-indices patch list (reference compressed_indices_scatter.cuh:133-231) as a
// plain list -- positions[k] is where patch value k goes -- AND the tile index of those positions, in one kernel, for lists
// whose anchoring period is short enough for a period's positions to sit in shared memory (the launcher checks).  A BLOCK
// per anchoring period: a column's patch list has too few periods (2.6 K for BASELINE configs[4]) to fill the device with a
// warp each.  The run metadata of up to 32 runs is staged at a time (one run per lane of the first warp, a warp scan of the
// run lengths); the runs' patches are then one flat list, each patch finding its run by a 5-step search of the scanned
// lengths.  Everything per patch is 32-bit arithmetic: the kernel is bound by instruction issue, not by bytes.
// (Lists with longer periods go through scatter.cuh's expand_compressed_indices and patch_tile_ranges.)
template <unsigned OutputIndexSize, unsigned InputIndexSize, unsigned RunLengthSize>
__global__ void expand_and_index_compressed_indices(
	uint_t<OutputIndexSize>*        __restrict__  positions,
	patch_range_t<OutputIndexSize>* __restrict__  ranges,
	const uint_t<RunLengthSize>*    __restrict__  scatter_position_run_lengths,
	const unsigned char*            __restrict__  scatter_position_run_individual_offset_sizes,
	const uint_t<OutputIndexSize>*  __restrict__  scatter_position_run_baseline_values,
	const uint_t<InputIndexSize>*   __restrict__  scatter_position_run_offsets_start_positions,
	const unsigned char*            __restrict__  scatter_position_offset_bytes,
	const uint_t<InputIndexSize>*   __restrict__  scatter_position_anchors,
	unsigned                                      anchoring_period,
	uint64_t                                      num_scatter_position_runs,
	uint64_t                                      data_length,
	unsigned                                      log2_tile_elements,
	uint64_t                                      length)
{
	launch_dependent_grids(); // (the decoder may be scheduled as this kernel drains; it waits for the index on the device)
	using position_t = uint_t<OutputIndexSize>;
	constexpr unsigned batch = 8; // patches per lane with their offset loads in flight at once
	__shared__ uint32_t   ends[warp_size];           // inclusive scan of the staged runs' lengths
	__shared__ position_t baselines[warp_size];
	__shared__ uint32_t   offset_starts[warp_size];  // byte offset of the run's first offset, relative to the chunk's first run's
	__shared__ uint8_t    widths[warp_size];         // bytes per offset
	__shared__ uint64_t   chunk_offsets_base;
	__shared__ position_t neighbours[2];             // positions of the patch before the period's first and after its last
	__shared__ bool       has_neighbour[2];
	position_t* const period_positions = reinterpret_cast<position_t*>(dynamic_shared_memory); // anchoring_period entries
	const unsigned lane = lane_index(), warp = warp_index_in_block();
	const uint64_t num_periods = (data_length + anchoring_period - 1) / anchoring_period;
	for (uint64_t period = blockIdx.x; period < num_periods; period += gridDim.x) {
		const uint64_t period_first = period * anchoring_period;
		const unsigned period_patches = (unsigned) minimum<uint64_t>(anchoring_period, data_length - period_first);
		const uint64_t first_run = __ldg(scatter_position_anchors + period);
		const uint64_t past_last_run = (period + 1 == num_periods) ? num_scatter_position_runs : (uint64_t) __ldg(scatter_position_anchors + period + 1);
		if (threadIdx.x >= blockDim.x - 2) {
			// the neighbours across the period's ends, decoded from their runs (runs never straddle a period): the last patch of
			// run first_run - 1, the first of run past_last_run -- by two threads of the LAST warp, overlapping the staging
			const bool before = threadIdx.x == blockDim.x - 2;
			const bool exists = before ? first_run > 0 : past_last_run < num_scatter_position_runs;
			position_t position = 0;
			if (exists) {
				const uint64_t run = before ? first_run - 1 : past_last_run;
				const uint64_t j = before ? (uint64_t) __ldg(scatter_position_run_lengths + run) - 1 : 0;
				position = (position_t)((uint64_t) __ldg(scatter_position_run_baseline_values + run) + ci_detail::load_offset(
					scatter_position_offset_bytes + (uint64_t) __ldg(scatter_position_run_offsets_start_positions + run) * OutputIndexSize, j,
					__ldg(scatter_position_run_individual_offset_sizes + run) / 8));
			}
			neighbours[before ? 0 : 1] = position;
			has_neighbour[before ? 0 : 1] = exists;
		}
		unsigned done = 0; // patches of the period decoded so far
		for (uint64_t chunk_first_run = first_run; chunk_first_run < past_last_run && done < period_patches; chunk_first_run += warp_size) {
			__syncthreads(); // every warp is done with the previous chunk
			if (warp == 0) {
				const uint64_t run = chunk_first_run + lane;
				const bool valid = run < past_last_run;
				const uint32_t my_length = valid ? (uint32_t) minimum<uint64_t>(__ldg(scatter_position_run_lengths + run), anchoring_period) : 0u;
				const uint64_t start = valid ? (uint64_t) __ldg(scatter_position_run_offsets_start_positions + run) * OutputIndexSize : 0;
				const uint64_t chunk_start = shuffle_from(start, 0);
				widths[lane]        = valid ? (uint8_t)(__ldg(scatter_position_run_individual_offset_sizes + run) / 8) : (uint8_t) 1;
				baselines[lane]     = valid ? (position_t) __ldg(scatter_position_run_baseline_values + run) : (position_t) 0;
				offset_starts[lane] = (uint32_t)(start - chunk_start);
				ends[lane] = warp_inclusive_sum(my_length); // (lanes without a run repeat the total: the search never lands on them)
				if (lane == 0) { chunk_offsets_base = chunk_start; }
			}
			__syncthreads();
			const unsigned chunk_patches = minimum<unsigned>(ends[warp_size - 1], period_patches - done);
			const uint8_t* __restrict__ const chunk_offsets = scatter_position_offset_bytes + chunk_offsets_base;
			for (unsigned q0 = warp * warp_size * batch; q0 < chunk_patches; q0 += warps_per_block() * warp_size * batch) {
				position_t decoded[batch];
				#pragma unroll
				for (unsigned u = 0; u < batch; u++) {
					const unsigned q = q0 + u * warp_size + lane;
					if (q < chunk_patches) {
						unsigned r = 0; // the first staged run whose end exceeds q
						#pragma unroll
						for (unsigned step = warp_size / 2; step > 0; step >>= 1) { if (ends[r + step - 1] <= q) { r += step; } }
						const unsigned j = q - (r == 0 ? 0u : ends[r - 1]);
						const unsigned width = widths[r];
						const uint8_t* const at = chunk_offsets + (offset_starts[r] + j * width);
						position_t offset;
						if (width == 1) { offset = __ldg(at); }
						else if (width == 2) { offset = __ldg(reinterpret_cast<const uint16_t*>(at)); }
						else if (width == 4 || OutputIndexSize <= 4) { offset = (position_t) __ldg(reinterpret_cast<const uint32_t*>(at)); }
						else { offset = (position_t)((uint64_t) __ldg(reinterpret_cast<const uint32_t*>(at)) | ((uint64_t) __ldg(reinterpret_cast<const uint32_t*>(at) + 1) << 32)); }
						decoded[u] = (position_t)(baselines[r] + offset);
					}
				}
				#pragma unroll
				for (unsigned u = 0; u < batch; u++) {
					const unsigned q = q0 + u * warp_size + lane;
					if (q < chunk_patches) { positions[period_first + done + q] = decoded[u]; period_positions[done + q] = decoded[u]; }
				}
			}
			done += chunk_patches;
		}
		// (runs which cover fewer patches than the period holds -- a malformed list -- leave the rest at a position beyond any
		// column: never applied, as the reference's kernel would skip them)
		for (unsigned k = done + threadIdx.x; k < period_patches; k += blockDim.x) { positions[period_first + k] = ~(position_t) 0; }
		// the tile index for this period's patches (see patch_tile_ranges): patch k opens its tile's range if its predecessor lies
		// in an earlier tile and closes it if its successor lies in a later one.  (The index was filled with "empty" by the kernel
		// before this one, whose programmatic dependent this kernel is: nothing of it has been touched so far.)
		__syncthreads();
		wait_for_prerequisite_grids();
		for (unsigned k = threadIdx.x; k < done; k += blockDim.x) {
			const position_t position = period_positions[k];
			if ((uint64_t) position >= length) { continue; }
			const position_t tile = position >> log2_tile_elements;
			const bool opens = (k == 0) ? (!has_neighbour[0] || (neighbours[0] >> log2_tile_elements) != tile) : (period_positions[k - 1] >> log2_tile_elements) != tile;
			const bool closes = (k + 1 == done) ? (!has_neighbour[1] || (neighbours[1] >> log2_tile_elements) != tile) : (period_positions[k + 1] >> log2_tile_elements) != tile;
			if (opens) { ranges[tile].first = (size_type<OutputIndexSize>)(period_first + k); }
			if (closes) { ranges[tile].past_last = (size_type<OutputIndexSize>)(period_first + k + 1); }
		}
		__syncthreads(); // before the next period overwrites the positions
	}
}

// ---- the tile loop -------------------------------------------------------------------------------------

template <unsigned IndexSize, unsigned U, typename Decoder>
__device__ __forceinline__ void decode_and_patch(
	const Decoder&                                decoder,
	const uint_t<U>*             __restrict__     patch_values,
	const uint_t<IndexSize>*     __restrict__     patch_positions,
	const patch_range_t<IndexSize>*  __restrict__ tile_ranges,
	unsigned                                      log2_tile_elements,
	size_type<IndexSize>                          length)
{
	// (element, piece and tile numbers are kept in the column's own index width: 32-bit arithmetic -- and half the
	// registers -- for columns of up to 2^32 elements)
	using index_t = size_type<IndexSize>;
	constexpr unsigned E = Decoder::E; // elements per piece
	uint_t<U>* __restrict__ const decompressed = decoder.decompressed;
	const unsigned lane = lane_index();
	const index_t elements_per_tile = (index_t) 1 << log2_tile_elements;
	const bool vectorizable = elements_per_tile == (index_t)(warp_size * unroll * E) && decoder.vectorizable(); // (the launcher sized the tiles by the same test)
	const index_t num_tiles = (index_t)(((uint64_t) length + elements_per_tile - 1) >> log2_tile_elements);
	constexpr unsigned pieces_per_tile = warp_size * unroll;

	const uint64_t first_tile = global_warp_index(), tile_stride = num_warps_in_grid();
	if (first_tile >= num_tiles) { return; }

	if (!vectorizable) { // misaligned buffers: element by element, patches stored after their elements
		wait_for_prerequisite_grids(); // (the tile index is another kernel's: see below)
		#pragma unroll 1
		for (uint64_t tile = first_tile; tile < num_tiles; tile += tile_stride) {
			const uint64_t tile_start = tile << log2_tile_elements;
			const uint64_t tile_end = minimum<uint64_t>(tile_start + elements_per_tile, length);
			const patch_range_t<IndexSize> range = tile_ranges[tile];
			for (uint64_t i = tile_start + lane; i < tile_end; i += warp_size) { decompressed[i] = decoder.decode_one(i); }
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

	// Lane q holds patch q of the tile (position and value, requested a tile ahead) and stores it right after the tile's pieces
	// have been stored: the sector it lands in is still dirty in L2, where the partial write merges, so DRAM sees the patched
	// sector once.  (Measured alternatives, profiles/patched_r02_notes.md: patching the pieces in registers before they are
	// stored -- by shuffles to the owning lane, or through shared memory -- costs 15-30 instructions per patch in a kernel which is
	// bound by instruction issue, against ~1 for the store.)
	struct patches_t { uint_t<IndexSize> position; uint_t<U> value; };
	const patch_range_t<IndexSize> no_patches = { 0, 0 };
	auto request_patches = [&](const patch_range_t<IndexSize>& range, index_t chunk_first) {
		patches_t p = { 0, 0 };
		if (range.past_last > range.first && chunk_first + lane < range.past_last - range.first) {
			p.position = __ldg(patch_positions + range.first + chunk_first + lane);
			p.value = __ldg(patch_values + range.first + chunk_first + lane);
		}
		return p;
	};

	// The kernel is launched as a programmatic dependent of the kernel which builds the tile index (abi/abi_patched.cu): a warp
	// decodes and stores its first tiles without looking at the index -- work which overlaps that kernel's tail instead of
	// waiting behind it --, then waits for the index, applies those tiles' patches, and carries on with the pipelined loop.
	index_t tile = (index_t) first_tile;
	unsigned early = 0;
	#pragma unroll 1
	for (; early < early_tiles && (uint64_t) tile < num_tiles; early++, tile += (index_t) tile_stride) {
		const index_t tile_start = tile << log2_tile_elements;
		const index_t first_piece = tile * pieces_per_tile;
		const bool whole = tile + 1 != num_tiles || (uint64_t) tile_start + elements_per_tile == length;
		const index_t tile_end = whole ? tile_start + elements_per_tile : length;
		const unsigned num_pieces = whole ? pieces_per_tile : (unsigned)(tile_end / E - first_piece);
		byte_pack<Decoder::in_bytes> in[unroll];
		#pragma unroll
		for (unsigned k = 0; k < unroll; k++) {
			const unsigned v = k * warp_size + lane;
			if (v < num_pieces) { in[k] = decoder.load(first_piece + v); }
		}
		#pragma unroll
		for (unsigned k = 0; k < unroll; k++) {
			const unsigned v = k * warp_size + lane;
			if (v < num_pieces) { store_piece(decompressed + (uint64_t)(first_piece + v) * E, decoder.decode(in[k], first_piece + v)); }
		}
		if (!whole) {
			const uint64_t i = (uint64_t)(first_piece + num_pieces) * E + lane;
			if (i < tile_end) { decompressed[i] = decoder.decode_one(i); }
		}
	}
	wait_for_prerequisite_grids();
	__syncwarp();
	{
		index_t patched = (index_t) first_tile;
		#pragma unroll 1
		for (unsigned e = 0; e < early; e++, patched += (index_t) tile_stride) {
			const patch_range_t<IndexSize> early_range = tile_ranges[patched];
			const index_t tile_start = patched << log2_tile_elements;
			const index_t tile_length = (index_t) minimum<uint64_t>(elements_per_tile, (uint64_t) length - tile_start);
			if (early_range.past_last <= early_range.first) { continue; } // (an empty range is all ones: no arithmetic on it)
			for (uint64_t q = (uint64_t) early_range.first + lane; q < early_range.past_last; q += warp_size) {
				const index_t position = __ldg(patch_positions + q);
				if ((index_t)(position - tile_start) < tile_length) { decompressed[position] = __ldg(patch_values + q); }
			}
		}
	}
	if ((uint64_t) tile >= num_tiles) { return; }
	patch_range_t<IndexSize> range = tile_ranges[tile];
	bool has_next = (uint64_t) tile + tile_stride < num_tiles;
	patch_range_t<IndexSize> range_next = has_next ? tile_ranges[tile + (index_t) tile_stride] : no_patches;
	patches_t patches = request_patches(range, 0);

	#pragma unroll 1
	while (true) {
		const index_t tile_start = tile << log2_tile_elements;
		const index_t first_piece = tile * pieces_per_tile;
		const bool whole = tile + 1 != num_tiles || (uint64_t) tile_start + elements_per_tile == length;
		const index_t tile_end = whole ? tile_start + elements_per_tile : length;
		const unsigned num_pieces = whole ? pieces_per_tile : (unsigned)(tile_end / E - first_piece); // all of them but for the column's last tile
		// 1. the tile's compressed elements
		byte_pack<Decoder::in_bytes> in[unroll];
		#pragma unroll
		for (unsigned k = 0; k < unroll; k++) {
			const unsigned v = k * warp_size + lane;
			if (v < num_pieces) { in[k] = decoder.load(first_piece + v); }
		}
		// 2. one tile ahead: the next tile's first 32 patches; two ahead: the range of the tile after it
		patches_t patches_next = { 0, 0 };
		patch_range_t<IndexSize> range_after = no_patches;
		bool has_after = false;
		if (has_next) {
			patches_next = request_patches(range_next, 0);
			has_after = (uint64_t) tile + 2 * tile_stride < num_tiles;
			if (has_after) { range_after = tile_ranges[tile + (index_t)(2 * tile_stride)]; }
		}
		// 3. decode
		byte_pack<piece_bytes> out[unroll];
		#pragma unroll
		for (unsigned k = 0; k < unroll; k++) {
			const unsigned v = k * warp_size + lane;
			if (v < num_pieces) { out[k] = decoder.decode(in[k], first_piece + v); }
		}
		const index_t num_patches = range.past_last > range.first ? range.past_last - range.first : 0; // uniform over the warp
		// 4. store
		#pragma unroll
		for (unsigned k = 0; k < unroll; k++) {
			const unsigned v = k * warp_size + lane;
			if (v < num_pieces) { store_piece(decompressed + (uint64_t)(first_piece + v) * E, out[k]); }
		}
		// 5. the patches, after the pieces they fall into (ordered behind them by the warp barrier): partial writes of sectors which
		// L2 still holds, merged there -- no HBM traffic
		if (!whole) {
			const uint64_t i = (uint64_t)(first_piece + num_pieces) * E + lane;
			if (i < tile_end) { decompressed[i] = decoder.decode_one(i); }
		}
		if (num_patches != 0) {
			__syncwarp();
			if (lane < num_patches && (index_t)(patches.position - tile_start) < tile_end - tile_start) { decompressed[patches.position] = patches.value; }
			for (uint64_t q = (uint64_t) range.first + warp_size + lane; q < range.past_last; q += warp_size) { // (tiles with more than 32 patches)
				const index_t position = __ldg(patch_positions + q);
				if ((index_t)(position - tile_start) < tile_end - tile_start) { decompressed[position] = __ldg(patch_values + q); }
			}
		}
		if (!has_next) { break; }
		tile += (index_t) tile_stride;
		range = range_next; range_next = range_after;
		patches = patches_next;
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
	if (cache_dictionary_in_shared_memory) {
		auto cached = reinterpret_cast<uint_t<U>*>(dynamic_shared_memory);
		for (uint64_t i = threadIdx.x; i < num_dictionary_entries; i += blockDim.x) { cached[i] = __ldg(dictionary_entries + i); }
		__syncthreads();
		const patched_detail::dictionary_decoder<IndexSize, U, DictionaryIndexSize, true> decoder { decompressed, compressed_input, cached };
		patched_detail::decode_and_patch<IndexSize, U>(decoder, patch_data, patch_indices, tile_ranges, log2_tile_elements, length);
	}
	else {
		const patched_detail::dictionary_decoder<IndexSize, U, DictionaryIndexSize, false> decoder { decompressed, compressed_input, dictionary_entries };
		patched_detail::decode_and_patch<IndexSize, U>(decoder, patch_data, patch_indices, tile_ranges, log2_tile_elements, length);
	}
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
