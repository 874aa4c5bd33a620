// Delta decoder: a segment-anchored inclusive prefix sum.
//
// One warp decodes one anchored segment.  A "row" is 32 lanes x E elements (E = 16 / sizeof
// (Uncompressed), so a lane's output is one 16-byte vector and the warp's is 512 contiguous
// bytes).  `rows_per_step` rows are loaded before any is scanned, so each lane keeps that
// many independent loads in flight and the row scans -- a serial scan of the lane's E deltas
// in registers plus one warp-shuffle scan of the lane totals -- overlap.  The only serial
// dependency is the running carry handed from row to row, one add per row.  Per-segment
// anchors make segments independent, so no inter-CTA communication is needed: the reference's
// block-wide scan with two barriers per 1024 elements (src/kernels/reduction/scan.cuh:345-361)
// becomes barrier-free.
#pragma once
#ifndef GIDDY_B200_KERNELS_DELTA_CUH_
#define GIDDY_B200_KERNELS_DELTA_CUH_

#include "common.cuh"

namespace giddy {
namespace kernels {

using namespace giddy::device;

namespace delta_detail {

constexpr unsigned rows_per_step = GIDDY_DELTA_ROWS_PER_STEP;

// Decode one segment with E elements per lane per row.  E > 1 requires the segment's first input
// element to be E*C-aligned and its first output element to be 16-byte aligned.
template <unsigned U, unsigned C, unsigned E>
__device__ __forceinline__ void decode_segment(
	uint_t<U>*        __restrict__  out,
	const uint_t<C>*  __restrict__  in,
	uint64_t                        length,
	uint_t<U>                       carry)
{
	const unsigned lane = lane_index();
	constexpr unsigned row_elements = warp_size * E;

	for (uint64_t step_start = 0; step_start < length; step_start += rows_per_step * row_elements) {
		byte_pack<E * C> deltas[rows_per_step];
		#pragma unroll
		for (unsigned r = 0; r < rows_per_step; r++) {
			const uint64_t p = step_start + r * row_elements + lane * E;
			if (p + E <= length) { deltas[r] = load_pack<E * C>(in + p); }
			else {
				#pragma unroll
				for (unsigned j = 0; j < E; j++) { deltas[r].template set<C>(j, (p + j < length) ? in[p + j] : (uint_t<C>) 0); }
			}
		}
		#pragma unroll
		for (unsigned r = 0; r < rows_per_step; r++) {
			const uint64_t p = step_start + r * row_elements + lane * E;
			if (step_start + r * row_elements >= length) { break; } // warp-uniform
			uint_t<U> local[E];
			uint_t<U> sum = 0;
			#pragma unroll
			for (unsigned j = 0; j < E; j++) {
				sum = (uint_t<U>)(sum + (uint_t<U>) deltas[r].template get<C>(j));
				local[j] = sum;
			}
			const uint_t<U> inclusive = warp_inclusive_sum(sum);
			const uint_t<U> before = (uint_t<U>)(carry + (uint_t<U>)(inclusive - sum));
			carry = (uint_t<U>)(carry + shuffle_from(inclusive, warp_size - 1));
			if (E > 1 && p + E <= length) {
				byte_pack<16> vec;
				#pragma unroll
				for (unsigned j = 0; j < E; j++) { vec.template set<U>(j, (uint_t<U>)(before + local[j])); }
				store_pack_16(out + p, vec);
			}
			else {
				#pragma unroll
				for (unsigned j = 0; j < E; j++) { if (p + j < length) { out[p + j] = (uint_t<U>)(before + local[j]); } }
			}
		}
	}
}

} // namespace delta_detail

// -------------------------------------------------------------------------------------------------
// Delta.  Replaces cuda::kernels::decompression::delta::decompress,
// reference src/kernels/decompression/delta.cuh:66-108 (same parameter list):
//   decompressed[i] = anchor_values[s] + sum_{j = s*L .. i} zero_extend(compressed_input[j])
// -------------------------------------------------------------------------------------------------
template <unsigned IndexSize, unsigned U, unsigned C>
__global__ void delta_decompress(
	uint_t<U>*        __restrict__  decompressed,
	const uint_t<C>*  __restrict__  compressed_input,
	const uint_t<U>*  __restrict__  anchor_values,
	size_type<IndexSize>            length,
	size_type<IndexSize>            segment_length)
{
	constexpr unsigned E = 16 / U;
	const uint64_t num_segments = ((uint64_t) length + segment_length - 1) / segment_length;
	const bool vectorizable = is_aligned_to(decompressed, 16) && is_aligned_to(compressed_input, E * C) &&
		(segment_length % E == 0);

	for (uint64_t segment = global_warp_index(); segment < num_segments; segment += num_warps_in_grid()) {
		const uint64_t start = segment * segment_length;
		const uint64_t this_length = minimum<uint64_t>(segment_length, length - start);
		const uint_t<U> anchor = __ldg(anchor_values + segment);
		if (vectorizable) { delta_detail::decode_segment<U, C, E>(decompressed + start, compressed_input + start, this_length, anchor); }
		else              { delta_detail::decode_segment<U, C, 1>(decompressed + start, compressed_input + start, this_length, anchor); }
	}
}

// -------------------------------------------------------------------------------------------------
// Delta with LONG segments (few anchors, down to a single one: a plain prefix sum over the column).
// A warp per segment cannot fill the GPU then, so segments are cut into CTA-sized tiles which are scanned
// in a single pass with decoupled look-back: every tile publishes its aggregate as soon as it has it, and
// obtains its exclusive prefix by summing its predecessors' aggregates -- 32 descriptors per warp step --
// back to the nearest tile whose inclusive prefix is already known.  The first tile of every segment knows
// its prefix at once (the segment's anchor), so a look-back never leaves its segment.
//
// Tiles are handed out through a ticket counter, in order, to persistent CTAs, so every tile's
// predecessors are running or done: the spin-waits cannot deadlock under any grid size.
//
// workspace: 32 zeroed bytes (the ticket counter) followed by one zeroed 32-byte descriptor per tile;
// the launcher takes it from the stream-ordered allocator (cudaMallocAsync) for the duration of the launch.
// -------------------------------------------------------------------------------------------------
namespace delta_detail {

struct alignas(32) tile_descriptor_t {
	unsigned long long aggregate;  // sum of the tile's deltas             (valid when status >= 1)
	unsigned long long inclusive;  // value of the tile's last element     (valid when status == 2)
	uint32_t           status;
	uint32_t           padding[3];
};
enum : uint32_t { tile_unknown = 0, tile_has_aggregate = 1, tile_has_inclusive = 2 };

__device__ __forceinline__ void store_release(uint32_t* p, uint32_t v)
{
	asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t load_acquire(const uint32_t* p)
{
	uint32_t v;
	asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ unsigned long long load_relaxed(const unsigned long long* p)
{
	unsigned long long v;
	asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ void store_relaxed(unsigned long long* p, unsigned long long v)
{
	asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}

// The exclusive prefix of tile `tile`, the `tile_in_segment`-th (>= 1) of its segment; called by one whole warp.
template <unsigned U>
__device__ __forceinline__ uint_t<U> look_back(tile_descriptor_t* descriptors, uint64_t tile, uint64_t tile_in_segment)
{
	const unsigned lane = lane_index();
	uint_t<U> exclusive = 0;
	uint64_t looked = 0; // predecessors already summed
	while (true) {
		const bool valid = looked + lane < tile_in_segment;
		uint32_t status = tile_unknown;
		tile_descriptor_t* d = descriptors + (tile - 1 - looked - lane);
		if (valid) { do { status = load_acquire(&d->status); } while (status == tile_unknown); }
		const unsigned with_inclusive = __ballot_sync(full_mask, valid && status == tile_has_inclusive);
		const unsigned nearest = with_inclusive ? (unsigned) __ffs((int) with_inclusive) - 1 : warp_size; // lanes beyond it do not count
		uint_t<U> mine = 0;
		if (valid && lane <= nearest) { mine = (uint_t<U>) load_relaxed(lane == nearest ? &d->inclusive : &d->aggregate); }
		#pragma unroll
		for (unsigned delta = warp_size / 2; delta > 0; delta >>= 1) {
			if (sizeof(uint_t<U>) <= 4) { mine = (uint_t<U>)(mine + (uint_t<U>) __shfl_xor_sync(full_mask, (uint32_t) mine, delta)); }
			else { mine = (uint_t<U>)(mine + (uint_t<U>) __shfl_xor_sync(full_mask, (unsigned long long) mine, delta)); }
		}
		exclusive = (uint_t<U>)(exclusive + mine);
		if (with_inclusive) { return exclusive; }
		looked += warp_size;
	}
}

template <unsigned U, unsigned C, unsigned E>
__device__ __forceinline__ void decode_tiles_with_look_back(
	uint_t<U>*        __restrict__  decompressed,
	const uint_t<C>*  __restrict__  compressed_input,
	const uint_t<U>*  __restrict__  anchor_values,
	uint64_t                        length,
	uint64_t                        segment_length,
	void*                           workspace)
{
	__shared__ uint_t<U> warp_totals[32];
	__shared__ uint_t<U> shared_tile_exclusive;
	__shared__ uint64_t  shared_ticket;

	const unsigned lane = lane_index(), warp = warp_index_in_block(), num_warps = warps_per_block();
	constexpr unsigned row_elements = warp_size * E;
	const uint64_t tile_elements = (uint64_t) num_warps * rows_per_step * row_elements;
	const uint64_t tiles_per_segment = (segment_length + tile_elements - 1) / tile_elements;
	const uint64_t num_segments = (length + segment_length - 1) / segment_length;
	const uint64_t last_segment_length = length - (num_segments - 1) * segment_length;
	const uint64_t num_tiles = (num_segments - 1) * tiles_per_segment + (last_segment_length + tile_elements - 1) / tile_elements;
	unsigned long long* const ticket_counter = static_cast<unsigned long long*>(workspace);
	tile_descriptor_t* const descriptors = reinterpret_cast<tile_descriptor_t*>(static_cast<unsigned char*>(workspace) + 32);

	while (true) {
		if (threadIdx.x == 0) { shared_ticket = atomicAdd(ticket_counter, 1ull); }
		__syncthreads();
		const uint64_t tile = shared_ticket;
		if (tile >= num_tiles) { return; }
		const uint64_t segment = tile / tiles_per_segment, tile_in_segment = tile - segment * tiles_per_segment;
		const uint64_t segment_start = segment * segment_length;
		const uint64_t this_segment_length = minimum(segment_length, length - segment_start);
		const uint64_t tile_offset = tile_in_segment * tile_elements;
		const uint64_t tile_start = segment_start + tile_offset;
		const uint64_t tile_length = minimum(tile_elements, this_segment_length - tile_offset);
		const uint64_t warp_offset = (uint64_t) warp * rows_per_step * row_elements;

		byte_pack<E * C> deltas[rows_per_step];
		#pragma unroll
		for (unsigned r = 0; r < rows_per_step; r++) {
			const uint64_t p = warp_offset + r * row_elements + lane * E;
			if (p + E <= tile_length) { deltas[r] = load_pack<E * C>(compressed_input + tile_start + p); }
			else {
				#pragma unroll
				for (unsigned j = 0; j < E; j++) { deltas[r].template set<C>(j, (p + j < tile_length) ? compressed_input[tile_start + p + j] : (uint_t<C>) 0); }
			}
		}
		// within the warp: what precedes each lane's elements of each row
		uint_t<U> before[rows_per_step];
		uint_t<U> warp_total = 0;
		#pragma unroll
		for (unsigned r = 0; r < rows_per_step; r++) {
			uint_t<U> sum = 0;
			#pragma unroll
			for (unsigned j = 0; j < E; j++) { sum = (uint_t<U>)(sum + (uint_t<U>) deltas[r].template get<C>(j)); }
			const uint_t<U> inclusive = warp_inclusive_sum(sum);
			before[r] = (uint_t<U>)(warp_total + (uint_t<U>)(inclusive - sum));
			warp_total = (uint_t<U>)(warp_total + shuffle_from(inclusive, warp_size - 1));
		}
		if (lane == 0) { warp_totals[warp] = warp_total; }
		__syncthreads();
		uint_t<U> preceding_warps = 0, tile_aggregate = 0;
		for (unsigned w = 0; w < num_warps; w++) {
			const uint_t<U> t = warp_totals[w];
			if (w < warp) { preceding_warps = (uint_t<U>)(preceding_warps + t); }
			tile_aggregate = (uint_t<U>)(tile_aggregate + t);
		}
		if (warp == 0) {
			tile_descriptor_t* mine = descriptors + tile;
			uint_t<U> exclusive;
			if (tile_in_segment == 0) { exclusive = __ldg(anchor_values + segment); }
			else {
				if (lane == 0) {
					store_relaxed(&mine->aggregate, (unsigned long long) tile_aggregate);
					store_release(&mine->status, tile_has_aggregate);
				}
				exclusive = look_back<U>(descriptors, tile, tile_in_segment);
			}
			if (lane == 0) {
				store_relaxed(&mine->inclusive, (unsigned long long)(uint_t<U>)(exclusive + tile_aggregate));
				store_release(&mine->status, tile_has_inclusive);
				shared_tile_exclusive = exclusive;
			}
		}
		__syncthreads();
		const uint_t<U> base = (uint_t<U>)(shared_tile_exclusive + preceding_warps);
		#pragma unroll
		for (unsigned r = 0; r < rows_per_step; r++) {
			const uint64_t p = warp_offset + r * row_elements + lane * E;
			uint_t<U> running = (uint_t<U>)(base + before[r]);
			if (E > 1 && p + E <= tile_length) {
				byte_pack<16> vec;
				#pragma unroll
				for (unsigned j = 0; j < E; j++) {
					running = (uint_t<U>)(running + (uint_t<U>) deltas[r].template get<C>(j));
					vec.template set<U>(j, running);
				}
				store_pack_16(decompressed + tile_start + p, vec);
			}
			else {
				#pragma unroll
				for (unsigned j = 0; j < E; j++) {
					running = (uint_t<U>)(running + (uint_t<U>) deltas[r].template get<C>(j));
					if (p + j < tile_length) { decompressed[tile_start + p + j] = running; }
				}
			}
		}
		__syncthreads(); // shared_ticket and warp_totals are rewritten by the next tile
	}
}

} // namespace delta_detail

template <unsigned IndexSize, unsigned U, unsigned C>
__global__ void delta_decompress_long_segments(
	uint_t<U>*        __restrict__  decompressed,
	const uint_t<C>*  __restrict__  compressed_input,
	const uint_t<U>*  __restrict__  anchor_values,
	size_type<IndexSize>            length,
	size_type<IndexSize>            segment_length,
	void*                           workspace)
{
	constexpr unsigned E = 16 / U;
	const bool vectorizable = is_aligned_to(decompressed, 16) && is_aligned_to(compressed_input, E * C) && (segment_length % E == 0);
	if (vectorizable) { delta_detail::decode_tiles_with_look_back<U, C, E>(decompressed, compressed_input, anchor_values, length, segment_length, workspace); }
	else              { delta_detail::decode_tiles_with_look_back<U, C, 1>(decompressed, compressed_input, anchor_values, length, segment_length, workspace); }
}

} // namespace kernels
} // namespace giddy

#endif // GIDDY_B200_KERNELS_DELTA_CUH_
