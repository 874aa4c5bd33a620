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

constexpr unsigned rows_per_step = 4;

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

} // namespace kernels
} // namespace giddy

#endif // GIDDY_B200_KERNELS_DELTA_CUH_
