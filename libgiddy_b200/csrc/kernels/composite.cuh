// Fused composite decoders: two schemes of a chain decoded in one pass, so that the intermediate column never
// touches HBM.  These are extensions (the reference has no fused kernels; its README.md:55-60 describes
// composition as running the kernels one after the other); the per-scheme kernels remain available.
//
// Delta o Discard-Zero-Bytes-variable (BASELINE.json configs[3]): the deltas of a Delta-compressed column are
// themselves stored with their zero high bytes discarded.  The two-kernel path writes the widened deltas (C bytes
// each) and reads them back; here a warp stages a tile of the byte stream with TMA exactly as the DZB-variable
// kernel does (dzb_variable.cuh), and each row of 32 x E assembled deltas goes straight into the segment-anchored
// prefix sum of the Delta kernel (delta.cuh): in-register scan of the lane's E deltas, one warp-shuffle scan of
// the lane totals, a running carry from row to row.  A lane's E results are E * U contiguous bytes -- a multiple
// of 32 -- and leave with 256-bit stores (sm_100: STG.E.EF.ENL2.256), so every warp store covers whole sectors.
// HBM bytes per element: compressed bytes + size fields + U, against + 2 * C more for the two-kernel path.
#pragma once
#ifndef GIDDY_B200_KERNELS_COMPOSITE_CUH_
#define GIDDY_B200_KERNELS_COMPOSITE_CUH_

#include "dzb_variable.cuh"

namespace giddy {
namespace kernels {

namespace composite_detail {

using namespace giddy::device;

// The Delta stage as a sink of the DZB-variable decoder (see dzbv_detail::store_sink for the protocol).
template <unsigned U, unsigned C>
struct delta_sink {
	static_assert(U == 4 || U == 8, "the fused Delta stage produces 4- or 8-byte elements");
	static constexpr unsigned E = 16 / C;
	uint_t<U>* out;    // the segment's first output element
	uint_t<U>  carry;  // the value preceding the next row: the segment's anchor at first
	bool       wide;   // whether `out` is 32-byte aligned

	__device__ __forceinline__ void emit(uint64_t position, const byte_pack<16>& deltas, unsigned valid)
	{
		uint_t<U> local[E];
		uint_t<U> sum = 0;
		#pragma unroll
		for (unsigned j = 0; j < E; j++) {
			sum = (uint_t<U>)(sum + (uint_t<U>) deltas.template get<C>(j)); // (zeros from index `valid` on)
			local[j] = sum;
		}
		const uint_t<U> inclusive = warp_inclusive_sum(sum);
		const uint_t<U> before = (uint_t<U>)(carry + (uint_t<U>)(inclusive - sum));
		carry = (uint_t<U>)(carry + shuffle_from(inclusive, warp_size - 1));
		if (valid == E && wide) {
			constexpr unsigned words = E * U / 4;
			uint32_t w[words];
			#pragma unroll
			for (unsigned j = 0; j < E; j++) {
				const uint_t<U> value = (uint_t<U>)(before + local[j]);
				if (U == 8) { w[2 * j] = (uint32_t) value; w[2 * j + 1] = (uint32_t)((uint64_t) value >> 32); }
				else { w[j] = (uint32_t) value; }
			}
			#pragma unroll
			for (unsigned k = 0; k < words / 8; k++) {
				const uint32_t (&chunk)[8] = *reinterpret_cast<const uint32_t (*)[8]>(&w[8 * k]);
				store_streaming_32(reinterpret_cast<uint8_t*>(out + position) + 32 * k, chunk);
			}
		}
		else {
			#pragma unroll
			for (unsigned j = 0; j < E; j++) { if (j < valid) { out[position + j] = (uint_t<U>)(before + local[j]); } }
		}
	}
};

} // namespace composite_detail

#ifndef GIDDY_COMPOSITE_MIN_BLOCKS
#define GIDDY_COMPOSITE_MIN_BLOCKS 4
#endif

// -------------------------------------------------------------------------------------------------
//   d_i = the DZB-variable element i of (compressed_data, packed_element_sizes, position_anchors), C bytes wide
//   decompressed[i] = anchor_values[s] + sum_{j = s*L .. i} d_j        (s = i / L; mod 2^(8U))
// i.e. delta::decompress (reference src/kernels/decompression/delta.cuh:66-108) applied to the output of
// discard_zero_bytes::variable_width::decompress (discard_zero_bytes_variable.cuh:381-439), both with the same
// segment length.  Requirements (checked by the launcher): 1-, 2- or 4-bit size fields, segment_length *
// bits_per_element_size a multiple of 32, GIDDY_DZBV_SHARED_BYTES_PER_WARP(C) bytes of dynamic shared memory per warp.
// -------------------------------------------------------------------------------------------------
template <unsigned IndexSize, unsigned U, unsigned C, unsigned ElementSizesContainerSize>
__global__ void __launch_bounds__(GIDDY_SEGMENT_WARPS_PER_BLOCK * 32, GIDDY_COMPOSITE_MIN_BLOCKS) delta_of_discard_zero_bytes_variable_decompress(
	uint_t<U>*                                __restrict__  decompressed,
	const uint8_t*                            __restrict__  compressed_data,
	const uint_t<ElementSizesContainerSize>*  __restrict__  packed_element_sizes,
	const uint_t<IndexSize>*                  __restrict__  position_anchors,
	const uint_t<U>*                          __restrict__  anchor_values,
	size_type<IndexSize>                                    segment_length,
	size_type<IndexSize>                                    length,
	unsigned                                                min_represented_element_size,
	unsigned                                                bits_per_element_size)
{
	constexpr unsigned CS = ElementSizesContainerSize;
	constexpr unsigned E = 16 / C;
	using namespace dzbv_detail;
	using namespace giddy::device;
	const uint64_t num_segments = ((uint64_t) length + segment_length - 1) / segment_length;
	const unsigned bits = bits_per_element_size;

	const tile_storage<C> tile(warp_shared_memory(shared_memory_bytes_per_warp<C>()));
	if (lane_index() == 0) { mbarrier_init(tile.barrier, 1); }
	__syncwarp();
	uint32_t barrier_phase = 0;
	const uint32_t* field_words = reinterpret_cast<const uint32_t*>(packed_element_sizes);
	const uint64_t sizes_per_container = (8 * CS) / bits;
	const uint64_t num_field_words = (((uint64_t) length + sizes_per_container - 1) / sizes_per_container) * (CS / 4);
	const bool wide = is_aligned_to(decompressed, 32) && (((uint64_t) segment_length * U) % 32 == 0);

	for (uint64_t segment = global_warp_index(); segment < num_segments; segment += num_warps_in_grid()) {
		const uint64_t start = segment * segment_length;
		const uint64_t this_length = minimum<uint64_t>(segment_length, length - start);
		const uint64_t first_byte = __ldg(position_anchors + segment);
		const bool next_anchor_known = segment + 1 < num_segments;
		const uint64_t next_anchor = next_anchor_known ? (uint64_t) __ldg(position_anchors + segment + 1) : 0;
		composite_detail::delta_sink<U, C> sink { decompressed + start, __ldg(anchor_values + segment), wide };
		switch (bits) {
		case 1:
			decode_segment_staged<C, 1>(sink, compressed_data, first_byte, next_anchor_known, next_anchor,
				field_words, num_field_words, start, this_length, min_represented_element_size, tile, barrier_phase);
			break;
		case 2:
			decode_segment_staged<C, 2>(sink, compressed_data, first_byte, next_anchor_known, next_anchor,
				field_words, num_field_words, start, this_length, min_represented_element_size, tile, barrier_phase);
			break;
		default:
			decode_segment_staged<C, (E * 4 <= 32 ? 4 : 2)>(sink, compressed_data, first_byte, next_anchor_known, next_anchor,
				field_words, num_field_words, start, this_length, min_represented_element_size, tile, barrier_phase);
			break;
		}
	}
}

} // namespace kernels
} // namespace giddy

#endif // GIDDY_B200_KERNELS_COMPOSITE_CUH_
