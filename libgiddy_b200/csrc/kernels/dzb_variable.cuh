// Discard-Zero-Bytes, variable width (a.k.a. null suppression, variable): every element is
// stored in size_i of its low bytes, the sizes come from fixed-width bit fields packed into
// 32- or 64-bit containers, and every segment has the byte offset of its first element as
// an anchor.
//
// One warp decodes one anchored segment, a "row" of 32 lanes x E elements at a time (E = 16 /
// sizeof(Uncompressed): one 16-byte output vector per lane, 512 contiguous bytes per warp
// store).  For a row, each lane extracts its E element sizes from the size containers (one
// container load per lane when the E fields share a container), the lane totals go through
// one warp-shuffle exclusive scan to give every lane its byte offset, and each element is
// assembled from the (at most three) aligned 32-bit words which hold its bytes with a funnel
// shift -- no byte-granular loads, no local-memory staging (the reference memcpy's byte by
// byte, src/kernels/decompression/discard_zero_bytes_variable.cuh:196-213).  The byte stream
// of a row is contiguous, so neighbouring lanes read neighbouring words and the stream is
// fetched from HBM once and then served by L1.  Several rows are in flight per warp.
#pragma once
#ifndef GIDDY_B200_KERNELS_DZB_VARIABLE_CUH_
#define GIDDY_B200_KERNELS_DZB_VARIABLE_CUH_

#include "common.cuh"

namespace giddy {
namespace kernels {

using namespace giddy::device;

namespace dzbv_detail {

// Where the size field of an element lives; reference get_element_size,
// src/kernels/decompression/discard_zero_bytes_variable.cuh:44-58.
template <unsigned CS>
struct size_field_locator {
	unsigned bits;               // bits per size field
	unsigned sizes_per_container;
	unsigned log2_sizes_per_container;
	bool     power_of_2;
	uint_t<CS> field_mask;
	unsigned min_size;

	__device__ __forceinline__ size_field_locator(unsigned bits_, unsigned min_size_) : bits(bits_), min_size(min_size_)
	{
		sizes_per_container = (8 * CS) / bits;
		power_of_2 = (sizes_per_container & (sizes_per_container - 1)) == 0;
		log2_sizes_per_container = 31 - __clz((int) sizes_per_container);
		field_mask = (bits >= 8 * CS) ? ~(uint_t<CS>) 0 : (uint_t<CS>)(((uint_t<CS>) 1 << bits) - 1);
	}

	__device__ __forceinline__ void locate(uint64_t element, uint64_t& container, unsigned& field) const
	{
		if (power_of_2) {
			container = element >> log2_sizes_per_container;
			field = (unsigned)(element & (sizes_per_container - 1));
		}
		else if ((element >> 32) == 0) {
			uint32_t q = (uint32_t) element / sizes_per_container;
			container = q; field = (uint32_t) element - q * sizes_per_container;
		}
		else {
			container = element / sizes_per_container;
			field = (unsigned)(element - container * sizes_per_container);
		}
	}
};

// The `size` low bytes at byte address p, zero-extended.  Reads only aligned 32-bit words which
// contain at least one of the addressed bytes.
template <unsigned U>
__device__ __forceinline__ uint_t<U> load_low_bytes(const uint8_t* __restrict__ p, unsigned size)
{
	if (size == 0) { return 0; }
	const uintptr_t address = reinterpret_cast<uintptr_t>(p);
	const uint32_t* words = reinterpret_cast<const uint32_t*>(address & ~(uintptr_t) 3);
	const unsigned misalignment = (unsigned)(address & 3);
	const unsigned shift = misalignment * 8;
	const unsigned span = misalignment + size; // bytes from the first word's start to the element's end
	const uint32_t w0 = __ldg(words);
	const uint32_t w1 = (span > 4) ? __ldg(words + 1) : 0u;
	if (U <= 4) {
		uint32_t v = __funnelshift_r(w0, w1, shift);
		if (size < 4) { v &= (1u << (8 * size)) - 1u; }
		return (uint_t<U>) v;
	}
	const uint32_t w2 = (span > 8) ? __ldg(words + 2) : 0u;
	uint64_t v = (uint64_t) __funnelshift_r(w0, w1, shift) | ((uint64_t) __funnelshift_r(w1, w2, shift) << 32);
	if (size < 8) { v &= ((uint64_t) 1 << (8 * size)) - 1u; }
	return (uint_t<U>) v;
}

constexpr unsigned rows_per_step = 2;

template <unsigned U, unsigned CS, unsigned E, typename Offset>
__device__ __forceinline__ void decode_segment(
	uint_t<U>*         __restrict__  out,             // the segment's first output element
	const uint8_t*     __restrict__  segment_bytes,   // the segment's first compressed byte
	const uint_t<CS>*  __restrict__  packed_element_sizes,
	uint64_t                         first_element,   // global index of the segment's first element
	uint64_t                         length,
	const size_field_locator<CS>&    locator)
{
	const unsigned lane = lane_index();
	constexpr unsigned row_elements = warp_size * E;
	Offset row_offset = 0; // byte offset, within the segment, of the current row

	for (uint64_t step_start = 0; step_start < length; step_start += rows_per_step * row_elements) {
		unsigned sizes[rows_per_step][E];
		Offset totals[rows_per_step];
		#pragma unroll
		for (unsigned r = 0; r < rows_per_step; r++) {
			const uint64_t p = step_start + r * row_elements + lane * E;
			uint64_t container_index = ~(uint64_t) 0;
			uint_t<CS> container = 0;
			Offset total = 0;
			#pragma unroll
			for (unsigned j = 0; j < E; j++) {
				unsigned size = 0;
				if (p + j < length) {
					uint64_t c; unsigned field;
					locator.locate(first_element + p + j, c, field);
					if (c != container_index) { container_index = c; container = __ldg(packed_element_sizes + c); }
					size = (unsigned)((container >> (field * locator.bits)) & locator.field_mask) + locator.min_size;
				}
				sizes[r][j] = size;
				total += size;
			}
			totals[r] = total;
		}
		#pragma unroll
		for (unsigned r = 0; r < rows_per_step; r++) {
			const uint64_t p = step_start + r * row_elements + lane * E;
			if (step_start + r * row_elements >= length) { break; } // warp-uniform
			const Offset inclusive = warp_inclusive_sum(totals[r]);
			Offset offset = row_offset + (inclusive - totals[r]);
			row_offset += shuffle_from(inclusive, warp_size - 1);
			uint_t<U> elements[E];
			#pragma unroll
			for (unsigned j = 0; j < E; j++) {
				elements[j] = load_low_bytes<U>(segment_bytes + offset, sizes[r][j]);
				offset += sizes[r][j];
			}
			if (E > 1 && p + E <= length) {
				byte_pack<16> vec;
				#pragma unroll
				for (unsigned j = 0; j < E; j++) { vec.template set<U>(j, elements[j]); }
				store_pack_16(out + p, vec);
			}
			else {
				#pragma unroll
				for (unsigned j = 0; j < E; j++) { if (p + j < length) { out[p + j] = elements[j]; } }
			}
		}
	}
}

} // namespace dzbv_detail

// -------------------------------------------------------------------------------------------------
// Replaces cuda::kernels::decompression::discard_zero_bytes::variable_width::decompress,
// reference src/kernels/decompression/discard_zero_bytes_variable.cuh:381-439 (same parameters).
//   size_i = min_represented_element_size + field_i;  element i occupies size_i bytes right after
//   element i-1;  segment s's first byte is compressed_data[position_anchors[s]];
//   decompressed[i] = those bytes, zero-extended (little-endian).
// -------------------------------------------------------------------------------------------------
template <unsigned IndexSize, unsigned U, unsigned ElementSizesContainerSize>
__global__ void discard_zero_bytes_variable_decompress(
	uint_t<U>*                                __restrict__  decompressed,
	const uint8_t*                            __restrict__  compressed_data,
	const uint_t<ElementSizesContainerSize>*  __restrict__  packed_element_sizes,
	const uint_t<IndexSize>*                  __restrict__  position_anchors,
	size_type<IndexSize>                                    segment_length,
	size_type<IndexSize>                                    length,
	unsigned                                                min_represented_element_size,
	unsigned                                                bits_per_element_size)
{
	constexpr unsigned CS = ElementSizesContainerSize;
	constexpr unsigned E = 16 / U;
	using offset_t = size_type<IndexSize>;
	const dzbv_detail::size_field_locator<CS> locator(bits_per_element_size, min_represented_element_size);
	const uint64_t num_segments = ((uint64_t) length + segment_length - 1) / segment_length;
	const bool vectorizable = is_aligned_to(decompressed, 16) && (segment_length % E == 0);

	for (uint64_t segment = global_warp_index(); segment < num_segments; segment += num_warps_in_grid()) {
		const uint64_t start = segment * segment_length;
		const uint64_t this_length = minimum<uint64_t>(segment_length, length - start);
		const uint8_t* segment_bytes = compressed_data + __ldg(position_anchors + segment);
		if (vectorizable) {
			dzbv_detail::decode_segment<U, CS, E, offset_t>(decompressed + start, segment_bytes, packed_element_sizes, start, this_length, locator);
		}
		else {
			dzbv_detail::decode_segment<U, CS, 1, offset_t>(decompressed + start, segment_bytes, packed_element_sizes, start, this_length, locator);
		}
	}
}

} // namespace kernels
} // namespace giddy

#endif // GIDDY_B200_KERNELS_DZB_VARIABLE_CUH_
