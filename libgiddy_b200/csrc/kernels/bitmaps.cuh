// Incidence-bitmaps decoder: K bitmaps, one per distinct value; bit i of bitmap b is set iff
// element i has value b.
//
// A warp owns a tile of 32 bitmap containers (1024 elements).  Phase 1 -- lane l reads container
// l of every bitmap (each a fully-coalesced 128-byte warp load; eight in flight per lane) and
// folds them, bit-sliced, into ceil(log2 K) "planes": plane t is the OR of the containers of
// all bitmaps whose index has bit t set, so bit j of plane t is bit t of the value index of
// element 32*l + j.  That is K*ceil(log2 K)/2 logic ops per 32 elements and no data-dependent
// loop (the reference pops set bits one at a time into byte-wide shared memory,
// src/kernels/decompression/incidence_bitmaps.cuh:47-60).  Phase 2 -- output is written in
// whole 16-byte vectors, consecutive lanes writing consecutive vectors: for each group of 32*E
// elements a lane fetches the planes of the container its E elements live in with one shuffle
// per plane, re-assembles the E indices from the plane bits and translates them through the
// dictionary (read-only L1 path; at most 256 entries).
#pragma once
#ifndef GIDDY_B200_KERNELS_BITMAPS_CUH_
#define GIDDY_B200_KERNELS_BITMAPS_CUH_

#include "common.cuh"

namespace giddy {
namespace kernels {

using namespace giddy::device;

namespace bitmaps_detail {

constexpr unsigned max_planes = 8; // bitmap indices are 8-bit (reference incidence_bitmaps.cuh:25)
constexpr unsigned load_batch = 8;

template <unsigned U, bool UseDictionary, unsigned E>
__device__ __forceinline__ void write_tile(
	uint_t<U>*        __restrict__  tile_out,          // first element of the tile
	const uint_t<U>*  __restrict__  dictionary_entries,
	const uint32_t                  (&planes)[max_planes],
	unsigned                        num_planes,
	unsigned                        num_containers_in_tile)
{
	const unsigned lane = lane_index();
	constexpr unsigned lanes_per_container = warp_size / E;   // lanes sharing one 32-element container
	constexpr unsigned containers_per_step = E;               // 32*E elements per warp store
	for (unsigned step = 0; step < warp_size / containers_per_step; step++) {
		if (step * containers_per_step >= num_containers_in_tile) { break; } // warp-uniform
		const unsigned container = step * containers_per_step + lane / lanes_per_container;
		const unsigned bit_offset = (lane % lanes_per_container) * E;
		unsigned index[E];
		#pragma unroll
		for (unsigned j = 0; j < E; j++) { index[j] = 0; }
		#pragma unroll
		for (unsigned t = 0; t < max_planes; t++) {
			if (t >= num_planes) { break; } // warp-uniform
			const uint32_t field = __shfl_sync(full_mask, planes[t], container) >> bit_offset;
			#pragma unroll
			for (unsigned j = 0; j < E; j++) { index[j] |= ((field >> j) & 1u) << t; }
		}
		if (container >= num_containers_in_tile) { continue; }
		uint_t<U>* out = tile_out + (size_t) container * 32 + bit_offset;
		if (E > 1) {
			byte_pack<16> vec;
			#pragma unroll
			for (unsigned j = 0; j < E; j++) {
				vec.template set<U>(j, UseDictionary ? __ldg(dictionary_entries + index[j]) : (uint_t<U>) index[j]);
			}
			store_pack_16(out, vec);
		}
		else {
			out[0] = UseDictionary ? __ldg(dictionary_entries + index[0]) : (uint_t<U>) index[0];
		}
	}
}

} // namespace bitmaps_detail

// -------------------------------------------------------------------------------------------------
// Replaces cuda::kernels::decompression::incidence_bitmaps::decompress,
// reference src/kernels/decompression/incidence_bitmaps.cuh:69-131 (same parameter list).
//   decompressed[32*w + j] = dictionary_entries[b]  (or (U) b)  where bit j of container w of
//   bitmap b is set; bitmap b starts at container b * round_up(bitmap_length, BitmapAlignmentInInts).
//   Output is produced for whole containers, i.e. 32 * bitmap_length elements.
// -------------------------------------------------------------------------------------------------
template <unsigned IndexSize, unsigned U, unsigned BitmapAlignmentInInts, bool UseDictionary>
__global__ void incidence_bitmaps_decompress(
	uint_t<U>*        __restrict__  decompressed,
	const uint32_t*   __restrict__  bitmaps,
	const uint_t<U>*  __restrict__  dictionary_entries,
	size_type<IndexSize>            bitmap_length,
	uint32_t                        num_bitmaps)
{
	using namespace bitmaps_detail;
	constexpr unsigned E = 16 / U;
	const unsigned lane = lane_index();
	const uint64_t aligned_bitmap_length =
		((uint64_t) bitmap_length + BitmapAlignmentInInts - 1) / BitmapAlignmentInInts * BitmapAlignmentInInts;
	const uint64_t num_tiles = ((uint64_t) bitmap_length + warp_size - 1) / warp_size;
	const unsigned num_planes = (num_bitmaps <= 1) ? 0 : (32 - __clz((int)(num_bitmaps - 1)));
	const bool vectorizable = is_aligned_to(decompressed, 16);

	for (uint64_t tile = global_warp_index(); tile < num_tiles; tile += num_warps_in_grid()) {
		const uint64_t first_container = tile * warp_size;
		const unsigned num_containers_in_tile = (unsigned) minimum<uint64_t>(warp_size, bitmap_length - first_container);
		const bool lane_has_container = lane < num_containers_in_tile;
		const uint32_t* lane_containers = bitmaps + first_container + lane;

		uint32_t planes[max_planes];
		#pragma unroll
		for (unsigned t = 0; t < max_planes; t++) { planes[t] = 0; }

		for (unsigned first_bitmap = 0; first_bitmap < num_bitmaps; first_bitmap += load_batch) {
			uint32_t containers[load_batch];
			#pragma unroll
			for (unsigned k = 0; k < load_batch; k++) {
				const unsigned b = first_bitmap + k;
				containers[k] = (lane_has_container && b < num_bitmaps) ?
					load_streaming_4(lane_containers + (uint64_t) b * aligned_bitmap_length) : 0u;
			}
			#pragma unroll
			for (unsigned k = 0; k < load_batch; k++) {
				const unsigned b = first_bitmap + k; // warp-uniform
				#pragma unroll
				for (unsigned t = 0; t < max_planes; t++) {
					if ((b >> t) & 1u) { planes[t] |= containers[k]; }
				}
			}
		}
		uint_t<U>* tile_out = decompressed + first_container * 32;
		if (vectorizable) { write_tile<U, UseDictionary, E>(tile_out, dictionary_entries, planes, num_planes, num_containers_in_tile); }
		else              { write_tile<U, UseDictionary, 1>(tile_out, dictionary_entries, planes, num_planes, num_containers_in_tile); }
	}
}

} // namespace kernels
} // namespace giddy

#endif // GIDDY_B200_KERNELS_BITMAPS_CUH_
