// Incidence-bitmaps decoder: K bitmaps, one per distinct value; bit i of bitmap b is set iff
// element i has value b.
//
// A warp owns a tile of 32 bitmap containers (1024 elements).
//   Phase 1 -- lane l reads container l of every bitmap (each a fully-coalesced 128-byte warp
//   load, sixteen in flight per lane) and folds them, bit-sliced, into ceil(log2 K) "planes":
//   plane t is the OR of the containers of all bitmaps whose index has bit t set, so bit j of
//   plane t is bit t of the value index of element 32*l + j.  Bitmaps are taken sixteen at a
//   time, so which of the four low planes a container goes to is known at compile time.
//   Phase 2 -- the lane turns its planes into 32 byte-wide indices, a nibble of every plane at
//   a time: (nibble * 0x00204081) & 0x01010101 spreads four bits over four bytes, and a
//   multiply-add by 2^t drops them into bit t of each byte.  The 32 index bytes go to the warp's
//   1 KiB of shared memory.
//   Phase 3 -- output is written in whole 16-byte vectors, consecutive lanes writing consecutive
//   vectors (512 contiguous bytes per warp store): a lane reads its vector's E index bytes with
//   one shared-memory access and translates them through the dictionary (read-only L1 path; at
//   most 256 entries).
// No data-dependent loop and no per-bit work (the reference pops set bits one at a time into
// byte-wide shared memory, src/kernels/decompression/incidence_bitmaps.cuh:47-60).
#pragma once
#ifndef GIDDY_B200_KERNELS_BITMAPS_CUH_
#define GIDDY_B200_KERNELS_BITMAPS_CUH_

#include "common.cuh"

namespace giddy {
namespace kernels {

using namespace giddy::device;

namespace bitmaps_detail {

constexpr unsigned max_planes = 8;   // bitmap indices are 8-bit (reference incidence_bitmaps.cuh:25)
constexpr unsigned group_size = 16;  // bitmaps folded per step: their low four index bits are compile-time constants
constexpr unsigned max_warps_per_block = GIDDY_SEGMENT_WARPS_PER_BLOCK;
constexpr unsigned tile_containers = GIDDY_BITMAPS_CONTAINERS_PER_TILE;
static_assert(tile_containers == warp_size, "a lane owns one container of the tile");

// Phases 1 and 2 for one tile: returns with the tile's 1024 index bytes in `indices` (element order).
template <unsigned NumPlanes>
__device__ __forceinline__ void tile_indices(
	uint8_t*         __restrict__  indices,        // [1024], 16-byte aligned, this warp's
	const uint32_t*  __restrict__  lane_containers, // this lane's container of bitmap 0
	bool                           lane_has_container,
	uint64_t                       aligned_bitmap_length,
	unsigned                       num_bitmaps)
{
	uint32_t planes[NumPlanes];
	#pragma unroll
	for (unsigned t = 0; t < NumPlanes; t++) { planes[t] = 0; }

	for (unsigned first_bitmap = 0; first_bitmap < num_bitmaps; first_bitmap += group_size) {
		uint32_t containers[group_size];
		#pragma unroll
		for (unsigned k = 0; k < group_size; k++) {
			containers[k] = (lane_has_container && first_bitmap + k < num_bitmaps) ?
				load_streaming_4(lane_containers + (uint64_t)(first_bitmap + k) * aligned_bitmap_length) : 0u;
		}
		uint32_t any = 0;
		#pragma unroll
		for (unsigned k = 0; k < group_size; k++) {
			#pragma unroll
			for (unsigned t = 0; t < 4 && t < NumPlanes; t++) {
				if ((k >> t) & 1u) { planes[t] |= containers[k]; }
			}
			any |= containers[k];
		}
		if (NumPlanes > 4) {
			const unsigned high = first_bitmap / group_size; // warp-uniform
			#pragma unroll
			for (unsigned t = 4; t < NumPlanes; t++) { planes[t] |= ((high >> (t - 4)) & 1u) ? any : 0u; }
		}
	}

	// 32 index bytes, four (one nibble of every plane) at a time
	uint32_t words[8];
	#pragma unroll
	for (unsigned n = 0; n < 8; n++) {
		uint32_t acc = 0;
		#pragma unroll
		for (unsigned t = 0; t < NumPlanes; t++) {
			const uint32_t nibble = (planes[t] >> (4 * n)) & 0xFu;
			acc = ((nibble * 0x00204081u) & 0x01010101u) * (1u << t) + acc;
		}
		words[n] = acc;
	}
	uint4* mine = reinterpret_cast<uint4*>(indices + 32 * lane_index());
	mine[0] = make_uint4(words[0], words[1], words[2], words[3]);
	mine[1] = make_uint4(words[4], words[5], words[6], words[7]);
}

template <unsigned U, bool UseDictionary>
__device__ __forceinline__ uint_t<U> translate(const uint_t<U>* __restrict__ dictionary_entries, unsigned index)
{
	return UseDictionary ? __ldg(dictionary_entries + index) : (uint_t<U>) index;
}

// Phase 3 for one tile.
template <unsigned U, bool UseDictionary, unsigned E>
__device__ __forceinline__ void write_tile(
	uint_t<U>*        __restrict__  tile_out,          // first element of the tile
	const uint_t<U>*  __restrict__  dictionary_entries,
	const uint8_t*    __restrict__  indices,
	unsigned                        num_elements_in_tile)
{
	const unsigned lane = lane_index();
	#pragma unroll 4
	for (unsigned p = lane * E; p < num_elements_in_tile; p += warp_size * E) {
		// this vector's E index bytes, in one shared-memory access
		uint32_t mine[(E + 3) / 4];
		if (E == 16)     { const uint4 t = *reinterpret_cast<const uint4*>(indices + p); mine[0] = t.x; mine[1 % ((E + 3) / 4)] = t.y; mine[2 % ((E + 3) / 4)] = t.z; mine[3 % ((E + 3) / 4)] = t.w; }
		else if (E == 8) { const uint2 t = *reinterpret_cast<const uint2*>(indices + p); mine[0] = t.x; mine[1 % ((E + 3) / 4)] = t.y; }
		else if (E == 4) { mine[0] = *reinterpret_cast<const uint32_t*>(indices + p); }
		else if (E == 2) { mine[0] = *reinterpret_cast<const uint16_t*>(indices + p); }
		else             { mine[0] = indices[p]; }
		if (E > 1) {
			byte_pack<16> vec;
			#pragma unroll
			for (unsigned j = 0; j < E; j++) {
				vec.template set<U>(j, translate<U, UseDictionary>(dictionary_entries, (mine[j / 4] >> (8 * (j % 4))) & 0xFFu));
			}
			store_pack_16(tile_out + p, vec);
		}
		else { tile_out[p] = translate<U, UseDictionary>(dictionary_entries, mine[0] & 0xFFu); }
	}
}

} // namespace bitmaps_detail

// -------------------------------------------------------------------------------------------------
// Replaces cuda::kernels::decompression::incidence_bitmaps::decompress,
// reference src/kernels/decompression/incidence_bitmaps.cuh:69-131 (same parameter list).
//   decompressed[32*w + j] = dictionary_entries[b]  (or (U) b)  where bit j of container w of
//   bitmap b is set; bitmap b starts at container b * round_up(bitmap_length, BitmapAlignmentInInts).
//   Output is produced for whole containers, i.e. 32 * bitmap_length elements.
// Any grid; blocks of up to GIDDY_SEGMENT_WARPS_PER_BLOCK warps.
// -------------------------------------------------------------------------------------------------
template <unsigned IndexSize, unsigned U, unsigned BitmapAlignmentInInts, bool UseDictionary>
__global__ void __launch_bounds__(bitmaps_detail::max_warps_per_block * 32, 5) incidence_bitmaps_decompress(
	uint_t<U>*        __restrict__  decompressed,
	const uint32_t*   __restrict__  bitmaps,
	const uint_t<U>*  __restrict__  dictionary_entries,
	size_type<IndexSize>            bitmap_length,
	uint32_t                        num_bitmaps)
{
	using namespace bitmaps_detail;
	constexpr unsigned E = 16 / U;
	__shared__ __align__(16) uint8_t index_bytes[max_warps_per_block][tile_containers * 32];
	uint8_t* const indices = index_bytes[warp_index_in_block()];

	const unsigned lane = lane_index();
	const uint64_t aligned_bitmap_length =
		((uint64_t) bitmap_length + BitmapAlignmentInInts - 1) / BitmapAlignmentInInts * BitmapAlignmentInInts;
	const uint64_t num_tiles = ((uint64_t) bitmap_length + tile_containers - 1) / tile_containers;
	const bool vectorizable = is_aligned_to(decompressed, 16);

	for (uint64_t tile = global_warp_index(); tile < num_tiles; tile += num_warps_in_grid()) {
		const uint64_t first_container = tile * tile_containers;
		const unsigned num_containers_in_tile = (unsigned) minimum<uint64_t>(tile_containers, bitmap_length - first_container);
		const bool lane_has_container = lane < num_containers_in_tile;
		const uint32_t* lane_containers = bitmaps + first_container + lane;

		if (num_bitmaps <= 4)       { tile_indices<2>(indices, lane_containers, lane_has_container, aligned_bitmap_length, num_bitmaps); }
		else if (num_bitmaps <= 16) { tile_indices<4>(indices, lane_containers, lane_has_container, aligned_bitmap_length, num_bitmaps); }
		else                        { tile_indices<8>(indices, lane_containers, lane_has_container, aligned_bitmap_length, num_bitmaps); }
		__syncwarp();

		uint_t<U>* tile_out = decompressed + first_container * 32;
		if (vectorizable) { write_tile<U, UseDictionary, E>(tile_out, dictionary_entries, indices, num_containers_in_tile * 32); }
		else              { write_tile<U, UseDictionary, 1>(tile_out, dictionary_entries, indices, num_containers_in_tile * 32); }
		__syncwarp();
	}
}

} // namespace kernels
} // namespace giddy

#endif // GIDDY_B200_KERNELS_BITMAPS_CUH_
