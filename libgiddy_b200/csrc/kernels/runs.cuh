// Run-based decoders: Run-Length Encoding and Run-Position Encoding.
//
// One warp decodes one anchored segment, in tiles of tile_positions output positions.  For
// each tile the warp
//   1. *stages* the runs that begin inside the tile: 32 runs per step, one per lane, with
//      coalesced loads of the run lengths / start positions and run values; a warp-shuffle scan
//      turns lengths into start positions (RLE; RPE has them stored).  Each staged run sets the
//      bit of its start position in a 1024-bit tile bitmap and parks its value, in run order,
//      in a shared-memory array.  The run covering the tile's first position is carried in
//      registers from the previous tile (or comes from the segment's anchor) and is staged as
//      entry 0 with bit 0.
//   2. *expands*: the run covering position p is entry rank(p) - 1, where rank(p) is the number
//      of set bits at positions <= p: a per-word prefix popcount (one warp scan per tile) plus
//      one POPC per element.  Each lane assembles whole 16-byte output vectors, consecutive
//      lanes write consecutive vectors (512 contiguous bytes per warp store), and the value
//      lookups of neighbouring lanes hit the same or adjacent shared-memory words.
// There is no data-dependent inner search loop (the reference walks the run lengths once per
// thread, src/kernels/decompression/run_length_encoding.cuh:150-157), no divergence besides
// tile edges, and the compressed input is read exactly once.
//
// Format restrictions (same as the reference's): run lengths are >= 1; RPE start positions are
// strictly increasing.
#pragma once
#ifndef GIDDY_B200_KERNELS_RUNS_CUH_
#define GIDDY_B200_KERNELS_RUNS_CUH_

#include "common.cuh"

namespace giddy {
namespace kernels {

using namespace giddy::device;

namespace runs {

constexpr unsigned tile_positions = GIDDY_RUNS_TILE_POSITIONS; // 1024: one bitmap word per lane
constexpr unsigned tile_words = tile_positions / 32;
constexpr unsigned staged_capacity = GIDDY_RUNS_STAGED_CAPACITY;

template <unsigned U>
constexpr unsigned shared_memory_bytes_per_warp()
{
	return GIDDY_RUNS_SHARED_BYTES_PER_WARP(U);
}

template <unsigned U>
struct tile_storage {
	uint32_t*  bitmap;   // [tile_words]   bit p set <=> a staged run starts at tile position p
	uint32_t*  prefix;   // [tile_words]   number of set bits in the words before this one
	uint_t<U>* values;   // [staged_capacity] staged run values, in run order

	__device__ __forceinline__ explicit tile_storage(unsigned char* warp_memory) :
		bitmap(reinterpret_cast<uint32_t*>(warp_memory)),
		prefix(reinterpret_cast<uint32_t*>(warp_memory) + tile_words),
		values(reinterpret_cast<uint_t<U>*>(warp_memory + 2 * tile_words * sizeof(uint32_t))) { }
};

// Step 2 above, for one tile.  `out` points at the tile's first output element.  E elements per
// store; E > 1 requires `out` to be 16-byte aligned.
template <unsigned U, unsigned E>
__device__ __forceinline__ void expand_tile(uint_t<U>* __restrict__ out, const tile_storage<U>& tile, unsigned tile_length)
{
	const unsigned lane = lane_index();
	{
		uint32_t word = tile.bitmap[lane];
		uint32_t count = __popc(word);
		uint32_t inclusive = warp_inclusive_sum(count);
		tile.prefix[lane] = inclusive - count;
	}
	__syncwarp();
	const unsigned num_vectors = tile_length / E;
	for (unsigned v = lane; v < num_vectors; v += warp_size) {
		const unsigned p0 = v * E;
		const uint32_t word = tile.bitmap[p0 >> 5];
		const uint32_t before = tile.prefix[p0 >> 5];
		if (E == 1) {
			unsigned rank = before + __popc(word & (0xffffffffu >> (31 - (p0 & 31))));
			out[p0] = tile.values[rank - 1];
		}
		else {
			byte_pack<16> vec;
			// all E positions share the bitmap word since E divides 32 and p0 is a multiple of E
			const unsigned first_rank = before + __popc(word & (0xffffffffu >> (31 - (p0 & 31))));
			const unsigned last_rank  = before + __popc(word & (0xffffffffu >> (31 - ((p0 + E - 1) & 31))));
			uint_t<U> value = tile.values[first_rank - 1];
			if (first_rank == last_rank) {
				#pragma unroll
				for (unsigned e = 0; e < E; e++) { vec.template set<U>(e, value); }
			}
			else {
				vec.template set<U>(0, value);
				#pragma unroll
				for (unsigned e = 1; e < E; e++) {
					unsigned rank = before + __popc(word & (0xffffffffu >> (31 - ((p0 + e) & 31))));
					vec.template set<U>(e, tile.values[rank - 1]);
				}
			}
			store_pack_16(out + p0, vec);
		}
	}
	if (E > 1) { // the < E trailing elements of the column's last tile
		const unsigned p = num_vectors * E + lane;
		if (p < tile_length) {
			unsigned rank = tile.prefix[p >> 5] + __popc(tile.bitmap[p >> 5] & (0xffffffffu >> (31 - (p & 31))));
			out[p] = tile.values[rank - 1];
		}
	}
	__syncwarp();
}

} // namespace runs

// -------------------------------------------------------------------------------------------------
// Run-Length Encoding.  Replaces cuda::kernels::decompression::run_length_encoding::decompress,
// reference src/kernels/decompression/run_length_encoding.cuh:342-374 (same parameter list).
//   decompressed[i] = run_data[r(i)];  segment s begins inside run position_anchors[s],
//   intra_run_offsets[s] elements into it (:284-325).  The last run extends to the end of the
//   column whatever its stored length (:307-313).
// Dynamic shared memory: runs::shared_memory_bytes_per_warp<U>() per warp.
// -------------------------------------------------------------------------------------------------
template <unsigned UncompressedIndexSize, unsigned U, unsigned RunLengthSize, unsigned RunIndexSize>
__global__ void run_length_encoding_decompress(
	uint_t<U>*                            __restrict__  decompressed,
	const uint_t<U>*                      __restrict__  run_data,
	const uint_t<RunLengthSize>*          __restrict__  run_lengths,
	const uint_t<UncompressedIndexSize>*  __restrict__  segment_position_anchors,
	const uint_t<RunLengthSize>*          __restrict__  segment_position_anchor_intra_run_offsets,
	size_type<UncompressedIndexSize>                    segment_length,
	size_type<UncompressedIndexSize>                    num_segments,
	size_type<RunIndexSize>                             num_element_runs,
	size_type<UncompressedIndexSize>                    total_uncompressed_length)
{
	using position_t = size_type<UncompressedIndexSize>; // positions relative to the segment's start
	constexpr unsigned E = 16 / U;
	constexpr position_t unbounded = ~(position_t) 0;
	using namespace runs;

	const unsigned lane = lane_index();
	tile_storage<U> tile(warp_shared_memory(shared_memory_bytes_per_warp<U>()));
	const bool vectorizable = is_aligned_to(decompressed, 16) && (segment_length % E == 0);
	const uint64_t num_runs = num_element_runs;

	for (uint64_t segment = global_warp_index(); segment < num_segments; segment += num_warps_in_grid()) {
		const uint64_t segment_start = segment * segment_length;
		if (segment_start >= total_uncompressed_length) { break; }
		const position_t this_segment_length = (position_t) minimum<uint64_t>(segment_length, total_uncompressed_length - segment_start);
		uint_t<U>* const segment_out = decompressed + segment_start;

		// the run covering the next tile's first position, its value, and where it ends
		uint64_t covering_run = __ldg(segment_position_anchors + segment);
		uint_t<U> covering_value = __ldg(run_data + covering_run);
		position_t covering_end = (covering_run + 1 >= num_runs) ? unbounded :
			(position_t)(__ldg(run_lengths + covering_run) - __ldg(segment_position_anchor_intra_run_offsets + segment));

		for (position_t tile_start = 0; tile_start < this_segment_length; tile_start += tile_positions) {
			const unsigned tile_length = (unsigned) minimum<position_t>(tile_positions, this_segment_length - tile_start);
			const position_t tile_end = tile_start + tile_length;

			tile.bitmap[lane] = (lane == 0) ? 1u : 0u;
			if (lane == 0) { tile.values[0] = covering_value; }
			__syncwarp();

			unsigned num_staged = 1;
			uint64_t next_run = covering_run + 1;
			position_t next_start = covering_end; // where run `next_run` begins
			while (next_start <= tile_end && next_run < num_runs) {
				const uint64_t run = next_run + lane;
				const bool valid = run < num_runs;
				const bool is_last_run = (run + 1 == num_runs);
				const position_t true_length = valid ? (position_t) __ldg(run_lengths + run) : 0;
				const uint_t<U> value = valid ? __ldg(run_data + run) : 0;
				// no run needs to reach further than just past the tile, as far as this tile goes
				const position_t length_in_scan = (is_last_run) ? (position_t)(tile_length + 1) :
					minimum<position_t>(true_length, (position_t)(tile_length + 1));
				const position_t inclusive = warp_inclusive_sum(length_in_scan);
				const position_t start = next_start + (inclusive - length_in_scan);
				const bool consumed = valid && start <= tile_end; // begins in this tile, or exactly at its end
				const bool staged = valid && start < tile_end;
				const unsigned consumed_mask = __ballot_sync(full_mask, consumed);
				const unsigned num_consumed = __popc(consumed_mask);
				if (staged) {
					const unsigned bit = (unsigned)(start - tile_start);
					atomicOr(tile.bitmap + (bit >> 5), 1u << (bit & 31));
					tile.values[num_staged + lane] = value;
				}
				if (num_consumed > 0) {
					uint64_t true_end = (uint64_t) start + true_length;
					position_t saturated_end = (is_last_run || true_end > (uint64_t) unbounded) ? unbounded : (position_t) true_end;
					covering_run = next_run + num_consumed - 1;
					covering_end = shuffle_from(saturated_end, num_consumed - 1);
					covering_value = shuffle_from(value, num_consumed - 1);
				}
				num_staged += __popc(__ballot_sync(full_mask, staged));
				if (num_consumed < warp_size) { break; }
				next_run += warp_size;
				next_start = shuffle_from((position_t)(next_start + inclusive), warp_size - 1);
			}
			__syncwarp();
			if (vectorizable) { expand_tile<U, E>(segment_out + tile_start, tile, tile_length); }
			else              { expand_tile<U, 1>(segment_out + tile_start, tile, tile_length); }
		}
	}
}

// -------------------------------------------------------------------------------------------------
// Run-Position Encoding.  Replaces cuda::kernels::decompression::run_position_encoding::decompress,
// reference src/kernels/decompression/run_position_encoding.cuh:327-359 (same parameter list).
//   decompressed[i] = run_data[ max r : run_start_positions[r] <= i ];  segment s begins inside
//   run position_anchors[s].  Absolute start positions index the whole column; relative ones are
//   relative to the start of the segment in which the run begins (see oracle/giddy_oracle.cpp on
//   the reference's boundary case).
// Dynamic shared memory: runs::shared_memory_bytes_per_warp<U>() per warp.
// -------------------------------------------------------------------------------------------------
template <unsigned UncompressedIndexSize, unsigned U, unsigned PositionOffsetSize, bool PositionsAreRelative, unsigned RunIndexSize>
__global__ void run_position_encoding_decompress(
	uint_t<U>*                         __restrict__  decompressed,
	const uint_t<U>*                   __restrict__  run_data,
	const uint_t<PositionOffsetSize>*  __restrict__  run_start_positions,
	const uint_t<RunIndexSize>*        __restrict__  position_anchors,
	size_type<PositionOffsetSize>                    anchoring_period,
	size_type<RunIndexSize>                          num_anchors,
	size_type<RunIndexSize>                          num_element_runs,
	size_type<UncompressedIndexSize>                 uncompressed_length)
{
	constexpr unsigned E = 16 / U;
	using namespace runs;

	const unsigned lane = lane_index();
	tile_storage<U> tile(warp_shared_memory(shared_memory_bytes_per_warp<U>()));
	const uint64_t segment_length = anchoring_period;
	const bool vectorizable = is_aligned_to(decompressed, 16) && (segment_length % E == 0);
	const uint64_t num_runs = num_element_runs;

	for (uint64_t segment = global_warp_index(); segment < num_anchors; segment += num_warps_in_grid()) {
		const uint64_t segment_start = segment * segment_length;
		if (segment_start >= uncompressed_length) { break; }
		const uint64_t this_segment_length = minimum<uint64_t>(segment_length, uncompressed_length - segment_start);
		uint_t<U>* const segment_out = decompressed + segment_start;

		uint64_t covering_run = __ldg(position_anchors + segment);
		uint_t<U> covering_value = __ldg(run_data + covering_run);

		// one past the last run which begins in this segment
		uint64_t run_limit = num_runs;
		if (PositionsAreRelative && segment + 1 < num_anchors) {
			uint64_t run_at_next_anchor = __ldg(position_anchors + segment + 1);
			bool begins_before_next_segment = (run_at_next_anchor == covering_run) ||
				(__ldg(run_start_positions + run_at_next_anchor) != 0);
			run_limit = run_at_next_anchor + (begins_before_next_segment ? 1 : 0);
		}
		// subtracted from stored start positions to make them relative to the segment
		const uint64_t position_base = PositionsAreRelative ? 0 : segment_start;

		for (uint64_t tile_start = 0; tile_start < this_segment_length; tile_start += tile_positions) {
			const unsigned tile_length = (unsigned) minimum<uint64_t>(tile_positions, this_segment_length - tile_start);
			const uint64_t tile_end = tile_start + tile_length;

			tile.bitmap[lane] = (lane == 0) ? 1u : 0u;
			if (lane == 0) { tile.values[0] = covering_value; }
			__syncwarp();

			unsigned num_staged = 1;
			uint64_t next_run = covering_run + 1;
			while (next_run < run_limit) {
				const uint64_t run = next_run + lane;
				const bool valid = run < run_limit;
				const uint64_t start = valid ? (uint64_t) __ldg(run_start_positions + run) - position_base : ~(uint64_t) 0;
				const uint_t<U> value = valid ? __ldg(run_data + run) : 0;
				const bool consumed = valid && start <= tile_end;
				const bool staged = valid && start < tile_end;
				const unsigned num_consumed = __popc(__ballot_sync(full_mask, consumed));
				if (staged) {
					const unsigned bit = (unsigned)(start - tile_start);
					atomicOr(tile.bitmap + (bit >> 5), 1u << (bit & 31));
					tile.values[num_staged + lane] = value;
				}
				if (num_consumed > 0) {
					covering_run = next_run + num_consumed - 1;
					covering_value = shuffle_from(value, num_consumed - 1);
				}
				num_staged += __popc(__ballot_sync(full_mask, staged));
				if (num_consumed < warp_size) { break; }
				next_run += warp_size;
			}
			__syncwarp();
			if (vectorizable) { expand_tile<U, E>(segment_out + tile_start, tile, tile_length); }
			else              { expand_tile<U, 1>(segment_out + tile_start, tile, tile_length); }
		}
	}
}

} // namespace kernels
} // namespace giddy

#endif // GIDDY_B200_KERNELS_RUNS_CUH_
