// Run-based decoders: Run-Length Encoding and Run-Position Encoding.
//
// One warp decodes one anchored segment, in tiles of tile_positions output positions.  For
// each tile the warp
//   1. *stages* the runs which intersect the tile: 32 runs per step, one per lane, with coalesced
//      loads of the run lengths / start positions and run values; a warp-shuffle scan turns
//      lengths into start positions (RLE; RPE has them stored).  Each staged run sets the bit of
//      its (clamped) start position in a 1024-bit tile bitmap and parks its value, in run order,
//      in a shared-memory array.  The run covering the tile's first position is lane 0 of the
//      first step: in a segment's first tile it is the anchor's run, shortened by the anchor's
//      intra-run offset; in later tiles it is carried over in registers.
//   2. *expands*: the run covering position p is entry rank(p) - 1, where rank(p) is the number
//      of set bits at positions <= p: a per-word prefix popcount (one warp scan per tile), one
//      POPC per output vector and one bit test per further element.  Each lane assembles whole
//      16-byte output vectors, consecutive lanes write consecutive vectors (512 contiguous bytes
//      per warp store), and the value lookups of neighbouring lanes hit the same or adjacent
//      shared-memory words.  For a full tile the loop is unrolled with every shared-memory and
//      global offset an immediate.
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
	// [tile_words] .x: bit p set <=> a staged run starts at tile position 32 * word + p
	//              .y: number of set bits in the words before this one
	uint2*     index;
	uint_t<U>* values;   // [staged_capacity] staged run values, in run order

	__device__ __forceinline__ explicit tile_storage(unsigned char* warp_memory) :
		index(reinterpret_cast<uint2*>(warp_memory)),
		values(reinterpret_cast<uint_t<U>*>(warp_memory + tile_words * sizeof(uint2))) { }

	__device__ __forceinline__ void clear() const { index[lane_index()].x = 0u; }

	__device__ __forceinline__ void stage(unsigned slot, unsigned tile_position, uint_t<U> value) const
	{
		atomicOr(&index[tile_position >> 5].x, 1u << (tile_position & 31));
		values[slot] = value;
	}
};

// Step 2 above, for one tile.  `out` points at the tile's first output element.  E elements per
// store; E > 1 requires `out` to be 16-byte aligned.
template <unsigned U, unsigned E>
__device__ __forceinline__ void expand_tile(uint_t<U>* __restrict__ out, const tile_storage<U>& tile, unsigned tile_length)
{
	const unsigned lane = lane_index();
	{
		const uint32_t count = __popc(tile.index[lane].x);
		const uint32_t inclusive = warp_inclusive_sum(count);
		tile.index[lane].y = inclusive - count;
	}
	__syncwarp();
	const uint_t<U>* const values_by_rank = tile.values - 1; // rank r (>= 1) -> entry r - 1

	if (E > 1 && tile_length == tile_positions) {
		// Full tile: iteration `it` covers positions [it * 32 * E, (it + 1) * 32 * E); this lane's vector starts
		// at it * 32 * E + lane * E, i.e. in word it * E + lane * E / 32 at a bit offset which does not depend on `it`.
		const unsigned bit = (lane * E) & 31;
		const uint32_t up_to_bit = 0xffffffffu >> (31 - bit);
		const uint2* lane_index_entries = tile.index + ((lane * E) >> 5);
		uint_t<U>* lane_out = out + lane * E;
		#pragma unroll
		for (unsigned it = 0; it < tile_positions / (warp_size * E); it++) {
			const uint2 entry = lane_index_entries[it * E];
			unsigned rank = entry.y + __popc(entry.x & up_to_bit);
			const uint32_t following = entry.x >> bit; // bit e: does a run start at this vector's element e?
			byte_pack<16> vec;
			vec.template set<U>(0, values_by_rank[rank]);
			#pragma unroll
			for (unsigned e = 1; e < E; e++) {
				rank += (following >> e) & 1u;
				vec.template set<U>(e, values_by_rank[rank]);
			}
			store_pack_16(lane_out + it * warp_size * E, vec);
		}
		__syncwarp();
		return;
	}

	const unsigned num_vectors = tile_length / E;
	for (unsigned v = lane; v < num_vectors; v += warp_size) {
		const unsigned p0 = v * E;
		const uint2 entry = tile.index[p0 >> 5];
		unsigned rank = entry.y + __popc(entry.x & (0xffffffffu >> (31 - (p0 & 31))));
		if (E == 1) { out[p0] = values_by_rank[rank]; }
		else {
			const uint32_t following = entry.x >> (p0 & 31);
			byte_pack<16> vec;
			vec.template set<U>(0, values_by_rank[rank]);
			#pragma unroll
			for (unsigned e = 1; e < E; e++) {
				rank += (following >> e) & 1u;
				vec.template set<U>(e, values_by_rank[rank]);
			}
			store_pack_16(out + p0, vec);
		}
	}
	if (E > 1) { // the < E trailing elements of the column's last tile
		const unsigned p = num_vectors * E + lane;
		if (p < tile_length) {
			const uint2 entry = tile.index[p >> 5];
			out[p] = values_by_rank[entry.y + __popc(entry.x & (0xffffffffu >> (31 - (p & 31))))];
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
	using run_t = size_type<RunIndexSize>;
	constexpr unsigned E = 16 / U;
	constexpr position_t unbounded = ~(position_t) 0;
	using namespace runs;

	const unsigned lane = lane_index();
	const tile_storage<U> tile(warp_shared_memory(shared_memory_bytes_per_warp<U>()));
	const bool vectorizable = is_aligned_to(decompressed, 16) && (segment_length % E == 0);
	const run_t num_runs = num_element_runs;

	for (uint64_t segment = global_warp_index(); segment < num_segments; segment += num_warps_in_grid()) {
		const uint64_t segment_start = segment * segment_length;
		if (segment_start >= total_uncompressed_length) { break; }
		const position_t this_segment_length = (position_t) minimum<uint64_t>(segment_length, total_uncompressed_length - segment_start);
		uint_t<U>* const segment_out = decompressed + segment_start;

		// The run covering the next tile's first position; where it ends and its value are known from the second tile on.
		run_t covering_run = __ldg(segment_position_anchors + segment);
		const position_t anchor_offset = __ldg(segment_position_anchor_intra_run_offsets + segment);
		position_t covering_end = 0;
		uint_t<U> covering_value = 0;

		for (position_t tile_start = 0; tile_start < this_segment_length; tile_start += tile_positions) {
			const unsigned tile_length = (unsigned) minimum<position_t>(tile_positions, this_segment_length - tile_start);
			const position_t tile_end = tile_start + tile_length;
			tile.clear();
			__syncwarp();

			unsigned num_staged = 0;
			run_t chunk_first_run = covering_run; // lane 0 of the first chunk is the covering run
			position_t chunk_start = tile_start;  // where the chunk's first run begins (the covering run: clamped to the tile)
			bool first_chunk = true;
			do {
				const bool valid = lane < num_runs - chunk_first_run;
				const run_t run = chunk_first_run + lane;
				const bool is_last_run = valid && (run + 1 == num_runs);
				position_t true_length = valid ? (position_t) __ldg(run_lengths + run) : 0;
				uint_t<U> value = valid ? __ldg(run_data + run) : 0;
				if (first_chunk && lane == 0) {
					if (tile_start == 0) { true_length -= anchor_offset; }
					else { true_length = covering_end - tile_start; value = covering_value; }
				}
				if (is_last_run) { true_length = unbounded; }
				// as far as this tile goes, no run needs to reach further than just past it
				const position_t length_in_scan = minimum<position_t>(true_length, (position_t)(tile_length + 1));
				const position_t inclusive = warp_inclusive_sum(length_in_scan);
				const position_t start = chunk_start + (inclusive - length_in_scan);
				const bool consumed = valid && start <= tile_end; // begins in this tile, or exactly at its end
				const bool staged = valid && start < tile_end;
				const unsigned num_consumed = __popc(__ballot_sync(full_mask, consumed));
				if (staged) { tile.stage(num_staged + lane, (unsigned)(start - tile_start), value); }
				// (num_consumed >= 1: the first chunk's lane 0 begins at tile_start; later chunks are entered only if their first run
				// begins in the tile -- the test only guards against malformed input)
				if (num_consumed > 0) {
					const position_t room = unbounded - start;
					const position_t saturated_end = (true_length >= room) ? unbounded : (position_t)(start + true_length);
					covering_run = chunk_first_run + (num_consumed - 1);
					covering_end = shuffle_from(saturated_end, num_consumed - 1);
					covering_value = shuffle_from(value, num_consumed - 1);
				}
				num_staged += __popc(__ballot_sync(full_mask, staged));
				if (num_consumed < warp_size) { break; }
				chunk_first_run += warp_size;
				chunk_start = shuffle_from((position_t)(chunk_start + inclusive), warp_size - 1);
				first_chunk = false;
			} while (chunk_start <= tile_end && chunk_first_run < num_runs);
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
	// positions relative to the segment's start
	using position_t = typename std::conditional<(UncompressedIndexSize > 4 || PositionOffsetSize > 4), uint64_t, uint32_t>::type;
	using run_t = size_type<RunIndexSize>;
	constexpr unsigned E = 16 / U;
	using namespace runs;

	const unsigned lane = lane_index();
	const tile_storage<U> tile(warp_shared_memory(shared_memory_bytes_per_warp<U>()));
	const uint64_t segment_length = anchoring_period;
	const bool vectorizable = is_aligned_to(decompressed, 16) && (segment_length % E == 0);
	const run_t num_runs = num_element_runs;

	for (uint64_t segment = global_warp_index(); segment < num_anchors; segment += num_warps_in_grid()) {
		const uint64_t segment_start = segment * segment_length;
		if (segment_start >= uncompressed_length) { break; }
		const position_t this_segment_length = (position_t) minimum<uint64_t>(segment_length, uncompressed_length - segment_start);
		uint_t<U>* const segment_out = decompressed + segment_start;

		run_t covering_run = __ldg(position_anchors + segment);
		uint_t<U> covering_value = 0;

		// one past the last run which begins in this segment
		run_t run_limit = num_runs;
		if (PositionsAreRelative && segment + 1 < num_anchors) {
			const run_t run_at_next_anchor = __ldg(position_anchors + segment + 1);
			const bool begins_before_next_segment = (run_at_next_anchor == covering_run) ||
				(__ldg(run_start_positions + run_at_next_anchor) != 0);
			run_limit = run_at_next_anchor + (begins_before_next_segment ? 1 : 0);
		}
		// subtracted from stored start positions to make them relative to the segment
		const uint_t<PositionOffsetSize> position_base = PositionsAreRelative ? 0 : (uint_t<PositionOffsetSize>) segment_start;

		for (position_t tile_start = 0; tile_start < this_segment_length; tile_start += tile_positions) {
			const unsigned tile_length = (unsigned) minimum<position_t>(tile_positions, this_segment_length - tile_start);
			const position_t tile_end = tile_start + tile_length;
			tile.clear();
			__syncwarp();

			unsigned num_staged = 0;
			run_t chunk_first_run = covering_run; // lane 0 of the first chunk is the covering run
			bool first_chunk = true;
			do {
				const bool valid = lane < run_limit - chunk_first_run;
				const run_t run = chunk_first_run + lane;
				position_t start = valid ? (position_t)(uint_t<PositionOffsetSize>)(__ldg(run_start_positions + run) - position_base) : 0;
				uint_t<U> value = valid ? __ldg(run_data + run) : 0;
				if (first_chunk && lane == 0) { // the covering run begins at or before the tile: its stored start is not needed
					start = tile_start;
					if (tile_start != 0) { value = covering_value; }
				}
				const bool consumed = valid && start <= tile_end;
				const bool staged = valid && start < tile_end;
				const unsigned num_consumed = __popc(__ballot_sync(full_mask, consumed));
				if (staged) { tile.stage(num_staged + lane, (unsigned)(start - tile_start), value); }
				if (num_consumed > 0) { // (a later chunk's first run may begin beyond the tile)
					covering_run = chunk_first_run + (num_consumed - 1);
					covering_value = shuffle_from(value, num_consumed - 1);
				}
				num_staged += __popc(__ballot_sync(full_mask, staged));
				if (num_consumed < warp_size) { break; }
				chunk_first_run += warp_size;
				first_chunk = false;
			} while (chunk_first_run < run_limit);
			__syncwarp();
			if (vectorizable) { expand_tile<U, E>(segment_out + tile_start, tile, tile_length); }
			else              { expand_tile<U, 1>(segment_out + tile_start, tile, tile_length); }
		}
	}
}

} // namespace kernels
} // namespace giddy

#endif // GIDDY_B200_KERNELS_RUNS_CUH_
