// Patching kernels: plain scatter and compressed-indices scatter.
//
// Both read their inputs with warp-contiguous accesses and issue the (inherently scattered)
// element stores with several independent ones in flight per thread.  Patch targets were just
// written by the decode kernel, so they are usually still resident in the 126 MB L2 and the
// read-modify-write of a 32-byte sector happens there rather than in HBM.
#pragma once
#ifndef GIDDY_B200_KERNELS_SCATTER_CUH_
#define GIDDY_B200_KERNELS_SCATTER_CUH_

#include "common.cuh"

namespace giddy {
namespace kernels {

using namespace giddy::device;

// -------------------------------------------------------------------------------------------------
// Replaces cuda::kernels::scatter::scatter, reference src/kernels/data_layout/scatter.cuh:26-46
// (same parameter list):  target[indices[i]] = data[i]  for i < data_length.
// -------------------------------------------------------------------------------------------------
template <unsigned OutputIndexSize, unsigned ElementSize, unsigned InputIndexSize>
__global__ void scatter(
	uint_t<ElementSize>*            __restrict__  target,
	const uint_t<ElementSize>*      __restrict__  data,
	const uint_t<OutputIndexSize>*  __restrict__  indices,
	size_type<InputIndexSize>                     data_length)
{
	constexpr unsigned unroll = GIDDY_SCATTER_UNROLL;
	const uint64_t per_tile = (uint64_t) blockDim.x * unroll;
	for (uint64_t tile_start = (uint64_t) blockIdx.x * per_tile; tile_start < data_length; tile_start += (uint64_t) gridDim.x * per_tile) {
		uint_t<ElementSize> values[unroll];
		uint_t<OutputIndexSize> positions[unroll];
		#pragma unroll
		for (unsigned k = 0; k < unroll; k++) {
			const uint64_t i = tile_start + (uint64_t) k * blockDim.x + threadIdx.x;
			if (i < data_length) { values[k] = __ldg(data + i); positions[k] = __ldg(indices + i); }
		}
		#pragma unroll
		for (unsigned k = 0; k < unroll; k++) {
			const uint64_t i = tile_start + (uint64_t) k * blockDim.x + threadIdx.x;
			if (i < data_length) { target[positions[k]] = values[k]; }
		}
	}
}

namespace ci_detail {

// An offset of `width_bytes` (1, 2, 4 or 8) bytes, the j-th of a run whose offsets start at `run_offsets`
// (aligned to min(width, OutputIndexSize) bytes).
__device__ __forceinline__ uint64_t load_offset(const uint8_t* __restrict__ run_offsets, uint64_t j, unsigned width_bytes)
{
	switch (width_bytes) {
	case 1: return __ldg(run_offsets + j);
	case 2: return __ldg(reinterpret_cast<const uint16_t*>(run_offsets) + j);
	case 4: return __ldg(reinterpret_cast<const uint32_t*>(run_offsets) + j);
	default: {
		const uint32_t* words = reinterpret_cast<const uint32_t*>(run_offsets) + 2 * j;
		return (uint64_t) __ldg(words) | ((uint64_t) __ldg(words + 1) << 32);
	}
	}
}

} // namespace ci_detail

// -------------------------------------------------------------------------------------------------
// Replaces cuda::kernels::scatter::compressed_indices::scatter,
// reference src/kernels/data_layout/compressed_indices_scatter.cuh:219-271 (same parameter list).
// Patch positions come in runs; run r has a baseline position, a per-run offset width (8/16/32/64
// bits) and run_lengths[r] offsets stored at byte offsets_start[r] * OutputIndexSize of the offset
// bytes; anchoring period a covers the runs [anchors[a], anchors[a+1]) and patch values from a*A.
//   target[(OI)(baseline[r] + offset_r[j])] = data[k],  k advancing through the runs in order.
//
// A warp owns an anchoring period.  The metadata of 32 runs is loaded at once (one run per lane) and
// parked in the warp's slice of shared memory together with a warp scan of the run lengths; the patches
// of those runs are then treated as one flat list: lane-strided, each patch finds its run with a 5-step
// branch-free search of the scanned lengths, and ci_unroll patches per lane have their offset and value
// loads in flight at once.  So a period costs one metadata round trip per 32 runs plus one per 32 x
// ci_unroll patches, however short its runs are (the reference handles the runs one after the other,
// with a dependent metadata read and a block-stride loop each, :83-124).
// -------------------------------------------------------------------------------------------------
namespace ci_detail {

constexpr unsigned ci_unroll = 4;        // patches per lane per batch when scattering
constexpr unsigned ci_list_unroll = 8;   // ... when only writing the decoded positions out (no values to hold)

struct staged_runs_t {
	uint64_t ends[warp_size];           // inclusive scan of the staged runs' lengths
	uint64_t baselines[warp_size];
	uint64_t offset_starts[warp_size];  // byte offset of the run's first offset
	uint8_t  widths[warp_size];         // bytes per offset
};

} // namespace ci_detail

namespace ci_detail {

// Walk every patch of the run-compressed position list, anchoring period by anchoring period (a warp each): for patch
// number k (its index into the patch values) at decoded position p, `sink.prepare(u, k)` is called first for each of a
// lane's up to ci_unroll patches of a batch (request what the patch needs), then `sink.commit(u, k, p)` for each.
template <unsigned OutputIndexSize, unsigned InputIndexSize, unsigned RunLengthSize, unsigned Unroll, typename Sink>
__device__ __forceinline__ void for_each_patch(
	Sink&                                         sink,
	staged_runs_t&                                staged,
	const uint_t<RunLengthSize>*    __restrict__  scatter_position_run_lengths,
	const unsigned char*            __restrict__  scatter_position_run_individual_offset_sizes,
	const uint_t<OutputIndexSize>*  __restrict__  scatter_position_run_baseline_values,
	const uint_t<InputIndexSize>*   __restrict__  scatter_position_run_offsets_start_positions,
	const unsigned char*            __restrict__  scatter_position_offset_bytes,
	const uint_t<InputIndexSize>*   __restrict__  scatter_position_anchors,
	uint64_t                                      anchoring_period,
	uint64_t                                      num_scatter_position_runs,
	uint64_t                                      data_length)
{
	const unsigned lane = lane_index();
	const uint64_t period_length = anchoring_period;
	const uint64_t num_periods = (data_length + period_length - 1) / period_length;

	for (uint64_t period = global_warp_index(); period < num_periods; period += num_warps_in_grid()) {
		const uint64_t first_run = __ldg(scatter_position_anchors + period);
		const bool reaches_end = (period + 1) * period_length >= data_length;
		const uint64_t past_last_run = reaches_end ? num_scatter_position_runs : (uint64_t) __ldg(scatter_position_anchors + period + 1);
		uint64_t next_value = period * period_length; // index into the patch values of the next run's first patch

		for (uint64_t chunk_first_run = first_run; chunk_first_run < past_last_run && next_value < data_length; chunk_first_run += warp_size) {
			const uint64_t run = chunk_first_run + lane;
			const bool valid = run < past_last_run;
			const uint64_t my_length = valid ? (uint64_t) __ldg(scatter_position_run_lengths + run) : 0;
			__syncwarp(); // every lane is done with the previous chunk
			staged.widths[lane]        = valid ? (uint8_t)(__ldg(scatter_position_run_individual_offset_sizes + run) / 8) : (uint8_t) 1;
			staged.baselines[lane]     = valid ? (uint64_t) __ldg(scatter_position_run_baseline_values + run) : 0;
			staged.offset_starts[lane] = valid ? (uint64_t) __ldg(scatter_position_run_offsets_start_positions + run) * OutputIndexSize : 0;
			const uint64_t inclusive = warp_inclusive_sum(my_length);
			staged.ends[lane] = inclusive; // (lanes without a run repeat the total: the search never lands on them)
			__syncwarp();
			const uint64_t chunk_patches = shuffle_from(inclusive, warp_size - 1);
			const uint64_t limit = minimum<uint64_t>(chunk_patches, data_length - next_value);

			for (uint64_t q0 = 0; q0 < limit; q0 += warp_size * Unroll) {
				uint_t<OutputIndexSize> positions[Unroll];
				#pragma unroll
				for (unsigned u = 0; u < Unroll; u++) {
					const uint64_t q = q0 + u * warp_size + lane;
					if (q < limit) {
						unsigned r = 0; // the first staged run whose end exceeds q
						#pragma unroll
						for (unsigned step = warp_size / 2; step > 0; step >>= 1) { if (staged.ends[r + step - 1] <= q) { r += step; } }
						const uint64_t j = q - (r == 0 ? 0 : staged.ends[r - 1]);
						positions[u] = (uint_t<OutputIndexSize>)(staged.baselines[r] +
							load_offset(scatter_position_offset_bytes + staged.offset_starts[r], j, staged.widths[r]));
						sink.prepare(u, next_value + q);
					}
				}
				#pragma unroll
				for (unsigned u = 0; u < Unroll; u++) {
					const uint64_t q = q0 + u * warp_size + lane;
					if (q < limit) { sink.commit(u, next_value + q, positions[u]); }
				}
			}
			next_value += chunk_patches;
		}
	}
}

template <unsigned ElementSize, unsigned OutputIndexSize>
struct scatter_sink {
	uint_t<ElementSize>*        __restrict__ target;
	const uint_t<ElementSize>*  __restrict__ data;
	uint_t<ElementSize> values[ci_unroll];
	__device__ __forceinline__ void prepare(unsigned u, uint64_t k) { values[u] = __ldg(data + k); }
	__device__ __forceinline__ void commit(unsigned u, uint64_t, uint_t<OutputIndexSize> position) { target[position] = values[u]; }
};

// (the patched decoders decode the whole position list first: kernels/patched.cuh)
template <unsigned OutputIndexSize>
struct position_list_sink {
	uint_t<OutputIndexSize>* __restrict__ positions;
	__device__ __forceinline__ void prepare(unsigned, uint64_t) { }
	__device__ __forceinline__ void commit(unsigned, uint64_t k, uint_t<OutputIndexSize> position) { positions[k] = position; }
};

} // namespace ci_detail

template <unsigned OutputIndexSize, unsigned ElementSize, unsigned InputIndexSize, unsigned RunLengthSize>
__global__ void compressed_indices_scatter(
	uint_t<ElementSize>*            __restrict__  target,
	const uint_t<ElementSize>*      __restrict__  data,
	const uint_t<RunLengthSize>*    __restrict__  scatter_position_run_lengths,
	const unsigned char*            __restrict__  scatter_position_run_individual_offset_sizes,
	const uint_t<OutputIndexSize>*  __restrict__  scatter_position_run_baseline_values,
	const uint_t<InputIndexSize>*   __restrict__  scatter_position_run_offsets_start_positions,
	const unsigned char*            __restrict__  scatter_position_offset_bytes,
	const uint_t<InputIndexSize>*   __restrict__  scatter_position_anchors,
	uint_t<InputIndexSize>                        anchoring_period,
	uint_t<InputIndexSize>                        num_scatter_position_runs,
	size_type<InputIndexSize>                     data_length)
{
	__shared__ ci_detail::staged_runs_t staged_by_warp[32];
	ci_detail::scatter_sink<ElementSize, OutputIndexSize> sink;
	sink.target = target; sink.data = data;
	ci_detail::for_each_patch<OutputIndexSize, InputIndexSize, RunLengthSize, ci_detail::ci_unroll>(sink, staged_by_warp[warp_index_in_block()],
		scatter_position_run_lengths, scatter_position_run_individual_offset_sizes, scatter_position_run_baseline_values,
		scatter_position_run_offsets_start_positions, scatter_position_offset_bytes, scatter_position_anchors,
		anchoring_period, num_scatter_position_runs, data_length);
}

// The same walk, writing the decoded positions out as a plain list: positions[k] is where patch value k goes.
template <unsigned OutputIndexSize, unsigned InputIndexSize, unsigned RunLengthSize>
__global__ void expand_compressed_indices(
	uint_t<OutputIndexSize>*        __restrict__  positions,
	const uint_t<RunLengthSize>*    __restrict__  scatter_position_run_lengths,
	const unsigned char*            __restrict__  scatter_position_run_individual_offset_sizes,
	const uint_t<OutputIndexSize>*  __restrict__  scatter_position_run_baseline_values,
	const uint_t<InputIndexSize>*   __restrict__  scatter_position_run_offsets_start_positions,
	const unsigned char*            __restrict__  scatter_position_offset_bytes,
	const uint_t<InputIndexSize>*   __restrict__  scatter_position_anchors,
	uint_t<InputIndexSize>                        anchoring_period,
	uint_t<InputIndexSize>                        num_scatter_position_runs,
	size_type<InputIndexSize>                     data_length)
{
	__shared__ ci_detail::staged_runs_t staged_by_warp[32];
	ci_detail::position_list_sink<OutputIndexSize> sink;
	sink.positions = positions;
	ci_detail::for_each_patch<OutputIndexSize, InputIndexSize, RunLengthSize, ci_detail::ci_list_unroll>(sink, staged_by_warp[warp_index_in_block()],
		scatter_position_run_lengths, scatter_position_run_individual_offset_sizes, scatter_position_run_baseline_values,
		scatter_position_run_offsets_start_positions, scatter_position_offset_bytes, scatter_position_anchors,
		anchoring_period, num_scatter_position_runs, data_length);
}

} // namespace kernels
} // namespace giddy

#endif // GIDDY_B200_KERNELS_SCATTER_CUH_
