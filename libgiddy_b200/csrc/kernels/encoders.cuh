// Compressors: the inverse of every decompression scheme of this library, on the GPU, emitting exactly the layouts the
// decompressors read including the anchor arrays.
//
// The reference ships decompressors only ("compression ... is not included", README.md:128-130; its open Issue #3): there
// is no kernel here to mirror, only formats -- those of src/kernels/decompression/*.cuh, cited per encoder.  What an
// encoder must do beyond inverting the arithmetic is produce the side data which makes segments independently decodable:
// per-segment model coefficients (FOR), the value preceding each segment (Delta), the run index and intra-run offset of
// each segment's first element (RLE / RPE), the byte offset of each segment's first element (DZB-variable).
//
// Position-aligned encoders (Delta, DZB-fixed, Dictionary lookup, FOR) are single streaming kernels.  The encoders with
// data-dependent output positions (RLE, RPE, DZB-variable) are two passes over the column around an exclusive scan of
// per-tile (per-segment) counts: pass 1 counts, the scan places, pass 2 writes.  Every encoder raises bits of an
// optional status word instead of writing garbage when the data does not fit the format (a delta too wide for the
// compressed type, a value missing from the dictionary, more runs than the caller's arrays hold).
#pragma once
#ifndef GIDDY_B200_KERNELS_ENCODERS_CUH_
#define GIDDY_B200_KERNELS_ENCODERS_CUH_

#include "common.cuh"

namespace giddy {
namespace kernels {
namespace encoders {

using namespace giddy::device;

// bits of the status word (giddy_b200.h: GIDDY_COMPRESSION_*)
constexpr unsigned status_unrepresentable = 1u; // an element does not fit the compressed representation
constexpr unsigned status_capacity = 2u;        // the output arrays are too small for the column
constexpr unsigned status_index_overflow = 4u;  // an anchor does not fit its index type

__device__ __forceinline__ void raise(unsigned* status, unsigned bits) { if (status != nullptr) { atomicOr(status, bits); } }

template <unsigned Bytes>
__device__ __forceinline__ void store_pack(void* p, const byte_pack<Bytes>& v)
{
	if (Bytes == 1) { *static_cast<uint8_t*>(p) = v.u8[0]; }
	else if (Bytes == 2) { *static_cast<uint16_t*>(p) = v.u16[0]; }
	else if (Bytes == 4) { *static_cast<uint32_t*>(p) = v.u32[0]; }
	else if (Bytes == 8) { *static_cast<uint2*>(p) = make_uint2(v.u32[0], v.u32[1]); }
	else {
		#pragma unroll
		for (unsigned k = 0; k < Bytes / 16; k++) {
			store_streaming_16(static_cast<uint8_t*>(p) + 16 * k, make_uint4(v.u32[4 * k], v.u32[4 * k + 1], v.u32[4 * k + 2], v.u32[4 * k + 3]));
		}
	}
}

constexpr unsigned elementwise_unroll = 4;

// Skeleton of the position-aligned encoders: a thread reads 16 bytes of the column (E elements) per step, `unroll` steps
// in flight, and writes the E compressed elements with one access.  encode(i, x) -> the compressed element.
template <unsigned U, unsigned C, typename Encode>
__device__ __forceinline__ void encode_elementwise(uint_t<C>* __restrict__ compressed, const uint_t<U>* __restrict__ uncompressed, uint64_t length, Encode encode)
{
	constexpr unsigned E = 16 / U;
	const uint64_t global_thread = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
	const uint64_t num_threads = (uint64_t) gridDim.x * blockDim.x;
	if (!(is_aligned_to(uncompressed, 16) && is_aligned_to(compressed, E * C))) {
		for (uint64_t i = global_thread; i < length; i += num_threads) { compressed[i] = encode(i, uncompressed[i]); }
		return;
	}
	const uint64_t num_vectors = length / E;
	const uint64_t vectors_per_tile = (uint64_t) blockDim.x * elementwise_unroll;
	#pragma unroll 1
	for (uint64_t tile_start = (uint64_t) blockIdx.x * vectors_per_tile; tile_start < num_vectors; tile_start += (uint64_t) gridDim.x * vectors_per_tile) {
		byte_pack<16> in[elementwise_unroll];
		#pragma unroll
		for (unsigned k = 0; k < elementwise_unroll; k++) {
			const uint64_t v = tile_start + (uint64_t) k * blockDim.x + threadIdx.x;
			if (v < num_vectors) { in[k] = load_pack<16>(uncompressed + v * E); }
		}
		#pragma unroll
		for (unsigned k = 0; k < elementwise_unroll; k++) {
			const uint64_t v = tile_start + (uint64_t) k * blockDim.x + threadIdx.x;
			if (v >= num_vectors) { continue; }
			byte_pack<E * C> out;
			#pragma unroll
			for (unsigned j = 0; j < E; j++) { out.template set<C>(j, encode(v * E + j, in[k].template get<U>(j))); }
			store_pack<E * C>(compressed + v * E, out);
		}
	}
	const uint64_t tail = num_vectors * E + global_thread;
	if (tail < length) { compressed[tail] = encode(tail, uncompressed[tail]); }
}

// -------------------------------------------------------------------------------------------------
// Delta (format: reference src/kernels/decompression/delta.cuh:66-108):
//   compressed[i] = x[i] - x[i-1] (mod 2^(8U); x[-1] = base), anchor_values[s] = x[s*L - 1] (base for s = 0)
// -------------------------------------------------------------------------------------------------
template <unsigned IndexSize, unsigned U, unsigned C>
__global__ void delta_compress(
	uint_t<C>*        __restrict__  compressed,
	uint_t<U>*        __restrict__  anchor_values,
	const uint_t<U>*  __restrict__  uncompressed,
	size_type<IndexSize>            length,
	size_type<IndexSize>            segment_length,
	uint_t<U>                       base,
	unsigned*                       status)
{
	bool overflow = false;
	encode_elementwise<U, C>(compressed, uncompressed, length, [&](uint64_t i, uint_t<U> x) {
		const uint_t<U> previous = (i == 0) ? base : __ldg(uncompressed + i - 1);
		const uint_t<U> delta = (uint_t<U>)(x - previous);
		if (C < U && (uint64_t) delta >> (8 * C) != 0) { overflow = true; }
		return (uint_t<C>) delta;
	});
	if (overflow) { raise(status, status_unrepresentable); }
	const uint64_t num_segments = ((uint64_t) length + segment_length - 1) / segment_length;
	for (uint64_t s = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; s < num_segments; s += (uint64_t) gridDim.x * blockDim.x) {
		anchor_values[s] = (s == 0) ? base : __ldg(uncompressed + s * segment_length - 1);
	}
}

// -------------------------------------------------------------------------------------------------
// Discard zero bytes, fixed width (format: discard_zero_bytes_fixed.cuh:108-124): compressed[i] = the C low bytes of x[i]
// -------------------------------------------------------------------------------------------------
template <unsigned IndexSize, unsigned U, unsigned C>
__global__ void discard_zero_bytes_fixed_compress(
	uint_t<C>*        __restrict__  compressed,
	const uint_t<U>*  __restrict__  uncompressed,
	size_type<IndexSize>            length,
	unsigned*                       status)
{
	bool overflow = false;
	encode_elementwise<U, C>(compressed, uncompressed, length, [&](uint64_t, uint_t<U> x) {
		if ((uint64_t) x >> (8 * C) != 0) { overflow = true; }
		return (uint_t<C>) x;
	});
	if (overflow) { raise(status, status_unrepresentable); }
}

// -------------------------------------------------------------------------------------------------
// Dictionary (format: dictionary.cuh:61-79): compressed[i] = the index of x[i] in dictionary_entries, which must be
// ASCENDING (binary search); a value which is not an entry raises status_unrepresentable.  The dictionary is searched in
// shared memory when the launch brings enough of it.
// -------------------------------------------------------------------------------------------------
template <unsigned IndexSize, unsigned U, unsigned DictionaryIndexSize>
__global__ void dictionary_compress(
	uint_t<DictionaryIndexSize>*  __restrict__  compressed,
	const uint_t<U>*              __restrict__  uncompressed,
	const uint_t<U>*              __restrict__  dictionary_entries,
	size_type<DictionaryIndexSize>              num_dictionary_entries,
	size_type<IndexSize>                        length,
	unsigned*                                   status)
{
	const uint_t<U>* table = dictionary_entries;
	if ((uint64_t) dynamic_shared_memory_size() >= (uint64_t) num_dictionary_entries * U && num_dictionary_entries != 0) {
		auto cached = reinterpret_cast<uint_t<U>*>(dynamic_shared_memory);
		for (uint64_t i = threadIdx.x; i < num_dictionary_entries; i += blockDim.x) { cached[i] = __ldg(dictionary_entries + i); }
		__syncthreads();
		table = cached;
	}
	bool missing = false;
	encode_elementwise<U, DictionaryIndexSize>(compressed, uncompressed, length, [&](uint64_t, uint_t<U> x) {
		uint64_t low = 0, high = num_dictionary_entries; // first entry >= x
		while (low < high) {
			const uint64_t middle = (low + high) / 2;
			if (table[middle] < x) { low = middle + 1; } else { high = middle; }
		}
		if (low >= num_dictionary_entries || table[low] != x) { missing = true; low = 0; }
		return (uint_t<DictionaryIndexSize>) low;
	});
	if (missing) { raise(status, status_unrepresentable); }
}

// -------------------------------------------------------------------------------------------------
// Frame of Reference (format: frame_of_reference.cuh:35-80).  A warp per segment, two passes over it (the second one out
// of L1/L2): fit the model, then store the offsets.
//   constant  c0 = min(segment) + 2^(8C-1): offsets span the signed compressed range from its lower end
//   linear    c1 = round-half-even((last - first) / (L - 1)), c0 = first + min(residual) + 2^(8C-1), where the residual is
//             x - (first + c1 * position) and a short last segment counts as padded with the column's last element
// (exactly libgiddy_b200/encode.py frame_of_reference_encode).  An offset outside the compressed type raises
// status_unrepresentable.  Values are signed integers of U bytes; coefficients are int32.
// -------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T warp_min(T x)
{
	#pragma unroll
	for (unsigned d = 16; d > 0; d >>= 1) {
		const T y = (T) __shfl_xor_sync(full_mask, (long long) x, d);
		x = y < x ? y : x;
	}
	return x;
}

template <unsigned IndexSize, unsigned U, unsigned C, int Model>
__global__ void frame_of_reference_compress(
	int_t<C>*        __restrict__  compressed,
	int32_t*         __restrict__  segment_model_coefficients,
	const int_t<U>*  __restrict__  uncompressed,
	size_type<IndexSize>           length,
	size_type<IndexSize>           segment_length,
	unsigned*                      status)
{
	constexpr unsigned dim = (Model == model_linear) ? 2 : 1;
	const uint64_t L = segment_length, n = length;
	const uint64_t num_segments = (n + L - 1) / L;
	const int64_t lowest = -((int64_t) 1 << (8 * C - 1)), highest = ((int64_t) 1 << (8 * C - 1)) - 1;
	const unsigned lane = lane_index();
	bool overflow = false;
	for (uint64_t s = global_warp_index(); s < num_segments; s += num_warps_in_grid()) {
		const uint64_t start = s * L;
		const uint64_t count = minimum<uint64_t>(L, n - start);
		const int_t<U>* __restrict__ x = uncompressed + start;
		int64_t c0, c1 = 0;
		if (Model == model_constant) {
			int64_t lowest_value = INT64_MAX;
			for (uint64_t i = lane; i < count; i += warp_size) { const int64_t v = x[i]; lowest_value = v < lowest_value ? v : lowest_value; }
			c0 = warp_min(lowest_value) - lowest;
		}
		else {
			const int64_t first = x[0];
			const int64_t last = (count == L) ? (int64_t) x[L - 1] : (int64_t) uncompressed[n - 1];
			c1 = (int64_t) rint((double)(last - first) / (double)(L > 1 ? L - 1 : 1));
			int64_t lowest_residual = INT64_MAX;
			for (uint64_t i = lane; i < L; i += warp_size) {
				const int64_t v = (i < count) ? (int64_t) x[i] : (int64_t) uncompressed[n - 1];
				const int64_t residual = v - (first + c1 * (int64_t) i);
				lowest_residual = residual < lowest_residual ? residual : lowest_residual;
			}
			c0 = first + (warp_min(lowest_residual) - lowest);
		}
		if (lane == 0) {
			segment_model_coefficients[s * dim] = (int32_t) c0;
			if (dim == 2) { segment_model_coefficients[s * dim + 1] = (int32_t) c1; }
		}
		for (uint64_t i = lane; i < count; i += warp_size) {
			const int64_t offset = (int64_t) x[i] - (c0 + c1 * (int64_t) i);
			if (offset < lowest || offset > highest) { overflow = true; }
			compressed[start + i] = (int_t<C>) offset;
		}
	}
	if (overflow) { raise(status, status_unrepresentable); }
}

// -------------------------------------------------------------------------------------------------
// Exclusive prefix sums of a short array of counts by ONE block (the arrays scanned here hold one count per tile of
// 4096 elements or per segment: 2^18 .. 2^20 entries for a 2^30-element column).  out[i] = sum of in[0 .. i); out[count]
// = the total.  In place is fine.
// -------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T block_exclusive_sum(T value, T* shared_totals /* [33] */, T& block_total)
{
	const unsigned lane = lane_index(), warp = warp_index_in_block();
	const T inclusive = warp_inclusive_sum(value);
	if (lane == warp_size - 1) { shared_totals[warp] = inclusive; }
	__syncthreads();
	if (warp == 0) {
		const T total = (lane < warps_per_block()) ? shared_totals[lane] : (T) 0;
		const T scanned = warp_inclusive_sum(total);
		shared_totals[lane] = scanned - total;
		if (lane == warp_size - 1) { shared_totals[warp_size] = scanned; }
	}
	__syncthreads();
	const T result = shared_totals[warp] + inclusive - value;
	block_total = shared_totals[warp_size];
	__syncthreads();
	return result;
}

__global__ void exclusive_sum_by_one_block(uint64_t* __restrict__ out, const uint64_t* __restrict__ in, uint64_t count)
{
	__shared__ uint64_t totals[warp_size + 1];
	const uint64_t per_thread = (count + blockDim.x - 1) / blockDim.x;
	const uint64_t first = minimum<uint64_t>((uint64_t) threadIdx.x * per_thread, count);
	const uint64_t past_last = minimum<uint64_t>(first + per_thread, count);
	uint64_t sum = 0;
	for (uint64_t i = first; i < past_last; i++) { sum += in[i]; }
	uint64_t total;
	uint64_t running = block_exclusive_sum(sum, totals, total);
	for (uint64_t i = first; i < past_last; i++) { const uint64_t value = in[i]; out[i] = running; running += value; }
	if (threadIdx.x == 0) { out[count] = total; }
}

// -------------------------------------------------------------------------------------------------
// Run-Length and Run-Position encoding (formats: run_length_encoding.cuh:284-374, run_position_encoding.cuh:122-359).
//
// A run begins ("head") where an element differs from its predecessor, and -- when run lengths are limited to M elements
// (the run-length type's maximum, or the caller's) -- every M elements after that.  Tiles of 4096 elements:
//   pass 1  run_tile_summaries   per tile: position of its last natural head (or none), of its first, and the number of
//                                heads from its first natural head on (natural heads + the split points after them)
//   scan    run_tile_placement   one block: the run open at each tile's start (a max-scan of the last-head positions), hence
//                                the split points before the tile's first natural head, hence the tile's head count;
//                                exclusive sum -> the index of the tile's first run
//   pass 2  run_emit             per tile: heads again, a block-wide scan, and for every head run_data[r] = x[i],
//                                starts[r] = i (workspace)
//   then    run_finish           per run: its length (next start - start) or its stored start position
//           run_anchors          per segment: the run containing its first element (binary search of starts) and the offset
//                                within it
// -------------------------------------------------------------------------------------------------
constexpr unsigned run_tile_threads = 256;
constexpr unsigned run_tile_per_thread = 16;
constexpr unsigned run_tile_elements = run_tile_threads * run_tile_per_thread;
constexpr uint64_t no_position = ~(uint64_t) 0;

struct run_tile_summary_t { uint64_t first_head, last_head, heads_from_first; };

// number of multiples of `period` in [a, b) when counted from `origin` (positions origin + k * period, k >= 1)
__device__ __forceinline__ uint64_t split_points_between(uint64_t origin, uint64_t a, uint64_t b, uint64_t period)
{
	if (period == 0 || b <= a) { return 0; }
	// k ranges over positions p = origin + k * period with a <= p < b, k >= 1
	const uint64_t below_b = (b - 1 >= origin + period) ? (b - 1 - origin) / period : 0;  // k with p <= b - 1
	const uint64_t below_a = (a >= origin + period + 1) ? (a - 1 - origin) / period : 0;  // k with p <= a - 1
	return below_b - below_a;
}

template <unsigned U>
__global__ void __launch_bounds__(run_tile_threads) run_tile_summaries(
	run_tile_summary_t*  __restrict__  summaries,
	const uint_t<U>*     __restrict__  x,
	uint64_t                           length,
	uint64_t                           max_run_length)
{
	__shared__ uint64_t shared_first[run_tile_threads / warp_size], shared_last[run_tile_threads / warp_size];
	const uint64_t num_tiles = (length + run_tile_elements - 1) / run_tile_elements;
	for (uint64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
		const uint64_t tile_start = tile * run_tile_elements;
		const uint64_t tile_end = minimum<uint64_t>(tile_start + run_tile_elements, length);
		uint64_t first = no_position, last = no_position;
		// thread t looks at elements tile_start + t, + 256, ...: coalesced, and a thread's positions ascend
		for (unsigned k = 0; k < run_tile_per_thread; k++) {
			const uint64_t i = tile_start + (uint64_t) k * run_tile_threads + threadIdx.x;
			if (i >= tile_end) { break; }
			const bool head = i == 0 || x[i] != x[i - 1];
			if (head) { if (first == no_position) { first = i; } last = i; }
		}
		// block-wide min of first, max of last (no_position = "none" loses both ways)
		uint64_t last_key = (last == no_position) ? 0 : last + 1;
		#pragma unroll
		for (unsigned d = 16; d > 0; d >>= 1) {
			const uint64_t f = __shfl_xor_sync(full_mask, (unsigned long long) first, d);
			const uint64_t l = __shfl_xor_sync(full_mask, (unsigned long long) last_key, d);
			first = f < first ? f : first;
			last_key = l > last_key ? l : last_key;
		}
		if (lane_index() == 0) { shared_first[warp_index_in_block()] = first; shared_last[warp_index_in_block()] = last_key; }
		__syncthreads();
		if (threadIdx.x == 0) {
			for (unsigned w = 1; w < run_tile_threads / warp_size; w++) {
				first = shared_first[w] < first ? shared_first[w] : first;
				last_key = shared_last[w] > last_key ? shared_last[w] : last_key;
			}
			shared_first[0] = first; shared_last[0] = last_key;
		}
		__syncthreads();
		first = shared_first[0];
		last = shared_last[0] == 0 ? no_position : shared_last[0] - 1;
		__syncthreads();
		// heads from the first natural head on: natural heads, plus the split points inside each run (known locally: the
		// run's start is in this tile).  Counted per natural head: the split points between it and the next head (or tile end)
		uint64_t heads = 0;
		if (first != no_position) {
			for (unsigned k = 0; k < run_tile_per_thread; k++) {
				const uint64_t i = tile_start + (uint64_t) k * run_tile_threads + threadIdx.x;
				if (i >= tile_end) { break; }
				if (i == 0 || x[i] != x[i - 1]) {
					heads++;
					if (max_run_length != 0) { // this run's end within the tile: scan forward to the next head (runs are short: rare loop)
						uint64_t end = i + 1;
						while (end < tile_end && x[end] == x[end - 1]) { end++; }
						heads += split_points_between(i, i + 1, end, max_run_length);
					}
				}
			}
		}
		// block sum
		__shared__ uint64_t totals[warp_size + 1];
		uint64_t total;
		block_exclusive_sum(heads, totals, total);
		if (threadIdx.x == 0) { summaries[tile] = { first, last, total }; }
	}
}

// one block; placement[t] = index of tile t's first run; placement[num_tiles] = the number of runs; open_run_start[t] = start
// of the run which contains the element before tile t (undefined for tile 0)
__global__ void run_tile_placement(
	uint64_t*                  __restrict__  placement,
	uint64_t*                  __restrict__  open_run_start,
	const run_tile_summary_t*  __restrict__  summaries,
	uint64_t                                 num_tiles,
	uint64_t                                 length,
	uint64_t                                 max_run_length)
{
	__shared__ uint64_t totals[warp_size + 1];
	__shared__ uint64_t carry_last[warp_size + 1];
	const uint64_t per_thread = (num_tiles + blockDim.x - 1) / blockDim.x;
	const uint64_t first = minimum<uint64_t>((uint64_t) threadIdx.x * per_thread, num_tiles);
	const uint64_t past_last = minimum<uint64_t>(first + per_thread, num_tiles);
	// exclusive max-scan of (last natural head + 1): the natural run open at each tile's start
	uint64_t my_last = 0;
	for (uint64_t t = first; t < past_last; t++) { const uint64_t l = summaries[t].last_head; if (l != no_position) { my_last = l + 1; } }
	// scan of "last non-zero" across threads: carry = the nearest preceding thread's non-zero value
	{
		const unsigned lane = lane_index(), warp = warp_index_in_block();
		uint64_t v = my_last;
		#pragma unroll
		for (unsigned d = 1; d < warp_size; d <<= 1) {
			const uint64_t y = __shfl_up_sync(full_mask, (unsigned long long) v, d);
			if (lane >= d && v == 0) { v = y; }
		}
		if (lane == warp_size - 1) { carry_last[warp] = v; }
		__syncthreads();
		if (warp == 0) {
			uint64_t w = (lane < warps_per_block()) ? carry_last[lane] : 0;
			#pragma unroll
			for (unsigned d = 1; d < warp_size; d <<= 1) {
				const uint64_t y = __shfl_up_sync(full_mask, (unsigned long long) w, d);
				if (lane >= d && w == 0) { w = y; }
			}
			carry_last[lane] = w;
		}
		__syncthreads();
		uint64_t before = __shfl_up_sync(full_mask, (unsigned long long) v, 1); // inclusive value of the previous lane
		if (lane == 0) { before = (warp == 0) ? 0 : carry_last[warp - 1]; }
		else if (before == 0 && warp != 0) { before = carry_last[warp - 1]; }
		my_last = before; // the value carried INTO this thread's first tile
		__syncthreads();
	}
	// per tile: its head count given the open run; thread-local sums
	uint64_t sum = 0, open = my_last;
	for (uint64_t t = first; t < past_last; t++) {
		const run_tile_summary_t s = summaries[t];
		const uint64_t tile_start = t * run_tile_elements;
		const uint64_t tile_end = minimum<uint64_t>(tile_start + run_tile_elements, length);
		const uint64_t prefix_end = (s.first_head == no_position) ? tile_end : s.first_head;
		const uint64_t origin = (open == 0) ? 0 : open - 1; // (tile 0 begins with a head at position 0: no prefix there)
		sum += s.heads_from_first + ((t == 0) ? 0 : split_points_between(origin, tile_start, prefix_end, max_run_length));
		if (s.last_head != no_position) { open = s.last_head + 1; }
	}
	uint64_t total;
	uint64_t running = block_exclusive_sum(sum, totals, total);
	open = my_last;
	for (uint64_t t = first; t < past_last; t++) {
		const run_tile_summary_t s = summaries[t];
		const uint64_t tile_start = t * run_tile_elements;
		const uint64_t tile_end = minimum<uint64_t>(tile_start + run_tile_elements, length);
		const uint64_t prefix_end = (s.first_head == no_position) ? tile_end : s.first_head;
		const uint64_t origin = (open == 0) ? 0 : open - 1;
		placement[t] = running;
		open_run_start[t] = origin;
		running += s.heads_from_first + ((t == 0) ? 0 : split_points_between(origin, tile_start, prefix_end, max_run_length));
		if (s.last_head != no_position) { open = s.last_head + 1; }
	}
	if (threadIdx.x == 0) { placement[num_tiles] = total; }
}

// pass 2: every head writes its run's value and start.  A thread owns run_tile_per_thread CONSECUTIVE elements here, so
// that the start of the run it is in carries from element to element.
template <unsigned U, typename Start>
__global__ void __launch_bounds__(run_tile_threads) run_emit(
	uint_t<U>*        __restrict__  run_data,
	Start*            __restrict__  starts,          // workspace: start position of every run, + one entry = length
	const uint_t<U>*  __restrict__  x,
	const uint64_t*   __restrict__  placement,
	const uint64_t*   __restrict__  open_run_start,
	uint64_t                        length,
	uint64_t                        max_run_length,
	uint64_t                        run_capacity,
	unsigned*                       status)
{
	__shared__ uint64_t totals[warp_size + 1];
	__shared__ uint64_t shared_open[run_tile_threads];
	const uint64_t num_tiles = (length + run_tile_elements - 1) / run_tile_elements;
	for (uint64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
		const uint64_t tile_start = tile * run_tile_elements;
		const uint64_t tile_end = minimum<uint64_t>(tile_start + run_tile_elements, length);
		const uint64_t first = tile_start + (uint64_t) threadIdx.x * run_tile_per_thread;
		const uint64_t past_last = minimum<uint64_t>(first + run_tile_per_thread, tile_end);
		// the natural run open at this thread's first element: the last natural head before it within the tile, else the tile's
		uint64_t my_last_natural = 0; // position + 1 of this thread's last natural head
		for (uint64_t i = first; i < past_last; i++) { if (i == 0 || x[i] != x[i - 1]) { my_last_natural = i + 1; } }
		shared_open[threadIdx.x] = my_last_natural;
		__syncthreads();
		uint64_t open = 0;
		for (int t = (int) threadIdx.x - 1; t >= 0; t--) { if (shared_open[t] != 0) { open = shared_open[t]; break; } } // (short: heads are dense)
		uint64_t run_start = (open != 0) ? open - 1 : open_run_start[tile];
		__syncthreads();
		// count this thread's heads
		uint64_t heads = 0, start = run_start;
		for (uint64_t i = first; i < past_last; i++) {
			const bool natural = i == 0 || x[i] != x[i - 1];
			if (natural) { start = i; }
			if (natural || (max_run_length != 0 && (i - start) % max_run_length == 0)) { heads++; }
		}
		uint64_t total;
		uint64_t index = placement[tile] + block_exclusive_sum(heads, totals, total);
		start = run_start;
		for (uint64_t i = first; i < past_last; i++) {
			const bool natural = i == 0 || x[i] != x[i - 1];
			if (natural) { start = i; }
			if (natural || (max_run_length != 0 && (i - start) % max_run_length == 0)) {
				if (index < run_capacity) { run_data[index] = x[i]; starts[index] = (Start) i; }
				else { raise(status, status_capacity); }
				index++;
			}
		}
	}
}

// per run: the stored length (RLE) or start position (RPE); also the run count, narrowed to the caller's type
template <unsigned RL, typename Start>
__global__ void run_lengths_from_starts(
	uint_t<RL>*      __restrict__  run_lengths,
	const Start*     __restrict__  starts,
	const uint64_t*  __restrict__  placement_total, // &placement[num_tiles]
	uint64_t                       length,
	uint64_t                       run_capacity)
{
	const uint64_t num_runs = minimum<uint64_t>(*placement_total, run_capacity);
	for (uint64_t r = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; r < num_runs; r += (uint64_t) gridDim.x * blockDim.x) {
		const uint64_t next = (r + 1 < num_runs) ? (uint64_t) starts[r + 1] : length;
		run_lengths[r] = (uint_t<RL>)(next - (uint64_t) starts[r]);
	}
}

template <unsigned PO, typename Start>
__global__ void run_positions_from_starts(
	uint_t<PO>*      __restrict__  run_start_positions,
	const Start*     __restrict__  starts,
	const uint64_t*  __restrict__  placement_total,
	uint64_t                       segment_length,
	bool                           relative,
	uint64_t                       run_capacity,
	unsigned*                      status)
{
	const uint64_t num_runs = minimum<uint64_t>(*placement_total, run_capacity);
	bool overflow = false;
	for (uint64_t r = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; r < num_runs; r += (uint64_t) gridDim.x * blockDim.x) {
		const uint64_t start = starts[r];
		const uint64_t stored = relative ? start % segment_length : start;
		if (PO < 8 && (stored >> (8 * PO)) != 0) { overflow = true; }
		run_start_positions[r] = (uint_t<PO>) stored;
	}
	if (overflow) { raise(status, status_unrepresentable); }
}

// per segment: the run which contains its first element = (number of runs starting at or before s * L) - 1
template <unsigned RunIndexSize, unsigned RL, typename Start, bool WithOffsets>
__global__ void run_anchors(
	uint_t<RunIndexSize>*  __restrict__  position_anchors,
	uint_t<RL>*            __restrict__  intra_run_anchor_offsets,
	void*                                num_element_runs_out,  // size_type<RunIndexSize>*, optional
	const Start*           __restrict__  starts,
	const uint64_t*        __restrict__  placement_total,
	uint64_t                             length,
	uint64_t                             segment_length,
	uint64_t                             run_capacity,
	unsigned*                            status)
{
	const uint64_t total = *placement_total;
	const uint64_t num_runs = minimum<uint64_t>(total, run_capacity);
	const uint64_t num_segments = (length + segment_length - 1) / segment_length;
	const uint64_t global_thread = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
	if (global_thread == 0) {
		if (num_element_runs_out != nullptr) { *static_cast<size_type<RunIndexSize>*>(num_element_runs_out) = (size_type<RunIndexSize>) total; }
		if (total > run_capacity) { raise(status, status_capacity); }
		if (RunIndexSize < 8 && (total >> (8 * RunIndexSize)) != 0) { raise(status, status_index_overflow); }
	}
	for (uint64_t s = global_thread; s < num_segments; s += (uint64_t) gridDim.x * blockDim.x) {
		const uint64_t position = s * segment_length;
		uint64_t low = 0, high = num_runs; // first run with start > position
		while (low < high) {
			const uint64_t middle = (low + high) / 2;
			if ((uint64_t) starts[middle] <= position) { low = middle + 1; } else { high = middle; }
		}
		const uint64_t run = low - 1; // (run 0 starts at 0: low >= 1)
		position_anchors[s] = (uint_t<RunIndexSize>) run;
		if (WithOffsets) { intra_run_anchor_offsets[s] = (uint_t<RL>)(position - (uint64_t) starts[run]); }
	}
}

// -------------------------------------------------------------------------------------------------
// Discard zero bytes, variable width (format: discard_zero_bytes_variable.cuh:335-439).  Element i is stored in
// size_i = max(significant bytes of x[i], min_size) bytes; size_i - min_size goes into a `bits`-wide field.
//   pass 1  variable_width_segment_bytes   a warp per segment: the bytes its elements take
//   scan    exclusive_sum_by_one_block     -> byte offset of every segment = position_anchors
//   pass 2  variable_width_emit            a warp per segment: sizes again, a running offset, the bytes
//           variable_width_pack_sizes      a thread per size container
// -------------------------------------------------------------------------------------------------
template <unsigned U>
__device__ __forceinline__ unsigned significant_bytes(uint_t<U> x)
{
	if (U == 8) { return x == 0 ? 0u : (unsigned)((64 - __clzll((long long) x) + 7) / 8); }
	return x == 0 ? 0u : (unsigned)((32 - __clz((int)(uint32_t) x) + 7) / 8);
}

template <unsigned U>
__global__ void variable_width_segment_bytes(
	uint64_t*         __restrict__  segment_bytes,
	const uint_t<U>*  __restrict__  x,
	uint64_t                        length,
	uint64_t                        segment_length,
	unsigned                        min_size,
	unsigned                        bits,
	unsigned*                       status)
{
	const uint64_t num_segments = (length + segment_length - 1) / segment_length;
	bool unrepresentable = false;
	for (uint64_t s = global_warp_index(); s < num_segments; s += num_warps_in_grid()) {
		const uint64_t start = s * segment_length;
		const uint64_t count = minimum<uint64_t>(segment_length, length - start);
		uint64_t bytes = 0;
		for (uint64_t i = lane_index(); i < count; i += warp_size) {
			const unsigned size = max(significant_bytes<U>(x[start + i]), min_size);
			if (bits < 32 && ((size - min_size) >> bits) != 0) { unrepresentable = true; }
			bytes += size;
		}
		bytes = warp_inclusive_sum(bytes);
		if (lane_index() == warp_size - 1) { segment_bytes[s] = bytes; }
	}
	if (unrepresentable) { raise(status, status_unrepresentable); }
}

template <unsigned IndexSize, unsigned U>
__global__ void variable_width_emit(
	uint8_t*           __restrict__  compressed_data,
	uint_t<IndexSize>* __restrict__  position_anchors,
	uint64_t*                        num_bytes_out,      // optional
	const uint_t<U>*   __restrict__  x,
	const uint64_t*    __restrict__  segment_offsets,    // exclusive sums of segment_bytes, + the total
	uint64_t                         length,
	uint64_t                         segment_length,
	unsigned                         min_size,
	uint64_t                         data_capacity,
	unsigned*                        status)
{
	const uint64_t num_segments = (length + segment_length - 1) / segment_length;
	const unsigned lane = lane_index();
	if (blockIdx.x == 0 && threadIdx.x == 0) {
		const uint64_t total = segment_offsets[num_segments];
		if (num_bytes_out != nullptr) { *num_bytes_out = total; }
		if (total > data_capacity) { raise(status, status_capacity); }
	}
	bool index_overflow = false;
	for (uint64_t s = global_warp_index(); s < num_segments; s += num_warps_in_grid()) {
		const uint64_t start = s * segment_length;
		const uint64_t count = minimum<uint64_t>(segment_length, length - start);
		uint64_t offset = segment_offsets[s];
		if (lane == 0) {
			if (IndexSize < 8 && (offset >> (8 * IndexSize)) != 0) { index_overflow = true; }
			position_anchors[s] = (uint_t<IndexSize>) offset;
		}
		for (uint64_t base = 0; base < count; base += warp_size) {
			const uint64_t i = base + lane;
			const uint_t<U> value = (i < count) ? x[start + i] : (uint_t<U>) 0;
			const unsigned size = (i < count) ? max(significant_bytes<U>(value), min_size) : 0u;
			const unsigned inclusive = warp_inclusive_sum(size);
			const uint64_t mine = offset + inclusive - size;
			if (mine + size <= data_capacity) {
				for (unsigned b = 0; b < size; b++) { compressed_data[mine + b] = (uint8_t)((uint64_t) value >> (8 * b)); }
			}
			offset += __shfl_sync(full_mask, inclusive, warp_size - 1);
		}
	}
	if (index_overflow) { raise(status, status_index_overflow); }
}

template <unsigned U, unsigned CS>
__global__ void variable_width_pack_sizes(
	uint_t<CS>*       __restrict__  packed_element_sizes,
	const uint_t<U>*  __restrict__  x,
	uint64_t                        length,
	unsigned                        min_size,
	unsigned                        bits)
{
	const uint64_t per_container = (8 * CS) / bits;
	const uint64_t num_containers = (length + per_container - 1) / per_container;
	const uint64_t field_mask = (bits >= 64) ? ~0ull : ((1ull << bits) - 1);
	for (uint64_t c = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; c < num_containers; c += (uint64_t) gridDim.x * blockDim.x) {
		uint64_t container = 0;
		for (uint64_t k = 0; k < per_container; k++) {
			const uint64_t i = c * per_container + k;
			if (i >= length) { break; }
			const unsigned size = max(significant_bytes<U>(x[i]), min_size);
			container |= ((uint64_t)(size - min_size) & field_mask) << (k * bits);
		}
		packed_element_sizes[c] = (uint_t<CS>) container;
	}
}

// -------------------------------------------------------------------------------------------------
// Incidence bitmaps (format: incidence_bitmaps.cuh:69-131): bit (i mod 32) of word i / 32 of bitmap b is set iff
// x[i] == dictionary_entries[b] (UseDictionary) or x[i] == b.  A warp per 32-element container: one ballot per bitmap.
// -------------------------------------------------------------------------------------------------
template <unsigned U, bool UseDictionary>
__global__ void incidence_bitmaps_compress(
	uint32_t*         __restrict__  bitmaps,
	const uint_t<U>*  __restrict__  x,
	const uint_t<U>*  __restrict__  dictionary_entries,
	uint64_t                        length,
	uint64_t                        bitmap_pitch,   // containers between consecutive bitmaps
	unsigned                        num_bitmaps,
	unsigned*                       status)
{
	const uint64_t num_containers = (length + 31) / 32;
	bool missing = false;
	for (uint64_t w = global_warp_index(); w < num_containers; w += num_warps_in_grid()) {
		const uint64_t i = w * 32 + lane_index();
		const bool inside = i < length;
		const uint_t<U> value = inside ? x[i] : (uint_t<U>) 0;
		bool found = !inside;
		for (unsigned b = 0; b < num_bitmaps; b++) {
			const uint_t<U> entry = UseDictionary ? __ldg(dictionary_entries + b) : (uint_t<U>) b;
			const bool match = inside && value == entry && !found; // (duplicate dictionary entries: the first one wins)
			found = found || match;
			const uint32_t word = __ballot_sync(full_mask, match);
			if (lane_index() == 0) { bitmaps[(uint64_t) b * bitmap_pitch + w] = word; }
		}
		if (!found) { missing = true; }
	}
	if (missing) { raise(status, status_unrepresentable); }
}

} // namespace encoders
} // namespace kernels
} // namespace giddy

#endif // GIDDY_B200_KERNELS_ENCODERS_CUH_
