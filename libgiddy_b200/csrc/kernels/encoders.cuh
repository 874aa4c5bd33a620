// Compressors: the inverse of every decompression scheme of this library, on the GPU, emitting exactly the layouts the
// decompressors read including the anchor arrays.
//
// The reference ships decompressors only ("compression ... is not included", README.md:128-130; its open Issue #3): there
// is no kernel here to mirror, only formats -- those of src/kernels/decompression/*.cuh, cited per encoder.  What an
// encoder must do beyond inverting the arithmetic is produce the side data which makes segments independently decodable:
// per-segment model coefficients (FOR), the value preceding each segment (Delta), the run index and intra-run offset of
// each segment's first element (RLE / RPE), the byte offset of each segment's first element (DZB-variable).
//
// Position-aligned encoders (Delta, DZB-fixed, Dictionary lookup, FOR) are single streaming kernels.  Of the encoders with
// data-dependent output positions, RLE and RPE take ONE pass with decoupled look-back over super-tiles (run_encode_single_pass)
// unless runs are to be cut shorter than a super-tile, and then -- like DZB-variable always -- two passes over the column
// around an exclusive scan of per-tile (per-segment) counts: pass 1 counts, the scan places, pass 2 writes.  Every encoder raises bits of an
// optional status word instead of writing garbage when the data does not fit the format (a delta too wide for the
// compressed type, a value missing from the dictionary, more runs than the caller's arrays hold).
#pragma once
#ifndef GIDDY_B200_KERNELS_ENCODERS_CUH_
#define GIDDY_B200_KERNELS_ENCODERS_CUH_

#include "common.cuh"
#include "delta.cuh" // the tile descriptors of the decoupled look-back

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
// The skeleton of encode_elementwise; the element before a lane's vector comes from the lane below by shuffle (lane 0 of a
// warp loads it), so the column is read once.
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
	constexpr unsigned E = 16 / U;
	bool overflow = false;
	auto encode = [&](uint_t<U> x, uint_t<U> previous) {
		const uint_t<U> delta = (uint_t<U>)(x - previous);
		if (C < U && (uint64_t) delta >> (8 * C) != 0) { overflow = true; }
		return (uint_t<C>) delta;
	};
	const uint64_t global_thread = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
	const uint64_t num_threads = (uint64_t) gridDim.x * blockDim.x;
	const uint64_t n = length;
	if (!(is_aligned_to(uncompressed, 16) && is_aligned_to(compressed, E * C)) || blockDim.x % warp_size != 0) {
		for (uint64_t i = global_thread; i < n; i += num_threads) { compressed[i] = encode(uncompressed[i], (i == 0) ? base : uncompressed[i - 1]); }
	}
	else {
		const uint64_t num_vectors = n / E;
		const uint64_t vectors_per_tile = (uint64_t) blockDim.x * elementwise_unroll;
		#pragma unroll 1
		for (uint64_t tile_start = (uint64_t) blockIdx.x * vectors_per_tile; tile_start < num_vectors; tile_start += (uint64_t) gridDim.x * vectors_per_tile) {
			byte_pack<16> in[elementwise_unroll];
			uint_t<U> before[elementwise_unroll]; // lane 0: the element before the warp's vectors of step k
			#pragma unroll
			for (unsigned k = 0; k < elementwise_unroll; k++) {
				const uint64_t v = tile_start + (uint64_t) k * blockDim.x + threadIdx.x;
				if (v < num_vectors) { in[k] = load_pack<16>(uncompressed + v * E); }
				before[k] = 0;
				if (lane_index() == 0 && v < num_vectors) { before[k] = (v == 0) ? base : __ldg(uncompressed + v * E - 1); }
			}
			#pragma unroll
			for (unsigned k = 0; k < elementwise_unroll; k++) {
				const uint64_t v = tile_start + (uint64_t) k * blockDim.x + threadIdx.x;
				// (the shuffle is executed by the whole warp; lanes past the end hold nothing anybody needs)
				const uint_t<U> last = (v < num_vectors) ? in[k].template get<U>(E - 1) : (uint_t<U>) 0;
				const uint_t<U> from_below = shuffle_up(last, 1);
				if (v >= num_vectors) { continue; }
				uint_t<U> previous = (lane_index() == 0) ? before[k] : from_below;
				byte_pack<E * C> out;
				#pragma unroll
				for (unsigned j = 0; j < E; j++) {
					const uint_t<U> x = in[k].template get<U>(j);
					out.template set<C>(j, encode(x, previous));
					previous = x;
				}
				store_pack<E * C>(compressed + v * E, out);
			}
		}
		const uint64_t tail = num_vectors * E + global_thread;
		if (tail < n) { compressed[tail] = encode(uncompressed[tail], (tail == 0) ? base : uncompressed[tail - 1]); }
	}
	if (overflow) { raise(status, status_unrepresentable); }
	const uint64_t num_segments = (n + segment_length - 1) / segment_length;
	for (uint64_t s = global_thread; s < num_segments; s += num_threads) {
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
constexpr unsigned dictionary_hash_max_entries = 1024;              // dictionaries up to this size are looked up through a hash table
constexpr unsigned dictionary_hash_slots = 2 * dictionary_hash_max_entries;

template <unsigned U>
__device__ __forceinline__ unsigned dictionary_hash(uint_t<U> x)
{
	uint64_t h = (uint64_t) x * 0x9E3779B97F4A7C15ull; // Fibonacci hashing: the product's high bits
	return (unsigned)(h >> (64 - 11)) & (dictionary_hash_slots - 1);
}

template <unsigned IndexSize, unsigned U, unsigned DictionaryIndexSize>
__global__ void dictionary_compress(
	uint_t<DictionaryIndexSize>*  __restrict__  compressed,
	const uint_t<U>*              __restrict__  uncompressed,
	const uint_t<U>*              __restrict__  dictionary_entries,
	size_type<DictionaryIndexSize>              num_dictionary_entries,
	size_type<IndexSize>                        length,
	unsigned*                                   status)
{
	static_assert(dictionary_hash_slots == 2048, "the hash takes 11 bits");
	bool missing = false;
	if (num_dictionary_entries != 0 && num_dictionary_entries <= dictionary_hash_max_entries) {
		// Small dictionaries (the usual case: 256 entries for 1-byte indices): an open-addressing hash table in shared memory,
		// built by every block -- ~1.3 probes per element where a binary search takes log2(D) dependent steps.
		__shared__ uint_t<U> slot_keys[dictionary_hash_slots];
		__shared__ unsigned  slot_indices[dictionary_hash_slots];
		constexpr unsigned empty = 0xffffffffu;
		for (unsigned i = threadIdx.x; i < dictionary_hash_slots; i += blockDim.x) { slot_indices[i] = empty; }
		__syncthreads();
		// (the dictionary ascends, so equal entries are neighbours: only the first of them is inserted -- the lowest index wins, as
		// in the search -- and no insertion ever needs to look at another's key)
		for (unsigned i = threadIdx.x; i < num_dictionary_entries; i += blockDim.x) {
			const uint_t<U> key = __ldg(dictionary_entries + i);
			if (i != 0 && __ldg(dictionary_entries + i - 1) == key) { continue; }
			unsigned h = dictionary_hash<U>(key);
			while (atomicCAS(&slot_indices[h], empty, i) != empty) { h = (h + 1) & (dictionary_hash_slots - 1); }
			slot_keys[h] = key;
		}
		__syncthreads();
		encode_elementwise<U, DictionaryIndexSize>(compressed, uncompressed, length, [&](uint64_t, uint_t<U> x) {
			unsigned h = dictionary_hash<U>(x);
			while (true) {
				const unsigned index = slot_indices[h];
				if (index == empty) { missing = true; return (uint_t<DictionaryIndexSize>) 0; }
				if (slot_keys[h] == x) { return (uint_t<DictionaryIndexSize>) index; }
				h = (h + 1) & (dictionary_hash_slots - 1);
			}
		});
		if (missing) { raise(status, status_unrepresentable); }
		return;
	}
	const uint_t<U>* table = dictionary_entries;
	if ((uint64_t) dynamic_shared_memory_size() >= (uint64_t) num_dictionary_entries * U && num_dictionary_entries != 0) {
		auto cached = reinterpret_cast<uint_t<U>*>(dynamic_shared_memory);
		for (uint64_t i = threadIdx.x; i < num_dictionary_entries; i += blockDim.x) { cached[i] = __ldg(dictionary_entries + i); }
		__syncthreads();
		table = cached;
	}
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
	constexpr unsigned E = 16 / U; // elements per 16-byte load
	for (uint64_t s = global_warp_index(); s < num_segments; s += num_warps_in_grid()) {
		const uint64_t start = s * L;
		const uint64_t count = minimum<uint64_t>(L, n - start);
		const int_t<U>* __restrict__ x = uncompressed + start;
		// whole vectors of the segment go through 16-byte loads (and E * C-byte stores); the rest element by element
		const bool vectors = is_aligned_to(x, 16) && is_aligned_to(compressed + start, E * C);
		const uint64_t num_vectors = vectors ? count / E : 0;
		int64_t c0, c1 = 0;
		if (Model == model_constant) {
			int64_t lowest_value = INT64_MAX;
			for (uint64_t v = lane; v < num_vectors; v += warp_size) {
				const byte_pack<16> pack = load_pack<16>(x + v * E);
				#pragma unroll
				for (unsigned j = 0; j < E; j++) { const int64_t value = (int_t<U>) pack.template get<U>(j); lowest_value = value < lowest_value ? value : lowest_value; }
			}
			for (uint64_t i = num_vectors * E + lane; i < count; i += warp_size) { const int64_t value = x[i]; lowest_value = value < lowest_value ? value : lowest_value; }
			c0 = warp_min(lowest_value) - lowest;
		}
		else {
			const int64_t first = x[0];
			const int64_t last = (count == L) ? (int64_t) x[L - 1] : (int64_t) uncompressed[n - 1];
			c1 = (int64_t) rint((double)(last - first) / (double)(L > 1 ? L - 1 : 1));
			int64_t lowest_residual = INT64_MAX;
			for (uint64_t v = lane; v < num_vectors; v += warp_size) {
				const byte_pack<16> pack = load_pack<16>(x + v * E);
				#pragma unroll
				for (unsigned j = 0; j < E; j++) {
					const int64_t residual = (int64_t)(int_t<U>) pack.template get<U>(j) - (first + c1 * (int64_t)(v * E + j));
					lowest_residual = residual < lowest_residual ? residual : lowest_residual;
				}
			}
			for (uint64_t i = num_vectors * E + lane; i < L; i += warp_size) { // (a short last segment counts as padded with the column's last element)
				const int64_t value = (i < count) ? (int64_t) x[i] : (int64_t) uncompressed[n - 1];
				const int64_t residual = value - (first + c1 * (int64_t) i);
				lowest_residual = residual < lowest_residual ? residual : lowest_residual;
			}
			c0 = first + (warp_min(lowest_residual) - lowest);
		}
		if (lane == 0) {
			segment_model_coefficients[s * dim] = (int32_t) c0;
			if (dim == 2) { segment_model_coefficients[s * dim + 1] = (int32_t) c1; }
		}
		for (uint64_t v = lane; v < num_vectors; v += warp_size) { // the second pass over the segment: out of L1 / L2
			const byte_pack<16> pack = load_pack<16>(x + v * E);
			byte_pack<E * C> out;
			#pragma unroll
			for (unsigned j = 0; j < E; j++) {
				const int64_t offset = (int64_t)(int_t<U>) pack.template get<U>(j) - (c0 + c1 * (int64_t)(v * E + j));
				if (offset < lowest || offset > highest) { overflow = true; }
				out.template set<C>(j, (uint_t<C>)(int_t<C>) offset);
			}
			store_pack<E * C>(compressed + start + v * E, out);
		}
		for (uint64_t i = num_vectors * E + lane; i < count; i += warp_size) {
			const int64_t offset = (int64_t) x[i] - (c0 + c1 * (int64_t) i);
			if (offset < lowest || offset > highest) { overflow = true; }
			compressed[start + i] = (int_t<C>) offset;
		}
	}
	if (overflow) { raise(status, status_unrepresentable); }
}

constexpr unsigned for_block_threads = 256;
constexpr unsigned for_block_vectors = 8; // 16-byte vectors of a segment a thread holds

// A CTA per segment, the segment's 16-byte vectors held in registers between the fit and the store: the column is read once
// (frame_of_reference_compress reads a segment twice, the second time out of L2).  Whole segments only, of 16-byte aligned
// columns, with at most for_block_threads x for_block_vectors vectors each (the launcher checks, and leaves a short last
// segment to frame_of_reference_compress).
template <unsigned IndexSize, unsigned U, unsigned C, int Model>
__global__ void __launch_bounds__(for_block_threads) frame_of_reference_compress_by_block(
	int_t<C>*        __restrict__  compressed,
	int32_t*         __restrict__  segment_model_coefficients,
	const int_t<U>*  __restrict__  uncompressed,
	size_type<IndexSize>           length,
	size_type<IndexSize>           segment_length,
	unsigned*                      status)
{
	constexpr unsigned dim = (Model == model_linear) ? 2 : 1;
	constexpr unsigned E = 16 / U;
	__shared__ int64_t warp_lowest[2][for_block_threads / warp_size]; // (double-buffered by segment parity: one barrier per segment)
	const uint64_t L = segment_length;
	const uint64_t whole_segments = (uint64_t) length / L;
	const unsigned num_vectors = (unsigned)(L / E);
	const int64_t lowest = -((int64_t) 1 << (8 * C - 1)), highest = ((int64_t) 1 << (8 * C - 1)) - 1;
	bool overflow = false;
	unsigned parity = 0;
	for (uint64_t s = blockIdx.x; s < whole_segments; s += gridDim.x, parity ^= 1u) {
		const int_t<U>* __restrict__ x = uncompressed + s * L;
		byte_pack<16> held[for_block_vectors];
		#pragma unroll
		for (unsigned k = 0; k < for_block_vectors; k++) {
			const unsigned v = k * for_block_threads + threadIdx.x;
			if (v < num_vectors) { held[k] = load_pack<16>(x + (uint64_t) v * E); }
		}
		int64_t first = 0, c1 = 0;
		if (Model == model_linear) {
			first = __ldg(x);
			const int64_t last = __ldg(x + L - 1);
			c1 = (int64_t) rint((double)(last - first) / (double)(L > 1 ? L - 1 : 1));
		}
		int64_t lowest_residual = INT64_MAX;
		#pragma unroll
		for (unsigned k = 0; k < for_block_vectors; k++) {
			const unsigned v = k * for_block_threads + threadIdx.x;
			if (v < num_vectors) {
				#pragma unroll
				for (unsigned j = 0; j < E; j++) {
					const int64_t residual = (int64_t)(int_t<U>) held[k].template get<U>(j) - (first + c1 * (int64_t)(v * E + j));
					lowest_residual = residual < lowest_residual ? residual : lowest_residual;
				}
			}
		}
		lowest_residual = warp_min(lowest_residual);
		if (lane_index() == 0) { warp_lowest[parity][warp_index_in_block()] = lowest_residual; }
		__syncthreads();
		#pragma unroll
		for (unsigned w = 0; w < for_block_threads / warp_size; w++) {
			const int64_t other = warp_lowest[parity][w];
			lowest_residual = other < lowest_residual ? other : lowest_residual;
		}
		const int64_t c0 = first + (lowest_residual - lowest);
		if (threadIdx.x == 0) {
			segment_model_coefficients[s * dim] = (int32_t) c0;
			if (dim == 2) { segment_model_coefficients[s * dim + 1] = (int32_t) c1; }
		}
		#pragma unroll
		for (unsigned k = 0; k < for_block_vectors; k++) {
			const unsigned v = k * for_block_threads + threadIdx.x;
			if (v < num_vectors) {
				byte_pack<E * C> out;
				#pragma unroll
				for (unsigned j = 0; j < E; j++) {
					const int64_t offset = (int64_t)(int_t<U>) held[k].template get<U>(j) - (c0 + c1 * (int64_t)(v * E + j));
					if (offset < lowest || offset > highest) { overflow = true; }
					out.template set<C>(j, (uint_t<C>)(int_t<C>) offset);
				}
				store_pack<E * C>(compressed + s * L + (uint64_t) v * E, out);
			}
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

// placement[t] = index of tile t's first run; placement[num_tiles] = the number of runs; open_run_start[t] = start of the
// run which contains the element before tile t (undefined for tile 0).  Two scans over the tile summaries -- the run open at
// each tile's start is a max-scan of the last-head positions; with it the tile's head count is known, and the placement is
// the sum-scan of the counts -- each done in three steps over chunks of 1024 tiles (a block per chunk):
//   run_chunk_last_heads   chunk c: its last natural head                                       (reduce)
//   run_scan_chunks        one block: exclusive "last non-zero" scan of the chunks' values         (scan of <= a few thousand values)
//   run_tile_counts        chunk c: open_run_start and head count of every tile; the chunk's total (scan within the chunk + carry)
//   run_scan_chunks        one block: exclusive sum of the chunks' totals
//   run_tile_places        chunk c: placement of every tile                                      (scan within the chunk + carry)
constexpr unsigned placement_chunk_tiles = 1024;

template <typename T, bool LastNonZero>
__device__ __forceinline__ T block_exclusive_scan_1024(T value, T* shared /* [33] */, T& block_total)
{
	// LastNonZero: the nearest preceding non-zero value (0 if none); else the sum of the preceding values
	const unsigned lane = lane_index(), warp = warp_index_in_block();
	T inclusive = value;
	#pragma unroll
	for (unsigned d = 1; d < warp_size; d <<= 1) {
		const T y = shuffle_up(inclusive, d);
		if (lane >= d) { inclusive = LastNonZero ? (inclusive != 0 ? inclusive : y) : (T)(inclusive + y); }
	}
	if (lane == warp_size - 1) { shared[warp] = inclusive; }
	__syncthreads();
	if (warp == 0) {
		T w = (lane < warps_per_block()) ? shared[lane] : (T) 0;
		T scanned = w;
		#pragma unroll
		for (unsigned d = 1; d < warp_size; d <<= 1) {
			const T y = shuffle_up(scanned, d);
			if (lane >= d) { scanned = LastNonZero ? (scanned != 0 ? scanned : y) : (T)(scanned + y); }
		}
		// exclusive over warps
		T before = shuffle_up(scanned, 1);
		if (lane == 0) { before = 0; }
		shared[lane] = before;
		if (lane == warp_size - 1) { shared[warp_size] = scanned; }
	}
	__syncthreads();
	T before_lane = shuffle_up(inclusive, 1);
	if (lane == 0) { before_lane = 0; }
	const T carried = shared[warp];
	const T result = LastNonZero ? (before_lane != 0 ? before_lane : carried) : (T)(carried + before_lane);
	block_total = shared[warp_size];
	__syncthreads();
	return result;
}

__global__ void __launch_bounds__(placement_chunk_tiles) run_chunk_last_heads(
	uint64_t* __restrict__ chunk_last, const run_tile_summary_t* __restrict__ summaries, uint64_t num_tiles)
{
	__shared__ uint64_t shared[warp_size + 1];
	const uint64_t tile = (uint64_t) blockIdx.x * placement_chunk_tiles + threadIdx.x;
	const uint64_t last = (tile < num_tiles) ? summaries[tile].last_head : no_position;
	uint64_t total;
	block_exclusive_scan_1024<uint64_t, true>(last == no_position ? 0 : last + 1, shared, total);
	if (threadIdx.x == 0) { chunk_last[blockIdx.x] = total; } // position + 1 of the chunk's last natural head, 0 if it has none
}

template <bool LastNonZero>
__global__ void __launch_bounds__(1024) run_scan_chunks(uint64_t* __restrict__ values, uint64_t count, uint64_t* __restrict__ total_out)
{
	// in place: values[c] <- exclusive scan; sequential over batches of 1024 chunks with a running carry
	__shared__ uint64_t shared[warp_size + 1];
	uint64_t carry = 0;
	for (uint64_t base = 0; base < count; base += 1024) {
		const uint64_t c = base + threadIdx.x;
		const uint64_t mine = (c < count) ? values[c] : 0;
		uint64_t total;
		const uint64_t before = block_exclusive_scan_1024<uint64_t, LastNonZero>(mine, shared, total);
		if (c < count) { values[c] = LastNonZero ? (before != 0 ? before : carry) : carry + before; }
		carry = LastNonZero ? (total != 0 ? total : carry) : carry + total;
	}
	if (threadIdx.x == 0 && total_out != nullptr) { *total_out = carry; }
}

__global__ void __launch_bounds__(placement_chunk_tiles) run_tile_counts(
	uint64_t*                  __restrict__  counts,          // per tile: its number of heads
	uint64_t*                  __restrict__  open_run_start,
	uint64_t*                  __restrict__  chunk_totals,
	const uint64_t*            __restrict__  chunk_open,      // exclusive last-non-zero scan of chunk_last
	const run_tile_summary_t*  __restrict__  summaries,
	uint64_t                                 num_tiles,
	uint64_t                                 length,
	uint64_t                                 max_run_length)
{
	__shared__ uint64_t shared[warp_size + 1];
	const uint64_t tile = (uint64_t) blockIdx.x * placement_chunk_tiles + threadIdx.x;
	const bool valid = tile < num_tiles;
	run_tile_summary_t s = { no_position, no_position, 0 };
	if (valid) { s = summaries[tile]; }
	uint64_t unused;
	uint64_t open = block_exclusive_scan_1024<uint64_t, true>(s.last_head == no_position ? 0 : s.last_head + 1, shared, unused);
	if (open == 0) { open = chunk_open[blockIdx.x]; }
	const uint64_t origin = (open == 0) ? 0 : open - 1;
	uint64_t count = 0;
	if (valid) {
		const uint64_t tile_start = tile * run_tile_elements;
		const uint64_t tile_end = minimum<uint64_t>(tile_start + run_tile_elements, length);
		const uint64_t prefix_end = (s.first_head == no_position) ? tile_end : s.first_head;
		count = s.heads_from_first + ((tile == 0) ? 0 : split_points_between(origin, tile_start, prefix_end, max_run_length));
		counts[tile] = count;
		open_run_start[tile] = origin;
	}
	uint64_t total;
	block_exclusive_scan_1024<uint64_t, false>(count, shared, total);
	if (threadIdx.x == 0) { chunk_totals[blockIdx.x] = total; }
}

__global__ void __launch_bounds__(placement_chunk_tiles) run_tile_places(
	uint64_t* __restrict__ placement /* in: counts, out: exclusive sums; [num_tiles] = total */, const uint64_t* __restrict__ chunk_bases,
	const uint64_t* __restrict__ grand_total, uint64_t num_tiles)
{
	__shared__ uint64_t shared[warp_size + 1];
	const uint64_t tile = (uint64_t) blockIdx.x * placement_chunk_tiles + threadIdx.x;
	const uint64_t count = (tile < num_tiles) ? placement[tile] : 0;
	uint64_t unused;
	const uint64_t before = block_exclusive_scan_1024<uint64_t, false>(count, shared, unused);
	if (tile < num_tiles) { placement[tile] = chunk_bases[blockIdx.x] + before; }
	if (blockIdx.x == 0 && threadIdx.x == 0) { placement[num_tiles] = *grand_total; }
}

// Exclusive prefix sums of an array of counts, in place, in three steps over chunks of 1024 (values[count] = the total):
// chunk_exclusive_sums -> run_scan_chunks<false> on the chunk totals -> add_chunk_bases.
__global__ void __launch_bounds__(placement_chunk_tiles) chunk_exclusive_sums(uint64_t* __restrict__ values, uint64_t* __restrict__ chunk_totals, uint64_t count)
{
	__shared__ uint64_t shared[warp_size + 1];
	const uint64_t i = (uint64_t) blockIdx.x * placement_chunk_tiles + threadIdx.x;
	const uint64_t mine = (i < count) ? values[i] : 0;
	uint64_t total;
	const uint64_t before = block_exclusive_scan_1024<uint64_t, false>(mine, shared, total);
	if (i < count) { values[i] = before; }
	if (threadIdx.x == 0) { chunk_totals[blockIdx.x] = total; }
}

__global__ void __launch_bounds__(placement_chunk_tiles) add_chunk_bases(uint64_t* __restrict__ values, const uint64_t* __restrict__ chunk_bases,
	const uint64_t* __restrict__ grand_total, uint64_t count)
{
	const uint64_t i = (uint64_t) blockIdx.x * placement_chunk_tiles + threadIdx.x;
	if (i < count) { values[i] += chunk_bases[blockIdx.x]; }
	if (blockIdx.x == 0 && threadIdx.x == 0) { values[count] = *grand_total; }
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

// ---- the fast path of the run encoders: run lengths limited to at least a tile's length (or not at all) ------------------
// Then a run which begins in a tile has no split point inside that tile, and the run which is open at a tile's start has at most
// one: heads are the natural ones plus, per tile, at most one computed position.  A thread takes 16 CONSECUTIVE elements with
// 16-byte loads, turns them into a 16-bit mask of heads (the element before its first comes from the previous lane by shuffle),
// and everything else is bit arithmetic on masks.  run_emit_fast compacts the tile's (value, position) pairs through shared
// memory so that they leave as contiguous, coalesced stores.

template <unsigned U>
__device__ __forceinline__ void load_16_elements(uint_t<U> (&v)[run_tile_per_thread], const uint_t<U>* __restrict__ x, uint64_t first, uint64_t limit, bool vector_loads)
{
	static_assert(run_tile_per_thread == 16, "a thread's elements are loaded as 16 x U bytes");
	if (vector_loads && first + run_tile_per_thread <= limit) {
		// (through L1: a thread's 16-byte pieces are consecutive, so two of them share a 32-byte sector -- and a warp's load
		// instruction touches every other piece of the sectors it fetches -- which a no-allocate load would fetch twice)
		byte_pack<16 * U> pack;
		#pragma unroll
		for (unsigned k = 0; k < U; k++) {
			const uint4 t = __ldg(reinterpret_cast<const uint4*>(x + first) + k);
			pack.u32[4 * k] = t.x; pack.u32[4 * k + 1] = t.y; pack.u32[4 * k + 2] = t.z; pack.u32[4 * k + 3] = t.w;
		}
		#pragma unroll
		for (unsigned j = 0; j < run_tile_per_thread; j++) { v[j] = pack.template get<U>(j); }
	}
	else {
		#pragma unroll
		for (unsigned j = 0; j < run_tile_per_thread; j++) { v[j] = (first + j < limit) ? x[first + j] : (uint_t<U>) 0; }
	}
}

// bit j set: element first + j begins a natural run (elements at or beyond `limit` never do)
template <unsigned U>
__device__ __forceinline__ unsigned natural_heads_of_16(const uint_t<U> (&v)[run_tile_per_thread], const uint_t<U>* __restrict__ x, uint64_t first, uint64_t limit)
{
	// the element before this thread's first: the previous lane's last, except for a warp's first lane
	uint_t<U> before = shuffle_up(v[run_tile_per_thread - 1], 1);
	if (lane_index() == 0 && first != 0 && first < limit) { before = x[first - 1]; }
	unsigned mask = 0;
	#pragma unroll
	for (unsigned j = 0; j < run_tile_per_thread; j++) {
		const bool head = (j == 0) ? (first == 0 || v[0] != before) : (v[j] != v[j - 1]);
		if (head && first + j < limit) { mask |= 1u << j; }
	}
	return mask;
}

template <unsigned U>
__global__ void __launch_bounds__(run_tile_threads) run_tile_summaries_fast(
	run_tile_summary_t*  __restrict__  summaries,
	const uint_t<U>*     __restrict__  x,
	uint64_t                           length)
{
	__shared__ unsigned shared_first[2][run_tile_threads / warp_size], shared_last[2][run_tile_threads / warp_size], shared_count[2][run_tile_threads / warp_size];
	const bool vector_loads = is_aligned_to(x, 16);
	const uint64_t num_tiles = (length + run_tile_elements - 1) / run_tile_elements;
	unsigned iteration = 0;
	for (uint64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, iteration++) {
		const uint64_t tile_start = tile * run_tile_elements;
		const uint64_t first = tile_start + (uint64_t) threadIdx.x * run_tile_per_thread;
		uint_t<U> v[run_tile_per_thread];
		load_16_elements<U>(v, x, first, length, vector_loads);
		const unsigned mask = natural_heads_of_16<U>(v, x, first, length);
		// warp: first head (min), last head + 1 (max), head count (sum) -- tile-relative, 32-bit: one REDUX each
		const unsigned thread_first = threadIdx.x * run_tile_per_thread;
		unsigned lowest = mask ? thread_first + (__ffs(mask) - 1) : ~0u;
		unsigned highest = mask ? thread_first + (32 - __clz(mask)) : 0u; // position + 1
		unsigned count = __popc(mask);
		lowest = __reduce_min_sync(full_mask, lowest);
		highest = __reduce_max_sync(full_mask, highest);
		count = __reduce_add_sync(full_mask, count);
		const unsigned parity = iteration & 1u; // (double-buffered: one barrier per tile)
		if (lane_index() == 0) { shared_first[parity][warp_index_in_block()] = lowest; shared_last[parity][warp_index_in_block()] = highest; shared_count[parity][warp_index_in_block()] = count; }
		__syncthreads();
		if (threadIdx.x < warp_size) {
			const bool has = threadIdx.x < run_tile_threads / warp_size;
			lowest = __reduce_min_sync(full_mask, has ? shared_first[parity][threadIdx.x] : ~0u);
			highest = __reduce_max_sync(full_mask, has ? shared_last[parity][threadIdx.x] : 0u);
			count = __reduce_add_sync(full_mask, has ? shared_count[parity][threadIdx.x] : 0u);
			if (threadIdx.x == 0) {
				summaries[tile] = { lowest == ~0u ? no_position : tile_start + lowest, highest == 0 ? no_position : tile_start + highest - 1, count };
			}
		}
	}
}

template <unsigned U, typename Start>
__global__ void __launch_bounds__(run_tile_threads) run_emit_fast(
	uint_t<U>*        __restrict__  run_data,
	Start*            __restrict__  starts,
	const uint_t<U>*  __restrict__  x,
	const uint64_t*   __restrict__  placement,
	const uint64_t*   __restrict__  open_run_start,
	const run_tile_summary_t*  __restrict__  summaries,
	uint64_t                        length,
	uint64_t                        max_run_length,   // 0, or at least run_tile_elements
	uint64_t                        run_capacity,
	unsigned*                       status)
{
	__shared__ unsigned   warp_totals[run_tile_threads / warp_size + 1];
	__shared__ uint_t<U>  staged_values[run_tile_elements];
	__shared__ uint16_t   staged_positions[run_tile_elements];
	const bool vector_loads = is_aligned_to(x, 16);
	const uint64_t num_tiles = (length + run_tile_elements - 1) / run_tile_elements;
	for (uint64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
		const uint64_t tile_start = tile * run_tile_elements;
		const uint64_t first = tile_start + (uint64_t) threadIdx.x * run_tile_per_thread;
		uint_t<U> v[run_tile_per_thread];
		load_16_elements<U>(v, x, first, length, vector_loads);
		// (the tile's metadata are requested along with its elements: one round trip to memory per tile, not three)
		const uint64_t base = __ldg(placement + tile);
		const bool may_split = max_run_length != 0 && tile != 0;
		const uint64_t origin = may_split ? __ldg(open_run_start + tile) : 0;
		const uint64_t first_head = may_split ? __ldg(&summaries[tile].first_head) : no_position;
		unsigned mask = natural_heads_of_16<U>(v, x, first, length);
		// the one split point this tile can hold: in the part of the open run which precedes the tile's first natural head
		if (may_split) {
			// the first split point of the open run at or after the tile's start: origin + k * M for the smallest k >= 1 which gets
			// there (runs shorter than M, the usual case, need no 64-bit division -- a hundred instructions per thread and tile)
			const uint64_t distance = tile_start - origin;
			const uint64_t split = origin + (distance <= max_run_length ? max_run_length : (distance + max_run_length - 1) / max_run_length * max_run_length);
			const uint64_t prefix_end = (first_head == no_position) ? minimum<uint64_t>(tile_start + run_tile_elements, length) : first_head;
			if (split < prefix_end && split >= first && split < first + run_tile_per_thread) { mask |= 1u << (unsigned)(split - first); }
		}
		// exclusive scan of the head counts over the block
		const unsigned count = __popc(mask);
		const unsigned inclusive = warp_inclusive_sum(count);
		if (lane_index() == warp_size - 1) { warp_totals[warp_index_in_block()] = inclusive; }
		__syncthreads();
		unsigned slot = inclusive - count, tile_heads = 0;
		#pragma unroll
		for (unsigned w = 0; w < run_tile_threads / warp_size; w++) { // (every thread sums the warps' totals itself: no second barrier)
			const unsigned t = warp_totals[w];
			if (w < warp_index_in_block()) { slot += t; }
			tile_heads += t;
		}
		// (heads are sparse -- one in `mean run length` elements -- so the loop is over the mask's set bits, and a head's value is
		// read again, out of L1, rather than selected from 16 registers)
		for (unsigned heads = mask; heads != 0; heads &= heads - 1, slot++) {
			const unsigned j = __ffs((int) heads) - 1;
			staged_values[slot] = __ldg(x + first + j);
			staged_positions[slot] = (uint16_t)(threadIdx.x * run_tile_per_thread + j);
		}
		__syncthreads();
		for (unsigned e = threadIdx.x; e < tile_heads; e += run_tile_threads) {
			if (base + e < run_capacity) { run_data[base + e] = staged_values[e]; starts[base + e] = (Start)(tile_start + staged_positions[e]); }
			else { raise(status, status_capacity); }
		}
		__syncthreads();
	}
}

// ---- the fast path in ONE pass over the column: tiles with two chained decoupled look-backs ------------------------------------
// What the two passes compute around their scans, a tile's CTA obtains from its predecessors' descriptors (kernels/delta.cuh:
// self-validating 16-byte descriptors, relaxed stores and loads, no fences):
//   A  the position of the last run head at or before the tile's end -- "inclusive" at once for a tile which holds a head
//      (nearly all do), so the run open at a tile's start is one look at the previous tile's descriptor;
//   B  the number of run heads (natural ones + the at most one split point, known once A's look-back is) -- an ordinary
//      sum look-back gives the index of the tile's first run.
// Tiles are drawn from a ticket counter, in order: a tile's predecessors are running or done whatever the grid.
// workspace: descriptors A [num_super_tiles], descriptors B [num_super_tiles], the run count (8 bytes), the ticket counter (8
// bytes), zeroed.
__device__ __forceinline__ uint64_t look_back_to_nearest(const delta_detail::tile_descriptor_t* descriptors, uint64_t tile)
{
	// called by one whole warp for tile >= 1: the value of the nearest predecessor whose descriptor is inclusive, all nearer
	// ones having been seen to hold none (tile 0 is always inclusive: position 0 begins a run)
	const unsigned lane = lane_index();
	for (uint64_t looked = 0; ; looked += warp_size) {
		const bool valid = looked + lane < tile;
		uint64_t value = 0;
		uint32_t status = delta_detail::tile_unknown;
		while (true) {
			if (valid && status == delta_detail::tile_unknown) { status = delta_detail::peek<8>(descriptors + (tile - 1 - (looked + lane)), value); }
			const unsigned unseen = __ballot_sync(full_mask, valid && status == delta_detail::tile_unknown);
			const unsigned inclusive = __ballot_sync(full_mask, valid && status == delta_detail::tile_has_inclusive);
			const unsigned seen_prefix = unseen ? ((1u << (__ffs((int) unseen) - 1)) - 1u) : full_mask; // lanes before the first unseen one
			const unsigned usable = inclusive & seen_prefix;
			if (usable) { return shuffle_from(value, (unsigned) __ffs((int) usable) - 1); }
			if (!unseen) { break; } // none of these 32 holds a head: farther back
		}
	}
}

// A look-back covers a SUPER-TILE of run_super_tiles consecutive tiles -- small tiles mean as many descriptors in flight as
// there are CTAs, and look-backs over a thousand predecessors (measured: slower than the two passes) -- whose run heads are
// compacted into shared memory tile by tile while the CTA reads it; after the look-backs they leave as one contiguous,
// coalesced block.  The split point, if there is one, lies before the super-tile's first natural head: entry 0 of the block
// is kept for it.  (A super-tile with more heads than the shared arrays hold -- runs of fewer than 4-8 elements on
// average -- is read again, tile by tile, and written straight from registers.)
constexpr unsigned run_super_tiles = 8;
constexpr unsigned run_super_elements = run_super_tiles * run_tile_elements; // 32 768: positions within fit 16 bits
template <unsigned U> struct run_staging { static constexpr unsigned capacity = (U >= 8) ? 4096 : (U == 4 ? 6144 : 8192); };

template <unsigned U, typename Start>
__global__ void __launch_bounds__(run_tile_threads) run_encode_single_pass(
	uint_t<U>*        __restrict__  run_data,
	Start*            __restrict__  starts,
	const uint_t<U>*  __restrict__  x,
	void*                           workspace,
	uint64_t                        length,
	uint64_t                        max_run_length,   // 0, or at least run_super_elements
	uint64_t                        run_capacity,
	unsigned*                       status)
{
	using delta_detail::tile_descriptor_t;
	constexpr unsigned warps = run_tile_threads / warp_size;
	constexpr unsigned capacity = run_staging<U>::capacity;
	__shared__ unsigned   shared_first[warps], shared_last[warps];
	__shared__ unsigned   warp_totals[2][warps];
	__shared__ uint64_t   shared_ticket, shared_base, shared_split;
	__shared__ uint_t<U>  staged_values[capacity];
	__shared__ uint16_t   staged_positions[capacity];
	const uint64_t num_super_tiles = (length + run_super_elements - 1) / run_super_elements;
	tile_descriptor_t* const last_heads = static_cast<tile_descriptor_t*>(workspace);   // A
	tile_descriptor_t* const head_counts = last_heads + num_super_tiles;                 // B
	uint64_t* const total_out = reinterpret_cast<uint64_t*>(head_counts + num_super_tiles);
	unsigned long long* const ticket_counter = reinterpret_cast<unsigned long long*>(total_out + 1);
	const bool vector_loads = is_aligned_to(x, 16);
	const unsigned lane = lane_index(), warp = warp_index_in_block();
	while (true) {
		if (threadIdx.x == 0) { shared_ticket = atomicAdd(ticket_counter, 1ull); }
		__syncthreads();
		const uint64_t tile = shared_ticket;
		if (tile >= num_super_tiles) { return; }
		const uint64_t tile_start = tile * run_super_elements;
		// 1. the super-tile's natural heads, tile by tile: counted, and compacted into shared memory from entry 1 on
		unsigned natural = 0;                 // natural heads so far (the same in every thread)
		unsigned lowest = ~0u, highest = 0u;  // this thread's first head, and last head + 1, relative to the super-tile
		#pragma unroll 1
		for (unsigned k = 0; k < run_super_tiles; k++) {
			const uint64_t first = tile_start + (uint64_t) k * run_tile_elements + (uint64_t) threadIdx.x * run_tile_per_thread;
			if (first - (uint64_t) threadIdx.x * run_tile_per_thread >= length) { break; }
			uint_t<U> v[run_tile_per_thread];
			load_16_elements<U>(v, x, first, length, vector_loads);
			const unsigned mask = natural_heads_of_16<U>(v, x, first, length);
			const unsigned relative = k * run_tile_elements + threadIdx.x * run_tile_per_thread;
			if (mask) { if (lowest == ~0u) { lowest = relative + (__ffs((int) mask) - 1); } highest = relative + (32 - __clz((int) mask)); }
			const unsigned count = __popc(mask);
			const unsigned inclusive = warp_inclusive_sum(count);
			if (lane == warp_size - 1) { warp_totals[k & 1u][warp] = inclusive; }
			__syncthreads();
			unsigned slot = 1 + natural + inclusive - count, tile_heads = 0;
			#pragma unroll
			for (unsigned w = 0; w < warps; w++) {
				const unsigned t = warp_totals[k & 1u][w];
				if (w < warp) { slot += t; }
				tile_heads += t;
			}
			natural += tile_heads;
			if (1 + natural <= capacity) {
				for (unsigned heads = mask; heads != 0; heads &= heads - 1, slot++) {
					const unsigned j = __ffs((int) heads) - 1;
					staged_values[slot] = __ldg(x + first + j);
					staged_positions[slot] = (uint16_t)(relative + j);
				}
			}
		}
		lowest = __reduce_min_sync(full_mask, lowest);
		highest = __reduce_max_sync(full_mask, highest);
		if (lane == 0) { shared_first[warp] = lowest; shared_last[warp] = highest; }
		__syncthreads();
		// 2. warp 0: the run open at the super-tile's start (hence its split point) and the index of its first run
		if (warp == 0) {
			const bool has = lane < warps;
			lowest = __reduce_min_sync(full_mask, has ? shared_first[lane] : ~0u);
			highest = __reduce_max_sync(full_mask, has ? shared_last[lane] : 0u);
			const bool has_head = highest != 0;
			// A: a super-tile with a head knows its last head at once
			if (lane == 0) {
				if (has_head) { delta_detail::publish<8>(last_heads + tile, delta_detail::tile_has_inclusive, tile_start + highest - 1); }
				else { delta_detail::publish<8>(last_heads + tile, delta_detail::tile_has_aggregate, 0ull); }
			}
			uint64_t split = no_position;
			if (tile != 0 && (max_run_length != 0 || !has_head)) {
				const uint64_t origin = look_back_to_nearest(last_heads, tile); // where the run open at the super-tile's start began
				if (!has_head && lane == 0) { delta_detail::publish<8>(last_heads + tile, delta_detail::tile_has_inclusive, origin); }
				if (max_run_length != 0) {
					// the one split point a super-tile can hold: in the part of the open run which precedes its first natural head
					const uint64_t distance = tile_start - origin;
					const uint64_t candidate = origin + (distance <= max_run_length ? max_run_length : (distance + max_run_length - 1) / max_run_length * max_run_length);
					const uint64_t prefix_end = has_head ? tile_start + lowest : minimum<uint64_t>(tile_start + run_super_elements, length);
					if (candidate < prefix_end) { split = candidate; }
				}
			}
			// B: the super-tile's head count, then the index of its first run
			const uint64_t tile_heads = (uint64_t) natural + (split != no_position ? 1u : 0u);
			uint64_t base = 0;
			if (tile != 0) {
				if (lane == 0) { delta_detail::publish<8>(head_counts + tile, delta_detail::tile_has_aggregate, tile_heads); }
				base = delta_detail::look_back<8>(head_counts, tile, tile);
			}
			if (lane == 0) {
				delta_detail::publish<8>(head_counts + tile, delta_detail::tile_has_inclusive, base + tile_heads);
				shared_base = base; shared_split = split;
				if (tile + 1 == num_super_tiles) { *total_out = base + tile_heads; }
				if (split != no_position) { staged_values[0] = __ldg(x + split); staged_positions[0] = (uint16_t)(split - tile_start); }
			}
		}
		__syncthreads();
		// 3. the block of (value, start) pairs
		const uint64_t split = shared_split;
		const unsigned skipped = (split != no_position) ? 0u : 1u;          // entry 0 is the split point's
		const uint64_t base = shared_base - skipped;                          // run index of entry 0
		const unsigned past_last = 1 + natural;
		if (past_last <= capacity) {
			for (unsigned e = skipped + threadIdx.x; e < past_last; e += run_tile_threads) {
				if (base + e < run_capacity) { run_data[base + e] = staged_values[e]; starts[base + e] = (Start)(tile_start + staged_positions[e]); }
				else { raise(status, status_capacity); }
			}
		}
		else { // more heads than the shared arrays hold: the tiles again, written from registers
			uint64_t next = shared_base; // run index of the next head
			if (split != no_position) {
				if (threadIdx.x == 0) { if (next < run_capacity) { run_data[next] = __ldg(x + split); starts[next] = (Start) split; } else { raise(status, status_capacity); } }
				next++;
			}
			#pragma unroll 1
			for (unsigned k = 0; k < run_super_tiles; k++) {
				const uint64_t first = tile_start + (uint64_t) k * run_tile_elements + (uint64_t) threadIdx.x * run_tile_per_thread;
				if (first - (uint64_t) threadIdx.x * run_tile_per_thread >= length) { break; }
				uint_t<U> v[run_tile_per_thread];
				load_16_elements<U>(v, x, first, length, vector_loads);
				const unsigned mask = natural_heads_of_16<U>(v, x, first, length);
				const unsigned count = __popc(mask);
				const unsigned inclusive = warp_inclusive_sum(count);
				__syncthreads(); // (the totals of the tile before have been read)
				if (lane == warp_size - 1) { warp_totals[0][warp] = inclusive; }
				__syncthreads();
				uint64_t at = next + inclusive - count;
				unsigned tile_heads = 0;
				#pragma unroll
				for (unsigned w = 0; w < warps; w++) {
					const unsigned t = warp_totals[0][w];
					if (w < warp) { at += t; }
					tile_heads += t;
				}
				for (unsigned heads = mask; heads != 0; heads &= heads - 1, at++) {
					const unsigned j = __ffs((int) heads) - 1;
					if (at < run_capacity) { run_data[at] = __ldg(x + first + j); starts[at] = (Start)(first + j); } else { raise(status, status_capacity); }
				}
				next += tile_heads;
			}
		}
		__syncthreads(); // before the next super-tile reuses the shared arrays (and the ticket)
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
//   pass 1  variable_width_sizes           a warp per segment, 16-byte vectors of elements: the bytes the segment takes and
//                                          -- where containers straddle neither a lane's vector nor segments -- the size
//                                          containers, each OR-ed together from the fields of neighbouring lanes
//           variable_width_pack_sizes      (otherwise) a thread per size container
//   scan    chunk_exclusive_sums etc.      -> byte offset of every segment = position_anchors
//   pass 2  variable_width_emit            a warp per segment: sizes again, a warp scan per 512 bytes of elements, the bytes
//                                          compacted into a shared-memory window laid out in the phase of the destination
//                                          (window byte p <-> global byte base + p, base 16-byte aligned), whole 16-byte
//                                          chunks of which are flushed with vector stores every `emit_window_steps` steps
// -------------------------------------------------------------------------------------------------
template <unsigned U>
__device__ __forceinline__ unsigned significant_bytes(uint_t<U> x)
{
	if (U == 8) { return x == 0 ? 0u : (unsigned)((64 - __clzll((long long) x) + 7) / 8); }
	return x == 0 ? 0u : (unsigned)((32 - __clz((int)(uint32_t) x) + 7) / 8);
}

// a lane's 16 / U elements i ... of a segment of `count`, zeroes beyond its end
template <unsigned U>
__device__ __forceinline__ byte_pack<16> load_segment_vector(const uint_t<U>* __restrict__ segment, uint64_t i, uint64_t count, bool aligned)
{
	constexpr unsigned E = 16 / U;
	if (aligned && i + E <= count) { return load_pack<16>(segment + i); }
	byte_pack<16> v;
	#pragma unroll
	for (unsigned j = 0; j < E; j++) { v.template set<U>(j, (i + j < count) ? segment[i + j] : (uint_t<U>) 0); }
	return v;
}

constexpr unsigned sizes_steps = 4; // 16-byte loads a lane has in flight

// The loads of one iteration of a segment's warp -- `Steps` steps of 32 lanes x 16 bytes from element `first` of the segment
// on -- all issued before any is used.  Whole: every lane's vectors lie inside the segment and are 16-byte aligned.
template <unsigned U, unsigned Steps, bool Whole>
__device__ __forceinline__ void load_steps(byte_pack<16> (&v)[Steps], const uint_t<U>* __restrict__ segment, uint64_t first, uint64_t count, bool aligned)
{
	constexpr unsigned E = 16 / U;
	#pragma unroll
	for (unsigned k = 0; k < Steps; k++) {
		const uint64_t i = first + (uint64_t) k * warp_size * E + lane_index() * E;
		if (Whole) { v[k] = load_pack<16>(segment + i); }
		else if (i < count) { v[k] = load_segment_vector<U>(segment, i, count, aligned); }
	}
}

// packed_element_sizes: null unless the launcher has established that x is 16-byte aligned, that the segment length is a
// multiple both of the elements in a 16-byte vector and of those in a container, and that a container holds the elements of a
// power-of-two number of vectors
template <unsigned U, unsigned CS>
__global__ void variable_width_sizes(
	uint64_t*         __restrict__  segment_bytes,
	uint_t<CS>*       __restrict__  packed_element_sizes,
	const uint_t<U>*  __restrict__  x,
	uint64_t                        length,
	uint64_t                        segment_length,
	unsigned                        min_size,
	unsigned                        bits,
	unsigned*                       status)
{
	constexpr unsigned E = 16 / U;
	constexpr unsigned span = sizes_steps * warp_size * E; // elements per iteration
	using container_t = uint_t<CS>;
	const uint64_t num_segments = (length + segment_length - 1) / segment_length;
	const bool aligned = is_aligned_to(x, 16) && segment_length % E == 0;
	const unsigned lane = lane_index();
	const unsigned per_container = (8 * CS) / bits;
	const unsigned group = per_container / E;                      // lanes whose vectors share a container (a power of two)
	const unsigned log2_per_container = 31 - __clz((int) per_container);
	const unsigned field_shift = (lane & (group - 1)) * E * bits;
	const bool writes_container = (lane & (group - 1)) == 0;
	const unsigned field_mask = (bits >= 32) ? ~0u : ((1u << bits) - 1);
	unsigned overflow = 0; // bits of size - min_size beyond the field
	for (uint64_t s = global_warp_index(); s < num_segments; s += num_warps_in_grid()) {
		const uint64_t start = s * segment_length;
		const uint64_t count = minimum<uint64_t>(segment_length, length - start);
		const uint_t<U>* const segment = x + start;
		uint64_t bytes = 0;
		auto iteration = [&](uint64_t first, auto whole) {
			constexpr bool Whole = decltype(whole)::value;
			byte_pack<16> v[sizes_steps];
			load_steps<U, sizes_steps, Whole>(v, segment, first, count, aligned);
			unsigned iteration_bytes = 0;
			#pragma unroll
			for (unsigned k = 0; k < sizes_steps; k++) {
				const uint64_t i = first + (uint64_t) k * warp_size * E + lane * E;
				if (!Whole && i - lane * E >= count) { break; }
				container_t fields = 0;
				#pragma unroll
				for (unsigned j = 0; j < E; j++) {
					if (Whole || i + j < count) {
						const unsigned size = max(significant_bytes<U>(v[k].template get<U>(j)), min_size);
						const unsigned excess = size - min_size;
						if (bits < 32) { overflow |= excess >> bits; }
						iteration_bytes += size;
						fields |= (container_t)(excess & field_mask) << (j * bits);
					}
				}
				if (packed_element_sizes != nullptr) {
					container_t container = fields << field_shift;
					for (unsigned d = 1; d < group; d <<= 1) { container |= __shfl_xor_sync(full_mask, container, d); }
					if (writes_container && (Whole || i < count)) { packed_element_sizes[(start + i) >> log2_per_container] = container; }
				}
			}
			bytes += iteration_bytes;
		};
		const uint64_t whole_end = aligned ? count / span * span : 0;
		for (uint64_t first = 0; first < whole_end; first += span) { iteration(first, ::std::true_type{}); }
		for (uint64_t first = whole_end; first < count; first += span) { iteration(first, ::std::false_type{}); }
		bytes = warp_inclusive_sum(bytes);
		if (lane == warp_size - 1) { segment_bytes[s] = bytes; }
	}
	if (overflow != 0) { raise(status, status_unrepresentable); }
}

constexpr unsigned emit_window_steps = 4;                              // steps of 32 lanes x 16 bytes of elements between flushes
constexpr unsigned emit_window_bytes = emit_window_steps * 512 + 32;   // + the destination's phase or the carried partial chunk

// dynamic shared memory: emit_window_bytes per warp of the block
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
	extern __shared__ __align__(16) unsigned char emit_windows[];
	constexpr unsigned E = 16 / U;
	constexpr unsigned span = emit_window_steps * warp_size * E; // elements per window
	unsigned char* const window = emit_windows + warp_index_in_block() * emit_window_bytes;
	const uint64_t num_segments = (length + segment_length - 1) / segment_length;
	const bool aligned = is_aligned_to(x, 16) && segment_length % E == 0;
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
		const uint_t<U>* const segment = x + start;
		const uint64_t offset = segment_offsets[s];
		if (lane == 0) {
			if (IndexSize < 8 && (offset >> (8 * IndexSize)) != 0) { index_overflow = true; }
			position_anchors[s] = (uint_t<IndexSize>) offset;
		}
		// window byte p <-> compressed_data[window_offset + p], that address 16-byte aligned at p = 0; the segment's first byte
		// is window byte `phase`, the bytes before it belong to the previous segment (another warp's)
		const unsigned phase = (unsigned)((reinterpret_cast<uintptr_t>(compressed_data) + offset) & 15u);
		int64_t window_offset = (int64_t) offset - (int64_t) phase;
		unsigned fill = phase;
		bool foreign_head = phase != 0;
		auto fill_window = [&](uint64_t first, auto whole) {
			constexpr bool Whole = decltype(whole)::value;
			byte_pack<16> v[emit_window_steps];
			load_steps<U, emit_window_steps, Whole>(v, segment, first, count, aligned);
			#pragma unroll
			for (unsigned k = 0; k < emit_window_steps; k++) {
				const uint64_t i = first + (uint64_t) k * warp_size * E + lane * E;
				if (!Whole && i - lane * E >= count) { break; }
				unsigned size[E], total = 0;
				#pragma unroll
				for (unsigned j = 0; j < E; j++) {
					size[j] = (Whole || i + j < count) ? max(significant_bytes<U>(v[k].template get<U>(j)), min_size) : 0u;
					total += size[j];
				}
				const unsigned inclusive = warp_inclusive_sum(total);
				unsigned at = fill + inclusive - total;
				#pragma unroll
				for (unsigned j = 0; j < E; j++) {
					const uint_t<U> value = v[k].template get<U>(j);
					#pragma unroll
					for (unsigned b = 0; b < U; b++) { if (b < size[j]) { window[at + b] = (unsigned char)(value >> (8 * b)); } }
					at += size[j];
				}
				fill += __shfl_sync(full_mask, inclusive, warp_size - 1);
			}
		};
		const uint64_t whole_end = aligned ? count / span * span : 0;
		for (uint64_t first = 0; first < count; first += span) {
			if (first < whole_end) { fill_window(first, ::std::true_type{}); } else { fill_window(first, ::std::false_type{}); }
			__syncwarp();
			// flush the whole 16-byte chunks; the partial one at the end is carried to the front of the window
			const unsigned whole_chunks = fill / 16;
			for (unsigned c = lane; c < whole_chunks; c += warp_size) {
				const int64_t position = window_offset + 16 * (int64_t) c;
				const unsigned begin = (c == 0 && foreign_head) ? phase : 0u;
				if (begin == 0 && (uint64_t) position + 16 <= data_capacity) {
					store_streaming_16(compressed_data + position, *reinterpret_cast<const uint4*>(window + 16 * c));
				}
				else {
					for (unsigned b = begin; b < 16; b++) { if ((uint64_t)(position + b) < data_capacity) { compressed_data[position + b] = window[16 * c + b]; } }
				}
			}
			const unsigned remaining = fill - 16 * whole_chunks;
			const unsigned char carried = (whole_chunks > 0 && lane < remaining) ? window[16 * whole_chunks + lane] : (unsigned char) 0;
			__syncwarp();
			if (whole_chunks > 0) {
				if (lane < remaining) { window[lane] = carried; }
				window_offset += 16 * (int64_t) whole_chunks;
				fill = remaining;
				foreign_head = false;
			}
			__syncwarp();
		}
		for (unsigned b = lane; b < fill; b += warp_size) {
			if (!(foreign_head && b < phase) && (uint64_t)(window_offset + b) < data_capacity) { compressed_data[window_offset + b] = window[b]; }
		}
		__syncwarp();
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
	const bool sixteen = per_container == run_tile_per_thread && is_aligned_to(x, 16); // 2-bit fields in 32-bit containers, 4-bit in 64-bit: one 16-element load
	for (uint64_t c = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; c < num_containers; c += (uint64_t) gridDim.x * blockDim.x) {
		uint64_t container = 0;
		if (sixteen) {
			uint_t<U> v[run_tile_per_thread];
			load_16_elements<U>(v, x, c * run_tile_per_thread, length, true);
			#pragma unroll
			for (unsigned k = 0; k < run_tile_per_thread; k++) {
				if (c * run_tile_per_thread + k < length) {
					container |= ((uint64_t)(max(significant_bytes<U>(v[k]), min_size) - min_size) & field_mask) << (k * bits);
				}
			}
		}
		else {
			for (uint64_t k = 0; k < per_container; k++) {
				const uint64_t i = c * per_container + k;
				if (i >= length) { break; }
				const unsigned size = max(significant_bytes<U>(x[i]), min_size);
				container |= ((uint64_t)(size - min_size) & field_mask) << (k * bits);
			}
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
