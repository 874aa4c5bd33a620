// Discard-Zero-Bytes, variable width (a.k.a. null suppression, variable): every element is
// stored in size_i of its low bytes, the sizes come from fixed-width bit fields packed into
// 32- or 64-bit containers, and every segment has the byte offset of its first element as
// an anchor.
//
// One warp decodes one anchored segment, in tiles of 1024 elements.  Two paths:
//
// * staged (the common shapes: 1, 2 or 4 bits per size field, segments which begin on a 32-bit
//   word of the size-field stream): per tile the warp
//     1. reads the tile's size fields with one coalesced access per lane (a lane owns 32
//        consecutive elements), sums them with population counts, and a single warp-shuffle scan
//        gives every lane its byte offset and the warp the tile's byte count;
//     2. stages the tile's bytes in shared memory with ONE TMA bulk copy (cp.async.bulk, SASS
//        UBLKCP) completing on a per-warp mbarrier -- the 16-byte-aligned body of the byte range;
//        the < 16 trailing bytes are fetched by four lanes, so nothing past the stream's last
//        word is ever read.  When the tile is a whole segment its byte range is known from the
//        next anchor and the copy is issued before the size fields have even arrived;
//     3. writes the byte offset of every 16-byte output vector (2 bytes each) and the size words
//        to shared memory, and
//     4. expands: 32 lanes x E elements per row (E = 16 / sizeof(Uncompressed)); a lane looks up
//        its vector's offset and size fields, assembles each element from the two or three
//        aligned shared-memory words holding its bytes with a funnel shift, and the warp stores
//        512 contiguous bytes.
//   The compressed bytes are read from HBM exactly once, no global access is narrower than
//   a sector, and there is no data-dependent loop.
// * generic (any field width, any alignment, or a launch without the shared memory): rows of
//   32 lanes x E elements; sizes straight from the containers, one warp scan per row, elements
//   assembled from aligned global words served by L1.
// (The reference walks a segment with one warp, 32 elements per iteration, memcpy'ing byte by
// byte through local memory: src/kernels/decompression/discard_zero_bytes_variable.cuh:196-213,276-310.)
#pragma once
#ifndef GIDDY_B200_KERNELS_DZB_VARIABLE_CUH_
#define GIDDY_B200_KERNELS_DZB_VARIABLE_CUH_

#include "common.cuh"

namespace giddy {
namespace kernels {

using namespace giddy::device;

#ifndef GIDDY_DZBV_MIN_BLOCKS
#define GIDDY_DZBV_MIN_BLOCKS 5
#endif

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

// ---- the staged path ----------------------------------------------------------------------------------

constexpr unsigned tile_elements = GIDDY_DZBV_TILE_ELEMENTS;
constexpr unsigned elements_per_lane = tile_elements / warp_size; // 32: a lane's size fields fill `bits` 32-bit words
static_assert(elements_per_lane == 32, "a lane owns 32 consecutive elements of a tile");
constexpr unsigned max_staged_bits = 4;

template <unsigned U> constexpr unsigned shared_memory_bytes_per_warp() { return GIDDY_DZBV_SHARED_BYTES_PER_WARP(U); }

template <unsigned U>
struct tile_storage {
	static constexpr unsigned E = 16 / U;
	static constexpr unsigned data_bytes = tile_elements * U + 64;
	static constexpr unsigned num_vectors = tile_elements / E;
	uint8_t*  data;     // [data_bytes]   the tile's bytes; the first one is at data[head]
	uint16_t* offsets;  // [num_vectors]  byte offset, within the tile, of each output vector's first element
	uint32_t* fields;   // [tile_elements * bits / 32] the tile's size fields, a flat little-endian bit stream
	uint64_t* barrier;

	__device__ __forceinline__ explicit tile_storage(unsigned char* warp_memory) :
		data(warp_memory),
		offsets(reinterpret_cast<uint16_t*>(warp_memory + data_bytes)),
		fields(reinterpret_cast<uint32_t*>(warp_memory + data_bytes + 2 * num_vectors)),
		barrier(reinterpret_cast<uint64_t*>(warp_memory + data_bytes + 2 * num_vectors + tile_elements * max_staged_bits / 8)) { }
};

// Sum of the `bits`-wide fields of x (bits in {1, 2, 4}).
template <unsigned bits>
__device__ __forceinline__ unsigned sum_of_fields(uint32_t x)
{
	if (bits == 1) { return __popc(x); }
	if (bits == 2) { return __popc(x & 0x55555555u) + 2 * __popc(x & 0xAAAAAAAAu); }
	return __popc(x & 0x11111111u) + 2 * __popc(x & 0x22222222u) + 4 * __popc(x & 0x44444444u) + 8 * __popc(x & 0x88888888u);
}

// x << s, with s clamped to 32 (so that s == 32 gives 0, as PTX's shl does)
__device__ __forceinline__ uint32_t shift_left_clamped(uint32_t x, unsigned s)
{
	uint32_t r;
	asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(s));
	return r;
}

// Sums of the fields within each byte of x, for fields of `bits` bits (bits in {1, 2, 4}).
template <unsigned bits>
__device__ __forceinline__ uint32_t field_sums_per_byte(uint32_t x)
{
	if (bits == 1) {
		x = x - ((x >> 1) & 0x55555555u);
		x = (x & 0x33333333u) + ((x >> 2) & 0x33333333u);
		return (x + (x >> 4)) & 0x0F0F0F0Fu;
	}
	if (bits == 2) {
		x = (x & 0x33333333u) + ((x >> 2) & 0x33333333u);
		return (x + (x >> 4)) & 0x0F0F0F0Fu;
	}
	return (x & 0x0F0F0F0Fu) + ((x >> 4) & 0x0F0F0F0Fu);
}

// One 16-byte output vector: E elements whose bytes lie back to back in shared memory from byte
// offset `at` of `data` (16-byte aligned), their size fields in the low E * bits bits of `chunk`.
// The (at most) 16 bytes are read as five aligned words and funnel-shifted into a 4-word window
// which begins at the first element's first byte; each element is the window's low bytes under a
// mask, after which the window is shifted right by the element's size.
template <unsigned U, unsigned bits>
__device__ __forceinline__ byte_pack<16> assemble_vector(const uint8_t* data, unsigned at, uint32_t chunk, unsigned min_size)
{
	constexpr unsigned E = 16 / U;
	constexpr uint32_t field_mask = (1u << bits) - 1u;
	const uint32_t* aligned = reinterpret_cast<const uint32_t*>(data + (at & ~3u));
	const uint32_t w0 = aligned[0], w1 = aligned[1], w2 = aligned[2], w3 = aligned[3], w4 = aligned[4];
	const unsigned shift = at * 8; // funnel shifts in wrap mode use the low 5 bits: (at % 4) * 8
	uint32_t window[5] = { __funnelshift_r(w0, w1, shift), __funnelshift_r(w1, w2, shift), __funnelshift_r(w2, w3, shift),
		__funnelshift_r(w3, w4, shift), 0u };
	byte_pack<16> vec;
	#pragma unroll
	for (unsigned j = 0; j < E; j++) {
		const unsigned size_in_bits = (((chunk >> (j * bits)) & field_mask) + min_size) * 8;
		// how many window words the elements after this one may still need
		const unsigned words_after = ((E - 1 - j) * U + 3) / 4; // (a compile-time constant once unrolled)
		if (U <= 4) {
			vec.template set<U>(j, (uint_t<U>)(window[0] & ~shift_left_clamped(0xffffffffu, size_in_bits)));
			#pragma unroll
			for (unsigned k = 0; k < 4; k++) {
				if (k < words_after) { window[k] = __funnelshift_rc(window[k], window[k + 1], size_in_bits); }
			}
		}
		else {
			const unsigned high_bits = (size_in_bits > 32 ? size_in_bits : 32u) - 32u;
			const uint32_t lo = window[0] & ~shift_left_clamped(0xffffffffu, size_in_bits);
			const uint32_t hi = window[1] & ~shift_left_clamped(0xffffffffu, high_bits);
			vec.template set<U>(j, (uint_t<U>)((uint64_t) lo | ((uint64_t) hi << 32)));
			const unsigned low_bits = size_in_bits - high_bits; // <= 32
			#pragma unroll
			for (unsigned k = 0; k < 4; k++) { // (one word more than needed: the second shift pulls it in)
				if (k < words_after + 1) { window[k] = __funnelshift_rc(window[k], window[k + 1], low_bits); }
			}
			#pragma unroll
			for (unsigned k = 0; k < 4; k++) {
				if (k < words_after) { window[k] = __funnelshift_rc(window[k], window[k + 1], high_bits); }
			}
		}
	}
	return vec;
}

// What becomes of the decoded vectors.  emit() is called by all lanes of the warp together, once per row of 32 x E
// elements, with `position` the index within the segment of this lane's first element, `vec` its E elements (zeros
// from index `valid` on) and `valid` in [0, E] the number of them which lie inside the segment.
template <unsigned U>
struct store_sink {
	static constexpr unsigned E = 16 / U;
	uint_t<U>* out; // the segment's first output element; 16-byte aligned

	__device__ __forceinline__ void emit(uint64_t position, const byte_pack<16>& vec, unsigned valid) const
	{
		if (valid == E) { store_pack_16(out + position, vec); }
		else {
			#pragma unroll
			for (unsigned j = 0; j < E; j++) { if (j < valid) { out[position + j] = vec.template get<U>(j); } }
		}
	}
};

// Decode one segment, tile by tile.  Requires: bits in {1, 2, 4}; (first_element * bits) % 32 == 0;
// packed sizes readable as 32-bit words up to num_field_words.
template <unsigned U, unsigned bits, typename Sink>
__device__ __forceinline__ void decode_segment_staged(
	Sink&                           sink,
	const uint8_t*    __restrict__  compressed_data,
	uint64_t                        segment_first_byte,
	bool                            next_anchor_known,
	uint64_t                        next_anchor,
	const uint32_t*   __restrict__  field_words,      // the packed size fields as a stream of 32-bit words
	uint64_t                        num_field_words,
	uint64_t                        first_element,    // global index of the segment's first element
	uint64_t                        length,
	unsigned                        min_size,
	const tile_storage<U>&          tile,
	uint32_t&                       barrier_phase)
{
	constexpr unsigned E = tile_storage<U>::E;
	constexpr unsigned vectors_per_lane = elements_per_lane / E;
	const unsigned lane = lane_index();
	constexpr uint32_t field_mask = (1u << bits) - 1u;
	uint64_t tile_first_byte = segment_first_byte;

	for (uint64_t tile_start = 0; tile_start < length; tile_start += tile_elements) {
		const unsigned tile_length = (unsigned) minimum<uint64_t>(tile_elements, length - tile_start);
		const bool bytes_known = next_anchor_known && (tile_start + tile_elements >= length);

		// the body of the byte range goes through TMA, its < 16 trailing bytes through four lanes
		auto stage_bytes = [&](uint64_t past_last_byte) {
			const uintptr_t first = reinterpret_cast<uintptr_t>(compressed_data) + tile_first_byte;
			const uintptr_t past_last = reinterpret_cast<uintptr_t>(compressed_data) + past_last_byte;
			const uintptr_t body_first = first & ~(uintptr_t) 15;
			const uintptr_t body_past_last = (past_last & ~(uintptr_t) 15) > body_first ? (past_last & ~(uintptr_t) 15) : body_first;
			// (a next anchor is a hint: whatever lies between it and the tile's last byte -- padding between segments, or a
			// corrupt anchor -- is never decoded, so never staged either: a tile's elements end within tile x U bytes of the
			// first one.  Two 32-bit guards on what goes into shared memory; clamping the 64-bit address instead, ahead of the
			// copy's issue, costs the kernel 3 %.)
			constexpr uint32_t most_body_bytes = tile_elements * U + 32;
			static_assert(most_body_bytes % 16 == 0 && most_body_bytes + 32 <= tile_storage<U>::data_bytes, "the clamped copy and its trailing words fit the data area");
			const uint32_t body_bytes = minimum<uint32_t>((uint32_t)(body_past_last - body_first), most_body_bytes);
			if (lane == 0) {
				fence_proxy_async_shared();
				if (body_bytes > 0) {
					mbarrier_expect_bytes(tile.barrier, body_bytes);
					tma_bulk_load(tile.data, reinterpret_cast<const void*>(body_first), body_bytes, tile.barrier);
				}
				else { mbarrier_arrive(tile.barrier); }
			}
			// trailing words: [body_past_last, round_up(past_last, 4)), at most 4 of them -- or, when the body is empty,
			// from the word holding the first byte
			const uintptr_t tail_first = (body_bytes > 0) ? body_past_last : (first & ~(uintptr_t) 3);
			const uintptr_t tail_word = tail_first + 4 * lane;
			if (lane < 8 && tail_word < past_last && (uint64_t)(tail_word - body_first) < most_body_bytes + 32 - 3) {
				*reinterpret_cast<uint32_t*>(tile.data + (tail_word - body_first)) = __ldg(reinterpret_cast<const uint32_t*>(tail_word));
			}
		};
		// 1. this lane's 32 size fields: `bits` consecutive words of the field stream (requested before the byte range's
		// anchors are consumed, so that the two round trips overlap)
		const unsigned lane_elements = (tile_length > lane * elements_per_lane) ?
			(unsigned) minimum<unsigned>(elements_per_lane, tile_length - lane * elements_per_lane) : 0u;
		const uint64_t lane_first_word = ((first_element + tile_start) * bits) / 32 + (uint64_t) lane * bits;
		uint32_t words[bits];
		#pragma unroll
		for (unsigned k = 0; k < bits; k++) {
			words[k] = (lane_first_word + k < num_field_words) ? load_streaming_4(field_words + lane_first_word + k) : 0u;
		}
		if (bytes_known) { stage_bytes(next_anchor); }
		// fields of elements past the tile's end do not count
		unsigned lane_bytes = lane_elements * min_size;
		{
			const unsigned valid_bits = lane_elements * bits;
			#pragma unroll
			for (unsigned k = 0; k < bits; k++) {
				const unsigned bits_here = (valid_bits > 32 * k) ? minimum<unsigned>(32u, valid_bits - 32 * k) : 0u;
				const uint32_t keep = (bits_here >= 32) ? 0xffffffffu : ((1u << bits_here) - 1u);
				words[k] &= keep;
				lane_bytes += sum_of_fields<bits>(words[k]);
			}
		}
		const unsigned inclusive = warp_inclusive_sum(lane_bytes);
		const unsigned tile_bytes = __shfl_sync(full_mask, inclusive, warp_size - 1);
		if (!bytes_known) { stage_bytes(tile_first_byte + tile_bytes); }

		// 2. per-vector byte offsets and the size words, to shared memory
		if (E * bits == 8 && lane_elements == elements_per_lane) {
			// a vector's fields fill a byte of a field word: byte-wise sums, then a multiply gives the prefix sums
			unsigned running = inclusive - lane_bytes;
			#pragma unroll
			for (unsigned k = 0; k < bits; k++) {
				const uint32_t sums = field_sums_per_byte<bits>(words[k]) + (E * min_size) * 0x01010101u; // bytes of the 4 vectors
				const uint32_t before = sums * 0x01010100u;                // byte v: bytes of the vectors before vector v
				const uint32_t base = running * 0x00010001u;
				const uint32_t offsets01 = __byte_perm(before, 0u, 0x4140) + base;
				const uint32_t offsets23 = __byte_perm(before, 0u, 0x4342) + base;
				reinterpret_cast<uint2*>(tile.offsets)[lane * bits + k] = make_uint2(offsets01, offsets23);
				running += (sums * 0x01010101u) >> 24;
			}
		}
		else {
			unsigned running = inclusive - lane_bytes;
			constexpr unsigned fields_per_pack = (vectors_per_lane >= 2) ? 2 : 1;
			#pragma unroll
			for (unsigned v = 0; v < vectors_per_lane; v += fields_per_pack) {
				uint32_t packed = 0;
				#pragma unroll
				for (unsigned h = 0; h < fields_per_pack; h++) {
					packed |= (running & 0xffffu) << (16 * h);
					// add the sizes of this vector's E elements
					constexpr unsigned fields_mask = (E * bits >= 32) ? 0xffffffffu : ((1u << ((E * bits) % 32)) - 1u);
					const unsigned first_bit = (v + h) * E * bits; // a compile-time constant once unrolled; < 32 * bits
					const uint32_t chunk = (words[first_bit / 32] >> (first_bit % 32)) & fields_mask;
					const unsigned elements_before = (v + h) * E;
					const unsigned elements_here = (lane_elements > elements_before) ? minimum<unsigned>(E, lane_elements - elements_before) : 0u;
					running += sum_of_fields<bits>(chunk) + elements_here * min_size;
				}
				if (fields_per_pack == 2) { reinterpret_cast<uint32_t*>(tile.offsets)[(lane * vectors_per_lane + v) / 2] = packed; }
				else { tile.offsets[lane * vectors_per_lane + v] = (uint16_t) packed; }
			}
		}
		#pragma unroll
		for (unsigned k = 0; k < bits; k++) { tile.fields[lane * bits + k] = words[k]; }
		__syncwarp();
		mbarrier_wait(tile.barrier, barrier_phase);
		barrier_phase ^= 1u;

		// 3. expand
		const unsigned head = (unsigned)((reinterpret_cast<uintptr_t>(compressed_data) + tile_first_byte) & 15u);
		const uint8_t* const bytes = tile.data; // 16-byte aligned; the tile's first byte is bytes[head]
		if (tile_length == tile_elements) { // a full tile: every offset below is an immediate
			constexpr unsigned words_per_row = warp_size * E * bits / 32;
			const uint16_t* lane_offsets = tile.offsets + lane;
			const uint32_t* lane_fields = tile.fields + (lane * E * bits) / 32;
			const unsigned field_shift = (lane * E * bits) % 32;
			#pragma unroll
			for (unsigned row = 0; row < tile_elements / (warp_size * E); row++) {
				const unsigned at = head + lane_offsets[row * warp_size];
				const uint32_t chunk = lane_fields[row * words_per_row] >> field_shift;
				sink.emit(tile_start + (row * warp_size + lane) * E, assemble_vector<U, bits>(bytes, at, chunk, min_size), E);
			}
		}
		else {
			const unsigned num_rows = (tile_length + warp_size * E - 1) / (warp_size * E);
			for (unsigned row = 0; row < num_rows; row++) { // (warp-uniform trip count: sinks may use warp collectives)
				const unsigned vector = row * warp_size + lane;
				const unsigned p = vector * E;
				const unsigned valid = (p < tile_length) ? minimum<unsigned>(E, tile_length - p) : 0u;
				const unsigned at = head + tile.offsets[vector];
				const unsigned first_bit = p * bits;
				const uint32_t chunk = tile.fields[first_bit / 32] >> (first_bit % 32);
				byte_pack<16> vec = assemble_vector<U, bits>(bytes, at, chunk, min_size);
				if (valid < E) {
					#pragma unroll
					for (unsigned j = 0; j < E; j++) { if (j >= valid) { vec.template set<U>(j, 0); } }
				}
				sink.emit(tile_start + p, vec, valid);
			}
		}
		__syncwarp(); // the tile's shared memory is free again
		tile_first_byte += tile_bytes;
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
__global__ void __launch_bounds__(GIDDY_SEGMENT_WARPS_PER_BLOCK * 32, GIDDY_DZBV_MIN_BLOCKS) discard_zero_bytes_variable_decompress(
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
	using namespace dzbv_detail;
	const uint64_t num_segments = ((uint64_t) length + segment_length - 1) / segment_length;
	const bool vectorizable = is_aligned_to(decompressed, 16) && (segment_length % E == 0);

	// The staged path: a flat stream of 1-, 2- or 4-bit fields (no slack bits in a container), every segment beginning on
	// a 32-bit word of it, E fields never straddling a word, and the shared memory to stage tiles in.
	const unsigned bits = bits_per_element_size;
	const bool staged = vectorizable && (bits == 1 || bits == 2 || bits == 4) && (E * bits <= 32) &&
		(((uint64_t) segment_length * bits) % 32 == 0) &&
		(dynamic_shared_memory_size() >= warps_per_block() * shared_memory_bytes_per_warp<U>());

	if (staged) {
		const tile_storage<U> tile(warp_shared_memory(shared_memory_bytes_per_warp<U>()));
		if (lane_index() == 0) { mbarrier_init(tile.barrier, 1); }
		__syncwarp();
		uint32_t barrier_phase = 0;
		const uint32_t* field_words = reinterpret_cast<const uint32_t*>(packed_element_sizes);
		const uint64_t sizes_per_container = (8 * CS) / bits;
		const uint64_t num_field_words = (((uint64_t) length + sizes_per_container - 1) / sizes_per_container) * (CS / 4);
		for (uint64_t segment = global_warp_index(); segment < num_segments; segment += num_warps_in_grid()) {
			const uint64_t start = segment * segment_length;
			const uint64_t this_length = minimum<uint64_t>(segment_length, length - start);
			const uint64_t first_byte = __ldg(position_anchors + segment);
			const bool next_anchor_known = segment + 1 < num_segments;
			const uint64_t next_anchor = next_anchor_known ? (uint64_t) __ldg(position_anchors + segment + 1) : 0;
			store_sink<U> sink { decompressed + start };
			switch (bits) {
			case 1:
				decode_segment_staged<U, 1>(sink, compressed_data, first_byte, next_anchor_known, next_anchor,
					field_words, num_field_words, start, this_length, min_represented_element_size, tile, barrier_phase);
				break;
			case 2:
				decode_segment_staged<U, 2>(sink, compressed_data, first_byte, next_anchor_known, next_anchor,
					field_words, num_field_words, start, this_length, min_represented_element_size, tile, barrier_phase);
				break;
			default:
				decode_segment_staged<U, (E * 4 <= 32 ? 4 : 2)>(sink, compressed_data, first_byte, next_anchor_known, next_anchor,
					field_words, num_field_words, start, this_length, min_represented_element_size, tile, barrier_phase);
				break;
			}
		}
		return;
	}

	const size_field_locator<CS> locator(bits_per_element_size, min_represented_element_size);
	for (uint64_t segment = global_warp_index(); segment < num_segments; segment += num_warps_in_grid()) {
		const uint64_t start = segment * segment_length;
		const uint64_t this_length = minimum<uint64_t>(segment_length, length - start);
		const uint8_t* segment_bytes = compressed_data + __ldg(position_anchors + segment);
		if (vectorizable) {
			decode_segment<U, CS, E, offset_t>(decompressed + start, segment_bytes, packed_element_sizes, start, this_length, locator);
		}
		else {
			decode_segment<U, CS, 1, offset_t>(decompressed + start, segment_bytes, packed_element_sizes, start, this_length, locator);
		}
	}
}

} // namespace kernels
} // namespace giddy

#endif // GIDDY_B200_KERNELS_DZB_VARIABLE_CUH_
