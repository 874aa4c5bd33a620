// giddy_oracle.cpp -- TEST INFRASTRUCTURE ONLY.  NOT PART OF THE PRODUCT PATH.
//
// Scalar host (CPU) decoders for the eleven Giddy formats, restated from the
// reference's scheme definitions.  The reference (eyalroz/libgiddy @ 74d3543)
// ships no CPU path (README.md:128-130), so each decoder below is written from
// the semantics of the reference's device kernel, cited per function.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs may load this library.  Nothing under libgiddy_b200/ links,
// imports or falls back to it.
//
// Parity pinning: the FOR decoder is pinned by the reference's only known-answer
// test (examples/src/modes_of_use/common.cuh:29-105, see tests/test_oracle_golden.py);
// all decoders are additionally pinned against outputs of the reference's own
// kernels rebuilt for sm_100a (oracle/_ref, fixtures under tests/golden/, made
// by tests/golden/make_ref_golden.py on a B200).
//
// All arithmetic is unsigned modular (two's complement), as in the reference
// (uint_t<N> containers, src/util/integer.h:203-210).  Everything is
// little-endian.  Each decoder is OpenMP-parallel over anchored segments
// (or plain index ranges) so that it doubles as the timed CPU baseline.

#include <cstdint>
#include <cstring>
#include <cstddef>
#include <type_traits>

#if defined(_OPENMP)
#include <omp.h>
#endif

namespace {

template <unsigned N> struct uint_of;
template <> struct uint_of<1> { using type = uint8_t;  };
template <> struct uint_of<2> { using type = uint16_t; };
template <> struct uint_of<4> { using type = uint32_t; };
template <> struct uint_of<8> { using type = uint64_t; };
template <unsigned N> using uint_t = typename uint_of<N>::type;

template <unsigned N> struct int_of;
template <> struct int_of<1> { using type = int8_t;  };
template <> struct int_of<2> { using type = int16_t; };
template <> struct int_of<4> { using type = int32_t; };
template <> struct int_of<8> { using type = int64_t; };
template <unsigned N> using int_t = typename int_of<N>::type;

enum { MODEL_CONSTANT = 0, MODEL_LINEAR = 1 };

// ---------------------------------------------------------------------------
// Parametric models.  Follows src/cuda/model_functors.hpp:87-90 (linear:
// c[0] + c[1] * x with x an unsigned index of IndexSize bytes, evaluated under
// the usual arithmetic conversions and then converted to T) and :108-111
// (constant: c[0]).  With T integral the result is the low 8*sizeof(T) bits of
// the wrapped sum, which is what the unsigned computation below yields.
// ---------------------------------------------------------------------------
template <unsigned TSize>
inline uint_t<TSize> eval_model(int kind, unsigned model_index_size, const int_t<TSize>* c, uint64_t x)
{
	using UT = uint_t<TSize>;
	if (kind == MODEL_CONSTANT) { return (UT) c[0]; }
	// x is first narrowed to the model's index type (uint_t<IndexSize>)
	uint64_t xi = (model_index_size == 8) ? x : (uint64_t)(uint32_t) x;
	uint64_t v = (uint64_t)(int64_t) c[0] + (uint64_t)(int64_t) c[1] * xi;
	return (UT) v;
}

// ---------------------------------------------------------------------------
// A.1 Frame of reference.  src/kernels/decompression/frame_of_reference.cuh:62-67,72-76
//   out[i] = (U)( (U) M_s(i - s*L) + in[i] ),   s = i / L
// Uncompressed / Compressed are signed (int16/int32, int8/int16) and the model's
// coefficient type is int32 in every registered instantiation
// (src/kernel_wrappers/decompression/frame_of_reference.cu:132-142).
// ---------------------------------------------------------------------------
template <unsigned U, unsigned C>
int for_decompress(void* out_, const void* in_, const void* coeffs_, uint64_t n, uint64_t L,
	int model_kind, unsigned model_index_size)
{
	auto out = static_cast<uint_t<U>*>(out_);
	auto in = static_cast<const int_t<C>*>(in_);
	auto coeffs = static_cast<const int32_t*>(coeffs_);
	const unsigned dim = (model_kind == MODEL_LINEAR) ? 2 : 1;
	const int64_t S = (int64_t)((n + L - 1) / L);
	#pragma omp parallel for schedule(static)
	for (int64_t s = 0; s < S; s++) {
		uint64_t base = (uint64_t) s * L;
		uint64_t len = (n - base < L) ? n - base : L;
		const int32_t* c = coeffs + (uint64_t) s * dim;
		for (uint64_t x = 0; x < len; x++) {
			uint_t<U> predicted = (uint_t<U>) eval_model<4>(model_kind, model_index_size, c, x);
			out[base + x] = (uint_t<U>)(predicted + (uint_t<U>)(int_t<U>) in[base + x]);
		}
	}
	return 0;
}

// ---------------------------------------------------------------------------
// A.2 Model.  src/kernels/decompression/model.cuh:38-53:  out[i] = (U) M(i)
// Coefficient type == Uncompressed in every registered instantiation
// (src/kernel_wrappers/decompression/model.cu:92-96).
// ---------------------------------------------------------------------------
template <unsigned U>
int model_decompress(void* out_, const void* coeffs_, uint64_t n, int model_kind, unsigned model_index_size)
{
	auto out = static_cast<uint_t<U>*>(out_);
	auto c = static_cast<const int_t<U>*>(coeffs_);
	#pragma omp parallel for schedule(static)
	for (int64_t i = 0; i < (int64_t) n; i++) {
		out[i] = eval_model<U>(model_kind, model_index_size, c, (uint64_t) i);
	}
	return 0;
}

// ---------------------------------------------------------------------------
// A.3 Run-length encoding.  src/kernels/decompression/run_length_encoding.cuh:284-325
// (per-segment parameters from the two anchor arrays) and :150-157 (walk the run
// lengths until the run covers the output position).
//   out[i] = run_data[r(i)]
// Segment s starts inside run anchors[s], intra_offsets[s] elements into it.
// The last run's stored length is not trusted (:307-313): it extends to n.
// ---------------------------------------------------------------------------
template <unsigned U, unsigned RL, unsigned I>
int rle_decompress(void* out_, const void* run_data_, const void* run_lengths_,
	const void* anchors_, const void* intra_offsets_,
	uint64_t L, uint64_t S, uint64_t R, uint64_t n)
{
	auto out = static_cast<uint_t<U>*>(out_);
	auto run_data = static_cast<const uint_t<U>*>(run_data_);
	auto run_lengths = static_cast<const uint_t<RL>*>(run_lengths_);
	auto anchors = static_cast<const uint_t<I>*>(anchors_);
	auto intra = static_cast<const uint_t<RL>*>(intra_offsets_);
	#pragma omp parallel for schedule(static)
	for (int64_t s = 0; s < (int64_t) S; s++) {
		uint64_t start = (uint64_t) s * L;
		if (start >= n) { continue; }
		uint64_t len = ((uint64_t) s + 1 == S) ? n - start : L;
		// walk the runs from the anchor's, filling each run's part of the segment
		uint64_t run = anchors[s];
		uint64_t run_end = (uint64_t) run_lengths[run] - intra[s]; // segment-relative end of the current run
		uint64_t p = 0;
		while (p < len) {
			if (run + 1 >= R || run_end > len) { run_end = len; } // the last run extends to the end of the column
			const uint_t<U> v = run_data[run];
			for (uint64_t q = p; q < run_end; q++) { out[start + q] = v; }
			p = run_end;
			if (p < len) { run++; run_end += run_lengths[run]; }
		}
	}
	return 0;
}

// ---------------------------------------------------------------------------
// A.4 Run-position encoding.  src/kernels/decompression/run_position_encoding.cuh:271-307
// (which runs belong to a segment) and :138-147 (a run covers a position until
// the next run's start position exceeds it; the first run of a segment covers
// the segment start without its start position being read, :122).
//   out[i] = run_data[ max r : start[r] <= i ]
// Absolute mode: start positions index the whole column.  Relative mode
// (PositionOffsetSize < index size): a run's start position is relative to the
// start of the segment in which the run begins.  In relative mode the
// reference's "does the run at the next anchor still belong to me" test
// (:292-297) compares a relative with an absolute position; this
// decoder implements the evident intent: that run belongs to segment s iff it
// begins strictly before segment s+1.
// ---------------------------------------------------------------------------
template <unsigned U, unsigned PO, unsigned I>
int rpe_decompress(void* out_, const void* run_data_, const void* run_starts_, const void* anchors_,
	uint64_t L, uint64_t S, uint64_t R, uint64_t n, bool relative)
{
	auto out = static_cast<uint_t<U>*>(out_);
	auto run_data = static_cast<const uint_t<U>*>(run_data_);
	auto starts = static_cast<const uint_t<PO>*>(run_starts_);
	auto anchors = static_cast<const uint_t<I>*>(anchors_);
	#pragma omp parallel for schedule(static)
	for (int64_t s = 0; s < (int64_t) S; s++) {
		uint64_t seg_start = (uint64_t) s * L;
		if (seg_start >= n) { continue; }
		uint64_t len = ((uint64_t) s + 1 == S) ? n - seg_start : L;
		uint64_t first = anchors[s];
		uint64_t past_last;
		if ((uint64_t) s + 1 == S) { past_last = R; }
		else {
			uint64_t next = anchors[s + 1];
			bool next_run_begins_before_next_segment = relative ?
				(next == first || starts[next] != 0) :
				((uint64_t) starts[next] < seg_start + L);
			past_last = next + (next_run_begins_before_next_segment ? 1 : 0);
		}
		uint64_t base = relative ? 0 : seg_start;
		uint64_t p = 0;
		for (uint64_t run = first; p < len; run++) {
			// this run covers positions up to the next run's start (or the segment's end)
			uint64_t end = (run + 1 < past_last) ? (uint64_t) starts[run + 1] - base : len;
			if (end > len) { end = len; }
			const uint_t<U> v = run_data[run];
			for (uint64_t q = p; q < end; q++) { out[seg_start + q] = v; }
			if (end > p) { p = end; }
			if (run + 1 >= past_last) { break; }
		}
	}
	return 0;
}

// ---------------------------------------------------------------------------
// A.5 Delta.  src/kernels/decompression/delta.cuh:87-106 + the inclusive scan of
// src/kernels/reduction/scan.cuh:354-360.  Deltas are unsigned and zero-extended
// (delta.cuh:74,95,100).
//   out[i] = anchor[s] + sum_{j = s*L .. i} zext(d[j])    (mod 2^(8U))
// ---------------------------------------------------------------------------
template <unsigned U, unsigned C>
int delta_decompress(void* out_, const void* in_, const void* anchors_, uint64_t n, uint64_t L)
{
	auto out = static_cast<uint_t<U>*>(out_);
	auto in = static_cast<const uint_t<C>*>(in_);
	auto anchors = static_cast<const uint_t<U>*>(anchors_);
	const int64_t S = (int64_t)((n + L - 1) / L);
	#pragma omp parallel for schedule(static)
	for (int64_t s = 0; s < S; s++) {
		uint64_t base = (uint64_t) s * L;
		uint64_t len = (n - base < L) ? n - base : L;
		uint_t<U> acc = anchors[s];
		for (uint64_t x = 0; x < len; x++) {
			acc = (uint_t<U>)(acc + (uint_t<U>) in[base + x]);
			out[base + x] = acc;
		}
	}
	return 0;
}

// ---------------------------------------------------------------------------
// A.6 Dictionary.  src/kernels/data_layout/gather.cuh:140-142 (and :97-137 for
// 1- and 2-byte elements): out[i] = dictionary[in[i]].  No bounds check.
// ---------------------------------------------------------------------------
template <unsigned U, unsigned DI>
int dict_decompress(void* out_, const void* in_, const void* dict_, uint64_t n)
{
	auto out = static_cast<uint_t<U>*>(out_);
	auto in = static_cast<const uint_t<DI>*>(in_);
	auto dict = static_cast<const uint_t<U>*>(dict_);
	#pragma omp parallel for schedule(static)
	for (int64_t i = 0; i < (int64_t) n; i++) { out[i] = dict[in[i]]; }
	return 0;
}

// ---------------------------------------------------------------------------
// A.7 Discard zero bytes, fixed width.  src/kernels/decompression/discard_zero_bytes_fixed.cuh:55-76
// (little-endian branch): the compressed value occupies the low bytes, the high
// bytes are zero:  out[i] = zext(in[i]).
// ---------------------------------------------------------------------------
template <unsigned U, unsigned C>
int dzbf_decompress(void* out_, const void* in_, uint64_t n)
{
	auto out = static_cast<uint_t<U>*>(out_);
	auto in = static_cast<const uint_t<C>*>(in_);
	#pragma omp parallel for schedule(static)
	for (int64_t i = 0; i < (int64_t) n; i++) { out[i] = (uint_t<U>) in[i]; }
	return 0;
}

// ---------------------------------------------------------------------------
// A.8 Discard zero bytes, variable width.
// src/kernels/decompression/discard_zero_bytes_variable.cuh:52-57 (size of element e:
// field e % spc of container e / spc, fields counted from the LSB, plus the
// minimum represented size), :292-309 (an element's bytes start where the
// previous element's end; exclusive prefix sum of the sizes), :397-438 (segment s
// starts at byte position_anchors[s]), :196-213 (the size_i low bytes are copied,
// the rest are zero).
// ---------------------------------------------------------------------------
template <unsigned U, unsigned CS, unsigned I>
int dzbv_decompress(void* out_, const void* bytes_, const void* sizes_, const void* anchors_,
	uint64_t L, uint64_t n, unsigned min_size, unsigned bits)
{
	auto out = static_cast<uint_t<U>*>(out_);
	auto bytes = static_cast<const uint8_t*>(bytes_);
	auto sizes = static_cast<const uint_t<CS>*>(sizes_);
	auto anchors = static_cast<const uint_t<I>*>(anchors_);
	const uint64_t spc = (8 * CS) / bits;
	const uint_t<CS> field_mask = (bits >= 8 * CS) ? ~(uint_t<CS>) 0 : (uint_t<CS>)(((uint_t<CS>) 1 << bits) - 1);
	const int64_t S = (int64_t)((n + L - 1) / L);
	#pragma omp parallel for schedule(static)
	for (int64_t s = 0; s < S; s++) {
		uint64_t base = (uint64_t) s * L;
		uint64_t len = (n - base < L) ? n - base : L;
		uint64_t pos = anchors[s];
		for (uint64_t x = 0; x < len; x++) {
			uint64_t e = base + x;
			uint_t<CS> container = sizes[e / spc];
			unsigned size = (unsigned)((container >> ((e % spc) * bits)) & field_mask) + min_size;
			uint_t<U> v = 0;
			if (size > U) { size = U; } // out of format; clamp rather than smash the stack
			memcpy(&v, bytes + pos, size);
			out[e] = v;
			pos += size;
		}
	}
	return 0;
}

// ---------------------------------------------------------------------------
// A.9 Incidence bitmaps.  src/kernels/decompression/incidence_bitmaps.cuh:47-60
// (for each bitmap, each set bit of the container at this position marks an
// element belonging to that bitmap; bit j, LSB first, is element 32*w + j) and
// :118-126 (translate through the dictionary, or cast the bitmap index).
// Bitmap b starts at word b * round_up(bitmap_length, alignment) (:94-95).
// Output is produced for whole containers: 32 * bitmap_length elements.
// Elements claimed by no bitmap are undefined in the reference (stale shared
// memory); this decoder gives them bitmap index 0.  Later bitmaps overwrite
// earlier ones when several claim an element, as in the reference's loop order.
// ---------------------------------------------------------------------------
template <unsigned U>
int bitmaps_decompress(void* out_, const void* bitmaps_, const void* dict_,
	uint64_t bitmap_length, unsigned num_bitmaps, unsigned alignment_in_ints, bool use_dictionary)
{
	auto out = static_cast<uint_t<U>*>(out_);
	auto bitmaps = static_cast<const uint32_t*>(bitmaps_);
	auto dict = static_cast<const uint_t<U>*>(dict_);
	uint64_t aligned = (bitmap_length + alignment_in_ints - 1) / alignment_in_ints * alignment_in_ints;
	#pragma omp parallel for schedule(static)
	for (int64_t w = 0; w < (int64_t) bitmap_length; w++) {
		uint8_t idx[32];
		memset(idx, 0, sizeof(idx));
		for (unsigned b = 0; b < num_bitmaps; b++) {
			uint32_t word = bitmaps[(uint64_t) b * aligned + w];
			while (word) {
				int j = __builtin_ctz(word);
				idx[j] = (uint8_t) b;
				word &= word - 1;
			}
		}
		for (int j = 0; j < 32; j++) {
			out[(uint64_t) w * 32 + j] = use_dictionary ? dict[idx[j]] : (uint_t<U>) idx[j];
		}
	}
	return 0;
}

// ---------------------------------------------------------------------------
// A.10 Scatter.  src/kernels/data_layout/scatter.cuh:42-44:
//   target[indices[i]] = data[i],  i < data_length.  Duplicate indices unsupported.
// ---------------------------------------------------------------------------
template <unsigned U, unsigned OI>
int scatter(void* target_, const void* data_, const void* indices_, uint64_t p)
{
	auto target = static_cast<uint_t<U>*>(target_);
	auto data = static_cast<const uint_t<U>*>(data_);
	auto indices = static_cast<const uint_t<OI>*>(indices_);
	#pragma omp parallel for schedule(static)
	for (int64_t i = 0; i < (int64_t) p; i++) { target[indices[i]] = data[i]; }
	return 0;
}

// ---------------------------------------------------------------------------
// A.11 Compressed-indices scatter.
// src/kernels/data_layout/compressed_indices_scatter.cuh:236-252 (anchoring period a
// covers runs [anchors[a], anchors[a+1]) -- or up to R for the last period -- and
// patch values starting at a*A) and :83-124 (runs are consumed in order; run r
// holds run_lengths[r] offsets of offset_sizes[r] bits each, starting at byte
// offsets_start[r] * OutputIndexSize of offset_bytes;
// target[(OI)(baseline[r] + offset)] = data[k]).
// ---------------------------------------------------------------------------
template <unsigned U, unsigned OI, unsigned II, unsigned RL>
int ci_scatter(void* target_, const void* data_, const void* run_lengths_, const void* offset_sizes_,
	const void* baselines_, const void* offsets_start_, const void* offset_bytes_, const void* anchors_,
	uint64_t A, uint64_t R, uint64_t p)
{
	auto target = static_cast<uint_t<U>*>(target_);
	auto data = static_cast<const uint_t<U>*>(data_);
	auto run_lengths = static_cast<const uint_t<RL>*>(run_lengths_);
	auto offset_sizes = static_cast<const uint8_t*>(offset_sizes_);
	auto baselines = static_cast<const uint_t<OI>*>(baselines_);
	auto offsets_start = static_cast<const uint_t<II>*>(offsets_start_);
	auto offset_bytes = static_cast<const uint8_t*>(offset_bytes_);
	auto anchors = static_cast<const uint_t<II>*>(anchors_);
	const int64_t num_periods = (int64_t)((p + A - 1) / A);
	#pragma omp parallel for schedule(static)
	for (int64_t a = 0; a < num_periods; a++) {
		uint64_t first_run = anchors[a];
		bool reaches_end = ((uint64_t)(a + 1) * A >= p);
		uint64_t past_last_run = reaches_end ? R : (uint64_t) anchors[a + 1];
		uint64_t k = (uint64_t) a * A;
		for (uint64_t r = first_run; r < past_last_run; r++) {
			unsigned width = offset_sizes[r] / 8;
			const uint8_t* run_offsets = offset_bytes + (uint64_t) offsets_start[r] * OI;
			uint64_t len = run_lengths[r];
			for (uint64_t j = 0; j < len; j++, k++) {
				uint64_t off = 0;
				memcpy(&off, run_offsets + j * width, width);
				uint_t<OI> pos = (uint_t<OI>)(baselines[r] + (uint_t<OI>) off);
				target[pos] = data[k];
			}
		}
	}
	return 0;
}

} // anonymous namespace

// ===========================================================================
// extern "C" dispatch on run-time byte sizes.  Return 0 on success, -1 when the
// size combination is not one the oracle instantiates.
// ===========================================================================

#define CASE2(a, b, A, B, CALL) if ((a) == (A) && (b) == (B)) { return CALL; }

extern "C" {

int giddy_oracle_num_threads(void)
{
#if defined(_OPENMP)
	return omp_get_max_threads();
#else
	return 1;
#endif
}

void giddy_oracle_set_num_threads(int n)
{
#if defined(_OPENMP)
	omp_set_num_threads(n);
#else
	(void) n;
#endif
}

int giddy_oracle_frame_of_reference(void* out, const void* in, const void* coeffs, uint64_t n, uint64_t L,
	unsigned U, unsigned C, int model_kind, unsigned model_index_size)
{
	if (L == 0) { return -2; }
	CASE2(U, C, 2, 1, (for_decompress<2, 1>(out, in, coeffs, n, L, model_kind, model_index_size)))
	CASE2(U, C, 4, 1, (for_decompress<4, 1>(out, in, coeffs, n, L, model_kind, model_index_size)))
	CASE2(U, C, 4, 2, (for_decompress<4, 2>(out, in, coeffs, n, L, model_kind, model_index_size)))
	CASE2(U, C, 8, 4, (for_decompress<8, 4>(out, in, coeffs, n, L, model_kind, model_index_size)))
	return -1;
}

int giddy_oracle_model(void* out, const void* coeffs, uint64_t n, unsigned U, int model_kind, unsigned model_index_size)
{
	switch (U) {
	case 2: return model_decompress<2>(out, coeffs, n, model_kind, model_index_size);
	case 4: return model_decompress<4>(out, coeffs, n, model_kind, model_index_size);
	case 8: return model_decompress<8>(out, coeffs, n, model_kind, model_index_size);
	}
	return -1;
}

#define RLE_CASES(I) \
	CASE2(U, RL, 1, 1, (rle_decompress<1, 1, I>(out, rd, rl, an, io, L, S, R, n))) \
	CASE2(U, RL, 2, 1, (rle_decompress<2, 1, I>(out, rd, rl, an, io, L, S, R, n))) \
	CASE2(U, RL, 4, 1, (rle_decompress<4, 1, I>(out, rd, rl, an, io, L, S, R, n))) \
	CASE2(U, RL, 8, 1, (rle_decompress<8, 1, I>(out, rd, rl, an, io, L, S, R, n))) \
	CASE2(U, RL, 1, 2, (rle_decompress<1, 2, I>(out, rd, rl, an, io, L, S, R, n))) \
	CASE2(U, RL, 2, 2, (rle_decompress<2, 2, I>(out, rd, rl, an, io, L, S, R, n))) \
	CASE2(U, RL, 4, 2, (rle_decompress<4, 2, I>(out, rd, rl, an, io, L, S, R, n))) \
	CASE2(U, RL, 8, 2, (rle_decompress<8, 2, I>(out, rd, rl, an, io, L, S, R, n))) \
	CASE2(U, RL, 1, 4, (rle_decompress<1, 4, I>(out, rd, rl, an, io, L, S, R, n))) \
	CASE2(U, RL, 2, 4, (rle_decompress<2, 4, I>(out, rd, rl, an, io, L, S, R, n))) \
	CASE2(U, RL, 4, 4, (rle_decompress<4, 4, I>(out, rd, rl, an, io, L, S, R, n))) \
	CASE2(U, RL, 8, 4, (rle_decompress<8, 4, I>(out, rd, rl, an, io, L, S, R, n))) \
	CASE2(U, RL, 1, 8, (rle_decompress<1, 8, I>(out, rd, rl, an, io, L, S, R, n))) \
	CASE2(U, RL, 4, 8, (rle_decompress<4, 8, I>(out, rd, rl, an, io, L, S, R, n))) \
	CASE2(U, RL, 8, 8, (rle_decompress<8, 8, I>(out, rd, rl, an, io, L, S, R, n)))

int giddy_oracle_run_length_encoding(void* out, const void* rd, const void* rl, const void* an, const void* io,
	uint64_t L, uint64_t S, uint64_t R, uint64_t n, unsigned I, unsigned U, unsigned RL)
{
	if (L == 0) { return -2; }
	if (I == 4) { RLE_CASES(4) }
	if (I == 8) { RLE_CASES(8) }
	return -1;
}

#define RPE_CASES(PO, I) \
	if (U == 1) { return rpe_decompress<1, PO, I>(out, rd, rs, an, L, S, R, n, relative != 0); } \
	if (U == 2) { return rpe_decompress<2, PO, I>(out, rd, rs, an, L, S, R, n, relative != 0); } \
	if (U == 4) { return rpe_decompress<4, PO, I>(out, rd, rs, an, L, S, R, n, relative != 0); } \
	if (U == 8) { return rpe_decompress<8, PO, I>(out, rd, rs, an, L, S, R, n, relative != 0); }

int giddy_oracle_run_position_encoding(void* out, const void* rd, const void* rs, const void* an,
	uint64_t L, uint64_t S, uint64_t R, uint64_t n, unsigned I, unsigned U, unsigned PO, int relative)
{
	if (L == 0) { return -2; }
	if (I == 4 && PO == 2) { RPE_CASES(2, 4) }
	if (I == 4 && PO == 4) { RPE_CASES(4, 4) }
	if (I == 8 && PO == 2) { RPE_CASES(2, 8) }
	if (I == 8 && PO == 8) { RPE_CASES(8, 8) }
	return -1;
}

#define UC_CASES(FN, ...) \
	CASE2(U, C, 2, 1, (FN<2, 1>(__VA_ARGS__))) \
	CASE2(U, C, 4, 1, (FN<4, 1>(__VA_ARGS__))) \
	CASE2(U, C, 4, 2, (FN<4, 2>(__VA_ARGS__))) \
	CASE2(U, C, 8, 1, (FN<8, 1>(__VA_ARGS__))) \
	CASE2(U, C, 8, 2, (FN<8, 2>(__VA_ARGS__))) \
	CASE2(U, C, 8, 4, (FN<8, 4>(__VA_ARGS__)))

int giddy_oracle_delta(void* out, const void* in, const void* anchors, uint64_t n, uint64_t L, unsigned U, unsigned C)
{
	if (L == 0) { return -2; }
	UC_CASES(delta_decompress, out, in, anchors, n, L)
	CASE2(U, C, 1, 1, (delta_decompress<1, 1>(out, in, anchors, n, L)))
	CASE2(U, C, 2, 2, (delta_decompress<2, 2>(out, in, anchors, n, L)))
	CASE2(U, C, 4, 4, (delta_decompress<4, 4>(out, in, anchors, n, L)))
	CASE2(U, C, 8, 8, (delta_decompress<8, 8>(out, in, anchors, n, L)))
	return -1;
}

int giddy_oracle_discard_zero_bytes_fixed(void* out, const void* in, uint64_t n, unsigned U, unsigned C)
{
	UC_CASES(dzbf_decompress, out, in, n)
	return -1;
}

int giddy_oracle_dictionary(void* out, const void* in, const void* dict, uint64_t n, unsigned U, unsigned DI)
{
	#define DICT_CASE(UU, DD) CASE2(U, DI, UU, DD, (dict_decompress<UU, DD>(out, in, dict, n)))
	DICT_CASE(1, 1) DICT_CASE(2, 1) DICT_CASE(2, 2) DICT_CASE(4, 1) DICT_CASE(4, 2) DICT_CASE(4, 4)
	DICT_CASE(8, 1) DICT_CASE(8, 2) DICT_CASE(8, 4) DICT_CASE(8, 8)
	return -1;
}

int giddy_oracle_discard_zero_bytes_variable(void* out, const void* bytes, const void* sizes, const void* anchors,
	uint64_t L, uint64_t n, unsigned min_size, unsigned bits, unsigned I, unsigned U, unsigned CS)
{
	if (L == 0 || bits == 0 || bits > 8 * CS) { return -2; }
	#define DZBV_CASE(UU, CC, II) if (U == UU && CS == CC && I == II) { \
		return dzbv_decompress<UU, CC, II>(out, bytes, sizes, anchors, L, n, min_size, bits); }
	DZBV_CASE(2, 4, 4) DZBV_CASE(4, 4, 4) DZBV_CASE(8, 4, 4)
	DZBV_CASE(2, 4, 8) DZBV_CASE(4, 4, 8) DZBV_CASE(8, 4, 8)
	DZBV_CASE(2, 8, 4) DZBV_CASE(4, 8, 4) DZBV_CASE(8, 8, 4)
	return -1;
}

int giddy_oracle_incidence_bitmaps(void* out, const void* bitmaps, const void* dict, uint64_t bitmap_length,
	unsigned num_bitmaps, unsigned U, unsigned alignment_in_ints, int use_dictionary)
{
	if (alignment_in_ints == 0) { return -2; }
	switch (U) {
	case 1: return bitmaps_decompress<1>(out, bitmaps, dict, bitmap_length, num_bitmaps, alignment_in_ints, use_dictionary != 0);
	case 2: return bitmaps_decompress<2>(out, bitmaps, dict, bitmap_length, num_bitmaps, alignment_in_ints, use_dictionary != 0);
	case 4: return bitmaps_decompress<4>(out, bitmaps, dict, bitmap_length, num_bitmaps, alignment_in_ints, use_dictionary != 0);
	case 8: return bitmaps_decompress<8>(out, bitmaps, dict, bitmap_length, num_bitmaps, alignment_in_ints, use_dictionary != 0);
	}
	return -1;
}

int giddy_oracle_scatter(void* target, const void* data, const void* indices, uint64_t p, unsigned OI, unsigned U)
{
	#define SC_CASE(UU, OO) CASE2(U, OI, UU, OO, (scatter<UU, OO>(target, data, indices, p)))
	SC_CASE(1, 4) SC_CASE(2, 4) SC_CASE(4, 4) SC_CASE(8, 4)
	SC_CASE(1, 8) SC_CASE(2, 8) SC_CASE(4, 8) SC_CASE(8, 8)
	return -1;
}

int giddy_oracle_compressed_indices_scatter(void* target, const void* data, const void* run_lengths,
	const void* offset_sizes, const void* baselines, const void* offsets_start, const void* offset_bytes,
	const void* anchors, uint64_t A, uint64_t R, uint64_t p, unsigned OI, unsigned U, unsigned II)
{
	if (A == 0) { return -2; }
	#define CI_CASE(UU, OO, IIX) if (U == UU && OI == OO && II == IIX) { \
		return ci_scatter<UU, OO, IIX, IIX>(target, data, run_lengths, offset_sizes, baselines, offsets_start, offset_bytes, anchors, A, R, p); }
	#define CI_CASES_FOR_U(UU) \
		CI_CASE(UU, 4, 1) CI_CASE(UU, 4, 2) CI_CASE(UU, 4, 4) CI_CASE(UU, 4, 8) \
		CI_CASE(UU, 8, 1) CI_CASE(UU, 8, 2) CI_CASE(UU, 8, 4) CI_CASE(UU, 8, 8)
	CI_CASES_FOR_U(1) CI_CASES_FOR_U(2) CI_CASES_FOR_U(4) CI_CASES_FOR_U(8)
	return -1;
}

} // extern "C"
