// Shard planning (include/giddy_b200.h, group 5): a contiguous range [s0, s1) of a column's anchored segments as a
// self-contained column -- which parts of the compressed arrays it reads, what its scalar arguments are, and what must
// be subtracted from the anchors (run indices, byte offsets, absolute positions, patch targets) when those parts are
// copied into buffers of the shard's own.
//
// This is the "host-side exclusive scan of segment metadata" of the multi-GPU path, and it is O(1) per shard: the
// carry-ins a shard needs are already in the formats' anchor arrays --
//   RLE   run index + intra-run offset of the segment's first element   (reference run_length_encoding.cuh:284-325)
//   RPE   run index of the segment's first element                       (run_position_encoding.cuh:122,138-147)
//   DZB-V byte offset of the segment's first element                     (discard_zero_bytes_variable.cuh:397-438)
//   Delta the value preceding the segment                                (delta.cuh:80-92)
// -- so planning only reads two neighbouring anchors.  The reference itself is single-GPU
// (src/kernel_wrappers/registered_wrapper.h:82) and has no counterpart of this file.
#include "giddy_b200.h"

#include <cstring>
#include <string>

namespace giddy { namespace abi {
giddy_status_t fail(giddy_status_t status, const ::std::string& message);
} }

namespace {

using giddy::abi::fail;

uint64_t read_unsigned(const void* array, uint64_t index, unsigned element_size)
{
	const unsigned char* p = static_cast<const unsigned char*>(array) + index * element_size;
	switch (element_size) {
	case 1: return *p;
	case 2: { uint16_t v; ::std::memcpy(&v, p, 2); return v; }
	case 4: { uint32_t v; ::std::memcpy(&v, p, 4); return v; }
	default: { uint64_t v; ::std::memcpy(&v, p, 8); return v; }
	}
}

void write_unsigned(void* array, uint64_t index, unsigned element_size, uint64_t value)
{
	unsigned char* p = static_cast<unsigned char*>(array) + index * element_size;
	switch (element_size) {
	case 1: *p = (unsigned char) value; break;
	case 2: { uint16_t v = (uint16_t) value; ::std::memcpy(p, &v, 2); break; }
	case 4: { uint32_t v = (uint32_t) value; ::std::memcpy(p, &v, 4); break; }
	default: ::std::memcpy(p, &value, 8); break;
	}
}

bool is_element_size(unsigned size) { return size == 1 || size == 2 || size == 4 || size == 8; }

struct slice_list {
	giddy_shard_t& shard;
	void add(const char* array, uint64_t first, uint64_t count, unsigned element_size, uint64_t rebase = 0, bool saturate = false,
		uint64_t rows = 1, uint64_t pitch = 0)
	{
		giddy_slice_t& s = shard.slices[shard.num_slices++];
		s.array = array; s.first = first; s.count = count; s.element_size = element_size;
		s.rows = rows; s.pitch = pitch; s.rebase = rebase; s.saturate = saturate ? 1 : 0;
	}
};

// index of the first of `count` ascending elements which is >= value
uint64_t lower_bound(const void* array, uint64_t count, unsigned element_size, uint64_t value)
{
	uint64_t low = 0, high = count;
	while (low < high) {
		const uint64_t middle = low + (high - low) / 2;
		if (read_unsigned(array, middle, element_size) < value) { low = middle + 1; } else { high = middle; }
	}
	return low;
}

} // namespace

extern "C" {

giddy_status_t giddy_split_segments(uint64_t num_segments, unsigned parts, unsigned part, uint64_t* first, uint64_t* past_last)
{
	if (parts == 0 || part >= parts || first == nullptr || past_last == nullptr) { return fail(GIDDY_INVALID_ARGUMENT, "giddy_split_segments: bad part or null output"); }
	// (128-bit products: num_segments * part overflows 64 bits only for columns nobody has, but costs nothing to get right)
	*first = (uint64_t)(((unsigned __int128) num_segments * part) / parts);
	*past_last = (uint64_t)(((unsigned __int128) num_segments * (part + 1)) / parts);
	return GIDDY_SUCCESS;
}

giddy_status_t giddy_shard_plan(const giddy_column_layout_t* column, uint64_t s0, uint64_t s1, giddy_shard_t* out)
{
	if (column == nullptr || out == nullptr) { return fail(GIDDY_INVALID_ARGUMENT, "giddy_shard_plan: null column or output"); }
	::std::memset(out, 0, sizeof(*out));
	const uint64_t L = column->segment_length, n = column->length;
	if (L == 0) { return fail(GIDDY_INVALID_ARGUMENT, "The segment length must be a positive number"); }
	const uint64_t num_segments = (n + L - 1) / L;
	if (s0 > s1 || s1 > num_segments) { return fail(GIDDY_INVALID_ARGUMENT, "giddy_shard_plan: segment range outside the column"); }
	const uint64_t first_element = s0 * L;
	const uint64_t past_last_element = (s1 * L < n) ? s1 * L : n;
	const uint64_t num_elements = past_last_element > first_element ? past_last_element - first_element : 0;
	out->first_segment = s0; out->num_segments = s1 - s0;
	out->first_element = first_element; out->num_elements = num_elements;
	slice_list slices { *out };
	const bool is_last = s1 >= num_segments;
	const unsigned U = column->uncompressed_size, C = column->compressed_size, I = column->index_size;

	switch (column->scheme) {
	case GIDDY_SCHEME_FRAME_OF_REFERENCE:
		if (!is_element_size(C) || column->model_dimension == 0) { return fail(GIDDY_INVALID_ARGUMENT, "giddy_shard_plan: Frame of Reference needs compressed_size and model_dimension"); }
		// (the position inside the segment restarts at every segment: nothing to rebase)
		slices.add("compressed_input", first_element, num_elements, C);
		slices.add("segment_model_coefficients", s0, s1 - s0, 4 * column->model_dimension);
		return GIDDY_SUCCESS;
	case GIDDY_SCHEME_MODEL:
		return GIDDY_SUCCESS; // no arrays; see giddy_shard_model_coefficients
	case GIDDY_SCHEME_DELTA:
		if (!is_element_size(C) || !is_element_size(U)) { return fail(GIDDY_INVALID_ARGUMENT, "giddy_shard_plan: Delta needs uncompressed_size and compressed_size"); }
		// the carry-in of a shard is simply its first segment's anchor
		slices.add("compressed_input", first_element, num_elements, C);
		slices.add("anchor_values", s0, s1 - s0, U);
		return GIDDY_SUCCESS;
	case GIDDY_SCHEME_DICTIONARY:
	case GIDDY_SCHEME_DISCARD_ZERO_BYTES_FIXED:
		if (!is_element_size(C)) { return fail(GIDDY_INVALID_ARGUMENT, "giddy_shard_plan: compressed_size must be 1, 2, 4 or 8"); }
		slices.add("compressed_input", first_element, num_elements, C); // (the dictionary is replicated: broadcast once)
		return GIDDY_SUCCESS;
	case GIDDY_SCHEME_RUN_LENGTH_ENCODING:
	case GIDDY_SCHEME_RUN_POSITION_ENCODING: {
		if (column->anchors == nullptr || !is_element_size(I) || !is_element_size(U) || !is_element_size(C)) {
			return fail(GIDDY_INVALID_ARGUMENT, "giddy_shard_plan: run-based schemes need the host anchors and the element sizes");
		}
		if (s1 == s0) { return GIDDY_SUCCESS; }
		const uint64_t first_run = read_unsigned(column->anchors, s0, I);
		// the run holding the shard's last element is at or before the anchor run of the next segment (it IS that run
		// unless the next segment begins exactly on a run boundary: one run too many is copied then, and never read)
		uint64_t last_run = is_last ? column->num_runs - 1 : read_unsigned(column->anchors, s1, I);
		const bool rpe = column->scheme == GIDDY_SCHEME_RUN_POSITION_ENCODING;
		if (rpe && column->positions_are_relative && !is_last) {
			// Relative positions cannot tell "begins at the next segment's first element" from "begins at this segment's
			// first element" (both are stored as 0), and a column's last segment takes every remaining run as its own:
			// leave out the next shard's first run when it begins exactly on the boundary.
			if (column->run_start_positions == nullptr) { return fail(GIDDY_INVALID_ARGUMENT, "giddy_shard_plan: relative run positions need the host run_start_positions"); }
			const bool begins_in_this_shard = last_run == read_unsigned(column->anchors, s1 - 1, I) ||
				read_unsigned(column->run_start_positions, last_run, C) != 0;
			if (!begins_in_this_shard) { last_run--; }
		}
		if (last_run < first_run || last_run >= column->num_runs) { return fail(GIDDY_INVALID_ARGUMENT, "giddy_shard_plan: the anchors do not ascend within num_runs"); }
		const uint64_t num_runs = last_run - first_run + 1;
		out->num_runs = num_runs;
		slices.add("run_data", first_run, num_runs, U);
		if (rpe) {
			// absolute positions become relative to the shard's first element; the shard's first run may begin before it --
			// its start is never read (a segment's covering run is taken as beginning with the segment) -- so clamp it
			if (column->positions_are_relative) { slices.add("run_start_positions", first_run, num_runs, C); }
			else { slices.add("run_start_positions", first_run, num_runs, C, first_element, true); }
			slices.add("position_anchors", s0, s1 - s0, I, first_run);
		}
		else {
			slices.add("run_lengths", first_run, num_runs, C);
			slices.add("position_anchors", s0, s1 - s0, I, first_run);
			slices.add("intra_run_anchor_offsets", s0, s1 - s0, C);
		}
		return GIDDY_SUCCESS;
	}
	case GIDDY_SCHEME_DISCARD_ZERO_BYTES_VARIABLE: {
		if (column->anchors == nullptr || !is_element_size(I) || (C != 4 && C != 8) || column->bits_per_element_size == 0 || column->bits_per_element_size > 8 * C) {
			return fail(GIDDY_INVALID_ARGUMENT, "giddy_shard_plan: DZB-variable needs the host anchors, the container size and the bits per element size");
		}
		if (s1 == s0) { return GIDDY_SUCCESS; }
		const uint64_t sizes_per_container = (8 * C) / column->bits_per_element_size;
		if (first_element % sizes_per_container != 0) {
			return fail(GIDDY_INVALID_ARGUMENT, "a DZB-variable shard must begin on a size-container boundary (segment_length * first_segment must be a multiple of " +
				::std::to_string(sizes_per_container) + ")");
		}
		const uint64_t first_byte = read_unsigned(column->anchors, s0, I);
		const uint64_t past_last_byte = is_last ? column->num_bytes : read_unsigned(column->anchors, s1, I);
		if (past_last_byte < first_byte || past_last_byte > column->num_bytes) { return fail(GIDDY_INVALID_ARGUMENT, "giddy_shard_plan: the byte anchors do not ascend within num_bytes"); }
		out->num_bytes = past_last_byte - first_byte;
		const uint64_t first_container = first_element / sizes_per_container;
		const uint64_t past_last_container = (past_last_element + sizes_per_container - 1) / sizes_per_container;
		slices.add("compressed_data", first_byte, past_last_byte - first_byte, 1);
		slices.add("packed_element_sizes", first_container, past_last_container - first_container, C);
		slices.add("position_anchors", s0, s1 - s0, I, first_byte);
		return GIDDY_SUCCESS;
	}
	case GIDDY_SCHEME_INCIDENCE_BITMAPS: {
		// segments are ranges of 32-element containers here: segment_length must be a multiple of 32
		if (L % 32 != 0 || column->num_bitmaps == 0 || column->bitmap_alignment_in_ints == 0) {
			return fail(GIDDY_INVALID_ARGUMENT, "giddy_shard_plan: incidence bitmaps are cut at multiples of 32 elements and need num_bitmaps and the alignment");
		}
		const uint64_t a = column->bitmap_alignment_in_ints;
		const uint64_t bitmap_length = (n + 31) / 32;
		const uint64_t pitch = (bitmap_length + a - 1) / a * a;
		const uint64_t w0 = first_element / 32, w1 = (past_last_element + 31) / 32;
		out->bitmap_length = w1 - w0;
		slices.add("incidence_bitmaps", w0, w1 - w0, 4, 0, false, column->num_bitmaps, pitch);
		return GIDDY_SUCCESS;
	}
	case GIDDY_SCHEME_SCATTER: {
		// naive patches with ASCENDING target indices: the shard takes those whose target lies in its element range
		if (column->anchors == nullptr || !is_element_size(I) || !is_element_size(U)) { return fail(GIDDY_INVALID_ARGUMENT, "giddy_shard_plan: scatter needs the host indices and the element sizes"); }
		const uint64_t q0 = lower_bound(column->anchors, column->num_patches, I, first_element);
		const uint64_t q1 = lower_bound(column->anchors, column->num_patches, I, past_last_element);
		out->num_patches = q1 - q0;
		slices.add("data", q0, q1 - q0, U);
		slices.add("indices", q0, q1 - q0, I, first_element);
		return GIDDY_SUCCESS;
	}
	}
	return fail(GIDDY_INVALID_ARGUMENT, "giddy_shard_plan: unknown scheme");
}

size_t giddy_slice_bytes(const giddy_slice_t* slice)
{
	return slice ? (size_t)(slice->rows * slice->count * slice->element_size) : 0;
}

giddy_status_t giddy_slice_copy(const giddy_slice_t* slice, void* destination, const void* whole_array)
{
	if (slice == nullptr) { return fail(GIDDY_INVALID_ARGUMENT, "giddy_slice_copy: null slice"); }
	if (giddy_slice_bytes(slice) == 0) { return GIDDY_SUCCESS; }
	if (destination == nullptr || whole_array == nullptr) { return fail(GIDDY_INVALID_ARGUMENT, "giddy_slice_copy: null source or destination"); }
	const unsigned size = slice->element_size;
	const unsigned char* source = static_cast<const unsigned char*>(whole_array);
	unsigned char* target = static_cast<unsigned char*>(destination);
	for (uint64_t row = 0; row < slice->rows; row++) {
		const unsigned char* from = source + (row * slice->pitch + slice->first) * size;
		unsigned char* to = target + row * slice->count * size;
		if (slice->rebase == 0) { ::std::memcpy(to, from, (size_t)(slice->count * size)); continue; }
		if (!is_element_size(size)) { return fail(GIDDY_INVALID_ARGUMENT, "giddy_slice_copy: only integer elements can be rebased"); }
		for (uint64_t i = 0; i < slice->count; i++) {
			const uint64_t value = read_unsigned(from, i, size);
			write_unsigned(to, i, size, (slice->saturate && value < slice->rebase) ? 0 : value - slice->rebase);
		}
	}
	return GIDDY_SUCCESS;
}

giddy_status_t giddy_shard_model_coefficients(const int64_t* coefficients, unsigned dimension, unsigned coefficient_size, uint64_t first_element, int64_t* out)
{
	if (coefficients == nullptr || out == nullptr || dimension == 0 || dimension > 2 || !is_element_size(coefficient_size)) {
		return fail(GIDDY_INVALID_ARGUMENT, "giddy_shard_model_coefficients: 1 or 2 coefficients of 1, 2, 4 or 8 bytes");
	}
	out[0] = coefficients[0];
	if (dimension == 2) {
		out[1] = coefficients[1];
		uint64_t c0 = (uint64_t) coefficients[0] + (uint64_t) coefficients[1] * first_element; // mod 2^64, then narrowed
		const unsigned bits = 8 * coefficient_size;
		if (bits < 64) { // sign-extend from the coefficient type's width
			c0 &= (1ull << bits) - 1;
			if (c0 >> (bits - 1)) { c0 |= ~((1ull << bits) - 1); }
		}
		out[0] = (int64_t) c0;
	}
	return GIDDY_SUCCESS;
}

} // extern "C"
