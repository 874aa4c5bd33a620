// C ABI launchers for the run-based decoders (include/giddy_b200.h, kernel API).
#include "abi_internal.hpp"
#include "../kernels/runs.cuh"

using namespace giddy;
using namespace giddy::abi;

// ---- RLE: src/kernel_wrappers/decompression/run_length_encoding.cu:107-128 ---------------------------

namespace {

using rle_launcher_t = giddy_status_t (*)(void*, const giddy_launch_config_t*, void*, const void*, const void*, const void*, const void*,
	uint64_t, uint64_t, uint64_t, uint64_t);
struct rle_entry_t { unsigned UI, U, RL, RI; const void* kernel; rle_launcher_t launch; };

template <unsigned UI, unsigned U, unsigned RL, unsigned RI>
giddy_status_t rle_launch(void* stream, const giddy_launch_config_t* cfg, void* out, const void* run_data, const void* run_lengths,
	const void* anchors, const void* intra_offsets, uint64_t L, uint64_t S, uint64_t R, uint64_t n)
{
	return launch(kernels::run_length_encoding_decompress<UI, U, RL, RI>, stream, cfg,
		static_cast<uint_t<U>*>(out), static_cast<const uint_t<U>*>(run_data), static_cast<const uint_t<RL>*>(run_lengths),
		static_cast<const uint_t<UI>*>(anchors), static_cast<const uint_t<RL>*>(intra_offsets),
		(size_type<UI>) L, (size_type<UI>) S, (size_type<RI>) R, (size_type<UI>) n);
}
#define RLE_ENTRY(UI, U, RL) { UI, U, RL, UI, \
	reinterpret_cast<const void*>(&kernels::run_length_encoding_decompress<UI, U, RL, UI>), &rle_launch<UI, U, RL, UI> }

const rle_entry_t rle_table[] = {
	RLE_ENTRY(4, 1, 1), RLE_ENTRY(4, 2, 1), RLE_ENTRY(4, 4, 1), RLE_ENTRY(4, 8, 1),
	RLE_ENTRY(4, 1, 2), RLE_ENTRY(4, 2, 2), RLE_ENTRY(4, 4, 2), RLE_ENTRY(4, 8, 2),
	RLE_ENTRY(4, 1, 4), RLE_ENTRY(4, 2, 4), RLE_ENTRY(4, 4, 4), RLE_ENTRY(4, 8, 4),
	RLE_ENTRY(8, 4, 1), RLE_ENTRY(8, 4, 2), RLE_ENTRY(8, 4, 4),
	RLE_ENTRY(8, 1, 8), RLE_ENTRY(8, 4, 8), RLE_ENTRY(8, 8, 8),
};

const rle_entry_t* find_rle(unsigned UI, unsigned U, unsigned RL, unsigned RI)
{
	for (const auto& e : rle_table) { if (e.UI == UI && e.U == U && e.RL == RL && e.RI == RI) { return &e; } }
	return nullptr;
}

} // namespace

extern "C" const void* giddy_b200_run_length_encoding_kernel(unsigned UI, unsigned U, unsigned RL, unsigned RI)
{
	auto e = find_rle(UI, U, RL, RI);
	return e ? e->kernel : nullptr;
}

extern "C" giddy_status_t giddy_b200_run_length_encoding_launch(
	void* stream, const giddy_launch_config_t* config, unsigned UI, unsigned U, unsigned RL, unsigned RI,
	void* decompressed, const void* run_data, const void* run_lengths,
	const void* segment_position_anchors, const void* segment_position_anchor_intra_run_offsets,
	uint64_t segment_length, uint64_t num_segments, uint64_t num_element_runs, uint64_t total_uncompressed_length)
{
	auto e = find_rle(UI, U, RL, RI);
	giddy_status_t status;
	if (!check_launch(config, e ? e->kernel : nullptr, status)) { return status; }
	if (segment_length == 0) { return fail(GIDDY_INVALID_ARGUMENT, "The segment length must be a positive number"); }
	if (total_uncompressed_length == 0 || num_segments == 0) { return GIDDY_SUCCESS; }
	const unsigned warps = (config->block_dimensions[0] + 31) / 32;
	const unsigned needed = warps * (2 * kernels::runs::tile_words * 4 + kernels::runs::staged_capacity * U);
	if (config->dynamic_shared_memory_size < needed) {
		return fail(GIDDY_INVALID_ARGUMENT, "the launch configuration provides less dynamic shared memory than the kernel needs per warp");
	}
	return e->launch(stream, config, decompressed, run_data, run_lengths, segment_position_anchors,
		segment_position_anchor_intra_run_offsets, segment_length, num_segments, num_element_runs, total_uncompressed_length);
}

// ---- RPE: src/kernel_wrappers/decompression/run_position_encoding.cu:124-136 --------------------------

namespace {

using rpe_launcher_t = giddy_status_t (*)(void*, const giddy_launch_config_t*, void*, const void*, const void*, const void*,
	uint64_t, uint64_t, uint64_t, uint64_t);
struct rpe_entry_t { unsigned UI, U, PO; bool relative; unsigned RI; const void* kernel; rpe_launcher_t launch; };

template <unsigned UI, unsigned U, unsigned PO, bool Rel, unsigned RI>
giddy_status_t rpe_launch(void* stream, const giddy_launch_config_t* cfg, void* out, const void* run_data, const void* run_starts,
	const void* anchors, uint64_t L, uint64_t S, uint64_t R, uint64_t n)
{
	return launch(kernels::run_position_encoding_decompress<UI, U, PO, Rel, RI>, stream, cfg,
		static_cast<uint_t<U>*>(out), static_cast<const uint_t<U>*>(run_data), static_cast<const uint_t<PO>*>(run_starts),
		static_cast<const uint_t<RI>*>(anchors),
		(size_type<PO>) L, (size_type<RI>) S, (size_type<RI>) R, (size_type<UI>) n);
}
#define RPE_ENTRY(UI, U, PO, REL) { UI, U, PO, REL, UI, \
	reinterpret_cast<const void*>(&kernels::run_position_encoding_decompress<UI, U, PO, REL, UI>), &rpe_launch<UI, U, PO, REL, UI> }

const rpe_entry_t rpe_table[] = {
	RPE_ENTRY(4, 1, 2, true), RPE_ENTRY(4, 2, 2, true), RPE_ENTRY(4, 4, 2, true), RPE_ENTRY(4, 8, 2, true), RPE_ENTRY(8, 4, 2, true),
	RPE_ENTRY(4, 1, 4, false), RPE_ENTRY(4, 2, 4, false), RPE_ENTRY(4, 4, 4, false), RPE_ENTRY(4, 8, 4, false), RPE_ENTRY(8, 4, 8, false),
};

const rpe_entry_t* find_rpe(unsigned UI, unsigned U, unsigned PO, bool relative, unsigned RI)
{
	for (const auto& e : rpe_table) { if (e.UI == UI && e.U == U && e.PO == PO && e.relative == relative && e.RI == RI) { return &e; } }
	return nullptr;
}

} // namespace

extern "C" const void* giddy_b200_run_position_encoding_kernel(unsigned UI, unsigned U, unsigned PO, int relative, unsigned RI)
{
	auto e = find_rpe(UI, U, PO, relative != 0, RI);
	return e ? e->kernel : nullptr;
}

extern "C" giddy_status_t giddy_b200_run_position_encoding_launch(
	void* stream, const giddy_launch_config_t* config, unsigned UI, unsigned U, unsigned PO, int relative, unsigned RI,
	void* decompressed, const void* run_data, const void* run_start_positions, const void* position_anchors,
	uint64_t anchoring_period, uint64_t num_anchors, uint64_t num_element_runs, uint64_t uncompressed_length)
{
	auto e = find_rpe(UI, U, PO, relative != 0, RI);
	giddy_status_t status;
	if (!check_launch(config, e ? e->kernel : nullptr, status)) { return status; }
	if (anchoring_period == 0) { return fail(GIDDY_INVALID_ARGUMENT, "The segment length must be a positive number"); }
	if (uncompressed_length == 0 || num_anchors == 0) { return GIDDY_SUCCESS; }
	const unsigned warps = (config->block_dimensions[0] + 31) / 32;
	const unsigned needed = warps * (2 * kernels::runs::tile_words * 4 + kernels::runs::staged_capacity * U);
	if (config->dynamic_shared_memory_size < needed) {
		return fail(GIDDY_INVALID_ARGUMENT, "the launch configuration provides less dynamic shared memory than the kernel needs per warp");
	}
	return e->launch(stream, config, decompressed, run_data, run_start_positions, position_anchors,
		anchoring_period, num_anchors, num_element_runs, uncompressed_length);
}
