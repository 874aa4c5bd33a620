// C ABI launchers for the patching kernels (include/giddy_b200.h, kernel API).
// Instantiations: src/kernel_wrappers/data_layout/scatter.cu:92-124 and compressed_indices_scatter.cu:119-151:
// OutputIndexSize in {4, 8} x ElementSize in {1, 2, 4, 8} x InputIndexSize in {1, 2, 4, 8}.
#include "abi_internal.hpp"
#include "../kernels/scatter.cuh"

using namespace giddy;
using namespace giddy::abi;

namespace {

using scatter_launcher_t = giddy_status_t (*)(void*, const giddy_launch_config_t*, void*, const void*, const void*, uint64_t);
struct scatter_entry_t { unsigned OI, E, II; const void* kernel; scatter_launcher_t launch; };

template <unsigned OI, unsigned E, unsigned II>
giddy_status_t scatter_launch(void* stream, const giddy_launch_config_t* cfg, void* target, const void* data, const void* indices, uint64_t p)
{
	return launch(kernels::scatter<OI, E, II>, stream, cfg,
		static_cast<uint_t<E>*>(target), static_cast<const uint_t<E>*>(data), static_cast<const uint_t<OI>*>(indices), (size_type<II>) p);
}
#define SCATTER_ENTRY(OI, E, II) { OI, E, II, reinterpret_cast<const void*>(&kernels::scatter<OI, E, II>), &scatter_launch<OI, E, II> }
#define SCATTER_ENTRIES_E(OI, E) SCATTER_ENTRY(OI, E, 1), SCATTER_ENTRY(OI, E, 2), SCATTER_ENTRY(OI, E, 4), SCATTER_ENTRY(OI, E, 8)
#define SCATTER_ENTRIES(OI) SCATTER_ENTRIES_E(OI, 1), SCATTER_ENTRIES_E(OI, 2), SCATTER_ENTRIES_E(OI, 4), SCATTER_ENTRIES_E(OI, 8)

const scatter_entry_t scatter_table[] = { SCATTER_ENTRIES(4), SCATTER_ENTRIES(8) };

const scatter_entry_t* find_scatter(unsigned OI, unsigned E, unsigned II)
{
	for (const auto& e : scatter_table) { if (e.OI == OI && e.E == E && e.II == II) { return &e; } }
	return nullptr;
}

using ci_launcher_t = giddy_status_t (*)(void*, const giddy_launch_config_t*, void*, const void*, const void*, const void*, const void*,
	const void*, const void*, const void*, uint64_t, uint64_t, uint64_t);
struct ci_entry_t { unsigned OI, E, II, RL; const void* kernel; ci_launcher_t launch; };

template <unsigned OI, unsigned E, unsigned II, unsigned RL>
giddy_status_t ci_launch(void* stream, const giddy_launch_config_t* cfg, void* target, const void* data,
	const void* run_lengths, const void* offset_sizes, const void* baselines, const void* offsets_start,
	const void* offset_bytes, const void* anchors, uint64_t A, uint64_t R, uint64_t p)
{
	return launch(kernels::compressed_indices_scatter<OI, E, II, RL>, stream, cfg,
		static_cast<uint_t<E>*>(target), static_cast<const uint_t<E>*>(data),
		static_cast<const uint_t<RL>*>(run_lengths), static_cast<const unsigned char*>(offset_sizes),
		static_cast<const uint_t<OI>*>(baselines), static_cast<const uint_t<II>*>(offsets_start),
		static_cast<const unsigned char*>(offset_bytes), static_cast<const uint_t<II>*>(anchors),
		(uint_t<II>) A, (uint_t<II>) R, (size_type<II>) p);
}
#define CI_ENTRY(OI, E, II) { OI, E, II, II, \
	reinterpret_cast<const void*>(&kernels::compressed_indices_scatter<OI, E, II, II>), &ci_launch<OI, E, II, II> }
#define CI_ENTRIES_E(OI, E) CI_ENTRY(OI, E, 1), CI_ENTRY(OI, E, 2), CI_ENTRY(OI, E, 4), CI_ENTRY(OI, E, 8)
#define CI_ENTRIES(OI) CI_ENTRIES_E(OI, 1), CI_ENTRIES_E(OI, 2), CI_ENTRIES_E(OI, 4), CI_ENTRIES_E(OI, 8)

const ci_entry_t ci_table[] = { CI_ENTRIES(4), CI_ENTRIES(8) };

const ci_entry_t* find_ci(unsigned OI, unsigned E, unsigned II, unsigned RL)
{
	for (const auto& e : ci_table) { if (e.OI == OI && e.E == E && e.II == II && e.RL == RL) { return &e; } }
	return nullptr;
}

} // namespace

extern "C" const void* giddy_b200_scatter_kernel(unsigned OI, unsigned E, unsigned II)
{
	auto e = find_scatter(OI, E, II);
	return e ? e->kernel : nullptr;
}

extern "C" giddy_status_t giddy_b200_scatter_launch(
	void* stream, const giddy_launch_config_t* config, unsigned OI, unsigned E, unsigned II,
	void* target, const void* data, const void* indices, uint64_t data_length)
{
	auto e = find_scatter(OI, E, II);
	giddy_status_t status;
	if (!check_launch(config, e ? e->kernel : nullptr, status)) { return status; }
	if (data_length == 0) { return GIDDY_SUCCESS; }
	return e->launch(stream, config, target, data, indices, data_length);
}

extern "C" const void* giddy_b200_compressed_indices_scatter_kernel(unsigned OI, unsigned E, unsigned II, unsigned RL)
{
	auto e = find_ci(OI, E, II, RL);
	return e ? e->kernel : nullptr;
}

extern "C" giddy_status_t giddy_b200_compressed_indices_scatter_launch(
	void* stream, const giddy_launch_config_t* config, unsigned OI, unsigned E, unsigned II, unsigned RL,
	void* target, const void* data,
	const void* scatter_position_run_lengths, const void* scatter_position_run_individual_offset_sizes,
	const void* scatter_position_run_baseline_values, const void* scatter_position_run_offsets_start_positions,
	const void* scatter_position_offset_bytes, const void* scatter_position_anchors,
	uint64_t anchoring_period, uint64_t num_scatter_position_runs, uint64_t data_length)
{
	auto e = find_ci(OI, E, II, RL);
	giddy_status_t status;
	if (!check_launch(config, e ? e->kernel : nullptr, status)) { return status; }
	if (anchoring_period == 0) { return fail(GIDDY_INVALID_ARGUMENT, "The anchoring period must be a positive number"); }
	// an empty grid is silently a no-op in the reference (compressed_indices_scatter.cu:69-73)
	if (data_length == 0 || config->grid_dimensions[0] == 0) { return GIDDY_SUCCESS; }
	return e->launch(stream, config, target, data, scatter_position_run_lengths, scatter_position_run_individual_offset_sizes,
		scatter_position_run_baseline_values, scatter_position_run_offsets_start_positions, scatter_position_offset_bytes,
		scatter_position_anchors, anchoring_period, num_scatter_position_runs, data_length);
}
