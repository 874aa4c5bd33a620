// C ABI launchers for the fused composite decoders (include/giddy_b200.h, kernel API; extensions).
#include "abi_internal.hpp"
#include "../kernels/composite.cuh"

using namespace giddy;
using namespace giddy::abi;

namespace {

using launcher_t = giddy_status_t (*)(void*, const giddy_launch_config_t*, void*, const void*, const void*, const void*, const void*,
	uint64_t, uint64_t, unsigned, unsigned);
struct entry_t { unsigned I, U, C, CS; const void* kernel; launcher_t launch; };

template <unsigned I, unsigned U, unsigned C, unsigned CS>
giddy_status_t do_launch(void* stream, const giddy_launch_config_t* cfg, void* out, const void* bytes, const void* sizes,
	const void* position_anchors, const void* anchor_values, uint64_t L, uint64_t n, unsigned min_size, unsigned bits)
{
	return launch(kernels::delta_of_discard_zero_bytes_variable_decompress<I, U, C, CS>, stream, cfg,
		static_cast<uint_t<U>*>(out), static_cast<const uint8_t*>(bytes), static_cast<const uint_t<CS>*>(sizes),
		static_cast<const uint_t<I>*>(position_anchors), static_cast<const uint_t<U>*>(anchor_values),
		(size_type<I>) L, (size_type<I>) n, min_size, bits);
}
#define ENTRY(I, U, C, CS) { I, U, C, CS, \
	reinterpret_cast<const void*>(&kernels::delta_of_discard_zero_bytes_variable_decompress<I, U, C, CS>), &do_launch<I, U, C, CS> }

const entry_t table[] = { ENTRY(4, 8, 4, 4), ENTRY(4, 8, 4, 8), ENTRY(4, 8, 2, 4), ENTRY(4, 4, 2, 4), ENTRY(8, 8, 4, 4) };

const entry_t* find(unsigned I, unsigned U, unsigned C, unsigned CS)
{
	for (const auto& e : table) { if (e.I == I && e.U == U && e.C == C && e.CS == CS) { return &e; } }
	return nullptr;
}

} // namespace

extern "C" const void* giddy_b200_delta_of_discard_zero_bytes_variable_kernel(unsigned I, unsigned U, unsigned C, unsigned CS)
{
	auto e = find(I, U, C, CS);
	return e ? e->kernel : nullptr;
}

extern "C" giddy_status_t giddy_b200_delta_of_discard_zero_bytes_variable_launch(
	void* stream, const giddy_launch_config_t* config, unsigned I, unsigned U, unsigned C, unsigned CS,
	void* decompressed, const void* compressed_data, const void* packed_element_sizes, const void* position_anchors,
	const void* anchor_values, uint64_t segment_length, uint64_t length,
	unsigned min_represented_element_size, unsigned bits_per_element_size)
{
	auto e = find(I, U, C, CS);
	giddy_status_t status;
	if (!check_launch(config, e ? e->kernel : nullptr, status)) { return status; }
	if (segment_length == 0) { return fail(GIDDY_INVALID_ARGUMENT, "The segment length must be a positive number"); }
	const unsigned bits = bits_per_element_size;
	if (!(bits == 1 || bits == 2 || bits == 4) || (16 / C) * bits > 32) {
		return fail(GIDDY_INVALID_ARGUMENT, "the fused decoder takes 1-, 2- or 4-bit element size fields");
	}
	if ((segment_length * bits) % 32 != 0) {
		return fail(GIDDY_INVALID_ARGUMENT, "the fused decoder needs segment_length * bits_per_element_size to be a multiple of 32");
	}
	const unsigned warps = (config->block_dimensions[0] + 31) / 32;
	if (config->dynamic_shared_memory_size < warps * GIDDY_DZBV_SHARED_BYTES_PER_WARP(C)) {
		return fail(GIDDY_INVALID_ARGUMENT, "the launch configuration provides less dynamic shared memory than the kernel needs per warp");
	}
	if (config->block_dimensions[0] % 32 != 0 || config->block_dimensions[0] > GIDDY_SEGMENT_WARPS_PER_BLOCK * 32) {
		return fail(GIDDY_INVALID_ARGUMENT, "blocks must be made of 1 to 8 whole warps");
	}
	if (length == 0) { return GIDDY_SUCCESS; }
	return e->launch(stream, config, decompressed, compressed_data, packed_element_sizes, position_anchors, anchor_values,
		segment_length, length, min_represented_element_size, bits_per_element_size);
}
