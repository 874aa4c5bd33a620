// C ABI launcher for the Discard-Zero-Bytes variable-width decoder (include/giddy_b200.h, kernel API).
// Instantiations: src/kernel_wrappers/decompression/discard_zero_bytes_variable.cu:109-119.
#include "abi_internal.hpp"
#include "../kernels/dzb_variable.cuh"

using namespace giddy;
using namespace giddy::abi;

namespace {

using dzbv_launcher_t = giddy_status_t (*)(void*, const giddy_launch_config_t*, void*, const void*, const void*, const void*,
	uint64_t, uint64_t, unsigned, unsigned);
struct dzbv_entry_t { unsigned I, U, CS; const void* kernel; dzbv_launcher_t launch; };

template <unsigned I, unsigned U, unsigned CS>
giddy_status_t dzbv_launch(void* stream, const giddy_launch_config_t* cfg, void* out, const void* bytes, const void* sizes,
	const void* anchors, uint64_t L, uint64_t n, unsigned min_size, unsigned bits)
{
	return launch(kernels::discard_zero_bytes_variable_decompress<I, U, CS>, stream, cfg,
		static_cast<uint_t<U>*>(out), static_cast<const uint8_t*>(bytes), static_cast<const uint_t<CS>*>(sizes),
		static_cast<const uint_t<I>*>(anchors), (size_type<I>) L, (size_type<I>) n, min_size, bits);
}
#define DZBV_ENTRY(I, U, CS) { I, U, CS, \
	reinterpret_cast<const void*>(&kernels::discard_zero_bytes_variable_decompress<I, U, CS>), &dzbv_launch<I, U, CS> }

const dzbv_entry_t dzbv_table[] = {
	DZBV_ENTRY(4, 2, 4), DZBV_ENTRY(4, 4, 4), DZBV_ENTRY(4, 8, 4),
	DZBV_ENTRY(8, 2, 4), DZBV_ENTRY(8, 4, 4), DZBV_ENTRY(8, 8, 4),
	DZBV_ENTRY(4, 2, 8), DZBV_ENTRY(4, 4, 8), DZBV_ENTRY(4, 8, 8),
};

const dzbv_entry_t* find_dzbv(unsigned I, unsigned U, unsigned CS)
{
	for (const auto& e : dzbv_table) { if (e.I == I && e.U == U && e.CS == CS) { return &e; } }
	return nullptr;
}

} // namespace

extern "C" const void* giddy_b200_discard_zero_bytes_variable_kernel(unsigned I, unsigned U, unsigned CS)
{
	auto e = find_dzbv(I, U, CS);
	return e ? e->kernel : nullptr;
}

extern "C" giddy_status_t giddy_b200_discard_zero_bytes_variable_launch(
	void* stream, const giddy_launch_config_t* config, unsigned I, unsigned U, unsigned CS,
	void* decompressed, const void* compressed_data, const void* packed_element_sizes, const void* position_anchors,
	uint64_t segment_length, uint64_t length, unsigned min_represented_element_size, unsigned bits_per_element_size)
{
	auto e = find_dzbv(I, U, CS);
	giddy_status_t status;
	if (!check_launch(config, e ? e->kernel : nullptr, status)) { return status; }
	if (segment_length == 0) { return fail(GIDDY_INVALID_ARGUMENT, "The segment length must be a positive number"); }
	if (bits_per_element_size == 0 || bits_per_element_size > 8 * CS) {
		return fail(GIDDY_INVALID_ARGUMENT, "bits_per_element_size must be between 1 and the number of bits in a sizes container");
	}
	if (length == 0) { return GIDDY_SUCCESS; }
	return e->launch(stream, config, decompressed, compressed_data, packed_element_sizes, position_anchors,
		segment_length, length, min_represented_element_size, bits_per_element_size);
}
