// C ABI launcher for the incidence-bitmaps decoder (include/giddy_b200.h, kernel API).
// Instantiations: src/kernel_wrappers/decompression/incidence_bitmaps.cu:96-100.
#include "abi_internal.hpp"
#include "../kernels/bitmaps.cuh"

using namespace giddy;
using namespace giddy::abi;

namespace {

using bitmaps_launcher_t = giddy_status_t (*)(void*, const giddy_launch_config_t*, void*, const void*, const void*, uint64_t, unsigned);
struct bitmaps_entry_t { unsigned I, U, A; bool use_dictionary; const void* kernel; bitmaps_launcher_t launch; };

template <unsigned I, unsigned U, unsigned A, bool D>
giddy_status_t bitmaps_launch(void* stream, const giddy_launch_config_t* cfg, void* out, const void* bitmaps, const void* dict,
	uint64_t bitmap_length, unsigned num_bitmaps)
{
	return launch(kernels::incidence_bitmaps_decompress<I, U, A, D>, stream, cfg,
		static_cast<uint_t<U>*>(out), static_cast<const uint32_t*>(bitmaps), static_cast<const uint_t<U>*>(dict),
		(size_type<I>) bitmap_length, (uint32_t) num_bitmaps);
}
#define BITMAPS_ENTRY(I, U, A, D) { I, U, A, D, \
	reinterpret_cast<const void*>(&kernels::incidence_bitmaps_decompress<I, U, A, D>), &bitmaps_launch<I, U, A, D> }

const bitmaps_entry_t bitmaps_table[] = {
	BITMAPS_ENTRY(4, 1, 2, true), BITMAPS_ENTRY(4, 2, 2, true), BITMAPS_ENTRY(4, 4, 2, true), BITMAPS_ENTRY(4, 8, 2, true),
	BITMAPS_ENTRY(4, 1, 2, false),
};

const bitmaps_entry_t* find_bitmaps(unsigned I, unsigned U, unsigned A, bool D)
{
	for (const auto& e : bitmaps_table) { if (e.I == I && e.U == U && e.A == A && e.use_dictionary == D) { return &e; } }
	return nullptr;
}

} // namespace

extern "C" const void* giddy_b200_incidence_bitmaps_kernel(unsigned I, unsigned U, unsigned A, int use_dictionary)
{
	auto e = find_bitmaps(I, U, A, use_dictionary != 0);
	return e ? e->kernel : nullptr;
}

extern "C" giddy_status_t giddy_b200_incidence_bitmaps_launch(
	void* stream, const giddy_launch_config_t* config, unsigned I, unsigned U, unsigned A, int use_dictionary,
	void* decompressed, const void* bitmaps, const void* dictionary_entries, uint64_t bitmap_length, unsigned num_bitmaps)
{
	auto e = find_bitmaps(I, U, A, use_dictionary != 0);
	giddy_status_t status;
	if (!check_launch(config, e ? e->kernel : nullptr, status)) { return status; }
	if (num_bitmaps > 256) { return fail(GIDDY_INVALID_ARGUMENT, "at most 256 bitmaps are supported (bitmap indices are 8-bit)"); }
	if (bitmap_length == 0) { return GIDDY_SUCCESS; }
	return e->launch(stream, config, decompressed, bitmaps, dictionary_entries, bitmap_length, num_bitmaps);
}
