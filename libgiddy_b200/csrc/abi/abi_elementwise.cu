// C ABI launchers for the position-aligned decoders (include/giddy_b200.h, kernel API).
// The instantiated type-size combinations are exactly those the reference registers in its factory
// (SURVEY.md Appendix C), cited per table.
#include "abi_internal.hpp"
#include "../kernels/elementwise.cuh"

using namespace giddy;
using namespace giddy::abi;

// ---- Frame of Reference: src/kernel_wrappers/decompression/frame_of_reference.cu:132-142 ----------

namespace {

using for_launcher_t = giddy_status_t (*)(void*, const giddy_launch_config_t*, void*, const void*, const void*, uint64_t, uint64_t);
struct for_entry_t { unsigned I, U, C; int model; unsigned MI; const void* kernel; for_launcher_t launch; };

template <unsigned I, unsigned U, unsigned C, int M, unsigned MI>
giddy_status_t for_launch(void* stream, const giddy_launch_config_t* cfg, void* out, const void* in, const void* coeffs, uint64_t n, uint64_t L)
{
	return launch(kernels::frame_of_reference_decompress<I, U, C, M, MI>, stream, cfg,
		static_cast<uint_t<U>*>(out), static_cast<const int_t<C>*>(in), static_cast<const int32_t*>(coeffs),
		(size_type<I>) n, (size_type<I>) L, (uint_t<I>) 0);
}
#define FOR_ENTRY(I, U, C, M, MI) { I, U, C, M, MI, \
	reinterpret_cast<const void*>(&kernels::frame_of_reference_decompress<I, U, C, M, MI>), &for_launch<I, U, C, M, MI> }

const for_entry_t for_table[] = {
	FOR_ENTRY(4, 2, 1, model_linear,   4), FOR_ENTRY(4, 4, 1, model_linear,   4), FOR_ENTRY(4, 4, 2, model_linear,   4),
	FOR_ENTRY(8, 4, 2, model_linear,   4), FOR_ENTRY(8, 4, 2, model_linear,   8),
	FOR_ENTRY(4, 2, 1, model_constant, 4), FOR_ENTRY(4, 4, 1, model_constant, 4), FOR_ENTRY(4, 4, 2, model_constant, 4),
};

const for_entry_t* find_for(unsigned I, unsigned U, unsigned C, int model, unsigned MI)
{
	for (const auto& e : for_table) { if (e.I == I && e.U == U && e.C == C && e.model == model && e.MI == MI) { return &e; } }
	return nullptr;
}

} // namespace

extern "C" const void* giddy_b200_frame_of_reference_kernel(unsigned I, unsigned U, unsigned C, giddy_model_t model, unsigned MI)
{
	auto e = find_for(I, U, C, (int) model, MI);
	return e ? e->kernel : nullptr;
}

extern "C" giddy_status_t giddy_b200_frame_of_reference_launch(
	void* stream, const giddy_launch_config_t* config,
	unsigned I, unsigned U, unsigned C, giddy_model_t model, unsigned MI,
	void* decompressed, const void* compressed_input, const void* segment_model_coefficients,
	uint64_t length, uint64_t segment_length)
{
	auto e = find_for(I, U, C, (int) model, MI);
	giddy_status_t status;
	if (!check_launch(config, e ? e->kernel : nullptr, status)) { return status; }
	if (segment_length == 0) { return fail(GIDDY_INVALID_ARGUMENT, "The segment length must be a positive number"); }
	if (length == 0) { return GIDDY_SUCCESS; }
	return e->launch(stream, config, decompressed, compressed_input, segment_model_coefficients, length, segment_length);
}

// ---- Model: src/kernel_wrappers/decompression/model.cu:92-96 ----------------------------------------

namespace {

using model_launcher_t = giddy_status_t (*)(void*, const giddy_launch_config_t*, void*, const void*, uint64_t);
struct model_entry_t { unsigned I, U; int model; const void* kernel; model_launcher_t launch; };

template <unsigned I, unsigned U, int M>
giddy_status_t model_launch(void* stream, const giddy_launch_config_t* cfg, void* out, const void* coeffs, uint64_t n)
{
	constexpr unsigned dim = (M == model_linear) ? 2 : 1;
	pod_array<int_t<U>, dim> c;
	for (unsigned d = 0; d < dim; d++) { c.v[d] = static_cast<const int_t<U>*>(coeffs)[d]; }
	return launch(kernels::model_decompress<I, U, M>, stream, cfg, static_cast<uint_t<U>*>(out), c, (size_type<I>) n);
}
#define MODEL_ENTRY(I, U, M) { I, U, M, reinterpret_cast<const void*>(&kernels::model_decompress<I, U, M>), &model_launch<I, U, M> }

const model_entry_t model_table[] = {
	MODEL_ENTRY(4, 2, model_linear), MODEL_ENTRY(4, 4, model_linear), MODEL_ENTRY(8, 4, model_linear),
	MODEL_ENTRY(4, 2, model_constant), MODEL_ENTRY(4, 4, model_constant),
};

const model_entry_t* find_model(unsigned I, unsigned U, int model)
{
	for (const auto& e : model_table) { if (e.I == I && e.U == U && e.model == model) { return &e; } }
	return nullptr;
}

} // namespace

extern "C" const void* giddy_b200_model_kernel(unsigned I, unsigned U, giddy_model_t model)
{
	auto e = find_model(I, U, (int) model);
	return e ? e->kernel : nullptr;
}

extern "C" giddy_status_t giddy_b200_model_launch(
	void* stream, const giddy_launch_config_t* config, unsigned I, unsigned U, giddy_model_t model,
	void* decompressed, const void* model_coefficients, uint64_t length)
{
	auto e = find_model(I, U, (int) model);
	giddy_status_t status;
	if (!check_launch(config, e ? e->kernel : nullptr, status)) { return status; }
	if (model_coefficients == nullptr) { return fail(GIDDY_INVALID_ARGUMENT, "null model coefficients"); }
	if (length == 0) { return GIDDY_SUCCESS; }
	return e->launch(stream, config, decompressed, model_coefficients, length);
}

// ---- DZB fixed width: src/kernel_wrappers/decompression/discard_zero_bytes_fixed.cu:113-125 --------
// ---- Dictionary:      src/kernel_wrappers/decompression/dictionary.cu:113-139 ----------------------

namespace {

using dzbf_launcher_t = giddy_status_t (*)(void*, const giddy_launch_config_t*, void*, const void*, uint64_t);
struct dzbf_entry_t { unsigned I, U, C; const void* kernel; dzbf_launcher_t launch; };

template <unsigned I, unsigned U, unsigned C>
giddy_status_t dzbf_launch(void* stream, const giddy_launch_config_t* cfg, void* out, const void* in, uint64_t n)
{
	return launch(kernels::discard_zero_bytes_fixed_decompress<I, U, C>, stream, cfg,
		static_cast<uint_t<U>*>(out), static_cast<const uint_t<C>*>(in), (size_type<I>) n);
}
#define DZBF_ENTRY(I, U, C) { I, U, C, reinterpret_cast<const void*>(&kernels::discard_zero_bytes_fixed_decompress<I, U, C>), &dzbf_launch<I, U, C> }
#define DZBF_ENTRIES(I) DZBF_ENTRY(I, 2, 1), DZBF_ENTRY(I, 4, 1), DZBF_ENTRY(I, 4, 2), DZBF_ENTRY(I, 8, 1), DZBF_ENTRY(I, 8, 2), DZBF_ENTRY(I, 8, 4)

const dzbf_entry_t dzbf_table[] = { DZBF_ENTRIES(4), DZBF_ENTRIES(8) };

const dzbf_entry_t* find_dzbf(unsigned I, unsigned U, unsigned C)
{
	for (const auto& e : dzbf_table) { if (e.I == I && e.U == U && e.C == C) { return &e; } }
	return nullptr;
}

using dict_launcher_t = giddy_status_t (*)(void*, const giddy_launch_config_t*, void*, const void*, const void*, uint64_t, uint64_t, bool, unsigned short);
struct dict_entry_t { unsigned I, U, DI; const void* kernel; dict_launcher_t launch; };

template <unsigned I, unsigned U, unsigned DI>
giddy_status_t dict_launch(void* stream, const giddy_launch_config_t* cfg, void* out, const void* in, const void* dict,
	uint64_t n, uint64_t num_entries, bool cache, unsigned short serialization_factor)
{
	return launch(kernels::dictionary_decompress<I, U, DI>, stream, cfg,
		static_cast<uint_t<U>*>(out), static_cast<const uint_t<DI>*>(in), static_cast<const uint_t<U>*>(dict),
		(size_type<I>) n, (size_type<DI>) num_entries, cache, serialization_factor);
}
#define DICT_ENTRY(I, U, DI) { I, U, DI, reinterpret_cast<const void*>(&kernels::dictionary_decompress<I, U, DI>), &dict_launch<I, U, DI> }
#define DICT_ENTRIES(I) DICT_ENTRY(I, 1, 1), DICT_ENTRY(I, 2, 1), DICT_ENTRY(I, 2, 2), DICT_ENTRY(I, 4, 1), DICT_ENTRY(I, 4, 2), \
	DICT_ENTRY(I, 4, 4), DICT_ENTRY(I, 8, 1), DICT_ENTRY(I, 8, 2), DICT_ENTRY(I, 8, 4), DICT_ENTRY(I, 8, 8)

const dict_entry_t dict_table[] = { DICT_ENTRIES(4), DICT_ENTRIES(8) };

const dict_entry_t* find_dict(unsigned I, unsigned U, unsigned DI)
{
	for (const auto& e : dict_table) { if (e.I == I && e.U == U && e.DI == DI) { return &e; } }
	return nullptr;
}

} // namespace

extern "C" const void* giddy_b200_discard_zero_bytes_fixed_kernel(unsigned I, unsigned U, unsigned C)
{
	auto e = find_dzbf(I, U, C);
	return e ? e->kernel : nullptr;
}

extern "C" giddy_status_t giddy_b200_discard_zero_bytes_fixed_launch(
	void* stream, const giddy_launch_config_t* config, unsigned I, unsigned U, unsigned C,
	void* decompressed, const void* compressed_input, uint64_t length)
{
	auto e = find_dzbf(I, U, C);
	giddy_status_t status;
	if (!check_launch(config, e ? e->kernel : nullptr, status)) { return status; }
	if (length == 0) { return GIDDY_SUCCESS; }
	return e->launch(stream, config, decompressed, compressed_input, length);
}

extern "C" const void* giddy_b200_dictionary_kernel(unsigned I, unsigned U, unsigned DI)
{
	auto e = find_dict(I, U, DI);
	return e ? e->kernel : nullptr;
}

extern "C" giddy_status_t giddy_b200_dictionary_launch(
	void* stream, const giddy_launch_config_t* config, unsigned I, unsigned U, unsigned DI,
	void* decompressed, const void* compressed_input, const void* dictionary_entries,
	uint64_t length, uint64_t num_dictionary_entries, int cache_dictionary_in_shared_memory, unsigned serialization_factor)
{
	auto e = find_dict(I, U, DI);
	giddy_status_t status;
	if (!check_launch(config, e ? e->kernel : nullptr, status)) { return status; }
	if (length == 0) { return GIDDY_SUCCESS; }
	// (with caching asked for, the kernel keeps as many leading entries in shared memory as the configuration's dynamic
	// shared memory holds -- all of them, if it was resolved for this dictionary -- and reads the rest through L1/L2)
	return e->launch(stream, config, decompressed, compressed_input, dictionary_entries, length, num_dictionary_entries,
		cache_dictionary_in_shared_memory != 0, (unsigned short) serialization_factor);
}

// ---- Utility kernels the reference has but does not wrap --------------------------------------------

namespace {

using intervals_launcher_t = giddy_status_t (*)(void*, const giddy_launch_config_t*, void*, const void*, uint64_t, uint64_t);
struct intervals_entry_t { unsigned I, U; int model; const void* kernel; intervals_launcher_t launch; };
template <unsigned I, unsigned U, int M>
giddy_status_t intervals_launch(void* stream, const giddy_launch_config_t* cfg, void* out, const void* coeffs, uint64_t period, uint64_t n)
{
	return launch(kernels::model_decompress_with_intervals<I, U, M>, stream, cfg, static_cast<uint_t<U>*>(out), static_cast<const int_t<U>*>(coeffs),
		(size_type<I>) period, (size_type<I>) n);
}
#define INTERVALS_ENTRY(I, U, M) { I, U, M, reinterpret_cast<const void*>(&kernels::model_decompress_with_intervals<I, U, M>), &intervals_launch<I, U, M> }
const intervals_entry_t intervals_table[] = {
	INTERVALS_ENTRY(4, 2, model_linear), INTERVALS_ENTRY(4, 4, model_linear), INTERVALS_ENTRY(8, 4, model_linear),
	INTERVALS_ENTRY(4, 2, model_constant), INTERVALS_ENTRY(4, 4, model_constant),
};
const intervals_entry_t* find_intervals(unsigned I, unsigned U, int model)
{
	for (const auto& e : intervals_table) { if (e.I == I && e.U == U && e.model == model) { return &e; } }
	return nullptr;
}

} // namespace

extern "C" const void* giddy_b200_model_with_intervals_kernel(unsigned I, unsigned U, giddy_model_t model)
{
	auto e = find_intervals(I, U, (int) model);
	return e ? e->kernel : nullptr;
}

extern "C" giddy_status_t giddy_b200_model_with_intervals_launch(
	void* stream, const giddy_launch_config_t* config, unsigned I, unsigned U, giddy_model_t model,
	void* decompressed, const void* interval_model_coefficients, uint64_t modeling_period, uint64_t length)
{
	auto e = find_intervals(I, U, (int) model);
	giddy_status_t status;
	if (!check_launch(config, e ? e->kernel : nullptr, status)) { return status; }
	if (modeling_period == 0) { return fail(GIDDY_INVALID_ARGUMENT, "The modeling period must be a positive number"); }
	if (length == 0) { return GIDDY_SUCCESS; }
	return e->launch(stream, config, decompressed, interval_model_coefficients, modeling_period, length);
}

// gather<OutputIndexSize, ElementSize, InputIndexSize> (reference src/kernels/data_layout/gather.cuh:151-164) is the kernel
// behind dictionary::decompress (dictionary.cuh:61-75 calls gather::detail::gather): the same launcher under its own name.
extern "C" const void* giddy_b200_gather_kernel(unsigned output_index_size, unsigned element_size, unsigned input_index_size)
{
	return giddy_b200_dictionary_kernel(output_index_size, element_size, input_index_size);
}

extern "C" giddy_status_t giddy_b200_gather_launch(
	void* stream, const giddy_launch_config_t* config, unsigned output_index_size, unsigned element_size, unsigned input_index_size,
	void* reordered_data, const void* data, const void* indices, uint64_t data_length, uint64_t num_indices,
	int cache_data_in_shared_memory, unsigned serialization_factor)
{
	return giddy_b200_dictionary_launch(stream, config, output_index_size, element_size, input_index_size, reordered_data, indices, data,
		num_indices, data_length, cache_data_in_shared_memory, serialization_factor);
}
