// The registry half of the C ABI (include/giddy_b200.h, group 1): a flat, type-erased binding of
//   cuda::registered::kernel_t::produceSubclass / resolve_launch_configuration / enqueue_launch /
//   get_device_function
// for callers which cannot speak C++ (the Python harness, or a cgo/JNI/N-API binding).
//
// This file uses nothing but the wrapper API the reference defines.  It is therefore compiled twice:
// against include/giddy/ into libgiddy_b200.so (the product), and -- with
// -DGIDDY_REGISTRY_FOR_REFERENCE -- against the reference's own headers and kernels into
// oracle/_ref/libgiddy_ref.so (test infrastructure; see oracle/build_ref.sh).  That the same source
// drives both is the drop-in claim made concrete.
#ifdef GIDDY_REGISTRY_FOR_REFERENCE
#include "kernel_wrappers/registered_wrapper.h"
#include "cuda/stl_array.cuh"
#include <any>
#else
#include "registered_wrapper.hpp"
#endif

#include "giddy_b200.h"

#include <algorithm>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

namespace giddy { namespace abi { giddy_status_t fail(giddy_status_t status, const ::std::string& message); } }

struct giddy_kernel {
	::std::unique_ptr<cuda::registered::kernel_t> wrapper;
	::std::string key;
};

namespace {

using giddy::abi::fail;
using arguments_type = cuda::kernel_t::arguments_type;
using mapped_type = arguments_type::mapped_type;

// Run a call into the C++ API, translating the exceptions the wrapper API is specified to throw
// (SURVEY.md 8b "Errors") into status codes.
template <typename F>
giddy_status_t guarded(F&& f)
{
	try { f(); return GIDDY_SUCCESS; }
	catch (const ::std::bad_any_cast& e) { return fail(GIDDY_BAD_ANY_CAST, ::std::string("bad any_cast: a named argument does not have exactly the expected type (") + e.what() + ")"); }
	catch (const ::std::out_of_range& e) { return fail(GIDDY_OUT_OF_RANGE, ::std::string("a required named argument is missing: ") + e.what()); }
	catch (const ::std::invalid_argument& e) { return fail(GIDDY_INVALID_ARGUMENT, e.what()); }
	catch (const ::std::logic_error& e) { return fail(GIDDY_LOGIC_ERROR, e.what()); }
	catch (const ::std::exception& e) { return fail(GIDDY_CUDA_ERROR, e.what()); }
	catch (...) { return fail(GIDDY_INTERNAL_ERROR, "unknown exception"); }
}

template <typename T>
mapped_type pointer_argument(const giddy_arg_t& a)
{
	if (a.type & GIDDY_T_CONST) { return mapped_type(static_cast<const T*>(a.value.p)); }
	return mapped_type(const_cast<T*>(static_cast<const T*>(a.value.p)));
}

template <typename T, unsigned N>
cuda::array<T, N> array_value(const giddy_arg_t& a)
{
	cuda::array<T, N> r;
	for (unsigned i = 0; i < N; i++) { r[i] = (T) a.value.a[i]; }
	return r;
}

// The value of a named argument, with exactly the static type its tag names.
mapped_type to_any(const giddy_arg_t& a)
{
	const int element = a.type & 0xff;
	if (a.type & GIDDY_T_PTR) {
		switch (element) {
		case GIDDY_T_U8:  return pointer_argument<::std::uint8_t>(a);
		case GIDDY_T_U16: return pointer_argument<::std::uint16_t>(a);
		case GIDDY_T_U32: return pointer_argument<::std::uint32_t>(a);
		case GIDDY_T_U64: return pointer_argument<::std::uint64_t>(a);
		case GIDDY_T_I8:  return pointer_argument<::std::int8_t>(a);
		case GIDDY_T_I16: return pointer_argument<::std::int16_t>(a);
		case GIDDY_T_I32: return pointer_argument<::std::int32_t>(a);
		case GIDDY_T_I64: return pointer_argument<::std::int64_t>(a);
		case GIDDY_T_BOOL: return pointer_argument<bool>(a);
		case GIDDY_T_ARRAY_I16_1: return pointer_argument<cuda::array<::std::int16_t, 1>>(a);
		case GIDDY_T_ARRAY_I16_2: return pointer_argument<cuda::array<::std::int16_t, 2>>(a);
		case GIDDY_T_ARRAY_I32_1: return pointer_argument<cuda::array<::std::int32_t, 1>>(a);
		case GIDDY_T_ARRAY_I32_2: return pointer_argument<cuda::array<::std::int32_t, 2>>(a);
		default: throw ::std::invalid_argument(::std::string("argument \"") + a.name + "\": unknown pointee type tag");
		}
	}
	switch (element) {
	case GIDDY_T_U8:   return mapped_type((::std::uint8_t) a.value.u);
	case GIDDY_T_U16:  return mapped_type((::std::uint16_t) a.value.u);
	case GIDDY_T_U32:  return mapped_type((::std::uint32_t) a.value.u);
	case GIDDY_T_U64:  return mapped_type((::std::uint64_t) a.value.u);
	case GIDDY_T_I8:   return mapped_type((::std::int8_t) a.value.i);
	case GIDDY_T_I16:  return mapped_type((::std::int16_t) a.value.i);
	case GIDDY_T_I32:  return mapped_type((::std::int32_t) a.value.i);
	case GIDDY_T_I64:  return mapped_type((::std::int64_t) a.value.i);
	case GIDDY_T_BOOL: return mapped_type((bool) (a.value.u != 0));
	case GIDDY_T_ARRAY_I16_1: return mapped_type(array_value<::std::int16_t, 1>(a));
	case GIDDY_T_ARRAY_I16_2: return mapped_type(array_value<::std::int16_t, 2>(a));
	case GIDDY_T_ARRAY_I32_1: return mapped_type(array_value<::std::int32_t, 1>(a));
	case GIDDY_T_ARRAY_I32_2: return mapped_type(array_value<::std::int32_t, 2>(a));
	default: throw ::std::invalid_argument(::std::string("argument \"") + a.name + "\": unknown type tag");
	}
}

arguments_type to_arguments(const giddy_arg_t* args, size_t count)
{
	arguments_type result;
	for (size_t i = 0; i < count; i++) {
		if (args[i].name == nullptr) { throw ::std::invalid_argument("a named argument has no name"); }
		result[args[i].name] = to_any(args[i]);
	}
	return result;
}

cuda::device::properties_t to_properties(const giddy_device_properties_t& p)
{
	cudaDeviceProp raw;
	::std::memset(&raw, 0, sizeof(raw));
	cuda::device::properties_t r(raw);
	r.major = p.compute_capability_major;
	r.minor = p.compute_capability_minor;
	r.multiProcessorCount = p.multiprocessor_count;
	r.maxThreadsPerBlock = p.max_threads_per_block;
	r.maxThreadsDim[0] = p.max_threads_dim_x;
	r.maxThreadsDim[1] = r.maxThreadsDim[2] = 1;
	r.maxThreadsPerMultiProcessor = p.max_threads_per_multiprocessor;
	r.maxBlocksPerMultiProcessor = p.max_blocks_per_multiprocessor;
	r.warpSize = p.warp_size;
	r.sharedMemPerBlock = p.shared_memory_per_block;
	r.sharedMemPerBlockOptin = p.shared_memory_per_block_optin;
	r.sharedMemPerMultiprocessor = p.shared_memory_per_multiprocessor;
	r.l2CacheSize = (int) p.l2_cache_size;
	r.totalGlobalMem = p.total_global_memory;
	return r;
}

cuda::device_function::attributes_t to_attributes(const giddy_function_attributes_t& a)
{
	cuda::device_function::attributes_t r;
	::std::memset(static_cast<cudaFuncAttributes*>(&r), 0, sizeof(cudaFuncAttributes));
	r.maxThreadsPerBlock = a.max_threads_per_block;
	r.numRegs = a.num_regs;
	r.sharedSizeBytes = a.shared_size_bytes;
	r.constSizeBytes = a.const_size_bytes;
	r.localSizeBytes = a.local_size_bytes;
	r.maxDynamicSharedSizeBytes = a.max_dynamic_shared_size_bytes;
	r.ptxVersion = a.ptx_version;
	r.binaryVersion = a.binary_version;
	return r;
}

cuda::launch_configuration_limits_t to_limits(const giddy_launch_config_limits_t* l)
{
	cuda::launch_configuration_limits_t r { cuda::nullopt, cuda::nullopt, cuda::nullopt };
	if (l == nullptr) { return r; }
	if (l->has_block_size) { r.block_size = l->block_size; }
	if (l->has_grid_size) { r.grid_size = l->grid_size; }
	if (l->has_dynamic_shared_memory) { r.dynamic_shared_memory = (cuda::shared_memory_size_t) l->dynamic_shared_memory; }
	return r;
}

const ::std::vector<::std::string>& sorted_keys()
{
	static ::std::vector<::std::string> keys;
	static ::std::once_flag once;
	::std::call_once(once, [] {
		for (const auto& entry : cuda::registered::kernel_t::getSubclassFactory().getInsantiators()) { keys.push_back(entry.first); }
		::std::sort(keys.begin(), keys.end());
	});
	return keys;
}

} // namespace

extern "C" {

size_t giddy_num_registered_kernels(void) { return sorted_keys().size(); }

const char* giddy_registered_kernel_key(size_t index)
{
	const auto& keys = sorted_keys();
	return index < keys.size() ? keys[index].c_str() : nullptr;
}

int giddy_can_produce(const char* key)
{
	if (key == nullptr) { return 0; }
	const auto& keys = sorted_keys();
	return ::std::binary_search(keys.begin(), keys.end(), ::std::string(key)) ? 1 : 0;
}

giddy_status_t giddy_produce_subclass(const char* key, giddy_kernel_t** out)
{
	if (out == nullptr) { return fail(GIDDY_INVALID_ARGUMENT, "null output pointer"); }
	*out = nullptr;
	if (key == nullptr) { return fail(GIDDY_INVALID_ARGUMENT, "null key"); }
	return guarded([&] {
		auto wrapper = cuda::registered::kernel_t::produceSubclass(key);
		*out = new giddy_kernel { ::std::move(wrapper), key };
	});
}

void giddy_kernel_destroy(giddy_kernel_t* kernel) { delete kernel; }

const char* giddy_kernel_key(const giddy_kernel_t* kernel) { return kernel ? kernel->key.c_str() : nullptr; }

const void* giddy_kernel_get_device_function(const giddy_kernel_t* kernel)
{
	if (kernel == nullptr) { return nullptr; }
	const void* result = nullptr;
	guarded([&] { result = kernel->wrapper->get_device_function().ptr(); });
	return result;
}

giddy_status_t giddy_kernel_resolve_launch_configuration(
	const giddy_kernel_t* kernel, const giddy_device_properties_t* device_properties,
	const giddy_function_attributes_t* kernel_function_attributes,
	const giddy_arg_t* extra_arguments, size_t num_extra_arguments,
	const giddy_launch_config_limits_t* limits, giddy_launch_config_t* out)
{
	if (kernel == nullptr || out == nullptr) { return fail(GIDDY_INVALID_ARGUMENT, "null kernel or output pointer"); }
	if (extra_arguments == nullptr && num_extra_arguments != 0) { return fail(GIDDY_INVALID_ARGUMENT, "null argument array"); }
	giddy_device_properties_t current_device;
	if (device_properties == nullptr) {
		giddy_status_t status = giddy_device_properties(-1, &current_device);
		if (status != GIDDY_SUCCESS) { return status; }
		device_properties = &current_device;
	}
	giddy_function_attributes_t queried;
	if (kernel_function_attributes == nullptr) {
		giddy_status_t status = giddy_function_attributes(giddy_kernel_get_device_function(kernel), &queried);
		if (status != GIDDY_SUCCESS) { return status; }
		kernel_function_attributes = &queried;
	}
	return guarded([&] {
		auto config = kernel->wrapper->resolve_launch_configuration(
			to_properties(*device_properties), to_attributes(*kernel_function_attributes),
			to_arguments(extra_arguments, num_extra_arguments), to_limits(limits));
		out->grid_dimensions[0] = config.grid_dimensions.x;   out->grid_dimensions[1] = config.grid_dimensions.y;   out->grid_dimensions[2] = config.grid_dimensions.z;
		out->block_dimensions[0] = config.block_dimensions.x; out->block_dimensions[1] = config.block_dimensions.y; out->block_dimensions[2] = config.block_dimensions.z;
		out->dynamic_shared_memory_size = config.dynamic_shared_memory_size;
	});
}

giddy_status_t giddy_kernel_enqueue_launch(
	const giddy_kernel_t* kernel, void* stream, const giddy_launch_config_t* launch_config,
	const giddy_arg_t* arguments, size_t num_arguments)
{
	if (kernel == nullptr || launch_config == nullptr) { return fail(GIDDY_INVALID_ARGUMENT, "null kernel or launch configuration"); }
	if (arguments == nullptr && num_arguments != 0) { return fail(GIDDY_INVALID_ARGUMENT, "null argument array"); }
	return guarded([&] {
		cuda::launch_configuration_t config;
		config.grid_dimensions  = cuda::dimensions_t(launch_config->grid_dimensions[0], launch_config->grid_dimensions[1], launch_config->grid_dimensions[2]);
		config.block_dimensions = cuda::dimensions_t(launch_config->block_dimensions[0], launch_config->block_dimensions[1], launch_config->block_dimensions[2]);
		config.dynamic_shared_memory_size = (cuda::shared_memory_size_t) launch_config->dynamic_shared_memory_size;
		kernel->wrapper->enqueue_launch(static_cast<cuda::stream::id_t>(stream), config, to_arguments(arguments, num_arguments));
	});
}

} // extern "C"
