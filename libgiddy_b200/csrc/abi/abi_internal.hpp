// Internal helpers shared by the translation units which implement the C ABI (include/giddy_b200.h).
#pragma once
#ifndef GIDDY_B200_ABI_INTERNAL_HPP_
#define GIDDY_B200_ABI_INTERNAL_HPP_

#include "giddy_b200.h"
#include <cuda_runtime.h>
#include <atomic>
#include <string>

namespace giddy {
namespace abi {

// Record a message for giddy_last_error() (thread-local) and return the status.
giddy_status_t fail(giddy_status_t status, const std::string& message);
giddy_status_t fail_cuda(cudaError_t error, const char* what);

// Stream-ordered scratch memory for the duration of a launch, from the library's own pool (abi_common.cu); release with
// cudaFreeAsync on the same stream.
cudaError_t allocate_workspace(void** workspace, size_t bytes, cudaStream_t stream);

inline dim3 to_dim3(const unsigned (&d)[3]) { return dim3(d[0], d[1] ? d[1] : 1, d[2] ? d[2] : 1); }

// A kernel may use 48 KiB of shared memory -- static plus dynamic -- without asking; beyond that it must opt in once per
// function (and per device; the attribute is sticky, so re-applying is harmless).  The static part is a property of the
// kernel (the patched decoders carry 33 KiB of it), asked of the runtime once per kernel and thread.
cudaError_t static_shared_memory_of(const void* kernel, size_t* bytes);

inline cudaError_t ensure_dynamic_shared_memory(const void* kernel, unsigned bytes)
{
	if (bytes == 0) { return cudaSuccess; }
	if (bytes <= 48 * 1024) {
		size_t static_bytes = 0;
		cudaError_t error = static_shared_memory_of(kernel, &static_bytes);
		if (error != cudaSuccess) { return error; }
		if (static_bytes + bytes <= 48 * 1024) { return cudaSuccess; }
	}
	return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) bytes);
}

// Common prologue of every launcher.  Returns true when there is something to launch.
inline bool check_launch(const giddy_launch_config_t* config, const void* kernel, giddy_status_t& status)
{
	if (config == nullptr) { status = fail(GIDDY_INVALID_ARGUMENT, "null launch configuration"); return false; }
	if (kernel == nullptr) { status = fail(GIDDY_UNSUPPORTED, "this combination of type sizes is not instantiated"); return false; }
	status = GIDDY_SUCCESS;
	return true;
}

template <typename Kernel, typename... Args>
inline giddy_status_t launch(Kernel kernel, void* stream, const giddy_launch_config_t* config, Args... args)
{
	cudaError_t error = ensure_dynamic_shared_memory(reinterpret_cast<const void*>(kernel), config->dynamic_shared_memory_size);
	if (error != cudaSuccess) { return fail_cuda(error, "cudaFuncSetAttribute(MaxDynamicSharedMemorySize)"); }
	kernel<<<to_dim3(config->grid_dimensions), to_dim3(config->block_dimensions),
		config->dynamic_shared_memory_size, static_cast<cudaStream_t>(stream)>>>(args...);
	error = cudaGetLastError();
	if (error != cudaSuccess) { return fail_cuda(error, "kernel launch"); }
	return GIDDY_SUCCESS;
}

// The same, as a programmatic dependent of the kernel before it on the stream (kernels/common.cuh): it may be scheduled
// while that kernel drains, and waits for it on the device where it first needs its results.  An ordinary launch where the
// attribute is refused.
template <typename... Params, typename... Args>
inline giddy_status_t launch_as_dependent(void (*kernel)(Params...), void* stream, const giddy_launch_config_t* config, Args... args)
{
	cudaError_t error = ensure_dynamic_shared_memory(reinterpret_cast<const void*>(kernel), config->dynamic_shared_memory_size);
	if (error != cudaSuccess) { return fail_cuda(error, "cudaFuncSetAttribute(MaxDynamicSharedMemorySize)"); }
	cudaLaunchConfig_t launch_config = { };
	launch_config.gridDim = to_dim3(config->grid_dimensions);
	launch_config.blockDim = to_dim3(config->block_dimensions);
	launch_config.dynamicSmemBytes = config->dynamic_shared_memory_size;
	launch_config.stream = static_cast<cudaStream_t>(stream);
	cudaLaunchAttribute attribute = { };
	attribute.id = cudaLaunchAttributeProgrammaticStreamSerialization;
	attribute.val.programmaticStreamSerializationAllowed = 1;
	launch_config.attrs = &attribute;
	launch_config.numAttrs = 1;
	error = cudaLaunchKernelEx(&launch_config, kernel, static_cast<Params>(args)...);
	if (error == cudaSuccess) { return GIDDY_SUCCESS; }
	cudaGetLastError();
	return launch(kernel, stream, config, args...);
}

} // namespace abi
} // namespace giddy

#endif // GIDDY_B200_ABI_INTERNAL_HPP_
