// C ABI: error reporting, device / function queries, launch-by-pointer.
#include "abi_internal.hpp"
#include <cstring>
#include <map>
#include <mutex>

namespace giddy {
namespace abi {

static thread_local std::string last_error_message;

giddy_status_t fail(giddy_status_t status, const std::string& message)
{
	last_error_message = message;
	return status;
}

giddy_status_t fail_cuda(cudaError_t error, const char* what)
{
	return fail(GIDDY_CUDA_ERROR, std::string(what) + ": " + cudaGetErrorName(error) + " (" + cudaGetErrorString(error) + ")");
}

// Workspaces (the Delta look-back's tile descriptors, the patched decoders' tile index) come from a pool of this
// library's own, one per device, which keeps what it has been given (release threshold: never): the default pool hands
// its memory back to the driver at every synchronisation, and the next launch then pays for mapping it again.
static cudaError_t workspace_pool(cudaMemPool_t* pool)
{
	static std::mutex guard;
	static cudaMemPool_t pools[64] = { };
	int device = 0;
	cudaError_t error = cudaGetDevice(&device);
	if (error != cudaSuccess) { return error; }
	if (device < 0 || device >= 64) { return cudaDeviceGetDefaultMemPool(pool, device); }
	std::lock_guard<std::mutex> lock(guard);
	if (pools[device] == nullptr) {
		cudaMemPoolProps properties = { };
		properties.allocType = cudaMemAllocationTypePinned;
		properties.handleTypes = cudaMemHandleTypeNone;
		properties.location.type = cudaMemLocationTypeDevice;
		properties.location.id = device;
		error = cudaMemPoolCreate(&pools[device], &properties);
		if (error != cudaSuccess) { pools[device] = nullptr; return error; }
		unsigned long long keep = ~0ull;
		cudaMemPoolSetAttribute(pools[device], cudaMemPoolAttrReleaseThreshold, &keep);
	}
	*pool = pools[device];
	return cudaSuccess;
}

cudaError_t static_shared_memory_of(const void* kernel, size_t* bytes)
{
	static thread_local std::map<const void*, size_t> known; // (a kernel's static shared memory is the same on every device)
	auto it = known.find(kernel);
	if (it == known.end()) {
		cudaFuncAttributes attributes;
		cudaError_t error = cudaFuncGetAttributes(&attributes, kernel);
		if (error != cudaSuccess) { return error; }
		it = known.emplace(kernel, attributes.sharedSizeBytes).first;
	}
	*bytes = it->second;
	return cudaSuccess;
}

cudaError_t allocate_workspace(void** workspace, size_t bytes, cudaStream_t stream)
{
	cudaMemPool_t pool = nullptr;
	cudaError_t error = workspace_pool(&pool);
	if (error != cudaSuccess) { return error; }
	return cudaMallocFromPoolAsync(workspace, bytes, pool, stream);
}

} // namespace abi
} // namespace giddy

using namespace giddy::abi;

extern "C" {

const char* giddy_last_error(void) { return last_error_message.c_str(); }

const char* giddy_version(void) { return "giddy_b200 0.1 (sm_100a)"; }

giddy_status_t giddy_device_properties(int device_id, giddy_device_properties_t* out)
{
	if (out == nullptr) { return fail(GIDDY_INVALID_ARGUMENT, "null output pointer"); }
	cudaError_t error;
	if (device_id < 0) {
		error = cudaGetDevice(&device_id);
		if (error != cudaSuccess) { return fail_cuda(error, "cudaGetDevice"); }
	}
	// cudaGetDeviceProperties costs milliseconds; a device's properties do not change, so ask once
	static std::mutex cache_mutex;
	static std::map<int, giddy_device_properties_t> cache;
	{
		std::lock_guard<std::mutex> lock(cache_mutex);
		auto it = cache.find(device_id);
		if (it != cache.end()) { *out = it->second; return GIDDY_SUCCESS; }
	}
	cudaDeviceProp p;
	error = cudaGetDeviceProperties(&p, device_id);
	if (error != cudaSuccess) { return fail_cuda(error, "cudaGetDeviceProperties"); }
	std::memset(out, 0, sizeof(*out));
	out->compute_capability_major = p.major;
	out->compute_capability_minor = p.minor;
	out->multiprocessor_count = p.multiProcessorCount;
	out->max_threads_per_block = p.maxThreadsPerBlock;
	out->max_threads_dim_x = p.maxThreadsDim[0];
	out->max_threads_per_multiprocessor = p.maxThreadsPerMultiProcessor;
	out->max_blocks_per_multiprocessor = p.maxBlocksPerMultiProcessor;
	out->warp_size = p.warpSize;
	out->shared_memory_per_block = p.sharedMemPerBlock;
	out->shared_memory_per_block_optin = p.sharedMemPerBlockOptin;
	out->shared_memory_per_multiprocessor = p.sharedMemPerMultiprocessor;
	out->l2_cache_size = (size_t) p.l2CacheSize;
	out->total_global_memory = p.totalGlobalMem;
	{
		std::lock_guard<std::mutex> lock(cache_mutex);
		cache[device_id] = *out;
	}
	return GIDDY_SUCCESS;
}

giddy_status_t giddy_function_attributes(const void* kernel, giddy_function_attributes_t* out)
{
	if (kernel == nullptr || out == nullptr) { return fail(GIDDY_INVALID_ARGUMENT, "null kernel or output pointer"); }
	cudaFuncAttributes a;
	cudaError_t error = cudaFuncGetAttributes(&a, kernel);
	if (error != cudaSuccess) { return fail_cuda(error, "cudaFuncGetAttributes"); }
	out->max_threads_per_block = a.maxThreadsPerBlock;
	out->num_regs = a.numRegs;
	out->shared_size_bytes = a.sharedSizeBytes;
	out->const_size_bytes = a.constSizeBytes;
	out->local_size_bytes = a.localSizeBytes;
	out->max_dynamic_shared_size_bytes = a.maxDynamicSharedSizeBytes;
	out->ptx_version = a.ptxVersion;
	out->binary_version = a.binaryVersion;
	return GIDDY_SUCCESS;
}

giddy_status_t giddy_launch_kernel(const void* kernel, void* stream, const giddy_launch_config_t* config, void** args)
{
	if (kernel == nullptr || config == nullptr) { return fail(GIDDY_INVALID_ARGUMENT, "null kernel or launch configuration"); }
	cudaError_t error = ensure_dynamic_shared_memory(kernel, config->dynamic_shared_memory_size);
	if (error != cudaSuccess) { return fail_cuda(error, "cudaFuncSetAttribute(MaxDynamicSharedMemorySize)"); }
	error = cudaLaunchKernel(kernel, to_dim3(config->grid_dimensions), to_dim3(config->block_dimensions), args,
		config->dynamic_shared_memory_size, static_cast<cudaStream_t>(stream));
	if (error != cudaSuccess) { return fail_cuda(error, "cudaLaunchKernel"); }
	return GIDDY_SUCCESS;
}

} // extern "C"
