// Host memory for host-resident columns (include/giddy_b200.h, group 6).
//
// A compressed column which lives in host memory reaches the GPU through cudaMemcpyAsync from page-locked pages.  On a
// multi-socket box the pages should sit on the NUMA node the GPU's PCIe root hangs off -- otherwise eight GPUs pull
// their chunks through one socket's memory controllers and the inter-socket link.  giddy_host_alloc therefore maps
// anonymous memory, binds it to the device's node (the mbind system call: libnuma is not required), touches it, and
// registers it with CUDA.  Where the platform reports no node (sysfs says -1: single-node hosts, most containers) the
// binding is skipped and the memory is plain page-locked memory.  No counterpart in the reference (its examples use
// cudaMallocHost through cuda-api-wrappers, examples/src/modes_of_use/common.cuh:60-75).
#include "giddy_b200.h"

#include <cuda_runtime.h>

#include <sched.h>
#include <sys/mman.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>

namespace giddy { namespace abi {
giddy_status_t fail(giddy_status_t status, const ::std::string& message);
giddy_status_t fail_cuda(cudaError_t error, const char* what);
} }

namespace {

using giddy::abi::fail;
using giddy::abi::fail_cuda;

struct allocation_t { size_t bytes; bool mapped; };
::std::mutex allocations_mutex;
::std::map<void*, allocation_t> allocations;

int resolve_device(int device)
{
	if (device >= 0) { return device; }
	int current = 0;
	return cudaGetDevice(&current) == cudaSuccess ? current : -1;
}

// "0-15,32-47" -> a cpu_set_t; false if nothing could be parsed
bool parse_cpu_list(const char* text, cpu_set_t* set)
{
	CPU_ZERO(set);
	bool any = false;
	const char* p = text;
	while (*p) {
		char* end = nullptr;
		long low = ::std::strtol(p, &end, 10);
		if (end == p) { break; }
		long high = low;
		if (*end == '-') { p = end + 1; high = ::std::strtol(p, &end, 10); }
		for (long c = low; c <= high && c < CPU_SETSIZE; c++) { CPU_SET((int) c, set); any = true; }
		p = end;
		while (*p == ',' || *p == '\n' || *p == ' ') { p++; }
	}
	return any;
}

} // namespace

extern "C" {

int giddy_host_numa_node_of_device(int device)
{
	device = resolve_device(device);
	if (device < 0) { return -1; }
	char bus_id[64] = { };
	if (cudaDeviceGetPCIBusId(bus_id, sizeof(bus_id), device) != cudaSuccess) { cudaGetLastError(); return -1; }
	for (char* c = bus_id; *c; c++) { if (*c >= 'A' && *c <= 'F') { *c = (char)(*c - 'A' + 'a'); } } // sysfs spells addresses in lower case
	const ::std::string path = ::std::string("/sys/bus/pci/devices/") + bus_id + "/numa_node";
	FILE* f = ::std::fopen(path.c_str(), "r");
	if (f == nullptr) { return -1; }
	int node = -1;
	if (::std::fscanf(f, "%d", &node) != 1) { node = -1; }
	::std::fclose(f);
	return node;
}

int giddy_host_bind_thread_to_device(int device)
{
	const int node = giddy_host_numa_node_of_device(device);
	if (node < 0) { return -1; }
	char path[96];
	::std::snprintf(path, sizeof(path), "/sys/devices/system/node/node%d/cpulist", node);
	FILE* f = ::std::fopen(path, "r");
	if (f == nullptr) { return -1; }
	char text[4096] = { };
	const size_t got = ::std::fread(text, 1, sizeof(text) - 1, f);
	::std::fclose(f);
	cpu_set_t wanted, allowed;
	if (got == 0 || !parse_cpu_list(text, &wanted)) { return -1; }
	if (sched_getaffinity(0, sizeof(allowed), &allowed) != 0) { return -1; }
	CPU_AND(&wanted, &wanted, &allowed);
	if (CPU_COUNT(&wanted) == 0) { return -1; }
	return sched_setaffinity(0, sizeof(wanted), &wanted) == 0 ? node : -1;
}

giddy_status_t giddy_host_alloc(size_t bytes, int device, void** out, int* numa_node)
{
	if (out == nullptr) { return fail(GIDDY_INVALID_ARGUMENT, "giddy_host_alloc: null output pointer"); }
	*out = nullptr;
	if (numa_node) { *numa_node = -1; }
	if (bytes == 0) { return GIDDY_SUCCESS; }
	const int node = giddy_host_numa_node_of_device(device);
	const size_t page = (size_t) sysconf(_SC_PAGESIZE);
	const size_t mapped_bytes = (bytes + page - 1) / page * page;
	void* memory = mmap(nullptr, mapped_bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
	if (memory != MAP_FAILED) {
		int bound = -1;
#ifdef SYS_mbind
		if (node >= 0 && node < 1024) {
			unsigned long mask[1024 / (8 * sizeof(unsigned long))] = { };
			mask[node / (8 * sizeof(unsigned long))] = 1ul << (node % (8 * sizeof(unsigned long)));
			const int MPOL_PREFERRED_ = 1; // (preferred, not strict: a full node must not fail the allocation)
			if (syscall(SYS_mbind, memory, mapped_bytes, MPOL_PREFERRED_, mask, (unsigned long) 1024, 0u) == 0) { bound = node; }
		}
#endif
		::std::memset(memory, 0, mapped_bytes); // first touch: the pages now exist, on the preferred node
		cudaError_t error = cudaHostRegister(memory, mapped_bytes, cudaHostRegisterPortable);
		if (error == cudaSuccess) {
			::std::lock_guard<::std::mutex> lock(allocations_mutex);
			allocations[memory] = { mapped_bytes, true };
			*out = memory;
			if (numa_node) { *numa_node = bound; }
			return GIDDY_SUCCESS;
		}
		cudaGetLastError();
		munmap(memory, mapped_bytes); // (registration can be refused -- locked-memory limits --: fall back to the runtime's own allocator)
	}
	void* pinned = nullptr;
	cudaError_t error = cudaHostAlloc(&pinned, bytes, cudaHostAllocPortable);
	if (error != cudaSuccess) { return fail_cuda(error, "cudaHostAlloc"); }
	{
		::std::lock_guard<::std::mutex> lock(allocations_mutex);
		allocations[pinned] = { bytes, false };
	}
	*out = pinned;
	return GIDDY_SUCCESS;
}

giddy_status_t giddy_host_free(void* memory)
{
	if (memory == nullptr) { return GIDDY_SUCCESS; }
	allocation_t allocation;
	{
		::std::lock_guard<::std::mutex> lock(allocations_mutex);
		auto it = allocations.find(memory);
		if (it == allocations.end()) { return fail(GIDDY_INVALID_ARGUMENT, "giddy_host_free: not a giddy_host_alloc allocation"); }
		allocation = it->second;
		allocations.erase(it);
	}
	if (allocation.mapped) {
		cudaError_t error = cudaHostUnregister(memory);
		munmap(memory, allocation.bytes);
		if (error != cudaSuccess) { return fail_cuda(error, "cudaHostUnregister"); }
		return GIDDY_SUCCESS;
	}
	cudaError_t error = cudaFreeHost(memory);
	if (error != cudaSuccess) { return fail_cuda(error, "cudaFreeHost"); }
	return GIDDY_SUCCESS;
}

} // extern "C"
