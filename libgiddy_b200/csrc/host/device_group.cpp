// Device groups (include/giddy_b200.h, group 7): the multi-GPU driver of the decompression path.
//
// Decoding shards without communication (csrc/host/shard_plan.cpp), so the driver is plain: one worker thread per
// GPU, each with its device current and bound to the CPUs of the device's NUMA node, all running the caller's work
// function on their own shard at the same time.  The one exchange the path has is the broadcast of a dictionary from
// the device which loaded it to the others: NCCL over NVLink / NVSwitch.  NCCL is loaded at first use with dlopen
// (libnccl.so.2 -- the copy already in the process, when the caller is a framework which brings its own) and never
// linked, so the library has no NCCL dependency for single-GPU users.  The reference is single-GPU
// (src/kernel_wrappers/registered_wrapper.h:82): no counterpart there.
#include "giddy_b200.h"

#include <cuda_runtime.h>
#include <dlfcn.h>

#include <condition_variable>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace giddy { namespace abi {
giddy_status_t fail(giddy_status_t status, const ::std::string& message);
giddy_status_t fail_cuda(cudaError_t error, const char* what);
} }

namespace {

using giddy::abi::fail;
using giddy::abi::fail_cuda;

// ---- NCCL, by name ------------------------------------------------------------------------------------

struct nccl_t {
	using comm_t = void*;
	int (*CommInitAll)(comm_t*, int, const int*) = nullptr;
	int (*CommDestroy)(comm_t) = nullptr;
	int (*Broadcast)(const void*, void*, size_t, int /* ncclDataType_t */, int, comm_t, cudaStream_t) = nullptr;
	int (*GroupStart)() = nullptr;
	int (*GroupEnd)() = nullptr;
	const char* (*GetErrorString)(int) = nullptr;
	bool loaded = false;
	::std::string why_not;
};

const nccl_t& nccl()
{
	static nccl_t api;
	static ::std::once_flag once;
	::std::call_once(once, [] {
		void* handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
		if (handle == nullptr) { handle = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL); }
		if (handle == nullptr) { api.why_not = ::std::string("libnccl.so.2 could not be loaded: ") + dlerror(); return; }
		api.CommInitAll = reinterpret_cast<decltype(api.CommInitAll)>(dlsym(handle, "ncclCommInitAll"));
		api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(handle, "ncclCommDestroy"));
		api.Broadcast = reinterpret_cast<decltype(api.Broadcast)>(dlsym(handle, "ncclBroadcast"));
		api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(dlsym(handle, "ncclGroupStart"));
		api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(dlsym(handle, "ncclGroupEnd"));
		api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(handle, "ncclGetErrorString"));
		api.loaded = api.CommInitAll && api.CommDestroy && api.Broadcast && api.GroupStart && api.GroupEnd && api.GetErrorString;
		if (!api.loaded) { api.why_not = "libnccl.so.2 lacks one of the entry points the dictionary broadcast needs"; }
	});
	return api;
}

constexpr int nccl_uint8 = 1; // ncclUint8 (ncclChar = ncclInt8 = 0), nccl.h

giddy_status_t fail_nccl(int result, const char* what)
{
	return fail(GIDDY_CUDA_ERROR, ::std::string(what) + ": " + nccl().GetErrorString(result));
}

} // namespace

// ---- the group -------------------------------------------------------------------------------------------

struct giddy_device_group {
	struct worker_t {
		int device = 0;
		::std::thread thread;
		giddy_status_t status = GIDDY_SUCCESS;
		::std::string message;
	};
	::std::vector<::std::unique_ptr<worker_t>> workers;
	::std::vector<void*> communicators;

	::std::mutex mutex;
	::std::condition_variable wake, done;
	uint64_t generation = 0;     // bumped for every round of work
	unsigned remaining = 0;      // workers which have not finished the current round
	bool stopping = false;
	giddy_device_work_t work = nullptr;
	void* user_data = nullptr;

	void loop(unsigned rank)
	{
		worker_t& me = *workers[rank];
		cudaSetDevice(me.device);
		giddy_host_bind_thread_to_device(me.device);
		uint64_t seen = 0;
		while (true) {
			giddy_device_work_t todo; void* data;
			{
				::std::unique_lock<::std::mutex> lock(mutex);
				wake.wait(lock, [&] { return stopping || generation != seen; });
				if (stopping) { return; }
				seen = generation; todo = work; data = user_data;
			}
			me.status = todo(rank, (unsigned) workers.size(), me.device, data);
			me.message = (me.status != GIDDY_SUCCESS) ? giddy_last_error() : ""; // (the message is thread-local: carry it over)
			{
				::std::lock_guard<::std::mutex> lock(mutex);
				if (--remaining == 0) { done.notify_all(); }
			}
		}
	}

	~giddy_device_group()
	{
		{
			::std::lock_guard<::std::mutex> lock(mutex);
			stopping = true;
		}
		wake.notify_all();
		for (auto& w : workers) { if (w->thread.joinable()) { w->thread.join(); } }
		for (void* c : communicators) { if (c) { nccl().CommDestroy(c); } }
	}
};

extern "C" {

giddy_status_t giddy_device_group_create(const int* devices, unsigned num_devices, giddy_device_group_t** out)
{
	if (out == nullptr) { return fail(GIDDY_INVALID_ARGUMENT, "giddy_device_group_create: null output pointer"); }
	*out = nullptr;
	int visible = 0;
	cudaError_t error = cudaGetDeviceCount(&visible);
	if (error != cudaSuccess) { return fail_cuda(error, "cudaGetDeviceCount"); }
	::std::vector<int> chosen;
	if (devices == nullptr) { for (int d = 0; d < visible; d++) { chosen.push_back(d); } }
	else { chosen.assign(devices, devices + num_devices); }
	if (chosen.empty()) { return fail(GIDDY_INVALID_ARGUMENT, "a device group needs at least one device"); }
	for (size_t i = 0; i < chosen.size(); i++) {
		if (chosen[i] < 0 || chosen[i] >= visible) { return fail(GIDDY_INVALID_ARGUMENT, "giddy_device_group_create: no such device"); }
		for (size_t j = 0; j < i; j++) { if (chosen[j] == chosen[i]) { return fail(GIDDY_INVALID_ARGUMENT, "giddy_device_group_create: a device is listed twice"); } }
	}
	::std::unique_ptr<giddy_device_group> group(new giddy_device_group);
	for (int device : chosen) {
		group->workers.emplace_back(new giddy_device_group::worker_t);
		group->workers.back()->device = device;
	}
	for (unsigned rank = 0; rank < group->workers.size(); rank++) {
		giddy_device_group* g = group.get();
		group->workers[rank]->thread = ::std::thread([g, rank] { g->loop(rank); });
	}
	*out = group.release();
	return GIDDY_SUCCESS;
}

void giddy_device_group_destroy(giddy_device_group_t* group) { delete group; }

unsigned giddy_device_group_size(const giddy_device_group_t* group) { return group ? (unsigned) group->workers.size() : 0; }

int giddy_device_group_device(const giddy_device_group_t* group, unsigned rank)
{
	return (group && rank < group->workers.size()) ? group->workers[rank]->device : -1;
}

giddy_status_t giddy_device_group_run(giddy_device_group_t* group, giddy_device_work_t work, void* user_data)
{
	if (group == nullptr || work == nullptr) { return fail(GIDDY_INVALID_ARGUMENT, "giddy_device_group_run: null group or work function"); }
	{
		::std::unique_lock<::std::mutex> lock(group->mutex);
		if (group->remaining != 0) { return fail(GIDDY_LOGIC_ERROR, "giddy_device_group_run: the group is already running work"); }
		group->work = work; group->user_data = user_data;
		group->remaining = (unsigned) group->workers.size();
		group->generation++;
		group->wake.notify_all();
		group->done.wait(lock, [&] { return group->remaining == 0; });
	}
	for (auto& w : group->workers) {
		if (w->status != GIDDY_SUCCESS) { return fail(w->status, "device " + ::std::to_string(w->device) + ": " + w->message); }
	}
	return GIDDY_SUCCESS;
}

giddy_status_t giddy_device_group_broadcast(giddy_device_group_t* group, void* const* buffers, size_t bytes, unsigned root, void* const* streams)
{
	if (group == nullptr || buffers == nullptr) { return fail(GIDDY_INVALID_ARGUMENT, "giddy_device_group_broadcast: null group or buffers"); }
	const unsigned size = (unsigned) group->workers.size();
	if (root >= size) { return fail(GIDDY_INVALID_ARGUMENT, "giddy_device_group_broadcast: no such root"); }
	if (bytes == 0 || size == 1) { return GIDDY_SUCCESS; }
	const nccl_t& api = nccl();
	if (!api.loaded) { return fail(GIDDY_UNSUPPORTED, api.why_not); }
	int previous = 0;
	cudaGetDevice(&previous);
	if (group->communicators.empty()) {
		::std::vector<int> devices;
		for (auto& w : group->workers) { devices.push_back(w->device); }
		group->communicators.assign(size, nullptr);
		const int result = api.CommInitAll(group->communicators.data(), (int) size, devices.data());
		if (result != 0) { group->communicators.clear(); cudaSetDevice(previous); return fail_nccl(result, "ncclCommInitAll"); }
	}
	int result = api.GroupStart();
	for (unsigned r = 0; r < size && result == 0; r++) {
		cudaSetDevice(group->workers[r]->device);
		result = api.Broadcast(buffers[r], buffers[r], bytes, nccl_uint8, (int) root, group->communicators[r],
			static_cast<cudaStream_t>(streams ? streams[r] : nullptr));
	}
	const int end = api.GroupEnd();
	cudaSetDevice(previous);
	if (result != 0) { return fail_nccl(result, "ncclBroadcast"); }
	if (end != 0) { return fail_nccl(end, "ncclGroupEnd"); }
	return GIDDY_SUCCESS;
}

giddy_status_t giddy_broadcast_dictionary(void* nccl_communicator, void* buffer, size_t bytes, int root, void* stream)
{
	if (nccl_communicator == nullptr || (buffer == nullptr && bytes != 0)) { return fail(GIDDY_INVALID_ARGUMENT, "giddy_broadcast_dictionary: null communicator or buffer"); }
	if (bytes == 0) { return GIDDY_SUCCESS; }
	const nccl_t& api = nccl();
	if (!api.loaded) { return fail(GIDDY_UNSUPPORTED, api.why_not); }
	const int result = api.Broadcast(buffer, buffer, bytes, nccl_uint8, root, nccl_communicator, static_cast<cudaStream_t>(stream));
	if (result != 0) { return fail_nccl(result, "ncclBroadcast"); }
	return GIDDY_SUCCESS;
}

} // extern "C"
