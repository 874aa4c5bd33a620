// giddy/multi_gpu.hpp -- C++ face of the shard planner, the pinned host allocator and the thread-per-GPU device
// groups (include/giddy_b200.h, groups 5-7).
//
// The reference is single-GPU (src/kernel_wrappers/registered_wrapper.h:82: "the stream is assumed to belong to the
// current device"); what lets a column spread over the GPUs of a box is the reference's own formats, whose anchored
// segments decode independently.  A caller
//
//   cuda::device_group_t gpus;                                           // one worker thread per visible GPU
//   gpus.run([&](unsigned rank, unsigned size, int device) {
//       auto range = cuda::sharding::split_segments(num_segments, size, rank);
//       auto shard = cuda::sharding::plan(layout, range.first, range.second);   // which runs / bytes / anchors it reads
//       for (unsigned k = 0; k < shard.num_slices; k++) { ... giddy_slice_copy into cuda::host_buffer_t staging, H2D ... }
//       kernel->enqueue_launch(stream, config, {...shard scalars...});          // exactly as for a whole column
//   });
//
// examples/sharded_scan.cpp is this, complete.
#pragma once
#ifndef GIDDY_MULTI_GPU_HPP_
#define GIDDY_MULTI_GPU_HPP_

#include "kernel_wrapper.hpp"

#include <functional>
#include <utility>
#include <vector>

namespace cuda {

namespace sharding {

/// Segments [first, second) of part `part` of `parts` nearly equal contiguous parts
inline ::std::pair<uint64_t, uint64_t> split_segments(uint64_t num_segments, unsigned parts, unsigned part)
{
	::std::pair<uint64_t, uint64_t> range;
	detail::throw_if_error(giddy_split_segments(num_segments, parts, part, &range.first, &range.second));
	return range;
}

/// The slices of the column's arrays which segments [first_segment, past_last_segment) read, and the shard's scalars
inline giddy_shard_t plan(const giddy_column_layout_t& column, uint64_t first_segment, uint64_t past_last_segment)
{
	giddy_shard_t shard;
	detail::throw_if_error(giddy_shard_plan(&column, first_segment, past_last_segment, &shard));
	return shard;
}

inline const giddy_slice_t& slice_of(const giddy_shard_t& shard, const ::std::string& array)
{
	for (unsigned k = 0; k < shard.num_slices; k++) { if (array == shard.slices[k].array) { return shard.slices[k]; } }
	throw ::std::out_of_range("the shard has no slice of " + array);
}

/// Host-to-host copy of a slice out of the whole column's array, anchors rebased on the way
inline void copy_slice(const giddy_slice_t& slice, void* destination, const void* whole_array)
{
	detail::throw_if_error(giddy_slice_copy(&slice, destination, whole_array));
}

} // namespace sharding

/// Page-locked host memory on the NUMA node of a device (giddy_host_alloc)
template <typename T>
class host_buffer_t {
public:
	host_buffer_t() = default;
	explicit host_buffer_t(::std::size_t count, int device = -1) : count_(count)
	{
		void* memory = nullptr;
		detail::throw_if_error(giddy_host_alloc(count * sizeof(T), device, &memory, &numa_node_));
		data_ = static_cast<T*>(memory);
	}
	~host_buffer_t() { giddy_host_free(data_); }
	host_buffer_t(const host_buffer_t&) = delete;
	host_buffer_t& operator=(const host_buffer_t&) = delete;
	host_buffer_t(host_buffer_t&& other) noexcept : data_(other.data_), count_(other.count_), numa_node_(other.numa_node_) { other.data_ = nullptr; other.count_ = 0; }
	host_buffer_t& operator=(host_buffer_t&& other) noexcept
	{
		if (this != &other) { giddy_host_free(data_); data_ = other.data_; count_ = other.count_; numa_node_ = other.numa_node_; other.data_ = nullptr; other.count_ = 0; }
		return *this;
	}
	T* data() { return data_; }
	const T* data() const { return data_; }
	::std::size_t size() const { return count_; }
	T& operator[](::std::size_t i) { return data_[i]; }
	const T& operator[](::std::size_t i) const { return data_[i]; }
	int numa_node() const { return numa_node_; }
private:
	T* data_ = nullptr;
	::std::size_t count_ = 0;
	int numa_node_ = -1;
};

/// One worker thread per GPU (giddy_device_group_*)
class device_group_t {
public:
	/// all visible devices
	device_group_t() { detail::throw_if_error(giddy_device_group_create(nullptr, 0, &handle_)); }
	explicit device_group_t(const ::std::vector<int>& devices) { detail::throw_if_error(giddy_device_group_create(devices.data(), (unsigned) devices.size(), &handle_)); }
	~device_group_t() { giddy_device_group_destroy(handle_); }
	device_group_t(const device_group_t&) = delete;
	device_group_t& operator=(const device_group_t&) = delete;

	unsigned size() const { return giddy_device_group_size(handle_); }
	int device(unsigned rank) const { return giddy_device_group_device(handle_, rank); }

	/// work(rank, size, device) on every worker at once; an exception thrown by any of them is rethrown here
	void run(::std::function<void(unsigned, unsigned, int)> work)
	{
		call_t call { ::std::move(work), { } };
		giddy_status_t status = giddy_device_group_run(handle_, &device_group_t::trampoline, &call);
		if (!call.what.empty()) { throw ::std::runtime_error(call.what); }
		detail::throw_if_error(status);
	}

	/// ncclBroadcast of `bytes` bytes from buffers[root] to every other rank's buffer (asynchronous on the given streams)
	void broadcast(const ::std::vector<void*>& buffers, ::std::size_t bytes, unsigned root = 0, const ::std::vector<stream::id_t>& streams = { })
	{
		::std::vector<void*> handles(streams.begin(), streams.end());
		detail::throw_if_error(giddy_device_group_broadcast(handle_, buffers.data(), bytes, root, handles.empty() ? nullptr : handles.data()));
	}

private:
	struct call_t { ::std::function<void(unsigned, unsigned, int)> work; ::std::string what; };
	static giddy_status_t trampoline(unsigned rank, unsigned size, int device, void* user_data)
	{
		auto* call = static_cast<call_t*>(user_data);
		try { call->work(rank, size, device); return GIDDY_SUCCESS; }
		catch (const ::std::exception& e) { call->what = e.what(); return GIDDY_INTERNAL_ERROR; } // (first writer wins is good enough for a message)
	}
	giddy_device_group_t* handle_ = nullptr;
};

} // namespace cuda

#endif // GIDDY_MULTI_GPU_HPP_
