// giddy/host_resident_scan.hpp -- C++ face of the host-resident scan executor (include/giddy_b200.h, group 4).
//
// The reference stops at the kernel launch; what surrounds it in its intended use -- columns kept compressed in host
// memory, shipped over the bus compressed and decompressed on the GPU (README.md:18-24) -- is left to the calling
// DBMS.  This is that caller-side piece as a library facility: a plan of (pinned host -> device copy, decode launch)
// operations per chunk of segments, replayed on side streams so that chunk c+1's copy overlaps chunk c's decode.
//
//   cuda::host_resident::scan_t scan(4);                                      // 4 side streams
//   for (unsigned c = 0; c < num_chunks; c++) {
//       scan.add_copy(c, device_offsets + e0, pinned_offsets + e0, (e1 - e0) * sizeof(short));
//       scan.add_launch(c, *kernel, config, {{"decompressed", out + e0}, ...});   // a cuda::kernel_t from the factory
//   }
//   scan.run(stream);                                                         // asynchronous; replay at will
#pragma once
#ifndef GIDDY_HOST_RESIDENT_SCAN_HPP_
#define GIDDY_HOST_RESIDENT_SCAN_HPP_

#include "kernel_wrapper.hpp"

#include <exception>
#include <functional>
#include <memory>

namespace cuda {
namespace host_resident {

class scan_t {
public:
	explicit scan_t(unsigned num_streams = 4) { detail::throw_if_error(giddy_scan_create(num_streams, &handle_)); }
	~scan_t() { giddy_scan_destroy(handle_); }
	scan_t(const scan_t&) = delete;
	scan_t& operator=(const scan_t&) = delete;
	scan_t(scan_t&& other) noexcept : handle_(other.handle_) { other.handle_ = nullptr; }

	unsigned num_streams() const { return giddy_scan_num_streams(handle_); }
	unsigned num_chunks() const { return giddy_scan_num_chunks(handle_); }
	::std::size_t num_operations() const { return giddy_scan_num_operations(handle_); }

	/// cudaMemcpyAsync host -> device on the chunk's stream; the source should be pinned
	void add_copy(unsigned chunk, void* device_destination, const void* host_source, ::std::size_t bytes)
	{
		detail::throw_if_error(giddy_scan_add_copy(handle_, chunk, device_destination, host_source, bytes));
	}

	/// cudaMemcpyAsync device -> host on the chunk's stream, for (part of) a chunk's result
	void add_copy_back(unsigned chunk, void* host_destination, const void* device_source, ::std::size_t bytes)
	{
		detail::throw_if_error(giddy_scan_add_copy_back(handle_, chunk, host_destination, device_source, bytes));
	}

	/// kernel.enqueue_launch(<the chunk's stream>, launch_config, arguments) at every replay; the kernel must outlive the scan
	void add_launch(unsigned chunk, const kernel_t& kernel, const launch_configuration_t& launch_config, kernel_t::arguments_type arguments)
	{
		const kernel_t* k = &kernel;
		add(chunk, [k, launch_config, arguments](stream::id_t stream) { k->enqueue_launch(stream, launch_config, arguments); });
	}

	/// Anything else: `enqueue(stream)` must only enqueue asynchronous work on the stream it is given
	void add(unsigned chunk, ::std::function<void(stream::id_t)> enqueue)
	{
		auto* boxed = new ::std::function<void(stream::id_t)>(::std::move(enqueue));
		giddy_status_t status = giddy_scan_add_callback(handle_, chunk, &scan_t::trampoline, boxed, &scan_t::destroy);
		if (status != GIDDY_SUCCESS) { delete boxed; detail::throw_if_error(status); }
	}

	/// What is added to `chunk` from now on waits for what has been added to `on_chunk` (< chunk) so far
	void add_dependency(unsigned chunk, unsigned on_chunk) { detail::throw_if_error(giddy_scan_add_dependency(handle_, chunk, on_chunk)); }

	/// Replay as one captured CUDA graph instead of re-enqueueing every operation
	void set_graph_replay(bool enabled) { detail::throw_if_error(giddy_scan_set_graph_replay(handle_, enabled ? 1 : 0)); }

	/// Fork the side streams from `stream`, run the chunks, join back into `stream`.  Asynchronous.
	void run(stream::id_t stream = stream::default_stream_id) { detail::throw_if_error(giddy_scan_run(handle_, stream)); }

private:
	static giddy_status_t trampoline(void* stream, void* user_data)
	{
		try { (*static_cast<::std::function<void(stream::id_t)>*>(user_data))(static_cast<stream::id_t>(stream)); return GIDDY_SUCCESS; }
		catch (const ::std::invalid_argument&) { return GIDDY_INVALID_ARGUMENT; }
		catch (const ::std::out_of_range&) { return GIDDY_OUT_OF_RANGE; }
		catch (const ::std::exception&) { return GIDDY_CUDA_ERROR; }
	}
	static void destroy(void* user_data) { delete static_cast<::std::function<void(stream::id_t)>*>(user_data); }

	giddy_scan_t* handle_ = nullptr;
};

} // namespace host_resident
} // namespace cuda

#endif // GIDDY_HOST_RESIDENT_SCAN_HPP_
