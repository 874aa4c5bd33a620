// Host-resident scans (include/giddy_b200.h, group 4): a replayable plan which pipelines pinned host -> device
// copies of compressed chunks with their decode launches on side streams.
//
// This is the caller-side piece the reference leaves to the DBMS ("the data is transferred over PCIe compressed and
// decompressed on the GPU", README.md:18-24): it owns no column data, only streams, events and the recorded
// operations.  A plan is built once per column set -- chunk by chunk: copies, kernel launches through the registry
// API, cross-chunk dependencies -- and then replayed per scan, either by re-enqueueing every operation or, if asked,
// as one CUDA graph captured from the first replay.  Chunk c runs on side stream c % num_streams, so that the copy of
// chunk c+1 overlaps the decode of chunk c; all side streams are forked from, and joined back into, the caller's stream.
//
// Written against the C ABI only (giddy_kernel_enqueue_launch), not the wrapper classes.
#include "giddy_b200.h"

#include <cuda_runtime.h>

#include <memory>
#include <string>
#include <vector>

namespace giddy { namespace abi {
giddy_status_t fail(giddy_status_t status, const ::std::string& message);
giddy_status_t fail_cuda(cudaError_t error, const char* what);
} }

namespace {

using giddy::abi::fail;
using giddy::abi::fail_cuda;

struct operation_t {
	enum kind_t { copy, copy_back, launch, callback, record, wait } kind;
	// copy
	void* destination = nullptr; const void* source = nullptr; size_t bytes = 0;
	// launch: a deep copy of the caller's argument array
	const giddy_kernel_t* kernel = nullptr;
	giddy_launch_config_t config = { };
	::std::vector<giddy_arg_t> arguments;
	::std::vector<::std::string> names;
	// callback
	giddy_scan_callback_t enqueue = nullptr;
	::std::shared_ptr<void> user_data;
	// record / wait
	size_t event = 0;
};

} // namespace

struct giddy_scan {
	int device = 0;
	::std::vector<cudaStream_t> streams;
	::std::vector<::std::vector<operation_t>> chunks;
	::std::vector<cudaEvent_t> dependency_events;
	::std::vector<cudaEvent_t> join_events;
	cudaEvent_t fork_event = nullptr;
	bool graph_replay = false;
	cudaStream_t capture_stream = nullptr;
	cudaGraphExec_t executable = nullptr;

	~giddy_scan()
	{
		if (executable) { cudaGraphExecDestroy(executable); }
		if (capture_stream) { cudaStreamDestroy(capture_stream); }
		for (auto e : dependency_events) { cudaEventDestroy(e); }
		for (auto e : join_events) { cudaEventDestroy(e); }
		if (fork_event) { cudaEventDestroy(fork_event); }
		for (auto s : streams) { cudaStreamDestroy(s); }
	}

	void invalidate_graph()
	{
		if (executable) { cudaGraphExecDestroy(executable); executable = nullptr; }
	}

	::std::vector<operation_t>& chunk(unsigned index)
	{
		if (index >= chunks.size()) { chunks.resize(index + 1); }
		invalidate_graph();
		return chunks[index];
	}

	// fork from `origin`, enqueue every operation, join back into `origin`
	giddy_status_t enqueue_all(cudaStream_t origin)
	{
		cudaError_t error = cudaEventRecord(fork_event, origin);
		if (error != cudaSuccess) { return fail_cuda(error, "cudaEventRecord (scan fork)"); }
		for (auto s : streams) {
			error = cudaStreamWaitEvent(s, fork_event, 0);
			if (error != cudaSuccess) { return fail_cuda(error, "cudaStreamWaitEvent (scan fork)"); }
		}
		for (size_t c = 0; c < chunks.size(); c++) {
			cudaStream_t stream = streams[c % streams.size()];
			for (const auto& op : chunks[c]) {
				giddy_status_t status = GIDDY_SUCCESS;
				switch (op.kind) {
				case operation_t::copy:
					error = cudaMemcpyAsync(op.destination, op.source, op.bytes, cudaMemcpyHostToDevice, stream);
					if (error != cudaSuccess) { return fail_cuda(error, "cudaMemcpyAsync (scan chunk copy)"); }
					break;
				case operation_t::copy_back:
					error = cudaMemcpyAsync(op.destination, op.source, op.bytes, cudaMemcpyDeviceToHost, stream);
					if (error != cudaSuccess) { return fail_cuda(error, "cudaMemcpyAsync (scan result copy)"); }
					break;
				case operation_t::launch:
					status = giddy_kernel_enqueue_launch(op.kernel, stream, &op.config, op.arguments.data(), op.arguments.size());
					break;
				case operation_t::callback:
					status = op.enqueue(stream, op.user_data.get());
					break;
				case operation_t::record:
					error = cudaEventRecord(dependency_events[op.event], stream);
					if (error != cudaSuccess) { return fail_cuda(error, "cudaEventRecord (scan dependency)"); }
					break;
				case operation_t::wait:
					error = cudaStreamWaitEvent(stream, dependency_events[op.event], 0);
					if (error != cudaSuccess) { return fail_cuda(error, "cudaStreamWaitEvent (scan dependency)"); }
					break;
				}
				if (status != GIDDY_SUCCESS) { return status; }
			}
		}
		for (size_t s = 0; s < streams.size(); s++) {
			error = cudaEventRecord(join_events[s], streams[s]);
			if (error == cudaSuccess) { error = cudaStreamWaitEvent(origin, join_events[s], 0); }
			if (error != cudaSuccess) { return fail_cuda(error, "scan join"); }
		}
		return GIDDY_SUCCESS;
	}
};

extern "C" {

giddy_status_t giddy_scan_create(unsigned num_streams, giddy_scan_t** out)
{
	if (out == nullptr) { return fail(GIDDY_INVALID_ARGUMENT, "null output pointer"); }
	*out = nullptr;
	if (num_streams == 0 || num_streams > 64) { return fail(GIDDY_INVALID_ARGUMENT, "a scan needs between 1 and 64 side streams"); }
	::std::unique_ptr<giddy_scan> scan(new giddy_scan);
	cudaError_t error = cudaGetDevice(&scan->device);
	if (error != cudaSuccess) { return fail_cuda(error, "cudaGetDevice"); }
	for (unsigned s = 0; s < num_streams; s++) {
		cudaStream_t stream; cudaEvent_t event;
		error = cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking);
		if (error != cudaSuccess) { return fail_cuda(error, "cudaStreamCreateWithFlags (scan side stream)"); }
		scan->streams.push_back(stream);
		error = cudaEventCreateWithFlags(&event, cudaEventDisableTiming);
		if (error != cudaSuccess) { return fail_cuda(error, "cudaEventCreateWithFlags (scan join)"); }
		scan->join_events.push_back(event);
	}
	error = cudaEventCreateWithFlags(&scan->fork_event, cudaEventDisableTiming);
	if (error != cudaSuccess) { return fail_cuda(error, "cudaEventCreateWithFlags (scan fork)"); }
	*out = scan.release();
	return GIDDY_SUCCESS;
}

void giddy_scan_destroy(giddy_scan_t* scan) { delete scan; }

unsigned giddy_scan_num_streams(const giddy_scan_t* scan) { return scan ? (unsigned) scan->streams.size() : 0; }
unsigned giddy_scan_num_chunks(const giddy_scan_t* scan) { return scan ? (unsigned) scan->chunks.size() : 0; }

size_t giddy_scan_num_operations(const giddy_scan_t* scan)
{
	size_t count = 0;
	if (scan) { for (const auto& c : scan->chunks) { count += c.size(); } }
	return count;
}

giddy_status_t giddy_scan_add_copy(giddy_scan_t* scan, unsigned chunk, void* device_destination, const void* host_source, size_t bytes)
{
	if (scan == nullptr) { return fail(GIDDY_INVALID_ARGUMENT, "null scan"); }
	if (bytes == 0) { return GIDDY_SUCCESS; }
	if (device_destination == nullptr || host_source == nullptr) { return fail(GIDDY_INVALID_ARGUMENT, "null copy source or destination"); }
	operation_t op; op.kind = operation_t::copy;
	op.destination = device_destination; op.source = host_source; op.bytes = bytes;
	scan->chunk(chunk).push_back(::std::move(op));
	return GIDDY_SUCCESS;
}

giddy_status_t giddy_scan_add_copy_back(giddy_scan_t* scan, unsigned chunk, void* host_destination, const void* device_source, size_t bytes)
{
	if (scan == nullptr) { return fail(GIDDY_INVALID_ARGUMENT, "null scan"); }
	if (bytes == 0) { return GIDDY_SUCCESS; }
	if (host_destination == nullptr || device_source == nullptr) { return fail(GIDDY_INVALID_ARGUMENT, "null copy source or destination"); }
	operation_t op; op.kind = operation_t::copy_back;
	op.destination = host_destination; op.source = device_source; op.bytes = bytes;
	scan->chunk(chunk).push_back(::std::move(op));
	return GIDDY_SUCCESS;
}

giddy_status_t giddy_scan_add_launch(giddy_scan_t* scan, unsigned chunk, const giddy_kernel_t* kernel,
	const giddy_launch_config_t* launch_config, const giddy_arg_t* arguments, size_t num_arguments)
{
	if (scan == nullptr || kernel == nullptr || launch_config == nullptr) { return fail(GIDDY_INVALID_ARGUMENT, "null scan, kernel or launch configuration"); }
	if (arguments == nullptr && num_arguments != 0) { return fail(GIDDY_INVALID_ARGUMENT, "null argument array"); }
	operation_t op; op.kind = operation_t::launch;
	op.kernel = kernel; op.config = *launch_config;
	op.arguments.assign(arguments, arguments + num_arguments);
	op.names.reserve(num_arguments);
	for (size_t i = 0; i < num_arguments; i++) {
		if (arguments[i].name == nullptr) { return fail(GIDDY_INVALID_ARGUMENT, "a named argument has no name"); }
		op.names.emplace_back(arguments[i].name);
	}
	auto& operations = scan->chunk(chunk);
	operations.push_back(::std::move(op));
	auto& stored = operations.back(); // (moving a vector of strings keeps the strings where they are; set the pointers last all the same)
	for (size_t i = 0; i < num_arguments; i++) { stored.arguments[i].name = stored.names[i].c_str(); }
	return GIDDY_SUCCESS;
}

giddy_status_t giddy_scan_add_callback(giddy_scan_t* scan, unsigned chunk, giddy_scan_callback_t enqueue, void* user_data, void (*destroy_user_data)(void*))
{
	if (scan == nullptr || enqueue == nullptr) { return fail(GIDDY_INVALID_ARGUMENT, "null scan or callback"); }
	operation_t op; op.kind = operation_t::callback;
	op.enqueue = enqueue;
	if (destroy_user_data) { op.user_data = ::std::shared_ptr<void>(user_data, destroy_user_data); }
	else { op.user_data = ::std::shared_ptr<void>(user_data, [](void*) { }); }
	scan->chunk(chunk).push_back(::std::move(op));
	return GIDDY_SUCCESS;
}

giddy_status_t giddy_scan_add_dependency(giddy_scan_t* scan, unsigned chunk, unsigned on_chunk)
{
	if (scan == nullptr) { return fail(GIDDY_INVALID_ARGUMENT, "null scan"); }
	if (on_chunk >= chunk) { return fail(GIDDY_INVALID_ARGUMENT, "a chunk can only depend on an earlier chunk"); }
	if (on_chunk % scan->streams.size() == chunk % scan->streams.size()) { return GIDDY_SUCCESS; } // same stream: already ordered
	cudaEvent_t event;
	cudaError_t error = cudaEventCreateWithFlags(&event, cudaEventDisableTiming);
	if (error != cudaSuccess) { return fail_cuda(error, "cudaEventCreateWithFlags (scan dependency)"); }
	scan->dependency_events.push_back(event);
	operation_t record; record.kind = operation_t::record; record.event = scan->dependency_events.size() - 1;
	operation_t wait;   wait.kind = operation_t::wait;     wait.event = record.event;
	scan->chunk(on_chunk).push_back(::std::move(record));
	scan->chunk(chunk).push_back(::std::move(wait));
	return GIDDY_SUCCESS;
}

giddy_status_t giddy_scan_set_graph_replay(giddy_scan_t* scan, int enabled)
{
	if (scan == nullptr) { return fail(GIDDY_INVALID_ARGUMENT, "null scan"); }
	scan->graph_replay = enabled != 0;
	if (!scan->graph_replay) { scan->invalidate_graph(); }
	return GIDDY_SUCCESS;
}

giddy_status_t giddy_scan_run(giddy_scan_t* scan, void* stream)
{
	if (scan == nullptr) { return fail(GIDDY_INVALID_ARGUMENT, "null scan"); }
	int device = 0;
	cudaError_t error = cudaGetDevice(&device);
	if (error != cudaSuccess) { return fail_cuda(error, "cudaGetDevice"); }
	if (device != scan->device) { return fail(GIDDY_INVALID_ARGUMENT, "a scan must be run with the device it was created on current"); }
	cudaStream_t origin = static_cast<cudaStream_t>(stream);
	if (!scan->graph_replay) { return scan->enqueue_all(origin); }

	if (scan->executable == nullptr) {
		if (scan->capture_stream == nullptr) {
			error = cudaStreamCreateWithFlags(&scan->capture_stream, cudaStreamNonBlocking);
			if (error != cudaSuccess) { return fail_cuda(error, "cudaStreamCreateWithFlags (scan capture)"); }
		}
		error = cudaStreamBeginCapture(scan->capture_stream, cudaStreamCaptureModeRelaxed);
		if (error != cudaSuccess) { return fail_cuda(error, "cudaStreamBeginCapture"); }
		giddy_status_t status = scan->enqueue_all(scan->capture_stream);
		cudaGraph_t graph = nullptr;
		error = cudaStreamEndCapture(scan->capture_stream, &graph);
		if (status != GIDDY_SUCCESS) { if (graph) { cudaGraphDestroy(graph); } return status; }
		if (error != cudaSuccess) { return fail_cuda(error, "cudaStreamEndCapture"); }
		error = cudaGraphInstantiate(&scan->executable, graph, 0);
		cudaGraphDestroy(graph);
		if (error != cudaSuccess) { scan->executable = nullptr; return fail_cuda(error, "cudaGraphInstantiate"); }
	}
	error = cudaGraphLaunch(scan->executable, origin);
	if (error != cudaSuccess) { return fail_cuda(error, "cudaGraphLaunch"); }
	return GIDDY_SUCCESS;
}

} // extern "C"
