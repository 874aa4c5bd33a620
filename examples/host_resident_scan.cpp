// The host-resident mode of use: a compressed column which lives in PINNED HOST memory is shipped to the GPU
// chunk by chunk on side streams and decoded as the chunks land (the use Giddy is made for: README.md:18-24,
// "transfer the data over the bus compressed, decompress it on the GPU").  The reference leaves this to its
// caller; here it is a library facility, cuda::host_resident::scan_t (include/giddy/host_resident_scan.hpp).
//
// The column is the reference's known-answer Frame-of-Reference data (examples/src/modes_of_use/common.cuh:
// 29-105: int32 out, int16 offsets iota per segment, constant<4, int> model with reference value 100 * segment
// index), made 64 segments long and cut into 8 chunks over 3 streams.  The plan is built once and replayed three
// times -- re-enqueued, then twice as a captured CUDA graph -- and every replay must reproduce the known answer.
// Exit code 0 on success.
#include <giddy/kernel_wrappers.hpp>
#include <giddy/host_resident_scan.hpp>

#include <cuda_runtime_api.h>

#include <cstdio>
#include <cstdlib>
#include <numeric>
#include <string>
#include <vector>

using coefficients_type = cuda::functors::unary::parametric_model::constant<4, int32_t>::coefficients_type;

#define CHECK_CUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
	std::fprintf(stderr, "%s failed: %s\n", #call, cudaGetErrorString(e_)); std::exit(2); } } while (0)

int main()
{
	const uint32_t segment_length = 2 << 10, num_segments = 64, length = num_segments * segment_length;
	const unsigned num_chunks = 8, num_streams = 3, segments_per_chunk = num_segments / num_chunks;

	int16_t* host_offsets; coefficients_type* host_coefficients;
	CHECK_CUDA(cudaHostAlloc((void**) &host_offsets, length * sizeof(int16_t), cudaHostAllocDefault));
	CHECK_CUDA(cudaHostAlloc((void**) &host_coefficients, num_segments * sizeof(coefficients_type), cudaHostAllocDefault));
	for (uint32_t s = 0; s < num_segments; s++) {
		host_coefficients[s][0] = (int32_t)(s * 100);
		std::iota(host_offsets + s * segment_length, host_offsets + (s + 1) * segment_length, (int16_t) 0);
	}
	int32_t* decompressed; int16_t* offsets; coefficients_type* coefficients;
	CHECK_CUDA(cudaMalloc((void**) &decompressed, length * sizeof(int32_t)));
	CHECK_CUDA(cudaMalloc((void**) &offsets, length * sizeof(int16_t)));
	CHECK_CUDA(cudaMalloc((void**) &coefficients, num_segments * sizeof(coefficients_type)));
	cudaStream_t stream;
	CHECK_CUDA(cudaStreamCreate(&stream));

	int failures = 0;
	try {
		auto kernel = cuda::registered::kernel_t::produceSubclass(
			"decompression::frame_of_reference::kernel_t<4u, int, short, cuda::functors::unary::parametric_model::constant<4u, int> >");
		auto properties = cuda::device::current::get().properties();
		const uint32_t chunk_length = segments_per_chunk * segment_length;
		auto config = kernel->resolve_launch_configuration(properties, {
			{ "length", cuda::any((size_t) chunk_length) }, { "segment_length", cuda::any((size_t) segment_length) } });

		cuda::host_resident::scan_t scan(num_streams);
		for (unsigned c = 0; c < num_chunks; c++) {
			const uint32_t s0 = c * segments_per_chunk, e0 = s0 * segment_length;
			scan.add_copy(c, offsets + e0, host_offsets + e0, chunk_length * sizeof(int16_t));
			scan.add_copy(c, coefficients + s0, host_coefficients + s0, segments_per_chunk * sizeof(coefficients_type));
			scan.add_launch(c, *kernel, config, {
				{ "decompressed",               cuda::any(decompressed + e0) },
				{ "compressed_input",           cuda::any((const int16_t*)(offsets + e0)) },
				{ "segment_model_coefficients", cuda::any((const coefficients_type*)(coefficients + s0)) },
				{ "length",                     cuda::any(chunk_length) },
				{ "segment_length",             cuda::any(segment_length) } });
		}
		std::printf("plan: %u chunks, %zu operations, %u side streams\n", scan.num_chunks(), scan.num_operations(), scan.num_streams());

		const char* names[] = { "re-enqueued", "captured as a CUDA graph", "graph replayed" };
		for (int replay = 0; replay < 3; replay++) {
			scan.set_graph_replay(replay > 0);
			CHECK_CUDA(cudaMemsetAsync(decompressed, 0xFF, length * sizeof(int32_t), stream));
			CHECK_CUDA(cudaMemsetAsync(offsets, 0, length * sizeof(int16_t), stream));
			scan.run(stream);
			CHECK_CUDA(cudaStreamSynchronize(stream));
			std::vector<int32_t> result(length);
			CHECK_CUDA(cudaMemcpy(result.data(), decompressed, length * sizeof(int32_t), cudaMemcpyDeviceToHost));
			uint32_t mismatches = 0;
			for (uint32_t s = 0; s < num_segments; s++) {
				for (uint32_t i = 0; i < segment_length; i++) { mismatches += result[s * segment_length + i] != (int32_t)(s * 100 + i); }
			}
			std::printf("%s  host-resident scan, %s (%u mismatches of %u)\n", mismatches ? "FAILURE" : "SUCCESS", names[replay], mismatches, length);
			failures += mismatches ? 1 : 0;
		}
	}
	catch (const std::exception& e) { std::printf("FAILURE  exception: %s\n", e.what()); failures++; }
	cudaFree(decompressed); cudaFree(offsets); cudaFree(coefficients); cudaFreeHost(host_offsets); cudaFreeHost(host_coefficients);
	return failures ? 1 : 0;
}
