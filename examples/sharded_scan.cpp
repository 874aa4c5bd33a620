// Multi-GPU mode of use: ONE compressed column, kept in pinned host memory, is cut into contiguous ranges of anchored
// segments -- one per visible GPU -- and every GPU decodes its range as a self-contained column.  No data-path
// collective: the carry-ins at the shard boundaries (byte offset of the first delta, value preceding the first
// element) are already in the formats' anchor arrays, and the library's planner (giddy_shard_plan) picks them out.
// The one exchange is the broadcast of a dictionary from the GPU which loaded it (NCCL), shown on the second column.
//
// Columns (BASELINE.json configs[3] and configs[2] in small):
//   A  Delta int64 with per-segment anchors over Discard-Zero-Bytes variable-width deltas (1-4 bytes, 2-bit size
//      fields), L = 1024: the fused decoder "decompression::composite::delta_of_discard_zero_bytes_variable_width"
//      AND the reference's two-kernel chain (DZB-variable, then Delta), both on every shard;
//   B  Dictionary uint8 -> int64, 256 entries; the dictionary is copied to GPU 0 only and broadcast to the others.
// Each GPU is driven by a worker thread of a cuda::device_group_t; a thread plans its shard, stages the slices in
// pinned memory on its GPU's NUMA node (rebasing the byte anchors), copies them in, decodes, copies the result back.
// The concatenated results must equal a scalar host decode of the whole column.  Exit code 0 on success.
//
//   sharded_scan [log2 elements = 24] [GPUs = all]
#include <giddy/kernel_wrappers.hpp>
#include <giddy/multi_gpu.hpp>

#include <cuda_runtime_api.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>

#define CHECK_CUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
	throw std::runtime_error(std::string(#call) + " failed: " + cudaGetErrorString(e_)); } } while (0)

namespace {

struct column_a_t { // Delta over DZB-variable
	uint64_t n, L, num_segments;
	std::vector<uint8_t>  compressed_data;        // + 8 bytes of padding: the kernels may read the word holding the last byte
	std::vector<uint32_t> packed_element_sizes;   // 16 two-bit fields per container: size - 1
	std::vector<uint32_t> position_anchors;       // byte offset of each segment's first delta
	std::vector<uint64_t> anchor_values;          // the value preceding each segment
	std::vector<uint64_t> expected;
	uint64_t num_bytes;
};

column_a_t make_column_a(uint64_t n, uint64_t L)
{
	column_a_t c;
	c.n = n; c.L = L; c.num_segments = (n + L - 1) / L;
	std::mt19937_64 random(8);
	c.packed_element_sizes.assign((n + 15) / 16, 0);
	c.expected.resize(n);
	c.compressed_data.reserve(n * 2);
	uint64_t value = random() >> 24;
	for (uint64_t i = 0; i < n; i++) {
		if (i % L == 0) { c.position_anchors.push_back((uint32_t) c.compressed_data.size()); c.anchor_values.push_back(value); }
		const unsigned roll = (unsigned)(random() % 100);
		const unsigned width = roll < 50 ? 1 : roll < 80 ? 2 : roll < 95 ? 3 : 4;           // P = {.50, .30, .15, .05}
		const uint64_t low = width == 1 ? 0 : 1ull << (8 * (width - 1)), high = 1ull << (8 * width);
		const uint32_t delta = (uint32_t)(low + random() % (high - low));
		for (unsigned b = 0; b < width; b++) { c.compressed_data.push_back((uint8_t)(delta >> (8 * b))); }
		c.packed_element_sizes[i / 16] |= (uint32_t)(width - 1) << (2 * (i % 16));
		value += delta;
		c.expected[i] = value;
	}
	c.num_bytes = c.compressed_data.size();
	c.compressed_data.resize(c.num_bytes + 8, 0);
	return c;
}

struct stopwatch_t {
	std::chrono::steady_clock::time_point start = std::chrono::steady_clock::now();
	double ms() const { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - start).count(); }
};

} // namespace

int main(int argc, char** argv)
{
	const unsigned log2_n = argc > 1 ? (unsigned) std::atoi(argv[1]) : 24;
	const uint64_t n = 1ull << log2_n, L = 1024;
	int failures = 0;
	try {
		int visible = 0;
		CHECK_CUDA(cudaGetDeviceCount(&visible));
		const int wanted = argc > 2 ? std::atoi(argv[2]) : visible;
		std::vector<int> devices;
		for (int d = 0; d < visible && d < wanted; d++) { devices.push_back(d); }
		cuda::device_group_t gpus(devices);
		const unsigned world = gpus.size();
		std::printf("%u GPU(s); column A: 2^%u int64 elements, Delta over DZB-variable, L = %llu\n", world, log2_n, (unsigned long long) L);

		const column_a_t a = make_column_a(n, L);
		auto fused = cuda::registered::kernel_t::produceSubclass("decompression::composite::delta_of_discard_zero_bytes_variable_width::kernel_t<4u, 8u, 4u, 4u>");
		auto dzbv = cuda::registered::kernel_t::produceSubclass("decompression::discard_zero_bytes::variable_width::kernel_t<4u, 4u, 4u>");
		auto delta = cuda::registered::kernel_t::produceSubclass("decompression::delta::kernel_t<4u, 8u, 4u>");

		// what the planner needs to know about the whole column
		giddy_column_layout_t layout = { };
		layout.scheme = GIDDY_SCHEME_DISCARD_ZERO_BYTES_VARIABLE;
		layout.length = n; layout.segment_length = L;
		layout.index_size = 4; layout.uncompressed_size = 4; layout.compressed_size = 4 /* size containers */;
		layout.bits_per_element_size = 2; layout.num_bytes = a.num_bytes;
		layout.anchors = a.position_anchors.data();

		std::vector<uint64_t> result_fused(n), result_chain(n);
		std::vector<double> milliseconds(world);
		gpus.run([&](unsigned rank, unsigned size, int device) {
			const auto range = cuda::sharding::split_segments(a.num_segments, size, rank);
			const giddy_shard_t shard = cuda::sharding::plan(layout, range.first, range.second);
			if (shard.num_elements == 0) { return; }
			// stage the shard's slices in pinned memory (byte anchors rebased to the shard's first byte) ...
			const auto& data_slice = cuda::sharding::slice_of(shard, "compressed_data");
			const auto& sizes_slice = cuda::sharding::slice_of(shard, "packed_element_sizes");
			const auto& anchors_slice = cuda::sharding::slice_of(shard, "position_anchors");
			cuda::host_buffer_t<uint8_t> host_data(giddy_slice_bytes(&data_slice) + 8, device);
			cuda::host_buffer_t<uint32_t> host_sizes(sizes_slice.count, device), host_anchors(anchors_slice.count, device);
			cuda::sharding::copy_slice(data_slice, host_data.data(), a.compressed_data.data());
			std::memset(host_data.data() + giddy_slice_bytes(&data_slice), 0, 8);
			cuda::sharding::copy_slice(sizes_slice, host_sizes.data(), a.packed_element_sizes.data());
			cuda::sharding::copy_slice(anchors_slice, host_anchors.data(), a.position_anchors.data());
			// ... the Delta stage's carry-in is the shard's first anchor value: a plain slice by segment
			cuda::host_buffer_t<uint64_t> host_values(shard.num_segments, device), host_out(shard.num_elements, device);
			std::memcpy(host_values.data(), a.anchor_values.data() + shard.first_segment, shard.num_segments * sizeof(uint64_t));

			uint8_t* data; uint32_t *sizes, *anchors, *widened; uint64_t *values, *out;
			CHECK_CUDA(cudaMalloc((void**) &data, host_data.size()));
			CHECK_CUDA(cudaMalloc((void**) &sizes, host_sizes.size() * 4));
			CHECK_CUDA(cudaMalloc((void**) &anchors, host_anchors.size() * 4));
			CHECK_CUDA(cudaMalloc((void**) &values, host_values.size() * 8));
			CHECK_CUDA(cudaMalloc((void**) &widened, shard.num_elements * 4));
			CHECK_CUDA(cudaMalloc((void**) &out, shard.num_elements * 8));
			cudaStream_t stream;
			CHECK_CUDA(cudaStreamCreate(&stream));
			auto properties = cuda::device::current::get().properties();
			const cuda::kernel_t::arguments_type shape = {
				{ "length", cuda::any((size_t) shard.num_elements) }, { "segment_length", cuda::any((size_t) L) } };
			const uint32_t length = (uint32_t) shard.num_elements, segment_length = (uint32_t) L;

			const stopwatch_t timer;
			CHECK_CUDA(cudaMemcpyAsync(data, host_data.data(), host_data.size(), cudaMemcpyHostToDevice, stream));
			CHECK_CUDA(cudaMemcpyAsync(sizes, host_sizes.data(), host_sizes.size() * 4, cudaMemcpyHostToDevice, stream));
			CHECK_CUDA(cudaMemcpyAsync(anchors, host_anchors.data(), host_anchors.size() * 4, cudaMemcpyHostToDevice, stream));
			CHECK_CUDA(cudaMemcpyAsync(values, host_values.data(), host_values.size() * 8, cudaMemcpyHostToDevice, stream));
			// the fused decoder, on the shard exactly as on a whole column
			fused->enqueue_launch(stream, fused->resolve_launch_configuration(properties, shape), {
				{ "decompressed", cuda::any(out) }, { "compressed_data", cuda::any((const unsigned char*) data) },
				{ "packed_element_sizes", cuda::any((const uint32_t*) sizes) }, { "position_anchors", cuda::any((const uint32_t*) anchors) },
				{ "anchor_values", cuda::any((const uint64_t*) values) }, { "segment_length", cuda::any(segment_length) }, { "length", cuda::any(length) },
				{ "min_represented_element_size", cuda::any(1u) }, { "bits_per_element_size", cuda::any(2u) } });
			CHECK_CUDA(cudaMemcpyAsync(host_out.data(), out, shard.num_elements * 8, cudaMemcpyDeviceToHost, stream));
			CHECK_CUDA(cudaStreamSynchronize(stream));
			milliseconds[rank] = timer.ms();
			std::memcpy(result_fused.data() + shard.first_element, host_out.data(), shard.num_elements * 8);

			// the reference's way: two kernels, the widened deltas going through device memory
			CHECK_CUDA(cudaMemsetAsync(out, 0xFF, shard.num_elements * 8, stream));
			dzbv->enqueue_launch(stream, dzbv->resolve_launch_configuration(properties, shape), {
				{ "decompressed", cuda::any(widened) }, { "compressed_data", cuda::any((const unsigned char*) data) },
				{ "packed_element_sizes", cuda::any((const uint32_t*) sizes) }, { "position_anchors", cuda::any((const uint32_t*) anchors) },
				{ "segment_length", cuda::any(segment_length) }, { "length", cuda::any(length) },
				{ "min_represented_element_size", cuda::any(1u) }, { "bits_per_element_size", cuda::any(2u) } });
			delta->enqueue_launch(stream, delta->resolve_launch_configuration(properties, shape), {
				{ "decompressed", cuda::any(out) }, { "compressed_input", cuda::any((const uint32_t*) widened) },
				{ "anchor_values", cuda::any((const uint64_t*) values) }, { "length", cuda::any(length) }, { "segment_length", cuda::any(segment_length) } });
			CHECK_CUDA(cudaMemcpyAsync(host_out.data(), out, shard.num_elements * 8, cudaMemcpyDeviceToHost, stream));
			CHECK_CUDA(cudaStreamSynchronize(stream));
			std::memcpy(result_chain.data() + shard.first_element, host_out.data(), shard.num_elements * 8);
			cudaFree(data); cudaFree(sizes); cudaFree(anchors); cudaFree(values); cudaFree(widened); cudaFree(out);
			cudaStreamDestroy(stream);
		});
		double slowest = 0;
		for (double ms : milliseconds) { slowest = ms > slowest ? ms : slowest; }
		uint64_t wrong_fused = 0, wrong_chain = 0;
		for (uint64_t i = 0; i < n; i++) { wrong_fused += result_fused[i] != a.expected[i]; wrong_chain += result_chain[i] != a.expected[i]; }
		std::printf("%s  column A, fused decoder, %u shard(s): %llu mismatches of %llu; copy in + decode + copy out %.2f ms on the slowest GPU\n",
			wrong_fused ? "FAILURE" : "SUCCESS", world, (unsigned long long) wrong_fused, (unsigned long long) n, slowest);
		std::printf("%s  column A, DZB-variable then Delta, %u shard(s): %llu mismatches\n", wrong_chain ? "FAILURE" : "SUCCESS", world, (unsigned long long) wrong_chain);
		failures += (wrong_fused != 0) + (wrong_chain != 0);

		// ---- column B: Dictionary; the dictionary reaches GPUs 1.. by broadcast from GPU 0 ------------------------------
		std::mt19937_64 random(5);
		std::vector<uint64_t> dictionary(256);
		for (auto& entry : dictionary) { entry = random(); }
		std::vector<uint8_t> indices(n);
		for (auto& index : indices) { index = (uint8_t) random(); }
		std::vector<void*> device_dictionaries(world, nullptr);
		gpus.run([&](unsigned rank, unsigned, int) {
			CHECK_CUDA(cudaMalloc(&device_dictionaries[rank], 256 * 8));
			if (rank == 0) { CHECK_CUDA(cudaMemcpy(device_dictionaries[0], dictionary.data(), 256 * 8, cudaMemcpyHostToDevice)); }
			else { CHECK_CUDA(cudaMemset(device_dictionaries[rank], 0, 256 * 8)); }
		});
		gpus.broadcast(device_dictionaries, 256 * 8, 0); // asynchronous on every device's default stream; the launches below follow on it
		auto dictionary_kernel = cuda::registered::kernel_t::produceSubclass("decompression::dictionary::kernel_t<4u, 8u, 1u>");
		giddy_column_layout_t layout_b = { };
		layout_b.scheme = GIDDY_SCHEME_DICTIONARY; layout_b.length = n; layout_b.segment_length = L; layout_b.compressed_size = 1;
		std::vector<uint64_t> result_b(n);
		gpus.run([&](unsigned rank, unsigned size, int device) {
			const auto range = cuda::sharding::split_segments((n + L - 1) / L, size, rank);
			const giddy_shard_t shard = cuda::sharding::plan(layout_b, range.first, range.second);
			if (shard.num_elements == 0) { return; }
			cuda::host_buffer_t<uint8_t> host_in(shard.num_elements, device);
			cuda::host_buffer_t<uint64_t> host_out(shard.num_elements, device);
			cuda::sharding::copy_slice(cuda::sharding::slice_of(shard, "compressed_input"), host_in.data(), indices.data());
			uint8_t* in; uint64_t* out;
			CHECK_CUDA(cudaMalloc((void**) &in, shard.num_elements));
			CHECK_CUDA(cudaMalloc((void**) &out, shard.num_elements * 8));
			CHECK_CUDA(cudaMemcpyAsync(in, host_in.data(), shard.num_elements, cudaMemcpyHostToDevice, nullptr));
			auto config = dictionary_kernel->resolve_launch_configuration(cuda::device::current::get().properties(), {
				{ "num_dictionary_entries", cuda::any((size_t) 256) }, { "data_length", cuda::any((size_t) shard.num_elements) } });
			dictionary_kernel->enqueue_launch(nullptr, config, {
				{ "decompressed", cuda::any(out) }, { "compressed_input", cuda::any((const uint8_t*) in) },
				{ "dictionary_entries", cuda::any((const uint64_t*) device_dictionaries[rank]) },
				{ "length", cuda::any((uint32_t) shard.num_elements) }, { "num_dictionary_entries", cuda::any((uint32_t) 256) } });
			CHECK_CUDA(cudaMemcpyAsync(host_out.data(), out, shard.num_elements * 8, cudaMemcpyDeviceToHost, nullptr));
			CHECK_CUDA(cudaStreamSynchronize(nullptr));
			std::memcpy(result_b.data() + shard.first_element, host_out.data(), shard.num_elements * 8);
			cudaFree(in); cudaFree(out); cudaFree(device_dictionaries[rank]);
		});
		uint64_t wrong_b = 0;
		for (uint64_t i = 0; i < n; i++) { wrong_b += result_b[i] != dictionary[indices[i]]; }
		std::printf("%s  column B, Dictionary with the dictionary broadcast from GPU 0%s, %u shard(s): %llu mismatches\n", wrong_b ? "FAILURE" : "SUCCESS",
			world > 1 ? " (NCCL)" : "", world, (unsigned long long) wrong_b);
		failures += wrong_b != 0;
	}
	catch (const std::exception& e) { std::printf("FAILURE  exception: %s\n", e.what()); failures++; }
	return failures ? 1 : 0;
}
