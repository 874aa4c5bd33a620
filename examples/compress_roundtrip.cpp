// A column compressed AND decompressed on the GPU through the factory, in the reference's mode (iii) idiom
// (examples/src/modes_of_use/factory_provided_type_erased_wrapper.cu:8-46): the compressors are registered under keys of
// their own, `compression::<scheme>::kernel_t<...>`, next to the reference's `decompression::...` keys, take the same
// kind of type-erased argument maps and go through the same launch-configuration resolver.  (The reference ships no
// compressors -- README.md:128-130, its open Issue #3 -- so there is no reference counterpart to mirror, only its formats.)
//
//   1. Run-Length Encoding: uint32 column with runs of 1..31 elements, uint16 run lengths, anchored segments of 1024
//      elements; `compression::run_length_encoding::kernel_t<4u, 4u, 2u, 4u>` writes run_data, run_lengths,
//      position_anchors and intra_run_anchor_offsets; `decompression::run_length_encoding::kernel_t<4u, 4u, 2u, 4u>`
//      reads them back.
//   2. Delta over uint64 with uint32 deltas, then the deltas through Discard-Zero-Bytes (variable width, 2-bit size
//      fields) -- BASELINE configs[3]'s chain, compressed by two kernels and decompressed by the ONE fused decoder
//      `decompression::composite::delta_of_discard_zero_bytes_variable_width::kernel_t<4u, 8u, 4u, 4u>`.
//
// Host-only C++17 (g++), linked against libgiddy_b200.so and the CUDA runtime.  Exit code 0 when both columns come back
// bit for bit.
#include <giddy/kernel_wrappers.hpp>
#include <giddy/compression_wrappers.hpp>

#include <cuda_runtime_api.h>

#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <random>
#include <string>
#include <vector>

#define CHECK_CUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
	std::fprintf(stderr, "%s failed: %s\n", #call, cudaGetErrorString(e_)); std::exit(2); } } while (0)

template <typename T> static T* device_array(size_t count)
{
	T* p = nullptr;
	CHECK_CUDA(cudaMalloc(reinterpret_cast<void**>(&p), (count ? count : 1) * sizeof(T)));
	return p;
}

template <typename T> static T* device_copy(const std::vector<T>& host)
{
	T* p = device_array<T>(host.size());
	CHECK_CUDA(cudaMemcpy(p, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice));
	return p;
}

// resolve + enqueue, as a caller of the reference's factory would
static void launch(const std::string& key, cudaStream_t stream, cuda::kernel_t::arguments_type resolution_arguments, cuda::kernel_t::arguments_type launch_arguments)
{
	auto kernel = cuda::registered::kernel_t::produceSubclass(key);
	auto properties = cuda::device::current::get().properties();
	auto attributes = kernel->get_device_function().attributes();
	auto config = kernel->resolve_launch_configuration(properties, attributes, resolution_arguments);
	kernel->enqueue_launch(stream, config, launch_arguments);
}

template <typename T> static size_t mismatches(const T* device, const std::vector<T>& expected)
{
	std::vector<T> got(expected.size());
	CHECK_CUDA(cudaMemcpy(got.data(), device, got.size() * sizeof(T), cudaMemcpyDeviceToHost));
	size_t bad = 0;
	for (size_t i = 0; i < got.size(); i++) { if (got[i] != expected[i]) { bad++; } }
	return bad;
}

static bool run_length_encoding_round_trip(cudaStream_t stream, uint32_t length)
{
	constexpr uint32_t segment_length = 1024;
	const uint32_t num_segments = (length + segment_length - 1) / segment_length;
	std::mt19937_64 rng(3);
	std::vector<uint32_t> column(length);
	size_t num_runs = 0;
	for (uint32_t i = 0; i < length; ) {
		uint32_t run = 1 + (uint32_t)(rng() % 31), value = (uint32_t) rng();
		if (i != 0 && value == column[i - 1]) { value++; } // (equal neighbours would merge two runs: keep the count exact)
		for (uint32_t j = 0; j < run && i < length; j++, i++) { column[i] = value; }
		num_runs++;
	}
	const size_t run_capacity = num_runs + 16;
	auto* uncompressed = device_copy(column);
	auto* run_data = device_array<uint32_t>(run_capacity);
	auto* run_lengths = device_array<uint16_t>(run_capacity);
	auto* position_anchors = device_array<uint32_t>(num_segments);
	auto* intra_run_anchor_offsets = device_array<uint16_t>(num_segments);
	auto* num_element_runs = device_array<uint32_t>(1);
	auto* status = device_array<unsigned>(1);
	CHECK_CUDA(cudaMemset(status, 0, sizeof(unsigned)));

	launch("compression::run_length_encoding::kernel_t<4u, 4u, 2u, 4u>", stream,
		{ { "length", cuda::any((size_t) length) }, { "segment_length", cuda::any((size_t) segment_length) } },
		{
			{ "run_data", cuda::any(run_data) }, { "run_lengths", cuda::any(run_lengths) },
			{ "position_anchors", cuda::any(position_anchors) }, { "intra_run_anchor_offsets", cuda::any(intra_run_anchor_offsets) },
			{ "num_element_runs", cuda::any(num_element_runs) }, { "uncompressed", cuda::any((const uint32_t*) uncompressed) },
			{ "length", cuda::any(length) }, { "segment_length", cuda::any(segment_length) },
			{ "run_capacity", cuda::any(run_capacity) }, { "status", cuda::any(status) },
		});
	uint32_t runs_found = 0; unsigned flags = 0;
	CHECK_CUDA(cudaStreamSynchronize(stream));
	CHECK_CUDA(cudaMemcpy(&runs_found, num_element_runs, sizeof(runs_found), cudaMemcpyDeviceToHost));
	CHECK_CUDA(cudaMemcpy(&flags, status, sizeof(flags), cudaMemcpyDeviceToHost));

	auto* decompressed = device_array<uint32_t>(length);
	CHECK_CUDA(cudaMemset(decompressed, 0xA5, (size_t) length * sizeof(uint32_t)));
	launch("decompression::run_length_encoding::kernel_t<4u, 4u, 2u, 4u>", stream,
		{ { "length", cuda::any((size_t) length) }, { "segment_length", cuda::any((size_t) segment_length) } },
		{
			{ "decompressed", cuda::any(decompressed) }, { "run_data", cuda::any((const uint32_t*) run_data) },
			{ "run_lengths", cuda::any((const uint16_t*) run_lengths) },
			{ "position_anchors", cuda::any((const uint32_t*) position_anchors) },
			{ "intra_run_anchor_offsets", cuda::any((const uint16_t*) intra_run_anchor_offsets) },
			{ "segment_length", cuda::any(segment_length) }, { "num_segments", cuda::any(num_segments) },
			{ "num_element_runs", cuda::any(runs_found) }, { "length", cuda::any(length) },
		});
	CHECK_CUDA(cudaStreamSynchronize(stream));
	const size_t bad = mismatches(decompressed, column);
	const bool ok = bad == 0 && flags == 0 && runs_found == num_runs;
	std::printf("%s  RLE uint32 / uint16 lengths, %u elements: %u runs found (%zu generated), status %u, %zu mismatches after the round trip; "
		"%.2f bytes per element compressed\n", ok ? "SUCCESS" : "FAILURE", length, runs_found, num_runs, flags, bad,
		((double) runs_found * 6 + (double) num_segments * 6) / length);
	for (void* p : { (void*) uncompressed, (void*) run_data, (void*) run_lengths, (void*) position_anchors, (void*) intra_run_anchor_offsets,
		(void*) num_element_runs, (void*) status, (void*) decompressed }) { cudaFree(p); }
	return ok;
}

static bool delta_of_variable_width_round_trip(cudaStream_t stream, uint32_t length)
{
	constexpr uint32_t segment_length = 1024;
	const uint32_t num_segments = (length + segment_length - 1) / segment_length;
	std::mt19937_64 rng(8);
	std::vector<uint64_t> column(length);
	const uint64_t base = rng();
	uint64_t value = base;
	for (uint32_t i = 0; i < length; i++) { // deltas of 1-4 significant bytes with P = .50 / .30 / .15 / .05
		const unsigned roll = (unsigned)(rng() % 100), width = roll < 50 ? 1 : roll < 80 ? 2 : roll < 95 ? 3 : 4;
		value += (rng() & 0xffffffffull) >> (8 * (4 - width));
		column[i] = value;
	}
	auto* uncompressed = device_copy(column);
	auto* deltas = device_array<uint32_t>(length);
	auto* anchor_values = device_array<uint64_t>(num_segments);
	const size_t data_capacity = (size_t) length * 4 + 16;
	auto* compressed_data = device_array<unsigned char>(data_capacity);
	auto* packed_element_sizes = device_array<uint32_t>((length + 15) / 16 + 1);
	auto* position_anchors = device_array<uint32_t>(num_segments);
	auto* num_bytes = device_array<std::uint64_t>(1);
	auto* status = device_array<unsigned>(1);
	CHECK_CUDA(cudaMemset(status, 0, sizeof(unsigned)));

	launch("compression::delta::kernel_t<4u, 8u, 4u>", stream,
		{ { "length", cuda::any((size_t) length) }, { "segment_length", cuda::any((size_t) segment_length) } },
		{
			{ "compressed_output", cuda::any(deltas) }, { "anchor_values", cuda::any(anchor_values) },
			{ "uncompressed", cuda::any((const uint64_t*) uncompressed) }, { "length", cuda::any(length) }, { "segment_length", cuda::any(segment_length) },
			{ "base_value", cuda::any(base) }, { "status", cuda::any(status) },
		});
	launch("compression::discard_zero_bytes::variable_width::kernel_t<4u, 4u, 4u>", stream,
		{ { "length", cuda::any((size_t) length) }, { "segment_length", cuda::any((size_t) segment_length) } },
		{
			{ "compressed_data", cuda::any(compressed_data) }, { "packed_element_sizes", cuda::any(packed_element_sizes) },
			{ "position_anchors", cuda::any(position_anchors) }, { "num_bytes", cuda::any(num_bytes) },
			{ "uncompressed", cuda::any((const uint32_t*) deltas) }, { "length", cuda::any(length) }, { "segment_length", cuda::any(segment_length) },
			{ "min_represented_element_size", cuda::any(1u) }, { "bits_per_element_size", cuda::any(2u) },
			{ "data_capacity", cuda::any(data_capacity) }, { "status", cuda::any(status) },
		});
	std::uint64_t bytes = 0; unsigned flags = 0;
	CHECK_CUDA(cudaStreamSynchronize(stream));
	CHECK_CUDA(cudaMemcpy(&bytes, num_bytes, sizeof(bytes), cudaMemcpyDeviceToHost));
	CHECK_CUDA(cudaMemcpy(&flags, status, sizeof(flags), cudaMemcpyDeviceToHost));

	auto* decompressed = device_array<uint64_t>(length);
	CHECK_CUDA(cudaMemset(decompressed, 0xA5, (size_t) length * sizeof(uint64_t)));
	launch("decompression::composite::delta_of_discard_zero_bytes_variable_width::kernel_t<4u, 8u, 4u, 4u>", stream,
		{ { "length", cuda::any((size_t) length) }, { "segment_length", cuda::any((size_t) segment_length) } },
		{
			{ "decompressed", cuda::any(decompressed) }, { "compressed_data", cuda::any((const unsigned char*) compressed_data) },
			{ "packed_element_sizes", cuda::any((const uint32_t*) packed_element_sizes) },
			{ "position_anchors", cuda::any((const uint32_t*) position_anchors) }, { "anchor_values", cuda::any((const uint64_t*) anchor_values) },
			{ "segment_length", cuda::any(segment_length) }, { "length", cuda::any(length) },
			{ "min_represented_element_size", cuda::any(1u) }, { "bits_per_element_size", cuda::any(2u) },
		});
	CHECK_CUDA(cudaStreamSynchronize(stream));
	const size_t bad = mismatches(decompressed, column);
	const bool ok = bad == 0 && flags == 0;
	std::printf("%s  Delta uint64 <- uint32 over DZB-variable (2-bit sizes), %u elements: %llu bytes of deltas (%.2f per element), status %u, "
		"%zu mismatches after the round trip through the fused decoder\n", ok ? "SUCCESS" : "FAILURE", length, (unsigned long long) bytes, (double) bytes / length, flags, bad);
	for (void* p : { (void*) uncompressed, (void*) deltas, (void*) anchor_values, (void*) compressed_data, (void*) packed_element_sizes,
		(void*) position_anchors, (void*) num_bytes, (void*) status, (void*) decompressed }) { cudaFree(p); }
	return ok;
}

int main(int argc, char** argv)
{
	const unsigned log2_length = argc > 1 ? (unsigned) std::atoi(argv[1]) : 24;
	const uint32_t length = (1u << log2_length) + 77; // ragged on purpose
	cudaStream_t stream;
	CHECK_CUDA(cudaStreamCreate(&stream));
	bool ok = true;
	try {
		ok = run_length_encoding_round_trip(stream, length) && ok;
		ok = delta_of_variable_width_round_trip(stream, length) && ok;
	}
	catch (const std::exception& error) {
		std::fprintf(stderr, "exception: %s\n", error.what());
		return 3;
	}
	cudaStreamDestroy(stream);
	return ok ? 0 : 1;
}
