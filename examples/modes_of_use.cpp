// The reference's only known-answer test -- Frame of Reference, int32 out, int16 offsets, constant<4, int>
// model, length 2<<12, segment length 2<<10, offsets iota per segment, reference value 100 * segment index,
// expected out[s*L + i] == 100*s + i (reference examples/src/modes_of_use/common.cuh:29-105) -- run through
// this library in the reference's three modes of use (README.md:92-126):
//
//   (i)   direct use of the kernel: build the scheme's launch_config_resolution_params_t, call the free
//         resolver, launch the kernel function yourself      (examples/src/modes_of_use/direct_use_of_kernel.cu:8-38)
//   (ii)  instantiate the templated wrapper, typed overloads  (.../instantiation_of_wrapper.cu:7-40)
//   (iii) factory-produced, type-erased wrapper, any-maps     (.../factory_provided_type_erased_wrapper.cu:8-46)
//
// Host-only C++17: compiles with g++ against include/ and links libgiddy_b200.so + the CUDA runtime (for
// cudaMalloc/cudaMemcpy only).  Unlike the reference's example, all segments (the last one too) are
// checked, and only num_segments coefficients are copied.
// Exit code: 0 when all three modes reproduce the known answer.
#include <giddy/kernel_wrappers.hpp>

#include <cuda_runtime_api.h>

#include <cstdio>
#include <cstdlib>
#include <numeric>
#include <string>
#include <vector>

using uncompressed_type = int32_t;
using compressed_type   = int16_t;
using model_type        = cuda::functors::unary::parametric_model::constant<4, int32_t>;
using coefficients_type = model_type::coefficients_type;
constexpr unsigned index_size = 4;

namespace for_ns = cuda::kernels::decompression::frame_of_reference;

#define CHECK_CUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
	std::fprintf(stderr, "%s failed: %s\n", #call, cudaGetErrorString(e_)); std::exit(2); } } while (0)

struct device_buffers {
	uncompressed_type* decompressed;
	compressed_type*   compressed;
	coefficients_type* coefficients;
};

static void mode_i(const device_buffers& d, uint32_t length, uint32_t segment_length, cudaStream_t stream)
{
	auto properties = cuda::device::current::get().properties();
	for_ns::launch_config_resolution_params_t<index_size, uncompressed_type, compressed_type, model_type>
		params(properties, length, segment_length);
	auto config = cuda::kernels::resolve_launch_configuration(params);
	cuda::device_function_t kernel_function(
		for_ns::kernel_address<index_size, uncompressed_type, compressed_type, model_type>());
	uint32_t segments_per_block = 1; // a launch-shape hint of the reference's kernel signature; unused here
	cuda::enqueue_launch(kernel_function, stream, config,
		d.decompressed, (const compressed_type*) d.compressed, (const int32_t*) d.coefficients, length, segment_length, segments_per_block);
}

static void mode_ii(const device_buffers& d, uint32_t length, uint32_t segment_length, cudaStream_t stream)
{
	for_ns::kernel_t<index_size, uncompressed_type, compressed_type, model_type> kernel;
	auto properties = cuda::device::current::get().properties();
	auto attributes = kernel.get_device_function().attributes();
	auto config = kernel.resolve_launch_configuration(properties, attributes, (size_t) length, (size_t) segment_length);
	kernel.enqueue_launch(stream, config, d.decompressed, d.compressed, d.coefficients, length, segment_length);
}

static void mode_iii(const device_buffers& d, uint32_t length, uint32_t segment_length, cudaStream_t stream)
{
	std::string key = "decompression::frame_of_reference::kernel_t<4u, int, short, "
		"cuda::functors::unary::parametric_model::constant<4u, int> >";
	auto kernel = cuda::registered::kernel_t::produceSubclass(key);
	auto properties = cuda::device::current::get().properties();
	auto attributes = kernel->get_device_function().attributes();
	auto config = kernel->resolve_launch_configuration(properties, attributes, {
		{ "length",         cuda::any((size_t) length) },
		{ "segment_length", cuda::any((size_t) segment_length) },
	});
	kernel->enqueue_launch(stream, config, {
		{ "decompressed",               cuda::any(d.decompressed) },
		{ "compressed_input",           cuda::any((const compressed_type*) d.compressed) },
		{ "segment_model_coefficients", cuda::any((const coefficients_type*) d.coefficients) },
		{ "length",                     cuda::any(length) },
		{ "segment_length",             cuda::any(segment_length) },
	});
}

int main()
{
	const uint32_t length = 2 << 12, segment_length = 2 << 10;
	const uint32_t num_segments = length / segment_length;

	std::vector<compressed_type> compressed(length);
	std::vector<coefficients_type> coefficients(num_segments);
	for (uint32_t s = 0; s < num_segments; s++) {
		coefficients[s][0] = (int32_t)(s * 100);
		std::iota(compressed.begin() + s * segment_length, compressed.begin() + (s + 1) * segment_length, (compressed_type) 0);
	}

	device_buffers d;
	CHECK_CUDA(cudaMalloc((void**) &d.decompressed, length * sizeof(uncompressed_type)));
	CHECK_CUDA(cudaMalloc((void**) &d.compressed, length * sizeof(compressed_type)));
	CHECK_CUDA(cudaMalloc((void**) &d.coefficients, num_segments * sizeof(coefficients_type)));
	CHECK_CUDA(cudaMemcpy(d.compressed, compressed.data(), length * sizeof(compressed_type), cudaMemcpyHostToDevice));
	CHECK_CUDA(cudaMemcpy(d.coefficients, coefficients.data(), num_segments * sizeof(coefficients_type), cudaMemcpyHostToDevice));
	cudaStream_t stream;
	CHECK_CUDA(cudaStreamCreate(&stream));

	struct { const char* name; void (*run)(const device_buffers&, uint32_t, uint32_t, cudaStream_t); } modes[] = {
		{ "direct use of the kernel", mode_i }, { "instantiation of the wrapper", mode_ii },
		{ "factory-provided type-erased wrapper", mode_iii },
	};
	int failures = 0;
	for (auto& mode : modes) {
		CHECK_CUDA(cudaMemset(d.decompressed, 0xFF, length * sizeof(uncompressed_type)));
		try { mode.run(d, length, segment_length, stream); }
		catch (const std::exception& e) { std::printf("FAILURE  %s: exception: %s\n", mode.name, e.what()); failures++; continue; }
		CHECK_CUDA(cudaStreamSynchronize(stream));
		std::vector<uncompressed_type> decompressed(length);
		CHECK_CUDA(cudaMemcpy(decompressed.data(), d.decompressed, length * sizeof(uncompressed_type), cudaMemcpyDeviceToHost));
		uint32_t mismatches = 0;
		for (uint32_t s = 0; s < num_segments; s++) {
			for (uint32_t i = 0; i < segment_length; i++) {
				if (decompressed[s * segment_length + i] != (int32_t)(s * 100 + i)) { mismatches++; }
			}
		}
		std::printf("%s  %s (%u mismatches of %u)\n", mismatches ? "FAILURE" : "SUCCESS", mode.name, mismatches, length);
		failures += mismatches ? 1 : 0;
	}
	cudaFree(d.decompressed); cudaFree(d.compressed); cudaFree(d.coefficients);
	return failures ? 1 : 0;
}
