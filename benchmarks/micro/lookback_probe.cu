// Micro-benchmark of the Delta look-back kernel (kernels/delta.cuh) under different compile-time settings
// (-DGIDDY_DELTA_LOOK_BACK_ROWS / _WINDOW / _PREFETCH_TICKET / _BLOCKS_PER_MULTIPROCESSOR): one long segment,
// checked against a host prefix sum, timed with CUDA events.  Development tool, not part of the library.
//   nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -Iinclude -Iinclude/giddy -D... benchmarks/micro/lookback_probe.cu -o probe
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <random>
#include <algorithm>
#include "../../libgiddy_b200/csrc/kernels/delta.cuh"

#define CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s failed: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

template <unsigned U, unsigned C>
void run(size_t n, int blocks_per_sm, int static_tiles, const char* label)
{
	using namespace giddy;
	std::vector<uint_t<C>> deltas(n);
	std::mt19937_64 rng(1);
	for (auto& d : deltas) { d = (uint_t<C>) rng(); }
	uint_t<C>* in; uint_t<U>* out; uint_t<U>* anchor; void* workspace;
	CHECK(cudaMalloc(&in, n * C)); CHECK(cudaMalloc(&out, n * U)); CHECK(cudaMalloc(&anchor, U));
	const uint_t<U> a = 12345;
	CHECK(cudaMemcpy(in, deltas.data(), n * C, cudaMemcpyHostToDevice));
	CHECK(cudaMemcpy(anchor, &a, U, cudaMemcpyHostToDevice));
	const unsigned threads = GIDDY_DELTA_LOOK_BACK_BLOCK_THREADS;
	const size_t tile = GIDDY_DELTA_TILE_ELEMENTS(U, threads);
	const size_t tiles = (n + tile - 1) / tile;
	const size_t workspace_bytes = 32 + tiles * 16;
	CHECK(cudaMalloc(&workspace, workspace_bytes));
	int sms; CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
	const unsigned grid = sms * blocks_per_sm;
	uint32_t n32 = (uint32_t) n;
	auto launch = [&]() {
		CHECK(cudaMemsetAsync(workspace, 0, workspace_bytes));
		if (static_tiles) {
			void* arguments[] = { &out, &in, &anchor, &n32, &n32, &workspace };
			const size_t staging = GIDDY_DELTA_LOOK_BACK_SHARED_BYTES(U, C, threads);
			const void* kernel = reinterpret_cast<const void*>(kernels::delta_decompress_long_segments_staged<4, U, C>);
			CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) staging));
			if (static_tiles == 2) { // a plain launch of the same grid (all of it resident here): what the cooperative launch itself costs
				kernels::delta_decompress_long_segments_staged<4, U, C><<<grid, threads, staging>>>(out, in, anchor, n32, n32, workspace);
			}
			else { CHECK(cudaLaunchCooperativeKernel(kernel, dim3(grid), dim3(threads), arguments, staging, 0)); }
		}
		else { kernels::delta_decompress_long_segments<4, U, C><<<grid, threads>>>(out, in, anchor, n32, n32, workspace); }
	};
	launch();
	CHECK(cudaDeviceSynchronize());
	std::vector<uint_t<U>> got(n);
	CHECK(cudaMemcpy(got.data(), out, n * U, cudaMemcpyDeviceToHost));
	uint_t<U> sum = a; size_t bad = 0;
	for (size_t i = 0; i < n; i++) { sum = (uint_t<U>)(sum + deltas[i]); if (got[i] != sum) { bad++; } }
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	for (int i = 0; i < 3; i++) { launch(); }
	cudaEventRecord(e0);
	const int reps = 20;
	for (int i = 0; i < reps; i++) { launch(); }
	cudaEventRecord(e1);
	CHECK(cudaDeviceSynchronize());
	float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
	{ // per-launch times: is there a slow mode?
		std::vector<float> each;
		for (int i = 0; i < 60; i++) {
			cudaEventRecord(e0); launch(); cudaEventRecord(e1); CHECK(cudaDeviceSynchronize());
			float t; cudaEventElapsedTime(&t, e0, e1); each.push_back(t);
		}
		std::sort(each.begin(), each.end());
		printf("  single launches (ms): min %.3f  p25 %.3f  median %.3f  p75 %.3f  p90 %.3f  max %.3f\n", each[0], each[15], each[30], each[45], each[54], each[59]);
	}
	printf("%-12s static=%d U=%u C=%u n=%zu tiles=%zu grid=%u: %.3f ms  %.0f GB/s algorithmic  mismatches=%zu\n", label, (int) static_tiles, U, C, n, tiles, grid, ms,
		(double) n * (U + C) / ms / 1e6, bad);
#ifdef GIDDY_DELTA_TRACE
	{
		launch();
		CHECK(cudaDeviceSynchronize());
		std::vector<unsigned long long> trace(tiles * 8);
		CHECK(cudaMemcpyFromSymbol(trace.data(), kernels::delta_detail::trace, tiles * 8 * sizeof(unsigned long long)));
		unsigned long long t_min = ~0ull;
		for (size_t t = 0; t < tiles; t++) { if (trace[t * 8] < t_min) { t_min = trace[t * 8]; } }
		double sums[5] = { 0, 0, 0, 0, 0 }, polls = 0, rounds = 0;
		for (size_t t = 1; t < tiles; t++) {
			for (int k = 0; k < 5; k++) { sums[k] += (double)(trace[t * 8 + k + 1] - trace[t * 8 + k]); }
			polls += trace[t * 8 + 6]; rounds += trace[t * 8 + 7];
		}
		printf("  per tile (ns): look-back (thread 0) %.0f | reduce next %.0f | barrier %.0f | publish next %.0f | scan+store %.0f ; polls %.2f rounds %.2f\n",
			sums[0] / tiles, sums[1] / tiles, sums[2] / tiles, sums[3] / tiles, sums[4] / tiles, polls / tiles, rounds / tiles);
		for (size_t t : { (size_t) 1, (size_t) 300, (size_t) 600, (size_t) 5000, (size_t) 5001, (size_t) 5002, (size_t) 5130, (size_t) 5300, (size_t) 5592 }) {
			if (t >= tiles) { continue; }
			printf("  tile %5zu: start %8llu", t, trace[t * 8] - t_min);
			for (int k = 1; k < 6; k++) { printf(" +%6llu", trace[t * 8 + k] - trace[t * 8 + k - 1]); }
			printf("  polls %llu rounds %llu\n", trace[t * 8 + 6], trace[t * 8 + 7]);
		}
	}
#endif
	cudaFree(in); cudaFree(out); cudaFree(anchor); cudaFree(workspace);
}

int main(int argc, char** argv)
{
	const char* label = argc > 1 ? argv[1] : "";
	const int blocks_per_sm = argc > 2 ? atoi(argv[2]) : GIDDY_DELTA_LOOK_BACK_BLOCKS_PER_MULTIPROCESSOR;
	const int static_tiles = argc > 3 ? atoi(argv[3]) : 1;
	run<8, 4>((size_t) 1 << 26, blocks_per_sm, static_tiles, label);
	run<4, 1>((size_t) 1 << 27, blocks_per_sm, static_tiles, label);
	return 0;
}
