// C ABI launcher for the Delta decoder (include/giddy_b200.h, kernel API).
// Instantiations: src/kernel_wrappers/decompression/delta.cu:94-106.
#include "abi_internal.hpp"
#include "../kernels/delta.cuh"
#include <vector>

using namespace giddy;
using namespace giddy::abi;

namespace {

using delta_launcher_t = giddy_status_t (*)(void*, const giddy_launch_config_t*, void*, const void*, const void*, uint64_t, uint64_t);
using delta_long_launcher_t = giddy_status_t (*)(void*, const giddy_launch_config_t*, void*, const void*, const void*, uint64_t, uint64_t, void*, size_t);
struct delta_entry_t { unsigned I, U, C; const void* kernel; delta_launcher_t launch; delta_long_launcher_t launch_long; };

template <unsigned I, unsigned U, unsigned C>
giddy_status_t delta_launch(void* stream, const giddy_launch_config_t* cfg, void* out, const void* in, const void* anchors, uint64_t n, uint64_t L)
{
	return launch(kernels::delta_decompress<I, U, C>, stream, cfg,
		static_cast<uint_t<U>*>(out), static_cast<const uint_t<C>*>(in), static_cast<const uint_t<U>*>(anchors),
		(size_type<I>) n, (size_type<I>) L);
}
// How many CTAs of `block` threads of `kernel` fit the current device at once, if grids can be launched cooperatively there
// (0 if not).  Asked of the runtime once per (device, kernel, block size) and thread.
unsigned resident_blocks(const void* kernel, unsigned block, unsigned dynamic_shared_memory)
{
	struct key_t { int device; const void* kernel; unsigned block, shared; };
	struct entry_t { key_t key; unsigned resident; };
	static thread_local ::std::vector<entry_t> known;
	int device = 0;
	if (cudaGetDevice(&device) != cudaSuccess) { return 0; }
	for (const auto& e : known) { if (e.key.device == device && e.key.kernel == kernel && e.key.block == block && e.key.shared == dynamic_shared_memory) { return e.resident; } }
	int supported = 0, multiprocessors = 0, per_multiprocessor = 0;
	unsigned resident = 0;
	if (cudaDeviceGetAttribute(&supported, cudaDevAttrCooperativeLaunch, device) == cudaSuccess && supported != 0 &&
		cudaDeviceGetAttribute(&multiprocessors, cudaDevAttrMultiProcessorCount, device) == cudaSuccess &&
		ensure_dynamic_shared_memory(kernel, dynamic_shared_memory) == cudaSuccess &&
		cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_multiprocessor, kernel, (int) block, dynamic_shared_memory) == cudaSuccess) {
		resident = (unsigned) multiprocessors * (unsigned) per_multiprocessor;
	}
	else { cudaGetLastError(); }
	known.push_back({ { device, kernel, block, dynamic_shared_memory }, resident });
	return resident;
}

template <unsigned I, unsigned U, unsigned C>
giddy_status_t delta_long_launch(void* stream, const giddy_launch_config_t* cfg, void* out, const void* in, const void* anchors, uint64_t n, uint64_t L,
	void* workspace, size_t workspace_bytes)
{
	auto out_ = static_cast<uint_t<U>*>(out);
	auto in_ = static_cast<const uint_t<C>*>(in);
	auto anchors_ = static_cast<const uint_t<U>*>(anchors);
	size_type<I> n_ = (size_type<I>) n, L_ = (size_type<I>) L;
	cudaError_t error = cudaMemsetAsync(workspace, 0, workspace_bytes, static_cast<cudaStream_t>(stream));
	if (error != cudaSuccess) { return fail_cuda(error, "cudaMemsetAsync (Delta look-back workspace)"); }
	// Columns aligned for bulk copies: statically assigned tiles, the deltas staged in shared memory.  Every CTA must be
	// resident, which a cooperative launch of at most (CTAs per multiprocessor) x (multiprocessors) CTAs guarantees.
	if (kernels::look_back_shapes<U, C>::wide_pieces(out, in, L)) {
		const void* kernel = reinterpret_cast<const void*>(&kernels::delta_decompress_long_segments_staged<I, U, C>);
		const unsigned block = cfg->block_dimensions[0];
		const unsigned staging = (unsigned) GIDDY_DELTA_LOOK_BACK_SHARED_BYTES(U, C, block);
		const unsigned resident = resident_blocks(kernel, block, staging);
		if (resident > 0) {
			unsigned grid = cfg->grid_dimensions[0] < resident ? cfg->grid_dimensions[0] : resident;
			// whole waves: CTA b takes tiles b, b + grid, ..., so with T tiles in ceil(T / grid) waves, ceil(T / waves) CTAs
			// finish together where `grid` CTAs would leave most of the last wave idle
			const uint64_t tile_elements = GIDDY_DELTA_TILE_ELEMENTS(U, block);
			const uint64_t num_segments = (n + L - 1) / L;
			const uint64_t num_tiles = (num_segments - 1) * ((L + tile_elements - 1) / tile_elements) +
				((n - (num_segments - 1) * L) + tile_elements - 1) / tile_elements;
			if (num_tiles > grid) {
				const uint64_t waves = (num_tiles + grid - 1) / grid;
				grid = (unsigned) ((num_tiles + waves - 1) / waves);
			}
			void* arguments[] = { &out_, &in_, &anchors_, &n_, &L_, &workspace };
			error = cudaLaunchCooperativeKernel(kernel, dim3(grid), dim3(block), arguments, staging, static_cast<cudaStream_t>(stream));
			if (error == cudaSuccess) { return GIDDY_SUCCESS; }
			// (MPS / green-context SM limits can refuse a cooperative grid which the occupancy query allowed: fall back to the
			// ticket-counter kernel, which is correct under any residency)
			cudaGetLastError();
		}
	}
	return launch(kernels::delta_decompress_long_segments<I, U, C>, stream, cfg, out_, in_, anchors_, n_, L_, workspace);
}
#define DELTA_ENTRY(I, U, C) { I, U, C, reinterpret_cast<const void*>(&kernels::delta_decompress<I, U, C>), &delta_launch<I, U, C>, &delta_long_launch<I, U, C> }
#define DELTA_ENTRIES(I) DELTA_ENTRY(I, 2, 1), DELTA_ENTRY(I, 4, 1), DELTA_ENTRY(I, 4, 2), DELTA_ENTRY(I, 8, 1), DELTA_ENTRY(I, 8, 2), DELTA_ENTRY(I, 8, 4)

const delta_entry_t delta_table[] = { DELTA_ENTRIES(4), DELTA_ENTRIES(8) };

const delta_entry_t* find_delta(unsigned I, unsigned U, unsigned C)
{
	for (const auto& e : delta_table) { if (e.I == I && e.U == U && e.C == C) { return &e; } }
	return nullptr;
}

} // namespace

extern "C" const void* giddy_b200_delta_kernel(unsigned I, unsigned U, unsigned C)
{
	auto e = find_delta(I, U, C);
	return e ? e->kernel : nullptr;
}

extern "C" giddy_status_t giddy_b200_delta_launch(
	void* stream, const giddy_launch_config_t* config, unsigned I, unsigned U, unsigned C,
	void* decompressed, const void* compressed_input, const void* anchor_values, uint64_t length, uint64_t segment_length)
{
	auto e = find_delta(I, U, C);
	giddy_status_t status;
	if (!check_launch(config, e ? e->kernel : nullptr, status)) { return status; }
	if (segment_length == 0) { return fail(GIDDY_INVALID_ARGUMENT, "The segment length must be a positive number"); }
	if (length == 0) { return GIDDY_SUCCESS; }
	if (!GIDDY_DELTA_USES_LOOK_BACK(length, segment_length, U)) {
		return e->launch(stream, config, decompressed, compressed_input, anchor_values, length, segment_length);
	}
	// Few, long segments: tiles with decoupled look-back (kernels/delta.cuh).  The tile descriptors live in a
	// workspace taken from the stream-ordered allocator for the duration of this launch -- allocation, zeroing,
	// kernel and release are all enqueued on the caller's stream; nothing synchronises and nothing is kept.
	if (config->block_dimensions[0] % 32 != 0 || config->block_dimensions[0] == 0) { return fail(GIDDY_INVALID_ARGUMENT, "blocks must be made of whole warps"); }
	// the look-back kernel is compiled for CTAs of at most GIDDY_DELTA_LOOK_BACK_BLOCK_THREADS threads; its CTAs are
	// persistent and draw tiles from a ticket counter, so a larger requested block is simply cut down
	giddy_launch_config_t long_config = *config;
	if (long_config.block_dimensions[0] > GIDDY_DELTA_LOOK_BACK_BLOCK_THREADS) { long_config.block_dimensions[0] = GIDDY_DELTA_LOOK_BACK_BLOCK_THREADS; }
	const unsigned threads = long_config.block_dimensions[0];
	const unsigned E = 16 / U;
	const bool vectorizable = reinterpret_cast<uintptr_t>(decompressed) % 16 == 0 &&
		reinterpret_cast<uintptr_t>(compressed_input) % (E * C) == 0 && segment_length % E == 0;
	const uint64_t tile_elements = (uint64_t) threads * GIDDY_DELTA_LOOK_BACK_ROWS * (vectorizable ? E : 1);
	const uint64_t num_segments = (length + segment_length - 1) / segment_length;
	const uint64_t num_tiles = num_segments * ((segment_length + tile_elements - 1) / tile_elements); // an upper bound
	const size_t workspace_bytes = 32 + (size_t) num_tiles * 16;
	void* workspace = nullptr;
	cudaStream_t s = static_cast<cudaStream_t>(stream);
	cudaError_t error = allocate_workspace(&workspace, workspace_bytes, s);
	if (error != cudaSuccess) { return fail_cuda(error, "cudaMallocFromPoolAsync (Delta look-back workspace)"); }
	status = e->launch_long(stream, &long_config, decompressed, compressed_input, anchor_values, length, segment_length, workspace, workspace_bytes);
	error = cudaFreeAsync(workspace, s);
	if (status == GIDDY_SUCCESS && error != cudaSuccess) { return fail_cuda(error, "cudaFreeAsync (Delta look-back workspace)"); }
	return status;
}


// ---- Inclusive prefix sums (reference src/kernels/reduction/scan.cuh: reduce_segments :127 + scan_using_segment_reductions
// :412, or scan_single_segment :525 -- kernels the reference has but does not wrap) -----------------------------------------
// A prefix sum is a Delta column with a single segment whose anchor is the initial value: one launch of the look-back
// kernel (one pass over the data) where the reference needs two kernels and two passes, or a single block.

namespace { __device__ unsigned long long giddy_scan_zero_anchor[1] = { 0ull }; }

extern "C" const void* giddy_b200_inclusive_scan_kernel(unsigned I, unsigned result_size, unsigned input_size)
{
	return giddy_b200_delta_kernel(I, result_size, input_size);
}

extern "C" giddy_status_t giddy_b200_inclusive_scan_launch(
	void* stream, const giddy_launch_config_t* config, unsigned I, unsigned result_size, unsigned input_size,
	void* result, const void* data, uint64_t length, const void* initial_value)
{
	if (initial_value == nullptr) { // a zero of the library's own, on the current device
		void* zero = nullptr;
		cudaError_t error = cudaGetSymbolAddress(&zero, giddy_scan_zero_anchor);
		if (error != cudaSuccess) { return fail_cuda(error, "cudaGetSymbolAddress (scan initial value)"); }
		initial_value = zero;
	}
	if (length == 0) { return GIDDY_SUCCESS; }
	return giddy_b200_delta_launch(stream, config, I, result_size, input_size, result, data, initial_value, length, length);
}
