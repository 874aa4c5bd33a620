// C ABI launchers for the patched decoders: a position-aligned decoder and its patches in one pass
// (include/giddy_b200.h, kernel API; extensions).
#include "abi_internal.hpp"
#include "../kernels/patched.cuh"

using namespace giddy;
using namespace giddy::abi;

namespace {

// ---- Dictionary + index-list patches ------------------------------------------------------------------

using dict_launcher_t = giddy_status_t (*)(void*, const giddy_launch_config_t*, void*, const void*, const void*, uint64_t, uint64_t, bool,
	const void*, const void*, uint64_t);
struct dict_entry_t { unsigned I, U, DI; const void* kernel; dict_launcher_t launch; };

// The tile index (kernels/patched.cuh): log2 of the warp tile's element count for this pair of buffers, and the launch
// of patch_tile_ranges into `ranges` (num_tiles entries).
template <unsigned U>
unsigned log2_tile_elements(const void* decompressed, const void* compressed_input, unsigned compressed_element_size)
{
	constexpr unsigned E = 32 / U; // elements per 32-byte piece (kernels/patched.cuh)
	const unsigned in_bytes = E * compressed_element_size;
	const bool vectorizable = reinterpret_cast<uintptr_t>(decompressed) % 32 == 0 &&
		reinterpret_cast<uintptr_t>(compressed_input) % (in_bytes < 16 ? in_bytes : 16) == 0;
	unsigned elements = 32 * GIDDY_PATCHED_UNROLL * (vectorizable ? E : 1), log2 = 0;
	while ((1u << log2) < elements) { log2++; }
	return log2;
}

template <unsigned I>
size_t tile_index_bytes(uint64_t num_tiles) { return (size_t) num_tiles * sizeof(kernels::patched_detail::patch_range_t<I>); }

template <unsigned I>
giddy_status_t build_tile_index(cudaStream_t stream, void* ranges, const void* positions, uint64_t num_patches, unsigned log2_tile, uint64_t num_tiles, uint64_t length)
{
	// (the caller has filled the index with ones: tiles without patches read first == past_last)
	if (num_patches == 0) { return GIDDY_SUCCESS; }
	const unsigned threads = 256;
	const uint64_t blocks = (num_patches + 4 * threads - 1) / (4 * threads); // 4 patches per thread
	cudaError_t error;
	kernels::patched_detail::patch_tile_ranges<I><<<(unsigned) (blocks < 0x7fffffffull ? blocks : 0x7fffffffull), threads, 0, stream>>>(
		static_cast<kernels::patched_detail::patch_range_t<I>*>(ranges), static_cast<const uint_t<I>*>(positions), (size_type<I>) num_patches, log2_tile, length);
	error = cudaGetLastError();
	if (error != cudaSuccess) { return fail_cuda(error, "kernel launch (patch tile index)"); }
	return GIDDY_SUCCESS;
}

template <unsigned I, unsigned U, unsigned DI>
giddy_status_t dict_launch(void* stream, const giddy_launch_config_t* cfg, void* out, const void* in, const void* dict,
	uint64_t n, uint64_t num_entries, bool cache, const void* patch_data, const void* patch_indices, uint64_t num_patches)
{
	cudaStream_t s = static_cast<cudaStream_t>(stream);
	const unsigned log2_tile = log2_tile_elements<U>(out, in, DI);
	const uint64_t num_tiles = (n + (1ull << log2_tile) - 1) >> log2_tile;
	void* offsets = nullptr;
	cudaError_t error = allocate_workspace(&offsets, tile_index_bytes<I>(num_tiles) + 16, s);
	if (error != cudaSuccess) { return fail_cuda(error, "cudaMallocFromPoolAsync (patch tile index)"); }
	error = cudaMemsetAsync(offsets, 0xff, tile_index_bytes<I>(num_tiles), s);
	if (error != cudaSuccess) { cudaFreeAsync(offsets, s); return fail_cuda(error, "cudaMemsetAsync (patch tile index)"); }
	giddy_status_t status = build_tile_index<I>(s, offsets, patch_indices, num_patches, log2_tile, num_tiles, n);
	if (status == GIDDY_SUCCESS) {
		status = launch_as_dependent(kernels::patched_dictionary_decompress<I, U, DI>, stream, cfg,
			static_cast<uint_t<U>*>(out), static_cast<const uint_t<DI>*>(in), static_cast<const uint_t<U>*>(dict),
			(size_type<I>) n, (size_type<DI>) num_entries, cache,
			static_cast<const uint_t<U>*>(patch_data), static_cast<const uint_t<I>*>(patch_indices),
			static_cast<const kernels::patched_detail::patch_range_t<I>*>(offsets), log2_tile);
	}
	error = cudaFreeAsync(offsets, s);
	if (status == GIDDY_SUCCESS && error != cudaSuccess) { return fail_cuda(error, "cudaFreeAsync (patch tile index)"); }
	return status;
}
#define DICT_ENTRY(I, U, DI) { I, U, DI, reinterpret_cast<const void*>(&kernels::patched_dictionary_decompress<I, U, DI>), &dict_launch<I, U, DI> }

const dict_entry_t dict_table[] = {
	DICT_ENTRY(4, 1, 1), DICT_ENTRY(4, 2, 1), DICT_ENTRY(4, 4, 1), DICT_ENTRY(4, 8, 1),
	DICT_ENTRY(4, 2, 2), DICT_ENTRY(4, 4, 2), DICT_ENTRY(4, 8, 2), DICT_ENTRY(4, 8, 4), DICT_ENTRY(8, 8, 1), DICT_ENTRY(8, 4, 2),
};

const dict_entry_t* find_dict(unsigned I, unsigned U, unsigned DI)
{
	for (const auto& e : dict_table) { if (e.I == I && e.U == U && e.DI == DI) { return &e; } }
	return nullptr;
}

// ---- Frame of Reference + compressed-indices patches -----------------------------------------------------

using for_launcher_t = giddy_status_t (*)(void*, const giddy_launch_config_t*, void*, const void*, const void*, uint64_t, uint64_t,
	const void*, const void*, const void*, const void*, const void*, const void*, const void*, uint64_t, uint64_t, uint64_t);
struct for_entry_t { unsigned I, U, C; int model; unsigned MI, II; const void* kernel; for_launcher_t launch; };

template <unsigned I, unsigned U, unsigned C, int M, unsigned MI, unsigned II>
giddy_status_t for_launch(void* stream, const giddy_launch_config_t* cfg, void* out, const void* in, const void* coeffs, uint64_t n, uint64_t L,
	const void* data, const void* run_lengths, const void* offset_sizes, const void* baselines, const void* offsets_start,
	const void* offset_bytes, const void* anchors, uint64_t anchoring_period, uint64_t num_runs, uint64_t data_length)
{
	cudaStream_t s = static_cast<cudaStream_t>(stream);
	const unsigned log2_tile = log2_tile_elements<U>(out, in, C);
	const uint64_t num_tiles = (n + (1ull << log2_tile) - 1) >> log2_tile;
	// workspace: the tile index, then the decoded patch positions
	const size_t offsets_bytes = (tile_index_bytes<I>(num_tiles) + 15) / 16 * 16;
	void* workspace = nullptr;
	cudaError_t error = allocate_workspace(&workspace, offsets_bytes + data_length * I + 16, s);
	if (error != cudaSuccess) { return fail_cuda(error, "cudaMallocFromPoolAsync (patch positions and tile index)"); }
	void* positions = static_cast<unsigned char*>(workspace) + offsets_bytes;
	giddy_status_t status = GIDDY_SUCCESS;
	// index entries of all ones are empty ranges; positions of all ones lie beyond any column (runs which cover fewer than
	// data_length patches leave the rest of the list never applied, as the reference's kernel would).  The fused kernel fills
	// the uncovered positions itself; the generic path needs the whole list filled beforehand.
	const bool fused_index = data_length != 0 && anchoring_period * I <= 32 * 1024;
	if (fused_index) { // (a kernel, so that the position decode can run as its programmatic dependent)
		const uint64_t vectors = offsets_bytes / 16;
		kernels::patched_detail::fill_tile_index<><<<(unsigned) ((vectors + 255) / 256 < 1184 ? (vectors + 255) / 256 : 1184), 256, 0, s>>>(static_cast<uint4*>(workspace), vectors);
		error = cudaGetLastError();
	}
	else { error = cudaMemsetAsync(workspace, 0xff, offsets_bytes + data_length * I, s); }
	if (error != cudaSuccess) { cudaFreeAsync(workspace, s); return fail_cuda(error, "filling the patch positions and tile index"); }
	bool index_built = false;
	if (data_length != 0) {
		const uint64_t periods = (data_length + anchoring_period - 1) / anchoring_period;
		// periods whose decoded positions fit 32 KiB keep them in shared memory and build the tile index in the same kernel
		// (one launch and one pass over the positions less); longer ones leave it to patch_tile_ranges
		if (fused_index) {
			giddy_launch_config_t index_config = { };
			index_config.grid_dimensions[0] = (unsigned) (periods < 0x7fffffffull ? periods : 0x7fffffffull);
			index_config.grid_dimensions[1] = index_config.grid_dimensions[2] = 1;
			index_config.block_dimensions[0] = GIDDY_PATCHED_INDEX_BLOCK_THREADS;
			index_config.block_dimensions[1] = index_config.block_dimensions[2] = 1;
			index_config.dynamic_shared_memory_size = (unsigned)(anchoring_period * I);
			status = launch_as_dependent(kernels::patched_detail::expand_and_index_compressed_indices<I, II, II>, stream, &index_config,
				static_cast<uint_t<I>*>(positions), static_cast<kernels::patched_detail::patch_range_t<I>*>(workspace),
				static_cast<const uint_t<II>*>(run_lengths), static_cast<const unsigned char*>(offset_sizes),
				static_cast<const uint_t<I>*>(baselines), static_cast<const uint_t<II>*>(offsets_start),
				static_cast<const unsigned char*>(offset_bytes), static_cast<const uint_t<II>*>(anchors),
				(unsigned) anchoring_period, (uint64_t) num_runs, (uint64_t) data_length, log2_tile, (uint64_t) n);
		}
		else {
			const unsigned warps_per_block = GIDDY_CI_SCATTER_WARPS_PER_BLOCK;
			const uint64_t blocks = (periods + warps_per_block - 1) / warps_per_block;
			kernels::expand_compressed_indices<I, II, II><<<(unsigned) (blocks < 0x7fffffffull ? blocks : 0x7fffffffull), warps_per_block * 32, 0, s>>>(
				static_cast<uint_t<I>*>(positions), static_cast<const uint_t<II>*>(run_lengths), static_cast<const unsigned char*>(offset_sizes),
				static_cast<const uint_t<I>*>(baselines), static_cast<const uint_t<II>*>(offsets_start),
				static_cast<const unsigned char*>(offset_bytes), static_cast<const uint_t<II>*>(anchors),
				(uint_t<II>) anchoring_period, (uint_t<II>) num_runs, (size_type<II>) data_length);
		}
		error = cudaGetLastError();
		if (error != cudaSuccess) { status = fail_cuda(error, "kernel launch (patch position decode)"); }
		index_built = fused_index;
	}
	if (status == GIDDY_SUCCESS && !index_built) { status = build_tile_index<I>(s, workspace, positions, data_length, log2_tile, num_tiles, n); }
	if (status == GIDDY_SUCCESS) {
		status = launch_as_dependent(kernels::patched_frame_of_reference_decompress<I, U, C, M, MI>, stream, cfg,
			static_cast<uint_t<U>*>(out), static_cast<const int_t<C>*>(in), static_cast<const int32_t*>(coeffs),
			(size_type<I>) n, (size_type<I>) L,
			static_cast<const uint_t<U>*>(data), static_cast<const uint_t<I>*>(positions), static_cast<const kernels::patched_detail::patch_range_t<I>*>(workspace), log2_tile);
	}
	error = cudaFreeAsync(workspace, s);
	if (status == GIDDY_SUCCESS && error != cudaSuccess) { return fail_cuda(error, "cudaFreeAsync (patch positions and tile index)"); }
	return status;
}
#define FOR_ENTRY(I, U, C, M, MI, II) { I, U, C, M, MI, II, \
	reinterpret_cast<const void*>(&kernels::patched_frame_of_reference_decompress<I, U, C, M, MI>), &for_launch<I, U, C, M, MI, II> }

const for_entry_t for_table[] = {
	FOR_ENTRY(4, 2, 1, model_linear,   4, 4), FOR_ENTRY(4, 4, 1, model_linear,   4, 4), FOR_ENTRY(4, 4, 2, model_linear,   4, 4),
	FOR_ENTRY(8, 4, 2, model_linear,   8, 8),
	FOR_ENTRY(4, 2, 1, model_constant, 4, 4), FOR_ENTRY(4, 4, 1, model_constant, 4, 4), FOR_ENTRY(4, 4, 2, model_constant, 4, 4),
};

const for_entry_t* find_for(unsigned I, unsigned U, unsigned C, int model, unsigned MI, unsigned II)
{
	for (const auto& e : for_table) { if (e.I == I && e.U == U && e.C == C && e.model == model && e.MI == MI && e.II == II) { return &e; } }
	return nullptr;
}

giddy_status_t check_block(const giddy_launch_config_t* config)
{
	if (config->block_dimensions[0] % 32 != 0 || config->block_dimensions[0] == 0 || config->block_dimensions[0] > GIDDY_ELEMENTWISE_BLOCK_THREADS) {
		return fail(GIDDY_INVALID_ARGUMENT, "blocks must be made of 1 to 8 whole warps");
	}
	return GIDDY_SUCCESS;
}

} // namespace

extern "C" const void* giddy_b200_patched_dictionary_kernel(unsigned I, unsigned U, unsigned DI)
{
	auto e = find_dict(I, U, DI);
	return e ? e->kernel : nullptr;
}

extern "C" giddy_status_t giddy_b200_patched_dictionary_launch(
	void* stream, const giddy_launch_config_t* config, unsigned I, unsigned U, unsigned DI,
	void* decompressed, const void* compressed_input, const void* dictionary_entries,
	uint64_t length, uint64_t num_dictionary_entries, int cache_dictionary_in_shared_memory,
	const void* patch_data, const void* patch_indices, uint64_t num_patches)
{
	auto e = find_dict(I, U, DI);
	giddy_status_t status;
	if (!check_launch(config, e ? e->kernel : nullptr, status)) { return status; }
	if ((status = check_block(config)) != GIDDY_SUCCESS) { return status; }
	if (length == 0) { return GIDDY_SUCCESS; }
	if (cache_dictionary_in_shared_memory && (uint64_t) config->dynamic_shared_memory_size < num_dictionary_entries * U) {
		return fail(GIDDY_INVALID_ARGUMENT, "caching the dictionary in shared memory needs num_dictionary_entries * sizeof(element) "
			"bytes of dynamic shared memory in the launch configuration");
	}
	return e->launch(stream, config, decompressed, compressed_input, dictionary_entries, length, num_dictionary_entries,
		cache_dictionary_in_shared_memory != 0, patch_data, patch_indices, num_patches);
}

extern "C" const void* giddy_b200_patched_frame_of_reference_kernel(
	unsigned I, unsigned U, unsigned C, giddy_model_t model, unsigned MI, unsigned II)
{
	auto e = find_for(I, U, C, (int) model, MI, II);
	return e ? e->kernel : nullptr;
}

extern "C" giddy_status_t giddy_b200_patched_frame_of_reference_launch(
	void* stream, const giddy_launch_config_t* config,
	unsigned I, unsigned U, unsigned C, giddy_model_t model, unsigned MI, unsigned II,
	void* decompressed, const void* compressed_input, const void* segment_model_coefficients,
	uint64_t length, uint64_t segment_length,
	const void* data_to_scatter,
	const void* scatter_position_run_lengths, const void* scatter_position_run_individual_offset_sizes,
	const void* scatter_position_run_baseline_values, const void* scatter_position_run_offsets_start_positions,
	const void* scatter_position_offset_bytes, const void* scatter_position_anchors,
	uint64_t anchoring_period, uint64_t num_scatter_position_runs, uint64_t data_length)
{
	auto e = find_for(I, U, C, (int) model, MI, II);
	giddy_status_t status;
	if (!check_launch(config, e ? e->kernel : nullptr, status)) { return status; }
	if ((status = check_block(config)) != GIDDY_SUCCESS) { return status; }
	if (segment_length == 0) { return fail(GIDDY_INVALID_ARGUMENT, "The segment length must be a positive number"); }
	if (data_length != 0 && anchoring_period == 0) { return fail(GIDDY_INVALID_ARGUMENT, "The anchoring period must be a positive number"); }
	if (length == 0) { return GIDDY_SUCCESS; }
	return e->launch(stream, config, decompressed, compressed_input, segment_model_coefficients, length, segment_length,
		data_to_scatter, scatter_position_run_lengths, scatter_position_run_individual_offset_sizes,
		scatter_position_run_baseline_values, scatter_position_run_offsets_start_positions, scatter_position_offset_bytes,
		scatter_position_anchors, anchoring_period, num_scatter_position_runs, data_length);
}
