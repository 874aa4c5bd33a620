// C ABI launchers of the compressors (include/giddy_b200.h, group 8; kernels/encoders.cuh).
// The reference has no compressors (README.md:128-130): each launcher cites the decompressor whose format it emits.
#include "abi_internal.hpp"
#include "../kernels/encoders.cuh"

using namespace giddy;
using namespace giddy::abi;
namespace enc = giddy::kernels::encoders;

namespace {

cudaStream_t as_stream(void* stream) { return static_cast<cudaStream_t>(stream); }

// helper kernels size their own grids: enough CTAs of 256 threads to fill the device a few times over, never more than the work
unsigned helper_blocks(uint64_t work_items, unsigned items_per_block)
{
	static int multiprocessors[64] = { };
	int device = 0;
	cudaGetDevice(&device);
	if (device < 0 || device >= 64) { device = 0; }
	if (multiprocessors[device] == 0) { cudaDeviceGetAttribute(&multiprocessors[device], cudaDevAttrMultiProcessorCount, device); }
	const uint64_t wanted = (work_items + items_per_block - 1) / items_per_block;
	const uint64_t most = (uint64_t)(multiprocessors[device] > 0 ? multiprocessors[device] : 148) * 16;
	return (unsigned)(wanted < 1 ? 1 : (wanted < most ? wanted : most));
}

giddy_status_t launched(const char* what)
{
	const cudaError_t error = cudaGetLastError();
	return error == cudaSuccess ? GIDDY_SUCCESS : fail_cuda(error, what);
}

struct workspace_t { // stream-ordered scratch, released when the launcher returns (i.e. after the kernels, in stream order)
	void* memory = nullptr; cudaStream_t stream;
	explicit workspace_t(cudaStream_t s) : stream(s) { }
	cudaError_t take(size_t bytes) { return allocate_workspace(&memory, bytes, stream); }
	~workspace_t() { if (memory) { cudaFreeAsync(memory, stream); } }
};

// ---- tables ---------------------------------------------------------------------------------------------

using uc_launcher_t = giddy_status_t (*)(void*, const giddy_launch_config_t*, void*, void*, const void*, uint64_t, uint64_t, uint64_t, unsigned*);
struct uc_entry_t { unsigned I, U, C; const void* delta_kernel; uc_launcher_t delta; const void* dzbf_kernel; uc_launcher_t dzbf; };

template <unsigned I, unsigned U, unsigned C>
giddy_status_t delta_launch(void* stream, const giddy_launch_config_t* cfg, void* compressed, void* anchors, const void* in, uint64_t n, uint64_t L, uint64_t base, unsigned* status)
{
	return launch(enc::delta_compress<I, U, C>, stream, cfg, static_cast<uint_t<C>*>(compressed), static_cast<uint_t<U>*>(anchors),
		static_cast<const uint_t<U>*>(in), (size_type<I>) n, (size_type<I>) L, (uint_t<U>) base, status);
}
template <unsigned I, unsigned U, unsigned C>
giddy_status_t dzbf_launch(void* stream, const giddy_launch_config_t* cfg, void* compressed, void*, const void* in, uint64_t n, uint64_t, uint64_t, unsigned* status)
{
	return launch(enc::discard_zero_bytes_fixed_compress<I, U, C>, stream, cfg, static_cast<uint_t<C>*>(compressed),
		static_cast<const uint_t<U>*>(in), (size_type<I>) n, status);
}
#define UC_ENTRY(I, U, C) { I, U, C, reinterpret_cast<const void*>(&enc::delta_compress<I, U, C>), &delta_launch<I, U, C>, \
	reinterpret_cast<const void*>(&enc::discard_zero_bytes_fixed_compress<I, U, C>), &dzbf_launch<I, U, C> }
#define UC_ENTRIES(I) UC_ENTRY(I, 2, 1), UC_ENTRY(I, 4, 1), UC_ENTRY(I, 4, 2), UC_ENTRY(I, 8, 1), UC_ENTRY(I, 8, 2), UC_ENTRY(I, 8, 4)
const uc_entry_t uc_table[] = { UC_ENTRIES(4), UC_ENTRIES(8) };
const uc_entry_t* find_uc(unsigned I, unsigned U, unsigned C)
{
	for (const auto& e : uc_table) { if (e.I == I && e.U == U && e.C == C) { return &e; } }
	return nullptr;
}

using dict_launcher_t = giddy_status_t (*)(void*, const giddy_launch_config_t*, void*, const void*, const void*, uint64_t, uint64_t, unsigned*);
struct dict_entry_t { unsigned I, U, DI; const void* kernel; dict_launcher_t launch; };
template <unsigned I, unsigned U, unsigned DI>
giddy_status_t dict_launch(void* stream, const giddy_launch_config_t* cfg, void* compressed, const void* in, const void* entries, uint64_t num_entries, uint64_t n, unsigned* status)
{
	return launch(enc::dictionary_compress<I, U, DI>, stream, cfg, static_cast<uint_t<DI>*>(compressed), static_cast<const uint_t<U>*>(in),
		static_cast<const uint_t<U>*>(entries), (size_type<DI>) num_entries, (size_type<I>) n, status);
}
#define DICT_ENTRY(I, U, DI) { I, U, DI, reinterpret_cast<const void*>(&enc::dictionary_compress<I, U, DI>), &dict_launch<I, U, DI> }
#define DICT_ENTRIES(I) DICT_ENTRY(I, 1, 1), DICT_ENTRY(I, 2, 1), DICT_ENTRY(I, 2, 2), DICT_ENTRY(I, 4, 1), DICT_ENTRY(I, 4, 2), DICT_ENTRY(I, 4, 4), \
	DICT_ENTRY(I, 8, 1), DICT_ENTRY(I, 8, 2), DICT_ENTRY(I, 8, 4), DICT_ENTRY(I, 8, 8)
const dict_entry_t dict_table[] = { DICT_ENTRIES(4), DICT_ENTRIES(8) };
const dict_entry_t* find_dict(unsigned I, unsigned U, unsigned DI)
{
	for (const auto& e : dict_table) { if (e.I == I && e.U == U && e.DI == DI) { return &e; } }
	return nullptr;
}

using for_launcher_t = giddy_status_t (*)(void*, const giddy_launch_config_t*, void*, void*, const void*, uint64_t, uint64_t, unsigned*);
struct for_entry_t { unsigned I, U, C; int model; const void* kernel; for_launcher_t launch; };
template <unsigned I, unsigned U, unsigned C, int M>
giddy_status_t for_launch(void* stream, const giddy_launch_config_t* cfg, void* compressed, void* coefficients, const void* in, uint64_t n, uint64_t L, unsigned* status)
{
	constexpr unsigned E = 16 / U, dim = (M == model_linear) ? 2 : 1;
	auto out = static_cast<int_t<C>*>(compressed);
	auto coefficients_ = static_cast<int32_t*>(coefficients);
	auto x = static_cast<const int_t<U>*>(in);
	// Segments of a few KB to a few tens of KB, aligned: a CTA per segment holds it in registers between the fit and the store
	// (one read of the column); a short last segment goes through the general kernel.
	const uint64_t vectors = L / E, whole_segments = L ? n / L : 0;
	const bool by_block = L % E == 0 && vectors > enc::for_block_threads && vectors <= (uint64_t) enc::for_block_threads * enc::for_block_vectors &&
		reinterpret_cast<uintptr_t>(in) % 16 == 0 && reinterpret_cast<uintptr_t>(compressed) % (E * C) == 0 && whole_segments > 0 &&
		cfg->grid_dimensions[0] != 0 && cfg->block_dimensions[0] != 0;
	if (!by_block) {
		return launch(enc::frame_of_reference_compress<I, U, C, M>, stream, cfg, out, coefficients_, x, (size_type<I>) n, (size_type<I>) L, status);
	}
	const unsigned blocks = helper_blocks(whole_segments, 1);
	enc::frame_of_reference_compress_by_block<I, U, C, M><<<blocks, enc::for_block_threads, 0, as_stream(stream)>>>(out, coefficients_, x,
		(size_type<I>)(whole_segments * L), (size_type<I>) L, status);
	giddy_status_t result = launched("kernel launch (frame of reference, a block per segment)");
	if (result != GIDDY_SUCCESS || whole_segments * L == n) { return result; }
	giddy_launch_config_t one_warp = *cfg;
	one_warp.grid_dimensions[0] = 1; one_warp.block_dimensions[0] = 32;
	return launch(enc::frame_of_reference_compress<I, U, C, M>, stream, &one_warp, out + whole_segments * L, coefficients_ + whole_segments * dim,
		x + whole_segments * L, (size_type<I>)(n - whole_segments * L), (size_type<I>) L, status);
}
#define FOR_ENTRY(I, U, C, M) { I, U, C, M, reinterpret_cast<const void*>(&enc::frame_of_reference_compress<I, U, C, M>), &for_launch<I, U, C, M> }
const for_entry_t for_table[] = {
	FOR_ENTRY(4, 2, 1, model_linear), FOR_ENTRY(4, 4, 1, model_linear), FOR_ENTRY(4, 4, 2, model_linear), FOR_ENTRY(8, 4, 2, model_linear),
	FOR_ENTRY(4, 2, 1, model_constant), FOR_ENTRY(4, 4, 1, model_constant), FOR_ENTRY(4, 4, 2, model_constant), FOR_ENTRY(8, 4, 2, model_constant),
};
const for_entry_t* find_for(unsigned I, unsigned U, unsigned C, int model)
{
	for (const auto& e : for_table) { if (e.I == I && e.U == U && e.C == C && e.model == model) { return &e; } }
	return nullptr;
}

// ---- run-based encoders -----------------------------------------------------------------------------------

// what RLE and RPE share: summaries, placement, emission of values and start positions into the workspace
struct run_plan_t {
	uint64_t num_tiles;
	enc::run_tile_summary_t* summaries;
	uint64_t* placement;       // num_tiles + 1
	uint64_t* open_run_start;  // num_tiles
	uint64_t* starts;          // run_capacity + 1
	uint64_t num_chunks;       // chunks of 1024 tiles
	uint64_t* chunk_open;      // num_chunks: the chunks' last natural heads, then their exclusive scan
	uint64_t* chunk_totals;    // num_chunks + 1: the chunks' run counts, then their exclusive scan; [num_chunks] = the number of runs
	uint64_t* total;           // where the emission leaves the number of runs for the kernels after it
};

size_t run_workspace_bytes(uint64_t n, uint64_t run_capacity, run_plan_t& plan)
{
	plan.num_tiles = (n + enc::run_tile_elements - 1) / enc::run_tile_elements;
	plan.num_chunks = (plan.num_tiles + enc::placement_chunk_tiles - 1) / enc::placement_chunk_tiles;
	return plan.num_tiles * sizeof(enc::run_tile_summary_t) + (plan.num_tiles + 1) * 8 + plan.num_tiles * 8 + (run_capacity + 1) * 8 + (2 * plan.num_chunks + 1) * 8 + 64;
}

void run_workspace_layout(void* memory, uint64_t run_capacity, run_plan_t& plan)
{
	auto* p = static_cast<unsigned char*>(memory);
	plan.summaries = reinterpret_cast<enc::run_tile_summary_t*>(p); p += plan.num_tiles * sizeof(enc::run_tile_summary_t);
	plan.placement = reinterpret_cast<uint64_t*>(p); p += (plan.num_tiles + 1) * 8;
	plan.open_run_start = reinterpret_cast<uint64_t*>(p); p += plan.num_tiles * 8;
	plan.starts = reinterpret_cast<uint64_t*>(p); p += (run_capacity + 1) * 8;
	plan.chunk_open = reinterpret_cast<uint64_t*>(p); p += plan.num_chunks * 8;
	plan.chunk_totals = reinterpret_cast<uint64_t*>(p);
	plan.total = plan.placement + plan.num_tiles;
}

template <unsigned U>
giddy_status_t run_common(cudaStream_t s, const giddy_launch_config_t* cfg, run_plan_t& plan, void* run_data, const void* in, uint64_t n,
	uint64_t max_run_length, uint64_t run_capacity, unsigned* status)
{
	const unsigned blocks = cfg->grid_dimensions[0] ? cfg->grid_dimensions[0] : helper_blocks(plan.num_tiles, 1);
	auto x = static_cast<const uint_t<U>*>(in);
	// runs limited to at least a tile's length (or not at all) have at most one split point per tile: the mask-based kernels
	const bool fast = max_run_length == 0 || max_run_length >= enc::run_tile_elements;
	if (GIDDY_RUN_ENCODERS_SINGLE_PASS && (max_run_length == 0 || max_run_length >= enc::run_super_elements)) {
		// ONE pass over the column: a super-tile's first run index and open run come from its predecessors' descriptors
		// (decoupled look-back), which live where the two-pass path keeps its summaries: 32 bytes per super-tile, then the run
		// count and the ticket counter, all zeroed
		const uint64_t num_super_tiles = (n + enc::run_super_elements - 1) / enc::run_super_elements;
		cudaError_t error = cudaMemsetAsync(plan.summaries, 0, num_super_tiles * 32 + 16, s);
		if (error != cudaSuccess) { return fail_cuda(error, "cudaMemsetAsync (run encoder tile descriptors)"); }
		const unsigned enough = helper_blocks(num_super_tiles, 1);
		enc::run_encode_single_pass<U, uint64_t><<<blocks < enough ? blocks : enough, enc::run_tile_threads, 0, s>>>(static_cast<uint_t<U>*>(run_data), plan.starts, x,
			plan.summaries, n, max_run_length, run_capacity, status);
		plan.total = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(plan.summaries) + num_super_tiles * 32);
		return launched("kernel launch (run encoding, one pass)");
	}
	if (fast) { enc::run_tile_summaries_fast<U><<<blocks, enc::run_tile_threads, 0, s>>>(plan.summaries, x, n); }
	else { enc::run_tile_summaries<U><<<blocks, enc::run_tile_threads, 0, s>>>(plan.summaries, x, n, max_run_length); }
	giddy_status_t result = launched("kernel launch (run tile summaries)");
	if (result != GIDDY_SUCCESS) { return result; }
	const unsigned chunks = (unsigned) plan.num_chunks;
	enc::run_chunk_last_heads<<<chunks, enc::placement_chunk_tiles, 0, s>>>(plan.chunk_open, plan.summaries, plan.num_tiles);
	enc::run_scan_chunks<true><<<1, 1024, 0, s>>>(plan.chunk_open, plan.num_chunks, nullptr);
	enc::run_tile_counts<<<chunks, enc::placement_chunk_tiles, 0, s>>>(plan.placement, plan.open_run_start, plan.chunk_totals, plan.chunk_open, plan.summaries,
		plan.num_tiles, n, max_run_length);
	enc::run_scan_chunks<false><<<1, 1024, 0, s>>>(plan.chunk_totals, plan.num_chunks, plan.chunk_totals + plan.num_chunks);
	enc::run_tile_places<<<chunks, enc::placement_chunk_tiles, 0, s>>>(plan.placement, plan.chunk_totals, plan.chunk_totals + plan.num_chunks, plan.num_tiles);
	if ((result = launched("kernel launch (run tile placement)")) != GIDDY_SUCCESS) { return result; }
	if (fast) {
		enc::run_emit_fast<U, uint64_t><<<blocks, enc::run_tile_threads, 0, s>>>(static_cast<uint_t<U>*>(run_data), plan.starts, x,
			plan.placement, plan.open_run_start, plan.summaries, n, max_run_length, run_capacity, status);
	}
	else {
		enc::run_emit<U, uint64_t><<<blocks, enc::run_tile_threads, 0, s>>>(static_cast<uint_t<U>*>(run_data), plan.starts, x,
			plan.placement, plan.open_run_start, n, max_run_length, run_capacity, status);
	}
	return launched("kernel launch (run emission)");
}

using rle_launcher_t = giddy_status_t (*)(void*, const giddy_launch_config_t*, void*, void*, void*, void*, void*, const void*, uint64_t, uint64_t, uint64_t, uint64_t, unsigned*);
struct rle_entry_t { unsigned UI, U, RL, RI; const void* kernel; rle_launcher_t launch; };

template <unsigned UI, unsigned U, unsigned RL, unsigned RI>
giddy_status_t rle_launch(void* stream, const giddy_launch_config_t* cfg, void* run_data, void* run_lengths, void* anchors, void* offsets, void* num_runs_out,
	const void* in, uint64_t n, uint64_t L, uint64_t max_run_length, uint64_t run_capacity, unsigned* status)
{
	cudaStream_t s = as_stream(stream);
	const uint64_t type_limit = (RL >= 8) ? ~0ull : ((1ull << (8 * RL)) - 1);
	if (max_run_length == 0 || max_run_length > type_limit) { max_run_length = type_limit; }
	if (max_run_length >= n) { max_run_length = 0; } // no run can be that long: no split points
	run_plan_t plan;
	workspace_t workspace(s);
	cudaError_t error = workspace.take(run_workspace_bytes(n, run_capacity, plan));
	if (error != cudaSuccess) { return fail_cuda(error, "cudaMallocFromPoolAsync (run encoder workspace)"); }
	run_workspace_layout(workspace.memory, run_capacity, plan);
	giddy_status_t result = run_common<U>(s, cfg, plan, run_data, in, n, max_run_length, run_capacity, status);
	if (result != GIDDY_SUCCESS) { return result; }
	enc::run_lengths_from_starts<RL, uint64_t><<<helper_blocks(run_capacity < n ? run_capacity : n, 256), 256, 0, s>>>(
		static_cast<uint_t<RL>*>(run_lengths), plan.starts, plan.total, n, run_capacity);
	if ((result = launched("kernel launch (run lengths)")) != GIDDY_SUCCESS) { return result; }
	enc::run_anchors<RI, RL, uint64_t, true><<<helper_blocks((n + L - 1) / L, 256), 256, 0, s>>>(static_cast<uint_t<RI>*>(anchors), static_cast<uint_t<RL>*>(offsets),
		num_runs_out, plan.starts, plan.total, n, L, run_capacity, status);
	return launched("kernel launch (run anchors)");
}
#define RLE_ENTRY(UI, U, RL, RI) { UI, U, RL, RI, reinterpret_cast<const void*>(&enc::run_tile_summaries<U>), &rle_launch<UI, U, RL, RI> }
const rle_entry_t rle_table[] = {
	RLE_ENTRY(4, 1, 1, 4), RLE_ENTRY(4, 2, 1, 4), RLE_ENTRY(4, 4, 1, 4), RLE_ENTRY(4, 8, 1, 4),
	RLE_ENTRY(4, 1, 2, 4), RLE_ENTRY(4, 2, 2, 4), RLE_ENTRY(4, 4, 2, 4), RLE_ENTRY(4, 8, 2, 4),
	RLE_ENTRY(4, 1, 4, 4), RLE_ENTRY(4, 2, 4, 4), RLE_ENTRY(4, 4, 4, 4), RLE_ENTRY(4, 8, 4, 4),
	RLE_ENTRY(8, 4, 1, 8), RLE_ENTRY(8, 4, 2, 8), RLE_ENTRY(8, 4, 4, 8), RLE_ENTRY(8, 4, 8, 8), RLE_ENTRY(8, 1, 8, 8), RLE_ENTRY(8, 8, 8, 8),
};
const rle_entry_t* find_rle(unsigned UI, unsigned U, unsigned RL, unsigned RI)
{
	for (const auto& e : rle_table) { if (e.UI == UI && e.U == U && e.RL == RL && e.RI == RI) { return &e; } }
	return nullptr;
}

using rpe_launcher_t = giddy_status_t (*)(void*, const giddy_launch_config_t*, void*, void*, void*, void*, const void*, uint64_t, uint64_t, bool, uint64_t, uint64_t, unsigned*);
struct rpe_entry_t { unsigned UI, U, PO, RI; const void* kernel; rpe_launcher_t launch; };

template <unsigned UI, unsigned U, unsigned PO, unsigned RI>
giddy_status_t rpe_launch(void* stream, const giddy_launch_config_t* cfg, void* run_data, void* positions, void* anchors, void* num_runs_out,
	const void* in, uint64_t n, uint64_t L, bool relative, uint64_t max_run_length, uint64_t run_capacity, unsigned* status)
{
	cudaStream_t s = as_stream(stream);
	if (max_run_length >= n) { max_run_length = 0; }
	run_plan_t plan;
	workspace_t workspace(s);
	cudaError_t error = workspace.take(run_workspace_bytes(n, run_capacity, plan));
	if (error != cudaSuccess) { return fail_cuda(error, "cudaMallocFromPoolAsync (run encoder workspace)"); }
	run_workspace_layout(workspace.memory, run_capacity, plan);
	giddy_status_t result = run_common<U>(s, cfg, plan, run_data, in, n, max_run_length, run_capacity, status);
	if (result != GIDDY_SUCCESS) { return result; }
	enc::run_positions_from_starts<PO, uint64_t><<<helper_blocks(run_capacity < n ? run_capacity : n, 256), 256, 0, s>>>(
		static_cast<uint_t<PO>*>(positions), plan.starts, plan.total, L, relative, run_capacity, status);
	if ((result = launched("kernel launch (run start positions)")) != GIDDY_SUCCESS) { return result; }
	enc::run_anchors<RI, 1, uint64_t, false><<<helper_blocks((n + L - 1) / L, 256), 256, 0, s>>>(static_cast<uint_t<RI>*>(anchors), nullptr,
		num_runs_out, plan.starts, plan.total, n, L, run_capacity, status);
	return launched("kernel launch (run anchors)");
}
#define RPE_ENTRY(UI, U, PO, RI) { UI, U, PO, RI, reinterpret_cast<const void*>(&enc::run_tile_summaries<U>), &rpe_launch<UI, U, PO, RI> }
const rpe_entry_t rpe_table[] = {
	RPE_ENTRY(4, 1, 2, 4), RPE_ENTRY(4, 2, 2, 4), RPE_ENTRY(4, 4, 2, 4), RPE_ENTRY(4, 8, 2, 4), RPE_ENTRY(8, 4, 2, 8),
	RPE_ENTRY(4, 1, 4, 4), RPE_ENTRY(4, 2, 4, 4), RPE_ENTRY(4, 4, 4, 4), RPE_ENTRY(4, 8, 4, 4), RPE_ENTRY(8, 4, 8, 8),
};
const rpe_entry_t* find_rpe(unsigned UI, unsigned U, unsigned PO, unsigned RI)
{
	for (const auto& e : rpe_table) { if (e.UI == UI && e.U == U && e.PO == PO && e.RI == RI) { return &e; } }
	return nullptr;
}

// ---- DZB-variable ---------------------------------------------------------------------------------------

using dzbv_launcher_t = giddy_status_t (*)(void*, const giddy_launch_config_t*, void*, void*, void*, void*, const void*, uint64_t, uint64_t, unsigned, unsigned, uint64_t, unsigned*);
struct dzbv_entry_t { unsigned I, U, CS; const void* kernel; dzbv_launcher_t launch; };

template <unsigned I, unsigned U, unsigned CS>
giddy_status_t dzbv_launch(void* stream, const giddy_launch_config_t* cfg, void* data, void* packed, void* anchors, void* num_bytes_out,
	const void* in, uint64_t n, uint64_t L, unsigned min_size, unsigned bits, uint64_t data_capacity, unsigned* status)
{
	cudaStream_t s = as_stream(stream);
	const uint64_t num_segments = (n + L - 1) / L;
	workspace_t workspace(s);
	const uint64_t num_chunks = (num_segments + enc::placement_chunk_tiles - 1) / enc::placement_chunk_tiles;
	cudaError_t error = workspace.take((num_segments + 1) * 8 + (num_chunks + 1) * 8 + 64);
	if (error != cudaSuccess) { return fail_cuda(error, "cudaMallocFromPoolAsync (variable-width encoder workspace)"); }
	auto* segment_bytes = static_cast<uint64_t*>(workspace.memory);
	const unsigned block = cfg->block_dimensions[0] ? cfg->block_dimensions[0] : 256;
	// (warps take segments grid-stride: more CTAs than fill the device a few times over only cost their launches)
	const unsigned enough = helper_blocks(num_segments, block / 32);
	const unsigned grid = (cfg->grid_dimensions[0] && cfg->grid_dimensions[0] < enough) ? cfg->grid_dimensions[0] : enough;
	auto x = static_cast<const uint_t<U>*>(in);
	// the size containers come out of the first pass where they straddle neither a lane's 16-byte vector nor segments and are
	// made of a power-of-two number of lanes' fields
	constexpr unsigned E = 16 / U;
	const unsigned per_container = (8 * CS) / bits;
	const unsigned lanes_per_container = per_container / E; // (a power of two for 1, 2, 4 and 8-bit fields)
	const bool pack_in_first_pass = reinterpret_cast<uintptr_t>(in) % 16 == 0 && L % E == 0 && per_container % E == 0 && L % per_container == 0 &&
		lanes_per_container >= 1 && (lanes_per_container & (lanes_per_container - 1)) == 0;
	enc::variable_width_sizes<U, CS><<<grid, block, 0, s>>>(segment_bytes, pack_in_first_pass ? static_cast<uint_t<CS>*>(packed) : nullptr,
		x, n, L, min_size, bits, status);
	giddy_status_t result = launched("kernel launch (variable-width segment sizes)");
	if (result != GIDDY_SUCCESS) { return result; }
	uint64_t* chunk_totals = segment_bytes + num_segments + 1;
	enc::chunk_exclusive_sums<<<(unsigned) num_chunks, enc::placement_chunk_tiles, 0, s>>>(segment_bytes, chunk_totals, num_segments);
	enc::run_scan_chunks<false><<<1, 1024, 0, s>>>(chunk_totals, num_chunks, chunk_totals + num_chunks);
	enc::add_chunk_bases<<<(unsigned) num_chunks, enc::placement_chunk_tiles, 0, s>>>(segment_bytes, chunk_totals, chunk_totals + num_chunks, num_segments);
	if ((result = launched("kernel launch (segment byte offsets)")) != GIDDY_SUCCESS) { return result; }
	const unsigned windows = (block / 32) * enc::emit_window_bytes;
	error = ensure_dynamic_shared_memory(reinterpret_cast<const void*>(&enc::variable_width_emit<I, U>), windows);
	if (error != cudaSuccess) { return fail_cuda(error, "cudaFuncSetAttribute(MaxDynamicSharedMemorySize)"); }
	enc::variable_width_emit<I, U><<<grid, block, windows, s>>>(static_cast<uint8_t*>(data), static_cast<uint_t<I>*>(anchors), static_cast<uint64_t*>(num_bytes_out),
		x, segment_bytes, n, L, min_size, data_capacity, status);
	if ((result = launched("kernel launch (variable-width emission)")) != GIDDY_SUCCESS) { return result; }
	if (pack_in_first_pass) { return GIDDY_SUCCESS; }
	const uint64_t containers = (n + per_container - 1) / per_container;
	enc::variable_width_pack_sizes<U, CS><<<helper_blocks(containers, 256), 256, 0, s>>>(static_cast<uint_t<CS>*>(packed), x, n, min_size, bits);
	return launched("kernel launch (size containers)");
}
#define DZBV_ENTRY(I, U, CS) { I, U, CS, reinterpret_cast<const void*>(&enc::variable_width_emit<I, U>), &dzbv_launch<I, U, CS> }
const dzbv_entry_t dzbv_table[] = {
	DZBV_ENTRY(4, 2, 4), DZBV_ENTRY(4, 4, 4), DZBV_ENTRY(4, 8, 4), DZBV_ENTRY(8, 2, 4), DZBV_ENTRY(8, 4, 4), DZBV_ENTRY(8, 8, 4),
	DZBV_ENTRY(4, 2, 8), DZBV_ENTRY(4, 4, 8), DZBV_ENTRY(4, 8, 8),
};
const dzbv_entry_t* find_dzbv(unsigned I, unsigned U, unsigned CS)
{
	for (const auto& e : dzbv_table) { if (e.I == I && e.U == U && e.CS == CS) { return &e; } }
	return nullptr;
}

// ---- incidence bitmaps -----------------------------------------------------------------------------------

using bitmaps_launcher_t = giddy_status_t (*)(void*, const giddy_launch_config_t*, void*, const void*, const void*, uint64_t, uint64_t, unsigned, unsigned*);
struct bitmaps_entry_t { unsigned U; bool use_dictionary; const void* kernel; bitmaps_launcher_t launch; };
template <unsigned U, bool D>
giddy_status_t bitmaps_launch(void* stream, const giddy_launch_config_t* cfg, void* bitmaps, const void* in, const void* entries, uint64_t n, uint64_t pitch,
	unsigned num_bitmaps, unsigned* status)
{
	return launch(enc::incidence_bitmaps_compress<U, D>, stream, cfg, static_cast<uint32_t*>(bitmaps), static_cast<const uint_t<U>*>(in),
		static_cast<const uint_t<U>*>(entries), n, pitch, num_bitmaps, status);
}
#define BITMAPS_ENTRY(U, D) { U, D, reinterpret_cast<const void*>(&enc::incidence_bitmaps_compress<U, D>), &bitmaps_launch<U, D> }
const bitmaps_entry_t bitmaps_table[] = { BITMAPS_ENTRY(1, true), BITMAPS_ENTRY(2, true), BITMAPS_ENTRY(4, true), BITMAPS_ENTRY(8, true), BITMAPS_ENTRY(1, false) };
const bitmaps_entry_t* find_bitmaps(unsigned U, bool use_dictionary)
{
	for (const auto& e : bitmaps_table) { if (e.U == U && e.use_dictionary == use_dictionary) { return &e; } }
	return nullptr;
}

giddy_status_t check_warps(const giddy_launch_config_t* config)
{
	if (config->block_dimensions[0] % 32 != 0 || config->block_dimensions[0] == 0) { return fail(GIDDY_INVALID_ARGUMENT, "blocks must be made of whole warps"); }
	return GIDDY_SUCCESS;
}

} // namespace

extern "C" {

const void* giddy_b200_delta_compress_kernel(unsigned I, unsigned U, unsigned C) { auto e = find_uc(I, U, C); return e ? e->delta_kernel : nullptr; }

giddy_status_t giddy_b200_delta_compress_launch(void* stream, const giddy_launch_config_t* config, unsigned I, unsigned U, unsigned C,
	void* compressed, void* anchor_values, const void* uncompressed, uint64_t length, uint64_t segment_length, uint64_t base_value, unsigned* status_flags)
{
	auto e = find_uc(I, U, C);
	giddy_status_t status;
	if (!check_launch(config, e ? e->delta_kernel : nullptr, status)) { return status; }
	if (segment_length == 0) { return fail(GIDDY_INVALID_ARGUMENT, "The segment length must be a positive number"); }
	if (length == 0) { return GIDDY_SUCCESS; }
	return e->delta(stream, config, compressed, anchor_values, uncompressed, length, segment_length, base_value, status_flags);
}

const void* giddy_b200_discard_zero_bytes_fixed_compress_kernel(unsigned I, unsigned U, unsigned C) { auto e = find_uc(I, U, C); return e ? e->dzbf_kernel : nullptr; }

giddy_status_t giddy_b200_discard_zero_bytes_fixed_compress_launch(void* stream, const giddy_launch_config_t* config, unsigned I, unsigned U, unsigned C,
	void* compressed, const void* uncompressed, uint64_t length, unsigned* status_flags)
{
	auto e = find_uc(I, U, C);
	giddy_status_t status;
	if (!check_launch(config, e ? e->dzbf_kernel : nullptr, status)) { return status; }
	if (length == 0) { return GIDDY_SUCCESS; }
	return e->dzbf(stream, config, compressed, nullptr, uncompressed, length, 0, 0, status_flags);
}

const void* giddy_b200_dictionary_compress_kernel(unsigned I, unsigned U, unsigned DI) { auto e = find_dict(I, U, DI); return e ? e->kernel : nullptr; }

giddy_status_t giddy_b200_dictionary_compress_launch(void* stream, const giddy_launch_config_t* config, unsigned I, unsigned U, unsigned DI,
	void* compressed, const void* uncompressed, const void* dictionary_entries, uint64_t num_dictionary_entries, uint64_t length, unsigned* status_flags)
{
	auto e = find_dict(I, U, DI);
	giddy_status_t status;
	if (!check_launch(config, e ? e->kernel : nullptr, status)) { return status; }
	if (DI < 8 && num_dictionary_entries > (1ull << (8 * DI))) { return fail(GIDDY_INVALID_ARGUMENT, "more dictionary entries than the index type can address"); }
	if (length == 0) { return GIDDY_SUCCESS; }
	return e->launch(stream, config, compressed, uncompressed, dictionary_entries, num_dictionary_entries, length, status_flags);
}

const void* giddy_b200_frame_of_reference_compress_kernel(unsigned I, unsigned U, unsigned C, giddy_model_t model)
{
	auto e = find_for(I, U, C, (int) model);
	return e ? e->kernel : nullptr;
}

giddy_status_t giddy_b200_frame_of_reference_compress_launch(void* stream, const giddy_launch_config_t* config, unsigned I, unsigned U, unsigned C, giddy_model_t model,
	void* compressed, void* segment_model_coefficients, const void* uncompressed, uint64_t length, uint64_t segment_length, unsigned* status_flags)
{
	auto e = find_for(I, U, C, (int) model);
	giddy_status_t status;
	if (!check_launch(config, e ? e->kernel : nullptr, status)) { return status; }
	if ((status = check_warps(config)) != GIDDY_SUCCESS) { return status; }
	if (segment_length == 0) { return fail(GIDDY_INVALID_ARGUMENT, "The segment length must be a positive number"); }
	if (length == 0) { return GIDDY_SUCCESS; }
	return e->launch(stream, config, compressed, segment_model_coefficients, uncompressed, length, segment_length, status_flags);
}

const void* giddy_b200_run_length_encoding_compress_kernel(unsigned UI, unsigned U, unsigned RL, unsigned RI) { auto e = find_rle(UI, U, RL, RI); return e ? e->kernel : nullptr; }

giddy_status_t giddy_b200_run_length_encoding_compress_launch(void* stream, const giddy_launch_config_t* config, unsigned UI, unsigned U, unsigned RL, unsigned RI,
	void* run_data, void* run_lengths, void* position_anchors, void* intra_run_anchor_offsets, void* num_element_runs,
	const void* uncompressed, uint64_t length, uint64_t segment_length, uint64_t max_run_length, uint64_t run_capacity, unsigned* status_flags)
{
	auto e = find_rle(UI, U, RL, RI);
	giddy_status_t status;
	if (!check_launch(config, e ? e->kernel : nullptr, status)) { return status; }
	if (segment_length == 0) { return fail(GIDDY_INVALID_ARGUMENT, "The segment length must be a positive number"); }
	if (length == 0) { return GIDDY_SUCCESS; }
	return e->launch(stream, config, run_data, run_lengths, position_anchors, intra_run_anchor_offsets, num_element_runs, uncompressed, length, segment_length,
		max_run_length, run_capacity, status_flags);
}

const void* giddy_b200_run_position_encoding_compress_kernel(unsigned UI, unsigned U, unsigned PO, unsigned RI) { auto e = find_rpe(UI, U, PO, RI); return e ? e->kernel : nullptr; }

giddy_status_t giddy_b200_run_position_encoding_compress_launch(void* stream, const giddy_launch_config_t* config, unsigned UI, unsigned U, unsigned PO,
	int positions_are_relative, unsigned RI, void* run_data, void* run_start_positions, void* position_anchors, void* num_element_runs,
	const void* uncompressed, uint64_t length, uint64_t segment_length, uint64_t max_run_length, uint64_t run_capacity, unsigned* status_flags)
{
	auto e = find_rpe(UI, U, PO, RI);
	giddy_status_t status;
	if (!check_launch(config, e ? e->kernel : nullptr, status)) { return status; }
	if (segment_length == 0) { return fail(GIDDY_INVALID_ARGUMENT, "The segment length must be a positive number"); }
	if (length == 0) { return GIDDY_SUCCESS; }
	return e->launch(stream, config, run_data, run_start_positions, position_anchors, num_element_runs, uncompressed, length, segment_length,
		positions_are_relative != 0, max_run_length, run_capacity, status_flags);
}

const void* giddy_b200_discard_zero_bytes_variable_compress_kernel(unsigned I, unsigned U, unsigned CS) { auto e = find_dzbv(I, U, CS); return e ? e->kernel : nullptr; }

giddy_status_t giddy_b200_discard_zero_bytes_variable_compress_launch(void* stream, const giddy_launch_config_t* config, unsigned I, unsigned U, unsigned CS,
	void* compressed_data, void* packed_element_sizes, void* position_anchors, void* num_bytes,
	const void* uncompressed, uint64_t length, uint64_t segment_length, unsigned min_represented_element_size, unsigned bits_per_element_size,
	uint64_t data_capacity, unsigned* status_flags)
{
	auto e = find_dzbv(I, U, CS);
	giddy_status_t status;
	if (!check_launch(config, e ? e->kernel : nullptr, status)) { return status; }
	if ((status = check_warps(config)) != GIDDY_SUCCESS) { return status; }
	if (segment_length == 0) { return fail(GIDDY_INVALID_ARGUMENT, "The segment length must be a positive number"); }
	if (bits_per_element_size == 0 || bits_per_element_size > 8 * CS) { return fail(GIDDY_INVALID_ARGUMENT, "bits_per_element_size must be between 1 and the container's width"); }
	if (min_represented_element_size > U) { return fail(GIDDY_INVALID_ARGUMENT, "min_represented_element_size exceeds the element size"); }
	if (length == 0) { return GIDDY_SUCCESS; }
	return e->launch(stream, config, compressed_data, packed_element_sizes, position_anchors, num_bytes, uncompressed, length, segment_length,
		min_represented_element_size, bits_per_element_size, data_capacity, status_flags);
}

const void* giddy_b200_incidence_bitmaps_compress_kernel(unsigned U, int use_dictionary) { auto e = find_bitmaps(U, use_dictionary != 0); return e ? e->kernel : nullptr; }

giddy_status_t giddy_b200_incidence_bitmaps_compress_launch(void* stream, const giddy_launch_config_t* config, unsigned U, unsigned bitmap_alignment_in_ints, int use_dictionary,
	void* bitmaps, const void* uncompressed, const void* dictionary_entries, uint64_t length, unsigned num_bitmaps, unsigned* status_flags)
{
	auto e = find_bitmaps(U, use_dictionary != 0);
	giddy_status_t status;
	if (!check_launch(config, e ? e->kernel : nullptr, status)) { return status; }
	if ((status = check_warps(config)) != GIDDY_SUCCESS) { return status; }
	if (bitmap_alignment_in_ints == 0) { return fail(GIDDY_INVALID_ARGUMENT, "the bitmap alignment must be positive"); }
	if (length == 0) { return GIDDY_SUCCESS; }
	const uint64_t bitmap_length = (length + 31) / 32;
	const uint64_t pitch = (bitmap_length + bitmap_alignment_in_ints - 1) / bitmap_alignment_in_ints * bitmap_alignment_in_ints;
	return e->launch(stream, config, bitmaps, uncompressed, dictionary_entries, length, pitch, num_bitmaps, status_flags);
}

} // extern "C"
