// Delta decoder: a segment-anchored inclusive prefix sum.
//
// One warp decodes one anchored segment.  A "row" is 32 lanes x E elements (E = 16 / sizeof
// (Uncompressed), so a lane's output is one 16-byte vector and the warp's is 512 contiguous
// bytes).  `rows_per_step` rows are loaded before any is scanned, so each lane keeps that
// many independent loads in flight and the row scans -- a serial scan of the lane's E deltas
// in registers plus one warp-shuffle scan of the lane totals -- overlap.  The only serial
// dependency is the running carry handed from row to row, one add per row.  Per-segment
// anchors make segments independent, so no inter-CTA communication is needed: the reference's
// block-wide scan with two barriers per 1024 elements (src/kernels/reduction/scan.cuh:345-361)
// becomes barrier-free.
#pragma once
#ifndef GIDDY_B200_KERNELS_DELTA_CUH_
#define GIDDY_B200_KERNELS_DELTA_CUH_

#include "common.cuh"

namespace giddy {
namespace kernels {

using namespace giddy::device;

namespace delta_detail {

constexpr unsigned rows_per_step = GIDDY_DELTA_ROWS_PER_STEP;

// Decode one segment with E elements per lane per row.  E > 1 requires the segment's first input
// element to be E*C-aligned and its first output element to be 16-byte aligned.
template <unsigned U, unsigned C, unsigned E>
__device__ __forceinline__ void decode_segment(
	uint_t<U>*        __restrict__  out,
	const uint_t<C>*  __restrict__  in,
	uint64_t                        length,
	uint_t<U>                       carry)
{
	const unsigned lane = lane_index();
	constexpr unsigned row_elements = warp_size * E;

	for (uint64_t step_start = 0; step_start < length; step_start += rows_per_step * row_elements) {
		byte_pack<E * C> deltas[rows_per_step];
		#pragma unroll
		for (unsigned r = 0; r < rows_per_step; r++) {
			const uint64_t p = step_start + r * row_elements + lane * E;
			if (p + E <= length) { deltas[r] = load_pack<E * C>(in + p); }
			else {
				#pragma unroll
				for (unsigned j = 0; j < E; j++) { deltas[r].template set<C>(j, (p + j < length) ? in[p + j] : (uint_t<C>) 0); }
			}
		}
		#pragma unroll
		for (unsigned r = 0; r < rows_per_step; r++) {
			const uint64_t p = step_start + r * row_elements + lane * E;
			if (step_start + r * row_elements >= length) { break; } // warp-uniform
			uint_t<U> local[E];
			uint_t<U> sum = 0;
			#pragma unroll
			for (unsigned j = 0; j < E; j++) {
				sum = (uint_t<U>)(sum + (uint_t<U>) deltas[r].template get<C>(j));
				local[j] = sum;
			}
			const uint_t<U> inclusive = warp_inclusive_sum(sum);
			const uint_t<U> before = (uint_t<U>)(carry + (uint_t<U>)(inclusive - sum));
			carry = (uint_t<U>)(carry + shuffle_from(inclusive, warp_size - 1));
			if (E > 1 && p + E <= length) {
				byte_pack<E * U> vec; // 16 bytes, or 32: one 256-bit store
				#pragma unroll
				for (unsigned j = 0; j < E; j++) { vec.template set<U>(j, (uint_t<U>)(before + local[j])); }
				if constexpr (E * U == 32) { store_streaming_32(out + p, vec.u32); }
				else { store_streaming_16(out + p, make_uint4(vec.u32[0], vec.u32[1], vec.u32[2], vec.u32[3])); }
			}
			else {
				#pragma unroll
				for (unsigned j = 0; j < E; j++) { if (p + j < length) { out[p + j] = (uint_t<U>)(before + local[j]); } }
			}
		}
	}
}

} // namespace delta_detail

// -------------------------------------------------------------------------------------------------
// Delta.  Replaces cuda::kernels::decompression::delta::decompress,
// reference src/kernels/decompression/delta.cuh:66-108 (same parameter list):
//   decompressed[i] = anchor_values[s] + sum_{j = s*L .. i} zero_extend(compressed_input[j])
// -------------------------------------------------------------------------------------------------
template <unsigned IndexSize, unsigned U, unsigned C>
__global__ void delta_decompress(
	uint_t<U>*        __restrict__  decompressed,
	const uint_t<C>*  __restrict__  compressed_input,
	const uint_t<U>*  __restrict__  anchor_values,
	size_type<IndexSize>            length,
	size_type<IndexSize>            segment_length)
{
	constexpr unsigned E = 16 / U;
	const uint64_t num_segments = ((uint64_t) length + segment_length - 1) / segment_length;
	const bool vectorizable = is_aligned_to(decompressed, 16) && is_aligned_to(compressed_input, E * C) &&
		(segment_length % E == 0);
	// (32 output bytes per lane per row where alignment allows: one 256-bit store, half the load and shuffle instructions per byte)
	const bool vectorizable_32 = vectorizable && is_aligned_to(decompressed, 32) && is_aligned_to(compressed_input, 2 * E * C) &&
		(segment_length % (2 * E) == 0);

	for (uint64_t segment = global_warp_index(); segment < num_segments; segment += num_warps_in_grid()) {
		const uint64_t start = segment * segment_length;
		const uint64_t this_length = minimum<uint64_t>(segment_length, length - start);
		const uint_t<U> anchor = __ldg(anchor_values + segment);
		if (vectorizable_32)   { delta_detail::decode_segment<U, C, 2 * E>(decompressed + start, compressed_input + start, this_length, anchor); }
		else if (vectorizable) { delta_detail::decode_segment<U, C, E>(decompressed + start, compressed_input + start, this_length, anchor); }
		else                   { delta_detail::decode_segment<U, C, 1>(decompressed + start, compressed_input + start, this_length, anchor); }
	}
}

// -------------------------------------------------------------------------------------------------
// Delta with LONG segments (few anchors, down to a single one: a plain prefix sum over the column).
// A warp per segment cannot fill the GPU then, so segments are cut into CTA-sized tiles which are scanned
// in a single pass with decoupled look-back: every tile publishes its aggregate as soon as it has it, and
// obtains its exclusive prefix by summing its predecessors' aggregates back to the nearest tile whose
// inclusive prefix is already known.  The first tile of every segment knows its prefix at once (the
// segment's anchor), so a look-back never leaves its segment.
//
// What keeps the look-back off the critical path:
//   * a descriptor is self-validating -- each 64-bit word carries the status in its upper half and 32 bits
//     of the value in its lower half -- so publishing is one relaxed store and polling one relaxed load:
//     no fences, and status and value arrive in the same round trip (a 64-bit value takes two such words
//     in one 16-byte access; a torn pair shows two different statuses and is polled again);
//   * a warp inspects look_back_window x 32 predecessors per round trip, so that the distance to the
//     inclusive-prefix wavefront (tiles started per L2 round trip, ~100 at HBM speed) is covered in one round;
//   * tiles are large (look_back_rows 16-byte vectors per thread);
//   * (cooperative launches) the deltas of the next `look_back_stages` tiles are staged in shared memory by TMA
//     bulk copies, and a tile's aggregate is published a whole scan-and-store phase before its own look-back
//     begins -- by then the neighbouring CTAs' aggregates are normally there.
//
// Persistent CTAs; a tile's predecessors are always running or done, so the spin-waits cannot deadlock: either
// the grid is launched cooperatively (all CTAs resident) and CTA b takes tiles b, b + grid, ..., or tiles are
// handed out in order through a ticket counter (any grid size; drawing tickets ahead of time is not an option: a
// drawn but unstarted tile stalls all its successors, and the stalls chain up -- measured 20x slower).
//
// workspace: 32 zeroed bytes (the ticket counter) followed by one zeroed 16-byte descriptor per tile;
// the launcher takes it from the stream-ordered allocator (cudaMallocAsync) for the duration of the launch.
// -------------------------------------------------------------------------------------------------
namespace delta_detail {

#ifdef GIDDY_DELTA_TRACE  // development aid (benchmarks/micro/lookback_probe.cu): per-tile phase timestamps
__device__ unsigned long long trace[1 << 20];
__device__ __forceinline__ unsigned long long trace_now() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define GIDDY_DELTA_TRACE_POINT(tile, k) do { if (threadIdx.x == 0 && (tile) < (1u << 17)) { trace[(tile) * 8 + (k)] = trace_now(); } } while (0)
#define GIDDY_DELTA_TRACE_VALUE(tile, k, v) do { if (threadIdx.x == 0 && (tile) < (1u << 17)) { trace[(tile) * 8 + (k)] = (v); } } while (0)
#else
#define GIDDY_DELTA_TRACE_POINT(tile, k) do { } while (0)
#define GIDDY_DELTA_TRACE_VALUE(tile, k, v) do { } while (0)
#endif

constexpr unsigned look_back_rows   = GIDDY_DELTA_LOOK_BACK_ROWS;
constexpr unsigned look_back_window = GIDDY_DELTA_LOOK_BACK_WINDOW; // descriptors per lane per round
constexpr unsigned look_back_stages = GIDDY_DELTA_LOOK_BACK_STAGES; // tiles of deltas staged in shared memory

struct alignas(16) tile_descriptor_t { unsigned long long word[2]; };
enum : uint32_t { tile_unknown = 0, tile_has_aggregate = 1, tile_has_inclusive = 2 };

template <unsigned U>
__device__ __forceinline__ void publish(tile_descriptor_t* d, uint32_t status, uint_t<U> value)
{
	const unsigned long long tag = (unsigned long long) status << 32;
	if (U <= 4) {
		asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(d), "l"(tag | (unsigned long long)(uint32_t) value) : "memory");
	}
	else {
		const unsigned long long v = (unsigned long long) value;
		asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" :: "l"(d), "l"(tag | (v & 0xffffffffull)), "l"(tag | (v >> 32)) : "memory");
	}
}

// the descriptor's status, and its value if the status is not tile_unknown
template <unsigned U>
__device__ __forceinline__ uint32_t peek(const tile_descriptor_t* d, uint_t<U>& value)
{
	if (U <= 4) {
		unsigned long long w;
		asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w) : "l"(d) : "memory");
		value = (uint_t<U>) w;
		return (uint32_t)(w >> 32);
	}
	else {
		unsigned long long w0, w1;
		asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(d) : "memory");
		value = (uint_t<U>)((w0 & 0xffffffffull) | (w1 << 32));
		const uint32_t s0 = (uint32_t)(w0 >> 32), s1 = (uint32_t)(w1 >> 32);
		return s0 == s1 ? s0 : (uint32_t) tile_unknown;
	}
}

template <typename T>
__device__ __forceinline__ T warp_sum(T x)
{
	if constexpr (sizeof(T) <= 4) { return (T) __reduce_add_sync(full_mask, (unsigned) x); }
	else {
		#pragma unroll
		for (unsigned delta = warp_size / 2; delta > 0; delta >>= 1) { x = (T)(x + (T) __shfl_xor_sync(full_mask, (unsigned long long) x, delta)); }
		return x;
	}
}

// The exclusive prefix of tile `tile`, the `tile_in_segment`-th (>= 1) of its segment; called by one whole warp.
// A round inspects the look_back_window x 32 nearest predecessors not yet summed, all lanes polling at once; it
// ends as soon as the descriptors up to the nearest one with an inclusive prefix have been seen (farther ones are
// never waited for), or when all of them hold just aggregates.
template <unsigned U>
__device__ __forceinline__ uint_t<U> look_back(const tile_descriptor_t* descriptors, uint64_t tile, uint64_t tile_in_segment)
{
	unsigned polls = 0, rounds = 0; (void) polls; (void) rounds;
	const unsigned lane = lane_index();
	constexpr unsigned round_tiles = look_back_window * warp_size;
	uint_t<U> exclusive = 0;
	uint64_t looked = 0; // predecessors already summed
	while (true) {
		uint_t<U> value[look_back_window];
		uint32_t status[look_back_window];
		bool valid[look_back_window];
		#pragma unroll
		for (unsigned q = 0; q < look_back_window; q++) {
			valid[q] = looked + q * warp_size + lane < tile_in_segment; // distance 0 = the immediate predecessor
			status[q] = tile_unknown; value[q] = 0;
		}
		unsigned nearest; // distance (within this round) of the nearest tile with an inclusive prefix; round_tiles: none
		rounds++;
		while (true) {
			polls++;
			#pragma unroll
			for (unsigned q = 0; q < look_back_window; q++) {
				if (valid[q] && status[q] == tile_unknown) {
					status[q] = peek<U>(descriptors + (tile - 1 - (looked + q * warp_size + lane)), value[q]);
				}
			}
			nearest = round_tiles;
			bool all_seen = true;
			#pragma unroll
			for (unsigned q = 0; q < look_back_window; q++) {
				const unsigned unseen = __ballot_sync(full_mask, valid[q] && status[q] == tile_unknown);
				const unsigned with_inclusive = __ballot_sync(full_mask, valid[q] && status[q] == tile_has_inclusive);
				if (all_seen && nearest == round_tiles) {
					const unsigned seen_prefix = unseen ? ((1u << (__ffs((int) unseen) - 1)) - 1u) : full_mask; // lanes before the first unseen one
					const unsigned usable = with_inclusive & seen_prefix;
					if (usable) { nearest = q * warp_size + (unsigned) __ffs((int) usable) - 1; }
					else if (unseen) { all_seen = false; }
				}
			}
			if (nearest != round_tiles || all_seen) { break; }
		}
		uint_t<U> mine = 0;
		#pragma unroll
		for (unsigned q = 0; q < look_back_window; q++) {
			if (valid[q] && q * warp_size + lane <= nearest) { mine = (uint_t<U>)(mine + value[q]); }
		}
		exclusive = (uint_t<U>)(exclusive + warp_sum(mine));
		if (nearest != round_tiles) { GIDDY_DELTA_TRACE_VALUE(tile, 6, polls); GIDDY_DELTA_TRACE_VALUE(tile, 7, rounds); return exclusive; }
		looked += round_tiles;
	}
}

// How a column is cut into tiles (tiles never straddle segments: a segment's last tile may be ragged).
struct tile_t { uint64_t segment, tile_in_segment, start, length; };

struct tiling_t {
	uint64_t length, segment_length, tile_elements, tiles_per_segment, num_tiles;

	__device__ __forceinline__ tiling_t(uint64_t length_, uint64_t segment_length_, uint64_t tile_elements_) :
		length(length_), segment_length(segment_length_), tile_elements(tile_elements_)
	{
		tiles_per_segment = (segment_length + tile_elements - 1) / tile_elements;
		const uint64_t num_segments = (length + segment_length - 1) / segment_length;
		const uint64_t last_segment_length = length - (num_segments - 1) * segment_length;
		num_tiles = (num_segments - 1) * tiles_per_segment + (last_segment_length + tile_elements - 1) / tile_elements;
	}

	__device__ __forceinline__ tile_t locate(uint64_t tile) const
	{
		tile_t g;
		g.segment = tile / tiles_per_segment;
		g.tile_in_segment = tile - g.segment * tiles_per_segment;
		const uint64_t segment_start = g.segment * segment_length;
		const uint64_t offset = g.tile_in_segment * tile_elements;
		g.start = segment_start + offset;
		g.length = minimum(tile_elements, minimum(segment_length, length - segment_start) - offset);
		return g;
	}
};

// A warp scans its Rows rows of a tile (a lane's piece of a row: E deltas, the warp's rows start `warp_offset` elements
// into the tile) onto `carry` -- the sum of everything before them -- and stores the results.
// (`piece(r)`: the lane's E deltas of row r, as a byte_pack<E * C>)
template <unsigned U, unsigned C, unsigned E, unsigned Rows, typename Piece>
__device__ __forceinline__ void scan_and_store(
	uint_t<U>* __restrict__ decompressed, Piece piece, const tile_t& g, uint64_t warp_offset, uint_t<U> carry)
{
	constexpr unsigned row_elements = warp_size * E;
	#pragma unroll
	for (unsigned r = 0; r < Rows; r++) {
		const uint64_t p = warp_offset + r * row_elements + lane_index() * E;
		const byte_pack<E * C> deltas = piece(r);
		uint_t<U> local[E];
		uint_t<U> sum = 0;
		#pragma unroll
		for (unsigned j = 0; j < E; j++) {
			sum = (uint_t<U>)(sum + (uint_t<U>) deltas.template get<C>(j));
			local[j] = sum;
		}
		const uint_t<U> inclusive = warp_inclusive_sum(sum);
		const uint_t<U> before = (uint_t<U>)(carry + (uint_t<U>)(inclusive - sum));
		carry = (uint_t<U>)(carry + shuffle_from(inclusive, warp_size - 1));
		if (E > 1 && p + E <= g.length) {
			byte_pack<E * U> vec; // 16 bytes, or a multiple of 32: 256-bit stores
			#pragma unroll
			for (unsigned j = 0; j < E; j++) { vec.template set<U>(j, (uint_t<U>)(before + local[j])); }
			if constexpr (E * U >= 32) {
				#pragma unroll
				for (unsigned b = 0; b < E * U / 32; b++) {
					store_streaming_32(reinterpret_cast<unsigned char*>(decompressed + g.start + p) + 32 * b, *reinterpret_cast<const uint32_t (*)[8]>(&vec.u32[8 * b]));
				}
			}
			else { store_streaming_16(decompressed + g.start + p, make_uint4(vec.u32[0], vec.u32[1], vec.u32[2], vec.u32[3])); }
		}
		else {
			#pragma unroll
			for (unsigned j = 0; j < E; j++) { if (p + j < g.length) { decompressed[g.start + p + j] = (uint_t<U>)(before + local[j]); } }
		}
	}
}

// One tile at a time, its deltas in registers: load, reduce, barrier, look back (warp 0), barrier, scan and store.  Tiles are
// drawn from the ticket counter, in order, so a tile's predecessors are running or done whatever the grid: the kernel for
// columns too little aligned for bulk copies, and the fall-back where a cooperative launch is not to be had.
template <unsigned U, unsigned C, unsigned E, unsigned Rows>
__device__ __forceinline__ void decode_tiles_one_at_a_time(
	uint_t<U>*        __restrict__  decompressed,
	const uint_t<C>*  __restrict__  compressed_input,
	const uint_t<U>*  __restrict__  anchor_values,
	uint64_t                        length,
	uint64_t                        segment_length,
	void*                           workspace)
{
	// double-buffered by tile parity, so that two barriers per tile are enough
	__shared__ uint_t<U> warp_totals[2][32];
	__shared__ uint_t<U> shared_tile_exclusive[2];
	__shared__ uint64_t  shared_ticket[2];

	const unsigned lane = lane_index(), warp = warp_index_in_block(), num_warps = warps_per_block();
	constexpr unsigned row_elements = warp_size * E;
	const tiling_t tiling(length, segment_length, (uint64_t) num_warps * Rows * row_elements);
	unsigned long long* const ticket_counter = static_cast<unsigned long long*>(workspace);
	tile_descriptor_t* const descriptors = reinterpret_cast<tile_descriptor_t*>(static_cast<unsigned char*>(workspace) + 32);
	const uint64_t warp_offset = (uint64_t) warp * Rows * row_elements;

	if (threadIdx.x == 0) { shared_ticket[0] = atomicAdd(ticket_counter, 1ull); }
	__syncthreads();
	for (unsigned parity = 0; ; parity ^= 1u) {
		const uint64_t tile = shared_ticket[parity];
		if (tile >= tiling.num_tiles) { return; }
		const tile_t g = tiling.locate(tile);
		byte_pack<E * C> deltas[Rows];
		#pragma unroll
		for (unsigned r = 0; r < Rows; r++) {
			const uint64_t p = warp_offset + r * row_elements + lane * E;
			if (p + E <= g.length) { deltas[r] = load_pack<E * C>(compressed_input + g.start + p); }
			else {
				#pragma unroll
				for (unsigned j = 0; j < E; j++) { deltas[r].template set<C>(j, (p + j < g.length) ? compressed_input[g.start + p + j] : (uint_t<C>) 0); }
			}
		}
		// the warp's total first (one reduction): the tile's aggregate is published as early as possible
		uint_t<U> lane_total = 0;
		#pragma unroll
		for (unsigned r = 0; r < Rows; r++) {
			#pragma unroll
			for (unsigned j = 0; j < E; j++) { lane_total = (uint_t<U>)(lane_total + (uint_t<U>) deltas[r].template get<C>(j)); }
		}
		const uint_t<U> warp_total = warp_sum(lane_total);
		if (lane == 0) { warp_totals[parity][warp] = warp_total; }
		__syncthreads();
		uint_t<U> preceding_warps = 0, tile_aggregate = 0;
		for (unsigned w = 0; w < num_warps; w++) {
			const uint_t<U> t = warp_totals[parity][w];
			if (w < warp) { preceding_warps = (uint_t<U>)(preceding_warps + t); }
			tile_aggregate = (uint_t<U>)(tile_aggregate + t);
		}
		if (warp == 0) { // settles the tile's exclusive prefix
			tile_descriptor_t* mine = descriptors + tile;
			uint_t<U> exclusive;
			if (g.tile_in_segment == 0) { exclusive = __ldg(anchor_values + g.segment); }
			else {
				if (lane == 0) { publish<U>(mine, tile_has_aggregate, tile_aggregate); }
				exclusive = look_back<U>(descriptors, tile, g.tile_in_segment);
			}
			if (lane == 0) {
				publish<U>(mine, tile_has_inclusive, (uint_t<U>)(exclusive + tile_aggregate));
				shared_tile_exclusive[parity] = exclusive;
				shared_ticket[parity ^ 1u] = atomicAdd(ticket_counter, 1ull);
			}
		}
		__syncthreads();
		scan_and_store<U, C, E, Rows>(decompressed, [&](unsigned r) { return deltas[r]; }, g, warp_offset, (uint_t<U>)(shared_tile_exclusive[parity] + preceding_warps));
	}
}

// Cooperative launch, statically assigned tiles, deltas staged in shared memory by TMA: a warp owns a contiguous chunk of
// every tile (Rows x 512 bytes of deltas) and fetches it with ONE bulk copy (cp.async.bulk, SASS UBLKCP) completing on its
// own mbarrier, `look_back_stages` tiles ahead -- the copy of tile k + stages is issued as soon as the warp has read tile
// k's chunk into registers.  So a CTA has loads in flight during its scan-and-store phases too, nothing but one tile's
// deltas is ever held in registers (no spills), and a tile's loads are one request per warp instead of Rows per thread.
//
// Software pipeline, one CTA barrier per tile.  While warp 0 looks back for tile k, the other warps reduce tile k+1's
// staged deltas; after the barrier tile k+1's aggregate is published -- a whole scan-and-store phase before its own
// look-back begins -- and only then tile k is scanned and stored.
//
// Requires the alignment of the wide-piece shape: 16 bytes of deltas per lane and row (E * C == 16).
// Dynamic shared memory: look_back_stages x (the tile's deltas) = GIDDY_DELTA_LOOK_BACK_SHARED_BYTES.
template <unsigned U, unsigned C, unsigned E, unsigned Rows>
__device__ __forceinline__ void decode_tiles_staged(
	uint_t<U>*        __restrict__  decompressed,
	const uint_t<C>*  __restrict__  compressed_input,
	const uint_t<U>*  __restrict__  anchor_values,
	uint64_t                        length,
	uint64_t                        segment_length,
	void*                           workspace)
{
	static_assert(E * C == 16, "a lane's piece of a row is one 16-byte vector of deltas");
	constexpr unsigned stages = look_back_stages;
	extern __shared__ __align__(128) unsigned char staged_deltas[]; // [stages][warps of the CTA][Rows][32 lanes] x 16 bytes
	__shared__ uint_t<U> totals[3][32];          // [tile ordinal % 3][warp]: written an iteration ahead, read in two iterations
	__shared__ uint_t<U> exclusives[2];
	__shared__ uint64_t  arrived[stages][32];    // [stage][warp]: the mbarrier a warp's chunk completes on

	const unsigned lane = lane_index(), warp = warp_index_in_block(), num_warps = warps_per_block();
	constexpr unsigned row_elements = warp_size * E;
	constexpr unsigned chunk_elements = Rows * row_elements;
	constexpr unsigned chunk_bytes = chunk_elements * C;
	const tiling_t tiling(length, segment_length, (uint64_t) num_warps * chunk_elements);
	tile_descriptor_t* const descriptors = reinterpret_cast<tile_descriptor_t*>(static_cast<unsigned char*>(workspace) + 32);
	const uint64_t warp_offset = (uint64_t) warp * chunk_elements;

	auto chunk = [&](unsigned stage) { return staged_deltas + ((size_t) stage * num_warps + warp) * chunk_bytes; };
	// the whole warp calls this, having finished with whatever the stage held before
	auto request = [&](const tile_t& g, unsigned stage) {
		const uint64_t available = g.length > warp_offset ? minimum<uint64_t>(g.length - warp_offset, chunk_elements) : 0;
		if (available == chunk_elements) {
			if (lane == 0) {
				fence_proxy_async_shared();
				mbarrier_expect_bytes(&arrived[stage][warp], chunk_bytes);
				tma_bulk_load(chunk(stage), compressed_input + g.start + warp_offset, chunk_bytes, &arrived[stage][warp]);
			}
		}
		else { // a segment's last, ragged tile: element by element, zero-filled
			uint_t<C>* const destination = reinterpret_cast<uint_t<C>*>(chunk(stage));
			for (unsigned i = lane; i < chunk_elements; i += warp_size) {
				destination[i] = (i < available) ? compressed_input[g.start + warp_offset + i] : (uint_t<C>) 0;
			}
			__syncwarp();
			if (lane == 0) { mbarrier_arrive(&arrived[stage][warp]); }
		}
	};
	auto staged_piece = [&](unsigned stage, unsigned r) {
		const uint4 t = *reinterpret_cast<const uint4*>(chunk(stage) + ((size_t) r * warp_size + lane) * 16);
		byte_pack<16> v;
		v.u32[0] = t.x; v.u32[1] = t.y; v.u32[2] = t.z; v.u32[3] = t.w;
		return v;
	};
	auto publish_totals = [&](unsigned stage, unsigned slot) {
		uint_t<U> lane_total = 0;
		#pragma unroll
		for (unsigned r = 0; r < Rows; r++) {
			const byte_pack<16> v = staged_piece(stage, r);
			#pragma unroll
			for (unsigned j = 0; j < E; j++) { lane_total = (uint_t<U>)(lane_total + (uint_t<U>) v.template get<C>(j)); }
		}
		const uint_t<U> warp_total = warp_sum(lane_total);
		if (lane == 0) { totals[slot][warp] = warp_total; }
	};
	// thread 0, after the barrier: the tile's aggregate -- or, for a segment's first tile, already its inclusive prefix
	auto publish_early = [&](uint64_t tile, const tile_t& g, unsigned slot) {
		uint_t<U> aggregate = 0;
		for (unsigned w = 0; w < num_warps; w++) { aggregate = (uint_t<U>)(aggregate + totals[slot][w]); }
		if (g.tile_in_segment == 0) { publish<U>(descriptors + tile, tile_has_inclusive, (uint_t<U>)(__ldg(anchor_values + g.segment) + aggregate)); }
		else { publish<U>(descriptors + tile, tile_has_aggregate, aggregate); }
	};

	uint64_t tile = blockIdx.x;
	if (tile >= tiling.num_tiles) { return; }
	if (lane == 0) {
		#pragma unroll
		for (unsigned s = 0; s < stages; s++) { mbarrier_init(&arrived[s][warp], 1); }
	}
	__syncwarp();
	tile_t g = tiling.locate(tile);
	#pragma unroll
	for (unsigned s = 0; s < stages; s++) {
		const uint64_t ahead = tile + (uint64_t) s * gridDim.x;
		if (ahead < tiling.num_tiles) { request(tiling.locate(ahead), s); }
	}
	mbarrier_wait(&arrived[0][warp], 0);
	publish_totals(0, 0);
	__syncthreads();
	if (threadIdx.x == 0) { publish_early(tile, g, 0); }

	for (unsigned ordinal = 0; ; ordinal++) {
		const unsigned slot = ordinal % 3, next_slot = (ordinal + 1) % 3, parity = ordinal & 1u;
		const unsigned stage = ordinal % stages, next_stage = (ordinal + 1) % stages;
		const uint64_t next_tile = tile + gridDim.x;
		const bool has_next = next_tile < tiling.num_tiles;
		tile_t next_g = g;
		if (has_next) { next_g = tiling.locate(next_tile); }
		GIDDY_DELTA_TRACE_POINT(tile, 0);
		if (warp == 0) {
			uint_t<U> exclusive;
			if (g.tile_in_segment == 0) { exclusive = __ldg(anchor_values + g.segment); }
			else {
				exclusive = look_back<U>(descriptors, tile, g.tile_in_segment);
				if (lane == 0) {
					uint_t<U> aggregate = 0;
					for (unsigned w = 0; w < num_warps; w++) { aggregate = (uint_t<U>)(aggregate + totals[slot][w]); }
					publish<U>(descriptors + tile, tile_has_inclusive, (uint_t<U>)(exclusive + aggregate));
				}
			}
			if (lane == 0) { exclusives[parity] = exclusive; }
		}
		GIDDY_DELTA_TRACE_POINT(tile, 1);
		if (has_next) {
			mbarrier_wait(&arrived[next_stage][warp], ((ordinal + 1) / stages) & 1u);
			publish_totals(next_stage, next_slot);
		}
		GIDDY_DELTA_TRACE_POINT(tile, 2);
		__syncthreads();
		GIDDY_DELTA_TRACE_POINT(tile, 3);
		if (has_next && threadIdx.x == 0) { publish_early(next_tile, next_g, next_slot); }
		GIDDY_DELTA_TRACE_POINT(tile, 4);
		uint_t<U> carry = exclusives[parity];
		for (unsigned w = 0; w < warp; w++) { carry = (uint_t<U>)(carry + totals[slot][w]); }
		// this tile's deltas: shared memory -> registers, which frees the stage for the tile `stages` ahead
		byte_pack<16> deltas[Rows];
		#pragma unroll
		for (unsigned r = 0; r < Rows; r++) { deltas[r] = staged_piece(stage, r); }
		__syncwarp();
		const uint64_t ahead = tile + (uint64_t) stages * gridDim.x;
		if (ahead < tiling.num_tiles) { request(tiling.locate(ahead), stage); }
		scan_and_store<U, C, E, Rows>(decompressed, [&](unsigned r) { return deltas[r]; }, g, warp_offset, carry);
		GIDDY_DELTA_TRACE_POINT(tile, 5);
		if (!has_next) { return; }
		tile = next_tile; g = next_g;
	}
}

} // namespace delta_detail

// A lane's piece of a row is "wide" where alignment allows -- at least 32 output bytes (256-bit stores) and a whole 16-byte
// vector of deltas: half (or less) the load, store and shuffle instructions per byte -- else 16 output bytes, else one
// element; a tile holds the same number of elements in all three shapes.
template <unsigned U, unsigned C>
struct look_back_shapes {
	static constexpr unsigned E = 16 / U;
	static constexpr unsigned rows = delta_detail::look_back_rows;
	static constexpr unsigned wide = (16 / C > 2 * E) ? 16 / C : 2 * E; // elements of a wide piece
	static constexpr unsigned wide_rows = rows * E / wide;
	static_assert((rows * E) % wide == 0, "a tile is also cut into rows of wide pieces");

	static __host__ __device__ bool vectorizable(const void* decompressed, const void* compressed_input, uint64_t segment_length)
	{
		return reinterpret_cast<uintptr_t>(decompressed) % 16 == 0 && reinterpret_cast<uintptr_t>(compressed_input) % (E * C) == 0 && segment_length % E == 0;
	}
	static __host__ __device__ bool wide_pieces(const void* decompressed, const void* compressed_input, uint64_t segment_length)
	{
		return reinterpret_cast<uintptr_t>(decompressed) % 32 == 0 && reinterpret_cast<uintptr_t>(compressed_input) % (wide * C) == 0 && segment_length % wide == 0;
	}
};

// Any grid; workspace as above.
template <unsigned IndexSize, unsigned U, unsigned C>
__global__ void __launch_bounds__(GIDDY_DELTA_LOOK_BACK_BLOCK_THREADS, GIDDY_DELTA_LOOK_BACK_BLOCKS_PER_MULTIPROCESSOR)
delta_decompress_long_segments(
	uint_t<U>*        __restrict__  decompressed,
	const uint_t<C>*  __restrict__  compressed_input,
	const uint_t<U>*  __restrict__  anchor_values,
	size_type<IndexSize>            length,
	size_type<IndexSize>            segment_length,
	void*                           workspace)
{
	using shapes = look_back_shapes<U, C>;
	if (shapes::wide_pieces(decompressed, compressed_input, segment_length)) {
		delta_detail::decode_tiles_one_at_a_time<U, C, shapes::wide, shapes::wide_rows>(decompressed, compressed_input, anchor_values, length, segment_length, workspace);
	}
	else if (shapes::vectorizable(decompressed, compressed_input, segment_length)) {
		delta_detail::decode_tiles_one_at_a_time<U, C, shapes::E, shapes::rows>(decompressed, compressed_input, anchor_values, length, segment_length, workspace);
	}
	else {
		delta_detail::decode_tiles_one_at_a_time<U, C, 1, shapes::rows>(decompressed, compressed_input, anchor_values, length, segment_length, workspace);
	}
}

// Columns aligned for wide pieces (look_back_shapes::wide_pieces -- the launcher checks) only; to be launched cooperatively,
// with at most as many CTAs as are resident at once and GIDDY_DELTA_LOOK_BACK_SHARED_BYTES of dynamic shared memory.
template <unsigned IndexSize, unsigned U, unsigned C>
__global__ void __launch_bounds__(GIDDY_DELTA_LOOK_BACK_BLOCK_THREADS, GIDDY_DELTA_LOOK_BACK_BLOCKS_PER_MULTIPROCESSOR)
delta_decompress_long_segments_staged(
	uint_t<U>*        __restrict__  decompressed,
	const uint_t<C>*  __restrict__  compressed_input,
	const uint_t<U>*  __restrict__  anchor_values,
	size_type<IndexSize>            length,
	size_type<IndexSize>            segment_length,
	void*                           workspace)
{
	using shapes = look_back_shapes<U, C>;
	delta_detail::decode_tiles_staged<U, C, shapes::wide, shapes::wide_rows>(decompressed, compressed_input, anchor_values, length, segment_length, workspace);
}

} // namespace kernels
} // namespace giddy

#endif // GIDDY_B200_KERNELS_DELTA_CUH_
