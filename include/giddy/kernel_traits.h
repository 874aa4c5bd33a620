/*
 * giddy/kernel_traits.h -- the few compile-time constants of the device kernels which the host-side
 * launch-configuration policy must agree with.  Included by both sides (C, C++ and CUDA compatible).
 */
#ifndef GIDDY_KERNEL_TRAITS_H_
#define GIDDY_KERNEL_TRAITS_H_

/* Position-aligned decoders (FOR, Model, DZB-fixed, Dictionary): a thread produces this many
 * 16-byte output vectors per tile; a CTA owns blockDim.x * this many vectors per grid-stride step. */
#define GIDDY_ELEMENTWISE_UNROLL 4
#define GIDDY_ELEMENTWISE_BLOCK_THREADS 256

/* Run-based decoders (RLE, RPE): a warp decodes a segment in tiles of this many positions ... */
#define GIDDY_RUNS_TILE_POSITIONS 1024
/* ... staging at most this many runs per tile ... */
#define GIDDY_RUNS_STAGED_CAPACITY (GIDDY_RUNS_TILE_POSITIONS + 32)
/* ... in this much dynamic shared memory per warp, for elements of U bytes. */
#define GIDDY_RUNS_SHARED_BYTES_PER_WARP(U) (2 * (GIDDY_RUNS_TILE_POSITIONS / 32) * 4 + GIDDY_RUNS_STAGED_CAPACITY * (U))
#define GIDDY_RUNS_WARPS_PER_BLOCK 4

/* Warp-per-segment decoders (Delta, DZB-variable) and warp-per-tile decoders (incidence bitmaps) */
#define GIDDY_SEGMENT_WARPS_PER_BLOCK 8
#define GIDDY_BITMAPS_CONTAINERS_PER_TILE 32

/* DZB-variable: a warp decodes a segment in tiles of this many elements, staging each tile's bytes,
 * per-vector byte offsets and size fields in this much dynamic shared memory per warp (U = uncompressed
 * element bytes): data (tile x U + 64), offsets (2 bytes per 16-byte output vector), size fields (tile x 4
 * bits), one mbarrier; rounded up to a multiple of 128. */
#define GIDDY_DZBV_TILE_ELEMENTS 1024
#define GIDDY_DZBV_SHARED_BYTES_PER_WARP(U) \
	((((GIDDY_DZBV_TILE_ELEMENTS * (U) + 64) + (2 * GIDDY_DZBV_TILE_ELEMENTS * (U) / 16) + (GIDDY_DZBV_TILE_ELEMENTS / 2) + 16) + 127) / 128 * 128)

/* Patching kernels */
#define GIDDY_SCATTER_UNROLL 8
#define GIDDY_SCATTER_BLOCK_THREADS 256
#define GIDDY_CI_SCATTER_WARPS_PER_BLOCK 4

/* Delta: a lane scans this many rows of vectors per step (a warp per anchored segment).  Columns with few, long segments
 * (so that a warp per segment could not fill the GPU) are decoded tile by tile with decoupled look-back instead: CTAs of
 * GIDDY_DELTA_LOOK_BACK_BLOCK_THREADS threads, tiles of threads * GIDDY_DELTA_LOOK_BACK_ROWS * (16 / U) elements, this many
 * CTAs per multiprocessor, launched cooperatively with statically assigned tiles (a ticket counter where cooperative
 * launches are not to be had).  The #ifndef guards let benchmarks/micro/lookback_probe.cu build the kernel under other
 * settings. */
#define GIDDY_DELTA_ROWS_PER_STEP 4
#ifndef GIDDY_DELTA_LOOK_BACK_BLOCK_THREADS
#define GIDDY_DELTA_LOOK_BACK_BLOCK_THREADS 256
#endif
#ifndef GIDDY_DELTA_LOOK_BACK_BLOCKS_PER_MULTIPROCESSOR
#define GIDDY_DELTA_LOOK_BACK_BLOCKS_PER_MULTIPROCESSOR 2
#endif
#ifndef GIDDY_DELTA_LOOK_BACK_ROWS
#define GIDDY_DELTA_LOOK_BACK_ROWS 24
#endif
#ifndef GIDDY_DELTA_LOOK_BACK_WINDOW
#define GIDDY_DELTA_LOOK_BACK_WINDOW 4             /* descriptors inspected per lane per look-back round */
#endif
#ifndef GIDDY_DELTA_LOOK_BACK_STAGES
#define GIDDY_DELTA_LOOK_BACK_STAGES 2             /* tiles of deltas a CTA of the cooperative kernel stages in shared memory */
#endif
/* dynamic shared memory of the staged kernel: the deltas of a tile are tile elements x C bytes */
#define GIDDY_DELTA_LOOK_BACK_SHARED_BYTES(U, C, threads) (GIDDY_DELTA_LOOK_BACK_STAGES * GIDDY_DELTA_TILE_ELEMENTS(U, threads) * (C))
#define GIDDY_DELTA_TILE_ELEMENTS(U, threads) ((threads) * GIDDY_DELTA_LOOK_BACK_ROWS * (16 / (U)))
#define GIDDY_DELTA_USES_LOOK_BACK(length, segment_length, U) \
	((segment_length) >= 2 * GIDDY_DELTA_TILE_ELEMENTS(U, GIDDY_DELTA_LOOK_BACK_BLOCK_THREADS) && \
	 ((length) + (segment_length) - 1) / (segment_length) < 8192)

/* Compressors: RLE / RPE whose runs may be at least a tile long in one pass with decoupled look-back (0: two passes around scans) */
#ifndef GIDDY_RUN_ENCODERS_SINGLE_PASS
#define GIDDY_RUN_ENCODERS_SINGLE_PASS 1
#endif

/* Patched decoders (decode + patches in one pass): warp tiles of 32 x this many output vectors go grid-stride to the
 * warps of this many CTAs (of GIDDY_ELEMENTWISE_BLOCK_THREADS threads) per multiprocessor; an index of one patch range per
 * tile is built into a workspace before the decode. */
#ifndef GIDDY_PATCHED_UNROLL
#define GIDDY_PATCHED_UNROLL 2
#endif
#ifndef GIDDY_PATCHED_INDEX_BLOCK_THREADS
#define GIDDY_PATCHED_INDEX_BLOCK_THREADS 64       /* expand_and_index_compressed_indices: a CTA per anchoring period */
#endif
#ifndef GIDDY_PATCHED_EARLY_TILES
#define GIDDY_PATCHED_EARLY_TILES 2                /* tiles a warp decodes ahead of the tile index (programmatic dependent launch) */
#endif
#ifndef GIDDY_PATCHED_BLOCKS_PER_MULTIPROCESSOR
#define GIDDY_PATCHED_BLOCKS_PER_MULTIPROCESSOR 4
#endif

#endif /* GIDDY_KERNEL_TRAITS_H_ */
