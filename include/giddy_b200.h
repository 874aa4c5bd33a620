/*
 * giddy_b200.h -- the C ABI of libgiddy_b200.so
 *
 * A Blackwell-native (sm_100a) implementation of the decompression path of
 * eyalroz/libgiddy.  Everything exported by the shared library is declared in
 * this header, `extern "C"`, with plain pointers and sizes only.
 *
 * Groups of entry points ((1) and (2) mirror the reference; (3) fused decoders and (4) host-resident scans are extensions):
 *
 *  (1) The *registry* API ("mode iii" of the reference, README.md:116-126):
 *      string-keyed factory -> type-erased kernel wrapper -> resolve a launch
 *      configuration from named scalar arguments -> enqueue a launch with named,
 *      exactly-typed arguments.  This is what a foreign-language binding
 *      (ctypes / cgo / JNI / N-API) of the reference would bind; it replaces
 *        cuda::registered::kernel_t::produceSubclass   src/util/FactoryProducible.h:63-70
 *        cuda::kernel_t::resolve_launch_configuration  src/cuda/kernel_wrapper.cuh:94-123
 *        cuda::kernel_t::enqueue_launch                src/cuda/kernel_wrapper.cuh:68-71
 *        cuda::kernel_t::get_device_function           src/cuda/kernel_wrapper.cuh:125
 *
 *  (2) The *kernel* API: one launcher per scheme taking the reference kernel's
 *      parameters in the reference kernel's order, preceded by the stream, the
 *      launch configuration and the template parameters as run-time byte sizes.
 *      The C++ wrapper classes (include/giddy/) call only these.  Each launcher
 *      cites the reference __global__ function it replaces.
 *
 *  (3) Fused composite decoders, (4) host-resident scans, (5) shard planning, (6) host memory placement,
 *      (7) device groups, (8) compressors, (9) utility kernels: see the section comments below.
 *
 * Conventions: every function returns a giddy_status_t (0 = success) unless it
 * returns a pointer/count; on failure giddy_last_error() holds a message for the
 * calling thread.  Launches are asynchronous on the given stream and -- like the
 * reference (external/cuda-api-wrappers/src/cuda/api/kernel_launch.cuh:107-112) --
 * execution errors surface at the caller's next synchronisation; launch
 * *configuration* errors are reported immediately.  The library owns no column
 * data and never synchronises; the only device memory it ever takes is the Delta look-back workspace (a few
 * KB per launch, stream-ordered, from a pool of its own) and the only copies it makes are those a caller records
 * in a scan (group 4).  There is no CPU fallback.
 */
#ifndef GIDDY_B200_H_
#define GIDDY_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define GIDDY_API
#else
#define GIDDY_API __attribute__((visibility("default")))
#endif

/* ------------------------------------------------------------------------- */
/* Status codes: one per exception type the reference's C++ API throws        */
/* (SURVEY.md 8b "Errors").                                                    */
/* ------------------------------------------------------------------------- */
typedef enum giddy_status_t {
	GIDDY_SUCCESS            = 0,
	GIDDY_INVALID_ARGUMENT   = 1, /* std::invalid_argument: unknown factory key, bad scheme parameter, impossible configuration */
	GIDDY_OUT_OF_RANGE       = 2, /* std::out_of_range: a required named argument is missing (unordered_map::at) */
	GIDDY_BAD_ANY_CAST       = 3, /* boost::bad_any_cast: a named argument does not have exactly the expected type */
	GIDDY_LOGIC_ERROR        = 4, /* std::logic_error: e.g. a serialization factor for a kernel which admits none */
	GIDDY_CUDA_ERROR         = 5, /* cuda::runtime_error: a CUDA runtime API call failed */
	GIDDY_UNSUPPORTED        = 6, /* the kernel API was asked for a size combination which is not instantiated */
	GIDDY_INTERNAL_ERROR     = 7
} giddy_status_t;

GIDDY_API const char* giddy_last_error(void);
GIDDY_API const char* giddy_version(void);

/* ------------------------------------------------------------------------- */
/* Plain-data mirrors of the types in the reference's signatures              */
/* ------------------------------------------------------------------------- */

/* cuda::launch_configuration_t, external/cuda-api-wrappers/src/cuda/api/types.h:99-103.
 * Only linear grids are ever produced (src/cuda/linear_grid.h).  The dynamic shared
 * memory size is widened from the reference's unsigned short (types.h:93) to 32 bits:
 * B200 allows 227 KiB per block. */
typedef struct giddy_launch_config_t {
	unsigned grid_dimensions[3];
	unsigned block_dimensions[3];
	unsigned dynamic_shared_memory_size;
} giddy_launch_config_t;

/* cuda::launch_configuration_limits_t, src/cuda/kernel_wrapper.cuh:20-33.
 * A zero "has_" flag means boost::none. */
typedef struct giddy_launch_config_limits_t {
	int      has_block_size;             unsigned block_size;
	int      has_grid_size;              unsigned grid_size;
	int      has_dynamic_shared_memory;  unsigned dynamic_shared_memory;
} giddy_launch_config_limits_t;

/* The subset of cuda::device::properties_t (cudaDeviceProp) the resolver reads. */
typedef struct giddy_device_properties_t {
	int    compute_capability_major;
	int    compute_capability_minor;
	int    multiprocessor_count;
	int    max_threads_per_block;
	int    max_threads_dim_x;
	int    max_threads_per_multiprocessor;
	int    max_blocks_per_multiprocessor;
	int    warp_size;
	size_t shared_memory_per_block;
	size_t shared_memory_per_block_optin;
	size_t shared_memory_per_multiprocessor;
	size_t l2_cache_size;
	size_t total_global_memory;
} giddy_device_properties_t;

/* The subset of cuda::device_function::attributes_t (cudaFuncAttributes). */
typedef struct giddy_function_attributes_t {
	int    max_threads_per_block;
	int    num_regs;
	size_t shared_size_bytes;               /* static shared memory */
	size_t const_size_bytes;
	size_t local_size_bytes;
	int    max_dynamic_shared_size_bytes;
	int    ptx_version;
	int    binary_version;
} giddy_function_attributes_t;

/* Properties of a device / of the current device (device_id < 0). */
GIDDY_API giddy_status_t giddy_device_properties(int device_id, giddy_device_properties_t* out);

/* Attributes of a kernel returned by one of the *_kernel() getters below;
 * replaces cuda::device_function_t::attributes(), external/.../device_function.hpp:73-79.
 * The kernels are registered with this library's CUDA runtime, so callers must ask here. */
GIDDY_API giddy_status_t giddy_function_attributes(const void* kernel, giddy_function_attributes_t* out);

/* Launch a kernel returned by one of the *_kernel() getters with a caller-built argument
 * pack ("mode ii", src/cuda/kernel_wrapper.cuh:135-147).  `args` is cudaLaunchKernel's. */
GIDDY_API giddy_status_t giddy_launch_kernel(const void* kernel, void* stream,
	const giddy_launch_config_t* config, void** args);

/* ------------------------------------------------------------------------- */
/* (1) Registry API                                                           */
/* ------------------------------------------------------------------------- */

/* Exact static type of a named argument.  boost::any_cast demands the exact type
 * (src/cuda/kernel_wrapper.cuh:47-49), and so does this API: passing GIDDY_T_U32
 * where the wrapper expects size_t fails with GIDDY_BAD_ANY_CAST.
 * size_t is GIDDY_T_U64; unsigned is GIDDY_T_U32; unsigned short is GIDDY_T_U16;
 * unsigned char is GIDDY_T_U8. */
typedef enum giddy_type_t {
	GIDDY_T_U8 = 1, GIDDY_T_U16, GIDDY_T_U32, GIDDY_T_U64,
	GIDDY_T_I8,     GIDDY_T_I16, GIDDY_T_I32, GIDDY_T_I64,
	GIDDY_T_BOOL,
	/* cuda::array<T, N> by value (model coefficients, src/cuda/stl_array.cuh:15-18) */
	GIDDY_T_ARRAY_I16_1, GIDDY_T_ARRAY_I16_2, GIDDY_T_ARRAY_I32_1, GIDDY_T_ARRAY_I32_2,
	/* pointers: T* = GIDDY_T_PTR | element type; const T* additionally | GIDDY_T_CONST */
	GIDDY_T_PTR   = 0x100,
	GIDDY_T_CONST = 0x200
} giddy_type_t;

typedef struct giddy_arg_t {
	const char* name;
	int         type;      /* a giddy_type_t, possibly or-ed with GIDDY_T_PTR [| GIDDY_T_CONST] */
	union {
		uint64_t    u;     /* unsigned scalars and bool */
		int64_t     i;     /* signed scalars */
		const void* p;     /* pointers (device memory) */
		int32_t     a[2];  /* cuda::array<int16|int32, 1|2> by value */
	} value;
} giddy_arg_t;

typedef struct giddy_kernel giddy_kernel_t;

/* The factory's key set (163 keys), in sorted order.
 * Replaces getSubclassFactory().getInsantiators(), src/util/Factory.h:98-100. */
GIDDY_API size_t      giddy_num_registered_kernels(void);
GIDDY_API const char* giddy_registered_kernel_key(size_t index);
GIDDY_API int         giddy_can_produce(const char* key);

/* produceSubclass(key).  Unknown key -> GIDDY_INVALID_ARGUMENT, *out = NULL. */
GIDDY_API giddy_status_t giddy_produce_subclass(const char* key, giddy_kernel_t** out);
GIDDY_API void           giddy_kernel_destroy(giddy_kernel_t* kernel);
GIDDY_API const char*    giddy_kernel_key(const giddy_kernel_t* kernel);

/* get_device_function(): the address of the __global__ instantiation. */
GIDDY_API const void*    giddy_kernel_get_device_function(const giddy_kernel_t* kernel);

/* resolve_launch_configuration(device_properties, kernel_function_attributes, extra_arguments, limits).
 * device_properties == NULL: use the current device's.  attributes == NULL: query them
 * (the 3-argument overload, src/cuda/kernel_wrapper.cuh:94-102).  limits == NULL: none. */
GIDDY_API giddy_status_t giddy_kernel_resolve_launch_configuration(
	const giddy_kernel_t*               kernel,
	const giddy_device_properties_t*    device_properties,
	const giddy_function_attributes_t*  kernel_function_attributes,
	const giddy_arg_t*                  extra_arguments,
	size_t                              num_extra_arguments,
	const giddy_launch_config_limits_t* limits,
	giddy_launch_config_t*              out);

/* enqueue_launch(stream, launch_config, arguments).  The stream is assumed to belong to
 * the current device (src/kernel_wrappers/registered_wrapper.h:82). */
GIDDY_API giddy_status_t giddy_kernel_enqueue_launch(
	const giddy_kernel_t*         kernel,
	void*                         stream,
	const giddy_launch_config_t*  launch_config,
	const giddy_arg_t*            arguments,
	size_t                        num_arguments);

/* ------------------------------------------------------------------------- */
/* (2) Kernel API.  Sizes are template parameters of the reference, in bytes.  */
/* Lengths are passed as 64-bit values and narrowed to size_type_by_index_size */
/* <IndexSize> (src/kernels/common.cuh:49-53) inside.                          */
/* ------------------------------------------------------------------------- */

typedef enum giddy_model_t { GIDDY_MODEL_CONSTANT = 0, GIDDY_MODEL_LINEAR = 1 } giddy_model_t;

/* decompression::frame_of_reference::decompress<IndexSize, Uncompressed, Compressed, UnaryModelFunction>
 * src/kernels/decompression/frame_of_reference.cuh:35-80.  Coefficients are int32 (constant: 1, linear: 2
 * per segment).  The reference's derived parameter num_segments_to_decompress_per_block is not needed. */
GIDDY_API const void* giddy_b200_frame_of_reference_kernel(
	unsigned index_size, unsigned uncompressed_size, unsigned compressed_size,
	giddy_model_t model, unsigned model_index_size);
GIDDY_API giddy_status_t giddy_b200_frame_of_reference_launch(
	void* stream, const giddy_launch_config_t* config,
	unsigned index_size, unsigned uncompressed_size, unsigned compressed_size,
	giddy_model_t model, unsigned model_index_size,
	void* decompressed, const void* compressed_input, const void* segment_model_coefficients,
	uint64_t length, uint64_t segment_length);

/* decompression::model::decompress<IndexSize, Uncompressed, UnaryModelFunction>
 * src/kernels/decompression/model.cuh:38-53.  Coefficient type == Uncompressed; passed by value. */
GIDDY_API const void* giddy_b200_model_kernel(
	unsigned index_size, unsigned uncompressed_size, giddy_model_t model);
GIDDY_API giddy_status_t giddy_b200_model_launch(
	void* stream, const giddy_launch_config_t* config,
	unsigned index_size, unsigned uncompressed_size, giddy_model_t model,
	void* decompressed, const void* model_coefficients /* host pointer to 1 or 2 coefficients */,
	uint64_t length);

/* decompression::run_length_encoding::decompress<UncompressedIndexSize, UncompressedSize, RunLengthSize, RunIndexSize>
 * src/kernels/decompression/run_length_encoding.cuh:342-374 */
GIDDY_API const void* giddy_b200_run_length_encoding_kernel(
	unsigned uncompressed_index_size, unsigned uncompressed_size, unsigned run_length_size, unsigned run_index_size);
GIDDY_API giddy_status_t giddy_b200_run_length_encoding_launch(
	void* stream, const giddy_launch_config_t* config,
	unsigned uncompressed_index_size, unsigned uncompressed_size, unsigned run_length_size, unsigned run_index_size,
	void* decompressed, const void* run_data, const void* run_lengths,
	const void* segment_position_anchors, const void* segment_position_anchor_intra_run_offsets,
	uint64_t segment_length, uint64_t num_segments, uint64_t num_element_runs, uint64_t total_uncompressed_length);

/* decompression::run_position_encoding::decompress<UncompressedIndexSize, UncompressedSize, PositionOffsetSize, PositionsAreRelative, RunIndexSize>
 * src/kernels/decompression/run_position_encoding.cuh:327-359 */
GIDDY_API const void* giddy_b200_run_position_encoding_kernel(
	unsigned uncompressed_index_size, unsigned uncompressed_size, unsigned position_offset_size,
	int positions_are_relative, unsigned run_index_size);
GIDDY_API giddy_status_t giddy_b200_run_position_encoding_launch(
	void* stream, const giddy_launch_config_t* config,
	unsigned uncompressed_index_size, unsigned uncompressed_size, unsigned position_offset_size,
	int positions_are_relative, unsigned run_index_size,
	void* decompressed, const void* run_data, const void* run_start_positions, const void* position_anchors,
	uint64_t anchoring_period, uint64_t num_anchors, uint64_t num_element_runs, uint64_t uncompressed_length);

/* decompression::delta::decompress<IndexSize, UncompressedSize, CompressedSize>
 * src/kernels/decompression/delta.cuh:66-108 */
GIDDY_API const void* giddy_b200_delta_kernel(
	unsigned index_size, unsigned uncompressed_size, unsigned compressed_size);
GIDDY_API giddy_status_t giddy_b200_delta_launch(
	void* stream, const giddy_launch_config_t* config,
	unsigned index_size, unsigned uncompressed_size, unsigned compressed_size,
	void* decompressed, const void* compressed_input, const void* anchor_values,
	uint64_t length, uint64_t segment_length);

/* decompression::dictionary::decompress<IndexSize, UncompressedSize, DictionaryIndexSize>
 * src/kernels/decompression/dictionary.cuh:61-75.  cache_dictionary_in_shared_memory is honoured
 * (the reference's shared-memory path is dead code, gather.cuh:33-90); serialization_factor is
 * accepted and ignored, as in the reference (gather.cuh:24). */
GIDDY_API const void* giddy_b200_dictionary_kernel(
	unsigned index_size, unsigned uncompressed_size, unsigned dictionary_index_size);
GIDDY_API giddy_status_t giddy_b200_dictionary_launch(
	void* stream, const giddy_launch_config_t* config,
	unsigned index_size, unsigned uncompressed_size, unsigned dictionary_index_size,
	void* decompressed, const void* compressed_input, const void* dictionary_entries,
	uint64_t length, uint64_t num_dictionary_entries,
	int cache_dictionary_in_shared_memory, unsigned serialization_factor);

/* decompression::discard_zero_bytes::fixed_width::decompress<IndexSize, UncompressedSize, CompressedSize, little, 32>
 * src/kernels/decompression/discard_zero_bytes_fixed.cuh:108-124 */
GIDDY_API const void* giddy_b200_discard_zero_bytes_fixed_kernel(
	unsigned index_size, unsigned uncompressed_size, unsigned compressed_size);
GIDDY_API giddy_status_t giddy_b200_discard_zero_bytes_fixed_launch(
	void* stream, const giddy_launch_config_t* config,
	unsigned index_size, unsigned uncompressed_size, unsigned compressed_size,
	void* decompressed, const void* compressed_input, uint64_t length);

/* decompression::discard_zero_bytes::variable_width::decompress<IndexSize, UncompressedSize, ElementSizesContainerSize>
 * src/kernels/decompression/discard_zero_bytes_variable.cuh:381-439 */
GIDDY_API const void* giddy_b200_discard_zero_bytes_variable_kernel(
	unsigned index_size, unsigned uncompressed_size, unsigned element_sizes_container_size);
GIDDY_API giddy_status_t giddy_b200_discard_zero_bytes_variable_launch(
	void* stream, const giddy_launch_config_t* config,
	unsigned index_size, unsigned uncompressed_size, unsigned element_sizes_container_size,
	void* decompressed, const void* compressed_data, const void* packed_element_sizes,
	const void* position_anchors, uint64_t segment_length, uint64_t length,
	unsigned min_represented_element_size, unsigned bits_per_element_size);

/* decompression::incidence_bitmaps::decompress<IndexSize, UncompressedSize, BitmapAlignmentInInts, UseDictionary>
 * src/kernels/decompression/incidence_bitmaps.cuh:69-131 */
GIDDY_API const void* giddy_b200_incidence_bitmaps_kernel(
	unsigned index_size, unsigned uncompressed_size, unsigned bitmap_alignment_in_ints, int use_dictionary);
GIDDY_API giddy_status_t giddy_b200_incidence_bitmaps_launch(
	void* stream, const giddy_launch_config_t* config,
	unsigned index_size, unsigned uncompressed_size, unsigned bitmap_alignment_in_ints, int use_dictionary,
	void* decompressed, const void* bitmaps, const void* dictionary_entries,
	uint64_t bitmap_length, unsigned num_bitmaps);

/* scatter::scatter<OutputIndexSize, ElementSize, InputIndexSize, 16>
 * src/kernels/data_layout/scatter.cuh:26-46 */
GIDDY_API const void* giddy_b200_scatter_kernel(
	unsigned output_index_size, unsigned element_size, unsigned input_index_size);
GIDDY_API giddy_status_t giddy_b200_scatter_launch(
	void* stream, const giddy_launch_config_t* config,
	unsigned output_index_size, unsigned element_size, unsigned input_index_size,
	void* target, const void* data, const void* indices, uint64_t data_length);

/* scatter::compressed_indices::scatter<OutputIndexSize, ElementSize, InputIndexSize, RunLengthSize>
 * src/kernels/data_layout/compressed_indices_scatter.cuh:219-271 */
GIDDY_API const void* giddy_b200_compressed_indices_scatter_kernel(
	unsigned output_index_size, unsigned element_size, unsigned input_index_size, unsigned run_length_size);
GIDDY_API giddy_status_t giddy_b200_compressed_indices_scatter_launch(
	void* stream, const giddy_launch_config_t* config,
	unsigned output_index_size, unsigned element_size, unsigned input_index_size, unsigned run_length_size,
	void* target, const void* data,
	const void* scatter_position_run_lengths,
	const void* scatter_position_run_individual_offset_sizes,
	const void* scatter_position_run_baseline_values,
	const void* scatter_position_run_offsets_start_positions,
	const void* scatter_position_offset_bytes,
	const void* scatter_position_anchors,
	uint64_t anchoring_period, uint64_t num_scatter_position_runs, uint64_t data_length);

/* ------------------------------------------------------------------------- */
/* (3) Extensions: fused composite decoders (no counterpart kernel in the      */
/* reference, which composes schemes by running one kernel after another,      */
/* README.md:55-60).                                                           */
/* ------------------------------------------------------------------------- */

/* delta::decompress<IndexSize, UncompressedSize, CompressedSize> (src/kernels/decompression/delta.cuh:66-108) applied to
 * the output of discard_zero_bytes::variable_width::decompress<IndexSize, CompressedSize, ElementSizesContainerSize>
 * (src/kernels/decompression/discard_zero_bytes_variable.cuh:381-439), same segment length for both, in one pass:
 *   decompressed[i] = anchor_values[s] + sum_{j = s*L .. i} dzbv_element[j]
 * Needs 1-, 2- or 4-bit size fields and segment_length * bits_per_element_size % 32 == 0. */
GIDDY_API const void* giddy_b200_delta_of_discard_zero_bytes_variable_kernel(
	unsigned index_size, unsigned uncompressed_size, unsigned compressed_size, unsigned element_sizes_container_size);
GIDDY_API giddy_status_t giddy_b200_delta_of_discard_zero_bytes_variable_launch(
	void* stream, const giddy_launch_config_t* config,
	unsigned index_size, unsigned uncompressed_size, unsigned compressed_size, unsigned element_sizes_container_size,
	void* decompressed, const void* compressed_data, const void* packed_element_sizes, const void* position_anchors,
	const void* anchor_values, uint64_t segment_length, uint64_t length,
	unsigned min_represented_element_size, unsigned bits_per_element_size);

/* Patched decoders: a decoder and the a-posteriori patching kernel which the reference would run after it
 * (README.md:55-60), in one pass.  A patch is stored right after the tile which contains it, while that tile is
 * still in L2, instead of costing a read-modify-write of a 32-byte HBM sector in a second pass.
 * Both need the patch positions in ASCENDING order. */

/* dictionary::decompress<IndexSize, UncompressedSize, DictionaryIndexSize> (src/kernels/decompression/dictionary.cuh:61-75)
 * then scatter::scatter<IndexSize, UncompressedSize, IndexSize> (src/kernels/data_layout/scatter.cuh:26-46) onto its output:
 *   decompressed[i] = dictionary_entries[compressed_input[i]];  decompressed[patch_indices[k]] = patch_data[k], k < num_patches */
GIDDY_API const void* giddy_b200_patched_dictionary_kernel(
	unsigned index_size, unsigned uncompressed_size, unsigned dictionary_index_size);
GIDDY_API giddy_status_t giddy_b200_patched_dictionary_launch(
	void* stream, const giddy_launch_config_t* config,
	unsigned index_size, unsigned uncompressed_size, unsigned dictionary_index_size,
	void* decompressed, const void* compressed_input, const void* dictionary_entries,
	uint64_t length, uint64_t num_dictionary_entries, int cache_dictionary_in_shared_memory,
	const void* patch_data, const void* patch_indices, uint64_t num_patches);

/* frame_of_reference::decompress<IndexSize, Uncompressed, Compressed, Model> (src/kernels/decompression/frame_of_reference.cuh:35-80)
 * then scatter::compressed_indices::scatter<IndexSize, Uncompressed, InputIndexSize, InputIndexSize>
 * (src/kernels/data_layout/compressed_indices_scatter.cuh:219-271) onto its output.  The patch arguments are that kernel's. */
GIDDY_API const void* giddy_b200_patched_frame_of_reference_kernel(
	unsigned index_size, unsigned uncompressed_size, unsigned compressed_size,
	giddy_model_t model, unsigned model_index_size, unsigned input_index_size);
GIDDY_API giddy_status_t giddy_b200_patched_frame_of_reference_launch(
	void* stream, const giddy_launch_config_t* config,
	unsigned index_size, unsigned uncompressed_size, unsigned compressed_size,
	giddy_model_t model, unsigned model_index_size, unsigned input_index_size,
	void* decompressed, const void* compressed_input, const void* segment_model_coefficients,
	uint64_t length, uint64_t segment_length,
	const void* data_to_scatter,
	const void* scatter_position_run_lengths,
	const void* scatter_position_run_individual_offset_sizes,
	const void* scatter_position_run_baseline_values,
	const void* scatter_position_run_offsets_start_positions,
	const void* scatter_position_offset_bytes,
	const void* scatter_position_anchors,
	uint64_t anchoring_period, uint64_t num_scatter_position_runs, uint64_t data_length);

/* ------------------------------------------------------------------------- */
/* (4) Host-resident scans: compressed columns in pinned host memory, copied   */
/* to the device chunk by chunk on side streams and decoded as they land.      */
/* No counterpart in the reference, which leaves this to the calling DBMS      */
/* (README.md:18-24: ship the data over the bus compressed, decompress on the  */
/* GPU); it is the caller-side half of the drop-in path (BASELINE.json         */
/* configs[4]).  A scan owns streams, events and the recorded operations --    */
/* never column data.                                                          */
/* ------------------------------------------------------------------------- */

typedef struct giddy_scan giddy_scan_t;

/* A scan object is not internally synchronised: build it and run it from one thread at a time (different scans may be
 * used from different threads).  Pointers recorded in a plan (buffers, kernels) must stay valid for as long as it is
 * replayed.
 *
 * A plan with `num_streams` side streams on the current device.  Operations are appended chunk by chunk (chunks are
 * created on first mention) and replayed in the order appended; chunk c runs on side stream c % num_streams. */
GIDDY_API giddy_status_t giddy_scan_create(unsigned num_streams, giddy_scan_t** out);
GIDDY_API void           giddy_scan_destroy(giddy_scan_t* scan);
GIDDY_API unsigned       giddy_scan_num_streams(const giddy_scan_t* scan);
GIDDY_API unsigned       giddy_scan_num_chunks(const giddy_scan_t* scan);
GIDDY_API size_t         giddy_scan_num_operations(const giddy_scan_t* scan);

/* cudaMemcpyAsync(device_destination, host_source, bytes, HostToDevice) on the chunk's stream; host_source should be
 * pinned (cudaHostAlloc / cudaHostRegister) or the copy will not overlap anything. */
GIDDY_API giddy_status_t giddy_scan_add_copy(giddy_scan_t* scan, unsigned chunk,
	void* device_destination, const void* host_source, size_t bytes);

/* The other direction, for (part of) a chunk's result: cudaMemcpyAsync(host_destination, device_source, bytes, DeviceToHost). */
GIDDY_API giddy_status_t giddy_scan_add_copy_back(giddy_scan_t* scan, unsigned chunk,
	void* host_destination, const void* device_source, size_t bytes);

/* giddy_kernel_enqueue_launch(kernel, <the chunk's stream>, launch_config, arguments) at this point of the chunk.
 * The argument array is copied; the kernel object must outlive the scan. */
GIDDY_API giddy_status_t giddy_scan_add_launch(giddy_scan_t* scan, unsigned chunk, const giddy_kernel_t* kernel,
	const giddy_launch_config_t* launch_config, const giddy_arg_t* arguments, size_t num_arguments);

/* Anything else a chunk needs enqueued (a caller's own kernel, a device-to-device copy): enqueue(stream, user_data) is
 * called at every replay and must only enqueue asynchronous work on `stream`. */
typedef giddy_status_t (*giddy_scan_callback_t)(void* stream, void* user_data);
GIDDY_API giddy_status_t giddy_scan_add_callback(giddy_scan_t* scan, unsigned chunk,
	giddy_scan_callback_t enqueue, void* user_data, void (*destroy_user_data)(void*));

/* What is appended to `chunk` from now on waits for what has been appended to `on_chunk` (< chunk) so far -- e.g. the
 * patches of a chunk whose first targets lie in the previous chunk's output. */
GIDDY_API giddy_status_t giddy_scan_add_dependency(giddy_scan_t* scan, unsigned chunk, unsigned on_chunk);

/* Replay as one CUDA graph, captured at the next giddy_scan_run() and re-captured after the plan changes, instead of
 * re-enqueueing every operation (off by default). */
GIDDY_API giddy_status_t giddy_scan_set_graph_replay(giddy_scan_t* scan, int enabled);

/* Enqueue one scan: the side streams fork from `stream`, run their chunks, and are joined back into `stream`.
 * Asynchronous; the current device must be the one the scan was created on. */
GIDDY_API giddy_status_t giddy_scan_run(giddy_scan_t* scan, void* stream);

/* ------------------------------------------------------------------------- */
/* (5) Shard planning: a contiguous range of anchored segments of a column as  */
/* a self-contained column -- for the GPUs of a box, or the chunks of a        */
/* host-resident scan.  No counterpart in the reference, which is single-GPU   */
/* (src/kernel_wrappers/registered_wrapper.h:82); what makes it possible is    */
/* the reference's own formats: every segment decodes from its anchor side     */
/* data alone (run_length_encoding.cuh:284-325, discard_zero_bytes_variable    */
/* .cuh:397-438, delta.cuh:80-92).  Planning reads host copies of the anchor   */
/* arrays only; it never touches a device.                                     */
/* ------------------------------------------------------------------------- */

typedef enum giddy_scheme_t {
	GIDDY_SCHEME_FRAME_OF_REFERENCE = 0, GIDDY_SCHEME_MODEL, GIDDY_SCHEME_RUN_LENGTH_ENCODING, GIDDY_SCHEME_RUN_POSITION_ENCODING,
	GIDDY_SCHEME_DELTA, GIDDY_SCHEME_DICTIONARY, GIDDY_SCHEME_DISCARD_ZERO_BYTES_FIXED, GIDDY_SCHEME_DISCARD_ZERO_BYTES_VARIABLE,
	GIDDY_SCHEME_INCIDENCE_BITMAPS, GIDDY_SCHEME_SCATTER
} giddy_scheme_t;

/* What the planner needs to know about a whole compressed column.  Array names below are the wrappers' argument names
 * (SURVEY.md Appendix B).  Fields a scheme does not use are ignored. */
typedef struct giddy_column_layout_t {
	giddy_scheme_t scheme;
	uint64_t       length;                 /* decompressed elements (SCATTER: length of the target column) */
	uint64_t       segment_length;         /* anchoring period; for schemes without segments, the granularity shards are cut at */
	unsigned       index_size;             /* bytes per element of the anchor array (RLE/RPE: RunIndexSize; DZB-variable: IndexSize; SCATTER: OutputIndexSize) */
	unsigned       uncompressed_size;
	unsigned       compressed_size;        /* FOR/Delta/DZB-fixed: Compressed; Dictionary: DictionaryIndexSize; RLE: RunLengthSize; RPE: PositionOffsetSize; DZB-variable: ElementSizesContainerSize */
	unsigned       model_dimension;        /* FOR: coefficients per segment (1 constant, 2 linear) */
	int            positions_are_relative; /* RPE */
	unsigned       bits_per_element_size;  /* DZB-variable */
	uint64_t       num_runs;               /* RLE/RPE: num_element_runs */
	uint64_t       num_bytes;              /* DZB-variable: bytes of compressed_data */
	unsigned       num_bitmaps;            /* incidence bitmaps */
	unsigned       bitmap_alignment_in_ints;
	uint64_t       num_patches;            /* SCATTER: data_length; indices must be ascending */
	const void*    anchors;                /* HOST copy of position_anchors (RLE, RPE, DZB-variable); SCATTER: of indices */
	const void*    run_start_positions;    /* HOST copy; RPE with relative positions only */
} giddy_column_layout_t;

/* A part of one of the column's arrays: `rows` rows of `count` elements starting at element `first`, `pitch` elements apart
 * (rows == 1 but for incidence bitmaps).  Copied into a shard's own buffer, every element has `rebase` subtracted (anchors,
 * absolute positions, target indices; elements smaller than `rebase` become 0 when `saturate` is set). */
typedef struct giddy_slice_t {
	const char* array;
	uint64_t    first, count;
	unsigned    element_size;
	uint64_t    rows, pitch;
	uint64_t    rebase;
	int         saturate;
} giddy_slice_t;

#define GIDDY_MAX_SHARD_SLICES 6
typedef struct giddy_shard_t {
	uint64_t      first_segment, num_segments;
	uint64_t      first_element, num_elements;  /* of the decompressed column; the shard's `length` */
	uint64_t      num_runs;                     /* RLE/RPE: the shard's num_element_runs */
	uint64_t      num_bytes;                    /* DZB-variable: the shard's bytes of compressed_data */
	uint64_t      bitmap_length;                /* incidence bitmaps: containers per bitmap in the shard */
	uint64_t      num_patches;                  /* SCATTER: the shard's data_length */
	unsigned      num_slices;
	giddy_slice_t slices[GIDDY_MAX_SHARD_SLICES];
} giddy_shard_t;

/* Segments [first, past_last) of part `part` of `parts` nearly equal contiguous parts of `num_segments`. */
GIDDY_API giddy_status_t giddy_split_segments(uint64_t num_segments, unsigned parts, unsigned part, uint64_t* first, uint64_t* past_last);

/* The slices of the column's arrays which segments [first_segment, past_last_segment) read, and the shard's scalar
 * arguments.  GIDDY_INVALID_ARGUMENT when the range cannot stand on its own (a DZB-variable shard must begin on a
 * size-container boundary). */
GIDDY_API giddy_status_t giddy_shard_plan(const giddy_column_layout_t* column, uint64_t first_segment, uint64_t past_last_segment,
	giddy_shard_t* out);

/* Bytes of a slice once copied (rows * count * element_size), and the copy itself, host to host, rebasing on the way:
 * `whole_array` is the column's array the slice was planned on, `destination` a buffer of the shard's own (e.g. pinned
 * staging memory from giddy_host_alloc). */
GIDDY_API size_t         giddy_slice_bytes(const giddy_slice_t* slice);
GIDDY_API giddy_status_t giddy_slice_copy(const giddy_slice_t* slice, void* destination, const void* whole_array);

/* Model: the coefficients of a shard which begins at element first_element: c0 + c1 * (first_element + i) = (c0 + c1 * first_element) + c1 * i,
 * in the arithmetic of the model's coefficient type (coefficient_size bytes, signed). */
GIDDY_API giddy_status_t giddy_shard_model_coefficients(const int64_t* coefficients, unsigned dimension, unsigned coefficient_size,
	uint64_t first_element, int64_t* out);

/* ------------------------------------------------------------------------- */
/* (6) Host memory for host-resident columns: page-locked, and placed on the   */
/* NUMA node the device hangs off, so that the GPUs of a box do not all pull   */
/* their compressed chunks out of one socket's memory.                          */
/* ------------------------------------------------------------------------- */

/* The NUMA node of a CUDA device (sysfs), or -1 when the platform does not say (single-node hosts, containers). */
GIDDY_API int giddy_host_numa_node_of_device(int device);

/* `bytes` of page-locked host memory for copies to and from `device` (< 0: the current device): mapped, bound to the
 * device's NUMA node where there is one (mbind), touched, and registered with CUDA (cudaHostRegisterPortable).
 * *numa_node (optional) receives the node the memory was bound to, or -1. */
GIDDY_API giddy_status_t giddy_host_alloc(size_t bytes, int device, void** out, int* numa_node);
GIDDY_API giddy_status_t giddy_host_free(void* memory);

/* Restrict the calling thread to the CPUs of the device's NUMA node (no-op returning -1 where there is none). */
GIDDY_API int giddy_host_bind_thread_to_device(int device);

/* ------------------------------------------------------------------------- */
/* (7) Device groups: one worker thread per GPU of the box.  The decode path    */
/* needs no collective; the one exchange of the multi-GPU path is the          */
/* broadcast of a dictionary from the device which loaded it (NCCL over         */
/* NVLink; libnccl is loaded on first use, never linked).                       */
/* ------------------------------------------------------------------------- */

typedef struct giddy_device_group giddy_device_group_t;

/* A group over `devices[0 .. num_devices)` (NULL: all visible devices).  Each device gets a worker thread which makes
 * the device current, binds itself to the device's NUMA node, and waits for work. */
GIDDY_API giddy_status_t giddy_device_group_create(const int* devices, unsigned num_devices, giddy_device_group_t** out);
GIDDY_API void           giddy_device_group_destroy(giddy_device_group_t* group);
GIDDY_API unsigned       giddy_device_group_size(const giddy_device_group_t* group);
GIDDY_API int            giddy_device_group_device(const giddy_device_group_t* group, unsigned rank);

/* work(rank, group size, device, user_data) on every worker concurrently; returns when all have returned, with the
 * first non-success status (and its message in giddy_last_error()). */
typedef giddy_status_t (*giddy_device_work_t)(unsigned rank, unsigned num_devices, int device, void* user_data);
GIDDY_API giddy_status_t giddy_device_group_run(giddy_device_group_t* group, giddy_device_work_t work, void* user_data);

/* ncclBroadcast of `bytes` bytes from buffers[root] (on the root's device) to buffers[r] on every other device of the
 * group, each on streams[r] (NULL: the devices' default streams); asynchronous on those streams.  Communicators are
 * created on first use (ncclCommInitAll).  GIDDY_UNSUPPORTED when libnccl.so.2 cannot be loaded. */
GIDDY_API giddy_status_t giddy_device_group_broadcast(giddy_device_group_t* group, void* const* buffers, size_t bytes, unsigned root,
	void* const* streams);

/* The same over a communicator of the caller's (one rank per process, e.g. under MPI): ncclBroadcast(buffer, buffer, bytes,
 * ncclUint8, root, comm, stream). */
GIDDY_API giddy_status_t giddy_broadcast_dictionary(void* nccl_communicator, void* buffer, size_t bytes, int root, void* stream);

/* ------------------------------------------------------------------------- */
/* (8) Compressors.  The reference ships decompressors only ("compression ...  */
/* is not included", README.md:128-130; its open Issue #3); these are their    */
/* inverses, on the GPU, emitting exactly the layouts the decompressors of     */
/* groups (1)-(3) read, anchor arrays included.  Registry keys: the            */
/* decompressor's with "decompression::" replaced by "compression::".          */
/* Every launcher takes an optional `status_flags` device word which has bits  */
/* OR-ed into it when the column does not fit the format (it is never cleared: */
/* zero it beforehand); the outputs are then unspecified but in bounds.        */
/* ------------------------------------------------------------------------- */

#define GIDDY_COMPRESSION_UNREPRESENTABLE 1u  /* an element does not fit: delta / offset too wide, value not in the dictionary, size field too narrow */
#define GIDDY_COMPRESSION_CAPACITY        2u  /* more runs / bytes than the caller's arrays hold (the true count is still reported) */
#define GIDDY_COMPRESSION_INDEX_OVERFLOW  4u  /* an anchor (run index, byte offset) does not fit its index type */

/* format: delta::decompress, src/kernels/decompression/delta.cuh:66-108.  compressed[i] = x[i] - x[i-1] mod 2^(8U), x[-1] = base_value;
 * anchor_values[s] = x[s * segment_length - 1] (base_value for s = 0). */
GIDDY_API const void* giddy_b200_delta_compress_kernel(unsigned index_size, unsigned uncompressed_size, unsigned compressed_size);
GIDDY_API giddy_status_t giddy_b200_delta_compress_launch(
	void* stream, const giddy_launch_config_t* config, unsigned index_size, unsigned uncompressed_size, unsigned compressed_size,
	void* compressed, void* anchor_values, const void* uncompressed, uint64_t length, uint64_t segment_length, uint64_t base_value,
	unsigned* status_flags);

/* format: discard_zero_bytes::fixed_width::decompress, discard_zero_bytes_fixed.cuh:108-124.  compressed[i] = the low bytes of x[i]. */
GIDDY_API const void* giddy_b200_discard_zero_bytes_fixed_compress_kernel(unsigned index_size, unsigned uncompressed_size, unsigned compressed_size);
GIDDY_API giddy_status_t giddy_b200_discard_zero_bytes_fixed_compress_launch(
	void* stream, const giddy_launch_config_t* config, unsigned index_size, unsigned uncompressed_size, unsigned compressed_size,
	void* compressed, const void* uncompressed, uint64_t length, unsigned* status_flags);

/* format: dictionary::decompress, dictionary.cuh:61-79.  compressed[i] = index of x[i] in dictionary_entries, which must ascend. */
GIDDY_API const void* giddy_b200_dictionary_compress_kernel(unsigned index_size, unsigned uncompressed_size, unsigned dictionary_index_size);
GIDDY_API giddy_status_t giddy_b200_dictionary_compress_launch(
	void* stream, const giddy_launch_config_t* config, unsigned index_size, unsigned uncompressed_size, unsigned dictionary_index_size,
	void* compressed, const void* uncompressed, const void* dictionary_entries, uint64_t num_dictionary_entries, uint64_t length,
	unsigned* status_flags);

/* format: frame_of_reference::decompress, frame_of_reference.cuh:35-80.  Per segment, constant: c0 = min + 2^(8C-1); linear: c1 =
 * round-half-even((last - first) / (L - 1)), c0 = first + min residual + 2^(8C-1); compressed[i] = x[i] - model(i - s * L).
 * Values are signed; coefficients int32 (1 or 2 per segment). */
GIDDY_API const void* giddy_b200_frame_of_reference_compress_kernel(unsigned index_size, unsigned uncompressed_size, unsigned compressed_size, giddy_model_t model);
GIDDY_API giddy_status_t giddy_b200_frame_of_reference_compress_launch(
	void* stream, const giddy_launch_config_t* config, unsigned index_size, unsigned uncompressed_size, unsigned compressed_size, giddy_model_t model,
	void* compressed, void* segment_model_coefficients, const void* uncompressed, uint64_t length, uint64_t segment_length,
	unsigned* status_flags);

/* format: run_length_encoding::decompress, run_length_encoding.cuh:284-374.  Runs of equal adjacent elements, cut every
 * max_run_length elements (0: the run-length type's maximum).  run_data / run_lengths hold run_capacity entries;
 * *num_element_runs (device, size_type<run_index_size>, optional) receives the number of runs. */
GIDDY_API const void* giddy_b200_run_length_encoding_compress_kernel(unsigned uncompressed_index_size, unsigned uncompressed_size, unsigned run_length_size, unsigned run_index_size);
GIDDY_API giddy_status_t giddy_b200_run_length_encoding_compress_launch(
	void* stream, const giddy_launch_config_t* config,
	unsigned uncompressed_index_size, unsigned uncompressed_size, unsigned run_length_size, unsigned run_index_size,
	void* run_data, void* run_lengths, void* position_anchors, void* intra_run_anchor_offsets, void* num_element_runs,
	const void* uncompressed, uint64_t length, uint64_t segment_length, uint64_t max_run_length, uint64_t run_capacity,
	unsigned* status_flags);

/* format: run_position_encoding::decompress, run_position_encoding.cuh:122-359.  run_start_positions are absolute, or relative to
 * the segment in which the run begins; max_run_length 0: runs are never cut. */
GIDDY_API const void* giddy_b200_run_position_encoding_compress_kernel(unsigned uncompressed_index_size, unsigned uncompressed_size, unsigned position_offset_size, unsigned run_index_size);
GIDDY_API giddy_status_t giddy_b200_run_position_encoding_compress_launch(
	void* stream, const giddy_launch_config_t* config,
	unsigned uncompressed_index_size, unsigned uncompressed_size, unsigned position_offset_size, int positions_are_relative, unsigned run_index_size,
	void* run_data, void* run_start_positions, void* position_anchors, void* num_element_runs,
	const void* uncompressed, uint64_t length, uint64_t segment_length, uint64_t max_run_length, uint64_t run_capacity,
	unsigned* status_flags);

/* format: discard_zero_bytes::variable_width::decompress, discard_zero_bytes_variable.cuh:335-439.  Element i takes
 * max(significant bytes, min_represented_element_size) bytes of compressed_data (capacity data_capacity bytes); the size fields
 * go into containers of element_sizes_container_size bytes; *num_bytes (device, uint64, optional) receives the stream's length. */
GIDDY_API const void* giddy_b200_discard_zero_bytes_variable_compress_kernel(unsigned index_size, unsigned uncompressed_size, unsigned element_sizes_container_size);
GIDDY_API giddy_status_t giddy_b200_discard_zero_bytes_variable_compress_launch(
	void* stream, const giddy_launch_config_t* config, unsigned index_size, unsigned uncompressed_size, unsigned element_sizes_container_size,
	void* compressed_data, void* packed_element_sizes, void* position_anchors, void* num_bytes,
	const void* uncompressed, uint64_t length, uint64_t segment_length, unsigned min_represented_element_size, unsigned bits_per_element_size,
	uint64_t data_capacity, unsigned* status_flags);

/* format: incidence_bitmaps::decompress, incidence_bitmaps.cuh:69-131.  Bit (i mod 32) of word i / 32 of bitmap b is set iff
 * x[i] == dictionary_entries[b] (use_dictionary) or x[i] == b; bitmaps are round_up(ceil(length / 32), alignment) words apart. */
GIDDY_API const void* giddy_b200_incidence_bitmaps_compress_kernel(unsigned uncompressed_size, int use_dictionary);
GIDDY_API giddy_status_t giddy_b200_incidence_bitmaps_compress_launch(
	void* stream, const giddy_launch_config_t* config, unsigned uncompressed_size, unsigned bitmap_alignment_in_ints, int use_dictionary,
	void* bitmaps, const void* uncompressed, const void* dictionary_entries, uint64_t length, unsigned num_bitmaps,
	unsigned* status_flags);

/* ------------------------------------------------------------------------- */
/* (9) Utility kernels which the reference has but neither wraps nor registers */
/* (SURVEY.md 8f-4).                                                           */
/* ------------------------------------------------------------------------- */

/* decompression::model::decompress_with_intervals<IndexSize, Uncompressed, UnaryModelFunction>
 * src/kernels/decompression/model.cuh:55-98.  decompressed[i] = model_s(i - s * modeling_period), s = i / modeling_period;
 * coefficients of the uncompressed (signed) type, 1 (constant) or 2 (linear) per interval.  The reference's derived parameter
 * intervals_per_block is not needed. */
GIDDY_API const void* giddy_b200_model_with_intervals_kernel(unsigned index_size, unsigned uncompressed_size, giddy_model_t model);
GIDDY_API giddy_status_t giddy_b200_model_with_intervals_launch(
	void* stream, const giddy_launch_config_t* config, unsigned index_size, unsigned uncompressed_size, giddy_model_t model,
	void* decompressed, const void* interval_model_coefficients, uint64_t modeling_period, uint64_t length);

/* gather::gather<OutputIndexSize, ElementSize, InputIndexSize>, src/kernels/data_layout/gather.cuh:151-164:
 * reordered_data[i] = data[indices[i]], i < num_indices.  The kernel behind dictionary::decompress, under its own name. */
GIDDY_API const void* giddy_b200_gather_kernel(unsigned output_index_size, unsigned element_size, unsigned input_index_size);
GIDDY_API giddy_status_t giddy_b200_gather_launch(
	void* stream, const giddy_launch_config_t* config, unsigned output_index_size, unsigned element_size, unsigned input_index_size,
	void* reordered_data, const void* data, const void* indices, uint64_t data_length, uint64_t num_indices,
	int cache_data_in_shared_memory, unsigned serialization_factor);

/* Inclusive prefix sum with ReductionOp = plus: result[i] = *initial_value + data[0] + ... + data[i] (mod 2^(8 result_size)); the
 * job of the reference's reduce_segments + scan_using_segment_reductions (src/kernels/reduction/scan.cuh:127, 412) or
 * scan_single_segment (:525), in one launch and one pass (tiles with decoupled look-back).  initial_value: a device pointer to one
 * element of the result type, or NULL for zero.  Sizes: those of delta::decompress. */
GIDDY_API const void* giddy_b200_inclusive_scan_kernel(unsigned index_size, unsigned result_size, unsigned input_size);
GIDDY_API giddy_status_t giddy_b200_inclusive_scan_launch(
	void* stream, const giddy_launch_config_t* config, unsigned index_size, unsigned result_size, unsigned input_size,
	void* result, const void* data, uint64_t length, const void* initial_value);

#ifdef __cplusplus
} /* extern "C" */
#endif

#endif /* GIDDY_B200_H_ */
