"""Per-scheme convenience calls over the registry API: build the factory key from the type sizes,
resolve a launch configuration and enqueue the launch, with the argument names and exact types the
reference's wrappers expect (reference src/kernel_wrappers/**/*.cu).

Buffers are torch tensors on the current CUDA device (any dtype of the right element size; the
kernels treat them as raw unsigned containers, as the reference does), or raw device addresses.
Every function works identically against the product library and against the rebuilt reference
(oracle/_ref/libgiddy_ref.so) -- pass `library=`.
"""
from . import _native as N

_default_library = None


def default_library():
    global _default_library
    if _default_library is None:
        _default_library = N.Library()
    return _default_library


_SIGNED_NAMES = {1: "signed char", 2: "short", 4: "int", 8: "long"}
_UTAG = {1: "u8", 2: "u16", 4: "u32", 8: "u64"}
_ITAG = {1: "i8", 2: "i16", 4: "i32", 8: "i64"}


def _check_size(tensor, size, name):
    if hasattr(tensor, "element_size") and tensor.element_size() != size:
        raise ValueError("%s: expected elements of %d bytes, got %d" % (name, size, tensor.element_size()))


_kernels = {}         # (library path, key) -> Kernel: wrappers are stateless, so one instance per key serves all calls
_configurations = {}  # (library path, key, resolve arguments, limits) -> LaunchConfig, per current device


def _run(library, key, resolve_arguments, launch_arguments, stream, limits, device_properties=None):
    library = library or default_library()
    kernel = _kernels.get((library.path, key))
    if kernel is None:
        kernel = _kernels[(library.path, key)] = library.produce_subclass(key)
    signature = (library.path, key, tuple(sorted((k, v.tag, v.value) for k, v in resolve_arguments.items())),
                 tuple(sorted(limits.items())) if limits else None, _current_device())
    config = _configurations.get(signature)
    if config is None:
        config = kernel.resolve_launch_configuration(resolve_arguments, limits=limits, device_properties=device_properties)
        if len(_configurations) > 4096:
            _configurations.clear()
        _configurations[signature] = config
    if isinstance(stream, N.ScanChunk):  # building a host-resident scan: record the launch instead of enqueueing it
        stream.scan.add_launch(stream.index, kernel, config, launch_arguments)
    else:
        kernel.enqueue_launch(stream, config, launch_arguments)
    return config


def _current_device():
    try:
        import torch
        return torch.cuda.current_device()
    except Exception:
        return -1


# ---- keys -----------------------------------------------------------------

def model_type_name(model, index_size, coefficient_size=4):
    return "cuda::functors::unary::parametric_model::%s<%du, %s>" % (model, index_size, _SIGNED_NAMES[coefficient_size])


def frame_of_reference_key(index_size, uncompressed_size, compressed_size, model="constant", model_index_size=4):
    return "decompression::frame_of_reference::kernel_t<%du, %s, %s, %s >" % (
        index_size, _SIGNED_NAMES[uncompressed_size], _SIGNED_NAMES[compressed_size], model_type_name(model, model_index_size))


def model_key(index_size, uncompressed_size, model="linear"):
    return "decompression::model::kernel_t<%du, %s, %s >" % (
        index_size, _SIGNED_NAMES[uncompressed_size], model_type_name(model, index_size, uncompressed_size))


def run_length_encoding_key(index_size, uncompressed_size, run_length_size, run_index_size=None):
    return "decompression::run_length_encoding::kernel_t<%du, %du, %du, %du>" % (
        index_size, uncompressed_size, run_length_size, run_index_size or index_size)


def run_position_encoding_key(index_size, uncompressed_size, position_offset_size, relative, run_index_size=None):
    return "decompression::run_position_encoding::kernel_t<%du, %du, %du, %s, %du>" % (
        index_size, uncompressed_size, position_offset_size, "true" if relative else "false", run_index_size or index_size)


def delta_key(index_size, uncompressed_size, compressed_size):
    return "decompression::delta::kernel_t<%du, %du, %du>" % (index_size, uncompressed_size, compressed_size)


def dictionary_key(index_size, uncompressed_size, dictionary_index_size):
    return "decompression::dictionary::kernel_t<%du, %du, %du>" % (index_size, uncompressed_size, dictionary_index_size)


def discard_zero_bytes_fixed_key(index_size, uncompressed_size, compressed_size):
    return "decompression::discard_zero_bytes::fixed_width::kernel_t<%du, %du, %du, (util::endianness_t)1, (unsigned short)32>" % (
        index_size, uncompressed_size, compressed_size)


def discard_zero_bytes_variable_key(index_size, uncompressed_size, container_size):
    return "decompression::discard_zero_bytes::variable_width::kernel_t<%du, %du, %du>" % (index_size, uncompressed_size, container_size)


def incidence_bitmaps_key(index_size, uncompressed_size, alignment_in_ints=2, use_dictionary=True):
    return "decompression::incidence_bitmaps::kernel_t<%du, %du, %du, %s>" % (
        index_size, uncompressed_size, alignment_in_ints, "true" if use_dictionary else "false")


def scatter_key(output_index_size, element_size, input_index_size):
    return "scatter::kernel_t<%du, %du, %du, (unsigned short)16>" % (output_index_size, element_size, input_index_size)


def compressed_indices_scatter_key(output_index_size, element_size, input_index_size, run_length_size=None):
    return "scatter::compressed_indices::kernel_t<%du, %du, %du, %du>" % (
        output_index_size, element_size, input_index_size, run_length_size or input_index_size)


# ---- launches --------------------------------------------------------------------------------------

def frame_of_reference(decompressed, compressed_input, segment_model_coefficients, length, segment_length, *,
                       uncompressed_size, compressed_size, model="constant", index_size=4, model_index_size=4,
                       stream=None, limits=None, library=None):
    """reference wrapper: src/kernel_wrappers/decompression/frame_of_reference.cu:58-102"""
    _check_size(decompressed, uncompressed_size, "decompressed")
    _check_size(compressed_input, compressed_size, "compressed_input")
    dim = 2 if model == "linear" else 1
    index = N.uint_of_size(index_size)
    return _run(library, frame_of_reference_key(index_size, uncompressed_size, compressed_size, model, model_index_size),
                {"length": N.size_t(length), "segment_length": N.size_t(segment_length)},
                {"decompressed": N.ptr(decompressed, _ITAG[uncompressed_size]),
                 "compressed_input": N.cptr(compressed_input, _ITAG[compressed_size]),
                 "segment_model_coefficients": N.cptr(segment_model_coefficients, "array_i32_%d" % dim),
                 "length": index(length), "segment_length": index(segment_length)},
                stream, limits)


def model(decompressed, model_coefficients, length, *, uncompressed_size, model="linear", index_size=4,
          stream=None, limits=None, library=None):
    """reference wrapper: src/kernel_wrappers/decompression/model.cu:44-83"""
    _check_size(decompressed, uncompressed_size, "decompressed")
    return _run(library, model_key(index_size, uncompressed_size, model),
                {"length": N.size_t(length)},
                {"decompressed": N.ptr(decompressed, _ITAG[uncompressed_size]),
                 "model_coefficients": N.coefficients(_ITAG[uncompressed_size], model_coefficients),
                 "length": N.uint_of_size(index_size)(length)},
                stream, limits)


def run_length_encoding(decompressed, run_data, run_lengths, position_anchors, intra_run_anchor_offsets,
                        segment_length, num_segments, num_element_runs, length, *,
                        uncompressed_size, run_length_size, index_size=4, stream=None, limits=None, library=None):
    """reference wrapper: src/kernel_wrappers/decompression/run_length_encoding.cu:57-95"""
    _check_size(decompressed, uncompressed_size, "decompressed")
    _check_size(run_lengths, run_length_size, "run_lengths")
    _check_size(position_anchors, index_size, "position_anchors")
    size = N.size_type_by_index_size(index_size)
    u, rl, ri = _UTAG[uncompressed_size], _UTAG[run_length_size], _UTAG[index_size]
    return _run(library, run_length_encoding_key(index_size, uncompressed_size, run_length_size),
                {"length": N.size_t(length), "segment_length": N.size_t(segment_length)},
                {"decompressed": N.ptr(decompressed, u), "run_data": N.cptr(run_data, u), "run_lengths": N.cptr(run_lengths, rl),
                 "position_anchors": N.cptr(position_anchors, ri), "intra_run_anchor_offsets": N.cptr(intra_run_anchor_offsets, rl),
                 "segment_length": size(segment_length), "num_segments": size(num_segments),
                 "num_element_runs": size(num_element_runs), "length": size(length)},
                stream, limits)


def run_position_encoding(decompressed, run_data, run_start_positions, position_anchors,
                          segment_length, num_anchors, num_element_runs, length, *,
                          uncompressed_size, position_offset_size, relative, index_size=4, stream=None, limits=None, library=None):
    """reference wrapper: src/kernel_wrappers/decompression/run_position_encoding.cu:65-105"""
    _check_size(decompressed, uncompressed_size, "decompressed")
    _check_size(run_start_positions, position_offset_size, "run_start_positions")
    u, po, ri = _UTAG[uncompressed_size], _UTAG[position_offset_size], _UTAG[index_size]
    return _run(library, run_position_encoding_key(index_size, uncompressed_size, position_offset_size, relative),
                {"length": N.size_t(length), "segment_length": N.size_t(segment_length)},
                {"decompressed": N.ptr(decompressed, u), "run_data": N.cptr(run_data, u),
                 "run_start_positions": N.cptr(run_start_positions, po), "position_anchors": N.cptr(position_anchors, ri),
                 "segment_length": N.size_type_by_index_size(position_offset_size)(segment_length),
                 "num_anchors": N.size_type_by_index_size(index_size)(num_anchors),
                 "num_element_runs": N.size_type_by_index_size(index_size)(num_element_runs),
                 "length": N.size_type_by_index_size(index_size)(length)},
                stream, limits)


def delta(decompressed, compressed_input, anchor_values, length, segment_length, *,
          uncompressed_size, compressed_size, index_size=4, stream=None, limits=None, library=None):
    """reference wrapper: src/kernel_wrappers/decompression/delta.cu:49-85"""
    _check_size(decompressed, uncompressed_size, "decompressed")
    _check_size(compressed_input, compressed_size, "compressed_input")
    index = N.uint_of_size(index_size)
    return _run(library, delta_key(index_size, uncompressed_size, compressed_size),
                {"length": N.size_t(length), "segment_length": N.size_t(segment_length)},
                {"decompressed": N.ptr(decompressed, _UTAG[uncompressed_size]),
                 "compressed_input": N.cptr(compressed_input, _UTAG[compressed_size]),
                 "anchor_values": N.cptr(anchor_values, _UTAG[uncompressed_size]),
                 "length": index(length), "segment_length": index(segment_length)},
                stream, limits)


def dictionary(decompressed, compressed_input, dictionary_entries, length, num_dictionary_entries, *,
               uncompressed_size, dictionary_index_size, index_size=4, cache_in_shared_memory=None,
               stream=None, limits=None, library=None):
    """reference wrapper: src/kernel_wrappers/decompression/dictionary.cu:51-104"""
    _check_size(decompressed, uncompressed_size, "decompressed")
    _check_size(compressed_input, dictionary_index_size, "compressed_input")
    resolve = {"num_dictionary_entries": N.size_t(num_dictionary_entries), "data_length": N.size_t(length)}
    if cache_in_shared_memory is not None:
        resolve["cache_input_data_in_shared_mem"] = N.boolean(cache_in_shared_memory)
    return _run(library, dictionary_key(index_size, uncompressed_size, dictionary_index_size), resolve,
                {"decompressed": N.ptr(decompressed, _UTAG[uncompressed_size]),
                 "compressed_input": N.cptr(compressed_input, _UTAG[dictionary_index_size]),
                 "dictionary_entries": N.cptr(dictionary_entries, _UTAG[uncompressed_size]),
                 "length": N.size_type_by_index_size(index_size)(length),
                 "num_dictionary_entries": N.size_type_by_index_size(dictionary_index_size)(num_dictionary_entries)},
                stream, limits)


def discard_zero_bytes_fixed(decompressed, compressed_input, length, *, uncompressed_size, compressed_size, index_size=4,
                             stream=None, limits=None, library=None):
    """reference wrapper: src/kernel_wrappers/decompression/discard_zero_bytes_fixed.cu:62-99"""
    _check_size(decompressed, uncompressed_size, "decompressed")
    _check_size(compressed_input, compressed_size, "compressed_input")
    return _run(library, discard_zero_bytes_fixed_key(index_size, uncompressed_size, compressed_size),
                {"length": N.size_t(length)},
                {"decompressed": N.ptr(decompressed, _UTAG[uncompressed_size]),
                 "compressed_input": N.cptr(compressed_input, _UTAG[compressed_size]),
                 "length": N.uint_of_size(index_size)(length)},
                stream, limits)


def discard_zero_bytes_variable(decompressed, compressed_data, packed_element_sizes, position_anchors,
                                segment_length, length, min_represented_element_size, bits_per_element_size, *,
                                uncompressed_size, container_size=4, index_size=4, stream=None, limits=None, library=None):
    """reference wrapper: src/kernel_wrappers/decompression/discard_zero_bytes_variable.cu:52-99"""
    _check_size(decompressed, uncompressed_size, "decompressed")
    _check_size(packed_element_sizes, container_size, "packed_element_sizes")
    _check_size(position_anchors, index_size, "position_anchors")
    index = N.uint_of_size(index_size)
    return _run(library, discard_zero_bytes_variable_key(index_size, uncompressed_size, container_size),
                {"length": N.size_t(length), "segment_length": N.size_t(segment_length)},
                {"decompressed": N.ptr(decompressed, _UTAG[uncompressed_size]),
                 "compressed_data": N.cptr(compressed_data, "u8"),
                 "packed_element_sizes": N.cptr(packed_element_sizes, _UTAG[container_size]),
                 "position_anchors": N.cptr(position_anchors, _UTAG[index_size]),
                 "segment_length": index(segment_length), "length": index(length),
                 "min_represented_element_size": N.unsigned(min_represented_element_size),
                 "bits_per_element_size": N.unsigned(bits_per_element_size)},
                stream, limits)


def delta_of_discard_zero_bytes_variable_key(index_size, uncompressed_size, compressed_size, container_size=4):
    return "decompression::composite::delta_of_discard_zero_bytes_variable_width::kernel_t<%du, %du, %du, %du>" % (
        index_size, uncompressed_size, compressed_size, container_size)


def delta_of_discard_zero_bytes_variable(decompressed, compressed_data, packed_element_sizes, position_anchors, anchor_values,
                                         segment_length, length, min_represented_element_size, bits_per_element_size, *,
                                         uncompressed_size, compressed_size, container_size=4, index_size=4,
                                         stream=None, limits=None, library=None):
    """Extension: Delta applied to the output of DZB-variable in one pass (no reference counterpart)."""
    _check_size(decompressed, uncompressed_size, "decompressed")
    _check_size(anchor_values, uncompressed_size, "anchor_values")
    _check_size(packed_element_sizes, container_size, "packed_element_sizes")
    _check_size(position_anchors, index_size, "position_anchors")
    index = N.uint_of_size(index_size)
    return _run(library, delta_of_discard_zero_bytes_variable_key(index_size, uncompressed_size, compressed_size, container_size),
                {"length": N.size_t(length), "segment_length": N.size_t(segment_length)},
                {"decompressed": N.ptr(decompressed, _UTAG[uncompressed_size]),
                 "compressed_data": N.cptr(compressed_data, "u8"),
                 "packed_element_sizes": N.cptr(packed_element_sizes, _UTAG[container_size]),
                 "position_anchors": N.cptr(position_anchors, _UTAG[index_size]),
                 "anchor_values": N.cptr(anchor_values, _UTAG[uncompressed_size]),
                 "segment_length": index(segment_length), "length": index(length),
                 "min_represented_element_size": N.unsigned(min_represented_element_size),
                 "bits_per_element_size": N.unsigned(bits_per_element_size)},
                stream, limits)


def patched_dictionary_key(index_size, uncompressed_size, dictionary_index_size):
    return "decompression::composite::patched_dictionary::kernel_t<%du, %du, %du>" % (index_size, uncompressed_size, dictionary_index_size)


def patched_dictionary(decompressed, compressed_input, dictionary_entries, length, num_dictionary_entries,
                       patch_data, patch_indices, num_patches, *,
                       uncompressed_size, dictionary_index_size, index_size=4, cache_in_shared_memory=None,
                       stream=None, limits=None, library=None):
    """Extension: Dictionary decode + naive patching (scatter) in one pass; patch_indices ascending."""
    _check_size(decompressed, uncompressed_size, "decompressed")
    _check_size(compressed_input, dictionary_index_size, "compressed_input")
    _check_size(patch_data, uncompressed_size, "patch_data")
    _check_size(patch_indices, index_size, "patch_indices")
    resolve = {"num_dictionary_entries": N.size_t(num_dictionary_entries), "data_length": N.size_t(length)}
    if cache_in_shared_memory is not None:
        resolve["cache_input_data_in_shared_mem"] = N.boolean(cache_in_shared_memory)
    size = N.size_type_by_index_size(index_size)
    return _run(library, patched_dictionary_key(index_size, uncompressed_size, dictionary_index_size), resolve,
                {"decompressed": N.ptr(decompressed, _UTAG[uncompressed_size]),
                 "compressed_input": N.cptr(compressed_input, _UTAG[dictionary_index_size]),
                 "dictionary_entries": N.cptr(dictionary_entries, _UTAG[uncompressed_size]),
                 "length": size(length),
                 "num_dictionary_entries": N.size_type_by_index_size(dictionary_index_size)(num_dictionary_entries),
                 "patch_data": N.cptr(patch_data, _UTAG[uncompressed_size]),
                 "patch_indices": N.cptr(patch_indices, _UTAG[index_size]),
                 "num_patches": size(num_patches)},
                stream, limits)


def patched_frame_of_reference_key(index_size, uncompressed_size, compressed_size, model="constant", model_index_size=4, input_index_size=None):
    return "decompression::composite::patched_frame_of_reference::kernel_t<%du, %s, %s, %s, %du>" % (
        index_size, _SIGNED_NAMES[uncompressed_size], _SIGNED_NAMES[compressed_size], model_type_name(model, model_index_size),
        input_index_size or index_size)


def patched_frame_of_reference(decompressed, compressed_input, segment_model_coefficients, length, segment_length,
                               data_to_scatter, run_lengths, run_individual_offset_sizes, run_baseline_values,
                               run_offsets_start_pos, offset_bytes, anchors, anchoring_period, num_runs, input_data_length, *,
                               uncompressed_size, compressed_size, model="constant", index_size=4, model_index_size=4,
                               input_index_size=None, stream=None, limits=None, library=None):
    """Extension: Frame-of-Reference decode + compressed-indices patching in one pass; patch positions ascending."""
    input_index_size = input_index_size or index_size
    _check_size(decompressed, uncompressed_size, "decompressed")
    _check_size(compressed_input, compressed_size, "compressed_input")
    _check_size(data_to_scatter, uncompressed_size, "data_to_scatter")
    dim = 2 if model == "linear" else 1
    index, ii = N.uint_of_size(index_size), N.uint_of_size(input_index_size)
    return _run(library, patched_frame_of_reference_key(index_size, uncompressed_size, compressed_size, model, model_index_size, input_index_size),
                {"length": N.size_t(length), "segment_length": N.size_t(segment_length)},
                {"decompressed": N.ptr(decompressed, _ITAG[uncompressed_size]),
                 "compressed_input": N.cptr(compressed_input, _ITAG[compressed_size]),
                 "segment_model_coefficients": N.cptr(segment_model_coefficients, "array_i32_%d" % dim),
                 "length": index(length), "segment_length": index(segment_length),
                 "data_to_scatter": N.cptr(data_to_scatter, _ITAG[uncompressed_size]),
                 "scatter_position_run_lengths": N.cptr(run_lengths, _UTAG[input_index_size]),
                 "scatter_position_run_individual_offset_sizes": N.cptr(run_individual_offset_sizes, "u8"),
                 "scatter_position_run_baseline_values": N.cptr(run_baseline_values, _UTAG[index_size]),
                 "scatter_position_run_offsets_start_pos": N.cptr(run_offsets_start_pos, _UTAG[input_index_size]),
                 "scatter_position_offset_bytes": N.cptr(offset_bytes, "u8"),
                 "scatter_position_anchors": N.cptr(anchors, _UTAG[input_index_size]),
                 "anchoring_period": ii(anchoring_period), "num_scatter_position_runs": ii(num_runs),
                 "input_data_length": N.size_type_by_index_size(input_index_size)(input_data_length)},
                stream, limits)


def incidence_bitmaps(decompressed, bitmaps, dictionary_entries, length, bitmap_length, num_bitmaps, *,
                      uncompressed_size, alignment_in_ints=2, use_dictionary=True, index_size=4,
                      stream=None, limits=None, library=None):
    """reference wrapper: src/kernel_wrappers/decompression/incidence_bitmaps.cu:48-84.
    `length` is the number of elements (for the launch configuration); `bitmap_length` the number of
    32-bit containers per bitmap."""
    _check_size(decompressed, uncompressed_size, "decompressed")
    return _run(library, incidence_bitmaps_key(index_size, uncompressed_size, alignment_in_ints, use_dictionary),
                {"length": N.size_t(length)},
                {"decompressed": N.ptr(decompressed, _UTAG[uncompressed_size]),
                 "incidence_bitmaps": N.cptr(bitmaps, "u32"),
                 "dictionary_entries": N.cptr(dictionary_entries, _UTAG[uncompressed_size]),
                 "bitmap_length": N.size_type_by_index_size(index_size)(bitmap_length),
                 "num_bitmaps": N.u32(num_bitmaps)},
                stream, limits)


def scatter(target, data, indices, data_length, *, element_size, output_index_size=4, input_index_size=4,
            stream=None, limits=None, library=None):
    """reference wrapper: src/kernel_wrappers/data_layout/scatter.cu:49-81"""
    _check_size(target, element_size, "target")
    _check_size(indices, output_index_size, "indices")
    return _run(library, scatter_key(output_index_size, element_size, input_index_size),
                {"data_length": N.size_t(data_length)},
                {"target": N.ptr(target, _UTAG[element_size]), "data": N.cptr(data, _UTAG[element_size]),
                 "indices": N.cptr(indices, _UTAG[output_index_size]),
                 "data_length": N.size_type_by_index_size(input_index_size)(data_length)},
                stream, limits)


def compressed_indices_scatter(target, data_to_scatter, run_lengths, run_individual_offset_sizes, run_baseline_values,
                               run_offsets_start_pos, offset_bytes, anchors, anchoring_period, num_runs, input_data_length, *,
                               element_size, output_index_size=4, input_index_size=4, stream=None, limits=None, library=None):
    """reference wrapper: src/kernel_wrappers/data_layout/compressed_indices_scatter.cu:48-104"""
    _check_size(target, element_size, "target")
    ii = N.uint_of_size(input_index_size)
    return _run(library, compressed_indices_scatter_key(output_index_size, element_size, input_index_size),
                {"input_data_length": N.size_t(input_data_length), "anchoring_period": N.size_t(anchoring_period)},
                {"target": N.ptr(target, _UTAG[element_size]),
                 "data_to_scatter": N.cptr(data_to_scatter, _UTAG[element_size]),
                 "scatter_position_run_lengths": N.cptr(run_lengths, _UTAG[input_index_size]),
                 "scatter_position_run_individual_offset_sizes": N.cptr(run_individual_offset_sizes, "u8"),
                 "scatter_position_run_baseline_values": N.cptr(run_baseline_values, _UTAG[output_index_size]),
                 "scatter_position_run_offsets_start_pos": N.cptr(run_offsets_start_pos, _UTAG[input_index_size]),
                 "scatter_position_offset_bytes": N.cptr(offset_bytes, "u8"),
                 "scatter_position_anchors": N.cptr(anchors, _UTAG[input_index_size]),
                 "anchoring_period": ii(anchoring_period), "num_scatter_position_runs": ii(num_runs),
                 "input_data_length": N.size_type_by_index_size(input_index_size)(input_data_length)},
                stream, limits)


# ---- utility kernels which the reference has but does not register (include/giddy_b200.h group 9) --------------------------

def gather_key(output_index_size, element_size, input_index_size):
    return "gather::kernel_t<%du, %du, %du>" % (output_index_size, element_size, input_index_size)


def gather(reordered_data, data, indices, data_length, num_indices, *, element_size, input_index_size, output_index_size=4,
           cache_in_shared_memory=None, stream=None, limits=None, library=None):
    """reference kernel: src/kernels/data_layout/gather.cuh:151-164.  reordered_data[i] = data[indices[i]]"""
    _check_size(reordered_data, element_size, "reordered_data")
    _check_size(indices, input_index_size, "indices")
    resolve = {"data_length": N.size_t(data_length), "num_indices": N.size_t(num_indices)}
    if cache_in_shared_memory is not None:
        resolve["cache_input_data_in_shared_mem"] = N.boolean(cache_in_shared_memory)
    return _run(library, gather_key(output_index_size, element_size, input_index_size), resolve,
                {"reordered_data": N.ptr(reordered_data, _UTAG[element_size]), "data": N.cptr(data, _UTAG[element_size]),
                 "indices": N.cptr(indices, _UTAG[input_index_size]),
                 "data_length": N.size_type_by_index_size(input_index_size)(data_length),
                 "num_indices": N.size_type_by_index_size(output_index_size)(num_indices)},
                stream, limits)


def model_with_intervals_key(index_size, uncompressed_size, model="linear"):
    return "decompression::model::intervals::kernel_t<%du, %s, %s >" % (
        index_size, _SIGNED_NAMES[uncompressed_size], model_type_name(model, index_size, uncompressed_size))


def model_with_intervals(decompressed, interval_model_coefficients, modeling_period, length, *, uncompressed_size, model="linear", index_size=4,
                         stream=None, limits=None, library=None):
    """reference kernel: src/kernels/decompression/model.cuh:55-98.  decompressed[i] = model_s(i - s * modeling_period)"""
    _check_size(decompressed, uncompressed_size, "decompressed")
    dim = 2 if model == "linear" else 1
    index = N.uint_of_size(index_size)
    return _run(library, model_with_intervals_key(index_size, uncompressed_size, model),
                {"length": N.size_t(length), "modeling_period": N.size_t(modeling_period)},
                {"decompressed": N.ptr(decompressed, _ITAG[uncompressed_size]),
                 "interval_model_coefficients": N.cptr(interval_model_coefficients, "array_%s_%d" % (_ITAG[uncompressed_size], dim)),
                 "modeling_period": index(modeling_period), "length": index(length)},
                stream, limits)


def inclusive_scan_key(index_size, result_size, input_size):
    return "reduction::scan::inclusive::plus::kernel_t<%du, %du, %du>" % (index_size, result_size, input_size)


def inclusive_scan(result, data, length, *, result_size, input_size, initial_value=None, index_size=4, stream=None, limits=None, library=None):
    """result[i] = initial + data[0] + ... + data[i]: the reference's scan kernels (src/kernels/reduction/scan.cuh:127, 412, 525) in one pass.
    initial_value: a one-element device tensor of the result type, or None for zero."""
    _check_size(result, result_size, "result")
    _check_size(data, input_size, "data")
    arguments = {"result": N.ptr(result, _UTAG[result_size]), "data": N.cptr(data, _UTAG[input_size]), "length": N.uint_of_size(index_size)(length)}
    if initial_value is not None:
        arguments["initial_value"] = N.cptr(initial_value, _UTAG[result_size])
    return _run(library, inclusive_scan_key(index_size, result_size, input_size), {"length": N.size_t(length)}, arguments, stream, limits)
