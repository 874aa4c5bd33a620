"""Per-scheme convenience calls for the GPU compressors (include/giddy_b200.h group 8; keys "compression::..."): build the
factory key from the type sizes, resolve a launch configuration and enqueue the launch, with the named, exactly typed
arguments of include/giddy/compression_wrappers.hpp.  Buffers are torch tensors on the current CUDA device or raw device
addresses, as in libgiddy_b200.schemes.

The reference has no compressors (README.md:128-130); libgiddy_b200/encode.py is the host-side (numpy) statement of the
same layouts, and the GPU compressors reproduce it byte for byte (tests/test_compressors.py).

Each call takes `status`: an optional uint32 device tensor into which GIDDY_COMPRESSION_* bits are OR-ed when the column
does not fit the format.
"""
from . import _native as N
from .schemes import _run, _check_size, _SIGNED_NAMES, _UTAG, _ITAG, model_type_name

UNREPRESENTABLE, CAPACITY, INDEX_OVERFLOW = 1, 2, 4


def _with_status(arguments, status):
    if status is not None:
        arguments["status"] = N.ptr(status, "u32")
    return arguments


def delta_key(index_size, uncompressed_size, compressed_size):
    return "compression::delta::kernel_t<%du, %du, %du>" % (index_size, uncompressed_size, compressed_size)


def delta(compressed_output, anchor_values, uncompressed, length, segment_length, *, uncompressed_size, compressed_size, base_value=0, index_size=4,
          status=None, stream=None, limits=None, library=None):
    _check_size(uncompressed, uncompressed_size, "uncompressed")
    _check_size(compressed_output, compressed_size, "compressed_output")
    index = N.uint_of_size(index_size)
    return _run(library, delta_key(index_size, uncompressed_size, compressed_size),
                {"length": N.size_t(length), "segment_length": N.size_t(segment_length)},
                _with_status({"compressed_output": N.ptr(compressed_output, _UTAG[compressed_size]), "anchor_values": N.ptr(anchor_values, _UTAG[uncompressed_size]),
                              "uncompressed": N.cptr(uncompressed, _UTAG[uncompressed_size]), "length": index(length), "segment_length": index(segment_length),
                              "base_value": N.uint_of_size(uncompressed_size)(base_value)}, status),
                stream, limits)


def discard_zero_bytes_fixed_key(index_size, uncompressed_size, compressed_size):
    return "compression::discard_zero_bytes::fixed_width::kernel_t<%du, %du, %du>" % (index_size, uncompressed_size, compressed_size)


def discard_zero_bytes_fixed(compressed_output, uncompressed, length, *, uncompressed_size, compressed_size, index_size=4,
                             status=None, stream=None, limits=None, library=None):
    return _run(library, discard_zero_bytes_fixed_key(index_size, uncompressed_size, compressed_size), {"length": N.size_t(length)},
                _with_status({"compressed_output": N.ptr(compressed_output, _UTAG[compressed_size]), "uncompressed": N.cptr(uncompressed, _UTAG[uncompressed_size]),
                              "length": N.uint_of_size(index_size)(length)}, status),
                stream, limits)


def dictionary_key(index_size, uncompressed_size, dictionary_index_size):
    return "compression::dictionary::kernel_t<%du, %du, %du>" % (index_size, uncompressed_size, dictionary_index_size)


def dictionary(compressed_output, uncompressed, dictionary_entries, num_dictionary_entries, length, *, uncompressed_size, dictionary_index_size,
               index_size=4, status=None, stream=None, limits=None, library=None):
    """dictionary_entries must ascend (as numpy.unique returns them)"""
    return _run(library, dictionary_key(index_size, uncompressed_size, dictionary_index_size),
                {"num_dictionary_entries": N.size_t(num_dictionary_entries), "data_length": N.size_t(length)},
                _with_status({"compressed_output": N.ptr(compressed_output, _UTAG[dictionary_index_size]), "uncompressed": N.cptr(uncompressed, _UTAG[uncompressed_size]),
                              "dictionary_entries": N.cptr(dictionary_entries, _UTAG[uncompressed_size]),
                              "num_dictionary_entries": N.size_type_by_index_size(dictionary_index_size)(num_dictionary_entries),
                              "length": N.size_type_by_index_size(index_size)(length)}, status),
                stream, limits)


def frame_of_reference_key(index_size, uncompressed_size, compressed_size, model="constant", model_index_size=4):
    return "compression::frame_of_reference::kernel_t<%du, %s, %s, %s >" % (
        index_size, _SIGNED_NAMES[uncompressed_size], _SIGNED_NAMES[compressed_size], model_type_name(model, model_index_size))


def frame_of_reference(compressed_output, segment_model_coefficients, uncompressed, length, segment_length, *, uncompressed_size, compressed_size,
                       model="constant", index_size=4, model_index_size=4, status=None, stream=None, limits=None, library=None):
    dim = 2 if model == "linear" else 1
    index = N.uint_of_size(index_size)
    return _run(library, frame_of_reference_key(index_size, uncompressed_size, compressed_size, model, model_index_size),
                {"length": N.size_t(length), "segment_length": N.size_t(segment_length)},
                _with_status({"compressed_output": N.ptr(compressed_output, _ITAG[compressed_size]),
                              "segment_model_coefficients": N.ptr(segment_model_coefficients, "array_i32_%d" % dim),
                              "uncompressed": N.cptr(uncompressed, _ITAG[uncompressed_size]), "length": index(length), "segment_length": index(segment_length)}, status),
                stream, limits)


def run_length_encoding_key(index_size, uncompressed_size, run_length_size, run_index_size=None):
    return "compression::run_length_encoding::kernel_t<%du, %du, %du, %du>" % (index_size, uncompressed_size, run_length_size, run_index_size or index_size)


def run_length_encoding(run_data, run_lengths, position_anchors, intra_run_anchor_offsets, num_element_runs, uncompressed, length, segment_length, run_capacity, *,
                        uncompressed_size, run_length_size, index_size=4, max_run_length=0, status=None, stream=None, limits=None, library=None):
    """num_element_runs: a one-element device tensor of size_type<index_size> which receives the number of runs"""
    size = N.size_type_by_index_size(index_size)
    u, rl, ri = _UTAG[uncompressed_size], _UTAG[run_length_size], _UTAG[index_size]
    arguments = {"run_data": N.ptr(run_data, u), "run_lengths": N.ptr(run_lengths, rl), "position_anchors": N.ptr(position_anchors, ri),
                 "intra_run_anchor_offsets": N.ptr(intra_run_anchor_offsets, rl), "uncompressed": N.cptr(uncompressed, u),
                 "length": size(length), "segment_length": size(segment_length), "max_run_length": N.size_t(max_run_length), "run_capacity": N.size_t(run_capacity)}
    if num_element_runs is not None:
        arguments["num_element_runs"] = N.ptr(num_element_runs, "u64" if index_size > 4 else "u32")
    return _run(library, run_length_encoding_key(index_size, uncompressed_size, run_length_size),
                {"length": N.size_t(length), "segment_length": N.size_t(segment_length)}, _with_status(arguments, status), stream, limits)


def run_position_encoding_key(index_size, uncompressed_size, position_offset_size, relative, run_index_size=None):
    return "compression::run_position_encoding::kernel_t<%du, %du, %du, %s, %du>" % (
        index_size, uncompressed_size, position_offset_size, "true" if relative else "false", run_index_size or index_size)


def run_position_encoding(run_data, run_start_positions, position_anchors, num_element_runs, uncompressed, length, segment_length, run_capacity, *,
                          uncompressed_size, position_offset_size, relative, index_size=4, max_run_length=0, status=None, stream=None, limits=None, library=None):
    size = N.size_type_by_index_size(index_size)
    u, po, ri = _UTAG[uncompressed_size], _UTAG[position_offset_size], _UTAG[index_size]
    arguments = {"run_data": N.ptr(run_data, u), "run_start_positions": N.ptr(run_start_positions, po), "position_anchors": N.ptr(position_anchors, ri),
                 "uncompressed": N.cptr(uncompressed, u), "length": size(length), "segment_length": size(segment_length),
                 "max_run_length": N.size_t(max_run_length), "run_capacity": N.size_t(run_capacity)}
    if num_element_runs is not None:
        arguments["num_element_runs"] = N.ptr(num_element_runs, "u64" if index_size > 4 else "u32")
    return _run(library, run_position_encoding_key(index_size, uncompressed_size, position_offset_size, relative),
                {"length": N.size_t(length), "segment_length": N.size_t(segment_length)}, _with_status(arguments, status), stream, limits)


def discard_zero_bytes_variable_key(index_size, uncompressed_size, container_size):
    return "compression::discard_zero_bytes::variable_width::kernel_t<%du, %du, %du>" % (index_size, uncompressed_size, container_size)


def discard_zero_bytes_variable(compressed_data, packed_element_sizes, position_anchors, num_bytes, uncompressed, length, segment_length,
                                min_represented_element_size, bits_per_element_size, data_capacity, *, uncompressed_size, container_size=4, index_size=4,
                                status=None, stream=None, limits=None, library=None):
    """num_bytes: a one-element int64 device tensor which receives the length of the byte stream"""
    index = N.uint_of_size(index_size)
    arguments = {"compressed_data": N.ptr(compressed_data, "u8"), "packed_element_sizes": N.ptr(packed_element_sizes, _UTAG[container_size]),
                 "position_anchors": N.ptr(position_anchors, _UTAG[index_size]), "uncompressed": N.cptr(uncompressed, _UTAG[uncompressed_size]),
                 "length": index(length), "segment_length": index(segment_length),
                 "min_represented_element_size": N.unsigned(min_represented_element_size), "bits_per_element_size": N.unsigned(bits_per_element_size),
                 "data_capacity": N.size_t(data_capacity)}
    if num_bytes is not None:
        arguments["num_bytes"] = N.ptr(num_bytes, "u64")
    return _run(library, discard_zero_bytes_variable_key(index_size, uncompressed_size, container_size),
                {"length": N.size_t(length), "segment_length": N.size_t(segment_length)}, _with_status(arguments, status), stream, limits)


def incidence_bitmaps_key(index_size, uncompressed_size, alignment_in_ints=2, use_dictionary=True):
    return "compression::incidence_bitmaps::kernel_t<%du, %du, %du, %s>" % (index_size, uncompressed_size, alignment_in_ints, "true" if use_dictionary else "false")


def incidence_bitmaps(bitmaps, uncompressed, dictionary_entries, length, num_bitmaps, *, uncompressed_size, alignment_in_ints=2, use_dictionary=True,
                      index_size=4, status=None, stream=None, limits=None, library=None):
    return _run(library, incidence_bitmaps_key(index_size, uncompressed_size, alignment_in_ints, use_dictionary), {"length": N.size_t(length)},
                _with_status({"incidence_bitmaps": N.ptr(bitmaps, "u32"), "uncompressed": N.cptr(uncompressed, _UTAG[uncompressed_size]),
                              "dictionary_entries": N.cptr(dictionary_entries, _UTAG[uncompressed_size]),
                              "length": N.size_type_by_index_size(index_size)(length), "num_bitmaps": N.u32(num_bitmaps)}, status),
                stream, limits)
