"""Decoding a column in pieces: contiguous ranges of anchored segments, for the GPUs of a box or the chunks of a
host-resident pipeline.  A binding of the library's shard planner (include/giddy_b200.h group 5,
csrc/host/shard_plan.cpp): every decision -- which runs, bytes, size containers and patches a segment range reads,
what its scalar arguments are, what gets subtracted from its anchors -- is taken in the C++ host library; this module
only marshals numpy arrays in and out.

Every scheme's segments decode independently from their anchor side data, so a column splits without
any data-path communication, and planning is O(1) per shard: the carry-ins (run index + intra-run offset, byte
offset, anchor value) are already in the anchor arrays.

Two ways to place a shard's slices on a device:
  * in place  -- device arrays are as long as the whole column's and each slice is copied to its own offset, so the
    absolute indices stored in the anchors stay valid (the host-resident pipeline, where all chunks land on one GPU):
    RunLengthShard / ByteStreamShard give the ranges;
  * rebased   -- slices are copied to the start of shard-sized buffers with the slice's first index subtracted from
    the anchors (shards on different GPUs): the shard_* functions return such self-contained columns.
"""
import numpy as np

from . import _native as N
from .schemes import default_library


def split_segments(num_segments, parts, library=None):
    """Contiguous, nearly equal ranges [s0, s1) covering [0, num_segments)."""
    library = library or default_library()
    return [library.split_segments(num_segments, parts, part) for part in range(parts)]


def shard_of_rank(num_segments, rank, world_size, library=None):
    return (library or default_library()).split_segments(num_segments, world_size, rank)


def _layout(scheme, length, segment_length, **fields):
    layout = N.ColumnLayout()
    layout.scheme, layout.length, layout.segment_length = scheme, int(length), int(segment_length)
    keepalive = []
    for name, value in fields.items():
        if name in ("anchors", "run_start_positions"):
            value = np.ascontiguousarray(value)
            keepalive.append(value)
            setattr(layout, name, value.ctypes.data)
        else:
            setattr(layout, name, int(value))
    layout._keepalive = keepalive
    return layout


def plan(layout, s0, s1, library=None):
    """giddy_shard_plan: the Shard (slices + scalars) of segments [s0, s1)."""
    return (library or default_library()).shard_plan(layout, s0, s1)


def materialize(shard, arrays, library=None, out=None):
    """The shard's slices as arrays of its own: `arrays` maps the wrappers' argument names to the WHOLE column's numpy
    arrays; returns {name: numpy array} with anchors rebased.  With `out` ({name: preallocated array, e.g. views of a
    pinned staging buffer}) the slices are written there instead."""
    library = library or default_library()
    result = {}
    for k in range(shard.num_slices):
        piece = shard.slices[k]
        name = piece.name
        whole = np.ascontiguousarray(arrays[name])
        assert whole.dtype.itemsize * (whole.shape[1] if (whole.ndim > 1 and name == "segment_model_coefficients") else 1) == piece.element_size, name
        count = int(piece.rows * piece.count)
        if out is not None and name in out:
            target = out[name]
        elif whole.ndim > 1 and name == "segment_model_coefficients":
            target = np.empty((count, whole.shape[1]), dtype=whole.dtype)
        else:
            target = np.empty(count, dtype=whole.dtype)
        if count:
            library.slice_copy(piece, target, whole)
        result[name] = target
    return result


class RunLengthShard:
    """The part of an RLE / RPE column which segments [s0, s1) read, for in-place placement.  reference format: SURVEY.md A.3."""

    def __init__(self, position_anchors, s0, s1, segment_length, length, num_runs, library=None):
        anchors = np.ascontiguousarray(position_anchors)
        layout = _layout(N.SCHEME_RUN_LENGTH_ENCODING, length, segment_length, index_size=anchors.dtype.itemsize, uncompressed_size=4,
                         compressed_size=2, num_runs=num_runs, anchors=anchors)
        shard = plan(layout, s0, s1, library)
        self.s0, self.s1 = int(s0), int(s1)
        self.segment_length = int(segment_length)
        self.first_element, self.num_elements = shard.first_element, shard.num_elements
        runs = shard.slice("run_data")
        self.first_run, self.num_runs = runs.first, runs.count
        self.last_run = self.first_run + self.num_runs - 1
        self.is_last = self.s1 >= len(anchors)

    def run_slice(self):
        return slice(self.first_run, self.last_run + 1)

    def segment_slice(self):
        return slice(self.s0, self.s1)


class ByteStreamShard:
    """The part of a DZB-variable column which segments [s0, s1) read.  reference format: SURVEY.md A.8."""

    def __init__(self, position_anchors, s0, s1, segment_length, length, total_bytes, sizes_per_container, container_size=4, library=None):
        anchors = np.ascontiguousarray(position_anchors)
        layout = _layout(N.SCHEME_DISCARD_ZERO_BYTES_VARIABLE, length, segment_length, index_size=anchors.dtype.itemsize, uncompressed_size=4,
                         compressed_size=container_size, bits_per_element_size=(8 * container_size) // sizes_per_container,
                         num_bytes=total_bytes, anchors=anchors)
        self.s0, self.s1 = int(s0), int(s1)
        self.first_element = self.s0 * int(segment_length)
        # a shard can be rebased onto its own buffers only if it starts on a container boundary
        self.container_aligned = self.first_element % sizes_per_container == 0
        if not self.container_aligned:
            self.num_elements = min(self.s1 * int(segment_length), int(length)) - self.first_element
            return
        shard = plan(layout, s0, s1, library)
        self.num_elements = shard.num_elements
        data, sizes = shard.slice("compressed_data"), shard.slice("packed_element_sizes")
        self.first_byte, self.past_last_byte = data.first, data.first + data.count
        self.first_container, self.past_last_container = sizes.first, sizes.first + sizes.count


def delta_carry_in(anchor_values, s0):
    """Delta: the carry-in of a shard beginning at segment s0 is simply that segment's anchor."""
    return anchor_values[s0]


# ---------------------------------------------------------------------------------------------------
# Rebased shards: a segment range [s0, s1) of a column as a self-contained column (numpy, host side).
# Each function returns the compressed arrays of the shard and its scalar parameters.  Decoding the
# shards of a partition independently (on different GPUs) and concatenating the outputs gives the
# whole column.
# ---------------------------------------------------------------------------------------------------

def shard_position_aligned(arrays, s0, s1, segment_length, length, per_segment=(), per_element=(), library=None):
    """FOR / Delta / DZB-fixed / Dictionary: per-element arrays slice by element, per-segment ones by segment;
    everything else (e.g. the dictionary, replicated on every device) is passed through."""
    # (position-aligned schemes differ only in which arrays they have: plan them as Dictionary -- one per-element array)
    layout = _layout(N.SCHEME_DICTIONARY, length, segment_length, compressed_size=1)
    shard = plan(layout, s0, s1, library)
    first, count = shard.first_element, shard.num_elements
    result = dict(arrays)
    for name in per_element:
        result[name] = arrays[name][first: first + count]
    for name in per_segment:
        result[name] = arrays[name][s0:s1]
    return result, {"length": count, "segment_length": int(segment_length), "first_element": first}


def shard_frame_of_reference(arrays, s0, s1, segment_length, length, library=None):
    coefficients = np.ascontiguousarray(arrays["segment_model_coefficients"])
    layout = _layout(N.SCHEME_FRAME_OF_REFERENCE, length, segment_length, compressed_size=arrays["compressed_input"].dtype.itemsize,
                     model_dimension=coefficients.shape[1] if coefficients.ndim > 1 else 1)
    shard = plan(layout, s0, s1, library)
    return materialize(shard, arrays, library), {"length": shard.num_elements, "segment_length": int(segment_length), "first_element": shard.first_element}


def shard_delta(arrays, s0, s1, segment_length, length, library=None):
    layout = _layout(N.SCHEME_DELTA, length, segment_length, compressed_size=arrays["compressed_input"].dtype.itemsize,
                     uncompressed_size=arrays["anchor_values"].dtype.itemsize)
    shard = plan(layout, s0, s1, library)
    return materialize(shard, arrays, library), {"length": shard.num_elements, "segment_length": int(segment_length), "first_element": shard.first_element}


def shard_model(coefficients, s0, s1, segment_length, length, model="linear", coefficient_size=4, library=None):
    """Model: one global model; a shard starting at element e evaluates c0 + c1 * (e + i) = (c0 + c1 * e) + c1 * i."""
    library = library or default_library()
    shard = plan(_layout(N.SCHEME_MODEL, length, segment_length), s0, s1, library)
    c = library.shard_model_coefficients([int(x) for x in coefficients], coefficient_size, shard.first_element)
    return c, {"length": shard.num_elements, "first_element": shard.first_element}


def shard_run_length_encoding(arrays, s0, s1, segment_length, length, library=None):
    anchors = np.ascontiguousarray(arrays["position_anchors"])
    layout = _layout(N.SCHEME_RUN_LENGTH_ENCODING, length, segment_length, index_size=anchors.dtype.itemsize,
                     uncompressed_size=arrays["run_data"].dtype.itemsize, compressed_size=arrays["run_lengths"].dtype.itemsize,
                     num_runs=len(arrays["run_data"]), anchors=anchors)
    shard = plan(layout, s0, s1, library)
    return (materialize(shard, arrays, library),
            {"length": shard.num_elements, "segment_length": int(segment_length), "num_segments": shard.num_segments,
             "num_element_runs": shard.num_runs, "first_element": shard.first_element})


def shard_run_position_encoding(arrays, s0, s1, segment_length, length, relative, library=None):
    anchors = np.ascontiguousarray(arrays["position_anchors"])
    starts = np.ascontiguousarray(arrays["run_start_positions"])
    layout = _layout(N.SCHEME_RUN_POSITION_ENCODING, length, segment_length, index_size=anchors.dtype.itemsize,
                     uncompressed_size=arrays["run_data"].dtype.itemsize, compressed_size=starts.dtype.itemsize,
                     positions_are_relative=1 if relative else 0, num_runs=len(arrays["run_data"]), anchors=anchors,
                     run_start_positions=starts)
    shard = plan(layout, s0, s1, library)
    return (materialize(shard, arrays, library),
            {"length": shard.num_elements, "segment_length": int(segment_length), "num_anchors": shard.num_segments,
             "num_element_runs": shard.num_runs, "first_element": shard.first_element})


def shard_discard_zero_bytes_variable(arrays, s0, s1, segment_length, length, total_bytes, container_size, bits_per_element_size, library=None):
    anchors = np.ascontiguousarray(arrays["position_anchors"])
    layout = _layout(N.SCHEME_DISCARD_ZERO_BYTES_VARIABLE, length, segment_length, index_size=anchors.dtype.itemsize,
                     compressed_size=container_size, bits_per_element_size=bits_per_element_size, num_bytes=total_bytes, anchors=anchors)
    shard = plan(layout, s0, s1, library)  # raises InvalidArgument (a ValueError) unless the shard begins on a size-container boundary
    return (materialize(shard, arrays, library),
            {"length": shard.num_elements, "segment_length": int(segment_length), "first_element": shard.first_element,
             "num_bytes": shard.num_bytes})


def shard_incidence_bitmaps(bitmaps, num_bitmaps, bitmap_length, w0, w1, alignment_in_ints=2, library=None):
    """Containers [w0, w1) of every bitmap, re-laid out as a K x (w1 - w0) array (pad the row length to the kernel's
    alignment with shard_incidence_bitmaps_aligned when w1 - w0 is not a multiple of it)."""
    layout = _layout(N.SCHEME_INCIDENCE_BITMAPS, 32 * int(bitmap_length), 32, num_bitmaps=num_bitmaps, bitmap_alignment_in_ints=alignment_in_ints)
    shard = plan(layout, w0, w1, library)
    shard_length = int(shard.bitmap_length)
    compact = materialize(shard, {"incidence_bitmaps": bitmaps}, library)["incidence_bitmaps"]
    shard_aligned = -(-shard_length // alignment_in_ints) * alignment_in_ints
    if shard_aligned != shard_length:  # the kernel takes bitmaps `aligned(bitmap_length)` containers apart
        padded = np.zeros((num_bitmaps, shard_aligned), dtype=compact.dtype)
        padded[:, :shard_length] = compact.reshape(num_bitmaps, shard_length)
        compact = padded.reshape(-1)
    return compact, {"bitmap_length": shard_length, "first_element": 32 * int(w0)}


def shard_scatter(data, indices, first_element, num_elements, library=None):
    """Naive patches whose target lies in [first_element, first_element + num_elements), target indices rebased.
    Ascending indices are cut by the library's planner (two binary searches); an unsorted list is filtered here."""
    indices = np.ascontiguousarray(indices)
    if len(indices) > 1 and not (indices[1:] >= indices[:-1]).all():
        wanted = (indices >= first_element) & (indices < first_element + num_elements)
        return data[wanted], (indices[wanted].astype(np.int64) - first_element).astype(indices.dtype)
    # (for the planner a column's segment range; here any element range: segments of one element)
    layout = _layout(N.SCHEME_SCATTER, first_element + num_elements, 1, index_size=indices.dtype.itemsize,
                     uncompressed_size=data.dtype.itemsize, num_patches=len(indices), anchors=indices)
    shard = plan(layout, first_element, first_element + num_elements, library)
    pieces = materialize(shard, {"data": data, "indices": indices}, library)
    return pieces["data"], pieces["indices"]


def broadcast_dictionary(dictionary_tensor, source_rank=0):
    """The one collective of the multi-GPU path: the rank which loaded a dictionary sends it to the others
    (NCCL over NVLink on a GPU box, gloo in the CPU tests).  Decoding itself needs no communication.  One process per
    GPU goes through torch.distributed's communicator; a single process driving several GPUs uses the library's
    own (DeviceGroup.broadcast -> giddy_device_group_broadcast, ncclCommInitAll + ncclBroadcast)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(dictionary_tensor, src=source_rank)
    return dictionary_tensor
