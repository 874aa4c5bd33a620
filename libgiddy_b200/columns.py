"""Host-side planning for decoding a column in pieces: contiguous ranges of anchored segments.

Every scheme's segments decode independently from their anchor side data, so a column
splits -- across the GPUs of a box, or across the chunks of a host-resident pipeline -- without any
data-path communication.  Each function here maps a segment range [s0, s1) to the slices of the
compressed arrays which that range reads and the scalar arguments of the corresponding launch.  This
is the "host-side exclusive scan of segment metadata": the carry-ins (run index + intra-run offset,
byte offset, anchor value) are already in the anchor arrays, so planning is O(1) per shard.

Two ways to place a shard's slices on a device:
  * in place  -- device arrays are as long as the whole column's and each slice is copied to its own
    offset, so the absolute indices stored in the anchors stay valid (used by the host-resident
    pipeline, where all chunks land on one GPU);
  * rebased   -- slices are copied to the start of shard-sized buffers and the anchors have the
    slice's first index subtracted (used when shards live on different GPUs).
"""
import numpy as np


def split_segments(num_segments, parts):
    """Contiguous, nearly equal ranges [s0, s1) covering [0, num_segments)."""
    bounds = [(num_segments * k) // parts for k in range(parts + 1)]
    return [(bounds[k], bounds[k + 1]) for k in range(parts)]


def shard_of_rank(num_segments, rank, world_size):
    return split_segments(num_segments, world_size)[rank]


class RunLengthShard:
    """The part of an RLE column which segments [s0, s1) read.  reference format: SURVEY.md A.3."""

    def __init__(self, position_anchors, s0, s1, segment_length, length, num_runs):
        self.s0, self.s1 = int(s0), int(s1)
        self.segment_length = int(segment_length)
        self.first_element = self.s0 * self.segment_length
        self.num_elements = min(self.s1 * self.segment_length, int(length)) - self.first_element
        self.first_run = int(position_anchors[self.s0])
        # the run holding the shard's last element is the anchor run of the next segment, or the column's last run
        self.last_run = int(position_anchors[self.s1]) if self.s1 < len(position_anchors) else int(num_runs) - 1
        self.num_runs = self.last_run - self.first_run + 1
        self.is_last = self.s1 >= len(position_anchors)

    def run_slice(self):
        return slice(self.first_run, self.last_run + 1)

    def segment_slice(self):
        return slice(self.s0, self.s1)

    def rebased_anchors(self, position_anchors):
        dtype = position_anchors.dtype
        return (position_anchors[self.s0:self.s1].astype(np.int64) - self.first_run).astype(dtype)


class ByteStreamShard:
    """The part of a DZB-variable column which segments [s0, s1) read.  reference format: SURVEY.md A.8."""

    def __init__(self, position_anchors, s0, s1, segment_length, length, total_bytes, sizes_per_container):
        self.s0, self.s1 = int(s0), int(s1)
        self.first_element = self.s0 * int(segment_length)
        self.num_elements = min(self.s1 * int(segment_length), int(length)) - self.first_element
        self.first_byte = int(position_anchors[self.s0])
        self.past_last_byte = int(position_anchors[self.s1]) if self.s1 < len(position_anchors) else int(total_bytes)
        self.first_container = self.first_element // sizes_per_container
        self.past_last_container = -(-(self.first_element + self.num_elements) // sizes_per_container)
        # a shard can be rebased onto its own buffers only if it starts on a container boundary
        self.container_aligned = self.first_element % sizes_per_container == 0


def delta_carry_in(anchor_values, s0):
    """Delta: the carry-in of a shard beginning at segment s0 is simply that segment's anchor."""
    return anchor_values[s0]


# ---------------------------------------------------------------------------------------------------
# Rebased shards: a segment range [s0, s1) of a column as a self-contained column (numpy, host side).
# Each function returns the compressed arrays of the shard -- views where the format allows, small
# rebased copies of the anchors otherwise -- and its scalar parameters.  Decoding the shards of a
# partition independently (on different GPUs) and concatenating the outputs gives the whole column.
# ---------------------------------------------------------------------------------------------------

def _element_range(s0, s1, segment_length, length):
    first = int(s0) * int(segment_length)
    return first, min(int(s1) * int(segment_length), int(length)) - first


def shard_position_aligned(arrays, s0, s1, segment_length, length, per_segment=(), per_element=()):
    """FOR / Delta / DZB-fixed / Dictionary: per-element arrays slice by element, per-segment ones by segment;
    everything else (e.g. the dictionary, replicated on every device) is passed through."""
    first, count = _element_range(s0, s1, segment_length, length)
    shard = dict(arrays)
    for name in per_element:
        shard[name] = arrays[name][first: first + count]
    for name in per_segment:
        shard[name] = arrays[name][s0:s1]
    return shard, {"length": count, "segment_length": int(segment_length), "first_element": first}


def shard_model(coefficients, s0, s1, segment_length, length, model="linear"):
    """Model: one global model; a shard starting at element e evaluates c0 + c1 * (e + i) = (c0 + c1 * e) + c1 * i."""
    first, count = _element_range(s0, s1, segment_length, length)
    c = [int(x) for x in coefficients]
    if model == "linear":
        c = [c[0] + c[1] * first, c[1]]
    return c, {"length": count, "first_element": first}


def shard_run_length_encoding(arrays, s0, s1, segment_length, length):
    anchors = arrays["position_anchors"]
    shard = RunLengthShard(anchors, s0, s1, segment_length, length, len(arrays["run_data"]))
    runs = shard.run_slice()
    return ({"run_data": arrays["run_data"][runs], "run_lengths": arrays["run_lengths"][runs],
             "position_anchors": shard.rebased_anchors(anchors),
             "intra_run_anchor_offsets": arrays["intra_run_anchor_offsets"][s0:s1]},
            {"length": shard.num_elements, "segment_length": int(segment_length), "num_segments": shard.s1 - shard.s0,
             "num_element_runs": shard.num_runs, "first_element": shard.first_element})


def shard_run_position_encoding(arrays, s0, s1, segment_length, length, relative):
    anchors = arrays["position_anchors"]
    shard = RunLengthShard(anchors, s0, s1, segment_length, length, len(arrays["run_data"]))
    if relative and not shard.is_last:
        # Relative positions cannot tell "begins at the next segment's first element" from "begins at this segment's
        # first element" (both are stored as 0), and a column's last segment takes every remaining run as its own: leave
        # out the next shard's first run when it begins exactly on the boundary.
        next_run = shard.last_run
        begins_in_this_shard = (next_run == int(anchors[shard.s1 - 1])) or (int(arrays["run_start_positions"][next_run]) != 0)
        if not begins_in_this_shard:
            shard.last_run -= 1
            shard.num_runs -= 1
    runs = shard.run_slice()
    starts = arrays["run_start_positions"][runs]
    if not relative:
        # absolute positions become relative to the shard's first element; the shard's first run may begin before
        # it -- its start is never read (the covering run of a segment begins at or before the segment), so clamp it
        rebased = starts.astype(np.int64) - shard.first_element
        rebased[rebased < 0] = 0
        starts = rebased.astype(starts.dtype)
    return ({"run_data": arrays["run_data"][runs], "run_start_positions": starts,
             "position_anchors": shard.rebased_anchors(anchors)},
            {"length": shard.num_elements, "segment_length": int(segment_length), "num_anchors": shard.s1 - shard.s0,
             "num_element_runs": shard.num_runs, "first_element": shard.first_element})


def shard_discard_zero_bytes_variable(arrays, s0, s1, segment_length, length, total_bytes, container_size, bits_per_element_size):
    sizes_per_container = (8 * container_size) // bits_per_element_size
    anchors = arrays["position_anchors"]
    shard = ByteStreamShard(anchors, s0, s1, segment_length, length, total_bytes, sizes_per_container)
    if not shard.container_aligned:
        raise ValueError("a DZB-variable shard must begin on a size-container boundary "
                         "(segment_length * s0 must be a multiple of %d)" % sizes_per_container)
    rebased = (anchors[s0:s1].astype(np.int64) - shard.first_byte).astype(anchors.dtype)
    return ({"compressed_data": arrays["compressed_data"][shard.first_byte: shard.past_last_byte],
             "packed_element_sizes": arrays["packed_element_sizes"][shard.first_container: shard.past_last_container],
             "position_anchors": rebased},
            {"length": shard.num_elements, "segment_length": int(segment_length), "first_element": shard.first_element,
             "num_bytes": shard.past_last_byte - shard.first_byte})


def shard_incidence_bitmaps(bitmaps, num_bitmaps, bitmap_length, w0, w1, alignment_in_ints=2):
    """Containers [w0, w1) of every bitmap, re-laid out as a K x aligned(w1 - w0) array."""
    aligned = -(-int(bitmap_length) // alignment_in_ints) * alignment_in_ints
    shard_length = int(w1) - int(w0)
    shard_aligned = -(-shard_length // alignment_in_ints) * alignment_in_ints
    out = np.zeros(num_bitmaps * shard_aligned, dtype=bitmaps.dtype)
    for b in range(num_bitmaps):
        out[b * shard_aligned: b * shard_aligned + shard_length] = bitmaps[b * aligned + w0: b * aligned + w1]
    return out, {"bitmap_length": shard_length, "first_element": 32 * int(w0)}


def shard_scatter(data, indices, first_element, num_elements):
    """Naive patches whose target lies in [first_element, first_element + num_elements), target indices rebased.
    Patches need not be sorted."""
    wanted = (indices >= first_element) & (indices < first_element + num_elements)
    return data[wanted], (indices[wanted].astype(np.int64) - first_element).astype(indices.dtype)


def broadcast_dictionary(dictionary_tensor, source_rank=0):
    """The one collective of the multi-GPU path: the rank which loaded a dictionary sends it to the others
    (NCCL over NVLink on a GPU box, gloo in the CPU tests).  Decoding itself needs no communication."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(dictionary_tensor, src=source_rank)
    return dictionary_tensor
