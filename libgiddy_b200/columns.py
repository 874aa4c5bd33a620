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
