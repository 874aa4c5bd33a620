"""The C++ shard planner (include/giddy_b200.h group 5) against a plain numpy restatement of what a segment range of
each format reads.  CPU only: planning never touches a device."""
import numpy as np
import pytest

import cases
from libgiddy_b200 import _native as N
from libgiddy_b200 import columns
from libgiddy_b200 import encode as E


def test_split_segments_matches_integer_arithmetic(library):
    for num_segments in (0, 1, 7, 1024, 1000003, (1 << 40) + 5):
        for parts in (1, 2, 3, 8):
            for part in range(parts):
                assert library.split_segments(num_segments, parts, part) == ((num_segments * part) // parts, (num_segments * (part + 1)) // parts)
    with pytest.raises(ValueError):
        library.split_segments(10, 2, 2)


@pytest.mark.parametrize("world", [2, 3, 7])
def test_run_length_shards(library, world):
    case = cases.rle_case(21, 30000 + 7, 256, 4, 2, mean_run=40)
    n, L = case.params["n"], case.params["L"]
    anchors, lengths = case.inputs["position_anchors"], case.inputs["run_lengths"].astype(np.int64)
    starts = np.cumsum(lengths) - lengths
    S = E.num_segments(n, L)
    for rank in range(world):
        s0, s1 = columns.shard_of_rank(S, rank, world)
        a, p = columns.shard_run_length_encoding(case.inputs, s0, s1, L, n)
        first, last_element = s0 * L, min(s1 * L, n) - 1
        first_run = int(np.searchsorted(starts, first, side="right")) - 1
        last_run = int(np.searchsorted(starts, last_element, side="right")) - 1
        assert int(anchors[s0]) == first_run
        # the planner may take one run more than needed (the next segment's anchor run when it begins on the boundary)
        assert p["num_element_runs"] in (last_run - first_run + 1, last_run - first_run + 2)
        assert (a["run_data"] == case.inputs["run_data"][first_run: first_run + p["num_element_runs"]]).all()
        assert (a["position_anchors"].astype(np.int64) == anchors[s0:s1].astype(np.int64) - first_run).all()
        assert (a["intra_run_anchor_offsets"] == case.inputs["intra_run_anchor_offsets"][s0:s1]).all()
        assert p["length"] == last_element - first + 1 and p["first_element"] == first


def test_variable_width_shards_and_the_container_rule(library):
    case = cases.dzbv_case(22, 20000 + 3, 1024, 4, 4, 2, 1)
    n, L = case.params["n"], case.params["L"]
    total = len(case.inputs["compressed_data"]) - 8
    anchors = case.inputs["position_anchors"].astype(np.int64)
    for (s0, s1) in ((0, 5), (5, 20), (3, 3)):
        a, p = columns.shard_discard_zero_bytes_variable(case.inputs, s0, s1, L, n, total, 4, 2)
        if s1 == s0:
            assert p["length"] == 0
            continue
        first_byte = anchors[s0]
        past_last = anchors[s1] if s1 < len(anchors) else total
        assert p["num_bytes"] == past_last - first_byte
        assert (a["compressed_data"] == case.inputs["compressed_data"][first_byte:past_last]).all()
        assert (a["position_anchors"].astype(np.int64) == anchors[s0:s1] - first_byte).all()
        assert (a["packed_element_sizes"] == case.inputs["packed_element_sizes"][s0 * L // 16: -(-min(s1 * L, n) // 16)]).all()
    odd = {"position_anchors": np.arange(0, 100, 10, dtype=np.uint32), "compressed_data": np.zeros(100, np.uint8), "packed_element_sizes": np.zeros(10, np.uint32)}
    with pytest.raises(ValueError):
        columns.shard_discard_zero_bytes_variable(odd, 1, 2, 10, 100, 100, 4, 2)


def test_absolute_run_positions_are_rebased_and_clamped(library):
    case = cases.rpe_case(23, 20000, 512, 4, 4, False, mean_run=700)
    n, L = case.params["n"], case.params["L"]
    s0, s1 = 7, 19
    a, p = columns.shard_run_position_encoding(case.inputs, s0, s1, L, n, False)
    first_run = int(case.inputs["position_anchors"][s0])
    whole = case.inputs["run_start_positions"][first_run: first_run + p["num_element_runs"]].astype(np.int64) - s0 * L
    whole[whole < 0] = 0
    assert (a["run_start_positions"].astype(np.int64) == whole).all()


def test_scatter_shards_cut_an_ascending_list(library):
    r = np.random.default_rng(5)
    indices = np.sort(r.choice(100000, 5000, replace=False)).astype(np.uint32)
    data = r.integers(0, 1 << 60, 5000, dtype=np.uint64)
    for first, count in ((0, 1000), (1000, 33333), (99000, 1000), (500, 0)):
        d, i = columns.shard_scatter(data, indices, first, count)
        wanted = (indices >= first) & (indices < first + count)
        assert (d == data[wanted]).all() and (i.astype(np.int64) == indices[wanted].astype(np.int64) - first).all()


def test_bitmap_shards_are_two_dimensional_slices(library):
    case = cases.bitmaps_case(24, 32 * 90, 4, 5)
    bitmaps = case.inputs["incidence_bitmaps"]
    shard, p = columns.shard_incidence_bitmaps(bitmaps, 5, 90, 11, 40)
    assert p == {"bitmap_length": 29, "first_element": 32 * 11}
    assert (shard.reshape(5, 30)[:, :29] == bitmaps.reshape(5, 90)[:, 11:40]).all() and (shard.reshape(5, 30)[:, 29] == 0).all()


def test_model_coefficients_absorb_the_first_position(library):
    assert library.shard_model_coefficients([-1234, 7], 4, 1000) == [-1234 + 7000, 7]
    assert library.shard_model_coefficients([5], 4, 1000) == [5]
    wrapped = library.shard_model_coefficients([2 ** 31 - 1, 3], 4, 10)  # wraps in the coefficient type (int32)
    assert wrapped[0] == ((2 ** 31 - 1 + 30 + 2 ** 31) % 2 ** 32) - 2 ** 31


def test_planner_rejects_ranges_outside_the_column(library):
    layout = columns._layout(N.SCHEME_DELTA, 1000, 100, compressed_size=4, uncompressed_size=8)
    with pytest.raises(ValueError):
        library.shard_plan(layout, 5, 11)
    with pytest.raises(ValueError):
        library.shard_plan(columns._layout(N.SCHEME_DELTA, 1000, 0, compressed_size=4, uncompressed_size=8), 0, 1)
