"""The N > 1 path on the CPU: world_size-2 `gloo` processes each take a contiguous range of anchored
segments of the same seeded column (libgiddy_b200/columns.py), decode ONLY their shard as a
self-contained, rebased column, and rank 0 checks that the concatenation equals the whole column.
No data-path collective is involved in decoding; the dictionary is broadcast once from rank 0, as on
a GPU box (NCCL there, gloo here).  The decoder standing in for the GPU here is the oracle -- what is
under test is the host-side shard planning (run-index / byte-offset / anchor carry-ins)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORLD = 2


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _decode_shards(rank, world):
    """Returns {scheme: (this rank's decoded shard, the whole column's expected output)}."""
    import cases
    from libgiddy_b200 import columns
    from libgiddy_b200 import encode as E
    from oracle import giddy_oracle as O
    results = {}

    # Run-Length Encoding: a ragged column, runs straddling the shard boundary
    case = cases.rle_case(7, 50000 + 37, 1024, 4, 2)
    n, L = case.params["n"], case.params["L"]
    s0, s1 = columns.shard_of_rank(E.num_segments(n, L), rank, world)
    a, p = columns.shard_run_length_encoding(case.inputs, s0, s1, L, n)
    results["rle"] = (O.run_length_encoding(np.ascontiguousarray(a["run_data"]), np.ascontiguousarray(a["run_lengths"]), a["position_anchors"],
                                            np.ascontiguousarray(a["intra_run_anchor_offsets"]), L, p["length"]), case.expected)

    # Run-Position Encoding, absolute and relative positions
    for relative, PO in ((False, 4), (True, 2)):
        case = cases.rpe_case(8, 40000 + 5, 512, 4, PO, relative)
        n, L = case.params["n"], case.params["L"]
        s0, s1 = columns.shard_of_rank(E.num_segments(n, L), rank, world)
        a, p = columns.shard_run_position_encoding(case.inputs, s0, s1, L, n, relative)
        results["rpe-rel" if relative else "rpe-abs"] = (
            O.run_position_encoding(np.ascontiguousarray(a["run_data"]), np.ascontiguousarray(a["run_start_positions"]),
                                    a["position_anchors"], L, p["length"], relative), case.expected)

    # Delta: the carry-in of a shard is its first segment's anchor
    case = cases.delta_case(9, 30000 + 11, 1024, 8, 4)
    n, L = case.params["n"], case.params["L"]
    s0, s1 = columns.shard_of_rank(E.num_segments(n, L), rank, world)
    a, p = columns.shard_position_aligned(case.inputs, s0, s1, L, n, per_segment=("anchor_values",), per_element=("compressed_input",))
    assert a["anchor_values"][0] == columns.delta_carry_in(case.inputs["anchor_values"], s0)
    results["delta"] = (O.delta(np.ascontiguousarray(a["compressed_input"]), np.ascontiguousarray(a["anchor_values"]), p["length"], L), case.expected)

    # Discard-Zero-Bytes, variable width: byte-offset and size-container carry-ins
    case = cases.dzbv_case(10, 20000 + 3, 1024, 4, 4, 2, 1)
    n, L = case.params["n"], case.params["L"]
    total_bytes = len(case.inputs["compressed_data"]) - 8
    s0, s1 = columns.shard_of_rank(E.num_segments(n, L), rank, world)
    a, p = columns.shard_discard_zero_bytes_variable(case.inputs, s0, s1, L, n, total_bytes, 4, 2)
    data = np.concatenate((a["compressed_data"], np.zeros(8, np.uint8)))
    results["dzbv"] = (O.discard_zero_bytes_variable(data, np.ascontiguousarray(a["packed_element_sizes"]), a["position_anchors"], L,
                                                     p["length"], 1, 2, 4), case.expected)

    # Frame of Reference, linear model (position inside the segment restarts at every segment: nothing to rebase)
    case = cases.for_case(11, 25000 + 20, 1000, 4, 2, "linear")
    n, L = case.params["n"], case.params["L"]
    s0, s1 = columns.shard_of_rank(E.num_segments(n, L), rank, world)
    a, p = columns.shard_position_aligned(case.inputs, s0, s1, L, n, per_segment=("segment_model_coefficients",), per_element=("compressed_input",))
    results["for"] = (O.frame_of_reference(np.ascontiguousarray(a["compressed_input"]), np.ascontiguousarray(a["segment_model_coefficients"]),
                                           p["length"], L, 4, "linear"), case.expected)

    # Model: the shard's constant coefficient absorbs the shard's first position
    coefficients = [-1234, 7]
    n, L = 10000 + 1, 1024
    s0, s1 = columns.shard_of_rank(E.num_segments(n, L), rank, world)
    c, p = columns.shard_model(coefficients, s0, s1, L, n)
    results["model"] = (O.model(np.array(c), p["length"], 4, "linear"), O.model(np.array(coefficients), n, 4, "linear"))

    # Dictionary + naive patches (BASELINE configs[2]): rank 0 owns the dictionary and broadcasts it
    case = cases.dictionary_case(12, 30000 + 9, 8, 1, 256)
    n, L = case.params["n"], 1024
    dictionary = torch.from_numpy(case.inputs["dictionary_entries"].view(np.int64).copy())
    if rank != 0:
        dictionary.zero_()
    columns.broadcast_dictionary(dictionary, 0)
    entries = dictionary.numpy().view(np.uint64)
    s0, s1 = columns.shard_of_rank(E.num_segments(n, L), rank, world)
    a, p = columns.shard_position_aligned(case.inputs, s0, s1, L, n, per_element=("compressed_input",))
    decoded = O.dictionary(np.ascontiguousarray(a["compressed_input"]), entries, p["length"])
    patches = cases.scatter_case(13, n, n // 100, 8, 4, 4)
    data, indices = columns.shard_scatter(patches.inputs["data"], patches.inputs["indices"], p["first_element"], p["length"])
    O.scatter(decoded, np.ascontiguousarray(data), np.ascontiguousarray(indices))
    whole = O.scatter(case.expected.copy(), patches.inputs["data"], patches.inputs["indices"])
    results["dictionary+scatter"] = (decoded, whole)

    # Incidence bitmaps: shards are ranges of 32-element containers
    case = cases.bitmaps_case(14, 32 * 700, 4, 16)
    bitmap_length = 700
    w0, w1 = columns.shard_of_rank(bitmap_length, rank, world)
    shard_bitmaps, p = columns.shard_incidence_bitmaps(case.inputs["incidence_bitmaps"], 16, bitmap_length, w0, w1)
    results["bitmaps"] = (O.incidence_bitmaps(shard_bitmaps, case.inputs["dictionary_entries"], p["bitmap_length"], 16, 4), case.expected)
    return results


def _worker(rank, world, port, failures):
    for path in (ROOT, os.path.join(ROOT, "tests")):
        if path not in sys.path:
            sys.path.insert(0, path)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        results = _decode_shards(rank, world)
        for name in sorted(results):
            mine, expected = results[name]
            gathered = [None] * world
            dist.all_gather_object(gathered, np.ascontiguousarray(mine).view(np.uint8).tobytes())
            if rank == 0:
                whole = np.frombuffer(b"".join(gathered), dtype=np.uint8)
                if not np.array_equal(whole, np.ascontiguousarray(expected).view(np.uint8)):
                    failures.put(name)
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_ranks_decode_disjoint_shards_of_every_scheme():
    context = mp.get_context("spawn")
    failures = context.Queue()
    port = _free_port()
    processes = [context.Process(target=_worker, args=(rank, WORLD, port, failures)) for rank in range(WORLD)]
    for p in processes:
        p.start()
    for p in processes:
        p.join(300)
        assert p.exitcode == 0, "a rank failed (exit code %s)" % p.exitcode
    failed = []
    while not failures.empty():
        failed.append(failures.get())
    assert not failed, "sharded decode differs from the whole column for: %s" % failed


def test_split_segments_is_a_partition():
    from libgiddy_b200 import columns
    for num_segments in (0, 1, 7, 1024, 1000003):
        for parts in (1, 2, 3, 8):
            ranges = columns.split_segments(num_segments, parts)
            assert ranges[0][0] == 0 and ranges[-1][1] == num_segments
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
            sizes = [b - a for a, b in ranges]
            assert max(sizes) - min(sizes) <= 1


def test_variable_width_shard_must_start_on_a_container_boundary():
    from libgiddy_b200 import columns
    anchors = np.arange(0, 100, 10, dtype=np.uint32)
    with pytest.raises(ValueError):
        columns.shard_discard_zero_bytes_variable({"position_anchors": anchors, "compressed_data": np.zeros(100, np.uint8),
                                                   "packed_element_sizes": np.zeros(10, np.uint32)}, 1, 2, 10, 100, 100, 4, 2)
