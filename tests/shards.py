"""Segment-range shards of seeded columns, decoded on the GPU: shared by tests/test_sharding_gpu.py (one process
decoding every shard; two processes decoding one shard each) and __graft_entry__.smoke().

For every scheme: the whole column comes from tests/cases.py (expected output = the oracle's), a shard [s0, s1) of it
from the library's C++ planner through libgiddy_b200/columns.py (slices copied into buffers of the shard's own, anchors
rebased), and the shard is decoded by the CUDA kernels through the registry C ABI exactly like a whole column.
Concatenating the shards' outputs must give the whole column, bit for bit."""
import numpy as np

import cases
from libgiddy_b200 import columns
from libgiddy_b200 import encode as E
from libgiddy_b200 import schemes as S


def _to_device(arrays):
    import torch
    return {k: torch.from_numpy(cases._as_torch_compatible(np.ascontiguousarray(v))).cuda() for k, v in arrays.items() if isinstance(v, np.ndarray)}


def _output(count, dtype):
    import torch
    itemsize = np.dtype(dtype).itemsize
    # poisoned, with a guard zone: a shard must write its own elements and nothing else
    return torch.full(((count + 64) * itemsize,), 0x5A, dtype=torch.uint8, device="cuda").view(cases._torch_int(itemsize))


def _finish(out, count, dtype):
    import torch
    torch.cuda.synchronize()
    result = out.cpu().numpy().view(dtype)
    assert (result[count:].view(np.uint8) == 0x5A).all(), "a shard wrote past its own output"
    return result[:count]


def rle(rank, world, library):
    case = cases.rle_case(7, 50000 + 37, 1024, 4, 2)
    n, L = case.params["n"], case.params["L"]
    s0, s1 = columns.shard_of_rank(E.num_segments(n, L), rank, world)
    a, p = columns.shard_run_length_encoding(case.inputs, s0, s1, L, n)
    t, out = _to_device(a), _output(p["length"], case.expected.dtype)
    if p["length"]:
        S.run_length_encoding(out, t["run_data"], t["run_lengths"], t["position_anchors"], t["intra_run_anchor_offsets"], L, p["num_segments"],
                              p["num_element_runs"], p["length"], uncompressed_size=4, run_length_size=2, library=library)
    return _finish(out, p["length"], case.expected.dtype), case.expected


def rle_long_runs(rank, world, library):
    case = cases.rle_case(17, 90000 + 5, 1024, 8, 1, mean_run=3000)  # runs spanning several segments -- and shard boundaries
    n, L = case.params["n"], case.params["L"]
    s0, s1 = columns.shard_of_rank(E.num_segments(n, L), rank, world)
    a, p = columns.shard_run_length_encoding(case.inputs, s0, s1, L, n)
    t, out = _to_device(a), _output(p["length"], case.expected.dtype)
    if p["length"]:
        S.run_length_encoding(out, t["run_data"], t["run_lengths"], t["position_anchors"], t["intra_run_anchor_offsets"], L, p["num_segments"],
                              p["num_element_runs"], p["length"], uncompressed_size=8, run_length_size=1, library=library)
    return _finish(out, p["length"], case.expected.dtype), case.expected


def _rpe(rank, world, library, relative, PO):
    case = cases.rpe_case(8, 40000 + 5, 512, 4, PO, relative)
    n, L = case.params["n"], case.params["L"]
    s0, s1 = columns.shard_of_rank(E.num_segments(n, L), rank, world)
    a, p = columns.shard_run_position_encoding(case.inputs, s0, s1, L, n, relative)
    t, out = _to_device(a), _output(p["length"], case.expected.dtype)
    if p["length"]:
        S.run_position_encoding(out, t["run_data"], t["run_start_positions"], t["position_anchors"], L, p["num_anchors"], p["num_element_runs"],
                                p["length"], uncompressed_size=4, position_offset_size=PO, relative=relative, library=library)
    return _finish(out, p["length"], case.expected.dtype), case.expected


def rpe_absolute(rank, world, library):
    return _rpe(rank, world, library, False, 4)


def rpe_relative(rank, world, library):
    return _rpe(rank, world, library, True, 2)


def delta(rank, world, library):
    case = cases.delta_case(9, 30000 + 11, 1024, 8, 4)
    n, L = case.params["n"], case.params["L"]
    s0, s1 = columns.shard_of_rank(E.num_segments(n, L), rank, world)
    a, p = columns.shard_delta(case.inputs, s0, s1, L, n)
    assert p["length"] == 0 or a["anchor_values"][0] == columns.delta_carry_in(case.inputs["anchor_values"], s0)
    t, out = _to_device(a), _output(p["length"], case.expected.dtype)
    if p["length"]:
        S.delta(out, t["compressed_input"], t["anchor_values"], p["length"], L, uncompressed_size=8, compressed_size=4, library=library)
    return _finish(out, p["length"], case.expected.dtype), case.expected


def delta_long_segments(rank, world, library):
    case = cases.delta_case(19, 6 * 65536 + 99, 65536, 8, 4)  # shards which take the decoupled look-back path
    n, L = case.params["n"], case.params["L"]
    s0, s1 = columns.shard_of_rank(E.num_segments(n, L), rank, world)
    a, p = columns.shard_delta(case.inputs, s0, s1, L, n)
    t, out = _to_device(a), _output(p["length"], case.expected.dtype)
    if p["length"]:
        S.delta(out, t["compressed_input"], t["anchor_values"], p["length"], L, uncompressed_size=8, compressed_size=4, library=library)
    return _finish(out, p["length"], case.expected.dtype), case.expected


def dzbv(rank, world, library):
    case = cases.dzbv_case(10, 20000 + 3, 1024, 4, 4, 2, 1)
    n, L = case.params["n"], case.params["L"]
    total_bytes = len(case.inputs["compressed_data"]) - 8
    s0, s1 = columns.shard_of_rank(E.num_segments(n, L), rank, world)
    a, p = columns.shard_discard_zero_bytes_variable(case.inputs, s0, s1, L, n, total_bytes, 4, 2)
    a["compressed_data"] = np.concatenate((a["compressed_data"], np.zeros(8, np.uint8)))  # (the kernels may read the word holding the last byte)
    t, out = _to_device(a), _output(p["length"], case.expected.dtype)
    if p["length"]:
        S.discard_zero_bytes_variable(out, t["compressed_data"], t["packed_element_sizes"], t["position_anchors"], L, p["length"], 1, 2,
                                      uncompressed_size=4, container_size=4, library=library)
    return _finish(out, p["length"], case.expected.dtype), case.expected


def dzbv_delta_fused(rank, world, library):
    """BASELINE configs[3] in small: the fused Delta-of-DZB-variable decoder on a shard (byte-offset AND delta carry-ins)."""
    case = cases.dzbv_delta_case(11, 40 * 1024 + 7, 1024, 8, 4, 4, 2, 1)
    n, L = case.params["n"], case.params["L"]
    total_bytes = len(case.inputs["compressed_data"]) - 8
    s0, s1 = columns.shard_of_rank(E.num_segments(n, L), rank, world)
    a, p = columns.shard_discard_zero_bytes_variable(case.inputs, s0, s1, L, n, total_bytes, 4, 2)
    a["compressed_data"] = np.concatenate((a["compressed_data"], np.zeros(8, np.uint8)))
    a["anchor_values"] = case.inputs["anchor_values"][s0:s1]
    t, out = _to_device(a), _output(p["length"], case.expected.dtype)
    if p["length"]:
        S.delta_of_discard_zero_bytes_variable(out, t["compressed_data"], t["packed_element_sizes"], t["position_anchors"], t["anchor_values"],
                                               L, p["length"], 1, 2, uncompressed_size=8, compressed_size=4, container_size=4, library=library)
    return _finish(out, p["length"], case.expected.dtype), case.expected


def frame_of_reference(rank, world, library):
    case = cases.for_case(11, 25000 + 20, 1000, 4, 2, "linear")
    n, L = case.params["n"], case.params["L"]
    s0, s1 = columns.shard_of_rank(E.num_segments(n, L), rank, world)
    a, p = columns.shard_frame_of_reference(case.inputs, s0, s1, L, n)
    t, out = _to_device(a), _output(p["length"], case.expected.dtype)
    if p["length"]:
        S.frame_of_reference(out, t["compressed_input"], t["segment_model_coefficients"], p["length"], L, uncompressed_size=4, compressed_size=2,
                             model="linear", library=library)
    return _finish(out, p["length"], case.expected.dtype), case.expected


def model(rank, world, library):
    from oracle import giddy_oracle as O
    coefficients, n, L = [-1234, 7], 10000 + 1, 1024
    s0, s1 = columns.shard_of_rank(E.num_segments(n, L), rank, world)
    c, p = columns.shard_model(coefficients, s0, s1, L, n)
    expected = O.model(np.array(coefficients), n, 4, "linear")
    out = _output(p["length"], expected.dtype)
    if p["length"]:
        S.model(out, c, p["length"], uncompressed_size=4, model="linear", library=library)
    return _finish(out, p["length"], expected.dtype), expected


def dictionary_with_patches(rank, world, library, dictionary_entries=None):
    """BASELINE configs[2] in small: Dictionary + naive patches.  `dictionary_entries` (a device tensor) is what the rank
    received from the broadcast; by default the case's own."""
    from oracle import giddy_oracle as O
    case = cases.dictionary_case(12, 30000 + 9, 8, 1, 256)
    patches = cases.scatter_case(13, case.params["n"], case.params["n"] // 100, 8, 4, 4)
    order = np.argsort(patches.inputs["indices"], kind="stable")
    data, indices = patches.inputs["data"][order], patches.inputs["indices"][order]
    n, L = case.params["n"], 1024
    whole = O.scatter(case.expected.copy(), data, indices)
    s0, s1 = columns.shard_of_rank(E.num_segments(n, L), rank, world)
    a, p = columns.shard_position_aligned(case.inputs, s0, s1, L, n, per_element=("compressed_input",))
    shard_data, shard_indices = columns.shard_scatter(data, indices, p["first_element"], p["length"])
    t = _to_device({"compressed_input": np.ascontiguousarray(a["compressed_input"]), "dictionary_entries": case.inputs["dictionary_entries"],
                    "data": np.ascontiguousarray(shard_data) if len(shard_data) else np.zeros(1, data.dtype),
                    "indices": np.ascontiguousarray(shard_indices) if len(shard_indices) else np.zeros(1, indices.dtype)})
    if dictionary_entries is not None:
        t["dictionary_entries"] = dictionary_entries
    out = _output(p["length"], case.expected.dtype)
    if p["length"]:
        S.dictionary(out, t["compressed_input"], t["dictionary_entries"], p["length"], 256, uncompressed_size=8, dictionary_index_size=1, library=library)
        if len(shard_data):
            S.scatter(out, t["data"], t["indices"], len(shard_data), element_size=8, output_index_size=4, input_index_size=4, library=library)
    return _finish(out, p["length"], case.expected.dtype), whole


def dictionary_with_patches_fused(rank, world, library):
    from oracle import giddy_oracle as O
    case = cases.dictionary_case(12, 30000 + 9, 8, 1, 256)
    patches = cases.scatter_case(13, case.params["n"], case.params["n"] // 100, 8, 4, 4)
    order = np.argsort(patches.inputs["indices"], kind="stable")
    data, indices = patches.inputs["data"][order], patches.inputs["indices"][order]
    n, L = case.params["n"], 1024
    whole = O.scatter(case.expected.copy(), data, indices)
    s0, s1 = columns.shard_of_rank(E.num_segments(n, L), rank, world)
    a, p = columns.shard_position_aligned(case.inputs, s0, s1, L, n, per_element=("compressed_input",))
    shard_data, shard_indices = columns.shard_scatter(data, indices, p["first_element"], p["length"])
    t = _to_device({"compressed_input": np.ascontiguousarray(a["compressed_input"]), "dictionary_entries": case.inputs["dictionary_entries"],
                    "data": np.ascontiguousarray(shard_data) if len(shard_data) else np.zeros(1, data.dtype),
                    "indices": np.ascontiguousarray(shard_indices) if len(shard_indices) else np.zeros(1, indices.dtype)})
    out = _output(p["length"], case.expected.dtype)
    if p["length"]:
        S.patched_dictionary(out, t["compressed_input"], t["dictionary_entries"], p["length"], 256, t["data"], t["indices"], len(shard_data),
                             uncompressed_size=8, dictionary_index_size=1, library=library)
    return _finish(out, p["length"], case.expected.dtype), whole


def incidence_bitmaps(rank, world, library):
    case = cases.bitmaps_case(14, 32 * 700, 4, 16)
    bitmap_length = 700
    w0, w1 = columns.shard_of_rank(bitmap_length, rank, world)
    shard_bitmaps, p = columns.shard_incidence_bitmaps(case.inputs["incidence_bitmaps"], 16, bitmap_length, w0, w1)
    count = 32 * p["bitmap_length"]
    t = _to_device({"incidence_bitmaps": shard_bitmaps, "dictionary_entries": case.inputs["dictionary_entries"]})
    out = _output(count, case.expected.dtype)
    if count:
        S.incidence_bitmaps(out, t["incidence_bitmaps"], t["dictionary_entries"], count, p["bitmap_length"], 16, uncompressed_size=4, library=library)
    return _finish(out, count, case.expected.dtype), case.expected


SCHEMES = {
    "rle": rle, "rle-long-runs": rle_long_runs, "rpe-abs": rpe_absolute, "rpe-rel": rpe_relative, "delta": delta,
    "delta-look-back": delta_long_segments, "dzbv": dzbv, "dzbv+delta-fused": dzbv_delta_fused, "for": frame_of_reference, "model": model,
    "dictionary+scatter": dictionary_with_patches, "dictionary+patches-fused": dictionary_with_patches_fused, "bitmaps": incidence_bitmaps,
}
