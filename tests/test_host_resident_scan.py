"""Host-resident scans (include/giddy_b200.h group 4; libgiddy_b200.Scan): compressed columns in pinned host memory,
copied chunk by chunk on side streams and decoded as they land.  The decoded column must equal the oracle's, whether
the plan is re-enqueued or replayed as a captured CUDA graph, and however chunks depend on each other."""
import os
import subprocess

import numpy as np
import pytest

import cases
import libgiddy_b200 as g
from libgiddy_b200 import schemes as S
from libgiddy_b200 import encode as E
from oracle import giddy_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_scan_entry_points_are_exported(library):
    for name in ("create", "destroy", "add_copy", "add_copy_back", "add_launch", "add_callback", "add_dependency", "set_graph_replay", "run",
                 "num_streams", "num_chunks", "num_operations"):
        assert hasattr(library.lib, "giddy_scan_" + name)


@pytest.mark.gpu
@pytest.mark.parametrize("graph", [False, True], ids=["enqueued", "graph"])
@pytest.mark.parametrize("num_streams,num_chunks", [(1, 1), (3, 8), (4, 5)])
def test_frame_of_reference_column_from_pinned_host_memory(library, graph, num_streams, num_chunks):
    import torch
    n, L = 40 * 1024 + 512, 1024   # ragged last segment
    case = cases.for_case(11, n, L, 4, 2, "linear")
    device = torch.device("cuda")
    host = {k: torch.from_numpy(cases._as_torch_compatible(v)).pin_memory() for k, v in case.inputs.items()}
    dev = {k: torch.zeros_like(v, device=device) for k, v in host.items()}
    out = torch.full((n + 64,), -1, dtype=torch.int32, device=device)
    num_segments = (n + L - 1) // L
    scan = library.create_scan(num_streams)
    bounds = [num_segments * c // num_chunks for c in range(num_chunks + 1)]
    for c in range(num_chunks):
        s0, s1 = bounds[c], bounds[c + 1]
        e0, e1 = s0 * L, min(s1 * L, n)
        chunk = scan.chunk(c)
        chunk.copy(dev["compressed_input"][e0:e1], host["compressed_input"][e0:e1])
        coefficients = dev["segment_model_coefficients"].view(-1)
        host_coefficients = host["segment_model_coefficients"].view(-1)
        chunk.copy(coefficients[2 * s0:2 * s1], host_coefficients[2 * s0:2 * s1])
        S.frame_of_reference(out[e0:e1], dev["compressed_input"][e0:e1], coefficients[2 * s0:2 * s1], e1 - e0, L,
                             uncompressed_size=4, compressed_size=2, model="linear", stream=chunk, library=library)
    assert scan.num_chunks == num_chunks and scan.num_operations == 3 * num_chunks
    scan.set_graph_replay(graph)
    for _ in range(3):
        out.fill_(-1)
        for t in dev.values():
            t.zero_()
        scan.run()
        torch.cuda.synchronize()
        got = out[:n].cpu().numpy().view(case.expected.dtype)
        assert (got == case.expected).all()
    assert int(out[n]) == -1  # nothing past the column's end


@pytest.mark.gpu
def test_dependencies_and_callbacks_between_chunks(library):
    """Chunk c's patches land in chunk c-1's output: a dictionary column decoded per chunk, then patched by a scatter
    recorded in the NEXT chunk, which therefore depends on the previous one; a callback clears a flag buffer."""
    import torch
    n, chunk_elements, num_chunks = 1 << 16, 1 << 14, 4
    rng = np.random.default_rng(5)
    entries = rng.integers(0, 1 << 62, 256, dtype=np.uint64)
    indices = rng.integers(0, 256, n, dtype=np.uint64).astype(np.uint8)
    positions = np.sort(rng.choice(n - chunk_elements, 700, replace=False)).astype(np.uint32)  # none in the last chunk
    values = rng.integers(0, 1 << 62, len(positions), dtype=np.uint64)
    expected = O.dictionary(indices, entries, n)
    expected[positions] = values
    device = torch.device("cuda")
    as_t = lambda a: torch.from_numpy(cases._as_torch_compatible(a))
    host_indices, host_values, host_positions = as_t(indices).pin_memory(), as_t(values).pin_memory(), as_t(positions).pin_memory()
    dev_indices, dev_values, dev_positions = (torch.zeros_like(t, device=device) for t in (host_indices, host_values, host_positions))
    dev_entries = as_t(entries).to(device)
    out = torch.zeros(n, dtype=torch.int64, device=device)
    flag = torch.ones(4, dtype=torch.int32, device=device)
    calls = []
    scan = library.create_scan(3)
    for c in range(num_chunks):
        e0, e1 = c * chunk_elements, (c + 1) * chunk_elements
        chunk = scan.chunk(c)
        chunk.copy(dev_indices[e0:e1], host_indices[e0:e1])
        S.dictionary(out[e0:e1], dev_indices[e0:e1], dev_entries, e1 - e0, 256, uncompressed_size=8, dictionary_index_size=1,
                     stream=chunk, library=library)
        if c > 0:  # the patches whose targets lie in the previous chunk
            lo, hi = np.searchsorted(positions, [e0 - chunk_elements, e0])
            if hi > lo:
                chunk.after(c - 1)
                chunk.copy(dev_values[lo:hi], host_values[lo:hi])
                chunk.copy(dev_positions[lo:hi], host_positions[lo:hi])
                S.scatter(out, dev_values[lo:hi], dev_positions[lo:hi], int(hi - lo), element_size=8, output_index_size=4,
                          input_index_size=4, stream=chunk, library=library)
    scan.add_callback(0, lambda stream: calls.append(stream))
    with pytest.raises(g.InvalidArgument):
        scan.add_dependency(1, 2)  # only on earlier chunks
    for graph in (False, True):
        scan.set_graph_replay(graph)
        out.zero_()
        scan.run()
        torch.cuda.synchronize()
        assert (out.cpu().numpy().view(np.uint64) == expected).all()
    assert len(calls) == 2  # once per enqueue-replay, once while the graph was captured
    del flag


def test_host_resident_example_builds(library):
    subprocess.run(["make", "-C", ROOT, "examples"], check=True, capture_output=True)
    assert os.path.exists(os.path.join(ROOT, "examples", "bin", "host_resident_scan"))


@pytest.mark.gpu
def test_host_resident_example_reproduces_the_known_answer(library):
    binary = os.path.join(ROOT, "examples", "bin", "host_resident_scan")
    if not os.path.exists(binary):
        subprocess.run(["make", "-C", ROOT, "examples"], check=True, capture_output=True)
    result = subprocess.run([binary], capture_output=True, text=True)
    assert result.returncode == 0, result.stdout + result.stderr
    assert result.stdout.count("SUCCESS") == 3, result.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("graph", [False, True], ids=["enqueued", "graph"])
def test_scan_with_kernels_which_take_a_workspace(library, graph):
    """The long-segment Delta path (cooperative launch, look-back workspace) and a patched decoder (tile index workspace)
    inside a scan: both allocate stream-ordered scratch memory from the library's pool, which must also work while the
    scan is being captured into a CUDA graph."""
    import torch
    device = torch.device("cuda")
    delta = cases.delta_case(21, 3 << 16, 1 << 16, 8, 4)              # 3 long segments: the look-back kernel
    patched = cases.patched_dictionary_case(22, 1 << 17, 8, 1, 256, 1300)
    scan = library.create_scan(2)
    outs, expected = [], []
    for index, case in enumerate((delta, patched)):
        host = {k: torch.from_numpy(cases._as_torch_compatible(v)).pin_memory() for k, v in case.inputs.items()}
        dev = {k: torch.zeros_like(v, device=device) for k, v in host.items()}
        out = torch.zeros(case.output_elements + 64, dtype=cases._torch_int(case.expected.dtype.itemsize), device=device)
        chunk = scan.chunk(index)
        for k in host:
            chunk.copy(dev[k], host[k])
        outs.append((out, dev, host))
        expected.append(case.expected)
    # record the launches through the scheme helpers (stream = the chunk)
    out, dev, _ = outs[0]
    S.delta(out, dev["compressed_input"], dev["anchor_values"], 3 << 16, 1 << 16, uncompressed_size=8,
            compressed_size=4, stream=scan.chunk(0), library=library)
    out, dev, _ = outs[1]
    S.patched_dictionary(out, dev["compressed_input"], dev["dictionary_entries"], 1 << 17, 256, dev["patch_data"], dev["patch_indices"],
                         1300, uncompressed_size=8, dictionary_index_size=1, stream=scan.chunk(1), library=library)
    scan.set_graph_replay(graph)
    for _ in range(2):
        for out, dev, _host in outs:
            out.zero_()
        scan.run()
        torch.cuda.synchronize()
        for (out, _dev, _host), want in zip(outs, expected):
            assert (out[:len(want)].cpu().numpy().view(want.dtype) == want).all()
