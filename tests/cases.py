"""Seeded test cases for every scheme: compressed inputs (numpy), parameters, the expected output
(from the oracle) and a launcher which decodes the case on the GPU through the registry C ABI.

The same cases drive
  * tests/test_gpu_parity.py      -- product library vs the oracle, on a B200;
  * tests/golden/make_ref_golden.py -- the rebuilt reference kernels (oracle/_ref), whose outputs'
    digests are committed as tests/golden/ref_gpu_digests.json;
  * tests/test_oracle_golden.py   -- the oracle vs those committed digests, on the CPU.

`ref_ok` marks the cases on which the reference's kernels are well defined;
the others are covered by the oracle alone.
"""
import hashlib

import numpy as np

from libgiddy_b200 import encode as E
from libgiddy_b200 import schemes as S
from oracle import giddy_oracle as O

UINT, INT = E.UINT, E.INT


class Case:
    def __init__(self, name, scheme, inputs, params, expected, launch, ref_ok=True, output_elements=None, initial_output=None):
        self.name, self.scheme, self.inputs, self.params = name, scheme, inputs, params
        self.expected, self.launch, self.ref_ok = expected, launch, ref_ok
        self.output_elements = len(expected) if output_elements is None else output_elements
        self.initial_output = initial_output  # for the scatters: the target's contents before patching

    def __repr__(self):
        return "Case(%s)" % self.name


def digest(array):
    return hashlib.sha256(np.ascontiguousarray(array).tobytes()).hexdigest()


def rng(*seed):
    return np.random.default_rng([int(s) for s in seed])


def random_uint(r, size, n, bits=None):
    bits = 8 * size if bits is None else bits
    if bits >= 64:
        return r.integers(0, 1 << 63, n, dtype=np.uint64) * np.uint64(2) + r.integers(0, 2, n, dtype=np.uint64)
    return r.integers(0, 1 << bits, n, dtype=np.uint64).astype(UINT[size])


MISALIGNMENT = 65  # elements: a buffer starting this far into an allocation is element-aligned but never 16-byte aligned


def _misaligned(tensor):
    """A copy of the tensor which starts MISALIGNMENT elements into its allocation."""
    import torch
    holder = torch.zeros(tensor.numel() + MISALIGNMENT, dtype=tensor.dtype, device=tensor.device)
    view = holder[MISALIGNMENT:]
    view.copy_(tensor.reshape(-1))
    return view.reshape(tensor.shape)


def run_case_on_gpu(case, library=None, limits=None, misaligned=False):
    """Copy the case's inputs to the current CUDA device, decode, and return the output as numpy.
    misaligned: every buffer starts at an address which is a multiple of its element size only (the kernels'
    vector paths want 16-byte alignment; without it their element-wise paths must produce the same output)."""
    import torch
    device = torch.device("cuda")
    tensors = {k: torch.from_numpy(_as_torch_compatible(v)).to(device) for k, v in case.inputs.items()}
    itemsize = case.expected.dtype.itemsize
    if case.initial_output is not None:
        out = torch.from_numpy(_as_torch_compatible(case.initial_output)).to(device)
    else:
        # poison the output (plus a guard zone) so that unwritten or over-written elements show up
        out = torch.full(((case.output_elements + 64) * itemsize,), 0x5A, dtype=torch.uint8, device=device).view(_torch_int(itemsize))
    if misaligned:
        # (the compressed-indices offset bytes hold 16/32/64-bit items at offsets aligned relative to the array's start: the
        # format itself wants that array aligned, in the reference as here)
        tensors = {k: (v if k == "scatter_position_offset_bytes" else _misaligned(v)) for k, v in tensors.items()}
        out = _misaligned(out)
        assert out.data_ptr() % 16 != 0
    case.launch(out, tensors, library, limits)
    torch.cuda.synchronize()
    result = out.cpu().numpy().view(case.expected.dtype)
    if case.initial_output is None:
        guard = result[case.output_elements:]
        assert (guard.view(np.uint8) == 0x5A).all(), "%s: wrote past the end of the output" % case.name
        result = result[:case.output_elements]
    return result


def _torch_int(itemsize):
    import torch
    return {1: torch.uint8, 2: torch.int16, 4: torch.int32, 8: torch.int64}[itemsize]


def _as_torch_compatible(a):
    a = np.ascontiguousarray(a)
    if a.dtype.kind == "u" and a.dtype.itemsize > 1:
        return a.view(INT[a.dtype.itemsize])
    return a


# ---- per-scheme case builders --------------------------------------------------------------------

def for_case(seed, n, L, U, C, model, I=4, MI=4):
    r = rng(1, seed, n, L, U, C)
    num_segments = E.num_segments(n, L)
    dim = 2 if model == "linear" else 1
    coefficients = r.integers(-(1 << 30), 1 << 30, (num_segments, dim), dtype=np.int64).astype(np.int32)
    if model == "linear":
        coefficients[:, 1] = r.integers(-1000, 1000, num_segments)
    compressed = r.integers(-(1 << (8 * C - 1)), 1 << (8 * C - 1), n, dtype=np.int64).astype(INT[C])
    expected = O.frame_of_reference(compressed, coefficients, n, L, U, model, MI)

    def launch(out, t, library, limits):
        S.frame_of_reference(out, t["compressed_input"], t["segment_model_coefficients"], n, L, uncompressed_size=U, compressed_size=C,
                             model=model, index_size=I, model_index_size=MI, library=library, limits=limits)
    # reference: L must be >= the block's thread count for its one-segment-per-block path, or divide it
    return Case("for-%s-I%d-U%d-C%d-MI%d-n%d-L%d" % (model, I, U, C, MI, n, L), "frame_of_reference",
                {"compressed_input": compressed, "segment_model_coefficients": coefficients}, dict(n=n, L=L), expected, launch,
                ref_ok=(n % L == 0 and (L >= 1024 or 1024 % L == 0)))


def model_case(seed, n, U, model, I=4):
    r = rng(2, seed, n, U)
    dim = 2 if model == "linear" else 1
    coefficients = r.integers(-(1 << (8 * U - 2)), 1 << (8 * U - 2), dim, dtype=np.int64)
    if model == "linear":
        coefficients[1] = r.integers(-300, 300)
    expected = O.model(coefficients, n, U, model, I)

    def launch(out, t, library, limits):
        S.model(out, [int(c) for c in coefficients], n, uncompressed_size=U, model=model, index_size=I, library=library, limits=limits)
    return Case("model-%s-I%d-U%d-n%d" % (model, I, U, n), "model", {}, dict(n=n), expected, launch)


def _random_runs_column(r, n, U, mean_run, value_bits=None):
    """A column of runs with lengths uniform in [1, 2*mean_run - 1]."""
    lengths = r.integers(1, 2 * mean_run, max(2 * n // mean_run + 16, 16))
    lengths = lengths[: np.searchsorted(np.cumsum(lengths), n) + 1]
    values = random_uint(r, U, len(lengths), value_bits)
    # make sure adjacent runs differ so the encoder reproduces them (flip the low bit where they do not)
    same = np.flatnonzero(values[1:] == values[:-1]) + 1
    values[same] ^= UINT[U](1)
    same = np.flatnonzero(values[1:] == values[:-1]) + 1
    values[same] ^= UINT[U](2)
    return np.repeat(values, lengths)[:n]


def rle_case(seed, n, L, U, RL, I=4, mean_run=16, max_run_length=None):
    r = rng(3, seed, n, L, U, RL)
    column = _random_runs_column(r, n, U, mean_run)
    enc = E.run_length_encode(column, L, RL, I, max_run_length)
    expected = O.run_length_encoding(enc["run_data"], enc["run_lengths"], enc["position_anchors"], enc["intra_run_anchor_offsets"], L, n)
    assert (expected == column).all()
    num_runs = len(enc["run_data"])

    def launch(out, t, library, limits):
        S.run_length_encoding(out, t["run_data"], t["run_lengths"], t["position_anchors"], t["intra_run_anchor_offsets"],
                              L, E.num_segments(n, L), num_runs, n, uncompressed_size=U, run_length_size=RL, index_size=I,
                              library=library, limits=limits)
    per_write = max(1, 4 // U)
    return Case("rle-I%d-U%d-RL%d-n%d-L%d-run%d" % (I, U, RL, n, L, mean_run), "run_length_encoding", enc, dict(n=n, L=L), expected, launch,
                ref_ok=(L * U >= 4 and L % per_write == 0 and n % per_write == 0))


def rpe_case(seed, n, L, U, PO, relative, I=4, mean_run=16, avoid_boundary_runs=False):
    r = rng(4, seed, n, L, U, PO)
    column = _random_runs_column(r, n, U, mean_run)
    if avoid_boundary_runs:  # make no run begin exactly on a segment boundary (reference quirk E4)
        boundaries = np.arange(L, n, L)
        column[boundaries] = column[boundaries - 1]
    enc = E.run_position_encode(column, L, PO, relative, I)
    expected = O.run_position_encoding(enc["run_data"], enc["run_start_positions"], enc["position_anchors"], L, n, relative)
    assert (expected == column).all()
    num_runs = len(enc["run_data"])

    def launch(out, t, library, limits):
        S.run_position_encoding(out, t["run_data"], t["run_start_positions"], t["position_anchors"], L, E.num_segments(n, L), num_runs, n,
                                uncompressed_size=U, position_offset_size=PO, relative=relative, index_size=I, library=library, limits=limits)
    per_write = max(1, 4 // U)
    # reference: E14 -- all lanes of its 64-thread blocks must run the same number of iterations
    return Case("rpe-%s-I%d-U%d-PO%d-n%d-L%d-run%d" % ("rel" if relative else "abs", I, U, PO, n, L, mean_run), "run_position_encoding",
                enc, dict(n=n, L=L), expected, launch,
                ref_ok=(L % (64 * per_write) == 0 and n % L == 0 and (avoid_boundary_runs or not relative)))


def delta_case(seed, n, L, U, C, I=4):
    r = rng(5, seed, n, L, U, C)
    deltas = random_uint(r, C, n)
    base = int(random_uint(r, U, 1)[0])
    column = (np.cumsum(deltas.astype(np.uint64), dtype=np.uint64) + np.uint64(base)).astype(UINT[U])
    enc = E.delta_encode(column, L, C, base)
    expected = O.delta(enc["compressed_input"], enc["anchor_values"], n, L)
    assert (expected == column).all()

    def launch(out, t, library, limits):
        S.delta(out, t["compressed_input"], t["anchor_values"], n, L, uncompressed_size=U, compressed_size=C, index_size=I,
                library=library, limits=limits)
    return Case("delta-I%d-U%d-C%d-n%d-L%d" % (I, U, C, n, L), "delta", enc, dict(n=n, L=L), expected, launch,
                ref_ok=(L >= 1024 and L % 1024 == 0))


def dictionary_case(seed, n, U, DI, num_entries, I=4, cache=None):
    r = rng(6, seed, n, U, DI, num_entries)
    entries = random_uint(r, U, num_entries)
    indices = r.integers(0, num_entries, n, dtype=np.uint64).astype(UINT[DI])
    expected = O.dictionary(indices, entries, n)

    def launch(out, t, library, limits):
        S.dictionary(out, t["compressed_input"], t["dictionary_entries"], n, num_entries, uncompressed_size=U, dictionary_index_size=DI,
                     index_size=I, cache_in_shared_memory=cache, library=library, limits=limits)
    return Case("dict-I%d-U%d-DI%d-D%d-n%d-cache%s" % (I, U, DI, num_entries, n, cache), "dictionary",
                {"compressed_input": indices, "dictionary_entries": entries}, dict(n=n), expected, launch)


def dzbf_case(seed, n, U, C, I=4):
    r = rng(7, seed, n, U, C)
    compressed = random_uint(r, C, n)
    expected = O.discard_zero_bytes_fixed(compressed, n, U)

    def launch(out, t, library, limits):
        S.discard_zero_bytes_fixed(out, t["compressed_input"], n, uncompressed_size=U, compressed_size=C, index_size=I, library=library, limits=limits)
    return Case("dzbf-I%d-U%d-C%d-n%d" % (I, U, C, n), "discard_zero_bytes_fixed", {"compressed_input": compressed}, dict(n=n), expected, launch)


def dzbv_case(seed, n, L, U, CS, bits, min_size, I=4, ref_ok=None, gap=0):
    r = rng(8, seed, n, L, U, CS, bits)
    max_size = min(U, min_size + (1 << bits) - 1)
    widths = r.integers(min_size, max_size + 1, n)
    column = np.zeros(n, dtype=UINT[U])
    raw = random_uint(r, U, n)
    for w in range(1, U + 1):
        mask = widths == w
        column[mask] = raw[mask] >> UINT[U](8 * (U - w)) if w < U else raw[mask]
    enc = E.discard_zero_bytes_variable_encode(column, L, CS, bits, min_size, I)
    sizes = enc.pop("element_sizes")
    if gap:  # `gap` bytes of padding after every segment: a segment's bytes end before the next one's anchor
        anchors = enc["position_anchors"].astype(np.uint64)
        ends = np.concatenate((anchors[1:], [len(enc["compressed_data"])])).astype(np.uint64)
        pieces = []
        for a, b in zip(anchors, ends):
            pieces += [enc["compressed_data"][int(a):int(b)], np.full(gap, 0xEE, np.uint8)]
        enc["compressed_data"] = np.concatenate(pieces)
        enc["position_anchors"] = (anchors + np.arange(len(anchors), dtype=np.uint64) * np.uint64(gap)).astype(enc["position_anchors"].dtype)
    # the kernels may read the aligned word holding the stream's last byte: keep the buffer word-padded
    enc["compressed_data"] = np.concatenate((enc["compressed_data"], np.zeros(8, np.uint8)))
    expected = O.discard_zero_bytes_variable(enc["compressed_data"], enc["packed_element_sizes"], enc["position_anchors"], L, n, min_size, bits, U)
    assert (expected == column).all(), (sizes[:8], column[:8], expected[:8])

    def launch(out, t, library, limits):
        S.discard_zero_bytes_variable(out, t["compressed_data"], t["packed_element_sizes"], t["position_anchors"], L, n, min_size, bits,
                                      uncompressed_size=U, container_size=CS, index_size=I, library=library, limits=limits)
    return Case("dzbv-I%d-U%d-CS%d-b%d-m%d-n%d-L%d%s" % (I, U, CS, bits, min_size, n, L, "-gap%d" % gap if gap else ""), "discard_zero_bytes_variable", enc,
                dict(n=n, L=L), expected, launch, ref_ok=((n % L == 0) if ref_ok is None else ref_ok) and not gap)


def dzbv_delta_case(seed, n, L, U, C, CS, bits, min_size, I=4):
    """The fused extension: Delta applied to DZB-variable-compressed deltas.  Expected output: the oracle's two decoders chained."""
    r = rng(12, seed, n, L, U, C, CS, bits)
    max_size = min(C, min_size + (1 << bits) - 1)
    widths = r.integers(min_size, max_size + 1, n)
    deltas = np.zeros(n, dtype=UINT[C])
    raw = random_uint(r, C, n)
    for w in range(1, C + 1):
        mask = widths == w
        deltas[mask] = raw[mask] >> UINT[C](8 * (C - w)) if w < C else raw[mask]
    base = int(random_uint(r, U, 1)[0])
    column = (np.cumsum(deltas.astype(np.uint64), dtype=np.uint64) + np.uint64(base)).astype(UINT[U])
    delta_encoded = E.delta_encode(column, L, C, base)
    assert (delta_encoded["compressed_input"] == deltas).all()
    enc = E.discard_zero_bytes_variable_encode(deltas, L, CS, bits, min_size, I)
    enc.pop("element_sizes")
    enc["compressed_data"] = np.concatenate((enc["compressed_data"], np.zeros(8, np.uint8)))
    widened = O.discard_zero_bytes_variable(enc["compressed_data"], enc["packed_element_sizes"], enc["position_anchors"], L, n, min_size, bits, C)
    expected = O.delta(widened, delta_encoded["anchor_values"], n, L)
    assert (expected == column).all()
    inputs = dict(enc, anchor_values=delta_encoded["anchor_values"])

    def launch(out, t, library, limits):
        S.delta_of_discard_zero_bytes_variable(out, t["compressed_data"], t["packed_element_sizes"], t["position_anchors"], t["anchor_values"],
                                               L, n, min_size, bits, uncompressed_size=U, compressed_size=C, container_size=CS, index_size=I,
                                               library=library, limits=limits)
    return Case("dzbv+delta-I%d-U%d-C%d-CS%d-b%d-m%d-n%d-L%d" % (I, U, C, CS, bits, min_size, n, L), "delta_of_discard_zero_bytes_variable",
                inputs, dict(n=n, L=L), expected, launch, ref_ok=False)


def bitmaps_case(seed, n, U, K, use_dictionary=True, I=4, alignment=2):
    r = rng(9, seed, n, U, K)
    indices = r.integers(0, K, n)
    bitmaps, bitmap_length = E.incidence_bitmaps_encode(indices, K, alignment)
    # elements past n in the last container belong to no bitmap: undefined in the reference; make them bitmap 0
    tail = bitmap_length * 32 - n
    if tail:
        bitmaps[bitmap_length - 1] |= np.uint32(((1 << tail) - 1) << (32 - tail))
    entries = random_uint(r, U, K) if use_dictionary else np.zeros(1, dtype=UINT[U])
    expected = O.incidence_bitmaps(bitmaps, entries, bitmap_length, K, U, alignment, use_dictionary)

    def launch(out, t, library, limits):
        S.incidence_bitmaps(out, t["incidence_bitmaps"], t["dictionary_entries"], n, bitmap_length, K, uncompressed_size=U,
                            alignment_in_ints=alignment, use_dictionary=use_dictionary, index_size=I, library=library, limits=limits)
    # reference quirk E8: the last warp reads past bitmap_length unless it is a multiple of 32 containers
    return Case("bitmaps-I%d-U%d-K%d-dict%d-n%d" % (I, U, K, use_dictionary, n), "incidence_bitmaps",
                {"incidence_bitmaps": bitmaps, "dictionary_entries": entries}, dict(n=n), expected, launch,
                ref_ok=(bitmap_length % 32 == 0))


def scatter_case(seed, target_length, p, U, OI, II):
    r = rng(10, seed, target_length, p, U, OI, II)
    target = random_uint(r, U, target_length)
    positions = np.sort(r.choice(target_length, p, replace=False)).astype(UINT[OI])
    r.shuffle(positions)
    data = random_uint(r, U, p)
    expected = O.scatter(target.copy(), data, positions)

    def launch(out, t, library, limits):
        S.scatter(out, t["data"], t["indices"], p, element_size=U, output_index_size=OI, input_index_size=II, library=library, limits=limits)
    return Case("scatter-OI%d-U%d-II%d-p%d" % (OI, U, II, p), "scatter", {"data": data, "indices": positions}, dict(p=p), expected, launch,
                initial_output=target)


def ci_scatter_case(seed, target_length, p, U, OI, II, A, max_run_length=64, widths=(8, 16, 32, 64)):
    r = rng(11, seed, target_length, p, U, OI, II, A)
    target = random_uint(r, U, target_length)
    positions = np.sort(r.choice(target_length, p, replace=False))
    data = random_uint(r, U, p)
    enc = E.compressed_indices_encode(positions, A, OI, II, max_run_length, widths)
    expected = O.compressed_indices_scatter(target.copy(), data, enc["scatter_position_run_lengths"],
                                            enc["scatter_position_run_individual_offset_sizes"], enc["scatter_position_run_baseline_values"],
                                            enc["scatter_position_run_offsets_start_pos"], enc["scatter_position_offset_bytes"],
                                            enc["scatter_position_anchors"], A)
    check = target.copy()
    check[positions] = data
    assert (expected == check).all()
    num_runs = len(enc["scatter_position_run_lengths"])
    inputs = dict(enc, data_to_scatter=data)
    inputs["scatter_position_offset_bytes"] = np.concatenate((enc["scatter_position_offset_bytes"], np.zeros(8, np.uint8)))

    def launch(out, t, library, limits):
        S.compressed_indices_scatter(out, t["data_to_scatter"], t["scatter_position_run_lengths"],
                                     t["scatter_position_run_individual_offset_sizes"], t["scatter_position_run_baseline_values"],
                                     t["scatter_position_run_offsets_start_pos"], t["scatter_position_offset_bytes"],
                                     t["scatter_position_anchors"], A, num_runs, p, element_size=U, output_index_size=OI,
                                     input_index_size=II, library=library, limits=limits)
    return Case("ciscatter-OI%d-U%d-II%d-p%d-A%d" % (OI, U, II, p, A), "compressed_indices_scatter", inputs, dict(p=p, A=A), expected, launch,
                initial_output=target)


def patched_dictionary_case(seed, n, U, DI, num_entries, p, I=4, cache=None, clustered=False):
    """Extension: Dictionary + naive patches in one pass; expected = the oracle's dictionary decode, then its scatter."""
    r = rng(12, seed, n, U, DI, num_entries, p)
    entries = random_uint(r, U, num_entries)
    indices = r.integers(0, num_entries, n, dtype=np.uint64).astype(UINT[DI])
    if clustered and p:  # all patches inside a few tiles: more patches than threads in a tile
        start = int(r.integers(0, max(n - 2 * p, 1)))
        positions = start + np.sort(r.choice(min(2 * p, n - start), p, replace=False))
    else:
        positions = np.sort(r.choice(n, p, replace=False))
    positions = positions.astype(UINT[I])
    data = random_uint(r, U, p)
    expected = O.scatter(O.dictionary(indices, entries, n), data, positions)

    def launch(out, t, library, limits):
        S.patched_dictionary(out, t["compressed_input"], t["dictionary_entries"], n, num_entries, t["patch_data"], t["patch_indices"], p,
                             uncompressed_size=U, dictionary_index_size=DI, index_size=I, cache_in_shared_memory=cache,
                             library=library, limits=limits)
    # keep the device buffers non-empty when there are no patches
    inputs = {"compressed_input": indices, "dictionary_entries": entries,
              "patch_data": data if p else np.zeros(1, UINT[U]), "patch_indices": positions if p else np.zeros(1, UINT[I])}
    return Case("patched-dict-I%d-U%d-DI%d-D%d-n%d-p%d%s" % (I, U, DI, num_entries, n, p, "-clustered" if clustered else ""),
                "patched_dictionary", inputs, dict(n=n, p=p), expected, launch, ref_ok=False)


def patched_for_case(seed, n, L, U, C, model, p, A, I=4, MI=4, max_run_length=64, widths=(8, 16, 32), clustered=False):
    """Extension: FOR + compressed-indices patches in one pass; expected = the oracle's FOR decode, then its ci-scatter."""
    r = rng(13, seed, n, L, U, C, p, A)
    num_segments = E.num_segments(n, L)
    dim = 2 if model == "linear" else 1
    coefficients = r.integers(-(1 << 30), 1 << 30, (num_segments, dim), dtype=np.int64).astype(np.int32)
    if model == "linear":
        coefficients[:, 1] = r.integers(-1000, 1000, num_segments)
    compressed = r.integers(-(1 << (8 * C - 1)), 1 << (8 * C - 1), n, dtype=np.int64).astype(INT[C])
    if clustered and p:
        start = int(r.integers(0, max(n - 2 * p, 1)))
        positions = start + np.sort(r.choice(min(2 * p, n - start), p, replace=False))
    else:
        positions = np.sort(r.choice(n, p, replace=False))
    data = random_uint(r, U, p)
    enc = E.compressed_indices_encode(positions, A, I, I, max_run_length, widths)
    decoded = O.frame_of_reference(compressed, coefficients, n, L, U, model, MI)
    if p:
        expected = O.compressed_indices_scatter(decoded.view(UINT[U]).copy(), data, enc["scatter_position_run_lengths"],
                                                enc["scatter_position_run_individual_offset_sizes"], enc["scatter_position_run_baseline_values"],
                                                enc["scatter_position_run_offsets_start_pos"], enc["scatter_position_offset_bytes"],
                                                enc["scatter_position_anchors"], A).view(decoded.dtype)
    else:
        expected = decoded
    num_runs = len(enc["scatter_position_run_lengths"])
    inputs = {k: (v if len(v) else np.zeros(1, v.dtype)) for k, v in enc.items()}
    inputs["scatter_position_offset_bytes"] = np.concatenate((enc["scatter_position_offset_bytes"], np.zeros(8, np.uint8)))
    inputs.update(compressed_input=compressed, segment_model_coefficients=coefficients,
                  data_to_scatter=(data if p else np.zeros(1, UINT[U])).view(INT[U]))

    def launch(out, t, library, limits):
        S.patched_frame_of_reference(out, t["compressed_input"], t["segment_model_coefficients"], n, L,
                                     t["data_to_scatter"], t["scatter_position_run_lengths"], t["scatter_position_run_individual_offset_sizes"],
                                     t["scatter_position_run_baseline_values"], t["scatter_position_run_offsets_start_pos"],
                                     t["scatter_position_offset_bytes"], t["scatter_position_anchors"], A, num_runs, p,
                                     uncompressed_size=U, compressed_size=C, model=model, index_size=I, model_index_size=MI,
                                     library=library, limits=limits)
    return Case("patched-for-%s-I%d-U%d-C%d-n%d-L%d-p%d-A%d-r%d%s" % (model, I, U, C, n, L, p, A, max_run_length, "-clustered" if clustered else ""),
                "patched_frame_of_reference", inputs, dict(n=n, L=L, p=p, A=A), expected, launch, ref_ok=False)


# ---- the case list ---------------------------------------------------------------------------------

def all_cases():
    """Every registered type combination of every scheme at least once, plus ragged / tiny / odd shapes."""
    cases = []
    # Frame of Reference (8 keys)
    for (I, U, C, model, MI) in [(4, 2, 1, "linear", 4), (4, 4, 1, "linear", 4), (4, 4, 2, "linear", 4), (8, 4, 2, "linear", 4),
                                 (8, 4, 2, "linear", 8), (4, 2, 1, "constant", 4), (4, 4, 1, "constant", 4), (4, 4, 2, "constant", 4)]:
        cases.append(for_case(0, 8 * 2048, 2048, U, C, model, I, MI))
        cases.append(for_case(1, 40000 + 13, 1020, U, C, model, I, MI))  # ragged tail, L not a power of 2
    cases.append(for_case(2, 1000, 4, 2, 1, "linear"))       # segments shorter than a vector
    cases.append(for_case(3, 3, 4, 4, 2, "constant"))        # shorter than one segment
    cases.append(for_case(4, 1 << 20, 4096, 4, 2, "constant"))
    # Model (5 keys)
    for (I, U, model) in [(4, 2, "linear"), (4, 4, "linear"), (8, 4, "linear"), (4, 2, "constant"), (4, 4, "constant")]:
        cases.append(model_case(0, 100000 + 7, U, model, I))
    cases.append(model_case(1, 5, 4, "linear"))
    # RLE (18 keys)
    for (I, U, RL) in [(4, u, rl) for rl in (1, 2, 4) for u in (1, 2, 4, 8)] + [(8, 4, 1), (8, 4, 2), (8, 4, 4), (8, 1, 8), (8, 4, 8), (8, 8, 8)]:
        cases.append(rle_case(0, 64 * 1024, 1024, U, RL, I))
    cases.append(rle_case(1, 50000 + 3, 1000, 4, 2))                      # ragged tail; L not a multiple of the tile
    cases.append(rle_case(2, 70000, 4096, 4, 2, mean_run=2))              # short runs, several tiles per segment
    cases.append(rle_case(3, 70000, 2048, 8, 1, mean_run=300))            # runs longer than the length type allows: split runs
    cases.append(rle_case(4, 100000, 1024, 4, 2, mean_run=5000))          # runs spanning many segments
    cases.append(rle_case(5, 9000, 1024, 1, 1, mean_run=1))               # every element its own run
    cases.append(rle_case(6, 7, 1024, 4, 2))                              # a single partial segment
    cases.append(rle_case(7, 40000, 36, 2, 2, mean_run=7))                # tiny segments
    cases.append(rle_case(8, 1 << 20, 1024, 4, 2))
    # RPE (10 keys)
    for (I, U, PO, rel) in [(4, u, 2, True) for u in (1, 2, 4, 8)] + [(8, 4, 2, True)] + [(4, u, 4, False) for u in (1, 2, 4, 8)] + [(8, 4, 8, False)]:
        cases.append(rpe_case(0, 64 * 1024, 1024, U, PO, rel, I, avoid_boundary_runs=rel))
    cases.append(rpe_case(1, 50000 + 3, 1000, 4, 4, False))
    cases.append(rpe_case(2, 70000, 4096, 4, 4, False, mean_run=2))
    cases.append(rpe_case(3, 70000, 1024, 4, 2, True, mean_run=3))        # relative, runs beginning on segment boundaries
    cases.append(rpe_case(4, 100000, 1024, 8, 4, False, mean_run=5000))
    cases.append(rpe_case(5, 1 << 20, 1024, 4, 4, False, mean_run=64))
    # Delta (12 keys)
    for I in (4, 8):
        for (U, C) in [(2, 1), (4, 1), (4, 2), (8, 1), (8, 2), (8, 4)]:
            cases.append(delta_case(0, 32 * 1024, 1024, U, C, I))
    cases.append(delta_case(1, 50000 + 5, 1000, 8, 4))                    # ragged; L % E == 0
    cases.append(delta_case(2, 50000 + 5, 999, 8, 4))                     # L not a multiple of the vector width
    cases.append(delta_case(3, 10000, 24, 4, 2))                          # segments shorter than a row
    cases.append(delta_case(4, 300000, 100000, 8, 4))                     # long segments
    cases.append(delta_case(5, 1 << 20, 1024, 8, 4))
    # few, long segments: the decoupled look-back path (tiles of a segment scanned by different CTAs)
    cases.append(delta_case(6, 1 << 20, 1 << 20, 8, 4))                   # a single segment: a plain prefix sum over the column
    cases.append(delta_case(7, 4 << 16, 1 << 16, 4, 2))                   # the reference is defined here too
    cases.append(delta_case(8, 500000 + 7, 200000, 8, 4))                 # ragged last segment
    cases.append(delta_case(9, 600000, 131071, 4, 1))                     # L odd: element-wise tiles
    cases.append(delta_case(10, 1 << 20, 1 << 18, 2, 1))
    cases.append(delta_case(11, 3 << 18, 1 << 18, 8, 1, I=8))
    cases.append(delta_case(12, (1 << 22) + 5, 1 << 21, 8, 2))
    cases.append(delta_case(13, (1 << 23) + 77, 1 << 23, 4, 1))             # > 1000 tiles in one segment: look-backs of several rounds
    cases.append(delta_case(14, 3 * 24576 + 100, 24576, 8, 4))              # the shortest segments which take the look-back path
    cases.append(delta_case(15, 700000 + 1, 200002, 8, 4))                  # L a multiple of the 16-byte vector but not of a wide piece: no bulk copies
    # Dictionary (20 keys)
    for I in (4, 8):
        for (U, DI) in [(1, 1), (2, 1), (2, 2), (4, 1), (4, 2), (4, 4), (8, 1), (8, 2), (8, 4), (8, 8)]:
            cases.append(dictionary_case(0, 100000 + 11, U, DI, min(1 << (8 * DI), 1000), I))
    cases.append(dictionary_case(1, 1 << 20, 8, 1, 256, cache=True))
    cases.append(dictionary_case(2, 1 << 20, 8, 1, 256, cache=False))
    cases.append(dictionary_case(3, 1 << 18, 4, 2, 40000, cache=True))    # 160 KB dictionary in shared memory
    cases.append(dictionary_case(4, 9, 8, 1, 3))
    cases.append(dictionary_case(5, 1 << 19, 4, 2, 60000, cache=True))    # 240 KB: its head in shared memory, its tail through L1/L2
    cases.append(dictionary_case(6, 1 << 18, 8, 4, 100000, cache=True))   # 800 KB: mostly through L1/L2
    cases.append(dictionary_case(7, 1 << 18, 2, 2, 65536, cache=True))    # all 16-bit indices, 128 KB
    # DZB fixed (12 keys)
    for I in (4, 8):
        for (U, C) in [(2, 1), (4, 1), (4, 2), (8, 1), (8, 2), (8, 4)]:
            cases.append(dzbf_case(0, 100000 + 9, U, C, I))
    # DZB variable (9 keys)
    for (I, U, CS) in [(4, 2, 4), (4, 4, 4), (4, 8, 4), (8, 2, 4), (8, 4, 4), (8, 8, 4), (4, 2, 8), (4, 4, 8), (4, 8, 8)]:
        bits, min_size = (1, 1) if U == 2 else ((2, 1) if U == 4 else (3, 1))
        cases.append(dzbv_case(0, 32 * 1024, 1024, U, CS, bits, min_size, I))
    cases.append(dzbv_case(1, 50000 + 3, 1000, 4, 4, 2, 1))               # ragged tail
    cases.append(dzbv_case(2, 50000, 1001, 4, 4, 2, 1))                   # unaligned segments
    cases.append(dzbv_case(3, 40000, 1024, 4, 4, 3, 0))                   # 3-bit fields (10 per container), zero-size elements
    cases.append(dzbv_case(4, 40000, 1024, 8, 8, 4, 0))
    cases.append(dzbv_case(5, 1 << 20, 1024, 4, 4, 2, 1))
    # the staged (TMA) path's corners: several tiles per segment, partial tiles, ragged tails, tiny segments, zero-size elements
    cases.append(dzbv_case(6, 5 * 4096, 4096, 4, 4, 2, 1))
    cases.append(dzbv_case(7, 3 * 2048 + 1040, 2048, 8, 8, 4, 0, ref_ok=False))
    cases.append(dzbv_case(8, 50000 + 3, 1008, 4, 4, 2, 1, ref_ok=False))
    cases.append(dzbv_case(9, 40000 + 7, 1040, 4, 8, 2, 0, ref_ok=False))
    cases.append(dzbv_case(10, 4000, 16, 4, 4, 2, 1, ref_ok=False))
    cases.append(dzbv_case(11, 70000, 3000 * 16, 2, 4, 1, 1, ref_ok=False))
    cases.append(dzbv_case(12, 33000, 1024, 4, 4, 4, 0, ref_ok=False))
    cases.append(dzbv_case(13, 64 * 1024, 1024, 8, 4, 2, 5, ref_ok=False))
    cases.append(dzbv_case(14, 20, 1024, 4, 4, 2, 1, ref_ok=False))
    # padding between segments: what lies between a segment's last byte and the next anchor is never decoded (nor staged)
    cases.append(dzbv_case(15, 40 * 1024, 1024, 4, 4, 2, 1, gap=8))
    cases.append(dzbv_case(16, 20 * 1024 + 3, 1024, 4, 4, 2, 1, gap=5000))           # farther than a whole tile of bytes
    cases.append(dzbv_case(17, 16 * 1024, 2048, 8, 4, 2, 5, gap=13))                 # two tiles per segment, unaligned segments
    # Extension: Delta of DZB-variable, fused (5 keys)
    for (I, U, C, CS) in [(4, 8, 4, 4), (4, 8, 4, 8), (4, 8, 2, 4), (4, 4, 2, 4), (8, 8, 4, 4)]:
        bits = 2 if C == 4 else 1
        cases.append(dzbv_delta_case(0, 32 * 1024, 1024, U, C, CS, bits, 1, I))
    cases.append(dzbv_delta_case(1, 50000 + 3, 1008, 8, 4, 4, 2, 1))      # partial tiles, ragged tail
    cases.append(dzbv_delta_case(2, 5 * 4096, 4096, 8, 4, 4, 2, 0))       # several tiles per segment, zero-size deltas
    cases.append(dzbv_delta_case(3, 4000 + 5, 16, 8, 4, 4, 2, 1))         # tiny segments
    cases.append(dzbv_delta_case(4, 33000, 1024, 8, 4, 4, 4, 0))          # 4-bit fields
    cases.append(dzbv_delta_case(5, 1 << 20, 1024, 8, 4, 4, 2, 1))
    # Extension: Dictionary + naive patches, fused (10 keys)
    for (I, U, DI) in [(4, 1, 1), (4, 2, 1), (4, 4, 1), (4, 8, 1), (4, 2, 2), (4, 4, 2), (4, 8, 2), (4, 8, 4), (8, 8, 1), (8, 4, 2)]:
        cases.append(patched_dictionary_case(0, 100000 + 11, U, DI, min(1 << (8 * DI), 1000), 1000, I))
    cases.append(patched_dictionary_case(1, 1 << 20, 8, 1, 256, 10485, cache=True))      # 1 % exceptions (BASELINE configs[2] in small)
    cases.append(patched_dictionary_case(2, 1 << 20, 8, 1, 256, 10485, cache=False))
    cases.append(patched_dictionary_case(3, 300000, 8, 1, 256, 0))                       # no patches at all
    cases.append(patched_dictionary_case(4, 300000 + 1, 8, 1, 256, 1))                   # a single patch; one trailing element
    cases.append(patched_dictionary_case(5, 200000, 4, 1, 200, 3000, clustered=True))    # more patches in a tile than the CTA has threads
    cases.append(patched_dictionary_case(6, 50000, 2, 2, 5000, 50000))                   # every element patched
    cases.append(patched_dictionary_case(7, 9, 8, 1, 3, 2))
    cases.append(patched_dictionary_case(8, 1 << 22, 4, 1, 256, 200000))
    # cached dictionaries between 15 and 48 KiB: the kernel's own static shared memory plus the dictionary exceed the 48 KiB a
    # kernel gets without opting in (the launcher must opt in on the sum, not on the dynamic part alone)
    cases.append(patched_dictionary_case(9, 1 << 18, 8, 2, 4096, 3000, cache=True))       # 32 KiB
    cases.append(patched_dictionary_case(10, 1 << 18, 4, 2, 6000, 3000, cache=True))      # 23.4 KiB
    cases.append(patched_dictionary_case(11, 1 << 18, 8, 2, 5800, 3000, cache=True))      # 45.3 KiB
    # many patches per tile: several patches in one 16-byte vector, more than 32 (and more than 255) patches in a tile
    cases.append(patched_dictionary_case(12, 1 << 16, 1, 1, 256, 20000))                  # U = 1: 4096-element tiles, ~1250 patches each
    cases.append(patched_dictionary_case(13, 1 << 17, 2, 1, 256, 30000))
    cases.append(patched_dictionary_case(14, 1 << 18, 8, 1, 256, 100000))                 # 38 % of the elements patched
    # Extension: Frame of Reference + compressed-indices patches, fused (7 keys)
    for (I, U, C, model, MI) in [(4, 2, 1, "linear", 4), (4, 4, 1, "linear", 4), (4, 4, 2, "linear", 4), (8, 4, 2, "linear", 8),
                                 (4, 2, 1, "constant", 4), (4, 4, 1, "constant", 4), (4, 4, 2, "constant", 4)]:
        cases.append(patched_for_case(0, 40 * 2048, 2048, U, C, model, 2000, 256, I, MI))
    cases.append(patched_for_case(1, 1 << 20, 1024, 4, 2, "linear", 10485, 1024))                   # BASELINE configs[4] column A in small
    cases.append(patched_for_case(2, 100000 + 13, 1020, 4, 2, "linear", 3000, 1000, max_run_length=7))   # ragged; many short runs
    cases.append(patched_for_case(3, 300000, 4096, 4, 2, "constant", 0, 1024))                      # no patches
    cases.append(patched_for_case(4, 300000, 4096, 4, 2, "constant", 1, 1024))
    cases.append(patched_for_case(5, 200000, 1024, 4, 1, "linear", 5000, 4096, max_run_length=3000, clustered=True))  # runs longer than the window
    cases.append(patched_for_case(6, 60000, 1024, 2, 1, "constant", 60000, 1024, max_run_length=1))  # every element patched, one-patch runs
    cases.append(patched_for_case(7, 1 << 22, 4096, 4, 2, "constant", 150000, 100000, max_run_length=500))  # long anchoring period
    cases.append(patched_for_case(8, 1 << 21, 1024, 4, 2, "linear", 30000, 16, max_run_length=16))   # tiny anchoring period
    cases.append(patched_for_case(9, 1 << 18, 1024, 2, 1, "linear", 60000, 1024))                    # U = 2: ~470 patches per 2048-element tile
    cases.append(patched_for_case(10, 1 << 18, 1024, 4, 2, "constant", 80000, 512, max_run_length=32))
    # Incidence bitmaps (5 keys)
    for (U, use_dictionary) in [(1, True), (2, True), (4, True), (8, True), (1, False)]:
        cases.append(bitmaps_case(0, 64 * 1024, U, 16, use_dictionary))
    cases.append(bitmaps_case(1, 100000 + 21, 4, 5))                      # ragged
    cases.append(bitmaps_case(2, 1 << 18, 4, 200))
    cases.append(bitmaps_case(3, 40, 4, 2))
    cases.append(bitmaps_case(4, 1 << 16, 8, 1))                          # a single bitmap
    # Scatter (32 keys) and compressed-indices scatter (32 keys)
    for OI in (4, 8):
        for U in (1, 2, 4, 8):
            for II in (1, 2, 4, 8):
                p = min(2000, (1 << (8 * II)) - 1)
                cases.append(scatter_case(0, 200000, p, U, OI, II))
                A = min(1024, max(p // 3, 16))
                if A < (1 << (8 * II)):
                    cases.append(ci_scatter_case(0, 200000, p, U, OI, II, A, max_run_length=min(50, (1 << (8 * II)) - 1)))
    cases.append(scatter_case(1, 1 << 22, 41943, 8, 4, 4))
    cases.append(ci_scatter_case(1, 1 << 22, 41943, 4, 4, 4, 1024))
    cases.append(ci_scatter_case(2, 1 << 22, 3000, 8, 8, 4, 1000, widths=(8, 16, 32)))
    return cases


def small_cases():
    """A subset which finishes quickly: used for host-side (oracle-only) checks."""
    return [c for c in all_cases() if c.output_elements <= 300000]
