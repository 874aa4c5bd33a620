"""The GPU compressors (include/giddy_b200.h group 8; keys "compression::...").

The reference ships no compressor (README.md:128-130), so parity here has two legs:
  * byte-identical to libgiddy_b200/encode.py, the numpy statement of the formats of SURVEY.md Appendix A -- which is
    itself pinned to the reference through the oracle (every tests/cases.py case asserts oracle(encode(x)) == x, and the
    oracle is pinned to the reference's kernels by tests/golden/ref_gpu_digests.json);
  * round trip: decompressing the GPU-compressed column with the library's own decompressors gives the column back.
"""
import numpy as np
import pytest

import cases
from libgiddy_b200 import compress as K
from libgiddy_b200 import encode as E
from libgiddy_b200 import schemes as S

UINT, INT = E.UINT, E.INT


def _dev(a):
    import torch
    return torch.from_numpy(cases._as_torch_compatible(np.ascontiguousarray(a))).cuda()


def _empty(count, dtype, fill=0x5A):
    import torch
    itemsize = np.dtype(dtype).itemsize
    return torch.full((max(int(count), 1) * itemsize,), fill, dtype=torch.uint8, device="cuda").view(cases._torch_int(itemsize))


def _host(t, dtype, count=None):
    a = t.cpu().numpy().view(dtype)
    return a if count is None else a[:count]


def _status():
    import torch
    return torch.zeros(1, dtype=torch.int32, device="cuda")


def test_compressor_keys_are_registered(library):
    keys = set(library.registered_keys())
    expected = [K.delta_key(4, 8, 4), K.delta_key(8, 2, 1), K.discard_zero_bytes_fixed_key(4, 4, 1), K.dictionary_key(8, 8, 8),
                K.frame_of_reference_key(4, 4, 2, "linear"), K.frame_of_reference_key(4, 2, 1, "constant"), K.run_length_encoding_key(4, 4, 2),
                K.run_length_encoding_key(8, 8, 8), K.run_position_encoding_key(4, 4, 2, True), K.run_position_encoding_key(8, 4, 8, False),
                K.discard_zero_bytes_variable_key(4, 8, 8), K.incidence_bitmaps_key(4, 4, 2, True), K.incidence_bitmaps_key(4, 1, 2, False)]
    assert all(k in keys for k in expected), [k for k in expected if k not in keys]
    assert sum(k.startswith("compression::") for k in keys) == 7 + 12 + 12 + 20 + 18 + 10 + 9 + 5


@pytest.mark.gpu
@pytest.mark.parametrize("I", [4, 8])
@pytest.mark.parametrize("U,C", [(2, 1), (4, 1), (4, 2), (8, 1), (8, 2), (8, 4)])
@pytest.mark.parametrize("n,L", [(40000 + 7, 1000), (5, 1024)])
def test_delta(library, I, U, C, n, L):
    case = cases.delta_case(31, n, L, U, C, I)
    column, base = case.expected, int(case.inputs["anchor_values"][0])
    x, out, anchors, status = _dev(column), _empty(n, UINT[C]), _empty(E.num_segments(n, L), UINT[U]), _status()
    K.delta(out, anchors, x, n, L, uncompressed_size=U, compressed_size=C, base_value=base, index_size=I, status=status, library=library)
    assert (_host(out, UINT[C], n) == case.inputs["compressed_input"]).all()
    assert (_host(anchors, UINT[U], E.num_segments(n, L)) == case.inputs["anchor_values"]).all()
    assert int(status.item()) == 0
    back = _empty(n, UINT[U])
    S.delta(back, out, anchors, n, L, uncompressed_size=U, compressed_size=C, index_size=I, library=library)
    assert (_host(back, UINT[U], n) == column).all()


@pytest.mark.gpu
def test_delta_reports_differences_which_do_not_fit(library):
    column = np.arange(0, 3000 * 300, 300, dtype=np.uint32)  # differences of 300 do not fit one byte
    x, out, anchors, status = _dev(column), _empty(len(column), np.uint8), _empty(3, np.uint32), _status()
    K.delta(out, anchors, x, len(column), 1024, uncompressed_size=4, compressed_size=1, status=status, library=library)
    assert int(status.item()) & K.UNREPRESENTABLE


@pytest.mark.gpu
@pytest.mark.parametrize("I", [4, 8])
@pytest.mark.parametrize("U,C", [(2, 1), (4, 1), (4, 2), (8, 1), (8, 2), (8, 4)])
def test_discard_zero_bytes_fixed(library, I, U, C):
    n = 70000 + 3
    compressed = cases.random_uint(cases.rng(32, U, C), C, n)
    column = compressed.astype(UINT[U])
    x, out, status = _dev(column), _empty(n, UINT[C]), _status()
    K.discard_zero_bytes_fixed(out, x, n, uncompressed_size=U, compressed_size=C, index_size=I, status=status, library=library)
    assert (_host(out, UINT[C], n) == compressed).all() and int(status.item()) == 0
    column[n // 2] = UINT[U](1) << UINT[U](8 * C)  # one element with a non-zero high byte
    K.discard_zero_bytes_fixed(out, _dev(column), n, uncompressed_size=U, compressed_size=C, index_size=I, status=status, library=library)
    assert int(status.item()) & K.UNREPRESENTABLE


@pytest.mark.gpu
@pytest.mark.parametrize("I,U,DI,D", [(4, 1, 1, 200), (4, 2, 1, 256), (4, 2, 2, 5000), (4, 4, 1, 17), (4, 4, 2, 40000), (4, 4, 4, 100000), (4, 8, 1, 256),
                                      (8, 8, 2, 3000), (8, 8, 4, 70000), (8, 8, 8, 1000), (4, 8, 2, 4096)])
def test_dictionary(library, I, U, DI, D):
    r = cases.rng(33, U, DI, D)
    entries = np.unique(cases.random_uint(r, U, D))
    D = len(entries)
    n = 150000 + 11
    column = entries[r.integers(0, D, n)]
    enc = E.dictionary_encode(column, DI)
    # (numpy.unique of the column may have fewer entries than the dictionary we search: compare through the entries)
    x, out, table, status = _dev(column), _empty(n, UINT[DI]), _dev(entries), _status()
    K.dictionary(out, x, table, D, n, uncompressed_size=U, dictionary_index_size=DI, index_size=I, status=status, library=library)
    indices = _host(out, UINT[DI], n)
    assert (entries[indices] == column).all() and int(status.item()) == 0
    assert (enc["dictionary_entries"][enc["compressed_input"]] == entries[indices]).all()
    back = _empty(n, UINT[U])
    S.dictionary(back, out, table, n, D, uncompressed_size=U, dictionary_index_size=DI, index_size=I, library=library)
    assert (_host(back, UINT[U], n) == column).all()
    # a value which is not in the dictionary
    stranger = column.copy()
    candidate = UINT[U](entries[0] + UINT[U](1))
    if candidate not in entries:
        stranger[7] = candidate
        K.dictionary(out, _dev(stranger), table, D, n, uncompressed_size=U, dictionary_index_size=DI, index_size=I, status=status, library=library)
        assert int(status.item()) & K.UNREPRESENTABLE


@pytest.mark.gpu
@pytest.mark.parametrize("I,U,C,model", [(4, 2, 1, "linear"), (4, 4, 1, "linear"), (4, 4, 2, "linear"), (8, 4, 2, "linear"),
                                         (4, 2, 1, "constant"), (4, 4, 1, "constant"), (4, 4, 2, "constant")])
@pytest.mark.parametrize("n,L", [(64 * 1024, 1024), (50000 + 13, 1020), (7, 4096),
                                 (200000 + 5, 4096), (4 * 8192, 8192)])  # segments a CTA holds in registers (+ a ragged last one)
def test_frame_of_reference(library, I, U, C, model, n, L):
    r = cases.rng(34, U, C, n, L)
    # a column which a segment-wise model fits: a level (and, for the linear model, a slope) per segment plus noise within the
    # compressed range; the noise vanishes at both ends of a segment so that the fitted line through them is the true one
    num_segments = E.num_segments(n, L)
    noise = 1 << (8 * C - 2)
    level = np.repeat(r.integers(-(1 << (8 * U - 3)), 1 << (8 * U - 3), num_segments), L)[:n]
    slopes = r.integers(-3, 4, num_segments) if model == "linear" else np.zeros(num_segments, np.int64)
    if n % L:
        slopes[-1] = 0  # (a short last segment is fitted through the column's last element: keep it flat)
    x_in_segment = np.arange(n) % L
    wobble = r.integers(-noise // 2, noise // 2, n)
    wobble[(x_in_segment == 0) | (x_in_segment == L - 1)] = 0
    wobble[-1] = 0
    column = (level + np.repeat(slopes, L)[:n] * x_in_segment + wobble).astype(INT[U])
    enc = E.frame_of_reference_encode(column, L, C, model)
    x, out, coefficients, status = _dev(column), _empty(n, INT[C]), _empty(num_segments * (2 if model == "linear" else 1), np.int32), _status()
    K.frame_of_reference(out, coefficients, x, n, L, uncompressed_size=U, compressed_size=C, model=model, index_size=I, status=status, library=library)
    assert int(status.item()) == 0
    got = _host(coefficients, np.int32).reshape(num_segments, -1)
    assert (got == enc["segment_model_coefficients"]).all()
    assert (_host(out, INT[C], n) == enc["compressed_input"]).all()
    back = _empty(n, INT[U])
    S.frame_of_reference(back, out, coefficients, n, L, uncompressed_size=U, compressed_size=C, model=model, index_size=I, library=library)
    assert (_host(back, INT[U], n) == column).all()


@pytest.mark.gpu
def test_frame_of_reference_reports_offsets_which_do_not_fit(library):
    column = np.zeros(4096, dtype=np.int32)
    column[100] = 1 << 20  # one outlier: the segment's range exceeds an int8 offset
    out, coefficients, status = _empty(4096, np.int8), _empty(4, np.int32), _status()
    K.frame_of_reference(out, coefficients, _dev(column), 4096, 1024, uncompressed_size=4, compressed_size=1, model="constant", status=status, library=library)
    assert int(status.item()) & K.UNREPRESENTABLE


RLE_SHAPES = [(64 * 1024, 1024, 16), (50000 + 3, 1000, 16), (70000, 4096, 2), (100000, 1024, 5000), (9000, 1024, 1), (7, 1024, 16), (40000, 36, 7),
              (300000, 1024, 700)]


@pytest.mark.gpu
@pytest.mark.parametrize("I,U,RL", [(4, u, rl) for rl in (1, 2, 4) for u in (1, 2, 4, 8)] + [(8, 4, 1), (8, 4, 2), (8, 4, 4), (8, 1, 8), (8, 4, 8), (8, 8, 8)])
@pytest.mark.parametrize("n,L,mean_run", RLE_SHAPES[:4])
def test_run_length_encoding_types(library, I, U, RL, n, L, mean_run):
    _check_rle(library, I, U, RL, n, L, mean_run)


@pytest.mark.gpu
@pytest.mark.parametrize("n,L,mean_run", RLE_SHAPES)
@pytest.mark.parametrize("max_run_length", [0, 5, 255])
def test_run_length_encoding_shapes(library, n, L, mean_run, max_run_length):
    _check_rle(library, 4, 4, 2, n, L, mean_run, max_run_length)


@pytest.mark.gpu
@pytest.mark.parametrize("RL,max_run_length", [(2, 0), (4, 40000), (4, 32768), (2, 50000), (4, 0)])
@pytest.mark.parametrize("n,mean_run", [(500000 + 11, 120000), (1 << 20, 30000), (300000, 300000)])
def test_run_length_encoding_runs_longer_than_a_super_tile(library, RL, max_run_length, n, mean_run):
    """The one-pass encoder: super-tiles of 32768 elements without a run head, split points of runs which began several
    super-tiles earlier (at the length type's limit, or at a limit of the caller's)."""
    _check_rle(library, 4, 4, RL, n, 1024, mean_run, max_run_length)


def _check_rle(library, I, U, RL, n, L, mean_run, max_run_length=0):
    import torch
    column = cases._random_runs_column(cases.rng(35, n, L, U, RL), n, U, mean_run)
    enc = E.run_length_encode(column, L, RL, I, max_run_length or None)
    R, S_ = len(enc["run_data"]), E.num_segments(n, L)
    capacity = R + 5
    run_data, run_lengths = _empty(capacity, UINT[U]), _empty(capacity, UINT[RL])
    anchors, offsets = _empty(S_, UINT[I]), _empty(S_, UINT[RL])
    count, status = torch.zeros(1, dtype=torch.int64 if I > 4 else torch.int32, device="cuda"), _status()
    K.run_length_encoding(run_data, run_lengths, anchors, offsets, count, _dev(column), n, L, capacity, uncompressed_size=U, run_length_size=RL,
                          index_size=I, max_run_length=max_run_length, status=status, library=library)
    assert int(status.item()) == 0 and int(count.item()) == R
    assert (_host(run_data, UINT[U], R) == enc["run_data"]).all()
    assert (_host(run_lengths, UINT[RL], R) == enc["run_lengths"]).all()
    assert (_host(anchors, UINT[I], S_) == enc["position_anchors"]).all()
    assert (_host(offsets, UINT[RL], S_) == enc["intra_run_anchor_offsets"]).all()
    assert (_host(run_data, np.uint8)[R * U:] == 0x5A).all() and (_host(run_lengths, np.uint8)[R * RL:] == 0x5A).all()  # nothing past the last run
    back = _empty(n, UINT[U])
    S.run_length_encoding(back, run_data, run_lengths, anchors, offsets, L, S_, R, n, uncompressed_size=U, run_length_size=RL, index_size=I, library=library)
    assert (_host(back, UINT[U], n) == column).all()


@pytest.mark.gpu
def test_run_length_encoding_reports_too_small_a_capacity(library):
    import torch
    column = cases._random_runs_column(cases.rng(36), 50000, 4, 4)
    R = len(E.run_length_encode(column, 1024, 2, 4)["run_data"])
    run_data, run_lengths, anchors, offsets = _empty(100, np.uint32), _empty(100, np.uint16), _empty(49, np.uint32), _empty(49, np.uint16)
    count, status = torch.zeros(1, dtype=torch.int32, device="cuda"), _status()
    K.run_length_encoding(run_data, run_lengths, anchors, offsets, count, _dev(column), 50000, 1024, 100, uncompressed_size=4, run_length_size=2,
                          status=status, library=library)
    assert int(status.item()) & K.CAPACITY and int(count.item()) == R  # the true count is still reported


@pytest.mark.gpu
@pytest.mark.parametrize("I,U,PO,relative", [(4, u, 2, True) for u in (1, 2, 4, 8)] + [(8, 4, 2, True)] + [(4, u, 4, False) for u in (1, 2, 4, 8)] + [(8, 4, 8, False)])
@pytest.mark.parametrize("n,L,mean_run", [(64 * 1024, 1024, 16), (50000 + 3, 1000, 64), (100000, 4096, 5000), (9000, 512, 1)])
def test_run_position_encoding(library, I, U, PO, relative, n, L, mean_run):
    import torch
    column = cases._random_runs_column(cases.rng(37, n, L, U, PO), n, U, mean_run)
    enc = E.run_position_encode(column, L, PO, relative, I)
    R, S_ = len(enc["run_data"]), E.num_segments(n, L)
    run_data, starts, anchors = _empty(R + 3, UINT[U]), _empty(R + 3, UINT[PO]), _empty(S_, UINT[I])
    count, status = torch.zeros(1, dtype=torch.int64 if I > 4 else torch.int32, device="cuda"), _status()
    K.run_position_encoding(run_data, starts, anchors, count, _dev(column), n, L, R + 3, uncompressed_size=U, position_offset_size=PO, relative=relative,
                            index_size=I, status=status, library=library)
    assert int(status.item()) == 0 and int(count.item()) == R
    assert (_host(run_data, UINT[U], R) == enc["run_data"]).all()
    assert (_host(starts, UINT[PO], R) == enc["run_start_positions"]).all()
    assert (_host(anchors, UINT[I], S_) == enc["position_anchors"]).all()
    back = _empty(n, UINT[U])
    S.run_position_encoding(back, run_data, starts, anchors, L, S_, R, n, uncompressed_size=U, position_offset_size=PO, relative=relative, index_size=I,
                            library=library)
    # (relative positions: a run beginning exactly on a segment boundary is indistinguishable from one beginning a segment earlier in the
    # reference's format -- quirk E4 --; the library's decoder, like the oracle, resolves it as the format intends)
    assert (_host(back, UINT[U], n) == column).all()


@pytest.mark.gpu
@pytest.mark.parametrize("I,U,CS,bits,min_size", [(4, 2, 4, 1, 1), (4, 4, 4, 2, 1), (4, 8, 4, 3, 1), (8, 2, 4, 1, 1), (8, 4, 4, 2, 1), (8, 8, 4, 4, 0),
                                                  (4, 2, 8, 2, 0), (4, 4, 8, 3, 0), (4, 8, 8, 4, 0), (4, 4, 4, 4, 0), (4, 8, 4, 2, 5)])
@pytest.mark.parametrize("n,L", [(32 * 1024, 1024), (50000 + 3, 1000), (20, 1024), (70000, 48)])
def test_discard_zero_bytes_variable(library, I, U, CS, bits, min_size, n, L):
    import torch
    r = cases.rng(38, U, CS, bits, n, L)
    max_size = min(U, min_size + (1 << bits) - 1)
    widths = r.integers(min_size, max_size + 1, n)
    raw = cases.random_uint(r, U, n)
    column = np.zeros(n, dtype=UINT[U])
    for w in range(1, U + 1):
        mask = widths == w
        column[mask] = raw[mask] >> UINT[U](8 * (U - w)) if w < U else raw[mask]
    enc = E.discard_zero_bytes_variable_encode(column, L, CS, bits, min_size, I)
    B, S_ = len(enc["compressed_data"]), E.num_segments(n, L)
    data, packed, anchors = _empty(B + 16, np.uint8), _empty(len(enc["packed_element_sizes"]), UINT[CS]), _empty(S_, UINT[I])
    num_bytes, status = torch.zeros(1, dtype=torch.int64, device="cuda"), _status()
    K.discard_zero_bytes_variable(data, packed, anchors, num_bytes, _dev(column), n, L, min_size, bits, B + 16, uncompressed_size=U, container_size=CS,
                                  index_size=I, status=status, library=library)
    assert int(status.item()) == 0 and int(num_bytes.item()) == B
    assert (_host(data, np.uint8, B) == enc["compressed_data"]).all()
    assert (_host(data, np.uint8)[B:] == 0x5A).all()
    assert (_host(packed, UINT[CS], len(enc["packed_element_sizes"])) == enc["packed_element_sizes"]).all()
    assert (_host(anchors, UINT[I], S_) == enc["position_anchors"]).all()
    back = _empty(n, UINT[U])
    S.discard_zero_bytes_variable(back, data, packed, anchors, L, n, min_size, bits, uncompressed_size=U, container_size=CS, index_size=I, library=library)
    assert (_host(back, UINT[U], n) == column).all()


@pytest.mark.gpu
def test_discard_zero_bytes_variable_reports_sizes_which_do_not_fit(library):
    import torch
    column = np.full(5000, 0x01020304, dtype=np.uint32)  # 4 significant bytes; 1-bit size fields over a minimum of 1 hold sizes 1 and 2 only
    data, packed, anchors, status = _empty(20016, np.uint8), _empty(200, np.uint32), _empty(5, np.uint32), _status()
    K.discard_zero_bytes_variable(data, packed, anchors, torch.zeros(1, dtype=torch.int64, device="cuda"), _dev(column), 5000, 1024, 1, 1, 20016,
                                  uncompressed_size=4, status=status, library=library)
    assert int(status.item()) & K.UNREPRESENTABLE


@pytest.mark.gpu
@pytest.mark.parametrize("U,use_dictionary", [(1, True), (2, True), (4, True), (8, True), (1, False)])
@pytest.mark.parametrize("n,K_", [(64 * 1024, 16), (100000 + 21, 5), (40, 2), (1 << 16, 1), (70000, 200)])
def test_incidence_bitmaps(library, U, use_dictionary, n, K_):
    r = cases.rng(39, U, n, K_)
    indices = r.integers(0, K_, n)
    entries = np.unique(cases.random_uint(r, U, 4 * K_))[:K_] if use_dictionary else np.arange(K_, dtype=UINT[U])
    if len(entries) < K_:
        pytest.skip("not enough distinct values of this width")
    column = entries[indices]
    bitmaps, bitmap_length = E.incidence_bitmaps_encode(indices, K_, 2)
    out, status = _empty(len(bitmaps), np.uint32), _status()
    K.incidence_bitmaps(out, _dev(column), _dev(entries), n, K_, uncompressed_size=U, use_dictionary=use_dictionary, status=status, library=library)
    assert int(status.item()) == 0
    got = _host(out, np.uint32, len(bitmaps)).copy()
    aligned = len(bitmaps) // K_
    # (alignment padding words are not written; compare the bitmap_length words of every bitmap)
    assert (got.reshape(K_, aligned)[:, :bitmap_length] == bitmaps.reshape(K_, aligned)[:, :bitmap_length]).all()
