"""The kernels which the reference has but neither wraps nor registers (include/giddy_b200.h group 9), behind
factory keys of their own: gather (src/kernels/data_layout/gather.cuh:151-164), model::decompress_with_intervals
(src/kernels/decompression/model.cuh:55-98) and the prefix sum of src/kernels/reduction/scan.cuh.  Expected values: numpy
restatements of the reference kernels' definitions, written here (the reference never runs them: no fixture to borrow)."""
import numpy as np
import pytest

import cases
from libgiddy_b200 import encode as E
from libgiddy_b200 import schemes as S

UINT, INT = E.UINT, E.INT


def _dev(a):
    import torch
    return torch.from_numpy(cases._as_torch_compatible(np.ascontiguousarray(a))).cuda()


def _out(count, dtype):
    import torch
    itemsize = np.dtype(dtype).itemsize
    return torch.full(((count + 64) * itemsize,), 0x5A, dtype=torch.uint8, device="cuda").view(cases._torch_int(itemsize))


def _result(t, dtype, count):
    a = t.cpu().numpy().view(dtype)
    assert (a[count:].view(np.uint8) == 0x5A).all(), "wrote past the end of the output"
    return a[:count]


def test_utility_keys_are_registered(library):
    keys = set(library.registered_keys())
    assert sum(k.startswith("gather::") for k in keys) == 20
    assert sum(k.startswith("decompression::model::intervals::") for k in keys) == 5
    assert sum(k.startswith("reduction::scan::") for k in keys) == 12
    assert S.gather_key(4, 8, 1) in keys and S.model_with_intervals_key(8, 4, "linear") in keys and S.inclusive_scan_key(4, 8, 4) in keys


@pytest.mark.gpu
@pytest.mark.parametrize("OI", [4, 8])
@pytest.mark.parametrize("U,II,D", [(1, 1, 200), (2, 1, 256), (2, 2, 30000), (4, 1, 100), (4, 2, 65536), (4, 4, 300000), (8, 1, 256), (8, 2, 5000), (8, 4, 100000), (8, 8, 70000)])
def test_gather(library, OI, U, II, D):
    r = cases.rng(51, OI, U, II, D)
    data = cases.random_uint(r, U, D)
    n = 200000 + 13
    indices = r.integers(0, D, n, dtype=np.uint64).astype(UINT[II])
    out = _out(n, UINT[U])
    S.gather(out, _dev(data), _dev(indices), D, n, element_size=U, input_index_size=II, output_index_size=OI, library=library)
    assert (_result(out, UINT[U], n) == data[indices]).all()


@pytest.mark.gpu
@pytest.mark.parametrize("I,U,model", [(4, 2, "linear"), (4, 4, "linear"), (8, 4, "linear"), (4, 2, "constant"), (4, 4, "constant")])
@pytest.mark.parametrize("n,period", [(100000 + 7, 1000), (64 * 1024, 1024), (5000, 6), (3, 4096)])
def test_model_with_intervals(library, I, U, model, n, period):
    r = cases.rng(52, I, U, n, period)
    intervals = E.num_segments(n, period)
    dim = 2 if model == "linear" else 1
    coefficients = r.integers(-(1 << (8 * U - 2)), 1 << (8 * U - 2), (intervals, dim), dtype=np.int64)
    if model == "linear":
        coefficients[:, 1] = r.integers(-300, 300, intervals)
    coefficients = coefficients.astype(INT[U])
    x = (np.arange(n) % period).astype(np.int64)
    c = np.repeat(coefficients.astype(np.int64), period, axis=0)[:n]
    expected = (c[:, 0] + (c[:, 1] * x if model == "linear" else 0)).astype(INT[U])  # wraps modulo 2^(8U), as the model functors do
    out = _out(n, INT[U])
    S.model_with_intervals(out, _dev(coefficients), period, n, uncompressed_size=U, model=model, index_size=I, library=library)
    assert (_result(out, INT[U], n) == expected).all()


@pytest.mark.gpu
@pytest.mark.parametrize("I", [4, 8])
@pytest.mark.parametrize("R,D", [(2, 1), (4, 1), (4, 2), (8, 1), (8, 2), (8, 4)])
@pytest.mark.parametrize("n", [1, 1000, 100000 + 3, (1 << 21) + 77])
@pytest.mark.parametrize("with_initial", [False, True])
def test_inclusive_scan(library, I, R, D, n, with_initial):
    r = cases.rng(53, I, R, D, n)
    data = cases.random_uint(r, D, n)
    initial = int(cases.random_uint(r, R, 1)[0]) if with_initial else 0
    expected = (np.cumsum(data.astype(np.uint64), dtype=np.uint64) + np.uint64(initial)).astype(UINT[R])
    out = _out(n, UINT[R])
    S.inclusive_scan(out, _dev(data), n, result_size=R, input_size=D, index_size=I, library=library,
                     initial_value=_dev(np.array([initial], dtype=UINT[R])) if with_initial else None)
    assert (_result(out, UINT[R], n) == expected).all()
