"""Parity at BASELINE.json's full sizes (SURVEY.md 8d shapes), on the GPU, through the registry C ABI.

The oracle cannot decode 2^30-element columns in test time, so at these sizes the expected column is computed ON THE
DEVICE with plain torch operations (repeat_interleave, cumsum, gather, index_put) -- independent of both the library's
kernels and the oracle -- by benchmarks/configs.py, which is also what bench.py times; every row it returns carries
`bit_exact`.  The small-size twins of all these shapes are compared with the oracle in tests/test_gpu_parity.py.

Also here: what only shows at these sizes -- 64-bit index types with more than 2^32 elements, and a DZB-variable byte
stream longer than 4 GiB (byte offsets beyond 32 bits, IndexSize = 8 anchors).
"""
import pytest

pytestmark = pytest.mark.gpu


def _rows(name, library):
    import torch
    from benchmarks import configs as C
    torch.cuda.set_device(0)
    rows = C.run_configs([name], C.Options(repeats=1), torch.device("cuda", 0), library, None)
    torch.cuda.empty_cache()
    assert rows, name
    return rows


@pytest.mark.parametrize("name", ["cfg1", "cfg2", "cfg3", "cfg4", "cfg5"])
def test_baseline_configuration_is_bit_exact_at_full_size(name, library):
    rows = _rows(name, library)
    wrong = [r["chain"] for r in rows if not r["bit_exact"]]
    assert not wrong, "%s: not bit-exact at full size: %s" % (name, wrong)
    chains = " | ".join(r["chain"] for r in rows)
    if name == "cfg3":
        assert "p=5368709" in chains  # exactly floor(0.01 * 2^29) exceptions
        assert "fused patched Dictionary" in chains and "two kernels" in chains
    if name == "cfg4":
        assert "fused Delta(DZB-variable)" in chains and "two kernels" in chains
    if name == "cfg5":
        assert "A: fused patched FOR" in chains and "B: RPE" in chains and "C: incidence bitmaps" in chains and "host-resident scan" in chains


def _compare_in_pieces(torch, out, expected_of, n, piece=1 << 28):
    for a in range(0, n, piece):
        b = min(a + piece, n)
        if not torch.equal(out[a:b], expected_of(a, b)):
            return a
    return None


def test_dictionary_with_more_than_2_to_32_elements(library):
    """decompression::dictionary::kernel_t<8u, 1u, 1u>: positions and the length need 64 bits."""
    import torch
    from libgiddy_b200 import schemes as S
    n = (1 << 32) + 4099
    gen = torch.Generator(device="cuda")
    gen.manual_seed(41)
    indices = torch.randint(0, 256, (n,), generator=gen, device="cuda", dtype=torch.uint8)
    dictionary = torch.randperm(256, generator=gen, device="cuda").to(torch.uint8)
    out = torch.full((n + 64,), 0x5A, dtype=torch.uint8, device="cuda")
    S.dictionary(out, indices, dictionary, n, 256, uncompressed_size=1, dictionary_index_size=1, index_size=8, library=library)
    torch.cuda.synchronize()
    assert _compare_in_pieces(torch, out, lambda a, b: dictionary[indices[a:b].to(torch.int64)], n) is None
    assert bool((out[n:] == 0x5A).all())


def test_discard_zero_bytes_fixed_with_more_than_2_to_32_elements(library):
    """decompression::discard_zero_bytes::fixed_width::kernel_t<8u, 2u, 1u, ...>: an 8 GiB output."""
    import torch
    from libgiddy_b200 import schemes as S
    n = (1 << 32) + 4099
    gen = torch.Generator(device="cuda")
    gen.manual_seed(42)
    compressed = torch.randint(0, 256, (n,), generator=gen, device="cuda", dtype=torch.uint8)
    out = torch.full((n + 64,), 0x5A5A, dtype=torch.int16, device="cuda")
    S.discard_zero_bytes_fixed(out, compressed, n, uncompressed_size=2, compressed_size=1, index_size=8, library=library)
    torch.cuda.synchronize()
    assert _compare_in_pieces(torch, out, lambda a, b: compressed[a:b].to(torch.int16), n) is None
    assert bool((out[n:] == 0x5A5A).all())


@pytest.mark.parametrize("fused", [False, True], ids=["dzbv", "delta-of-dzbv-fused"])
def test_variable_width_stream_longer_than_4_gib(fused, library):
    """decompression::discard_zero_bytes::variable_width::kernel_t<8u, 4u, 4u> (and the fused Delta over it, <8u, 8u, 4u, 4u>):
    byte anchors beyond 2^32.  A random block of 2^26 elements is laid down 38 times (4.4 GB of stream); every copy
    must decode to the block (for Delta: to the block's prefix sums on top of that copy's anchors)."""
    import torch
    from benchmarks import configs as C
    from libgiddy_b200 import schemes as S
    block, L, copies = 1 << 26, 1024, 38
    gen = torch.Generator(device="cuda")
    gen.manual_seed(43)
    deltas, stream, packed, anchors, total = C.synthesize_dzbv_deltas(torch.device("cuda"), gen, block, L)
    assert copies * total > (1 << 32)
    n = copies * block
    whole_stream = torch.cat((stream[:total].repeat(copies), torch.zeros(16, dtype=torch.uint8, device="cuda")))
    whole_packed = packed.repeat(copies)
    block_anchors = anchors.to(torch.int64) & 0xFFFFFFFF
    whole_anchors = (block_anchors.view(1, -1) + (torch.arange(copies, device="cuda", dtype=torch.int64) * total).view(-1, 1)).reshape(-1)
    del stream, packed
    if not fused:
        out = torch.full((n + 64,), 0x5A5A5A5A, dtype=torch.int32, device="cuda")
        S.discard_zero_bytes_variable(out, whole_stream, whole_packed, whole_anchors, L, n, 1, 2, uncompressed_size=4, container_size=4, index_size=8,
                                      library=library)
        torch.cuda.synchronize()
        expected = deltas.to(torch.int32)  # (values below 2^32: the int32 bit pattern)
        for k in range(copies):
            assert torch.equal(out[k * block: (k + 1) * block], expected), "copy %d of the block decodes differently" % k
        assert bool((out[n:] == 0x5A5A5A5A).all())
    else:
        sums = torch.cumsum(deltas, 0)
        segment_carry = torch.cat((torch.zeros(1, dtype=torch.int64, device="cuda"), sums[L - 1: block - 1: L]))  # sum of the block's deltas before each segment
        starts = torch.arange(copies, device="cuda", dtype=torch.int64) * 1000003 + 17
        anchor_values = (starts.view(-1, 1) + segment_carry.view(1, -1)).reshape(-1)
        out = torch.full((n + 64,), 0x5A5A5A5A5A5A5A5A, dtype=torch.int64, device="cuda")
        S.delta_of_discard_zero_bytes_variable(out, whole_stream, whole_packed, whole_anchors, anchor_values, L, n, 1, 2,
                                               uncompressed_size=8, compressed_size=4, container_size=4, index_size=8, library=library)
        torch.cuda.synchronize()
        for k in range(copies):
            assert torch.equal(out[k * block: (k + 1) * block], sums + starts[k]), "copy %d of the block decodes differently" % k
        assert bool((out[n:] == 0x5A5A5A5A5A5A5A5A).all())
