#!/usr/bin/env python
"""benchmarks/configs.py -- the five BASELINE.json configurations at their named shapes, on synthetic columns
generated on the device from fixed seeds.  Needs a GPU; under torchrun every rank decodes its own columns (weak
scaling, no data-path collective) and rank 0 reports the max-over-ranks time.

    python benchmarks/configs.py [--only cfg1,cfg4] [--shrink K] [--out gpurun_out/configs.json] [--no-reference]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 benchmarks/configs.py --only cfg4

Per configuration and per kernel chain:
  decompressed GB/s   output bytes / time
  algorithmic GB/s    (compressed + side data read + decompressed written) / time, and its fraction of the measured
                      HBM copy bandwidth (MEASURED_PEAKS.json)
  reference           the same chain through eyalroz/libgiddy's own kernels rebuilt for sm_100a (oracle/_ref), when built
Every output is verified bit-for-bit on the device against an expansion written with plain torch ops before it is
timed.  Timing: CUDA events around `--repeats` back-to-back passes after 3 warm-ups; all working sets >> L2.
--shrink K divides every element count by 2^K (for quick runs).
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import libgiddy_b200 as g  # noqa: E402
from libgiddy_b200 import schemes as S  # noqa: E402

WORLD = int(os.environ.get("WORLD_SIZE", "1"))
RANK = int(os.environ.get("RANK", "0"))
LOCAL_RANK = int(os.environ.get("LOCAL_RANK", "0"))


LOG_FILE = None  # where progress lines go (None: stdout); bench.py sends them to stderr -- its stdout carries one JSON line


def log(text):
    print(text, file=LOG_FILE or sys.stdout, flush=True)


def generator(device, seed):
    gen = torch.Generator(device=device)
    gen.manual_seed(seed + 1000 * RANK)
    return gen


def nbytes(*tensors):
    return sum(t.numel() * t.element_size() for t in tensors)


def barrier():
    if WORLD > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()


def time_passes(run, repeats):
    """Seconds per pass: CUDA events around `repeats` passes, max over ranks."""
    for _ in range(3):
        run()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    start.record()
    for _ in range(repeats):
        run()
    end.record()
    barrier()
    ms = start.elapsed_time(end) / repeats
    if WORLD > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms = float(t.item())
    return ms / 1e3


def distinct_sorted_positions(device, gen, n, p):
    """Exactly p distinct positions in [0, n), ascending: draw a few too many, drop duplicates, then drop a random surplus."""
    if p == 0:
        return torch.empty(0, dtype=torch.int64, device=device)
    drawn = torch.unique(torch.randint(0, n, (p + p // 50 + 1024,), generator=gen, device=device, dtype=torch.int64))
    while drawn.numel() < p:  # (only for tiny n)
        drawn = torch.unique(torch.cat((drawn, torch.randint(0, n, (p,), generator=gen, device=device, dtype=torch.int64))))
    keep = torch.randperm(drawn.numel(), generator=gen, device=device)[:p]
    return drawn[torch.sort(keep).values]


def report(rows, peak, config, chain, output_bytes, algorithmic_bytes, seconds, verified, reference_seconds=None, extra=None):
    # (peak None: a row which is not bound by HBM -- the host-resident scan is bound by the host link -- has no HBM fraction)
    row = {"config": config, "chain": chain, "n_gpus": WORLD, "ms": seconds * 1e3,
           "decompressed_gbs": WORLD * output_bytes / seconds / 1e9, "algorithmic_gbs": WORLD * algorithmic_bytes / seconds / 1e9,
           "frac_of_measured_peak_per_gpu": (algorithmic_bytes / seconds / 1e9 / peak) if peak else None, "output_bytes_per_gpu": output_bytes,
           "algorithmic_bytes_per_gpu": algorithmic_bytes, "bit_exact": verified}
    if reference_seconds is not None:
        row.update({"reference_kernels_ms": reference_seconds * 1e3, "reference_kernels_decompressed_gbs": WORLD * output_bytes / reference_seconds / 1e9,
                    "speedup_vs_reference_kernels": reference_seconds / seconds})
    if extra:
        row.update(extra)
    rows.append(row)
    if RANK == 0:
        log("%-5s %-44s %9.3f ms  out %8.1f GB/s  alg %8.1f GB/s (%5.1f%% of %.0f per GPU)  exact=%s  ref %s" % (
            config, chain, row["ms"], row["decompressed_gbs"], row["algorithmic_gbs"], 100 * (row["frac_of_measured_peak_per_gpu"] or 0), peak or 0, verified,
            ("%.3f ms (x%.2f)" % (reference_seconds * 1e3, reference_seconds / seconds)) if reference_seconds else "-"))


def with_reference(reference, run_with, repeats):
    if reference is None:
        return None
    try:
        return time_passes(lambda: run_with(reference), max(3, repeats // 4))
    except Exception as error:  # the reference is not defined on every shape
        if RANK == 0:
            log("      (reference kernels failed: %s)" % str(error)[:120])
        return None


# ---------------------------------------------------------------------------------------------------------------
# cfg1: Frame of Reference int32 <- int16, constant model, L = 4096
# ---------------------------------------------------------------------------------------------------------------

def cfg1(args, device, product, reference, rows, peak):
    n, L = (1 << 26) >> args.shrink, 4096
    S_ = n // L
    gen = generator(device, 1)
    coefficients = torch.randint(-(1 << 30), 1 << 30, (S_, 1), generator=gen, device=device, dtype=torch.int64).to(torch.int32)
    offsets = torch.randint(-32768, 32768, (n,), generator=gen, device=device, dtype=torch.int32).to(torch.int16)
    out = torch.empty(n, dtype=torch.int32, device=device)

    def run(library=product):
        S.frame_of_reference(out, offsets, coefficients, n, L, uncompressed_size=4, compressed_size=2, model="constant", library=library)
    run()
    expected = coefficients[:, 0].repeat_interleave(L) + offsets.to(torch.int32)
    verified = bool(torch.equal(out, expected))
    del expected
    seconds = time_passes(run, args.repeats)
    report(rows, peak, "cfg1", "FOR int32<-int16 constant L=4096 (n=2^%d)" % (26 - args.shrink), n * 4, nbytes(coefficients, offsets) + n * 4,
           seconds, verified, with_reference(reference, run, args.repeats))


# ---------------------------------------------------------------------------------------------------------------
# cfg2: RLE uint32 values + uint16 run lengths, mean run 16, L = 1024
# ---------------------------------------------------------------------------------------------------------------

def synthesize_runs(device, gen, n, max_run, L, length_dtype):
    estimate = int(n / ((1 + max_run) / 2) * 1.02) + 4096
    lengths = torch.randint(1, max_run + 1, (estimate,), generator=gen, device=device, dtype=torch.int64)
    ends = torch.cumsum(lengths, 0)
    num_runs = int(torch.searchsorted(ends, torch.tensor([n], device=device, dtype=torch.int64)).item()) + 1
    assert num_runs <= estimate
    lengths, ends = lengths[:num_runs].clone(), ends[:num_runs].clone()
    lengths[-1] -= ends[-1] - n
    starts = ends - lengths
    starts[-1] = n - lengths[-1]
    values = torch.randint(-(1 << 31), 1 << 31, (num_runs,), generator=gen, device=device, dtype=torch.int64).to(torch.int32)
    segment_starts = torch.arange((n + L - 1) // L, device=device, dtype=torch.int64) * L
    anchors = torch.searchsorted(starts, segment_starts, right=True) - 1
    intra = segment_starts - starts[anchors]
    return values, lengths, starts, anchors.to(torch.int32), intra.to(length_dtype), num_runs


def cfg2(args, device, product, reference, rows, peak):
    n, L = (1 << 30) >> args.shrink, 1024
    gen = generator(device, 3)
    values, lengths64, _, anchors, intra, R = synthesize_runs(device, gen, n, 31, L, torch.int16)
    lengths = lengths64.to(torch.int16)
    out = torch.empty(n + 64, dtype=torch.int32, device=device)

    def run(library=product):
        S.run_length_encoding(out, values, lengths, anchors, intra, L, (n + L - 1) // L, R, n, uncompressed_size=4, run_length_size=2, library=library)
    run()
    expected = torch.repeat_interleave(values, lengths64)
    verified = bool(torch.equal(out[:n], expected))
    del expected, lengths64
    seconds = time_passes(run, args.repeats)
    report(rows, peak, "cfg2", "RLE u32 values + u16 lengths, run 16, L=1024 (n=2^%d)" % (30 - args.shrink), n * 4,
           nbytes(values, lengths, anchors, intra) + n * 4, seconds, verified, with_reference(reference, run, args.repeats))


# ---------------------------------------------------------------------------------------------------------------
# cfg3: Dictionary uint8 -> int64 (256 entries) + naive patch scatter, 1 % exceptions
# ---------------------------------------------------------------------------------------------------------------

def cfg3(args, device, product, reference, rows, peak):
    n = (1 << 29) >> args.shrink
    gen = generator(device, 5)
    dictionary = torch.randint(-(1 << 62), 1 << 62, (256,), generator=gen, device=device, dtype=torch.int64)
    columns_module = __import__("libgiddy_b200.columns", fromlist=["broadcast_dictionary"])
    columns_module.broadcast_dictionary(dictionary, 0)  # the one collective of the multi-GPU path (NCCL over NVLink)
    indices = torch.randint(0, 256, (n,), generator=gen, device=device, dtype=torch.int16).to(torch.uint8)
    p = n // 100  # exactly floor(0.01 n) distinct sorted positions (5 368 709 at full size)
    positions = distinct_sorted_positions(device, gen, n, p)
    patches = torch.randint(-(1 << 62), 1 << 62, (p,), generator=gen, device=device, dtype=torch.int64) | 1
    positions32 = positions.to(torch.int32)
    out = torch.empty(n, dtype=torch.int64, device=device)

    def decode(library=product):
        S.dictionary(out, indices, dictionary, n, 256, uncompressed_size=8, dictionary_index_size=1, library=library)

    def patch(library=product):
        S.scatter(out, patches, positions32, p, element_size=8, output_index_size=4, input_index_size=4, library=library)

    def both(library=product):
        decode(library)
        patch(library)
    both()
    expected = dictionary[indices.to(torch.int64)]
    expected[positions] = patches
    verified = bool(torch.equal(out, expected))
    del expected
    name = "(n=2^%d, p=%d)" % (29 - args.shrink, p)
    decode_bytes = nbytes(indices, dictionary) + n * 8
    patch_bytes = nbytes(patches, positions32) + p * 8
    seconds = time_passes(decode, args.repeats)
    report(rows, peak, "cfg3", "Dictionary u8->i64 D=256 " + name, n * 8, decode_bytes, seconds, verified, with_reference(reference, decode, args.repeats))
    seconds = time_passes(patch, args.repeats)
    report(rows, peak, "cfg3", "scatter 1% i64 patches (target hot in L2 or not)", p * 8, patch_bytes, seconds, verified, with_reference(reference, patch, args.repeats))
    seconds = time_passes(both, args.repeats)
    report(rows, peak, "cfg3", "Dictionary + scatter, two kernels " + name, n * 8, decode_bytes + patch_bytes, seconds, verified,
           with_reference(reference, both, args.repeats))

    # extension: the same column decoded and patched in one pass
    def fused(limits=None):
        S.patched_dictionary(out, indices, dictionary, n, 256, patches, positions32, p, uncompressed_size=8, dictionary_index_size=1, limits=limits,
                             library=product)
    out.zero_()
    fused()
    expected = dictionary[indices.to(torch.int64)]
    expected[positions] = patches
    verified_fused = bool(torch.equal(out, expected))
    del expected
    fused_bytes = decode_bytes + nbytes(patches, positions32)
    seconds = time_passes(fused, args.repeats)
    report(rows, peak, "cfg3", "fused patched Dictionary, one kernel " + name, n * 8, fused_bytes, seconds, verified_fused, extra={"extension": True})
    if args.sweep_grids:
        sms = torch.cuda.get_device_properties(device).multi_processor_count
        for per_sm in (4, 8, 16, 64, 128):
            seconds = time_passes(lambda: fused({"grid_size": sms * per_sm}), args.repeats)
            report(rows, peak, "cfg3", "fused patched Dictionary, %d CTAs/SM" % per_sm, n * 8, fused_bytes, seconds, verified_fused, extra={"extension": True})


# ---------------------------------------------------------------------------------------------------------------
# cfg4: Delta int64 (anchored) over DZB-variable deltas (1-4 bytes, 2-bit size fields), L = 1024
# ---------------------------------------------------------------------------------------------------------------

def synthesize_dzbv_deltas(device, gen, n, L):
    """widths {1,2,3,4} w.p. {.50,.30,.15,.05}; value uniform in [2^(8(w-1)), 2^(8w)) (w = 1: [0, 256))."""
    u = torch.rand(n, generator=gen, device=device)
    widths = (1 + (u >= 0.50).to(torch.int64) + (u >= 0.80).to(torch.int64) + (u >= 0.95).to(torch.int64))
    del u
    low = torch.where(widths > 1, torch.ones_like(widths) << (8 * (widths - 1)), torch.zeros_like(widths))
    high = torch.ones_like(widths) << (8 * widths)
    deltas = low + (torch.rand(n, generator=gen, device=device, dtype=torch.float64) * (high - low).to(torch.float64)).to(torch.int64)
    deltas = torch.minimum(deltas, high - 1)
    del low, high
    ends = torch.cumsum(widths, 0)
    offsets = ends - widths
    total = int(ends[-1].item())
    del ends
    stream = torch.zeros(total + 16, dtype=torch.uint8, device=device)
    for k in range(4):
        mask = widths > k
        stream[offsets[mask] + k] = ((deltas[mask] >> (8 * k)) & 0xFF).to(torch.uint8)
        del mask
    shifts = 2 * torch.arange(16, device=device, dtype=torch.int64)
    packed = (((widths - 1).view(-1, 16)) << shifts).sum(1).to(torch.int32)
    position_anchors = offsets[::L].to(torch.int32)  # (byte offsets < 2^32, stored in the uint32 bit pattern)
    del offsets, widths
    return deltas, stream, packed, position_anchors, total


def cfg4(args, device, product, reference, rows, peak):
    n, L = (1 << 30) >> args.shrink, 1024
    columns_per_gpu = max(1, args.cfg4_columns // WORLD) if WORLD > 1 else 1
    gen = generator(device, 8)
    deltas, stream, packed, position_anchors, total = synthesize_dzbv_deltas(device, gen, n, L)
    start_value = int(torch.randint(0, 1 << 40, (1,), generator=gen, device=device).item())
    expected = torch.cumsum(deltas, 0) + start_value
    anchor_values = torch.cat((torch.tensor([start_value], device=device, dtype=torch.int64), expected[L - 1:n - 1:L]))
    del deltas
    widened = torch.empty(n, dtype=torch.int32, device=device)
    out = torch.empty(n, dtype=torch.int64, device=device)

    def dzbv(library=product):
        S.discard_zero_bytes_variable(widened, stream, packed, position_anchors, L, n, 1, 2, uncompressed_size=4, container_size=4, library=library)

    def delta(library=product):
        S.delta(out, widened, anchor_values, n, L, uncompressed_size=8, compressed_size=4, library=library)

    def two_kernels(library=product):
        for _ in range(columns_per_gpu):
            dzbv(library)
            delta(library)

    def fused(library=product):
        for _ in range(columns_per_gpu):
            S.delta_of_discard_zero_bytes_variable(out, stream, packed, position_anchors, anchor_values, L, n, 1, 2,
                                                   uncompressed_size=8, compressed_size=4, container_size=4, library=library)
    two_kernels()
    verified_two = bool(torch.equal(out, expected))
    out.zero_()
    fused()
    verified_fused = bool(torch.equal(out, expected))
    del expected
    compressed = total + nbytes(packed, position_anchors)
    name = "(n=2^%d x %d column(s) per GPU, %.3f B/delta)" % (30 - args.shrink, columns_per_gpu, total / n)
    c = columns_per_gpu
    seconds = time_passes(dzbv, args.repeats)
    report(rows, peak, "cfg4", "DZB-variable u32, 2-bit sizes, L=1024 (1 column)", n * 4, compressed + n * 4, seconds, verified_two,
           with_reference(reference, dzbv, args.repeats))
    seconds = time_passes(delta, args.repeats)
    report(rows, peak, "cfg4", "Delta u64<-u32, L=1024 (1 column)", n * 8, n * 4 + nbytes(anchor_values) + n * 8, seconds, verified_two,
           with_reference(reference, delta, args.repeats))
    seconds = time_passes(two_kernels, args.repeats)
    report(rows, peak, "cfg4", "DZB-variable -> Delta, two kernels " + name, c * n * 8,
           c * (compressed + n * 4 + n * 4 + nbytes(anchor_values) + n * 8), seconds, verified_two, with_reference(reference, two_kernels, args.repeats))
    seconds = time_passes(fused, args.repeats)
    report(rows, peak, "cfg4", "fused Delta(DZB-variable), one kernel " + name, c * n * 8, c * (compressed + nbytes(anchor_values) + n * 8), seconds,
           verified_fused, extra={"extension": True})


# ---------------------------------------------------------------------------------------------------------------
# cfg5: mixed columnar scan from pinned host memory: FOR(linear) + ci-scatter patches, RPE, incidence bitmaps
# ---------------------------------------------------------------------------------------------------------------

def synthesize_ci_patches(device, gen, n, density, A, run_length):
    p = int(n * density)
    positions = distinct_sorted_positions(device, gen, n, p)
    values = torch.randint(-(1 << 31), 1 << 31, (p,), generator=gen, device=device, dtype=torch.int64).to(torch.int32)
    assert A % run_length == 0  # so that fixed-length runs never straddle an anchoring period
    num_runs = (p + run_length - 1) // run_length
    run_first = torch.arange(num_runs, device=device, dtype=torch.int64) * run_length
    run_lengths = torch.clamp(p - run_first, max=run_length)
    baselines = positions[run_first]
    run_of_patch = torch.arange(p, device=device, dtype=torch.int64) // run_length
    offsets = positions - baselines[run_of_patch]
    max_offset = torch.zeros(num_runs, device=device, dtype=torch.int64).scatter_reduce(0, run_of_patch, offsets, "amax")
    width_bytes = torch.where(max_offset < 256, 1, torch.where(max_offset < 65536, 2, 4)).to(torch.int64)
    run_bytes = (run_lengths * width_bytes + 3) // 4 * 4  # every run's offsets begin on an OutputIndexSize (4-byte) boundary
    run_byte_start = torch.cumsum(run_bytes, 0) - run_bytes
    total = int(run_bytes.sum().item())
    offset_bytes = torch.zeros(total + 16, dtype=torch.uint8, device=device)
    wb = width_bytes[run_of_patch]
    index_in_run = torch.arange(p, device=device, dtype=torch.int64) - run_first[run_of_patch]
    address = run_byte_start[run_of_patch] + index_in_run * wb
    for k in range(4):
        mask = wb > k
        offset_bytes[address[mask] + k] = ((offsets[mask] >> (8 * k)) & 0xFF).to(torch.uint8)
    anchors = torch.arange((p + A - 1) // A, device=device, dtype=torch.int64) * (A // run_length)
    arrays = {"run_lengths": run_lengths.to(torch.int32), "offset_sizes": (8 * width_bytes).to(torch.uint8), "baselines": baselines.to(torch.int32),
              "offsets_start": (run_byte_start // 4).to(torch.int32), "offset_bytes": offset_bytes, "anchors": anchors.to(torch.int32)}
    return positions, values, arrays, num_runs, p


def cfg5(args, device, product, reference, rows, peak):
    n, L, A, K = (1 << 28) >> args.shrink, 1024, 1024, 16
    S_ = n // L
    # (A) FOR, linear model, int32 <- int16, with 1 % exceptions patched by a compressed-indices scatter
    gen = generator(device, 20)
    coefficients = torch.randint(-(1 << 30), 1 << 30, (S_, 2), generator=gen, device=device, dtype=torch.int64).to(torch.int32)
    coefficients[:, 1] = torch.randint(-1000, 1000, (S_,), generator=gen, device=device, dtype=torch.int32)
    offsets = torch.randint(-32768, 32768, (n,), generator=gen, device=device, dtype=torch.int32).to(torch.int16)
    positions, patch_values, ci, num_runs, p = synthesize_ci_patches(device, gen, n, 0.01, A, 64)
    out_a = torch.empty(n, dtype=torch.int32, device=device)

    def for_linear(library=product):
        S.frame_of_reference(out_a, offsets, coefficients, n, L, uncompressed_size=4, compressed_size=2, model="linear", library=library)

    def ci_patch(library=product):
        S.compressed_indices_scatter(out_a, patch_values, ci["run_lengths"], ci["offset_sizes"], ci["baselines"], ci["offsets_start"],
                                     ci["offset_bytes"], ci["anchors"], A, num_runs, p, element_size=4, output_index_size=4, input_index_size=4,
                                     library=library)

    def column_a(library=product):
        for_linear(library)
        ci_patch(library)
    column_a()
    x = torch.arange(L, device=device, dtype=torch.int32).repeat(S_)
    expected = coefficients[:, 0].repeat_interleave(L) + coefficients[:, 1].repeat_interleave(L) * x + offsets.to(torch.int32)
    del x
    expected[positions] = patch_values
    verified_a = bool(torch.equal(out_a, expected))
    del expected
    bytes_a = nbytes(coefficients, offsets, patch_values, *ci.values()) - 16
    seconds = time_passes(column_a, args.repeats)
    report(rows, peak, "cfg5", "A: FOR linear i32<-i16 + ci-scatter 1%% (n=2^%d, resident)" % (28 - args.shrink), n * 4, bytes_a + n * 4 + p * 4, seconds, verified_a,
           with_reference(reference, column_a, args.repeats))
    seconds = time_passes(ci_patch, args.repeats)
    report(rows, peak, "cfg5", "A: ci-scatter alone (p=%d)" % p, p * 4, nbytes(patch_values, *ci.values()) - 16 + p * 4, seconds, verified_a,
           with_reference(reference, ci_patch, args.repeats))

    # extension: column A decoded and patched in one pass
    def fused_a(limits=None):
        S.patched_frame_of_reference(out_a, offsets, coefficients, n, L, patch_values, ci["run_lengths"], ci["offset_sizes"], ci["baselines"],
                                     ci["offsets_start"], ci["offset_bytes"], ci["anchors"], A, num_runs, p,
                                     uncompressed_size=4, compressed_size=2, model="linear", limits=limits, library=product)
    out_a.zero_()
    fused_a()
    x = torch.arange(L, device=device, dtype=torch.int32).repeat(S_)
    expected = coefficients[:, 0].repeat_interleave(L) + coefficients[:, 1].repeat_interleave(L) * x + offsets.to(torch.int32)
    del x
    expected[positions] = patch_values
    verified_fused = bool(torch.equal(out_a, expected))
    del expected
    seconds = time_passes(fused_a, args.repeats)
    report(rows, peak, "cfg5", "A: fused patched FOR (ci patches), one kernel", n * 4, bytes_a + n * 4, seconds, verified_fused, extra={"extension": True})
    if args.sweep_grids:
        sms = torch.cuda.get_device_properties(device).multi_processor_count
        for per_sm in (4, 8, 16, 64, 128):
            seconds = time_passes(lambda: fused_a({"grid_size": sms * per_sm}), args.repeats)
            report(rows, peak, "cfg5", "A: fused patched FOR, %d CTAs/SM" % per_sm, n * 4, bytes_a + n * 4, seconds, verified_fused, extra={"extension": True})

    # (B) RPE, absolute positions, mean run 64
    gen = generator(device, 23)
    values, lengths64, starts, anchors, _, R = synthesize_runs(device, gen, n, 127, L, torch.int16)
    starts32 = starts.to(torch.int32)
    out_b = torch.empty(n + 64, dtype=torch.int32, device=device)

    def rpe(library=product):
        S.run_position_encoding(out_b, values, starts32, anchors, L, S_, R, n, uncompressed_size=4, position_offset_size=4, relative=False, library=library)
    rpe()
    expected = torch.repeat_interleave(values, lengths64)
    verified_b = bool(torch.equal(out_b[:n], expected))
    del expected, lengths64, starts
    bytes_b = nbytes(values, starts32, anchors)
    seconds = time_passes(rpe, args.repeats)
    report(rows, peak, "cfg5", "B: RPE u32 absolute positions, run 64 (resident)", n * 4, bytes_b + n * 4, seconds, verified_b, with_reference(reference, rpe, args.repeats))

    # (C) incidence bitmaps, K = 16, skewed value distribution
    gen = generator(device, 26)
    index = (torch.rand(n, generator=gen, device=device) ** 2 * K).to(torch.int64).clamp_(max=K - 1)
    dictionary = torch.randint(-(1 << 31), 1 << 31, (K,), generator=gen, device=device, dtype=torch.int64).to(torch.int32)
    bitmap_length = n // 32
    weights = torch.ones(32, device=device, dtype=torch.int64) << torch.arange(32, device=device, dtype=torch.int64)
    bitmaps = torch.empty((K, bitmap_length), dtype=torch.int32, device=device)
    for b in range(K):
        bitmaps[b] = ((index == b).view(-1, 32).to(torch.int64) * weights).sum(1).to(torch.int32)
    out_c = torch.empty(n, dtype=torch.int32, device=device)

    def bitmaps_decode(library=product):
        S.incidence_bitmaps(out_c, bitmaps, dictionary, n, bitmap_length, K, uncompressed_size=4, alignment_in_ints=2, use_dictionary=True, library=library)
    bitmaps_decode()
    verified_c = bool(torch.equal(out_c, dictionary[index]))
    del index
    bytes_c = nbytes(bitmaps, dictionary)
    seconds = time_passes(bitmaps_decode, args.repeats)
    report(rows, peak, "cfg5", "C: incidence bitmaps u32, K=16 (resident)", n * 4, bytes_c + n * 4, seconds, verified_c,
           with_reference(reference, bitmaps_decode, args.repeats))

    # host-resident mode: the three compressed columns live in pinned host memory; segment-range chunks are copied on side
    # streams and decoded as they land.  Compared with copying the uncompressed columns.
    host_resident_scan(args, device, product, rows, n, L, A, K, S_, p, num_runs, R, bitmap_length,
                       {"coefficients": coefficients, "offsets": offsets, "patch_values": patch_values, **{"ci_" + k: v for k, v in ci.items()},
                        "positions": positions},
                       {"values": values, "starts": starts32, "anchors": anchors}, {"bitmaps": bitmaps, "dictionary": dictionary},
                       (out_a, out_b, out_c), bytes_a + bytes_b + bytes_c)


def host_resident_scan(args, device, product, rows, n, L, A, K, S_, p, num_runs, R, bitmap_length, col_a, col_b, col_c, outs, compressed_bytes):
    """The three compressed columns live in pinned host memory.  A pass copies them H2D in contiguous segment-range
    chunks on side streams and decodes each chunk as soon as it has landed; nothing is device-resident beforehand."""
    import numpy as np
    from libgiddy_b200 import columns
    out_a, out_b, out_c = outs
    num_chunks, num_streams = args.chunks, args.streams
    streams = [torch.cuda.Stream(device=device) for _ in range(num_streams)]
    runs_per_period = A // 64

    def pin(t):
        return t.cpu().pin_memory()

    positions = col_a.pop("positions")
    host_a, dev_a = {k: pin(v) for k, v in col_a.items()}, {k: torch.empty_like(v) for k, v in col_a.items()}
    host_b, dev_b = {k: pin(v) for k, v in col_b.items()}, {k: torch.empty_like(v) for k, v in col_b.items()}
    host_c = {k: pin(v) for k, v in col_c.items()}
    dev_dictionary = torch.empty_like(col_c["dictionary"])
    kernel_c = product.produce_subclass(S.incidence_bitmaps_key(4, 4, 2, True))

    # ---- planning (once per column, untimed, host side) ---------------------------------------------------------
    # A: a period of A patches is applied by the chunk which holds its LAST target (targets are sorted), after the previous
    #    chunk's decode when it reaches back.  B: absolute run start positions and run-index anchors are rebased per chunk.
    #    C: containers [w0, w1) of all K bitmaps are gathered into a compact K x (w1 - w0) buffer (the kernel takes one pitch).
    num_periods = (p + A - 1) // A
    period_last = positions[torch.clamp(torch.arange(1, num_periods + 1, device=device) * A, max=p) - 1].cpu().numpy()
    anchors_np = host_b["anchors"].numpy().view(np.uint32)
    starts_np = host_b["starts"].numpy().view(np.uint32)
    plans, next_period = [], 0
    for (s0, s1) in columns.split_segments(S_, num_chunks):
        e0, e1 = s0 * L, s1 * L
        k0 = next_period
        while next_period < num_periods and period_last[next_period] < e1:
            next_period += 1
        k1 = next_period
        shard = columns.RunLengthShard(anchors_np, s0, s1, L, n, R)
        r0, r1 = shard.first_run, shard.last_run + 1
        rebased_starts = starts_np[r0:r1].astype(np.int64) - e0
        rebased_starts[rebased_starts < 0] = 0
        u0, u1 = k0 * runs_per_period, min(k1 * runs_per_period, num_runs)
        w0, w1 = e0 // 32, e1 // 32
        plan = {"s": (s0, s1), "e": (e0, e1), "k": (k0, k1), "r": (r0, r1), "u": (u0, u1), "q": (k0 * A, min(k1 * A, p)), "w": (w0, w1),
                "host_starts": torch.from_numpy(rebased_starts.astype(np.uint32).view(np.int32)).pin_memory(),
                "host_anchors_b": torch.from_numpy((anchors_np[s0:s1].astype(np.int64) - r0).astype(np.uint32).view(np.int32)).pin_memory(),
                "host_anchors_a": (host_a["ci_anchors"][k0:k1].to(torch.int64) - u0).to(torch.int32).pin_memory(),
                "dev_anchors_a": torch.empty(k1 - k0, dtype=torch.int32, device=device),
                "compact_c": torch.empty((K, w1 - w0), dtype=torch.int32, device=device),
                "host_compact_c": host_c["bitmaps"][:, w0:w1].contiguous().pin_memory(),  # chunk-major on the host: one contiguous copy
                "config_c": kernel_c.resolve_launch_configuration({"length": g.size_t(e1 - e0)})}
        if k1 > k0:
            plan["b"] = (int(host_a["ci_offsets_start"][u0]) * 4,
                         int(host_a["ci_offsets_start"][u1]) * 4 if u1 < num_runs else host_a["ci_offset_bytes"].numel())
        plans.append(plan)
    assert next_period == num_periods
    h2d_bytes = (sum(t.numel() * t.element_size() for d in (host_a, host_b, host_c) for t in d.values()) - host_b["starts"].numel() * 0)

    # ---- the plan, recorded once in the library's host-resident scan executor (include/giddy_b200.h group 4) --------
    # chunk c: side stream c % num_streams; a pass (`step`) is one giddy_scan_run(): every copy and launch is enqueued
    # by the library, in C++.
    scan = product.create_scan(num_streams)

    def copy(chunk, dst, src, a, b):
        if b > a:
            chunk.copy(dst[a:b], src[a:b])

    for index, plan in enumerate(plans):
        chunk = scan.chunk(index)
        (s0, s1), (e0, e1), (k0, k1), (r0, r1), (u0, u1), (q0, q1), (w0, w1) = (plan[x] for x in "sekruqw")
        # column A: Frame of Reference, then the patches whose periods end in this chunk
        copy(chunk, dev_a["offsets"], host_a["offsets"], e0, e1)
        copy(chunk, dev_a["coefficients"], host_a["coefficients"], s0, s1)
        S.frame_of_reference(out_a[e0:e1], dev_a["offsets"][e0:e1], dev_a["coefficients"][s0:s1], e1 - e0, L, uncompressed_size=4,
                             compressed_size=2, model="linear", stream=chunk, library=product)
        if k1 > k0:
            if index > 0:
                chunk.after(index - 1)  # the first period may reach back into the previous chunk
            copy(chunk, dev_a["patch_values"], host_a["patch_values"], q0, q1)
            for name in ("ci_run_lengths", "ci_offset_sizes", "ci_baselines", "ci_offsets_start"):
                copy(chunk, dev_a[name], host_a[name], u0, u1)
            copy(chunk, dev_a["ci_offset_bytes"], host_a["ci_offset_bytes"], *plan["b"])
            chunk.copy(plan["dev_anchors_a"], plan["host_anchors_a"])
            S.compressed_indices_scatter(out_a, dev_a["patch_values"][q0:q1], dev_a["ci_run_lengths"][u0:u1], dev_a["ci_offset_sizes"][u0:u1],
                                         dev_a["ci_baselines"][u0:u1], dev_a["ci_offsets_start"][u0:u1], dev_a["ci_offset_bytes"],
                                         plan["dev_anchors_a"], A, u1 - u0, q1 - q0, element_size=4, output_index_size=4,
                                         input_index_size=4, stream=chunk, library=product)
        # column B: Run-Position Encoding
        copy(chunk, dev_b["values"], host_b["values"], r0, r1)
        chunk.copy(dev_b["starts"][r0:r1], plan["host_starts"])
        chunk.copy(dev_b["anchors"][s0:s1], plan["host_anchors_b"])
        S.run_position_encoding(out_b[e0:], dev_b["values"][r0:r1], dev_b["starts"][r0:r1], dev_b["anchors"][s0:s1], L, s1 - s0, r1 - r0, e1 - e0,
                                uncompressed_size=4, position_offset_size=4, relative=False, stream=chunk, library=product)
        # column C: incidence bitmaps
        if index == 0:
            chunk.copy(dev_dictionary, host_c["dictionary"])
        elif index < num_streams:
            chunk.after(0)  # (the dictionary arrives with the first chunk)
        chunk.copy(plan["compact_c"], plan["host_compact_c"])
        scan.add_launch(index, kernel_c, plan["config_c"], {
            "decompressed": g.ptr(out_c[e0:e1], "u32"), "incidence_bitmaps": g.cptr(plan["compact_c"], "u32"),
            "dictionary_entries": g.cptr(dev_dictionary, "u32"), "bitmap_length": g.u32(w1 - w0), "num_bitmaps": g.u32(K)})
    if args.graph_replay:
        scan.set_graph_replay(True)

    def step():
        scan.run()

    resident_outputs = [o.clone() for o in (out_a, out_b[:n], out_c)]
    for o in (out_a, out_b, out_c):
        o.zero_()
    step()
    torch.cuda.synchronize()
    verified = all(bool(torch.equal(a, b)) for a, b in zip(resident_outputs, (out_a, out_b[:n], out_c)))
    del resident_outputs
    repeats = max(3, args.repeats // 4)
    seconds = time_passes(step, repeats)
    output_bytes = 3 * n * 4
    extra = {"h2d_bytes_per_pass": h2d_bytes, "h2d_gbs_per_gpu": h2d_bytes / seconds / 1e9, "chunks": num_chunks, "streams": num_streams,
             "executor": "giddy_scan_run (%d recorded operations%s)" % (scan.num_operations, ", replayed as a CUDA graph" if args.graph_replay else ""),
             "compression_ratio": output_bytes / h2d_bytes}
    # the alternative Giddy argues against (README.md:18-24): ship the uncompressed columns over the same link
    raw_host = torch.empty(n, dtype=torch.int32).pin_memory()
    raw_dev = torch.empty(n, dtype=torch.int32, device=device)

    def raw_copy():
        for _ in range(3):
            raw_dev.copy_(raw_host, non_blocking=True)
    raw_seconds = time_passes(raw_copy, repeats)
    extra.update({"uncompressed_copy_ms": raw_seconds * 1e3, "uncompressed_copy_gbs": WORLD * output_bytes / raw_seconds / 1e9,
                  "gain_over_uncompressed_copy": raw_seconds / seconds})
    report(rows, None, "cfg5", "host-resident scan: 3 columns, pinned H2D + decode overlap", output_bytes, h2d_bytes + output_bytes, seconds, verified,
           extra=extra)
    if RANK == 0:
        log("      host-resident: %.1f GB/s decompressed over %d GPU(s) vs %.1f GB/s copying the uncompressed columns: gain x%.2f "
            "(compression ratio %.2f, H2D %.1f GB/s per GPU)" % (WORLD * output_bytes / seconds / 1e9, WORLD, WORLD * output_bytes / raw_seconds / 1e9,
                                                              raw_seconds / seconds, output_bytes / h2d_bytes, h2d_bytes / seconds / 1e9))


# ---------------------------------------------------------------------------------------------------------------
# strong scaling: ONE column of cfg2's shape cut into contiguous segment ranges, one per GPU, by the library's C++ planner
# ---------------------------------------------------------------------------------------------------------------

def strong(args, device, product, reference, rows, peak):
    """The same 2^30-element RLE column on every rank (rank-independent seed); rank r decodes segments [s0, s1) of it as a
    self-contained column: its runs, its segments' anchors with the first run's index subtracted (giddy_shard_plan says
    which and by how much).  Total work is fixed as the number of GPUs grows; time is the max over ranks."""
    import numpy as np
    from libgiddy_b200 import _native as N
    from libgiddy_b200 import columns
    n, L = (1 << 30) >> args.shrink, 1024
    gen = torch.Generator(device=device)
    gen.manual_seed(3)
    values, lengths64, starts, anchors, intra, R = synthesize_runs(device, gen, n, 31, L, torch.int16)
    lengths = lengths64.to(torch.int16)
    num_segments = (n + L - 1) // L
    host_anchors = anchors.cpu().numpy().view(np.uint32)  # planning reads a host copy of the anchors: 4 bytes per segment
    layout = columns._layout(N.SCHEME_RUN_LENGTH_ENCODING, n, L, index_size=4, uncompressed_size=4, compressed_size=2, num_runs=R, anchors=host_anchors)
    s0, s1 = product.split_segments(num_segments, WORLD, RANK)
    shard = product.shard_plan(layout, s0, s1)
    runs, rebase = shard.slice("run_data"), shard.slice("position_anchors").rebase
    # the shard as a column of its own, in buffers of its own (on one box the slices are cut on the device; across boxes
    # they would be copied out of the host-resident column with giddy_slice_copy)
    shard_values = values[runs.first: runs.first + runs.count].clone()
    shard_lengths = lengths[runs.first: runs.first + runs.count].clone()
    shard_anchors = (anchors[s0:s1].to(torch.int64) - rebase).to(torch.int32)
    shard_intra = intra[s0:s1].clone()
    count = int(shard.num_elements)
    expected = torch.repeat_interleave(values, lengths64)[shard.first_element: shard.first_element + count].clone()
    del values, lengths, lengths64, starts, anchors, intra
    torch.cuda.empty_cache()
    out = torch.empty(count + 64, dtype=torch.int32, device=device)

    def run(library=product):
        S.run_length_encoding(out, shard_values, shard_lengths, shard_anchors, shard_intra, L, s1 - s0, int(shard.num_runs), count,
                              uncompressed_size=4, run_length_size=2, library=library)
    run()
    verified = bool(torch.equal(out[:count], expected))
    del expected
    if WORLD > 1:
        ok = torch.tensor([1 if verified else 0], device=device)
        torch.distributed.all_reduce(ok, op=torch.distributed.ReduceOp.MIN)
        verified = bool(ok.item())
    seconds = time_passes(run, args.repeats)
    shard_bytes = nbytes(shard_values, shard_lengths, shard_anchors, shard_intra) + count * 4
    # (report() multiplies per-GPU bytes by the world size: pass the per-GPU share of the one column)
    report(rows, peak, "strong", "ONE RLE column (n=2^%d) split into %d segment-range shards" % (30 - args.shrink, WORLD), n * 4 // WORLD, shard_bytes, seconds, verified,
           extra={"scaling": "strong", "whole_column_decompressed_gbs": n * 4 / seconds / 1e9, "segments_per_shard": s1 - s0, "planner": "giddy_shard_plan"})


CONFIGS = {"cfg1": cfg1, "cfg2": cfg2, "cfg3": cfg3, "cfg4": cfg4, "cfg5": cfg5, "strong": strong}


class Options:
    """The knobs of a run (the command line's, or a caller's: bench.py, tests/test_baseline_sizes.py)."""

    def __init__(self, shrink=0, repeats=20, chunks=16, streams=4, graph_replay=False, cfg4_columns=8, sweep_grids=False):
        self.shrink, self.repeats, self.chunks, self.streams = shrink, repeats, chunks, streams
        self.graph_replay, self.cfg4_columns, self.sweep_grids = graph_replay, cfg4_columns, sweep_grids


def measured_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def run_configs(names, options, device, product=None, reference=None):
    """Run the named configurations on the current device (all ranks of a torchrun job must call this together) and
    return their rows."""
    product = product or g.default_library()
    peak, rows = measured_peak(), []
    for name in names:
        t0 = time.time()
        CONFIGS[name](options, device, product, reference, rows, peak)
        torch.cuda.synchronize()
        torch.cuda.empty_cache()
        if RANK == 0:
            log("      [%s took %.1f s]" % (name, time.time() - t0))
    return rows


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("--only", default="")
    parser.add_argument("--shrink", type=int, default=0, help="divide every element count by 2^K")
    parser.add_argument("--repeats", type=int, default=20)
    parser.add_argument("--chunks", type=int, default=16)
    parser.add_argument("--streams", type=int, default=4)
    parser.add_argument("--graph-replay", action="store_true", help="cfg5 host-resident scan: replay the plan as one captured CUDA graph")
    parser.add_argument("--cfg4-columns", type=int, default=8, help="columns of cfg4 spread over the ranks (1 per GPU at 8 GPUs)")
    parser.add_argument("--no-reference", action="store_true")
    parser.add_argument("--sweep-grids", action="store_true", help="also time the fused patched decoders under several grid sizes")
    parser.add_argument("--library", default="", help="path of another build of the product library (make variant ...)")
    parser.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "configs.json"))
    args = parser.parse_args()

    torch.cuda.set_device(LOCAL_RANK)
    device = torch.device("cuda", LOCAL_RANK)
    if WORLD > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.distributed.init_process_group("nccl", device_id=device)
    product = g.Library(args.library) if args.library else g.default_library()
    if WORLD > 1:
        product.host_bind_thread_to_device(LOCAL_RANK)
    reference_path = os.path.join(ROOT, "oracle", "_ref", "libgiddy_ref.so")
    reference = g.Library(reference_path) if os.path.exists(reference_path) and not args.no_reference else None
    only = [s for s in args.only.split(",") if s]
    names = [name for name in CONFIGS if (name in only if only else name != "strong" or WORLD > 1)]
    options = Options(args.shrink, args.repeats, args.chunks, args.streams, args.graph_replay, args.cfg4_columns, args.sweep_grids)
    rows = run_configs(names, options, device, product, reference)
    if RANK == 0:
        os.makedirs(os.path.dirname(args.out), exist_ok=True)
        with open(args.out, "w") as f:
            json.dump({"device": torch.cuda.get_device_name(device), "n_gpus": WORLD, "peak_gbs": measured_peak(), "shrink": args.shrink,
                       "library": product.path, "rows": rows}, f, indent=1)
    if WORLD > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
