#!/usr/bin/env python
"""benchmarks/patched_probe.py -- where the time of the patched decoders goes: BASELINE configs[2]'s column (Dictionary u8 -> i64,
2^29 elements) and configs[4]'s column A (FOR linear i32 <- i16, 2^28) decoded by the plain kernel and by the fused
decode-and-patch kernel under patch densities from none to 5 %, for the default build and every build under
benchmarks/variants/.  Prints a table; --out writes JSON."""
import argparse
import glob
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import libgiddy_b200 as g  # noqa: E402
from libgiddy_b200 import schemes as S  # noqa: E402
from benchmarks import configs as C  # noqa: E402


def timed(run, repeats=20):
    for _ in range(3):
        run()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    start.record()
    for _ in range(repeats):
        run()
    end.record()
    torch.cuda.synchronize()
    return start.elapsed_time(end) / repeats


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "patched_probe.json"))
    parser.add_argument("--shrink", type=int, default=0)
    args = parser.parse_args()
    device = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    libraries = [("default", g.default_library())] + [(os.path.basename(os.path.dirname(p)), g.Library(p))
                                                     for p in sorted(glob.glob(os.path.join(ROOT, "benchmarks", "variants", "*", "libgiddy_b200.so")))]
    peak = C.measured_peak()
    rows = []
    gen = torch.Generator(device=device)
    gen.manual_seed(5)
    # ---- cfg3's column
    n = (1 << 29) >> args.shrink
    dictionary = torch.randint(-(1 << 62), 1 << 62, (256,), generator=gen, device=device, dtype=torch.int64)
    indices = torch.randint(0, 256, (n,), generator=gen, device=device, dtype=torch.int16).to(torch.uint8)
    out = torch.empty(n, dtype=torch.int64, device=device)
    plain_bytes = n + 2048 + n * 8
    for name, library in libraries:
        ms = timed(lambda: S.dictionary(out, indices, dictionary, n, 256, uncompressed_size=8, dictionary_index_size=1, library=library))
        rows.append({"column": "cfg3", "library": name, "kernel": "plain", "density": None, "ms": ms, "frac": plain_bytes / ms / 1e6 / peak})
    for density in (0.0, 0.001, 0.01, 0.05):
        p = int(n * density)
        positions = C.distinct_sorted_positions(device, gen, n, p).to(torch.int32) if p else torch.zeros(1, dtype=torch.int32, device=device)
        values = torch.randint(-(1 << 62), 1 << 62, (max(p, 1),), generator=gen, device=device, dtype=torch.int64)
        for name, library in libraries:
            ms = timed(lambda: S.patched_dictionary(out, indices, dictionary, n, 256, values, positions, p, uncompressed_size=8, dictionary_index_size=1, library=library))
            rows.append({"column": "cfg3", "library": name, "kernel": "fused", "density": density, "ms": ms, "frac": (plain_bytes + p * 12) / ms / 1e6 / peak})
    del indices, out
    torch.cuda.empty_cache()
    # ---- cfg5's column A
    n, L, A = (1 << 28) >> args.shrink, 1024, 1024
    coefficients = torch.randint(-(1 << 30), 1 << 30, (n // L, 2), generator=gen, device=device, dtype=torch.int64).to(torch.int32)
    offsets = torch.randint(-32768, 32768, (n,), generator=gen, device=device, dtype=torch.int32).to(torch.int16)
    out = torch.empty(n, dtype=torch.int32, device=device)
    plain_bytes = n * 2 + coefficients.numel() * 4 + n * 4
    for name, library in libraries:
        ms = timed(lambda: S.frame_of_reference(out, offsets, coefficients, n, L, uncompressed_size=4, compressed_size=2, model="linear", library=library))
        rows.append({"column": "cfg5-A", "library": name, "kernel": "plain", "density": None, "ms": ms, "frac": plain_bytes / ms / 1e6 / peak})
    for density in (0.001, 0.01, 0.05):
        positions, values, ci, num_runs, p = C.synthesize_ci_patches(device, gen, n, density, A, 64)
        patch_bytes = C.nbytes(values, *ci.values()) - 16
        for name, library in libraries:
            ms = timed(lambda: S.patched_frame_of_reference(out, offsets, coefficients, n, L, values, ci["run_lengths"], ci["offset_sizes"], ci["baselines"],
                                                            ci["offsets_start"], ci["offset_bytes"], ci["anchors"], A, num_runs, p,
                                                            uncompressed_size=4, compressed_size=2, model="linear", library=library))
            rows.append({"column": "cfg5-A", "library": name, "kernel": "fused", "density": density, "ms": ms, "frac": (plain_bytes + patch_bytes) / ms / 1e6 / peak})
    for r in rows:
        print("%-7s %-10s %-6s density %-6s %8.3f ms  %5.1f%% of HBM" % (r["column"], r["library"], r["kernel"], r["density"], r["ms"], 100 * r["frac"]), flush=True)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump({"peak_gbs": peak, "rows": rows}, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
