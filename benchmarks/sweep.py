#!/usr/bin/env python
"""benchmarks/sweep.py -- per-scheme throughput of the product kernels and, beside them, of the
reference's own kernels rebuilt for sm_100a (oracle/_ref), on synthetic columns of BASELINE.json's
shapes scaled to fit a quick run.  Needs a GPU.

    python benchmarks/sweep.py [--log2-n 27] [--out gpurun_out/sweep.json] [--only rle,delta]

For every workload: decompressed GB/s (output bytes / time), algorithmic GB/s (compressed + side data
read + decompressed written) and its fraction of the measured HBM copy bandwidth.
Timing: CUDA events around `repeats` back-to-back launches after 3 warm-ups; every working set is far
larger than the 126 MB L2.  Outputs are checked bit-for-bit against the oracle before timing.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def input_bytes(case):
    return sum(v.nbytes for v in case.inputs.values())


def workloads(log2_n):
    import cases
    n = 1 << log2_n
    w = []
    w.append(("cfg1 FOR int32<-int16 constant L=4096", lambda: cases.for_case(1, n, 4096, 4, 2, "constant")))
    w.append(("cfg5 FOR int32<-int16 linear L=1024", lambda: cases.for_case(1, n, 1024, 4, 2, "linear")))
    w.append(("Model int32 linear", lambda: cases.model_case(1, n, 4, "linear")))
    w.append(("cfg2 RLE u32 run16 u16 lengths L=1024", lambda: cases.rle_case(1, n, 1024, 4, 2)))
    w.append(("RLE u32 run16 L=4096", lambda: cases.rle_case(1, n, 4096, 4, 2)))
    w.append(("RLE u64 run64 L=1024", lambda: cases.rle_case(1, n // 2, 1024, 8, 2, mean_run=64)))
    w.append(("RLE u8 run16 L=1024", lambda: cases.rle_case(1, n, 1024, 1, 1)))
    w.append(("RLE u16 run16 L=1024", lambda: cases.rle_case(1, n, 1024, 2, 2)))
    w.append(("RPE u16 abs run16 L=1024", lambda: cases.rpe_case(1, n, 1024, 2, 4, False)))
    w.append(("RPE u8 abs run16 L=1024", lambda: cases.rpe_case(1, n, 1024, 1, 4, False)))
    w.append(("cfg5 RPE u32 abs run64 L=1024", lambda: cases.rpe_case(1, n, 1024, 4, 4, False, mean_run=64)))
    w.append(("RPE u32 rel run16 L=1024", lambda: cases.rpe_case(1, n, 1024, 4, 2, True, avoid_boundary_runs=True)))
    w.append(("cfg4 Delta u64<-u32 L=1024", lambda: cases.delta_case(1, n // 2, 1024, 8, 4)))
    w.append(("Delta u32<-u8 L=1024", lambda: cases.delta_case(1, n, 1024, 4, 1)))
    w.append(("Delta u64<-u32 long segments L=2^20 (look-back)", lambda: cases.delta_case(1, n // 2, 1 << 20, 8, 4)))
    w.append(("Delta u64<-u32 one segment (look-back)", lambda: cases.delta_case(1, n // 2, n // 2, 8, 4)))
    w.append(("Delta u32<-u8 one segment (look-back)", lambda: cases.delta_case(1, n, n, 4, 1)))
    w.append(("cfg3 Dictionary u64<-u8 D=256", lambda: cases.dictionary_case(1, n // 2, 8, 1, 256)))
    w.append(("Dictionary u32<-u16 D=60000 (L2 gather)", lambda: cases.dictionary_case(1, n, 4, 2, 60000, cache=False)))
    w.append(("Dictionary u32<-u16 D=60000 (shared-memory head, persistent)", lambda: cases.dictionary_case(1, n, 4, 2, 60000, cache=True)))
    w.append(("Dictionary u32<-u16 D=40000 (shared memory, persistent)", lambda: cases.dictionary_case(1, n, 4, 2, 40000, cache=True)))
    w.append(("Dictionary u32<-u16 D=60000 (default policy)", lambda: cases.dictionary_case(1, n, 4, 2, 60000)))
    w.append(("DZB-fixed u32<-u16", lambda: cases.dzbf_case(1, n, 4, 2)))
    w.append(("DZB-fixed u64<-u8", lambda: cases.dzbf_case(1, n // 2, 8, 1)))
    w.append(("cfg4 DZB-variable u32 2-bit sizes L=1024", lambda: cases.dzbv_case(1, n, 1024, 4, 4, 2, 1)))
    w.append(("cfg4 fused Delta(DZB-variable) u64<-u32 L=1024", lambda: cases.dzbv_delta_case(1, n // 2, 1024, 8, 4, 4, 2, 1)))
    w.append(("cfg5 Incidence bitmaps u32 K=16", lambda: cases.bitmaps_case(1, n, 4, 16)))
    w.append(("cfg3 scatter u64 1% patches", lambda: cases.scatter_case(1, n // 2, n // 200, 8, 4, 4)))
    w.append(("cfg3 fused patched Dictionary u64<-u8 1%", lambda: cases.patched_dictionary_case(1, n // 2, 8, 1, 256, n // 200)))
    w.append(("diagnostic: patched Dictionary u64<-u8, no patches", lambda: cases.patched_dictionary_case(1, n // 2, 8, 1, 256, 0)))
    w.append(("diagnostic: patched Dictionary u64<-u8, 0.01% patches", lambda: cases.patched_dictionary_case(1, n // 2, 8, 1, 256, n // 20000)))
    w.append(("fused patched Dictionary u32<-u8 5%", lambda: cases.patched_dictionary_case(1, n, 4, 1, 256, n // 20)))
    w.append(("cfg5 fused patched FOR i32<-i16 linear ci 1%", lambda: cases.patched_for_case(1, n, 1024, 4, 2, "linear", n // 100, 1024)))
    w.append(("cfg5 ci-scatter u32 1% patches A=1024", lambda: cases.ci_scatter_case(1, n, n // 100, 4, 4, 4, 1024)))
    return w


def time_case(torch, cases, case, library, repeats, single_launch=False):
    device = torch.device("cuda")
    tensors = {k: torch.from_numpy(cases._as_torch_compatible(v)).to(device) for k, v in case.inputs.items()}
    itemsize = case.expected.dtype.itemsize
    if case.initial_output is not None:
        out = torch.from_numpy(cases._as_torch_compatible(case.initial_output)).to(device)
    else:
        out = torch.zeros(case.output_elements + 64, dtype=cases._torch_int(itemsize), device=device)
    case.launch(out, tensors, library, None)
    torch.cuda.synchronize()
    if single_launch:  # profiling mode (under ncu): exactly one launch per workload
        return float("nan"), True
    got = out.cpu().numpy().view(case.expected.dtype)[:case.output_elements]
    correct = bool((got == case.expected).all())
    for _ in range(3):
        case.launch(out, tensors, library, None)
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    start.record()
    for _ in range(repeats):
        case.launch(out, tensors, library, None)
    end.record()
    torch.cuda.synchronize()
    return start.elapsed_time(end) / repeats / 1e3, correct


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--log2-n", type=int, default=27)
    p.add_argument("--repeats", type=int, default=20)
    p.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sweep.json"))
    p.add_argument("--only", default="")
    p.add_argument("--no-reference", action="store_true")
    p.add_argument("--single-launch", action="store_true", help="one untimed launch per workload (for ncu captures)")
    p.add_argument("--library", default="", help="path of another build of the product library (make variant ...)")
    args = p.parse_args()

    import torch
    import cases
    import libgiddy_b200 as g
    product = g.Library(args.library) if args.library else g.default_library()
    reference_path = os.path.join(ROOT, "oracle", "_ref", "libgiddy_ref.so")
    reference = g.Library(reference_path) if os.path.exists(reference_path) and not args.no_reference else None
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        peak = 6650.0
    only = [s for s in args.only.split(",") if s]
    rows = []
    for name, build in workloads(args.log2_n):
        if only and not any(s.lower() in name.lower() for s in only):
            continue
        t0 = time.time()
        case = build()
        build_seconds = time.time() - t0
        is_scatter = case.initial_output is not None
        patch_elements = len(case.inputs["data"]) if "data" in case.inputs else (len(case.inputs["data_to_scatter"]) if is_scatter else 0)
        out_bytes = (patch_elements if is_scatter else case.output_elements) * case.expected.dtype.itemsize
        algorithmic = input_bytes(case) + out_bytes
        row = {"workload": name, "case": case.name, "output_bytes": out_bytes, "algorithmic_bytes": algorithmic}
        seconds, correct = time_case(torch, cases, case, product, args.repeats, args.single_launch)
        if args.single_launch:
            print("launched %s" % name, flush=True)
            continue
        row.update({"ms": seconds * 1e3, "decompressed_gbs": out_bytes / seconds / 1e9, "algorithmic_gbs": algorithmic / seconds / 1e9,
                    "frac_of_measured_peak": algorithmic / seconds / 1e9 / peak, "bit_exact": correct})
        if reference is not None and case.ref_ok:
            try:
                ref_seconds, ref_correct = time_case(torch, cases, case, reference, max(3, args.repeats // 4))
                row.update({"reference_ms": ref_seconds * 1e3, "reference_algorithmic_gbs": algorithmic / ref_seconds / 1e9,
                            "reference_bit_exact": ref_correct, "speedup_vs_reference_kernels": ref_seconds / seconds})
            except Exception as error:
                row["reference_error"] = str(error)[:200]
        rows.append(row)
        print("%-46s %8.3f ms  out %7.1f GB/s  alg %7.1f GB/s (%5.1f%% of %.0f)  exact=%s  ref %s  [built in %.1fs]" % (
            name, row["ms"], row["decompressed_gbs"], row["algorithmic_gbs"], 100 * row["frac_of_measured_peak"], peak, correct,
            ("%.3f ms (x%.1f)" % (row["reference_ms"], row["speedup_vs_reference_kernels"])) if "reference_ms" in row else "-", build_seconds), flush=True)
        del case
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump({"device": torch.cuda.get_device_name(0), "log2_n": args.log2_n, "peak_gbs": peak, "rows": rows}, f, indent=1)


if __name__ == "__main__":
    main()
