#!/usr/bin/env python
"""bench.py -- the headline benchmark of the decompression path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--configs cfg1,cfg3,...|none]

Workload (BASELINE.json configs[1]): Run-Length Encoding, uint32 values + uint16 run lengths, 2^30
decompressed elements per GPU, run lengths uniform in {1..31} (mean 16), anchored segments of 1024
elements, synthetic data generated on the device from a fixed seed.  A "step" is one decode of the
whole column through the library's registry C ABI (resolve_launch_configuration + enqueue_launch of
"decompression::run_length_encoding::kernel_t<4u, 4u, 2u, 4u>"): one kernel launch.

Prints ONE JSON line (rank 0):
  value     decompressed GB/s (10^9 output bytes per second), summed over the N GPUs, inputs resident in HBM
  roofline  algorithmic HBM bytes (compressed + side data read, decompressed written) / kernel time,
            against the measured copy bandwidth in MEASURED_PEAKS.json
  e2e       the same metric from pinned HOST buffers: the compressed column is copied H2D in segment-range
            chunks on side streams, each chunk decoded as it lands, and a sample of every chunk's output
            is read back D2H (e2e.value: the decompressed column stays in HBM for a GPU consumer);
            e2e.result_returning is the same with every chunk's WHOLE output copied back to pinned host
            memory, overlapped with the next chunks (for a host consumer); both with their measured
            raw-copy ceilings; all inside the timed region
  configs   every other BASELINE.json configuration at its named shape (benchmarks/configs.py), timed in
            this process after the headline: ms, algorithmic bytes, fraction of the HBM roofline,
            bit-exactness, the reference's own kernels beside it; under torchrun also ONE column split
            into segment-range shards over the GPUs (strong scaling)
  cpu_baseline  the oracle's scalar host decoder (OpenMP, all host threads) on a bounded sample
With --impl reference the CPU decoder alone is timed (the reference ships no CPU path; its kernels
are GPU-only, so the "reference arm" of this tier is the host decoder written from its format).

Multi-GPU (torchrun, one rank per GPU): every rank decodes its own column (weak scaling); decode needs
no collective; timing is the max over ranks of CUDA-event time between two barriers.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LOG2_N = 30
SEGMENT_LENGTH = 1024
MAX_RUN = 31
KEY = "decompression::run_length_encoding::kernel_t<4u, 4u, 2u, 4u>"
ORACLE_MARCH = None


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=200)
    p.add_argument("--warmup", type=int, default=10)
    p.add_argument("--impl", default="giddy_b200", choices=["giddy_b200", "reference"])
    p.add_argument("--log2-n", type=int, default=LOG2_N, help="log2 of the number of decompressed elements per GPU")
    p.add_argument("--segment-length", type=int, default=SEGMENT_LENGTH)
    p.add_argument("--cpu-sample-log2", type=int, default=30, help="log2 of the elements the CPU baseline decodes per pass")
    p.add_argument("--cpu-seconds", type=float, default=12.0, help="host time the cpu_baseline leg spends decoding (bounded sample)")
    p.add_argument("--skip-reference-kernels", action="store_true")
    p.add_argument("--skip-e2e", action="store_true")
    p.add_argument("--skip-cpu", action="store_true")
    p.add_argument("--skip-verify", action="store_true")
    p.add_argument("--configs", default="cfg1,cfg3,cfg4,cfg5", help="BASELINE configurations timed after the headline (comma-separated, or 'none')")
    p.add_argument("--configs-repeats", type=int, default=10)
    return p.parse_args()


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def known_traffic():
    """DRAM bytes per launch of the headline kernel from the committed ncu capture, if any."""
    try:
        with open(os.path.join(ROOT, "profiles", "headline_traffic.json")) as f:
            return json.load(f).get("dram_bytes_per_launch")
    except Exception:
        return None


# ---------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------

class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.samples, self.process, self.thread = [], None, None
        self.device_index = device_index

    def start(self):
        try:
            self.process = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.device_index), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.process = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.process.stdout:
            self.samples.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self):
        if self.process is not None:
            self.process.terminate()
            try:
                self.process.wait(timeout=2)
            except Exception:
                self.process.kill()

    def summary(self, t0, t1):
        rows = [r for (t, r) in self.samples if t0 <= t <= t1 and len(r) >= 7]
        if not rows:
            rows = [r for (_, r) in self.samples if len(r) >= 7]
        if not rows:
            return None
        def number(x):
            try:
                return float(x)
            except ValueError:
                return None
        sm = sorted(v for v in (number(r[0]) for r in rows) if v is not None)
        reasons = []
        for index, name in ((3, "hw_slowdown"), (4, "hw_thermal_slowdown"), (5, "sw_thermal_slowdown"), (6, "sw_power_cap")):
            if any(r[index].lower().startswith("active") for r in rows):
                reasons.append(name)
        power = [v for v in (number(r[2]) for r in rows) if v is not None]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": number(rows[0][1]), "reasons": reasons,
                "power_w_max": max(power) if power else None, "samples": len(rows)}


# ---------------------------------------------------------------------------------------------------
# synthetic column (device side)
# ---------------------------------------------------------------------------------------------------

def synthesize_rle_column(torch, n, segment_length, seed, device):
    """Run lengths i.i.d. uniform {1..31}, last run truncated to end exactly at n; values uniform uint32."""
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    estimate = int(n / ((1 + MAX_RUN) / 2) * 1.02) + 4096
    lengths = torch.randint(1, MAX_RUN + 1, (estimate,), generator=gen, device=device, dtype=torch.int64)
    ends = torch.cumsum(lengths, 0)
    num_runs = int(torch.searchsorted(ends, torch.tensor([n], device=device, dtype=torch.int64)).item()) + 1
    assert num_runs <= estimate
    lengths, ends = lengths[:num_runs].clone(), ends[:num_runs].clone()
    lengths[-1] -= ends[-1] - n
    ends[-1] = n
    starts = ends - lengths
    values = torch.randint(-(1 << 31), 1 << 31, (num_runs,), generator=gen, device=device, dtype=torch.int64).to(torch.int32)
    num_segments = (n + segment_length - 1) // segment_length
    segment_starts = torch.arange(num_segments, device=device, dtype=torch.int64) * segment_length
    anchors = torch.searchsorted(starts, segment_starts, right=True) - 1
    offsets = segment_starts - starts[anchors]
    column = {
        "run_data": values, "run_lengths": lengths.to(torch.int16), "position_anchors": anchors.to(torch.int32),
        "intra_run_anchor_offsets": offsets.to(torch.int16),
        "n": n, "segment_length": segment_length, "num_segments": num_segments, "num_runs": num_runs,
    }
    del ends, starts, segment_starts, anchors, offsets
    return column, lengths


def algorithmic_bytes(column):
    # decompressed write + run values and lengths read + anchors and intra-run offsets read
    return column["n"] * 4 + column["num_runs"] * (4 + 2) + column["num_segments"] * (4 + 2)


def rle_arguments(g, column, out_address, anchors_address, offsets_address, num_segments, length, run_data=None, run_lengths=None):
    size = g.size_type_by_index_size(4)
    return {
        "decompressed": g.ptr(out_address, "u32"),
        "run_data": g.cptr(column["run_data"] if run_data is None else run_data, "u32"),
        "run_lengths": g.cptr(column["run_lengths"] if run_lengths is None else run_lengths, "u16"),
        "position_anchors": g.cptr(anchors_address, "u32"),
        "intra_run_anchor_offsets": g.cptr(offsets_address, "u16"),
        "segment_length": size(column["segment_length"]), "num_segments": size(num_segments),
        "num_element_runs": size(column["num_runs"]), "length": size(length),
    }


# ---------------------------------------------------------------------------------------------------
# the reference arm: the scalar host decoder on the box's host cores
# ---------------------------------------------------------------------------------------------------

def host_sample(np, log2_elements, segment_length, seed):
    """The same distribution as the device-side generator, built with numpy (no GPU needed)."""
    from libgiddy_b200 import encode as E
    n = 1 << log2_elements
    r = np.random.default_rng(seed)
    estimate = int(n / 16 * 1.02) + 4096
    lengths = r.integers(1, MAX_RUN + 1, estimate, dtype=np.int64)
    ends = np.cumsum(lengths)
    num_runs = int(np.searchsorted(ends, n)) + 1
    lengths, ends = lengths[:num_runs], ends[:num_runs]
    lengths[-1] -= ends[-1] - n
    starts = np.concatenate(([0], ends[:-1]))
    starts[-1] = n - lengths[-1]
    values = r.integers(0, 1 << 32, num_runs, dtype=np.uint64).astype(np.uint32)
    segment_starts = np.arange(E.num_segments(n, segment_length), dtype=np.int64) * segment_length
    anchors = np.searchsorted(starts, segment_starts, side="right") - 1
    return {"run_data": values, "run_lengths": lengths.astype(np.uint16), "position_anchors": anchors.astype(np.uint32),
            "intra_run_anchor_offsets": (segment_starts - starts[anchors]).astype(np.uint16), "n": n}


def time_host_decoder(np, sample, segment_length, passes, seconds=None):
    """Decode the sample `passes` times -- or, with `seconds`, as many times as fit (at least 3, at most 400)."""
    from oracle import giddy_oracle as O  # the checker / CPU baseline: the only use of oracle/ in this file
    global ORACLE_MARCH
    if ORACLE_MARCH is None:
        ORACLE_MARCH = O.use_native_build()  # -O3 -march=native on the box it is timed on
    # all the host threads this process may use -- torchrun exports OMP_NUM_THREADS=1, which would leave the decoder on one core
    O.set_num_threads(len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    out = np.empty(sample["n"], dtype=np.uint32)
    out[:] = 0  # first touch
    times = []
    while True:
        t0 = time.perf_counter()
        O.run_length_encoding(sample["run_data"], sample["run_lengths"], sample["position_anchors"],
                              sample["intra_run_anchor_offsets"], segment_length, sample["n"], out=out)
        times.append(time.perf_counter() - t0)
        if seconds is None:
            if len(times) >= passes:
                break
        elif (len(times) >= 3 and sum(times) >= seconds) or len(times) >= 400:
            break
    return times, O.num_threads()


def run_reference_arm(args):
    import numpy as np
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = host_sample(np, min(args.cpu_sample_log2, args.log2_n), args.segment_length, 3)
    times, threads = time_host_decoder(np, sample, args.segment_length, args.warmup + args.steps)
    timed = times[args.warmup:]
    total = sum(timed)
    value = sample["n"] * 4 * len(timed) / total / 1e9
    description = "2^%d elements of the column's distribution (run lengths U{1..31}, L=%d) per step, %d timed steps, %.1f s of host time" % (
        min(args.cpu_sample_log2, args.log2_n), args.segment_length, len(timed), sum(times))
    print(json.dumps({
        "impl": "reference", "metric": "decompressed GB/s", "value": value, "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(timed), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": "GB/s", "cores": threads, "kind": "port", "sample": "each step decodes a bounded sample: " + description,
                         "march": ORACLE_MARCH},
        "e2e": {"value": value, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "the reference (eyalroz/libgiddy) has no CPU implementation; this is the scalar host decoder written from its format "
                "definitions (oracle/giddy_oracle.cpp), OpenMP over segments, on the box's host threads",
    }))


def workload_config(args):
    # (key for key the same in both arms: the driver compares them)
    return {
        "workload": "RLE decode: uint32 values + uint16 run lengths, 2^%d decompressed elements per GPU, run lengths uniform {1..31} (mean 16), "
                    "anchored segments of %d elements (BASELINE.json configs[1])" % (args.log2_n, args.segment_length),
        "factory_key": KEY,
        "elements_per_gpu": 1 << args.log2_n, "segment_length": args.segment_length,
        "cache": "working set per step (4.7 GB) is far larger than the 126 MB L2; no flush needed",
        "parallelism": "segment ranges / columns sharded over GPUs, no data-path collective",
    }


# ---------------------------------------------------------------------------------------------------
# the product arm
# ---------------------------------------------------------------------------------------------------

def run_product_arm(args):
    import numpy as np
    import torch
    import libgiddy_b200 as g
    from libgiddy_b200 import columns

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    distributed = world > 1
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libgiddy_b200 has no CPU path (use --impl reference for the host decoder)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if distributed:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # stdout carries the one JSON line and nothing else
        dist.init_process_group("nccl", device_id=device)

    library = g.default_library()
    # pinned buffers of the host-resident leg live on the NUMA node this rank's GPU hangs off (giddy_host_alloc), and so does
    # the thread which submits the copies; None where the platform reports no node
    numa_node = library.host_bind_thread_to_device(local_rank) if distributed else library.host_numa_node_of_device(local_rank)
    kernel = library.produce_subclass(KEY)
    n, L = 1 << args.log2_n, args.segment_length
    column, lengths64 = synthesize_rle_column(torch, n, L, 1234 + rank, device)
    out = torch.empty(n + 64, dtype=torch.int32, device=device)
    resolve_arguments = {"length": g.size_t(n), "segment_length": g.size_t(L)}
    launch_arguments = rle_arguments(g, column, out.data_ptr(), column["position_anchors"], column["intra_run_anchor_offsets"],
                                     column["num_segments"], n)
    launches = [0]

    def step():
        config = kernel.resolve_launch_configuration(resolve_arguments)
        kernel.enqueue_launch(None, config, launch_arguments)
        launches[0] += 1
        return config

    # correctness at full size: the decoded column must equal the run values repeated by the run lengths
    config = step()
    torch.cuda.synchronize()
    verified = None
    if not args.skip_verify:
        expected = torch.repeat_interleave(column["run_data"], lengths64)
        verified = bool(torch.equal(out[:n], expected))
        del expected
        if not verified:
            raise SystemExit("bench.py: decoded column differs from the expected expansion -- refusing to time a wrong kernel")
    del lengths64
    torch.cuda.empty_cache()

    sampler = ClockSampler(torch.cuda.current_device() if "CUDA_VISIBLE_DEVICES" not in os.environ else local_rank)
    if rank == 0:
        sampler.start()

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    launches[0] = 0
    t_begin = time.time()
    start.record()
    for _ in range(args.steps):
        step()
    end.record()
    barrier()
    t_end = time.time()
    elapsed_ms = start.elapsed_time(end)
    gpu_launches = launches[0]
    if distributed:
        t = torch.tensor([elapsed_ms], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
        totals = torch.tensor([float(algorithmic_bytes(column))], device=device, dtype=torch.float64)
        dist.all_reduce(totals, op=dist.ReduceOp.SUM)
        algorithmic_total = float(totals.item())
    else:
        algorithmic_total = float(algorithmic_bytes(column))
    seconds = elapsed_ms / 1e3
    value = world * n * 4 * args.steps / seconds / 1e9

    # keep the GPU under the same load a little longer if the timed region was too short to sample clocks
    if rank == 0 and t_end - t_begin < 1.0:
        probe_until = time.time() + 1.2
        while time.time() < probe_until:
            for _ in range(50):
                step()
            torch.cuda.synchronize()
        t_end = time.time()

    peak, peak_source = measured_peak()
    # one launch per step: the kernel's average duration is the step time on this rank's stream
    achieved = algorithmic_bytes(column) * args.steps / (start.elapsed_time(end) / 1e3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": known_traffic(),
                "peak_source": peak_source, "frac_of_nominal_8000": achieved / 8000.0,
                "algorithmic_bytes_per_launch": algorithmic_bytes(column), "kernel": "run_length_encoding_decompress<4,4,2,4>",
                "launch": {"grid": config.grid, "block": config.block, "dynamic_shared_memory": config.dynamic_shared_memory_size}}

    e2e = None
    if not args.skip_e2e:
        e2e = run_e2e(args, torch, np, g, columns, library, kernel, column, out, device, distributed, world)
    clocks = sampler.summary(t_begin, t_end) if rank == 0 else None
    sampler.stop()

    reference_kernels = None
    if rank == 0 and not args.skip_reference_kernels:
        reference_kernels = time_reference_kernels(args, torch, g, column, out, resolve_arguments, launch_arguments)

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.skip_cpu:
        sample = host_sample(np, args.cpu_sample_log2, L, 3)
        times, threads = time_host_decoder(np, sample, L, 0, seconds=args.cpu_seconds)
        timed = times[1:]
        cpu_baseline = {"value": sample["n"] * 4 * len(timed) / sum(timed) / 1e9, "unit": "GB/s", "cores": threads, "kind": "port", "march": ORACLE_MARCH,
                        "best_pass_value": sample["n"] * 4 / min(timed) / 1e9,
                        "sample": "2^%d elements of the same distribution (run lengths U{1..31}, L=%d) decoded %d times after one warm-up pass: "
                                  "%.1f s of host time; scalar C++ host decoder (oracle/giddy_oracle.cpp), OpenMP over segments" % (
                                      args.cpu_sample_log2, L, len(timed), sum(times))}

    # every other BASELINE configuration, in this process, after the headline (all ranks together under torchrun)
    configs = None
    if args.configs and args.configs != "none":
        del column, out, launch_arguments
        torch.cuda.empty_cache()
        configs = run_configs_block(args, torch, g, library, device, world, rank)

    if rank == 0:
        print(json.dumps({
            "metric": "decompressed GB/s", "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32", "data": "synthetic", "config": workload_config(args), "roofline": roofline, "cpu_baseline": cpu_baseline,
            "e2e": e2e, "gpu_launches": gpu_launches, "clocks": clocks, "verified_bit_exact": verified,
            "reference_kernels": reference_kernels, "numa_node_rank0": numa_node, "configs": configs,
            "algorithmic_gbs_all_gpus": algorithmic_total * args.steps / seconds / 1e9,
            "library": library.version(),
        }))
    if distributed:
        dist.destroy_process_group()


def time_reference_kernels(args, torch, g, column, out, resolve_arguments, launch_arguments):
    """A reported baseline, outside the timed region: the reference's OWN kernel and launch-configuration resolver for the
    same factory key (eyalroz/libgiddy rebuilt for sm_100a by oracle/build_ref.sh), on the same resident column, rank 0."""
    path = os.path.join(ROOT, "oracle", "_ref", "libgiddy_ref.so")
    if not os.path.exists(path):
        return {"unavailable": "oracle/_ref/libgiddy_ref.so has not been built"}
    try:
        kernel = g.Library(path).produce_subclass(KEY)
        steps = max(3, min(args.steps, 20))
        config = None
        for _ in range(3):
            config = kernel.resolve_launch_configuration(resolve_arguments)
            kernel.enqueue_launch(None, config, launch_arguments)
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        start.record()
        for _ in range(steps):
            config = kernel.resolve_launch_configuration(resolve_arguments)
            kernel.enqueue_launch(None, config, launch_arguments)
        end.record()
        torch.cuda.synchronize()
        seconds = start.elapsed_time(end) / 1e3 / steps
        return {"value": column["n"] * 4 / seconds / 1e9, "unit": "GB/s", "ms_per_step": seconds * 1e3, "steps": steps,
                "algorithmic_gbs": algorithmic_bytes(column) / seconds / 1e9,
                "launch": {"grid": config.grid, "block": config.block, "dynamic_shared_memory": config.dynamic_shared_memory_size},
                "what": "eyalroz/libgiddy's own run_length_encoding::decompress kernel + resolver, rebuilt for sm_100a, same column"}
    except Exception as error:
        return {"unavailable": str(error)[:200]}


def run_configs_block(args, torch, g, library, device, world, rank):
    """BASELINE configs[0], [2], [3], [4] (and, on several GPUs, one column split by segment range) through
    benchmarks/configs.py: same process, same library, after the headline."""
    from benchmarks import configs as C
    C.LOG_FILE = sys.stderr
    names = [name for name in args.configs.split(",") if name in C.CONFIGS]
    if world > 1 and "strong" not in names:
        names.append("strong")
    reference = None
    path = os.path.join(ROOT, "oracle", "_ref", "libgiddy_ref.so")
    if os.path.exists(path) and not args.skip_reference_kernels:
        reference = g.Library(path)
    options = C.Options(shrink=LOG2_N - args.log2_n if args.log2_n < LOG2_N else 0, repeats=args.configs_repeats)
    try:
        rows = C.run_configs(names, options, device, library, reference)
    except Exception as error:  # a failed side measurement must not cost the headline
        if world > 1:
            raise
        return {"failed": str(error)[:300]}
    keep = ("config", "chain", "ms", "decompressed_gbs", "algorithmic_gbs", "algorithmic_bytes_per_gpu", "frac_of_measured_peak_per_gpu", "bit_exact",
            "reference_kernels_ms", "speedup_vs_reference_kernels", "extension", "scaling", "whole_column_decompressed_gbs", "h2d_gbs_per_gpu",
            "gain_over_uncompressed_copy", "compression_ratio", "uncompressed_copy_gbs")
    return {"repeats": args.configs_repeats, "timing": "CUDA events around `repeats` back-to-back passes after 3 warm-ups, max over ranks; working sets >> L2",
            "frac": "algorithmic bytes / time / MEASURED_PEAKS.json hbm_gbs, per GPU", "all_bit_exact": all(bool(r["bit_exact"]) for r in rows),
            "rows": [{k: r[k] for k in keep if k in r} for r in rows]}


def run_e2e(args, torch, np, g, columns, library, kernel, column, out, device, distributed, world):
    """Host-resident mode.  The compressed column lives in pinned host memory in CHUNK-MAJOR form: for each of 16 segment
    ranges, the slices that range reads -- run values, run lengths, anchors rebased to its first run, intra-run offsets
    -- back to back in one giddy_host_alloc buffer (written once, untimed, by the library's planner: giddy_shard_plan +
    giddy_slice_copy).  A pass = per chunk ONE H2D copy, one decode launch, and the copy back of (a) a sample of the
    chunk's output or (b) all of it; chunks go round-robin over 4 side streams; planned once and replayed by the
    library's scan executor (giddy_scan_run)."""
    from libgiddy_b200 import _native as N
    n, L, S, R = column["n"], column["segment_length"], column["num_segments"], column["num_runs"]
    names = ("run_data", "run_lengths", "position_anchors", "intra_run_anchor_offsets")
    whole = {k: column[k].cpu().numpy() for k in names}
    whole = {k: (v.view("uint%d" % (8 * v.dtype.itemsize))) for k, v in whole.items()}
    num_chunks, num_streams, sample_elements = 16, 4, 16384
    layout = columns._layout(N.SCHEME_RUN_LENGTH_ENCODING, n, L, index_size=4, uncompressed_size=4, compressed_size=2, num_runs=R,
                             anchors=whole["position_anchors"])
    shards = [library.shard_plan(layout, *library.split_segments(S, num_chunks, c)) for c in range(num_chunks)]

    def aligned(x):
        return (x + 255) // 256 * 256
    placements, total = [], 0  # per chunk: {array name: (byte offset in the packed buffer, bytes)}
    for shard in shards:
        placement = {}
        for k in range(shard.num_slices):
            piece = shard.slices[k]
            placement[piece.name] = (total, library.slice_bytes(piece))
            total = aligned(total + library.slice_bytes(piece))
        placements.append(placement)
    host_packed = library.host_alloc(total, device.index)  # on the GPU's NUMA node where there is one
    for shard, placement in zip(shards, placements):
        for k in range(shard.num_slices):
            piece = shard.slices[k]
            library.slice_copy(piece, host_packed.address + placement[piece.name][0], whole[piece.name])
    del whole
    dev_packed = torch.empty(total, dtype=torch.uint8, device=device)
    samples_buffer = library.host_alloc(num_chunks * sample_elements * 4, device.index)
    samples = samples_buffer.tensor("int32").view(num_chunks, sample_elements)
    h2d_bytes = sum(nbytes for placement in placements for (_, nbytes) in placement.values())
    d2h_bytes = num_chunks * sample_elements * 4
    size = g.size_type_by_index_size(4)

    def chunk_span(placement):
        first = min(offset for (offset, _) in placement.values())
        last = max(offset + nbytes for (offset, nbytes) in placement.values())
        return first, last - first

    def build(copy_back):
        """copy_back: 'sample', 'all' or None (decode only); returns (scan, pinned host output or None)"""
        scan = library.create_scan(num_streams)
        host_out = library.host_alloc(n * 4, device.index) if copy_back == "all" else None
        for index, (shard, placement) in enumerate(zip(shards, placements)):
            first, span = chunk_span(placement)
            scan.add_copy(index, dev_packed.data_ptr() + first, host_packed.address + first, span)  # ONE copy per chunk
            base = dev_packed.data_ptr()
            first_element, count = int(shard.first_element), int(shard.num_elements)
            arguments = {
                "decompressed": g.ptr(out.data_ptr() + 4 * first_element, "u32"),
                "run_data": g.cptr(base + placement["run_data"][0], "u32"), "run_lengths": g.cptr(base + placement["run_lengths"][0], "u16"),
                "position_anchors": g.cptr(base + placement["position_anchors"][0], "u32"),
                "intra_run_anchor_offsets": g.cptr(base + placement["intra_run_anchor_offsets"][0], "u16"),
                "segment_length": size(L), "num_segments": size(int(shard.num_segments)), "num_element_runs": size(int(shard.num_runs)),
                "length": size(count)}
            config = kernel.resolve_launch_configuration({"length": g.size_t(count), "segment_length": g.size_t(L)})
            scan.add_launch(index, kernel, config, arguments)
            if copy_back == "sample":
                tail = first_element + count - sample_elements
                scan.add_copy_back(index, samples[index], out[tail: tail + sample_elements])
            elif copy_back == "all":
                scan.add_copy_back(index, host_out.address + 4 * first_element, out.data_ptr() + 4 * first_element, 4 * count)
        return scan, host_out

    def copies_only(direction, host_out=None):
        """The same transfers with nothing else: the ceiling the executor's overlap is measured against."""
        scan = library.create_scan(num_streams)
        for index, (shard, placement) in enumerate(zip(shards, placements)):
            if direction == "h2d":
                first, span = chunk_span(placement)
                scan.add_copy(index, dev_packed.data_ptr() + first, host_packed.address + first, span)
            else:
                first_element, count = int(shard.first_element), int(shard.num_elements)
                scan.add_copy_back(index, host_out.address + 4 * first_element, out.data_ptr() + 4 * first_element, 4 * count)
        return scan

    def barrier():
        if distributed:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def timed(scan, steps, warmup=3):
        """ms per pass, max over ranks: CUDA events around `steps` replays between two barriers"""
        for _ in range(warmup):
            scan.run()
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        start.record()
        for _ in range(steps):
            scan.run()
        end.record()
        barrier()
        elapsed_ms = start.elapsed_time(end)
        if distributed:
            import torch.distributed as dist
            t = torch.tensor([elapsed_ms], device=device, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            elapsed_ms = float(t.item())
        return elapsed_ms / steps

    def gbs(nbytes, ms):
        return nbytes / (ms / 1e3) / 1e9

    steps = max(3, min(args.steps, 20))
    resident = out[:n].clone()  # what the in-HBM decode of the timed region produced
    out.zero_()
    dev_packed.zero_()
    scan, _ = build("sample")
    ms = timed(scan, steps)
    matches = bool(torch.equal(out[:n], resident))
    last = shards[-1]
    tail = int(last.first_element + last.num_elements) - sample_elements
    sample_ok = bool(torch.equal(samples[-1], resident[tail: tail + sample_elements].cpu()))
    ceiling_ms = timed(copies_only("h2d"), steps)
    result = {"value": world * gbs(n * 4, ms), "unit": "GB/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes, "steps": steps,
              "ms_per_step": ms, "gpu_launches": num_chunks * steps, "chunks": num_chunks, "streams": num_streams, "copies_per_step": 2 * num_chunks,
              "sample_matches": sample_ok, "matches_resident_decode": matches,
              "what": "e2e.value: pinned host compressed column (chunk-major, giddy_host_alloc) -> ONE H2D copy per segment-range chunk (%d chunks on %d "
                      "streams) -> decode per chunk -> D2H of each chunk's last %d elements; the decompressed column stays in HBM for its GPU "
                      "consumer.  e2e.result_returning copies every chunk's whole output back instead.  Both replayed by the library's scan "
                      "executor (giddy_scan_run)" % (num_chunks, num_streams, sample_elements),
              "h2d_gbs_per_gpu": gbs(h2d_bytes, ms), "h2d_ceiling_gbs_per_gpu": gbs(h2d_bytes, ceiling_ms),
              "h2d_frac_of_ceiling": ceiling_ms / ms,
              "ceiling": "the same %d copies on the same streams with no kernels, all ranks at once" % num_chunks,
              "host_memory": {"numa_node": host_packed.numa_node, "allocator": "giddy_host_alloc (mmap + mbind + cudaHostRegister)"}}
    # the same plan replayed as one captured CUDA graph
    try:
        scan.set_graph_replay(True)
        graph_ms = timed(scan, steps, warmup=2)
        result["with_graph_replay"] = {"value": world * gbs(n * 4, graph_ms), "unit": "GB/s", "steps": steps}
    except Exception as error:
        result["with_graph_replay"] = {"skipped": str(error)[:100]}
    del scan
    # result-returning: every chunk's whole output goes back to pinned host memory, overlapped with the following chunks
    try:
        full_steps = max(5, min(args.steps, 10))
        out.zero_()
        scan, host_out = build("all")
        full_ms = timed(scan, full_steps, warmup=2)
        returned = host_out.tensor("int32")
        ok = True
        for shard in (shards[0], shards[num_chunks // 2], shards[-1]):
            a, b = int(shard.first_element), int(shard.first_element + shard.num_elements)
            ok = ok and bool(torch.equal(returned[a:b], resident[a:b].cpu()))
        d2h_ceiling_ms = timed(copies_only("d2h", host_out), full_steps, warmup=1)
        result["result_returning"] = {
            "value": world * gbs(n * 4, full_ms), "unit": "GB/s", "steps": full_steps, "ms_per_step": full_ms, "h2d_bytes_per_step": h2d_bytes,
            "d2h_bytes_per_step": n * 4, "returned_column_matches": ok, "d2h_gbs_per_gpu": gbs(n * 4, full_ms),
            "d2h_ceiling_gbs_per_gpu": gbs(n * 4, d2h_ceiling_ms), "d2h_frac_of_ceiling": d2h_ceiling_ms / full_ms,
            "what": "as e2e.value, but each chunk's whole decompressed output (256 MB) is copied back to pinned host memory on the chunk's "
                    "stream while later chunks are copied in and decoded; bound by the D2H link: 10x more bytes go back than come in"}
        del scan, host_out
    except Exception as error:  # e.g. 4 GiB of pinned memory refused on a small host
        result["result_returning"] = {"skipped": str(error)[:200]}
    del resident
    return result


class _quiet_stdout:
    """stdout carries the one JSON line and nothing else: while the benchmark runs, file descriptor 1 points at stderr (NCCL
    prints its version banner to fd 1 from C when the box sets NCCL_DEBUG), and Python-level prints are collected and
    written to the real stdout at the end."""

    def __enter__(self):
        import io
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        self.python_stdout, sys.stdout = sys.stdout, io.StringIO()
        return self

    def __exit__(self, *exc):
        collected, sys.stdout = sys.stdout.getvalue(), self.python_stdout
        os.dup2(self.saved, 1)
        os.close(self.saved)
        sys.stdout.write(collected)
        sys.stdout.flush()
        return False


def main():
    args = parse_args()
    with _quiet_stdout():
        if args.impl == "reference":
            run_reference_arm(args)
        else:
            run_product_arm(args)


if __name__ == "__main__":
    main()
