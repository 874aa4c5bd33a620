#!/usr/bin/env python
"""benchmarks/compress_sweep.py -- throughput of the GPU compressors on columns of BASELINE.json's shapes (generated on the device),
each checked by a round trip through the library's decompressor.  Roofline: HBM; algorithmic bytes = uncompressed column read ONCE +
compressed arrays written (the two-pass encoders -- RLE, RPE, DZB-variable -- read the column twice: their fraction shows it).

    python benchmarks/compress_sweep.py [--shrink K] [--out gpurun_out/compress_sweep.json]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import libgiddy_b200 as g  # noqa: E402
from libgiddy_b200 import compress as K  # noqa: E402
from libgiddy_b200 import schemes as S  # noqa: E402
from benchmarks import configs as C  # noqa: E402


def timed(run, repeats):
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
    parser.add_argument("--shrink", type=int, default=0)
    parser.add_argument("--repeats", type=int, default=10)
    parser.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "compress_sweep.json"))
    args = parser.parse_args()
    device = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    library = g.default_library()
    peak = C.measured_peak()
    rows = []
    status = torch.zeros(1, dtype=torch.int32, device=device)

    def report(name, column_bytes, compressed_bytes, ms, ok):
        row = {"workload": name, "ms": ms, "uncompressed_gbs": column_bytes / ms / 1e6, "algorithmic_gbs": (column_bytes + compressed_bytes) / ms / 1e6,
               "frac_of_measured_peak": (column_bytes + compressed_bytes) / ms / 1e6 / peak, "round_trip_exact": ok, "status_flags": int(status.item())}
        rows.append(row)
        print("%-58s %8.3f ms  in %7.1f GB/s  alg %7.1f GB/s (%5.1f%% of %.0f)  round trip=%s" % (
            name, ms, row["uncompressed_gbs"], row["algorithmic_gbs"], 100 * row["frac_of_measured_peak"], peak, ok), flush=True)

    gen = torch.Generator(device=device)
    gen.manual_seed(11)
    # cfg1: Frame of Reference int32 -> int16, constant model, L = 4096
    n, L = (1 << 26) >> args.shrink, 4096
    column = (torch.randint(-(1 << 29), 1 << 29, (n // L, 1), generator=gen, device=device, dtype=torch.int64) +
              torch.randint(-30000, 30000, (n // L, L), generator=gen, device=device, dtype=torch.int64)).reshape(-1).to(torch.int32)
    out, coefficients, back = torch.empty(n, dtype=torch.int16, device=device), torch.empty(n // L, dtype=torch.int32, device=device), torch.empty(n, dtype=torch.int32, device=device)
    run = lambda: K.frame_of_reference(out, coefficients, column, n, L, uncompressed_size=4, compressed_size=2, model="constant", status=status, library=library)
    run()
    S.frame_of_reference(back, out, coefficients, n, L, uncompressed_size=4, compressed_size=2, model="constant", library=library)
    report("cfg1 FOR int32->int16 constant L=4096 (n=2^%d)" % (26 - args.shrink), n * 4, n * 2 + n // L * 4, timed(run, args.repeats), bool(torch.equal(back, column)))
    del column, out, back
    # cfg2: RLE uint32, mean run 16, uint16 lengths, L = 1024
    n, L = (1 << 30) >> args.shrink, 1024
    values, lengths64, _, _, _, R = C.synthesize_runs(device, gen, n, 31, L, torch.int16)
    same = values[1:] == values[:-1]
    values[1:][same] ^= 1  # adjacent runs must differ for the encoder to find them again
    column = torch.repeat_interleave(values, lengths64)
    del lengths64
    S_ = n // L
    capacity = R + 1024
    run_data, run_lengths = torch.empty(capacity, dtype=torch.int32, device=device), torch.empty(capacity, dtype=torch.int16, device=device)
    anchors, offsets, count = torch.empty(S_, dtype=torch.int32, device=device), torch.empty(S_, dtype=torch.int16, device=device), torch.zeros(1, dtype=torch.int32, device=device)
    run = lambda: K.run_length_encoding(run_data, run_lengths, anchors, offsets, count, column, n, L, capacity, uncompressed_size=4, run_length_size=2, status=status, library=library)
    run()
    found = int(count.item())
    back = torch.empty(n + 64, dtype=torch.int32, device=device)
    S.run_length_encoding(back, run_data, run_lengths, anchors, offsets, L, S_, found, n, uncompressed_size=4, run_length_size=2, library=library)
    report("cfg2 RLE uint32, run 16, uint16 lengths, L=1024 (n=2^%d, %d runs)" % (30 - args.shrink, found), n * 4, found * 6 + S_ * 6, timed(run, args.repeats),
           bool(torch.equal(back[:n], column)))
    count.zero_()
    starts = torch.empty(capacity, dtype=torch.int32, device=device)
    run = lambda: K.run_position_encoding(run_data, starts, anchors, count, column, n, L, capacity, uncompressed_size=4, position_offset_size=4, relative=False, status=status, library=library)
    run()
    S.run_position_encoding(back, run_data, starts, anchors, L, S_, int(count.item()), n, uncompressed_size=4, position_offset_size=4, relative=False, library=library)
    report("RPE uint32 absolute positions, same column", n * 4, found * 8 + S_ * 4, timed(run, args.repeats), bool(torch.equal(back[:n], column)))
    del column, back, run_data, run_lengths, starts
    torch.cuda.empty_cache()
    # cfg4: Delta int64 -> uint32 deltas, then DZB-variable of the deltas (2-bit size fields), L = 1024
    n = (1 << 30) >> args.shrink
    deltas, stream, packed, position_anchors, total = C.synthesize_dzbv_deltas(device, gen, n, L)
    column = torch.cumsum(deltas, 0) + 12345
    del stream, packed, position_anchors
    deltas32 = deltas.to(torch.int32)
    del deltas
    out, anchor_values = torch.empty(n, dtype=torch.int32, device=device), torch.empty(n // L, dtype=torch.int64, device=device)
    run = lambda: K.delta(out, anchor_values, column, n, L, uncompressed_size=8, compressed_size=4, base_value=12345, status=status, library=library)
    run()
    report("cfg4 Delta int64 -> uint32 deltas, L=1024 (n=2^%d)" % (30 - args.shrink), n * 8, n * 4 + n // L * 8, timed(run, args.repeats), bool(torch.equal(out, deltas32)))
    del column
    data, sizes = torch.empty(total + 64, dtype=torch.uint8, device=device), torch.empty(n // 16, dtype=torch.int32, device=device)
    byte_anchors, num_bytes = torch.empty(n // L, dtype=torch.int32, device=device), torch.zeros(1, dtype=torch.int64, device=device)
    run = lambda: K.discard_zero_bytes_variable(data, sizes, byte_anchors, num_bytes, out, n, L, 1, 2, total + 64, uncompressed_size=4, container_size=4, status=status, library=library)
    run()
    back = torch.empty(n, dtype=torch.int32, device=device)
    S.discard_zero_bytes_variable(back, data, sizes, byte_anchors, L, n, 1, 2, uncompressed_size=4, container_size=4, library=library)
    report("cfg4 DZB-variable of the deltas, 2-bit sizes (%.2f B/delta)" % (total / n), n * 4, total + n // 4 + n // L * 4, timed(run, args.repeats),
           bool(torch.equal(back, deltas32)) and int(num_bytes.item()) == total)
    del data, sizes, back, out, deltas32
    torch.cuda.empty_cache()
    # cfg3: Dictionary lookup int64 -> uint8, 256 entries
    n = (1 << 29) >> args.shrink
    dictionary = torch.sort(torch.randint(-(1 << 62), 1 << 62, (256,), generator=gen, device=device, dtype=torch.int64).view(torch.int64)).values
    # (the kernel compares unsigned: sort the entries as unsigned 64-bit numbers)
    order = torch.argsort(dictionary.view(torch.int64) ^ (-(1 << 63)))
    dictionary = dictionary[order]
    indices = torch.randint(0, 256, (n,), generator=gen, device=device, dtype=torch.int16).to(torch.uint8)
    column = dictionary[indices.to(torch.int64)]
    out = torch.empty(n, dtype=torch.uint8, device=device)
    run = lambda: K.dictionary(out, column, dictionary, 256, n, uncompressed_size=8, dictionary_index_size=1, status=status, library=library)
    run()
    report("cfg3 Dictionary lookup int64 -> uint8, 256 entries (n=2^%d)" % (29 - args.shrink), n * 8, n, timed(run, args.repeats), bool(torch.equal(out, indices)))
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump({"peak_gbs": peak, "rows": rows}, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
