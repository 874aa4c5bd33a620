#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full capture) into a small committed table.

    python scripts/ncu_summary.py gpurun_out/r01a_schemes.ncu-rep profiles/ncu_r01_schemes.md [--json out.json]

Per profiled launch: kernel, grid x block, registers, duration, DRAM bytes read / written, DRAM
throughput (% of ncu's peak), achieved occupancy, issue-slot utilisation, L1/L2 hit rates and the
three largest warp-stall reasons (stalled warps per issue-active cycle).  Reads the report with
`ncu -i ... --page raw --csv` (works without a GPU)."""
import csv
import io
import json
import re
import subprocess
import sys


def to_bytes(value, unit):
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)
    return float(value.replace(",", "")) * scale


def to_us(value, unit):
    scale = {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6, "usecond": 1, "msecond": 1e3, "nsecond": 1e-3, "second": 1e6}.get(unit, 1)
    return float(value.replace(",", "")) * scale


def main():
    report, out = sys.argv[1], sys.argv[2]
    json_out = sys.argv[sys.argv.index("--json") + 1] if "--json" in sys.argv else None
    text = subprocess.run(["ncu", "-i", report, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(text)))
    header, units, launches = rows[0], rows[1], rows[2:]
    col = {name: i for i, name in enumerate(header)}

    def get(row, name, default=""):
        return row[col[name]] if name in col else default

    stall_columns = [h for h in header if re.match(r"smsp__average_warps_issue_stalled_(.*)_per_issue_active\.ratio", h)
                     and "selected" not in h]
    records = []
    for row in launches:
        name = re.sub(r"\(.*", "", get(row, "Kernel Name")).replace("void ", "").strip()
        stalls = sorted(((float(row[col[h]] or 0), re.match(r"smsp__average_warps_issue_stalled_(.*)_per_issue_active", h).group(1))
                         for h in stall_columns), reverse=True)[:3]
        read = to_bytes(get(row, "dram__bytes_read.sum"), units[col["dram__bytes_read.sum"]])
        write = to_bytes(get(row, "dram__bytes_write.sum"), units[col["dram__bytes_write.sum"]])
        duration = to_us(get(row, "gpu__time_duration.sum"), units[col["gpu__time_duration.sum"]])
        records.append({
            "kernel": name, "grid": get(row, "launch__grid_size"), "block": get(row, "launch__block_size"),
            "registers": get(row, "launch__registers_per_thread"),
            "dynamic_shared_memory": get(row, "launch__shared_mem_per_block_dynamic"),
            "duration_us": duration, "dram_read_bytes": read, "dram_write_bytes": write,
            "dram_gbs": (read + write) / duration / 1e3,
            "dram_throughput_pct": float(get(row, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed") or 0),
            "warps_active_pct": float(get(row, "sm__warps_active.avg.pct_of_peak_sustained_active") or 0),
            "issue_active_pct": float(get(row, "smsp__issue_active.avg.pct_of_peak_sustained_active") or 0),
            "l1_hit_pct": float(get(row, "l1tex__t_sector_hit_rate.pct") or 0),
            "l2_hit_pct": float(get(row, "lts__t_sector_hit_rate.pct") or 0),
            "top_stalls": [{"reason": r, "warps_per_issue": v} for v, r in stalls],
        })
    with open(out, "w") as f:
        f.write("Source: `%s` (ncu --set full --clock-control none; cold-cache, serialised replays: compare shares and byte counts, not absolute times)\n\n" % report)
        f.write("| kernel | grid x block | regs | dyn smem | time us | DRAM read MB | DRAM write MB | DRAM GB/s | DRAM % of ncu peak | warps active % | issue active % | L1 hit % | L2 hit % | top stalls (warps per issue) |\n")
        f.write("|---|---|---|---|---|---|---|---|---|---|---|---|---|---|\n")
        for r in records:
            f.write("| `%s` | %s x %s | %s | %s | %.1f | %.1f | %.1f | %.0f | %.1f | %.1f | %.1f | %.1f | %.1f | %s |\n" % (
                r["kernel"], r["grid"], r["block"], r["registers"], r["dynamic_shared_memory"], r["duration_us"], r["dram_read_bytes"] / 1e6,
                r["dram_write_bytes"] / 1e6, r["dram_gbs"], r["dram_throughput_pct"], r["warps_active_pct"], r["issue_active_pct"],
                r["l1_hit_pct"], r["l2_hit_pct"], ", ".join("%s %.1f" % (s["reason"], s["warps_per_issue"]) for s in r["top_stalls"])))
    if json_out:
        with open(json_out, "w") as f:
            json.dump(records, f, indent=1)
    print("wrote %s (%d launches)" % (out, len(records)))


if __name__ == "__main__":
    main()
