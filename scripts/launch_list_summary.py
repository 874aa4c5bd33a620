#!/usr/bin/env python
"""Condense an ncu launch list (`--metrics gpu__time_duration.sum --csv --log-file X`) into per-kernel totals.

    python scripts/launch_list_summary.py gpurun_out/r01a_launches.csv profiles/launches_r01a_bench.md
"""
import collections
import csv
import re
import sys


def main():
    source, out = sys.argv[1], sys.argv[2]
    lines = [l for l in open(source) if l.startswith('"')]
    rows = list(csv.DictReader(lines))
    totals, counts, order = collections.OrderedDict(), collections.Counter(), []
    for r in rows:
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "")
        name = re.sub(r"<.*", "", name) if not name.startswith("giddy") else name
        value = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        us = value * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(unit, 1.0)
        key = (name, r["Grid Size"], r["Block Size"])
        totals[key] = totals.get(key, 0.0) + us
        counts[key] += 1
    whole = sum(totals.values())
    with open(out, "w") as f:
        f.write("Source: `%s` (ncu --metrics gpu__time_duration.sum --clock-control none; per-launch times are cold-cache and serialised: "
                "compare shares)\n\n| kernel | grid | block | launches | total us | mean us | share of all device time |\n|---|---|---|---|---|---|---|\n" % source)
        for key, us in sorted(totals.items(), key=lambda kv: -kv[1]):
            f.write("| `%s` | %s | %s | %d | %.1f | %.1f | %.1f %% |\n" % (key[0], key[1], key[2], counts[key], us, us / counts[key], 100 * us / whole))
    print("wrote", out)


if __name__ == "__main__":
    main()
