#!/usr/bin/env python
"""Condense the SASS listings `make sass` writes (profiles/sass/*.sass, ~80 MB, git-ignored) into what is committed:

  profiles/sass_<tag>_summary.md   one row per kernel instantiation: registers, and how many of the instructions which
                                   carry the design are in it (128/256-bit global accesses, TMA bulk copies, mbarrier
                                   waits, shuffles, votes, shared-memory atomics, barriers)
  profiles/sass_<tag>/<kernel>.sass.gz the full listing of the instantiations the BASELINE configurations use

    make sass && python scripts/sass_summary.py r01
"""
import io
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATTERNS = [
    ("LDG.128", r"\bLDG\.E(\.\w+)*\.128"), ("LDG.64", r"\bLDG\.E(\.\w+)*\.64"), ("LDG other", r"\bLDG\.E(?!(\.\w+)*\.(128|64))"),
    ("STG.128", r"\bSTG\.E(\.\w+)*\.128"), ("STG.256", r"\bSTG\.E(\.\w+)*\.256"), ("STG other", r"\bSTG\.E(?!(\.\w+)*\.(128|256))"),
    ("UBLKCP (TMA)", r"\bUBLKCP"), ("SYNCS (mbarrier)", r"\bSYNCS"), ("LDS", r"\bLDS"), ("STS", r"\bSTS"), ("ATOMS", r"\bATOMS"),
    ("SHFL", r"\bSHFL"), ("VOTE", r"\bVOTE"), ("REDUX", r"\bREDUX"), ("POPC", r"\bPOPC"), ("BAR", r"\bBAR\."), ("LDL/STL", r"\b(LDL|STL)"),
]
# the instantiations the BASELINE.json configurations (and the headline bench) launch
HEADLINE = [
    "frame_of_reference_decompress<4, 4, 2, 0, 4>", "frame_of_reference_decompress<4, 4, 2, 1, 4>",
    "run_length_encoding_decompress<4, 4, 2, 4>", "run_position_encoding_decompress<4, 4, 4, false, 4>",
    "delta_decompress<4, 8, 4>", "delta_decompress_long_segments_staged<4, 8, 4>", "delta_decompress_long_segments<4, 8, 4>", "dictionary_decompress<4, 8, 1>",
    "discard_zero_bytes_variable_decompress<4, 4, 4>", "delta_of_discard_zero_bytes_variable_decompress<4, 8, 4, 4>",
    "incidence_bitmaps_decompress<4, 4, 2, true>", "scatter<4, 8, 4>", "compressed_indices_scatter<4, 4, 4, 4>",
    "patched_dictionary_decompress<4, 8, 1>", "patched_frame_of_reference_decompress<4, 4, 2, 1, 4>", "expand_compressed_indices<4, 4, 4>", "patched_detail::patch_tile_ranges<4>",
    "patched_detail::expand_and_index_compressed_indices<4, 4, 4>",
    "encoders::run_encode_single_pass<4, unsigned long>", "encoders::run_tile_summaries_fast<4>", "encoders::run_emit_fast<4, unsigned long>", "encoders::delta_compress<4, 8, 4>",
    "encoders::frame_of_reference_compress<4, 4, 2, 0>", "encoders::dictionary_compress<4, 8, 1>", "encoders::variable_width_emit<4, 4>",
]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return [re.sub(r"\(unsigned int\)|\(int\)|\(bool\)|\(unsigned short\)", "", re.sub(r"\(giddy::.*", "", d)).replace("void giddy::kernels::", "") for d in out]


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    source = os.path.join(ROOT, "profiles", "sass")
    out_dir = os.path.join(ROOT, "profiles", "sass_" + tag)
    os.makedirs(out_dir, exist_ok=True)
    rows = []
    for file in sorted(f for f in os.listdir(source) if f.endswith(".sass")):
        registers = {}
        log = os.path.join(source, file.replace(".sass", ".ptxas.txt"))
        if os.path.exists(log):
            for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'.*?(\d+) bytes spill stores.*?Used (\d+) registers", open(log).read(), re.S):
                registers[m.group(1)] = (m.group(3), m.group(2))
        text = open(os.path.join(source, file)).read()
        parts = re.split(r"\n\s*Function : (\S+)\n", text)
        names, bodies = parts[1::2], parts[2::2]
        for mangled, name, body in zip(names, demangle(names), bodies):
            counts = [len(re.findall(p, body)) for _, p in PATTERNS]
            instructions = len(re.findall(r"^\s+/\*[0-9a-f]{4}\*/", body, re.M))
            rows.append((name, registers.get(mangled, ("?", "?")), instructions, counts))
            if any(name.replace("u>", ">").replace("u,", ",") == h for h in HEADLINE):
                safe = re.sub(r"[^A-Za-z0-9]+", "_", name).strip("_")
                import gzip
                with open(os.path.join(out_dir, safe + ".sass.gz"), "wb") as raw, gzip.GzipFile(fileobj=raw, mode="wb", mtime=0) as binary, \
                        io.TextIOWrapper(binary) as f:  # (gzip'ed, reproducibly: listings are evidence, not reading matter)
                    f.write("Function : %s\n(%s; nvcc 12.9 -O3 -gencode arch=compute_100a,code=sm_100a)\n%s" % (name, mangled, body))
    rows.sort()
    with open(os.path.join(ROOT, "profiles", "sass_%s_summary.md" % tag), "w") as f:
        f.write("SASS census of every kernel instantiation in libgiddy_b200.so (`make sass && python scripts/sass_summary.py %s`; "
                "full listings of the BASELINE instantiations in `profiles/sass_%s/`).  Counts are static instruction counts.\n\n" % (tag, tag))
        f.write("| kernel | regs | spill B | instructions | " + " | ".join(n for n, _ in PATTERNS) + " |\n")
        f.write("|---|---|---|---|" + "---|" * len(PATTERNS) + "\n")
        for name, (regs, spills), instructions, counts in rows:
            f.write("| `%s` | %s | %s | %d | %s |\n" % (name, regs, spills, instructions, " | ".join(str(c) for c in counts)))
    print("wrote %d kernels; %d headline listings in %s" % (len(rows), len(os.listdir(out_dir)), out_dir))


if __name__ == "__main__":
    main()
