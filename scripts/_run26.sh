set -u
timeout 600 python benchmarks/sweep.py --only "diagnostic,cfg3" --no-reference --out gpurun_out/r01y.json 2>&1 | grep -v Warn
