set -u
timeout 900 python -m pytest tests -m gpu -x -q -k "delta or scan" 2>&1 | tail -3
timeout 900 python benchmarks/sweep.py --only "look-back" --out gpurun_out/r01af_sweep.json 2>&1 | grep -v Warn
