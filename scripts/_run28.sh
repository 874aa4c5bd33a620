set -u
timeout 900 python -m pytest tests -m gpu -x -q -k "patched or scatter" 2>&1 | tail -3
timeout 900 python benchmarks/sweep.py --only "patched" --no-reference --out gpurun_out/r01zz_sweep.json 2>&1 | grep -v Warn
timeout 1200 python benchmarks/configs.py --only cfg3,cfg5 --no-reference --out gpurun_out/r01zz_configs.json 2>&1 | grep -v Warning | grep "fused"
