set -u
timeout 900 python -m pytest tests -m gpu -x -q -k "delta" 2>&1 | tail -3
timeout 900 python benchmarks/sweep.py --only "Delta" --no-reference --out gpurun_out/r01ae_sweep.json 2>&1 | grep -v Warn
timeout 1200 python benchmarks/configs.py --only cfg4 --no-reference --out gpurun_out/r01ae_configs.json 2>&1 | grep -v Warning | grep "^cfg"
