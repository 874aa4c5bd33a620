set -u
timeout 600 python benchmarks/sweep.py --only "FOR int32" --out gpurun_out/r01final_sweep_for.json 2>&1 | grep -v Warn
timeout 600 python benchmarks/sweep.py --only "FOR int32" --out gpurun_out/r01final_sweep_for2.json 2>&1 | grep -v Warn
