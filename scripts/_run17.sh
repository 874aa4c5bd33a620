# GPU session r01p: full parity after look-back v3 / ci-scatter v2 / large-dictionary policy; full sweep
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r01p_pytest_gpu.txt
timeout 1200 python benchmarks/sweep.py --out gpurun_out/r01p_sweep.json 2>&1 | grep -v Warn | tee gpurun_out/r01p_sweep.txt
