# GPU session r01n: library-path look-back probe; full parity suite
set -u
mkdir -p gpurun_out
timeout 300 python benchmarks/micro/lookback_lib.py 2>&1 | grep -v Warning | tee gpurun_out/r01n_lookback_lib.txt
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r01n_pytest_gpu.txt
