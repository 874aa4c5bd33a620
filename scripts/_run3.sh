set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/r01c_pytest.txt; tail -3 gpurun_out/r01c_pytest.txt
timeout 900 python benchmarks/sweep.py --log2-n 27 --only "cfg4" --no-reference --out gpurun_out/r01c_sweep.json 2>&1 | tee gpurun_out/r01c_sweep.txt
