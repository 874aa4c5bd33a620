set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "dzbv" 2>&1 | tail -3
timeout 900 python benchmarks/sweep.py --log2-n 27 --only "DZB-variable" --no-reference --out gpurun_out/r01c_sweep.json 2>&1 | tee gpurun_out/r01c_sweep.txt
