set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r01final_pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -8 | tee gpurun_out/r01final_smoke.txt
timeout 900 python bench.py > gpurun_out/r01final_bench.json 2> gpurun_out/r01final_bench.err; cut -c1-200 gpurun_out/r01final_bench.json
