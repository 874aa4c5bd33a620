# GPU session r01z: full validation of the round's final kernels: parity, smoke, bench (both arms), sweep, all five configs
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r01z_gpu.txt 2>&1
(nproc; lscpu | head -20) > gpurun_out/r01z_host.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r01z_pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6 | tee gpurun_out/r01z_smoke.txt
timeout 900 python bench.py > gpurun_out/r01z_bench.json 2> gpurun_out/r01z_bench.err; cut -c1-400 gpurun_out/r01z_bench.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r01z_bench_reference.json 2>> gpurun_out/r01z_bench.err; cut -c1-200 gpurun_out/r01z_bench_reference.json
timeout 1200 python benchmarks/sweep.py --out gpurun_out/r01z_sweep.json 2>&1 | grep -v Warn | tee gpurun_out/r01z_sweep.txt
timeout 1800 python benchmarks/configs.py --out gpurun_out/r01z_configs.json 2>&1 | grep -v Warning | tee gpurun_out/r01z_configs.txt
