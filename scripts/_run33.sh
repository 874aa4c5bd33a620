# GPU session r01-final: the round's final build -- parity, smoke, bench (both arms), launch list, sweep, all five configs
set -u
T=r01final
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${T}_gpu.txt 2>&1
(nproc; lscpu | head -20) > gpurun_out/${T}_host.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/${T}_pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -9 | tee gpurun_out/${T}_smoke.txt
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; cut -c1-300 gpurun_out/${T}_bench.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${T}_bench_reference.json 2>> gpurun_out/${T}_bench.err; cut -c1-200 gpurun_out/${T}_bench_reference.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${T}_launches.csv \
	python bench.py --steps 5 --warmup 3 --skip-cpu --skip-verify > gpurun_out/${T}_launches_bench.log 2>&1
timeout 1200 python benchmarks/sweep.py --out gpurun_out/${T}_sweep.json 2>&1 | grep -v Warn | tee gpurun_out/${T}_sweep.txt
timeout 1800 python benchmarks/configs.py --out gpurun_out/${T}_configs.json 2>&1 | grep -v Warning | tee gpurun_out/${T}_configs.txt
