# GPU session r01h: re-validation after container re-creation -- full parity, bench line, patched/look-back timings, ncu
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r01h_gpu.txt 2>&1
(nproc; lscpu | head -20) > gpurun_out/r01h_host.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 | tee gpurun_out/r01h_pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6 | tee gpurun_out/r01h_smoke.txt
timeout 900 python bench.py > gpurun_out/r01h_bench.json 2> gpurun_out/r01h_bench.err; cat gpurun_out/r01h_bench.json
timeout 900 python benchmarks/sweep.py --only "delta,patched" --out gpurun_out/r01h_sweep.json 2>&1 | tee gpurun_out/r01h_sweep.txt
timeout 900 python benchmarks/configs.py --only cfg3,cfg5 --sweep-grids --out gpurun_out/r01h_configs.json 2>&1 | grep -v Warning | tee gpurun_out/r01h_configs.txt
timeout 600 python tests/golden/make_ref_golden.py gpurun_out/r01h_ref_gpu_digests.json 2>&1 | tail -5
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:patched|delta_decompress_long" -c 8 -f -o gpurun_out/r01h_patched \
	python benchmarks/sweep.py --log2-n 26 --single-launch --no-reference --only "patched,look-back" > gpurun_out/r01h_ncu.log 2>&1
ls -la gpurun_out | tail -12
