# GPU session r01g: warp-granular patched decoders, Delta with decoupled look-back -- parity, timings, ncu
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "patched or delta" 2>&1 | tail -15 | tee gpurun_out/r01g_pytest_new.txt
timeout 900 python benchmarks/configs.py --only cfg3,cfg5 --sweep-grids --out gpurun_out/r01g_configs.json 2>&1 | grep -v Warning | tee gpurun_out/r01g_configs.txt
timeout 900 python benchmarks/sweep.py --only "delta,patched" --out gpurun_out/r01g_sweep.json 2>&1 | tee gpurun_out/r01g_sweep.txt
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r01g_pytest_gpu.txt
timeout 600 python tests/golden/make_ref_golden.py gpurun_out/r01g_ref_gpu_digests.json 2>&1 | tail -5
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:patched|delta_decompress_long" -c 8 -f -o gpurun_out/r01g_patched \
	python benchmarks/sweep.py --log2-n 26 --single-launch --no-reference --only "patched,look-back" > gpurun_out/r01g_ncu.log 2>&1
ls -la gpurun_out | tail -12
