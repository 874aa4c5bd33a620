# GPU session r01f: patched (fused decode + patch) decoders -- parity, cfg3/cfg5 timings with a grid sweep, ncu capture
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 900 python -m pytest tests -m gpu -x -q -k "patched" 2>&1 | tail -15 | tee gpurun_out/r01f_pytest_patched.txt
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r01f_pytest_gpu.txt
timeout 900 python benchmarks/configs.py --only cfg3,cfg5 --sweep-grids --out gpurun_out/r01f_configs.json 2>&1 | grep -v Warning | tee gpurun_out/r01f_configs.txt
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:patched|^scatter|compressed_indices_scatter" -c 12 -f -o gpurun_out/r01f_patched \
	python benchmarks/configs.py --only cfg3,cfg5 --repeats 1 --no-reference --shrink 1 --out gpurun_out/r01f_configs_ncu.json > gpurun_out/r01f_ncu.log 2>&1
ls -la gpurun_out | tail
