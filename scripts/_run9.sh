# GPU session r01i: look-back v2 (self-validating descriptors, 128-wide window, 8-row tiles)
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r01i_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q -k "delta" 2>&1 | tail -8 | tee gpurun_out/r01i_pytest_delta.txt
timeout 900 python benchmarks/sweep.py --only "delta" --out gpurun_out/r01i_sweep.json 2>&1 | tee gpurun_out/r01i_sweep.txt
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:delta_decompress_long" -c 3 -f -o gpurun_out/r01i_lookback \
	python benchmarks/sweep.py --log2-n 26 --single-launch --no-reference --only "look-back" > gpurun_out/r01i_ncu.log 2>&1
ls -la gpurun_out | tail -8
