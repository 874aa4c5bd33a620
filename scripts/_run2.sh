set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r01b_pytest.txt; tail -5 gpurun_out/r01b_pytest.txt
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "test_matches_oracle and (dzbv or bitmaps) and not limits" 2>&1 | tail -25 > gpurun_out/r01b_memcheck.txt; tail -8 gpurun_out/r01b_memcheck.txt
timeout 900 python benchmarks/sweep.py --log2-n 27 --only DZB,bitmaps,Dictionary --out gpurun_out/r01b_sweep.json 2>&1 | tee gpurun_out/r01b_sweep.txt
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:variable_decompress|bitmaps_decompress" -f -o gpurun_out/r01b_dzbv_bitmaps \
	python benchmarks/sweep.py --log2-n 26 --single-launch --no-reference --only "DZB-variable,bitmaps" > gpurun_out/r01b_ncu.log 2>&1
