# GPU session r01l: traced look-back probe; dictionary with large shared-memory dictionaries
set -u
mkdir -p gpurun_out
for v in T TW1 TR16; do timeout 120 benchmarks/micro/bin/lb_$v $v; done 2>&1 | tee gpurun_out/r01l_lookback_trace.txt
timeout 900 python -m pytest tests -m gpu -x -q -k "dict" 2>&1 | tail -8 | tee gpurun_out/r01l_pytest_dict.txt
timeout 900 python benchmarks/sweep.py --only "dictionary" --out gpurun_out/r01l_sweep.json 2>&1 | tee gpurun_out/r01l_sweep.txt
