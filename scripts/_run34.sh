# GPU session r01ad: look-back whose first poll is an asynchronous copy started during the previous tile's scan-and-store
set -u
mkdir -p gpurun_out
(timeout 120 benchmarks/micro/bin/lb_T T; timeout 120 benchmarks/micro/bin/lb_A A; timeout 120 benchmarks/micro/bin/lb_R16 R16 2; timeout 120 benchmarks/micro/bin/lb_R8 R8 4) 2>&1 | grep -v "^  tile" | tee gpurun_out/r01ad_lookback_probe.txt
timeout 900 python -m pytest tests -m gpu -x -q -k "delta" 2>&1 | tail -4 | tee gpurun_out/r01ad_pytest_delta.txt
timeout 900 python benchmarks/sweep.py --only "look-back" --no-reference --out gpurun_out/r01ad_sweep.json 2>&1 | grep -v Warn | tee gpurun_out/r01ad_sweep.txt
