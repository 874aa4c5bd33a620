# GPU session r01ac: look-back with tiles staged in shared memory (TMA), 6-8 CTAs per SM
set -u
mkdir -p gpurun_out
(timeout 120 benchmarks/micro/bin/lb_A staged 6 2; timeout 120 benchmarks/micro/bin/lb_A pipelined 3 1; timeout 120 benchmarks/micro/bin/lb_A staged4 4 2
 timeout 120 benchmarks/micro/bin/lb_R16 R16 6 2; timeout 120 benchmarks/micro/bin/lb_R8B8 R8B8 8 2; timeout 120 benchmarks/micro/bin/lb_R12B8 R12B8 8 2; timeout 120 benchmarks/micro/bin/lb_R20B5 R20B5 5 2; timeout 120 benchmarks/micro/bin/lb_R12B8W2 R12B8W2 8 2) 2>&1 | grep -v "^  tile" | tee gpurun_out/r01ac_lookback_probe.txt
timeout 900 python -m pytest tests -m gpu -x -q -k "delta" 2>&1 | tail -6 | tee gpurun_out/r01ac_pytest_delta.txt
timeout 900 python benchmarks/sweep.py --only "look-back" --no-reference --out gpurun_out/r01ac_sweep.json 2>&1 | grep -v Warn | tee gpurun_out/r01ac_sweep.txt
