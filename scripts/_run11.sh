# GPU session r01k: look-back with statically assigned tiles (cooperative launch) + load prefetch: variants, then parity
set -u
mkdir -p gpurun_out
for v in A E W8 W8E R16 R12; do timeout 120 benchmarks/micro/bin/lb_$v $v; done 2>&1 | tee gpurun_out/r01k_lookback_probe.txt
timeout 120 benchmarks/micro/bin/lb_A Adyn 4 0 2>&1 | tee -a gpurun_out/r01k_lookback_probe.txt
timeout 120 benchmarks/micro/bin/lb_A A3 3 2>&1 | tee -a gpurun_out/r01k_lookback_probe.txt
timeout 120 benchmarks/micro/bin/lb_A A2 2 2>&1 | tee -a gpurun_out/r01k_lookback_probe.txt
timeout 900 python -m pytest tests -m gpu -x -q -k "delta" 2>&1 | tail -8 | tee gpurun_out/r01k_pytest_delta.txt
timeout 900 python benchmarks/sweep.py --only "look-back" --no-reference --out gpurun_out/r01k_sweep.json 2>&1 | tee gpurun_out/r01k_sweep.txt
