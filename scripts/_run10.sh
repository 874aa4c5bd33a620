# GPU session r01j: look-back variants (micro probe), then ci-scatter v2 parity + timing
set -u
mkdir -p gpurun_out
for v in A B C D E F G; do timeout 120 benchmarks/micro/bin/lb_$v $v; done 2>&1 | tee gpurun_out/r01j_lookback_probe.txt
timeout 120 benchmarks/micro/bin/lb_A A8 8 2>&1 | tee -a gpurun_out/r01j_lookback_probe.txt
timeout 120 benchmarks/micro/bin/lb_D D8 8 2>&1 | tee -a gpurun_out/r01j_lookback_probe.txt
timeout 900 python -m pytest tests -m gpu -x -q -k "scatter or patched" 2>&1 | tail -8 | tee gpurun_out/r01j_pytest_scatter.txt
timeout 900 python benchmarks/sweep.py --only "scatter" --out gpurun_out/r01j_sweep.json 2>&1 | tee gpurun_out/r01j_sweep.txt
