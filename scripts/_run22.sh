# GPU session r01u: look-back with a control warp (look-back of tile k+1 overlaps scan+store of tile k): probes, parity, sweep
set -u
mkdir -p gpurun_out
for v in T A R8B4 R8B3 R16B2 R16B3; do timeout 120 benchmarks/micro/bin/lb_$v $v; done 2>&1 | grep -v "^  tile" | tee gpurun_out/r01u_lookback_probe.txt
timeout 900 python -m pytest tests -m gpu -x -q -k "delta" 2>&1 | tail -6 | tee gpurun_out/r01u_pytest_delta.txt
timeout 900 python benchmarks/sweep.py --only "delta" --no-reference --out gpurun_out/r01u_sweep.json 2>&1 | grep -v Warn | tee gpurun_out/r01u_sweep.txt
# patched decoders with patch-driven tile ranges: parity, sweep, per-kernel times, full-size configs
timeout 900 python -m pytest tests -m gpu -x -q -k "patched" 2>&1 | tail -4 | tee gpurun_out/r01u_pytest_patched.txt
timeout 900 python benchmarks/sweep.py --only "patched" --no-reference --out gpurun_out/r01u_sweep_patched.json 2>&1 | grep -v Warn | tee gpurun_out/r01u_sweep_patched.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01u_patched_launches.csv \
	python benchmarks/sweep.py --log2-n 27 --single-launch --no-reference --only "patched" > /dev/null 2>&1
grep -v "^==" gpurun_out/r01u_patched_launches.csv | python -c "
import csv,sys
for r in csv.reader(sys.stdin):
    if len(r)>10 and r[0]!='ID': print(r[4][:70].ljust(72), r[-1], r[-2])
" | tail -20
timeout 1200 python benchmarks/configs.py --only cfg3,cfg5 --no-reference --out gpurun_out/r01u_configs.json 2>&1 | grep -v Warning | grep "fused\|two kernels\|ci-scatter 1%" | tee gpurun_out/r01u_configs.txt
