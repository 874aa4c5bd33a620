#!/bin/bash
# One GPU-box session: parity tests, the headline bench, the per-scheme sweep and the ncu captures.
# Usage (from the repo root, under gpurun):  bash scripts/gpu_round.sh <tag> [tests|bench|sweep|ncu ...]
# Everything lands in gpurun_out/<tag>_*.
set -u
TAG=${1:-r01}; shift || true
WHAT=${*:-tests bench sweep ncu}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_gpu.txt 2>&1
(nproc; lscpu | head -20) > $OUT/${TAG}_host.txt 2>&1
for what in $WHAT; do
case $what in
tests)
	timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $OUT/${TAG}_pytest_gpu.txt
	tail -3 $OUT/${TAG}_pytest_gpu.txt ;;
smoke)
	timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.txt 2>&1; tail -5 $OUT/${TAG}_smoke.txt ;;
bench)
	timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; cat $OUT/${TAG}_bench.json
	timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/${TAG}_bench_reference.json 2>> $OUT/${TAG}_bench.err; cat $OUT/${TAG}_bench_reference.json ;;
sweep)
	timeout 1200 python benchmarks/sweep.py --log2-n 27 --out $OUT/${TAG}_sweep.json 2>&1 | tee $OUT/${TAG}_sweep.txt ;;
ncu)
	# (1) launch list of the bench command: every launch with its device time
	timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $OUT/${TAG}_launches.csv \
		python bench.py --steps 5 --warmup 3 --skip-cpu --skip-verify > $OUT/${TAG}_launches_bench.log 2>&1
	# (2) full capture of the headline kernel
	timeout 900 ncu --set full --clock-control none --import-source on -k regex:run_length_encoding -s 2 -c 2 -f -o $OUT/${TAG}_rle \
		python bench.py --steps 3 --warmup 1 --skip-cpu --skip-e2e --skip-verify > $OUT/${TAG}_ncu_rle.log 2>&1
	# (3) full capture of one launch of every scheme's kernel
	timeout 1500 ncu --set full --clock-control none --import-source on -k "regex:decompress|scatter" -f -o $OUT/${TAG}_schemes \
		python benchmarks/sweep.py --log2-n 26 --single-launch --no-reference > $OUT/${TAG}_ncu_schemes.log 2>&1
	ls -la $OUT | tail -20 ;;
esac
done
