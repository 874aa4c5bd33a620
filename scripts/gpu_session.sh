#!/bin/bash
# GPU-box sessions of round 2 (run under gpurun from the repo root):  bash scripts/gpu_session.sh <tag> <step> [<step> ...]
# Everything lands in gpurun_out/<tag>_*.  Steps:
#   newtests   the round's new GPU tests only (baseline sizes, shards on the GPU, examples), stop at first failure
#   tests      the whole GPU suite
#   smoke      __graft_entry__.smoke()
#   bench      bench.py (all legs) and the reference arm
#   variants   benchmarks/configs.py cfg3,cfg5 for the default build and every benchmarks/variants/*/libgiddy_b200.so
#   configs    benchmarks/configs.py, all configurations, with the reference's kernels beside
#   sweep      benchmarks/sweep.py at 2^27
#   ncu_patched / ncu_all   ncu --set full captures
#   launches   ncu launch list of the bench command
set -u
TAG=${1:-r02}; shift || true
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_gpu.txt 2>&1
(nproc; lscpu | head -25; free -g; cat /sys/devices/system/node/online 2>/dev/null) > $OUT/${TAG}_host.txt 2>&1
for what in "$@"; do
echo "=== $what ($(date +%T))"
case $what in
newtests)
	timeout 1500 python -m pytest tests/test_baseline_sizes.py tests/test_sharding_gpu.py tests/test_cpp_examples.py tests/test_host_resident_scan.py -m gpu -x -q 2>&1 | tail -30 > $OUT/${TAG}_pytest_new.txt
	tail -8 $OUT/${TAG}_pytest_new.txt ;;
tests)
	timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > $OUT/${TAG}_pytest_gpu.txt
	tail -5 $OUT/${TAG}_pytest_gpu.txt ;;
smoke)
	timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.txt 2>&1; tail -12 $OUT/${TAG}_smoke.txt ;;
bench)
	timeout 1500 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; cut -c1-3000 $OUT/${TAG}_bench.json; tail -40 $OUT/${TAG}_bench.err
	timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/${TAG}_bench_reference.json 2>> $OUT/${TAG}_bench.err; cut -c1-600 $OUT/${TAG}_bench_reference.json ;;
variants)
	timeout 600 python benchmarks/configs.py --only cfg3,cfg5 --no-reference --out $OUT/${TAG}_variant_default.json 2>&1 | grep -i "fused\|Dictionary u8\|FOR linear" | tee $OUT/${TAG}_variants.txt
	for lib in benchmarks/variants/*/libgiddy_b200.so; do
		name=$(basename $(dirname $lib)); echo "--- $name" | tee -a $OUT/${TAG}_variants.txt
		timeout 600 python benchmarks/configs.py --only cfg3,cfg5 --no-reference --library $lib --out $OUT/${TAG}_variant_$name.json 2>&1 | grep -i "fused\|Dictionary u8\|FOR linear" | tee -a $OUT/${TAG}_variants.txt
	done ;;
patchedtests)
	timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_sharding_gpu.py -m gpu -x -q -k "patched or fused" 2>&1 | tail -15 > $OUT/${TAG}_pytest_patched.txt
	tail -4 $OUT/${TAG}_pytest_patched.txt ;;
launches_patched)
	timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/${TAG}_launches_patched.csv -k "regex:patch|expand|dictionary|frame_of" \
		python benchmarks/configs.py --only cfg3,cfg5 --no-reference --repeats 2 --library ${VARIANT_LIB:-libgiddy_b200/lib/libgiddy_b200.so} > $OUT/${TAG}_launches_patched.log 2>&1
	python scripts/launch_list_summary.py $OUT/${TAG}_launches_patched.csv $OUT/${TAG}_launches_patched.md; cat $OUT/${TAG}_launches_patched.md | cut -c1-220 ;;
ncu_kernel)
	# NCU_KERNEL=<regex> NCU_COUNT=<n> NCU_CMD="python ..." : a full capture of a few launches of one kernel
	timeout 900 ncu --set full --clock-control none --import-source on -k "regex:${NCU_KERNEL}" -c ${NCU_COUNT:-2} -f -o $OUT/${TAG}_kernel \
		${NCU_CMD:-python benchmarks/configs.py --only cfg5 --no-reference --repeats 1} > $OUT/${TAG}_ncu_kernel.log 2>&1
	ls -la $OUT/${TAG}_kernel.ncu-rep ;;
probe)
	timeout 900 python benchmarks/patched_probe.py --out $OUT/${TAG}_patched_probe.json 2>&1 | tee $OUT/${TAG}_patched_probe.txt ;;
configs)
	timeout 1500 python benchmarks/configs.py --out $OUT/${TAG}_configs.json 2>&1 | tee $OUT/${TAG}_configs.txt ;;
launches_compress)
	timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/${TAG}_launches_compress.csv \
		python benchmarks/compress_sweep.py --repeats 1 --out $OUT/${TAG}_compress_ncu.json > $OUT/${TAG}_launches_compress.log 2>&1
	python scripts/launch_list_summary.py $OUT/${TAG}_launches_compress.csv $OUT/${TAG}_launches_compress.md; cut -c1-200 $OUT/${TAG}_launches_compress.md ;;
compress)
	timeout 900 python benchmarks/compress_sweep.py --out $OUT/${TAG}_compress_sweep.json 2>&1 | tee $OUT/${TAG}_compress_sweep.txt ;;
sweep)
	timeout 1500 python benchmarks/sweep.py --log2-n 27 --out $OUT/${TAG}_sweep.json 2>&1 | tee $OUT/${TAG}_sweep.txt | tail -70 ;;
ncu_patched)
	timeout 900 ncu --set full --clock-control none --import-source on -k "regex:patched_|patch_tile|expand_compressed" -c ${NCU_COUNT:-40} -f -o $OUT/${TAG}_patched \
		python benchmarks/configs.py --only cfg3,cfg5 --no-reference --repeats 1 --library ${VARIANT_LIB:-libgiddy_b200/lib/libgiddy_b200.so} > $OUT/${TAG}_ncu_patched.log 2>&1
	ls -la $OUT/${TAG}_patched.ncu-rep ;;
ncu_all)
	timeout 1500 ncu --set full --clock-control none --import-source on -k "regex:decompress|scatter|compress|encode" -f -o $OUT/${TAG}_schemes \
		python benchmarks/sweep.py --log2-n 26 --single-launch --no-reference > $OUT/${TAG}_ncu_schemes.log 2>&1
	ls -la $OUT/${TAG}_schemes.ncu-rep ;;
launches)
	timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $OUT/${TAG}_launches.csv \
		python bench.py --steps 5 --warmup 3 --skip-cpu --skip-verify --configs-repeats 2 > $OUT/${TAG}_launches_bench.log 2>&1
	wc -l $OUT/${TAG}_launches.csv ;;
sanitize)
	# compute-sanitizer over the product-vs-oracle tests, the compressors, the utility kernels and smoke (no reference kernels: their Delta races)
	for tool in memcheck racecheck synccheck; do
		timeout 2400 compute-sanitizer --tool $tool --error-exitcode 7 python -m pytest tests/test_gpu_parity.py tests/test_compressors.py tests/test_utility_kernels.py tests/test_host_resident_scan.py \
			-m gpu -q -x -k "not matches_reference_kernels" > $OUT/${TAG}_sanitize_$tool.txt 2>&1
		echo "$tool: rc=$? $(grep -E 'passed|failed' $OUT/${TAG}_sanitize_$tool.txt | tail -1) | $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $OUT/${TAG}_sanitize_$tool.txt | tail -1)"
	done
	timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_sanitize_smoke.txt 2>&1
	echo "smoke memcheck: rc=$? $(grep -E 'ERROR SUMMARY' $OUT/${TAG}_sanitize_smoke.txt | tail -1)" ;;
*) echo "unknown step $what" ;;
esac
done
echo "=== done ($(date +%T))"
