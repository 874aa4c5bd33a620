# compute-sanitizer (memcheck, racecheck, synccheck) over the whole GPU parity suite
set -u
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
	echo "== $tool"; timeout 900 compute-sanitizer --tool $tool --error-exitcode 1 python -m pytest tests/test_gpu_parity.py tests/test_host_resident_scan.py -x -q -m gpu > gpurun_out/r01final_sanitizer_${tool}_suite.txt 2>&1; echo "exit code $?"
	grep "passed\|failed\|SUMMARY" gpurun_out/r01final_sanitizer_${tool}_suite.txt | cut -c1-160
done
