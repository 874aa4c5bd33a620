# racecheck / synccheck over the product's GPU suite (the reference's own kernels excluded: its Delta kernel races under racecheck)
set -u
mkdir -p gpurun_out
for tool in racecheck synccheck; do
	echo "== $tool"; timeout 900 compute-sanitizer --tool $tool --error-exitcode 1 python -m pytest tests/test_gpu_parity.py tests/test_host_resident_scan.py -x -q -m gpu -k "not reference_kernels" > gpurun_out/r01final_sanitizer_${tool}_product.txt 2>&1; echo "exit code $?"
	grep "passed\|failed\|SUMMARY" gpurun_out/r01final_sanitizer_${tool}_product.txt | cut -c1-160
done
