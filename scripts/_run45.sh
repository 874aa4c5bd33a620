# compute-sanitizer over the smoke cases (RLE, FOR, Delta anchored + look-back, DZB-variable, patched Dictionary, host-resident scan)
set -u
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
	echo "== $tool"
	timeout 600 compute-sanitizer --tool $tool --error-exitcode 1 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r01final_sanitizer_$tool.txt 2>&1
	echo "exit code $?"; grep -c "smoke: " gpurun_out/r01final_sanitizer_$tool.txt; grep "ERROR SUMMARY\|RACECHECK SUMMARY\|error\b" gpurun_out/r01final_sanitizer_$tool.txt | head -5
done
