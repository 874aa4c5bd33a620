# compute-sanitizer over parity tests of the kernels written or rewritten in this session
set -u
mkdir -p gpurun_out
echo "== memcheck"; timeout 420 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_parity.py -x -q -k "test_matches_oracle and (patched or delta or scatter or dict)" > gpurun_out/r01final_sanitizer_memcheck_parity.txt 2>&1; echo "exit code $?"; tail -4 gpurun_out/r01final_sanitizer_memcheck_parity.txt | cut -c1-160
echo "== racecheck"; timeout 420 compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests/test_gpu_parity.py -x -q -k "test_matches_oracle[ and (patched or delta-I4-U8-C4 or ci-scatter)" > gpurun_out/r01final_sanitizer_racecheck_parity.txt 2>&1; echo "exit code $?"; tail -4 gpurun_out/r01final_sanitizer_racecheck_parity.txt | cut -c1-160
