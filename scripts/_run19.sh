# GPU session r01r: bench.py (e2e through the native scan executor), cfg5 host-resident scan via the executor (+ graph replay), scan tests
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_host_resident_scan.py -x -q 2>&1 | tail -4
timeout 900 python bench.py > gpurun_out/r01r_bench.json 2> gpurun_out/r01r_bench.err; cat gpurun_out/r01r_bench.json; tail -5 gpurun_out/r01r_bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r01r_bench_reference.json 2>> gpurun_out/r01r_bench.err; cat gpurun_out/r01r_bench_reference.json
timeout 900 python benchmarks/configs.py --only cfg5 --no-reference --out gpurun_out/r01r_cfg5.json 2>&1 | grep -v Warning | tee gpurun_out/r01r_cfg5.txt
timeout 900 python benchmarks/configs.py --only cfg5 --no-reference --graph-replay --out gpurun_out/r01r_cfg5_graph.json 2>&1 | grep -v Warning | tee gpurun_out/r01r_cfg5_graph.txt
