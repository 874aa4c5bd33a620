# GPU session r01q: host-resident scan executor (C++ example + Python tests)
set -u
mkdir -p gpurun_out
examples/bin/host_resident_scan 2>&1 | tee gpurun_out/r01q_example_scan.txt
timeout 900 python -m pytest tests/test_host_resident_scan.py tests/test_cpp_examples.py -x -q 2>&1 | tail -15 | tee gpurun_out/r01q_pytest_scan.txt
