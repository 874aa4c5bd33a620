# GPU session r01m: look-back with early aggregates (one barrier per tile); dictionary: single loop instance, large shared-memory dictionaries
set -u
mkdir -p gpurun_out
for v in T A T1024 512 R16 R12; do timeout 120 benchmarks/micro/bin/lb_$v $v; done 2>&1 | tee gpurun_out/r01m_lookback_probe.txt
timeout 900 python -m pytest tests -m gpu -x -q -k "dict or delta" 2>&1 | tail -8 | tee gpurun_out/r01m_pytest.txt
timeout 900 python benchmarks/sweep.py --only "dictionary,look-back" --out gpurun_out/r01m_sweep.json 2>&1 | tee gpurun_out/r01m_sweep.txt
