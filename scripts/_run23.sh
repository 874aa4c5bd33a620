# GPU session r01v: patched decoders, persistent grid + range prefetch; look-back back to the early-aggregate version
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "patched or delta or scatter" 2>&1 | tail -4 | tee gpurun_out/r01v_pytest.txt
timeout 900 python benchmarks/sweep.py --only "patched,look-back" --no-reference --out gpurun_out/r01v_sweep.json 2>&1 | grep -v Warn | tee gpurun_out/r01v_sweep.txt
timeout 1200 python benchmarks/configs.py --only cfg3,cfg5 --no-reference --sweep-grids --out gpurun_out/r01v_configs.json 2>&1 | grep -v Warning | grep "fused\|two kernels\|ci-scatter 1%\|Dictionary u8" | tee gpurun_out/r01v_configs.txt
