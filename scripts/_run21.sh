# GPU session r01t: patched decoders rebuilt around a tile index: parity, sweep, cfg3/cfg5 at full size
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "patched or scatter" 2>&1 | tail -6 | tee gpurun_out/r01t_pytest.txt
timeout 900 python benchmarks/sweep.py --only "patched,scatter" --out gpurun_out/r01t_sweep.json 2>&1 | grep -v Warn | tee gpurun_out/r01t_sweep.txt
timeout 1200 python benchmarks/configs.py --only cfg3,cfg5 --no-reference --out gpurun_out/r01t_configs.json 2>&1 | grep -v Warning | tee gpurun_out/r01t_configs.txt
