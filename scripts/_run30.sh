# GPU session r01aa: misaligned-buffer parity (element-wise paths of every kernel), workspaces inside scans / graphs, 2-GPU bench stdout
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -k "misaligned or workspace" 2>&1 | tail -15 | tee gpurun_out/r01aa_pytest.txt
