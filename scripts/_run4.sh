set -u
mkdir -p gpurun_out
timeout 1200 python benchmarks/configs.py --shrink ${1:-4} --only ${2:-cfg1,cfg2,cfg3,cfg4,cfg5} ${3:-} --out gpurun_out/r01d_configs.json 2>&1 | tail -60
