#!/bin/bash
# Multi-GPU session (run under `gpurun --gpus N`):  bash scripts/gpu_session_multi.sh <tag> <N> [tests] [example] [bench] [configs]
set -u
TAG=${1:-r02m}; N=${2:-2}; shift 2 || true
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=index,name,pci.bus_id --format=csv > $OUT/${TAG}_gpus.txt 2>&1
nvidia-smi topo -m >> $OUT/${TAG}_gpus.txt 2>&1
(nproc; lscpu | head -25; free -g; cat /sys/devices/system/node/online 2>/dev/null; for d in /sys/bus/pci/devices/*/numa_node; do :; done) > $OUT/${TAG}_host.txt 2>&1
for what in "$@"; do
echo "=== $what ($(date +%T))"
case $what in
tests)
	timeout 900 python -m pytest tests/test_sharding_gpu.py tests/test_cpp_examples.py -m gpu -x -q 2>&1 | tail -15 | tee $OUT/${TAG}_pytest_multi.txt ;;
example)
	timeout 600 examples/bin/sharded_scan 26 $N 2>&1 | tee $OUT/${TAG}_sharded_scan.txt ;;
bench)
	timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 50 --warmup 5 \
		> $OUT/${TAG}_bench_${N}gpu.json 2> $OUT/${TAG}_bench_${N}gpu.err; cut -c1-2500 $OUT/${TAG}_bench_${N}gpu.json; grep -v "^\[\|NCCL\|^$" $OUT/${TAG}_bench_${N}gpu.err | tail -30 ;;
configs)
	timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 benchmarks/configs.py --only cfg4,cfg5,strong \
		--no-reference --out $OUT/${TAG}_configs_${N}gpu.json 2>&1 | grep -v "NCCL\|^$" | tee $OUT/${TAG}_configs_${N}gpu.txt | tail -20 ;;
esac
done
echo "=== done ($(date +%T))"
