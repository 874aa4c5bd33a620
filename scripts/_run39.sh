# 2-GPU regression of the final build: bench.py (stdout must be the one JSON line), reference arm (all host threads), cfg4/cfg5 under torchrun
set -u
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/r01final_bench_2gpu.json 2> gpurun_out/r01final_bench_2gpu.err
echo "stdout lines: $(wc -l < gpurun_out/r01final_bench_2gpu.json)"; cut -c1-200 gpurun_out/r01final_bench_2gpu.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 2>/dev/null > gpurun_out/r01final_bench_reference_2gpu.json; cut -c1-160 gpurun_out/r01final_bench_reference_2gpu.json; grep -o '"cores": [0-9]*' gpurun_out/r01final_bench_reference_2gpu.json | head -1
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29523 benchmarks/configs.py --only cfg4,cfg5 --no-reference --cfg4-columns 2 --out gpurun_out/r01final_configs_2gpu.json 2>&1 | grep -v Warning | grep "^cfg\|host-resident" | cut -c1-200
