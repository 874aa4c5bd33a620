# 4-GPU point of the scaling series (final build)
set -u
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 4 --steps 100 --warmup 5 > gpurun_out/r01final_bench_4gpu.json 2> gpurun_out/r01final_bench_4gpu.err; cut -c1-160 gpurun_out/r01final_bench_4gpu.json
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29553 benchmarks/configs.py --only cfg4,cfg5 --no-reference --cfg4-columns 8 --out gpurun_out/r01final_configs_4gpu.json 2>&1 | grep -v Warning | grep "^cfg4\|host-resident" | cut -c1-200
