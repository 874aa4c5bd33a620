# 2-GPU regression: bench.py and the BASELINE configs under torchrun (one process per GPU, NCCL)
set -u
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/r01z_bench_2gpu.json 2> gpurun_out/r01z_bench_2gpu.err; cut -c1-600 gpurun_out/r01z_bench_2gpu.json; tail -3 gpurun_out/r01z_bench_2gpu.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 2>/dev/null | cut -c1-200
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 benchmarks/configs.py --only cfg2,cfg4,cfg5 --no-reference --cfg4-columns 2 --out gpurun_out/r01z_configs_2gpu.json 2>&1 | grep -v Warning | grep "^cfg\|host-resident" | cut -c1-200
