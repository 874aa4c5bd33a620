# 8-GPU run of the final build: bench.py (both arms) and cfg2/cfg4 (8 columns, one per GPU)/cfg5 under torchrun
set -u
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 100 --warmup 5 > gpurun_out/r01final_bench_8gpu.json 2> gpurun_out/r01final_bench_8gpu.err; cut -c1-160 gpurun_out/r01final_bench_8gpu.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 bench.py --impl reference --gpus 8 --steps 3 --warmup 1 2>/dev/null > gpurun_out/r01final_bench_reference_8gpu.json; cut -c1-120 gpurun_out/r01final_bench_reference_8gpu.json
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29543 benchmarks/configs.py --only cfg2,cfg4,cfg5 --no-reference --cfg4-columns 8 --out gpurun_out/r01final_configs_8gpu.json 2>&1 | grep -v Warning | grep "^cfg\|host-resident" | cut -c1-200
