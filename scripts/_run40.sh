set -u
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/r01final_bench_2gpu.json 2> gpurun_out/r01final_bench_2gpu.err
echo "2-GPU stdout lines: $(wc -l < gpurun_out/r01final_bench_2gpu.json)"; cut -c1-120 gpurun_out/r01final_bench_2gpu.json
CUDA_VISIBLE_DEVICES=0 timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/r01final_bench_check.json 2>/dev/null; echo "1-GPU stdout lines: $(wc -l < gpurun_out/r01final_bench_check.json)"; cut -c1-120 gpurun_out/r01final_bench_check.json
