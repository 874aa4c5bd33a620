set -u
mkdir -p gpurun_out
nvidia-smi -L | head -8
for N in 2 4 8; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 100 --warmup 5 --skip-cpu --skip-reference-kernels > gpurun_out/r01e_bench_n$N.json 2> gpurun_out/r01e_bench_n$N.err
  tail -c 1500 gpurun_out/r01e_bench_n$N.json | cut -c1-400; echo
done
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 benchmarks/configs.py --no-reference --out gpurun_out/r01e_configs_8gpu.json 2>&1 | grep -v Warning | tail -40
