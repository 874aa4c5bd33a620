set -u
timeout 900 python -m pytest tests -m gpu -x -q -k "rle or rpe or run_" 2>&1 | tail -3
timeout 900 python benchmarks/sweep.py --only "RLE,RPE" --no-reference --out gpurun_out/r01ag_sweep.json 2>&1 | grep -v Warn
timeout 900 python bench.py --skip-cpu --steps 100 --warmup 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bench', d['value'], d['roofline']['frac'], d['ms_per_step'])"
