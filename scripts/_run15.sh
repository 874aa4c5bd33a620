set -u
mkdir -p gpurun_out
echo "== sweep, no reference"; timeout 600 python benchmarks/sweep.py --only "look-back" --no-reference --out gpurun_out/r01o_a.json 2>&1 | grep -v Warn
echo "== sweep, with reference"; timeout 600 python benchmarks/sweep.py --only "look-back" --out gpurun_out/r01o_b.json 2>&1 | grep -v Warn
echo "== sweep, no reference, only one segment"; timeout 600 python benchmarks/sweep.py --only "one segment" --no-reference --out gpurun_out/r01o_c.json 2>&1 | grep -v Warn
echo "== lib probe"; timeout 300 python benchmarks/micro/lookback_lib.py 2>&1 | grep "limits=None"
