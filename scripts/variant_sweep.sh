#!/bin/bash
# On the GPU box: benchmarks/sweep.py rows matching $1 for the shipped library and every build under benchmarks/variants/
#   bash scripts/variant_sweep.sh <only> [log2-n ...]
ONLY=$1; shift; SIZES=${@:-27}
for n in $SIZES; do
for lib in libgiddy_b200/lib/libgiddy_b200.so benchmarks/variants/*/libgiddy_b200.so; do
	[ -f $lib ] || continue
	echo "== $lib  (log2 n = $n)"
	timeout 300 python benchmarks/sweep.py --log2-n $n --only "$ONLY" --no-reference --library $lib --out /tmp/variant_sweep.json 2>&1 | grep "GB/s" | cut -c1-150
done; done
