# GPU session r01s: ncu --set full of the kernels touched this session (one launch each), source-level counters imported
set -u
mkdir -p gpurun_out
timeout 1500 ncu --set full --clock-control none --import-source on -k "regex:patched|delta_decompress_long|dictionary_decompress|compressed_indices_scatter" -c 12 -f -o gpurun_out/r01s_kernels \
	python benchmarks/sweep.py --log2-n 26 --single-launch --no-reference --only "patched,look-back,Dictionary u32,cfg3 Dictionary,ci-scatter" > gpurun_out/r01s_ncu.log 2>&1
tail -5 gpurun_out/r01s_ncu.log; ls -la gpurun_out/*.ncu-rep
