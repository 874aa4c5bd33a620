set -u
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:patched_dictionary|patched_frame|frame_of_reference_decompress" -c 4 -f -o gpurun_out/r01w_patched \
	python benchmarks/sweep.py --log2-n 27 --single-launch --no-reference --only "cfg3 fused,cfg5 fused,cfg5 FOR" > gpurun_out/r01w_ncu.log 2>&1
tail -3 gpurun_out/r01w_ncu.log
