# ncu --set full of the round's final Delta kernels (anchored with 32-byte pieces; look-back with 24-row tiles) and patched decoders
set -u
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k "regex:delta_decompress|patched_dictionary|patched_frame" -c 8 -f -o gpurun_out/r01final_delta_patched \
	python benchmarks/sweep.py --log2-n 27 --single-launch --no-reference --only "cfg4 Delta,Delta u32<-u8 L=1024,look-back,cfg3 fused,cfg5 fused" > gpurun_out/r01final_ncu.log 2>&1
tail -2 gpurun_out/r01final_ncu.log
