set -u
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:patched_dictionary|dictionary_decompress" -c 3 -f -o gpurun_out/r01ab_dict \
	python benchmarks/sweep.py --log2-n 27 --single-launch --no-reference --only "cfg3 Dictionary,diagnostic: patched Dictionary u64<-u8, no patches,cfg3 fused" > gpurun_out/r01ab_ncu.log 2>&1
tail -3 gpurun_out/r01ab_ncu.log
