set -u
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r01x_cfg35_launches.csv \
	python benchmarks/configs.py --only cfg3,cfg5 --no-reference --repeats 2 > /dev/null 2>&1
python - <<'P'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r01x_cfg35_launches.csv')) if len(r)>10]
h=rows[0]; c={n:i for i,n in enumerate(h)}
agg={}
for r in rows[1:]:
    k=(r[c['Kernel Name']][:60], r[c['Metric Name']])
    agg.setdefault(k,[]).append(float(r[c['Metric Value']].replace(',','')))
for (k,m),v in agg.items():
    if any(s in k for s in ('patched','range','expand','dictionary','frame_of','scatter','memset','Memset')):
        print(k.ljust(62), m.ljust(28), len(v), min(v))
P
