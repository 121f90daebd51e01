#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/small_launches.csv python scripts/bench_small.py 3 > gpurun_out/small_ncu.log 2>&1
python - <<'PY'
import csv, io, collections
rows=[l for l in open('gpurun_out/small_launches.csv') if not l.startswith('==')]
agg=collections.OrderedDict()
for x in csv.DictReader(io.StringIO(''.join(rows))):
    agg.setdefault(x['Kernel Name'][:70],[]).append(float(x['Metric Value'])/1e3)
for k,v in agg.items(): print("%-72s n=%3d med=%8.2f min=%8.2f"%(k,len(v),sorted(v)[len(v)//2],min(v)))
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ctc_compress_kernel -s 2 -c 1 -f -o gpurun_out/r01b_compress python scripts/bench_small.py 3 ctc_compress > gpurun_out/r01b_compress.log 2>&1; echo "exit=$?"
