#!/bin/bash
# ncu launch list (per-launch device time) of a short forward bench; aggregated by kernel name.
set -x
mkdir -p gpurun_out
TAG=${TAG:-r2k}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-train-step --rows 2625000 > gpurun_out/ncu_list_$TAG.log 2>&1
python tools/agg_launches.py gpurun_out/launches_$TAG.csv > gpurun_out/launches_agg_$TAG.txt; head -30 gpurun_out/launches_agg_$TAG.txt
python - <<PY
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/launches_$TAG.csv')) if len(r)>5]
hdr=None
for i,r in enumerate(rows):
    if 'Kernel Name' in r: hdr=r; rows=rows[i+1:]; break
ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); gi=hdr.index('Grid Size')
att=[(r[ki][:40], r[gi], float(r[vi].replace(',',''))) for r in rows if 'attention' in r[ki]]
last=att[-60:]
agg=collections.defaultdict(lambda:[0,0.0])
for n,g,v in att[-(len(att)//5):]:
    agg[(n,g)][0]+=1; agg[(n,g)][1]+=v
for k,(c,t) in sorted(agg.items(), key=lambda x:-x[1][1]): print(k,c,round(t/1000,1),'us total', round(t/c/1000,1),'us each')
PY
