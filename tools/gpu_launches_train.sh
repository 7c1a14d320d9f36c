#!/bin/bash
# ncu launch list (per-launch device time) of a short TRAINING bench; aggregated by kernel name and, for the
# attention kernels, by launch shape.
set -x
mkdir -p gpurun_out
TAG=${TAG:-r2C}
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_train_$TAG.csv \
   python bench.py --train --steps 1 --warmup 1 --no-cpu-baseline --no-gpu-reference --rows 2625000 > gpurun_out/ncu_list_train_$TAG.log 2>&1
python tools/agg_launches.py gpurun_out/launches_train_$TAG.csv > gpurun_out/launches_train_agg_$TAG.txt; head -40 gpurun_out/launches_train_agg_$TAG.txt
python - <<PY
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/launches_train_$TAG.csv')) if len(r)>5]
hdr=None
for i,r in enumerate(rows):
    if 'Kernel Name' in r: hdr=r; rows=rows[i+1:]; break
ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); gi=hdr.index('Grid Size')
att=[(r[ki].split('::')[-1][:48], r[gi], float(r[vi].replace(',',''))) for r in rows if 'attention' in r[ki]]
half=att[len(att)//2:]          # the second (timed) step
agg=collections.defaultdict(lambda:[0,0.0])
for n,g,v in half:
    agg[(n,g, round(v/1000/25))][0]+=1; agg[(n,g, round(v/1000/25))][1]+=v
for k,(c,t) in sorted(agg.items(), key=lambda x:-x[1][1])[:40]: print(k[0],k[1],c,round(t/1e6,2),'ms total', round(t/c/1000,1),'us each')
PY
rm -f gpurun_out/launches_train_$TAG.csv.tmp
