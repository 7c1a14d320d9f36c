#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --no-cpu-baseline > gpurun_out/bench_r2f.json 2> gpurun_out/bench_r2f.err; tail -5 gpurun_out/bench_r2f.err
python - <<'PY'
import json
l=json.load(open('gpurun_out/bench_r2f.json'))
print('value',l['value'],'ms',l['ms_per_step'],'e2e',l['e2e']['value'])
print('parity',l['parity']['ids'])
print('gpu_reference',l.get('gpu_reference'))
t=l.get('train_step'); print('train', {k:t[k] for k in ('value','ms_per_step','kernel_time_ms_per_step','gradient_allreduce')} if t else None)
print('ktime',l['kernel_time_ms_per_step'],'roof',l['roofline']['frac'])
PY
