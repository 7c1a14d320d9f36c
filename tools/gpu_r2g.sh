#!/bin/bash
set -x
mkdir -p gpurun_out
N=${N:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --no-cpu-baseline > gpurun_out/bench_r2g_n$N.json 2> gpurun_out/bench_r2g_n$N.err; tail -5 gpurun_out/bench_r2g_n$N.err
python - <<PY
import json
l=json.load(open('gpurun_out/bench_r2g_n$N.json'))
print('value',l['value'],'ms',l['ms_per_step'],'e2e',l['e2e']['value'])
print('parity',l['parity'])
print('gpu_reference',l.get('gpu_reference',{}).get('value'))
t=l.get('train_step'); print('train', {k:t[k] for k in ('value','ms_per_step','kernel_time_ms_per_step','gradient_allreduce')} if t else None)
print('ktime',l['kernel_time_ms_per_step'],'roof',l['roofline']['frac'])
PY
timeout 300 python -m pytest tests/test_mips_gpu.py -q -m gpu -k nccl > gpurun_out/pytest_nccl_r2g.log 2>&1; tail -3 gpurun_out/pytest_nccl_r2g.log
