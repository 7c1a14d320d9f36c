#!/bin/bash
# N-GPU training step with 8 vs 4 SMs left to NCCL (forward steps shortened, no refresh block).
mkdir -p gpurun_out
N=${N:-8}
for C in 8 4; do
EMDR2_NCCL_CTAS=$C timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$C bench.py --gpus $N --steps 3 --no-cpu-baseline --no-gpu-reference --refresh-rows 0 > gpurun_out/bench_ncclctas_$C.json 2> gpurun_out/bench_ncclctas_$C.err
python - <<PY
import json
l=json.loads(open('gpurun_out/bench_ncclctas_$C.json').read().splitlines()[-1])
t=l['train_step']
print('NCCL_CTAS=$C value', round(l['value'],1), 'train ms', round(t['ms_per_step'],1), t['kernel_time_ms_per_step']['gemm'], t['kernel_time_ms_per_step']['attention'], 'allreduce alone', t['gradient_allreduce']['exposed_alone_ms'])
PY
done
