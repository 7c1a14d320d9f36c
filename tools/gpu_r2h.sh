#!/bin/bash
set -x
mkdir -p gpurun_out
N=${N:-1}
if [ "$N" = "1" ]; then
timeout 900 python bench.py --steps 6 --no-cpu-baseline --no-gpu-reference --rows ${ROWS:-21000000} > gpurun_out/bench_r2h_n$N.json 2> gpurun_out/bench_r2h_n$N.err
else
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 6 --no-cpu-baseline --no-gpu-reference > gpurun_out/bench_r2h_n$N.json 2> gpurun_out/bench_r2h_n$N.err
fi
tail -5 gpurun_out/bench_r2h_n$N.err
python - <<PY
import json
l=json.loads(open('gpurun_out/bench_r2h_n$N.json').read().splitlines()[-1])
print('value',l['value'],'ms',l['ms_per_step'])
print('train', l['train_step']['value'], l['train_step']['ms_per_step'])
print('refresh', json.dumps(l.get('index_refresh'), indent=1))
PY
