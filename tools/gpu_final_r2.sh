#!/bin/bash
# End-of-round evidence on one GPU: default bench line (with cpu_baseline, gpu_reference, train_step, index_refresh),
# the reference arm, and the ncu launch list of the forward bench (kernel shares of the step).
set -x
mkdir -p gpurun_out
TAG=${TAG:-r2I}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu_$TAG.txt
timeout 900 python bench.py > gpurun_out/bench_${TAG}_n1.json 2> gpurun_out/bench_${TAG}_n1.err; cut -c1-300 gpurun_out/bench_${TAG}_n1.json; tail -3 gpurun_out/bench_${TAG}_n1.err
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${TAG}_reference_arm.json 2>> gpurun_out/bench_${TAG}_n1.err; cut -c1-300 gpurun_out/bench_${TAG}_reference_arm.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-train-step --rows 2625000 > gpurun_out/ncu_list_$TAG.log 2>&1
python tools/agg_launches.py gpurun_out/launches_$TAG.csv > gpurun_out/launches_forward_agg_$TAG.txt; head -14 gpurun_out/launches_forward_agg_$TAG.txt
rm -f gpurun_out/launches_$TAG.csv
