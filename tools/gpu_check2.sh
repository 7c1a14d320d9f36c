#!/bin/bash
# Bounded GPU check: parity tests, smoke, default bench, attention timings, training-step bench,
# ncu launch list of the bench command, ncu full captures of the representative block-kernel shapes.
set -x
mkdir -p gpurun_out
TAG=${TAG:-r1c}
PER_FILE_TIMEOUT=${PER_FILE_TIMEOUT:-200} bash tools/gpu_all.sh --durations=3 > gpurun_out/pytest_summary_$TAG.log 2>&1
cp gpurun_out/pytest_all.log gpurun_out/pytest_gpu_$TAG.log; tail -14 gpurun_out/pytest_summary_$TAG.log
grep -E "^(FAILED|ERROR)|Error|assert" gpurun_out/pytest_all.log | head -20
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; tail -2 gpurun_out/smoke_$TAG.log
timeout 400 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; cut -c1-200 gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
timeout 200 python tools/gpu_perf_attn.py > gpurun_out/perf_attn_$TAG.txt 2>&1; cat gpurun_out/perf_attn_$TAG.txt
timeout 300 python bench.py --train --steps 6 --no-cpu-baseline > gpurun_out/bench_train_$TAG.json 2>> gpurun_out/bench_$TAG.err; cut -c1-200 gpurun_out/bench_train_$TAG.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline --rows 2625000 > gpurun_out/ncu_list_$TAG.log 2>&1
python tools/agg_launches.py gpurun_out/launches_$TAG.csv > gpurun_out/launches_agg_$TAG.txt; head -16 gpurun_out/launches_agg_$TAG.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"gemm_kernel|attention_fwd" -c 6 -f -o gpurun_out/prof_blocks_$TAG \
   python tools/gpu_prof_blocks.py > gpurun_out/ncu_full_$TAG.log 2>&1
tail -3 gpurun_out/ncu_full_$TAG.log
