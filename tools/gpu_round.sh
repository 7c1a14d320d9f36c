#!/bin/bash
# One gpurun call for the round's evidence: GPU parity tests, smoke, benches (both arms), ncu launch list
# of the default bench command, ncu full captures of the dominant kernels.  TAG names the outputs.
set -x
mkdir -p gpurun_out
TAG=${TAG:-r1}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu_$TAG.txt
lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket" >> gpurun_out/gpu_$TAG.txt
PER_FILE_TIMEOUT=${PER_FILE_TIMEOUT:-200} bash tools/gpu_all.sh > gpurun_out/pytest_summary_$TAG.log 2>&1
cp gpurun_out/pytest_all.log gpurun_out/pytest_gpu_$TAG.log; tail -16 gpurun_out/pytest_summary_$TAG.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; tail -2 gpurun_out/smoke_$TAG.log
timeout 400 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; cut -c1-200 gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
timeout 300 python bench.py --train --steps 8 --no-cpu-baseline > gpurun_out/bench_train_$TAG.json 2>> gpurun_out/bench_$TAG.err; cut -c1-200 gpurun_out/bench_train_$TAG.json
timeout 300 python bench.py --retrieve-only --steps 200 --warmup 5 > gpurun_out/bench_retrieve_$TAG.json 2>> gpurun_out/bench_$TAG.err; cut -c1-200 gpurun_out/bench_retrieve_$TAG.json
timeout 300 python bench.py --retrieve-only --rows 1000000 --dtype bf16 --steps 500 --warmup 5 --no-gpu-reference > gpurun_out/bench_c2_$TAG.json 2>> gpurun_out/bench_$TAG.err; cut -c1-200 gpurun_out/bench_c2_$TAG.json
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>> gpurun_out/bench_$TAG.err; cut -c1-200 gpurun_out/bench_ref_$TAG.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list_$TAG.log 2>&1
python tools/agg_launches.py gpurun_out/launches_$TAG.csv > gpurun_out/launches_agg_$TAG.txt; head -16 gpurun_out/launches_agg_$TAG.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mips_scan -c 1 -f -o gpurun_out/prof_scan_$TAG \
   python bench.py --retrieve-only --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-reference > gpurun_out/ncu_full_$TAG.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"gemm|attention_fwd" -c 6 -f -o gpurun_out/prof_blocks_$TAG \
   python tools/gpu_prof_blocks.py >> gpurun_out/ncu_full_$TAG.log 2>&1
tail -3 gpurun_out/ncu_full_$TAG.log
