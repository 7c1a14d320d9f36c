#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench (both arms), ncu launch list + full capture of the scan.
set -x
mkdir -p gpurun_out
TAG=${TAG:-r1}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu_$TAG.txt
lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket" >> gpurun_out/gpu_$TAG.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
tail -5 gpurun_out/pytest_gpu_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; tail -2 gpurun_out/smoke_$TAG.log
timeout 600 python bench.py --steps 200 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; cat gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
timeout 600 python bench.py --rows 1000000 --dtype bf16 --steps 500 --warmup 5 --no-gpu-reference > gpurun_out/bench_c2_$TAG.json 2>> gpurun_out/bench_$TAG.err; cat gpurun_out/bench_c2_$TAG.json
timeout 600 python bench.py --impl reference --steps 10 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>> gpurun_out/bench_$TAG.err; cat gpurun_out/bench_ref_$TAG.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:mips_ -c 40 --csv --log-file gpurun_out/launches_$TAG.csv \
   python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-reference > gpurun_out/ncu_list_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mips_scan -s 6 -c 2 -f -o gpurun_out/prof_scan_$TAG \
   python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-reference > gpurun_out/ncu_full_$TAG.log 2>&1
tail -3 gpurun_out/ncu_full_$TAG.log
