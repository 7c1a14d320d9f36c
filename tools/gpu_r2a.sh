#!/bin/bash
# Round-2 call A: new BASELINE-size parity tests, default bench with the parity key.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu_r2a.txt
free -g | head -2 >> gpurun_out/gpu_r2a.txt; nproc >> gpurun_out/gpu_r2a.txt
timeout 900 python -m pytest tests/test_baseline_configs_gpu.py -x -q -m gpu --durations=8 > gpurun_out/pytest_r2a.log 2>&1; tail -25 gpurun_out/pytest_r2a.log
timeout 400 python bench.py --steps 10 --no-cpu-baseline > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; cut -c1-300 gpurun_out/bench_r2a.json; tail -3 gpurun_out/bench_r2a.err
python -c "import json;print(json.load(open('gpurun_out/bench_r2a.json'))['parity'])"
timeout 300 python bench.py --retrieve-only --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/bench_retrieve_r2a.json 2>> gpurun_out/bench_r2a.err
python -c "import json;print(json.load(open('gpurun_out/bench_retrieve_r2a.json'))['parity'])"
