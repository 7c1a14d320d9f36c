#!/bin/bash
# CTA-pair GEMM bring-up: its parity tests under a short limit, then timings, then the bench with it on.
set -x
mkdir -p gpurun_out
TAG=${TAG:-r1d}
timeout -k 5 150 python -m pytest tests/test_gemm_pair_gpu.py -m gpu -x -q --timeout 60 > gpurun_out/pytest_pair_$TAG.log 2>&1; echo "pair tests rc=$?"; tail -25 gpurun_out/pytest_pair_$TAG.log
timeout -k 5 150 python -m pytest tests/test_blocks_gpu.py tests/test_ops_gpu.py -m gpu -x -q --timeout 60 > gpurun_out/pytest_blocks_$TAG.log 2>&1; echo "blocks/ops tests rc=$?"; tail -5 gpurun_out/pytest_blocks_$TAG.log
if grep -q " passed" gpurun_out/pytest_pair_$TAG.log && ! grep -q "failed\|error" gpurun_out/pytest_pair_$TAG.log; then
  timeout -k 5 200 python tools/gpu_perf_gemm.py > gpurun_out/perf_gemm_$TAG.txt 2>&1; cat gpurun_out/perf_gemm_$TAG.txt
  EMDR2_GEMM_PAIR=1 timeout -k 5 300 python bench.py --steps 10 --no-cpu-baseline > gpurun_out/bench_pair_$TAG.json 2> gpurun_out/bench_pair_$TAG.err; cut -c1-200 gpurun_out/bench_pair_$TAG.json; tail -3 gpurun_out/bench_pair_$TAG.err
  EMDR2_GEMM_PAIR=1 timeout -k 5 200 python -m pytest tests/test_e2e_gpu.py tests/test_backward_gpu.py -m gpu -x -q --timeout 100 > gpurun_out/pytest_e2e_pair_$TAG.log 2>&1; tail -3 gpurun_out/pytest_e2e_pair_$TAG.log
fi
nvidia-smi --query-gpu=name,clocks.sm,power.draw --format=csv
