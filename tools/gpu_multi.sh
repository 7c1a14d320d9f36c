#!/bin/bash
# Multi-GPU check: NCCL parity test + bench at N ranks (run under `gpurun --gpus N`).
N=${N:-2}; TAG=${TAG:-r1}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus_n${N}_$TAG.txt
timeout 600 python -m pytest tests -m gpu -x -q -k "nccl" > gpurun_out/pytest_nccl_n${N}_$TAG.log 2>&1; tail -3 gpurun_out/pytest_nccl_n${N}_$TAG.log
for n in $(seq 1 $N); do
  case $n in 1|2|4|8) ;; *) continue;; esac
  if [ $n -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 200 --warmup 5 --no-cpu-baseline --no-gpu-reference > gpurun_out/bench_n${n}_$TAG.json 2> gpurun_out/bench_n${n}_$TAG.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 200 --warmup 5 > gpurun_out/bench_n${n}_$TAG.json 2> gpurun_out/bench_n${n}_$TAG.err
  fi
  tail -1 gpurun_out/bench_n${n}_$TAG.json; tail -2 gpurun_out/bench_n${n}_$TAG.err
done
