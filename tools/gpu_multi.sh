#!/bin/bash
# Multi-GPU check: NCCL parity test + benches at 1..N ranks (run under `gpurun --gpus N`).
N=${N:-2}; TAG=${TAG:-r1}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus_n${N}_$TAG.txt
[ -z "$SKIP_TEST" ] && { timeout 300 python -m pytest tests -m gpu -x -q -k "nccl" --timeout 200 > gpurun_out/pytest_nccl_n${N}_$TAG.log 2>&1; tail -3 gpurun_out/pytest_nccl_n${N}_$TAG.log; }
run() {  # run <n> <outfile> <bench args...>
  local n=$1 out=$2; shift 2
  if [ $n -eq 1 ]; then
    timeout 300 python bench.py --gpus 1 "$@" > $out 2> ${out%.json}.err
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n "$@" > $out 2> ${out%.json}.err
  fi
  tail -1 $out | cut -c1-420; tail -2 ${out%.json}.err
}
for n in 1 2 4 8; do
  [ $n -le $N ] || continue
  [ -n "$ONLY_N" ] && [ "$ONLY_N" != "$n" ] && continue
  run $n gpurun_out/bench_retrieve_n${n}_$TAG.json --retrieve-only --steps 200 --warmup 5 --no-cpu-baseline --no-gpu-reference
  run $n gpurun_out/bench_read_n${n}_$TAG.json --steps 20 --warmup 3 --no-cpu-baseline
  [ -n "$TRAIN" ] && run $n gpurun_out/bench_train_n${n}_$TAG.json --train --steps 8 --warmup 3 --no-cpu-baseline
done
