#!/bin/bash
set -x
mkdir -p gpurun_out
python tools/gpu_prof_varlen.py > gpurun_out/varlen_timing_r2l.json 2>&1; cat gpurun_out/varlen_timing_r2l.json
NCU=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_varlen -s 3 -c 1 -f -o gpurun_out/prof_varlen_r2l python tools/gpu_prof_varlen.py > gpurun_out/ncu_varlen_r2l.log 2>&1; tail -2 gpurun_out/ncu_varlen_r2l.log
