#!/bin/bash
# Attention kernel change: parity tests under a short limit, then timings.
set -x
mkdir -p gpurun_out
TAG=${TAG:-r1h}
timeout -k 5 200 python -m pytest tests/test_ops_gpu.py tests/test_blocks_gpu.py tests/test_e2e_gpu.py tests/test_backward_gpu.py -m gpu -x -q --timeout 90 > gpurun_out/pytest_attn_$TAG.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/pytest_attn_$TAG.log
timeout -k 5 200 python tools/gpu_perf_attn.py > gpurun_out/perf_attn_$TAG.txt 2>&1; cat gpurun_out/perf_attn_$TAG.txt
timeout -k 5 300 python bench.py --steps 10 --no-cpu-baseline > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; cut -c1-200 gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_%s.json" % __import__("os").environ.get("TAG", "r1h")).read().strip().splitlines()[-1])
print(d["kernel_time_ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"])
PY
