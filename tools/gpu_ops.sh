#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py -x -q "$@" > gpurun_out/pytest_ops.log 2>&1; tail -30 gpurun_out/pytest_ops.log
