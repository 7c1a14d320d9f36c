#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_e2e_gpu.py tests/test_formatter_losses.py -q -m gpu "$@" > gpurun_out/pytest_e2e.log 2>&1; tail -40 gpurun_out/pytest_e2e.log
