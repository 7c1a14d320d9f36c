#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_blocks_gpu.py -q "$@" > gpurun_out/pytest_blocks.log 2>&1; tail -30 gpurun_out/pytest_blocks.log
