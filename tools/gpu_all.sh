#!/bin/bash
# Full GPU test suite, one process per test file, each under its own wall-clock limit so that a
# deadlocked kernel costs minutes, not the whole gpurun budget.
mkdir -p gpurun_out
: > gpurun_out/pytest_all.log
for f in tests/test_*.py; do
  echo "=== $f" >> gpurun_out/pytest_all.log
  timeout -k 10 ${PER_FILE_TIMEOUT:-420} python -m pytest "$f" -m gpu -q -x --timeout 180 "$@" >> gpurun_out/pytest_all.log 2>&1
  echo "rc=$? $f" | tee -a gpurun_out/pytest_all.log
done
grep -E "passed|failed|error|rc=" gpurun_out/pytest_all.log | tail -40
