#!/bin/bash
# Wall time of the GPU test suite file by file (run on the GPU box: gpurun -- 'bash tools/gpu_suite_times.sh').
# The driver's GPU test step has a 20 minute limit; this is the check that the suite stays far below it.
mkdir -p gpurun_out
: > gpurun_out/suite_times.log
for f in tests/test_gpu_*.py tests/test_boundary.py; do
  s=$SECONDS
  python -m pytest "$f" -m gpu -q -x --durations=5 > "gpurun_out/suite_$(basename "$f" .py).log" 2>&1
  rc=$?
  echo "$f rc=$rc $((SECONDS - s)) s: $(tail -1 "gpurun_out/suite_$(basename "$f" .py).log")" >> gpurun_out/suite_times.log
done
cat gpurun_out/suite_times.log
