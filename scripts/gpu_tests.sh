#!/bin/bash
# Runs every -m gpu test file in its own process (a sticky CUDA error cannot cascade across files).
# Usage (on the GPU box, via gpurun): bash scripts/gpu_tests.sh [extra pytest args]
mkdir -p gpurun_out
rc=0
for f in tests/test_gpu_*.py; do
  n=$(basename "$f" .py)
  timeout 900 python -m pytest "$f" -m gpu -q --timeout 600 -p no:cacheprovider "$@" > "gpurun_out/$n.log" 2>&1 || rc=1
  echo "== $n: $(tail -1 gpurun_out/$n.log)"
done
exit $rc
