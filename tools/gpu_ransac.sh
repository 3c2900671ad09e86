#!/bin/bash
# RANSAC iteration loop: parity tests, timing marks, old path vs cluster path on the same box
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for v in 0 1; do
  if [ $v = 1 ]; then export PLADE_NO_CLUSTER_REFINE=1; echo "== multi-kernel path"; else unset PLADE_NO_CLUSTER_REFINE; echo "== cluster path"; fi
  PLADE_TIMING=1 timeout 300 python bench.py --profile --steps 3 --warmup 2 --pairs-per-gpu 1 2>&1 | grep "plade timing\|plade ransac\|profile_run" | tail -4
  timeout 300 python tools/concurrency_probe.py 2000000 1,4 5 2>&1 | grep "B="
done
