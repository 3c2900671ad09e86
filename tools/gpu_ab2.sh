#!/bin/bash
mkdir -p gpurun_out
for rep in 1 2; do
for v in 0 1; do
  if [ $v = 1 ]; then export PLADE_NO_CLUSTER_REFINE=1; echo "== multi-kernel path (rep $rep)"; else unset PLADE_NO_CLUSTER_REFINE; echo "== cluster path (rep $rep)"; fi
  timeout 300 python tools/concurrency_probe.py 2000000 1,4 15 2>&1 | grep "B="
done
done
unset PLADE_NO_CLUSTER_REFINE
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01f_launches_raw.csv \
    python bench.py --profile --steps 1 --warmup 1 --pairs-per-gpu 1 > gpurun_out/r01f_profile_run.log 2>&1
echo "ncu exit $?"
