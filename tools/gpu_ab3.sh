#!/bin/bash
# cluster geometry vs multi-kernel path at B = 1 and B = 4 (diagnostic)
for rep in 1 2; do
for cfg in "16 1024" "8 1024" "16 512" "8 512" "4 1024" "none"; do
  set -- $cfg
  unset PLADE_NO_CLUSTER_REFINE PLADE_REFINE_CLUSTER PLADE_REFINE_THREADS
  if [ "$1" = "none" ]; then export PLADE_NO_CLUSTER_REFINE=1; else export PLADE_REFINE_CLUSTER=$1 PLADE_REFINE_THREADS=$2; fi
  echo "== cluster $cfg (rep $rep)"
  timeout 300 python tools/concurrency_probe.py 2000000 1,4 10 2>&1 | grep "B="
done
done
