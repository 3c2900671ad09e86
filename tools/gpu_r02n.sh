#!/bin/bash
# host-core scarcity on one GPU: 4 cores per GPU is what the 8-GPU box has (32 vCPUs)
mkdir -p gpurun_out
run() { # label, cores, extra args
  local label=$1 cores=$2; shift 2
  timeout 300 taskset -c $cores python bench.py --steps 8 --warmup 3 --skip-cpu-baseline --skip-config4 "$@" > gpurun_out/r02n_$label.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/r02n_$label.json')); print('$label', round(d['value'],1), 'pairs/s; e2e', round(d['e2e']['value'],1), 'cores', d['run']['host_cores'], 'blocking', d['run']['blocking_sync'], 'B', d['run']['pairs_per_step_per_gpu'])"
}
run all16_B4 0-15
run c4_B4_block 0-3 --blocking-sync 1
run c4_B4_spin 0-3 --blocking-sync 0
run c4_B2_block 0-3 --pairs-per-gpu 2 --blocking-sync 1
run c4_B2_spin 0-3 --pairs-per-gpu 2 --blocking-sync 0
run c4_B1_spin 0-3 --pairs-per-gpu 1 --blocking-sync 0
run c8_B4_block 0-7 --blocking-sync 1
run c8_B4_spin 0-7 --blocking-sync 0
