#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "not seed_sweep and not room_pair" > gpurun_out/r02g_pytest.log 2>&1; tail -6 gpurun_out/r02g_pytest.log
python tools/seed_sweep.py --margins 1.25,2.0 --maxcand 200,1000 --out gpurun_out/r02g_seed_sweep.json > gpurun_out/r02g_seed_sweep.log 2>&1; grep within gpurun_out/r02g_seed_sweep.log
for cfg in "detect_margin=1.25 max_candidates=200" "detect_margin=2.0 max_candidates=1000"; do
  set -- $cfg
  PLADE_TIMING=1 timeout 300 python bench.py --steps 8 --warmup 3 --skip-cpu-baseline --param $1 --param $2 > gpurun_out/r02g_bench.json 2> gpurun_out/r02g_bench.err
  python -c "
import json; d=json.load(open('gpurun_out/r02g_bench.json')); print('$cfg:', round(d['value'],1), 'pairs/s; e2e', round(d['e2e']['value'],1), 'launches/pair', d['gpu_launches']/d['steps']/4, 'stage', {k: round(v,2) for k,v in d['stage_ms'].items()}, d['result'])"
  grep "plade ransac" gpurun_out/r02g_bench.err | tail -2 | cut -c1-260
done
python tools/k3c_bench.py --out gpurun_out/r02g_k3c.json 2>&1 | tail -3
