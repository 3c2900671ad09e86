#!/bin/bash
mkdir -p gpurun_out
for R in 0 1; do
  echo "--- detect_resume=$R"
  python tools/seed_sweep.py --resume $R --cases room_decimated,polyhedron,room_full --margins 1.25,2.0 --maxcand 200,1000 --out gpurun_out/r02h_sweep_resume$R.json > gpurun_out/r02h_sweep_resume$R.log 2>&1; grep within gpurun_out/r02h_sweep_resume$R.log
done
for cfg in "detect_margin=1.25 max_candidates=200" "detect_margin=2.0 max_candidates=1000"; do
  set -- $cfg
  PLADE_TIMING=1 timeout 300 python bench.py --steps 8 --warmup 3 --skip-cpu-baseline --skip-config4 --param $1 --param $2 > gpurun_out/r02h_bench.json 2> gpurun_out/r02h_bench.err
  python -c "
import json; d=json.load(open('gpurun_out/r02h_bench.json')); print('$cfg:', round(d['value'],1), 'pairs/s; e2e', round(d['e2e']['value'],1), 'launches/pair', d['gpu_launches']/d['steps']/4, 'stage', {k: round(v,2) for k,v in d['stage_ms'].items()}, d['result']); print(json.dumps(d['roofline'])[:600])"
done
