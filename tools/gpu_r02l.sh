#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02l_pytest.log 2>&1; tail -8 gpurun_out/r02l_pytest.log | cut -c1-300
PLADE_TIMING=1 timeout 600 python bench.py --skip-cpu-baseline --steps 10 > gpurun_out/r02l_bench.json 2> gpurun_out/r02l_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r02l_bench.json')); print(round(d['value'],1), 'pairs/s; e2e', round(d['e2e']['value'],1), d['result'], {k: round(v,2) for k,v in d['stage_ms'].items()}); print('launches/pair', d['gpu_launches']/d['steps']/4); print([ (l['kernel'][:24], l['launches'], round(l['ms'],2)) for l in d['step_kernels']])"
grep "plade ransac" gpurun_out/r02l_bench.err | tail -2 | cut -c1-330
grep "plade timing" gpurun_out/r02l_bench.err | tail -1 | cut -c1-400
for B in 1 8; do timeout 300 python bench.py --steps 8 --warmup 3 --pairs-per-gpu $B --skip-cpu-baseline --skip-config4 > gpurun_out/r02l_bench_B$B.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r02l_bench_B$B.json')); print('B=$B', round(d['value'],1), 'pairs/s; e2e', round(d['e2e']['value'],1), 'lat', round(d['latency_ms_per_pair'],2))"; done
python tools/seed_sweep.py --cases room_decimated,polyhedron,room_full --out gpurun_out/r02l_sweep.json > gpurun_out/r02l_sweep.log 2>&1; grep within gpurun_out/r02l_sweep.log
