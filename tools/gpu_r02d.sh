#!/bin/bash
mkdir -p gpurun_out
python tools/seed_sweep.py --batch 1 --cases room_decimated,room_full,polyhedron --margins 1.25,1.5,2.0 --maxcand 200,1000 --out gpurun_out/r02d_seed_sweep.json > gpurun_out/r02d_seed_sweep.log 2>&1; grep within gpurun_out/r02d_seed_sweep.log
echo "--- batch 8 vs 1 on the same cases (margin 1.25)"
python tools/seed_sweep.py --batch 8 --cases room_decimated,polyhedron,synth_150k,synth_1m --seeds 1,2,3 --out gpurun_out/r02d_b8.json > gpurun_out/r02d_b8.log 2>&1; grep "seed" gpurun_out/r02d_b8.log | head -20
python tools/seed_sweep.py --batch 1 --cases room_decimated,polyhedron,synth_150k,synth_1m --seeds 1,2,3 --out gpurun_out/r02d_b1.json > gpurun_out/r02d_b1.log 2>&1; grep "seed" gpurun_out/r02d_b1.log | head -20
PLADE_TIMING=1 timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err; echo "bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/r02d_bench.json')); print(d['value'], d['e2e']['value'], d['stage_ms'], d['result'])"
grep "plade ransac" gpurun_out/r02d_bench.err | tail -4
