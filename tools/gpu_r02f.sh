#!/bin/bash
mkdir -p gpurun_out
nvcc -O2 -o /tmp/launch_rate tools/micro/launch_rate.cu 2>/dev/null && /tmp/launch_rate 20000 8 | tee gpurun_out/r02f_launch_rate.txt
timeout 1500 python -m pytest tests -m gpu -q -k "not seed_sweep and not room_pair" > gpurun_out/r02f_pytest.log 2>&1; tail -12 gpurun_out/r02f_pytest.log
for B in 1 2 4 8; do
  timeout 300 python bench.py --steps 8 --warmup 3 --pairs-per-gpu $B --skip-cpu-baseline > gpurun_out/r02f_bench_B$B.json 2> gpurun_out/r02f_bench_B$B.err
  python -c "
import json; d=json.load(open('gpurun_out/r02f_bench_B$B.json')); print('B=$B', round(d['value'],1), 'pairs/s; e2e', round(d['e2e']['value'],1), 'launches/pair', d['gpu_launches']/d['steps']/$B, 'planes ms', round(d['stage_ms']['planes'],2), 'lat', round(d['latency_ms_per_pair'],2))"
done
