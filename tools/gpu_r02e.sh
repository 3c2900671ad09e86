#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02e_pytest.log 2>&1; tail -25 gpurun_out/r02e_pytest.log
PLADE_TIMING=1 timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err; echo "bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/r02e_bench.json')); print(d['value'], d['e2e']['value'], d['stage_ms'], d['result']); print(d['step_kernels'])"
grep "plade ransac" gpurun_out/r02e_bench.err | tail -3
grep "plade timing" gpurun_out/r02e_bench.err | tail -2
