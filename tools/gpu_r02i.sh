#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02i_pytest.log 2>&1; tail -15 gpurun_out/r02i_pytest.log | cut -c1-300
( time timeout 900 python bench.py > gpurun_out/r02i_bench.json 2> gpurun_out/r02i_bench.err ) 2>&1 | grep real; python -c "
import json; d=json.load(open('gpurun_out/r02i_bench.json')); print(round(d['value'],1), 'pairs/s; e2e', round(d['e2e']['value'],1), d['result'], d['run']); print(json.dumps(d.get('config4_verify_sharded'))[:700]); print(json.dumps(d.get('cpu_baseline'))[:500]); print(json.dumps(d['roofline'])[:400])"
( time timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r02i_bench_reference.json 2> gpurun_out/r02i_bench_reference.err ) 2>&1 | grep real; cut -c1-600 gpurun_out/r02i_bench_reference.json
