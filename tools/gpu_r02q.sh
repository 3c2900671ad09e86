#!/bin/bash
# final validation of the round-2 head: GPU tests, smoke, default bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02q_pytest.log 2>&1; tail -4 gpurun_out/r02q_pytest.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time timeout 600 python bench.py > gpurun_out/r02q_bench.json 2> gpurun_out/r02q_bench.err ) 2>&1 | grep real; python -c "
import json; d=json.load(open('gpurun_out/r02q_bench.json')); print(round(d['value'],1), 'pairs/s; e2e', round(d['e2e']['value'],1), d['result'], d['run']['wait_mode']); print(json.dumps(d['roofline'])[:500]); print(json.dumps(d['cpu_baseline'])[:300])"
