#!/bin/bash
mkdir -p gpurun_out
ls /dev/nvidia* | tr '\n' ' '; echo; echo "CVD=$CUDA_VISIBLE_DEVICES"
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02o_pytest.log 2>&1; tail -8 gpurun_out/r02o_pytest.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time timeout 900 python bench.py > gpurun_out/r02o_bench.json 2> gpurun_out/r02o_bench.err ) 2>&1 | grep real; python -c "
import json; d=json.load(open('gpurun_out/r02o_bench.json')); print(round(d['value'],1), 'pairs/s; e2e', round(d['e2e']['value'],1), d['result'], d['run']['wait_mode']); print(json.dumps(d['roofline'])[:500]); print(json.dumps(d['cpu_baseline'])[:300])"
( time timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r02o_bench_reference.json 2> gpurun_out/r02o_bench_reference.err ) 2>&1 | grep real; python -c "
import json; d=json.load(open('gpurun_out/r02o_bench_reference.json')); print('reference arm', round(d['value'],2), 'pairs/s on', d['cpu_baseline']['cores'], 'cores;', d['cpu_baseline']['build'][:40])"
