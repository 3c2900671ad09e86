#!/bin/bash
# short run: scoring rounds on the live slots / stage 2 by candidate (score_live = 2) against scoring every slot + bench A/B
mkdir -p gpurun_out
timeout 100 python -m pytest tests -m gpu -q -x -k "live_slots or ransac_planes_vs_reference_golden or ransac_synthetic_ground_truth or end_to_end or accept_loop_on_the_device or kernel_times" > gpurun_out/r02t_pytest.log 2>&1; tail -5 gpurun_out/r02t_pytest.log | cut -c1-300
show() { python -c "
import json,sys; d=json.load(open(sys.argv[1])); k=[x for x in d['step_kernels'] if x['kernel'].startswith('score_candidates')][0]
print(sys.argv[2], round(d['value'],1), 'pairs/s; e2e', round(d['e2e']['value'],1), d['result']['ok'], '/', d['result']['registrations'], 'ok; K1a', round(k['avg_launch_ms']*1e3,1), 'us per launch; latency', round(d['latency_ms_per_pair'],2))" $1 $2; }
timeout 60 python bench.py > gpurun_out/r02t_bench.json 2> gpurun_out/r02t_bench.err; show gpurun_out/r02t_bench.json score_live=2
timeout 60 python bench.py --skip-cpu-baseline --skip-config4 --param score_live=1 > gpurun_out/r02t_bench_1.json 2> gpurun_out/r02t_bench_1.err; show gpurun_out/r02t_bench_1.json score_live=1
timeout 60 python bench.py --skip-cpu-baseline --skip-config4 > gpurun_out/r02t_bench_2b.json 2> gpurun_out/r02t_bench_2b.err; show gpurun_out/r02t_bench_2b.json score_live=2
