#!/bin/bash
# stage-1 scoring on the live candidate slots only (score_live): GPU tests with it on, A/B bench, launch list + ncu capture
mkdir -p gpurun_out
timeout 330 python -m pytest tests -m gpu -q > gpurun_out/r02r_pytest.log 2>&1; tail -8 gpurun_out/r02r_pytest.log | cut -c1-300
show() { python -c "
import json,sys; d=json.load(open(sys.argv[1])); k=[x for x in d['step_kernels'] if x['kernel'].startswith('score_candidates')][0]
print(sys.argv[2], round(d['value'],1), 'pairs/s; e2e', round(d['e2e']['value'],1), d['result']['ok'], '/', d['result']['registrations'], 'ok; K1a', round(k['avg_launch_ms']*1e3,1), 'us per launch; accept loop share', round(d['roofline']['share_of_step'],3), 'latency', round(d['latency_ms_per_pair'],2))" $1 $2; }
timeout 120 python bench.py > gpurun_out/r02r_bench.json 2> gpurun_out/r02r_bench.err; show gpurun_out/r02r_bench.json score_live=1
timeout 120 python bench.py --skip-cpu-baseline --skip-config4 --param score_live=0 > gpurun_out/r02r_bench_off.json 2> gpurun_out/r02r_bench_off.err; show gpurun_out/r02r_bench_off.json score_live=0
timeout 120 python bench.py --skip-cpu-baseline --skip-config4 > gpurun_out/r02r_bench_on2.json 2> gpurun_out/r02r_bench_on2.err; show gpurun_out/r02r_bench_on2.json score_live=1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02r_launches_raw.csv \
    python bench.py --profile --steps 1 --warmup 1 --pairs-per-gpu 1 > gpurun_out/r02r_profile_run.log 2>&1; echo "ncu launches exit $?"
timeout 80 ncu --set full --clock-control none --import-source on -k regex:score_candidates_kernel -s 6 -c 1 -o gpurun_out/r02r_score_candidates_kernel -f \
    python bench.py --profile --steps 1 --warmup 1 --pairs-per-gpu 1 > gpurun_out/r02r_ncu_score.log 2>&1; echo "ncu score exit $?"
