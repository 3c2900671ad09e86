#!/bin/bash
# Evidence run: parity tests, bench (with CPU baseline), reference arm, ncu launch list and --set full captures.
TAG=${1:-r01g}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
lscpu | head -20 > gpurun_out/${TAG}_lscpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; tail -3 gpurun_out/${TAG}_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"
cat gpurun_out/${TAG}_bench.json
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; echo "reference arm exit $?"
cat gpurun_out/${TAG}_bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_raw.csv \
    python bench.py --profile --steps 1 --warmup 1 --pairs-per-gpu 1 > gpurun_out/${TAG}_profile_run.log 2>&1; echo "ncu launches exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:verify_kernel -s 1 -c 1 -o gpurun_out/${TAG}_verify_kernel_h10000 -f \
    python tools/run_verify.py 2000000 10000 2 > gpurun_out/${TAG}_ncu_verify.log 2>&1; echo "ncu K5 exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:refine_cluster_kernel -s 70 -c 1 -o gpurun_out/${TAG}_refine_cluster_kernel -f \
    python bench.py --profile --steps 1 --warmup 1 --pairs-per-gpu 1 > gpurun_out/${TAG}_ncu_refine.log 2>&1; echo "ncu refine exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:score_candidates_kernel -s 60 -c 2 -o gpurun_out/${TAG}_score_candidates_kernel -f \
    python bench.py --profile --steps 1 --warmup 1 --pairs-per-gpu 1 > gpurun_out/${TAG}_ncu_score.log 2>&1; echo "ncu score exit $?"
