#!/bin/bash
# Evidence run of round 2: launch list of one registration + ncu --set full captures of the kernels the step spends its time in.
TAG=${1:-r02}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
lscpu | head -20 > gpurun_out/${TAG}_lscpu.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_raw.csv \
    python bench.py --profile --steps 1 --warmup 1 --pairs-per-gpu 1 > gpurun_out/${TAG}_profile_run.log 2>&1; echo "ncu launches exit $?"
cap() {   # name, kernel regex, skip, command...
  local name=$1 rx=$2 skip=$3; shift 3
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -o gpurun_out/${TAG}_$name -f "$@" > gpurun_out/${TAG}_ncu_$name.log 2>&1; echo "ncu $name exit $?"
}
cap accept_loop_kernel accept_loop_kernel 5 python bench.py --profile --steps 1 --warmup 1 --pairs-per-gpu 1
cap score_candidates_kernel score_candidates_kernel 6 python bench.py --profile --steps 1 --warmup 1 --pairs-per-gpu 1
cap score_points_kernel score_points_kernel 6 python bench.py --profile --steps 1 --warmup 1 --pairs-per-gpu 1
cap match_kernel match_kernel 1 python tools/run_match.py 118000 118000
cap verify_kernel_config4 verify_kernel 1 python tools/run_verify.py 5000000 10000 2 24
ls -la gpurun_out/${TAG}_*.ncu-rep
