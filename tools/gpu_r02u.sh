#!/bin/bash
# launch list of the head of the branch (one registration per step, two registrations)
mkdir -p gpurun_out
timeout 45 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02d_launches_raw.csv \
    python bench.py --profile --steps 1 --warmup 1 --pairs-per-gpu 1 > gpurun_out/r02d_profile_run.log 2>&1; echo "ncu launches exit $?"
