#!/bin/bash
# round-2 run A: parity suite + seed sweep (both extract() rules) + a short bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02a_pytest.log 2>&1; tail -15 gpurun_out/r02a_pytest.log
timeout 900 python tools/seed_sweep.py --margins 1.0,1.25 --planes --out gpurun_out/r02a_seed_sweep.json > gpurun_out/r02a_seed_sweep.log 2>&1; grep "within" gpurun_out/r02a_seed_sweep.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; echo "bench exit $?"; cat gpurun_out/r02a_bench.json | head -c 3000
