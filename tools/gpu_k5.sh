#!/bin/bash
# K5 iteration loop: parity tests of the verification kernel, timing at H = 10000 / 1000, one ncu --set full capture
TAG=${1:-k5}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "verify or sharded" 2>&1 | tail -3
timeout 300 python tools/run_verify.py 2000000 10000 3 2>&1 | grep "H="
timeout 300 python tools/run_verify.py 2000000 1000 3 2>&1 | grep "H=" | tail -1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:verify_kernel -s 1 -c 1 -o gpurun_out/${TAG}_verify_kernel -f python tools/run_verify.py 2000000 1000 2 > gpurun_out/${TAG}_ncu_verify.log 2>&1; tail -1 gpurun_out/${TAG}_ncu_verify.log
