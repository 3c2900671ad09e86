#!/bin/bash
# diagnostics: new GPU tests, fine-grained timing marks of one 2M registration, concurrency probe
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "batch or cli" > gpurun_out/probe_pytest.log 2>&1; tail -15 gpurun_out/probe_pytest.log
PLADE_TIMING=1 timeout 600 python bench.py --profile --steps 3 --warmup 2 > gpurun_out/probe_timing.json 2> gpurun_out/probe_timing.err
grep "plade timing\|plade ransac" gpurun_out/probe_timing.err | tail -4
timeout 900 python tools/concurrency_probe.py 2000000 4 6 2>&1 | grep "B=" | tee gpurun_out/probe_concurrency.txt
