#!/bin/bash
# quick loop: full GPU tests + timing marks + B=1/4 probe
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
PLADE_TIMING=1 timeout 300 python bench.py --profile --steps 3 --warmup 2 --pairs-per-gpu 1 2>&1 | grep "plade timing\|plade ransac" | tail -3
timeout 300 python tools/concurrency_probe.py 2000000 1,4 15 2>&1 | grep "B="
