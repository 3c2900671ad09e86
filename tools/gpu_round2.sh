#!/bin/bash
TAG=${1:-r01e}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; tail -4 gpurun_out/${TAG}_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 --skip-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"; tail -3 gpurun_out/${TAG}_bench.err
cat gpurun_out/${TAG}_bench.json
timeout 600 python tools/concurrency_probe.py 2000000 1,2,4,6,8 5 2>&1 | grep "B=" | tee gpurun_out/${TAG}_concurrency.txt
