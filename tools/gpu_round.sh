#!/bin/bash
# One GPU-box visit: parity tests, bench line, ncu launch list of one registration, ncu --set full of the top kernels.
# Everything lands in gpurun_out/; numbers printed under ncu are never bench values.
set -u
TAG=${1:-r01b}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
lscpu | head -20 > gpurun_out/${TAG}_lscpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/${TAG}_pytest.log
tail -5 gpurun_out/${TAG}_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench exit $?"
cat gpurun_out/${TAG}_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_raw.csv \
    python bench.py --profile --steps 1 --warmup 1 > gpurun_out/${TAG}_profile_run.log 2>&1
echo "ncu launches exit $?"
for k in verify_kernel score_candidates_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -o gpurun_out/${TAG}_$k -f \
      python bench.py --profile --steps 1 --warmup 1 > gpurun_out/${TAG}_ncu_$k.log 2>&1
  echo "ncu $k exit $?"
done
