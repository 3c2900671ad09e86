#!/bin/bash
# 8-GPU run: the scaling point N = 8 (+ N = 1 on the same box) with the wait modes, config 5 through the CLI
mkdir -p gpurun_out
nproc
timeout 600 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline --skip-config4 > gpurun_out/r02_bench_1of8.json 2> gpurun_out/r02_bench_1of8.err; echo "bench N=1 exit $?"
for variant in "default" "yield --blocking-sync 2"; do
  set -- $variant; tag=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 10 --warmup 3 --skip-config4 "$@" > gpurun_out/r02_bench_8gpu_$tag.json 2> gpurun_out/r02_bench_8gpu_$tag.err; echo "bench N=8 $tag exit $?"
  python -c "
import json
a=json.load(open('gpurun_out/r02_bench_1of8.json')); b=json.load(open('gpurun_out/r02_bench_8gpu_$tag.json'))
print('$tag: N=1', round(a['value'],1), 'N=8', round(b['value'],1), 'efficiency', round(b['value']/8/a['value'],3), 'e2e', round(b['e2e']['value'],1), b['run']['wait_mode'])"
done
timeout 900 python tools/config5_batch.py --pairs 64 --skip-reference --out gpurun_out/r02_config5b.json 2>&1 | tail -8
