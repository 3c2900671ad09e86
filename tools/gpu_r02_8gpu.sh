#!/bin/bash
# 8-GPU run: the scaling point N = 8 (+ N = 1 on the same box), config 4 over 8 ranks, config 5 through the CLI, multi-device batch test
mkdir -p gpurun_out
nproc; echo "CVD=$CUDA_VISIBLE_DEVICES"; ls /dev/nvidia* | tr '\n' ' '; echo
timeout 600 python -m pytest tests -m gpu -q -k "batch_mode or nccl" 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline > gpurun_out/r02_bench_1of8.json 2> gpurun_out/r02_bench_1of8.err; echo "bench N=1 exit $?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02_bench_8gpu.json 2> gpurun_out/r02_bench_8gpu.err; echo "bench N=8 exit $?"
python -c "
import json
a=json.load(open('gpurun_out/r02_bench_1of8.json')); b=json.load(open('gpurun_out/r02_bench_8gpu.json'))
print('N=1', round(a['value'],1), 'N=8', round(b['value'],1), 'efficiency', round(b['value']/8/a['value'],3), 'e2e', round(b['e2e']['value'],1), b['run']['wait_mode'])
print('config 4: N=1', round(a['config4_verify_sharded']['ms'],2), 'ms, N=8', round(b['config4_verify_sharded']['ms'],2), 'ms', b['config4_verify_sharded']['best_index'], b['config4_verify_sharded']['best_is_true_transform'])"
timeout 900 python tools/config5_batch.py --pairs 64 --out gpurun_out/r02_config5.json 2>&1 | grep -v "^registration failed\|^two few\|^too few" | tail -8
