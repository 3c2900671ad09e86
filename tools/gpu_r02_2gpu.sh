#!/bin/bash
# 2-GPU checks: NCCL inside the library (in-process two ranks, CLI --shard-verify, torchrun bench with config 4)
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests -m gpu -q -k "nccl or sharded or batch_mode" > gpurun_out/r02_2gpu_pytest.log 2>&1; tail -5 gpurun_out/r02_2gpu_pytest.log
python - <<'PY'
import os, sys, numpy as np
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
from plade_b200.synth import make_pair
from plyio import write_ply
t, s, gt = make_pair(n_points=500000, n_planes=20, seed=11)
write_ply("/tmp/t.ply", t); write_ply("/tmp/s.ply", s)
PY
./plade_b200/plade_b200_cli /tmp/t.ply /tmp/s.ply /tmp/r1.txt > /dev/null 2>&1; echo "cli exit $?"
PLADE_DEVICES=2 ./plade_b200/plade_b200_cli --shard-verify --report /tmp/rep.json /tmp/t.ply /tmp/s.ply /tmp/r2.txt > /tmp/cli2.log 2>&1; echo "cli --shard-verify exit $?"; tail -3 /tmp/cli2.log
diff /tmp/r1.txt /tmp/r2.txt && echo "sharded result == single-GPU result"; cat /tmp/rep.json | cut -c1-400
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/r02_bench_2gpu.json 2> gpurun_out/r02_bench_2gpu.err; echo "bench N=2 exit $?"
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_2gpu.json')); print(round(d['value'],1), 'pairs/s; e2e', round(d['e2e']['value'],1), d['result']); print(json.dumps(d.get('config4_verify_sharded'))[:600])"
