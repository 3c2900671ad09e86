"""BASELINE config 5 in small: N synthetic 1 M-point pairs written as PLY files, registered by plade_register_batch
(PLY parsing included) with W workers per GPU.  Prints whole-box pairs/s.  Diagnostic / DESIGN.md figure, not the bench line."""
import os, sys, tempfile, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import plade_b200
from plade_b200.synth import make_pair, transform_error
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from plyio import write_ply

n_pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n_points = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
workers = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [1, 4]
n_gpus = int(sys.argv[4]) if len(sys.argv) > 4 else 1
d = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
pairs, gts = [], []
for k in range(n_pairs):
    t, s, gt = make_pair(n_points=n_points, n_planes=20, seed=1000 + k)
    tp, sp = os.path.join(d, "t%d.ply" % k), os.path.join(d, "s%d.ply" % k)
    write_ply(tp, t); write_ply(sp, s)
    pairs.append((tp, sp)); gts.append((gt, float(np.linalg.norm(np.ptp(t[:, :3], axis=0)))))
devnull, saved = os.open(os.devnull, os.O_WRONLY), os.dup(1)
for W in workers:
    devs = [g for _ in range(W) for g in range(n_gpus)]
    os.dup2(devnull, 1)
    plade_b200.register_batch(pairs[:len(devs)], devices=devs)          # warm-up (contexts, scratch)
    t0 = time.perf_counter()
    ok, T = plade_b200.register_batch(pairs, devices=devs)
    dt = time.perf_counter() - t0
    os.dup2(saved, 1)
    errs = [transform_error(T[k], gts[k][0], gts[k][1]) for k in range(n_pairs)]
    if W == workers[0]:
        print("per pair (seed: ok rot_err_deg): " + ", ".join("%d: %s %.2f" % (1000 + k, "ok" if ok[k] else "FAIL", errs[k][0]) for k in range(n_pairs)), flush=True)
    print("workers/GPU=%d GPUs=%d: %d pairs of %d points in %.2f s = %.1f pairs/s incl. PLY read; ok %d/%d, max rot err %.3f deg"
          % (W, n_gpus, n_pairs, n_points, dt, n_pairs / dt, int(ok.sum()), n_pairs, max(e[0] for e in errs)), flush=True)
for tp, sp in pairs:
    os.unlink(tp); os.unlink(sp)
os.rmdir(d)
