"""Where does the time go on the random 1 M-point scenes of BASELINE config 5?  (diagnostic)"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import plade_b200
from plade_b200.synth import make_pair, transform_error
c = plade_b200.Context(0)
for seed in range(1000, 1008):
    t, s, gt = make_pair(n_points=1_000_000, n_planes=20, seed=seed)
    diag = float(np.linalg.norm(np.ptp(t[:, :3], axis=0)))
    c.register_clouds(t, s)
    t0 = time.perf_counter()
    ok, T = c.register_clouds(t, s)
    dt = time.perf_counter() - t0
    rot, tr = transform_error(T, gt, diag)
    st = c.stage_times()
    rep = c.last_report() if ok else {}
    print("seed %d: ok=%s rot %.2f  %.1f ms (planes %.1f ms)  planes %s + %s  launches %d" % (seed, ok, rot, dt * 1e3, st["planes"] * 1e3, rep.get("target_planes"), rep.get("source_planes"), c.launch_count()), flush=True)
