"""Dump the planes the GPU path extracts from the decimated room pair (diagnostic; product only, no oracle)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import plade_b200
g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "room_decimated.npz"))
tgt, src = g["tgt"], g["src"]
c = plade_b200.Context(0)
out = {}
for seed in (1, 2, 6):
    c.set_param("seed", seed)
    ok, T = c.register_clouds(src, tgt)
    pt, ps = c.extract_planes(tgt, 10000), c.extract_planes(src, 10000)
    for name, p in (("t", pt), ("s", ps)):
        out["%s%d_off" % (name, seed)] = p.offsets; out["%s%d_idx" % (name, seed)] = p.indices; out["%s%d_par" % (name, seed)] = p.params
    out["T%d" % seed] = T
    print(seed, ok, len(pt), len(ps), sorted(pt.sizes().tolist(), reverse=True), sorted(ps.sizes().tolist(), reverse=True))
np.savez_compressed(os.path.join("gpurun_out", "room_planes_gpu.npz"), **out)
