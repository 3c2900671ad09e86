"""How does the decimated room pair (tests/golden/room_decimated.npz) react to the RANSAC seed?  (diagnostic)"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import plade_b200
from plade_b200.synth import transform_error
g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "room_decimated.npz"))
tgt, src, gt = g["tgt"], g["src"], g["gt"]
diag = float(np.linalg.norm(np.ptp(tgt[:, :3], axis=0)))
c = plade_b200.Context(0)
fd = os.dup(1); dn = os.open(os.devnull, os.O_WRONLY)
for seed in [20240611, 1, 2, 3, 4, 5, 6, 7]:
    c.set_param("seed", seed)
    os.dup2(dn, 1)
    ok, T = c.register_clouds(src, tgt)          # swapped, as the file overload does
    pt, ps = c.extract_planes(src, 10000), c.extract_planes(tgt, 10000)
    os.dup2(fd, 1)
    Ti = np.linalg.inv(T.astype(np.float64)) if ok else T
    rot, tr = transform_error(Ti, gt, diag)
    print("seed %d: ok=%s rot %.2f deg, trans %.4f; planes %d (289K cloud) + %d (94K cloud)" % (seed, ok, rot, tr, len(pt), len(ps)), flush=True)
