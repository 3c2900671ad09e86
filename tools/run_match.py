"""K3c at the P = 40 worst-case shape for ncu captures (never a bench value)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import plade_b200
nq = int(sys.argv[1]) if len(sys.argv) > 1 else 118000
ndb = int(sys.argv[2]) if len(sys.argv) > 2 else 118000
rng = np.random.default_rng(0)
db = rng.uniform(-1, 1, size=(ndb, 8)).astype(np.float32); db[:, 0] = rng.uniform(0, 1, size=ndb)
q = rng.uniform(-1, 1, size=(nq, 8)).astype(np.float32); q[:, 0] = rng.uniform(0, 1, size=nq)
q[: nq // 20] = db[rng.integers(0, ndb, nq // 20)] + rng.normal(0, 0.005, size=(nq // 20, 8)).astype(np.float32)
ctx = plade_b200.Context()
for _ in range(2):
    off, idx, d2 = ctx.match_descriptors(db, q)
print("matches", len(idx))
