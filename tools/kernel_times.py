"""Print the CUDA-event kernel clock of one warm registration of the bench pair (diagnostic)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import plade_b200
from plade_b200.synth import make_pair
tgt, src, gt = make_pair(n_points=2_000_000, n_planes=20, seed=20240611)
c = plade_b200.Context(0)
ht, hs = c.upload(tgt), c.upload(src)
fd = os.dup(1); os.dup2(os.open(os.devnull, os.O_WRONLY), 1)
for _ in range(4):
    c.register_resident(ht, hs)
os.dup2(fd, 1)
k = c.kernel_times("score_candidates")
print("score_candidates: %d launches, %.3f ms total, %.1f us per launch" % (k["launches"], k["ms"], 1e3 * k["ms"] / max(k["launches"], 1)))
