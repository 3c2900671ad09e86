"""How much does one B200 gain from registering several pairs concurrently (one context + host thread per pair)?
Prints pairs/s for B = 1..Bmax contexts on the same resident 2M pair.  Diagnostic only (never a bench value)."""
import os, sys, threading, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import plade_b200
from plade_b200.synth import make_pair

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
bmax = int(sys.argv[2]) if len(sys.argv) > 2 else 4
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 6
tgt, src, gt = make_pair(n_points=n, n_planes=20, seed=20240611)
devnull = os.open(os.devnull, os.O_WRONLY)
saved = os.dup(1)
ctxs, clouds = [], []
for b in range(bmax):
    c = plade_b200.Context(0)
    ctxs.append(c)
    clouds.append((c.upload(tgt), c.upload(src)))
res = {}
for B in range(1, bmax + 1):
    def work(k):
        for _ in range(reps):
            ctxs[k].register_resident(*clouds[k])
    os.dup2(devnull, 1)
    for k in range(B):
        ctxs[k].register_resident(*clouds[k])
    th = [threading.Thread(target=work, args=(k,)) for k in range(B)]
    t0 = time.perf_counter()
    for t in th: t.start()
    for t in th: t.join()
    dt = time.perf_counter() - t0
    os.dup2(saved, 1)
    res[B] = B * reps / dt
    print("B=%d  %.1f pairs/s  (%.2f ms per pair per context)" % (B, res[B], 1e3 * dt / reps), flush=True)
