"""How much does one B200 gain from registering several pairs concurrently (one context + host thread per pair)?
Prints pairs/s for B = 1..Bmax contexts on the same resident 2M pair.  Diagnostic only (never a bench value)."""
import os, sys, threading, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes
import plade_b200
from plade_b200.synth import make_pair
if os.environ.get("PLADE_AB_LIB"):       # A/B runs against an older build of the library (tools/_ab/*.so)
    plade_b200.LIB_PATH = os.path.abspath(os.environ["PLADE_AB_LIB"])
    _l = ctypes.CDLL(plade_b200.LIB_PATH)
    for _n in list(plade_b200.SIGNATURES):
        if not hasattr(_l, _n):
            del plade_b200.SIGNATURES[_n]

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
bs = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 2, 4]
bmax = max(bs)
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
for B in bs:
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
