"""K3c (descriptor radius matching, match.cu) throughput in pair-distances/s: the room shape
(Q = 2 065 queries x N_db = 32 626) and the worst case P = 40 planes per cloud (1.4e10 pair distances).

    python tools/k3c_bench.py [--out gpurun_out/k3c.json]

The figure is the whole stage call (H2D of the descriptors, match_kernel, canonical ordering of the matches, D2H of the
match list) timed on the host around the C ABI call, plus the algorithmic rate pairs / time; B200 fp64 peak for reference:
~37 TFLOP/s non-tensor fp64 -> 24 flop per pair = 1.5e12 pairs/s as the roofline of an exhaustive scan."""
import argparse, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import plade_b200


def descriptors(rng, n):
    d = rng.uniform(-1, 1, size=(n, 8)).astype(np.float32)
    d[:, 0] = rng.uniform(0, 1, size=n)          # length / scale
    return d


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join("gpurun_out", "k3c.json"))
    a = ap.parse_args()
    ctx = plade_b200.Context(0)
    rng = np.random.default_rng(0)
    rows = []
    for name, nq, ndb in (("room pair (16 + 26 planes)", 2065, 32626), ("P = 40 worst case", 118000, 118000)):
        db = descriptors(rng, ndb)
        q = descriptors(rng, nq)
        q[: nq // 20] = db[rng.integers(0, ndb, nq // 20)] + rng.normal(0, 0.005, size=(nq // 20, 8)).astype(np.float32)   # some true matches
        for _ in range(2):
            off, idx, d2 = ctx.match_descriptors(db, q)
        t0 = time.perf_counter()
        reps = 5
        for _ in range(reps):
            off, idx, d2 = ctx.match_descriptors(db, q)
        dt = (time.perf_counter() - t0) / reps
        pairs = float(nq) * ndb
        rows.append({"case": name, "queries": nq, "db": ndb, "pair_distances": pairs, "matches": int(len(idx)), "ms_per_call": dt * 1e3,
                     "pair_distances_per_s": pairs / dt, "fp64_flop_per_s_if_no_early_exit": 24 * pairs / dt})
        print(rows[-1], flush=True)
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    json.dump(rows, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
