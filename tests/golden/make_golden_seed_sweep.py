"""Seed sweep of the REFERENCE (oracle/_ref) on the end-to-end cases the GPU path is compared with, and the
reference's plane-detection curve near min_support.  Run in the build container (needs /root/reference, oracle/_ref):

    python tests/golden/make_golden_seed_sweep.py [--full-room]

Writes tests/golden/seed_sweep_ref.json:
  cases.<name>.errors      [rot_deg, trans_rel_diag, ok] of the reference for RANSAC seeds 1..8 (time() interposed)
  detection_curve          per (cloud, min_support): every distinct plane the reference found over the seeds, its support and in
                           how many of the 8 runs it was found -> the detection frequency as a function of support / min_support
                           (the measurement behind Params::detect_margin, plade_b200/csrc/ransac.cu extract_planes_dev)
--full-room also writes tests/golden/_local/room_full.npz (the 2.3 M-point room pair of BASELINE config 2, 58 MB: git-ignored,
travels to the GPU box with the snapshot) with the reference's own results on it.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.ref import Ref, load_ply  # noqa: E402
from plade_b200.synth import make_pair, transform_error  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
SAMPLE = "/root/reference/sample_data"
SEEDS = list(range(1, 9))


def sweep_cases():
    """name -> (target, source, ground truth, swapped)"""
    g = np.load(os.path.join(OUT, "room_decimated.npz"))
    p = np.load(os.path.join(OUT, "polyhedron_pair.npz"))
    cases = {"room_decimated": (g["tgt"], g["src"], g["gt"], True),
             "polyhedron": (p["tgt"], p["src"], p["gt"], False)}
    t, s, gt = make_pair(n_points=150000, n_planes=20, seed=11)
    cases["synth_150k"] = (t, s, gt, False)
    t, s, gt = make_pair(n_points=1000000, n_planes=20, seed=5)
    cases["synth_1m"] = (t, s, gt, False)
    return cases


def run_case(ref, tgt, src, gt, swapped):
    diag = float(np.linalg.norm(np.ptp(tgt[:, :3], axis=0)))
    out = []
    for seed in SEEDS:
        ref.set_seed(seed)
        if swapped:      # the file overload: source >= 1.2 x target -> swap, return the inverse (PLADE/plade.cpp:689-704)
            ok, T = ref.registration_clouds(src, tgt)
            T = np.linalg.inv(T.astype(np.float64)) if ok else T
        else:
            ok, T = ref.registration_clouds(tgt, src)
        rot, tr = transform_error(T, gt, diag)
        out.append([float(rot), float(tr), bool(ok)])
        print("   seed %d: ok=%s rot %.3f deg trans %.5f" % (seed, ok, rot, tr), flush=True)
    return out


def detection_curve(ref):
    g = np.load(os.path.join(OUT, "room_decimated.npz"))
    p = np.load(os.path.join(OUT, "polyhedron_pair.npz"))
    clouds = {"room_tgt": (g["tgt"], [2500, 1250]), "room_src": (g["src"], [10000, 5000, 2500]), "poly_tgt": (p["tgt"], [5000, 2500, 1250])}
    rows = []
    for name, (c, supports) in clouds.items():
        for M in supports:
            planes = []      # [normal, d, sizes, seeds]
            for seed in SEEDS:
                ref.set_seed(seed)
                off, _, par = ref.detect(c, M)
                sz = np.diff(off)
                for k in range(len(sz)):
                    n, d = par[k, :3], par[k, 3]
                    for q in planes:
                        dn = float(n @ q[0])
                        if abs(dn) > 0.99 and abs(d - np.sign(dn) * q[1]) < 0.02 and abs(np.log(sz[k] / np.median(q[2]))) < 0.3 and seed not in q[3]:
                            q[2].append(int(sz[k])); q[3].add(seed)
                            break
                    else:
                        planes.append([n, d, [int(sz[k])], {seed}])
            for q in planes:
                rows.append({"cloud": name, "min_support": M, "support": int(np.median(q[2])), "found": len(q[3]), "runs": len(SEEDS)})
            print("   %s @ %d: %d distinct planes" % (name, M, len(planes)), flush=True)
    return rows


def main():
    ref = Ref(quiet=True)
    doc = {"seeds": SEEDS, "cases": {}, "detection_curve": detection_curve(ref)}
    for name, (t, s, gt, swapped) in sweep_cases().items():
        print(name, flush=True)
        doc["cases"][name] = {"errors": run_case(ref, t, s, gt, swapped), "swapped": swapped}
    if "--full-room" in sys.argv:
        tgt = load_ply(os.path.join(SAMPLE, "room_target.ply"))
        src = load_ply(os.path.join(SAMPLE, "room_source.ply"))
        gt = np.loadtxt(os.path.join(SAMPLE, "room_source_groundtruth.txt"))
        print("room_full", flush=True)
        errs = run_case(ref, tgt, src, gt, True)
        doc["cases"]["room_full"] = {"errors": errs, "swapped": True}
        os.makedirs(os.path.join(OUT, "_local"), exist_ok=True)
        np.savez(os.path.join(OUT, "_local", "room_full.npz"), tgt=tgt.astype(np.float32), src=src.astype(np.float32), gt=gt.astype(np.float64))
    elif os.path.exists(os.path.join(OUT, "seed_sweep_ref.json")):
        old = json.load(open(os.path.join(OUT, "seed_sweep_ref.json")))
        if "room_full" in old.get("cases", {}):
            doc["cases"]["room_full"] = old["cases"]["room_full"]
    json.dump(doc, open(os.path.join(OUT, "seed_sweep_ref.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
