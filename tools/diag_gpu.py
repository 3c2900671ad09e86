"""Ad-hoc GPU diagnostics (not a test): prints where the CUDA path and the reference dumps differ."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import plade_b200
from plade_b200 import Planes
from plade_b200.synth import make_pair, transform_error
from oracle.ref import Ref, have_ref
np.set_printoptions(linewidth=200, precision=6, suppress=True)
G = os.path.join(ROOT, "tests", "golden")
ctx = plade_b200.Context()
ref = Ref() if have_ref() else None

def umeyama():
    rng = np.random.default_rng(9); n = 500
    v1 = rng.normal(size=(n, 3)); v2 = rng.normal(size=(n, 3))
    v1 /= np.linalg.norm(v1, axis=1, keepdims=True); v2 /= np.linalg.norm(v2, axis=1, keepdims=True)
    Rs = []
    for _ in range(n):
        a = rng.normal(size=3); a /= np.linalg.norm(a); ang = rng.uniform(0, 3)
        K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
        Rs.append(np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * K @ K)
    R = np.array(Rs)
    w1 = np.einsum("nij,nj->ni", R, v1); w2 = np.einsum("nij,nj->ni", R, v2)
    sp, tp = rng.uniform(-1, 1, size=(n, 3)), rng.uniform(-1, 1, size=(n, 3))
    inp = np.concatenate([v1, v2, w1, w2, sp, tp], axis=1).astype(np.float32)
    Rg, Tg = ctx.transforms_from_matches(inp)
    Rr, Tr = ref.transform_from_two_vecs(inp)
    d = np.abs(Rg - Rr).reshape(n, -1).max(1)
    print("umeyama: max|dR|", d.max(), "worst", np.argsort(-d)[:5], d[np.argsort(-d)[:5]], "max|dT|", np.abs(Tg - Tr).max())
    k = int(np.argmax(d)); print(" ours\n", Rg[k], "\n ref\n", Rr[k], "\n true\n", R[k], "\n angle v1v2", np.degrees(np.arccos(v1[k] @ v2[k])))

def pipeline(which):
    g = dict(np.load(os.path.join(G, "polyhedron_stages.npz" if which == "poly" else "synth_small_stages.npz")))
    if which == "poly":
        p = np.load(os.path.join(G, "polyhedron_pair.npz")); tgt, src = p["tgt"], p["src"]
    else:
        tgt, src, _ = make_pair(n_points=200000, n_planes=20, seed=11)
    ctx.set_debug(True)
    ok, T = ctx.register_with_planes(tgt, src, Planes(g["t_off"], g["t_idx"], g["t_par"]), Planes(g["s_off"], g["s_idx"], g["s_par"]))
    ctx.set_debug(False)
    print(which, "ok", ok, "err vs ref T", transform_error(T, g["T"]), "stage ms", {k: round(v * 1e3, 2) for k, v in ctx.stage_times().items()})
    for name, dt in (("tgt_db_desc", np.float32), ("src_lines", np.float32), ("tgt_lines", np.float32), ("src_center", np.float32), ("tgt_center", np.float32),
                     ("src_plane_center", np.float32), ("tgt_plane_center", np.float32), ("src_plane_radius", np.float32), ("tgt_plane_radius", np.float32),
                     ("src_plane_corners4", np.float32)):
        a, b = ctx.blob(name, dt), g[name]
        if a.shape != b.shape: print("  ", name, "SHAPE", a.shape, b.shape); continue
        d = np.abs(a - b); print("  %-20s n=%7d max|d|=%.3g  n(>5e-5)=%d" % (name, len(a), d.max() if len(d) else 0, int((d > 5e-5).sum())))
        if name == "tgt_db_desc" and (d > 5e-5).any():
            bad = np.where(d.reshape(-1, 8).max(1) > 5e-5)[0]; print("     bad rows", bad[:10], "\n", a.reshape(-1, 8)[bad[:3]], "\n", b.reshape(-1, 8)[bad[:3]], "pairs", g["tgt_db_pair"].reshape(-1, 2)[bad[:3]])
    for name in ("src_plane_ds_offsets", "tgt_plane_ds_offsets", "lines_to_match", "tgt_db_pair", "src_line_planes"):
        a, b = ctx.blob(name, np.int32), g[name]; print("  %-20s equal=%s %s %s" % (name, a.shape == b.shape and np.array_equal(a, b), a.shape, b.shape))
    R, Tt, npl = ctx.blob("mr_R", np.float32).reshape(-1, 9), ctx.blob("mr_T", np.float32).reshape(-1, 3), ctx.blob("mr_nplanes", np.int32)
    Rr, Tr, nr = g["mr_R"].reshape(-1, 9), g["mr_T"].reshape(-1, 3), g["mr_nplanes"]
    print("  hypotheses ours", len(R), "ref", len(Rr), "nplanes ours", npl[:12], "ref", nr[:12])
    used = set()
    for i in range(len(Rr)):
        dd = np.abs(R - Rr[i]).max(1) + np.abs(Tt - Tr[i]).max(1) if len(R) else np.array([])
        j = int(np.argmin(dd)) if len(dd) else -1
        if j < 0 or dd[j] > 1e-3: print("   ref hyp", i, "nplanes", nr[i], "has no counterpart (best", dd[j] if j >= 0 else None, ")")
        else: used.add(j)
    print("   ours without counterpart:", [j for j in range(len(R)) if j not in used][:10])
    sc, scr = ctx.blob("ver_score", np.float32), g["ver_score"]
    print("  best ours", int(np.argmax(sc)), float(sc.max()), "ref", int(np.argmax(scr)), float(scr.max()))

def ransac_poly():
    p = np.load(os.path.join(G, "polyhedron_pair.npz")); g = dict(np.load(os.path.join(G, "polyhedron_stages.npz")))
    for name in ("tgt", "src"):
        pl = ctx.extract_planes(p[name], 10000)
        print("ransac", name, "planes", len(pl), "sizes", pl.sizes().tolist())
    print("ref sizes tgt", np.diff(g["t_off"]).tolist(), "src", np.diff(g["s_off"]).tolist())
    ctx.set_debug(True); ok, T = ctx.register_clouds(p["tgt"], p["src"]); ctx.set_debug(False)
    print("e2e poly ok", ok, transform_error(T, p["gt"]), "stage ms", {k: round(v * 1e3, 2) for k, v in ctx.stage_times().items()})
    sc = ctx.blob("ver_score", np.float32); cn = ctx.blob("ver_count", np.int32); npl = ctx.blob("mr_nplanes", np.int32)
    o = np.argsort(-sc)[:6]; print(" top scores", sc[o], cn[o], npl[o], "n hyp", len(sc))
    R, Tt = ctx.blob("mr_R", np.float32).reshape(-1, 3, 3), ctx.blob("mr_T", np.float32).reshape(-1, 3)
    errs = []
    for i in range(len(R)):
        M = np.eye(4); M[:3, :3] = R[i]; M[:3, 3] = Tt[i]; errs.append(transform_error(M, p["gt"])[0])
    errs = np.array(errs); print(" hyps within 1deg of gt:", int((errs < 1).sum()), "min err", errs.min() if len(errs) else None)

if __name__ == "__main__":
    for f in sys.argv[1:]:
        if f == "umeyama": umeyama()
        elif f == "ransac": ransac_poly()
        else: pipeline(f)
