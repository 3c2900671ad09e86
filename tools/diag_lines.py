"""Diagnostic: how far are our intersection lines / descriptors from the reference's golden dumps?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from plade_b200 import Context, Planes
from plade_b200.synth import make_pair

def planes(g, p):
    return Planes(g[p + "_off"], g[p + "_idx"], g[p + "_par"])

ctx = Context()
root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
for which in ("poly", "synth"):
    if which == "poly":
        g = np.load(os.path.join(root, "polyhedron_stages.npz"))
        pp = np.load(os.path.join(root, "polyhedron_pair.npz"))
        tgt, src = pp["tgt"], pp["src"]
    else:
        g = np.load(os.path.join(root, "synth_small_stages.npz"))
        tgt, src, _ = make_pair(n_points=200000, n_planes=20, seed=11)
    ctx.set_debug(True)
    ok, T = ctx.register_with_planes(tgt, src, planes(g, "t"), planes(g, "s"))
    for side in ("tgt", "src"):
        a, b = ctx.blob(side + "_lines", np.float32).reshape(-1, 6), g[side + "_lines"].reshape(-1, 6)
        d = np.abs(a - b)
        print(which, side, "lines", a.shape, "max|dvec|", d[:, :3].max(), "max|dpt|", d[:, 3:].max(), "exact rows", int((d.max(1) == 0).sum()))
        for k in ("_center", "_plane_center", "_plane_radius", "_plane_corners4"):
            x, y = ctx.blob(side + k, np.float32), g[side + k]
            print("   ", k, "max diff", np.abs(x - y.ravel()).max() if x.size == y.size else ("size", x.size, y.size))
    a, b = ctx.blob("tgt_db_desc", np.float32).reshape(-1, 8), g["tgt_db_desc"].reshape(-1, 8)
    print(which, "desc0 max diff", np.abs(a[:, 0] - b[:, 0]).max(), "exact", int((a[:, 0] == b[:, 0]).sum()), "/", len(a), " desc1-7 max", np.abs(a[:, 1:] - b[:, 1:]).max())
    R, Tt = ctx.blob("mr_R", np.float32).reshape(-1, 9), ctx.blob("mr_T", np.float32).reshape(-1, 3)
    Rr, Tr = g["mr_R"].reshape(-1, 9), g["mr_T"].reshape(-1, 3)
    print(which, "hyps ours", len(R), "ref", len(Rr))
    if len(R) == len(Rr):
        print("   max|dR|", np.abs(R - Rr).max(), "max|dT|", np.abs(Tt - Tr).max())
    sc, scr = ctx.blob("ver_score", np.float32), g["ver_score"]
    print("   best score ours", sc.max(), "ref", scr.max())
    npl, nplr = ctx.blob("mr_nplanes", np.int32), g["mr_nplanes"]
    def nearest(Ra, Ta, Rb, Tb):
        out = []
        for i in range(len(Ra)):
            d = np.abs(Rb - Ra[i]).max(1) + np.abs(Tb - Ta[i]).max(1)
            out.append((int(np.argmin(d)), float(d.min())))
        return out
    for i, (j, d) in enumerate(nearest(Rr, Tr, R, Tt)):
        if d > 1e-4:
            print("   ref hyp", i, "nplanes", nplr[i], "score", scr[i], "nearest ours", j, "dist", d, "ours nplanes", npl[j])
    for i, (j, d) in enumerate(nearest(R, Tt, Rr, Tr)):
        if d > 1e-4:
            print("   our hyp", i, "nplanes", npl[i], "score", sc[i], "nearest ref", j, "dist", d)
    print("   blobs:", [k for k in g.files if k.startswith(("match", "mr_", "ver_", "lines_to"))])
    mp, mpr = ctx.blob("match_params", np.float32), g["match_params"]
    print("   match_params ours", mp, "ref", mpr)
    # stage by stage against the compiled reference (when it travelled with the repo)
    from oracle.ref import Ref, have_ref
    if have_ref():
        ref = Ref()
        dbd, qd = ctx.blob("tgt_db_desc", np.float32).reshape(-1, 8), ctx.blob("src_q_desc", np.float32).reshape(-1, 8)
        off, idx, d2 = ctx.blob("match_offsets", np.int32), ctx.blob("match_idx", np.int32), ctx.blob("match_dist2", np.float64)
        roff, ridx, rd = ref.match_descriptors(dbd, qd, 0.04)
        print("   matches ours", len(idx), "ref", len(ridx), "offsets equal", np.array_equal(off, roff), "idx equal", np.array_equal(idx, ridx))
        in18 = ctx.blob("match_in18", np.float32).reshape(-1, 18)
        iR, iT = ctx.blob("init_R", np.float32).reshape(-1, 9), ctx.blob("init_T", np.float32).reshape(-1, 3)
        rR, rT = ref.transform_from_two_vecs(in18)
        rR = rR.reshape(-1, 9)
        print("   transforms max|dR|", np.abs(iR - rR).max(), "max|dT|", np.abs(iT - rT).max())
        cp = ctx.blob("cluster_params", np.float32)
        lab = ctx.blob("cluster_label", np.int32)
        nc, rlab = ref.cluster_transformations(rR, rT, cp[0], cp[1])
        nc2, rlab2 = ref.cluster_transformations(iR, iT, cp[0], cp[1])
        print("   clusters ours", len(np.unique(lab)), "ref on ref transforms", nc, "ref on our transforms", nc2)
