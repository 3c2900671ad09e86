"""Runs registration-with-the-reference's-planes on the two golden cases with debug dumps on and writes every
stage blob to gpurun_out/ours_<case>.npz, for offline comparison with tests/golden/*_stages.npz."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from plade_b200 import Context, Planes
from plade_b200.synth import make_pair

F, I, D = np.float32, np.int32, np.float64
NAMES = {"average_space": F, "downsample_distance": F, "tgt_db_desc": F, "src_q_desc": F, "init_R": F, "init_T": F,
         "match_in18": F, "cluster_params": F, "mr_R": F, "mr_T": F, "ver_score": F, "ver_center": F,
         "tgt_db_pair": I, "lines_to_match": I, "match_offsets": I, "match_idx": I, "cluster_label": I, "mr_nplanes": I,
         "ver_count": I, "cand_rt": I, "cand_nplanes": I, "cand_pen": I, "match_dist2": D}
for side in ("tgt_", "src_"):
    for k in ("lines", "ds", "center", "plane_ds", "plane_corners4", "plane_center", "plane_radius"):
        NAMES[side + k] = F
    for k in ("line_planes", "plane_ds_offsets"):
        NAMES[side + k] = I
    NAMES[side + "radius"] = D

ctx = Context()
gold = os.path.join(ROOT, "tests", "golden")
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
for which in ("poly", "synth"):
    if which == "poly":
        g = np.load(os.path.join(gold, "polyhedron_stages.npz"))
        pp = np.load(os.path.join(gold, "polyhedron_pair.npz"))
        tgt, src = pp["tgt"], pp["src"]
    else:
        g = np.load(os.path.join(gold, "synth_small_stages.npz"))
        tgt, src, _ = make_pair(n_points=200000, n_planes=20, seed=11)
    ctx.set_debug(True)
    ok, T = ctx.register_with_planes(tgt, src, Planes(g["t_off"], g["t_idx"], g["t_par"]), Planes(g["s_off"], g["s_idx"], g["s_par"]))
    out = {k: ctx.blob(k, dt) for k, dt in NAMES.items()}
    out["T"] = T
    np.savez_compressed(os.path.join(ROOT, "gpurun_out", "ours_%s.npz" % which), **out)
    print(which, ok, {k: v.shape for k, v in out.items() if v.size == 0})
