"""Generates the committed golden fixtures from the reference itself (run in the build container, where
/root/reference exists and oracle/_ref has been built):

    python tests/golden/make_golden.py

  polyhedron_pair.npz   the reference's sample_data polyhedron pair (config 1, the correctness gate),
                        its ground truth and the authors' published result (file_pairs_results.txt:3-7)
  polyhedron_stages.npz planes extracted by the reference's RANSAC (fixed seed) + every stage blob of
                        the reference's registration(T, tgt, src, planes, planes) on those planes
  synth_small_*.npz     same for a small synthetic planar scene (tests/synth.py)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.ref import Ref, load_ply  # noqa: E402

SAMPLE = "/root/reference/sample_data"
OUT = os.path.dirname(os.path.abspath(__file__))

BLOBS = {
    "src_ds": np.float32, "tgt_ds": np.float32, "src_center": np.float32, "tgt_center": np.float32,
    "src_lines": np.float32, "tgt_lines": np.float32, "src_line_planes": np.int32, "tgt_line_planes": np.int32,
    "src_planes": np.float32, "tgt_planes": np.float32,
    "src_plane_ds_offsets": np.int32, "tgt_plane_ds_offsets": np.int32,
    "src_plane_corners4": np.float32, "tgt_plane_corners4": np.float32,
    "src_plane_center": np.float32, "tgt_plane_center": np.float32,
    "src_plane_radius": np.float32, "tgt_plane_radius": np.float32,
    "lines_to_match": np.int32, "tgt_db_desc": np.float32, "tgt_db_pair": np.int32,
    "match_params": np.float64, "mr_R": np.float32, "mr_T": np.float32, "mr_nplanes": np.int32,
    "average_space": np.float32, "downsample_distance": np.float32, "src_radius": np.float64,
    "ver_overlap": np.float32, "ver_score": np.float32, "ver_center": np.float32, "final_T": np.float32,
}


def parse_result_file(path, which):
    rows = [l.split() for l in open(path).read().splitlines()]
    mats, cur = [], []
    for r in rows:
        if len(r) == 4:
            try:
                cur.append([float(x) for x in r])
            except ValueError:
                cur = []
            if len(cur) == 4:
                mats.append(np.array(cur))
                cur = []
        else:
            cur = []
    return mats[which]


def stage_dump(ref, tgt, src, seed, init_support=10000):
    ref.set_seed(seed)
    tp = ref.extract(tgt, init_support, "t_")
    sp = ref.extract(src, init_support, "s_")
    ok, T = ref.registration_planes(tgt, src, tp, sp, dump=True)
    d = {k: ref.blob(k, dt) for k, dt in BLOBS.items()}
    d.update(dict(t_off=tp[0], t_idx=tp[1], t_par=tp[2], s_off=sp[0], s_idx=sp[1], s_par=sp[2], ok=np.array([ok]), T=T, seed=np.array([seed])))
    return d


def main():
    ref = Ref()
    tgt = load_ply(os.path.join(SAMPLE, "polyhedron_target.ply"))
    src = load_ply(os.path.join(SAMPLE, "polyhedron_source.ply"))
    gt = np.loadtxt(os.path.join(SAMPLE, "polyhedron_source_groundtruth.txt"))
    published = parse_result_file(os.path.join(SAMPLE, "file_pairs_results.txt"), 0)
    np.savez_compressed(os.path.join(OUT, "polyhedron_pair.npz"), tgt=tgt, src=src, gt=gt, published=published)
    d = stage_dump(ref, tgt, src, seed=3)
    print("polyhedron: ok", d["ok"], "planes", len(d["t_par"]), len(d["s_par"]), "hyps", len(d["mr_nplanes"]))
    print(d["T"])
    np.savez_compressed(os.path.join(OUT, "polyhedron_stages.npz"), **d)

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import synth
    tgt, src, gt = synth.make_pair(n_points=200000, n_planes=20, seed=11)
    d = stage_dump(ref, tgt, src, seed=1)
    print("synth: ok", d["ok"], "planes", len(d["t_par"]), len(d["s_par"]), "hyps", len(d["mr_nplanes"]))
    print(d["T"], "\n", gt)
    np.savez_compressed(os.path.join(OUT, "synth_small_stages.npz"), gt=gt, **d)


if __name__ == "__main__":
    main()
