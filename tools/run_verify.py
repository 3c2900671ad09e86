"""K5 at the bench shape for ncu captures: ds clouds of the 2M pair, N perturbed hypotheses (never a bench value)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import plade_b200
from plade_b200.synth import make_pair, perturbed_hypotheses

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
H = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
planes = int(sys.argv[4]) if len(sys.argv) > 4 else 20
tgt, src, gt = make_pair(n_points=n, n_planes=planes, seed=20240611)
ctx = plade_b200.Context()
leaf = 4 * ctx.average_spacing(src)
ds_t, ds_s = ctx.voxel_downsample(tgt[:, :3], leaf), ctx.voxel_downsample(src[:, :3], leaf)
R, T, true_idx = perturbed_hypotheses(gt, H, seed=7)
rc, c, whd, _ = ctx.bounding_box(ds_s)
cen = (np.einsum("hij,j->hi", R, c) + T).astype(np.float32)
ctx.verify_upload(ds_s, ds_t, leaf)
for _ in range(reps):
    counts, ms = ctx.verify_resident(R, T, cen, float(max(whd) / 2), leaf)
    print("H=%d ns=%d nt=%d kernel_ms=%.3f best=%d true=%d GB/s(alg)=%.1f" % (H, len(ds_s), len(ds_t), ms, int(np.argmax(counts)), true_idx,
                                                                            (H * 16.0 * len(ds_s) + 16.0 * len(ds_t)) / ms / 1e6))
