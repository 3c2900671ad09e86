"""GPU parity tests: every CUDA stage, called through the C ABI, against the oracle on the same inputs.

Bars: bit-exact for integer / index work (inlier counts, match sets, voxel membership, plane-consensus
masks, cluster labels); floating-point stages carry their tolerance in the test.
"""
import os

import numpy as np
import pytest

from tests.conftest import synth_small_tgt

from plade_b200 import Planes
from plade_b200.synth import make_pair, perturbed_hypotheses, transform_error

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _rand_rigid(rng, n, rot_deg=20.0, trans=0.3):
    Rs, Ts = [], []
    for _ in range(n):
        a = rng.normal(size=3)
        a /= np.linalg.norm(a)
        ang = np.deg2rad(rng.uniform(0, rot_deg))
        K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
        Rs.append(np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * K @ K)
        Ts.append(rng.uniform(-trans, trans, size=3))
    return np.array(Rs, np.float32), np.array(Ts, np.float32)


# ---------------------------------------------------------------------------------------------- K5
def test_verify_counts_bit_exact_random(ctx, restate):
    rng = np.random.default_rng(0)
    tgt = rng.uniform(0, 1, size=(6000, 3)).astype(np.float32)
    src = (tgt[rng.permutation(6000)[:4000]] + rng.normal(0, 0.004, size=(4000, 3))).astype(np.float32)
    R, T = _rand_rigid(rng, 37, rot_deg=6, trans=0.05)
    R[0], T[0] = np.eye(3), 0
    centers = (np.einsum("hij,j->hi", R, src.mean(0)) + T).astype(np.float32)
    for ball, inl in ((0.6, 0.01), (0.25, 0.02), (5.0, 0.003)):
        got = ctx.verify_hypotheses(src, tgt, R, T, centers, ball, inl)
        want = restate.verify_counts(src, tgt, R, T, centers, ball, inl)
        assert np.array_equal(got, want)
    assert got[0] > 0


def test_verify_edge_cases(ctx, restate):
    rng = np.random.default_rng(1)
    tgt = rng.uniform(0, 1, size=(500, 3)).astype(np.float32)
    src = rng.uniform(0, 1, size=(300, 3)).astype(np.float32)
    R, T = _rand_rigid(rng, 3)
    c = np.zeros((3, 3), np.float32)
    # no hypotheses, empty clouds, ball that contains nothing, single points, duplicates
    assert len(ctx.verify_hypotheses(src, tgt, R[:0], T[:0], c[:0], 1.0, 0.05)) == 0
    assert np.array_equal(ctx.verify_hypotheses(src[:0], tgt, R, T, c, 1.0, 0.05), np.zeros(3, np.uint32))
    assert np.array_equal(ctx.verify_hypotheses(src, tgt[:0], R, T, c, 1.0, 0.05), np.zeros(3, np.uint32))
    far = np.full((3, 3), 100.0, np.float32)
    assert np.array_equal(ctx.verify_hypotheses(src, tgt, R, T, far, 0.5, 0.05), np.zeros(3, np.uint32))
    one = np.array([[0.5, 0.5, 0.5]], np.float32)
    I = np.eye(3, dtype=np.float32)[None]
    z = np.zeros((1, 3), np.float32)
    assert ctx.verify_hypotheses(one, one, I, z, one, 1.0, 0.01)[0] == 1
    dup = np.repeat(one, 64, axis=0)
    got = ctx.verify_hypotheses(dup, dup, I, z, one, 1.0, 0.01)
    assert got[0] == 64 and np.array_equal(got, restate.verify_counts(dup, dup, I, z, one, 1.0, 0.01))
    # ragged tile sizes around the 2048-point tile and the 32-hypothesis chunk
    for ns, H in ((2047, 31), (2048, 32), (2049, 33), (4097, 65)):
        s = rng.uniform(0, 1, size=(ns, 3)).astype(np.float32)
        Rr, Tr = _rand_rigid(rng, H, rot_deg=3, trans=0.02)
        cc = (np.einsum("hij,j->hi", Rr, s.mean(0)) + Tr).astype(np.float32)
        assert np.array_equal(ctx.verify_hypotheses(s, tgt, Rr, Tr, cc, 0.7, 0.06), restate.verify_counts(s, tgt, Rr, Tr, cc, 0.7, 0.06))


def test_verify_golden_polyhedron(ctx, restate, poly_stages):
    """The reference's own ComputeOverlap (FLANN kd-trees) on its own hypotheses -> overlap ratios."""
    g = poly_stages
    src, tgt = g["src_ds"].reshape(-1, 3), g["tgt_ds"].reshape(-1, 3)
    R, T, c = g["mr_R"].reshape(-1, 9), g["mr_T"].reshape(-1, 3), g["ver_center"].reshape(-1, 3)
    ball, inl = float(g["src_radius"][0]), float(g["downsample_distance"][0])
    got = ctx.verify_hypotheses(src, tgt, R, T, c, ball, inl)
    assert np.array_equal(got, restate.verify_counts(src, tgt, R, T, c, ball, inl))
    overlap = (got.astype(np.float64) / min(len(src), len(tgt))).astype(np.float32)
    assert np.array_equal(overlap, g["ver_overlap"])          # bit-exact against the reference's floats
    score = (0.2 * (g["mr_nplanes"] / float(len(g["s_par"]))) + 0.8 * overlap.astype(np.float64)).astype(np.float32)
    assert np.array_equal(score, g["ver_score"])


def test_verify_golden_synth(ctx, synth_stages):
    g = synth_stages
    src, tgt = g["src_ds"].reshape(-1, 3), g["tgt_ds"].reshape(-1, 3)
    got = ctx.verify_hypotheses(src, tgt, g["mr_R"].reshape(-1, 9), g["mr_T"].reshape(-1, 3), g["ver_center"].reshape(-1, 3),
                                float(g["src_radius"][0]), float(g["downsample_distance"][0]))
    overlap = (got.astype(np.float64) / min(len(src), len(tgt))).astype(np.float32)
    assert np.array_equal(overlap, g["ver_overlap"])


def test_verify_full_size_properties(ctx):
    """BASELINE-size property checks (no oracle at this size): identity on itself counts every point,
    a far translation counts none, the count is invariant to hypothesis order and chunking."""
    rng = np.random.default_rng(3)
    tgt, src, gt = make_pair(n_points=400000, n_planes=20, seed=5)
    leaf = 0.012
    ds_t = ctx.voxel_downsample(tgt[:, :3], leaf)
    ds_s = ctx.voxel_downsample(src[:, :3], leaf)
    I = np.eye(3, dtype=np.float32)[None]
    z = np.zeros((1, 3), np.float32)
    c = ds_t.mean(0, keepdims=True).astype(np.float32)
    assert ctx.verify_hypotheses(ds_t, ds_t, I, z, c, 10.0, leaf)[0] == len(ds_t)
    assert ctx.verify_hypotheses(ds_t, ds_t, I, z + 50.0, c + 50.0, 10.0, leaf)[0] == 0
    R, T, true_idx = perturbed_hypotheses(gt, 257, seed=7)
    cen = (np.einsum("hij,j->hi", R, ds_s.mean(0)) + T).astype(np.float32)
    a = ctx.verify_hypotheses(ds_s, ds_t, R, T, cen, 2.0, leaf)
    perm = rng.permutation(257)
    b = ctx.verify_hypotheses(ds_s, ds_t, R[perm], T[perm], cen[perm], 2.0, leaf)
    assert np.array_equal(a[perm], b)
    assert int(np.argmax(a)) == true_idx
    one = ctx.verify_hypotheses(ds_s, ds_t, R[true_idx:true_idx + 1], T[true_idx:true_idx + 1], cen[true_idx:true_idx + 1], 2.0, leaf)
    assert one[0] == a[true_idx]


# ---------------------------------------------------------------------------------------------- K2
def test_voxel_downsample_bit_exact(ctx, restate):
    rng = np.random.default_rng(4)
    for n, leaf in ((1, 0.1), (17, 0.5), (5000, 0.05), (60000, 0.013)):
        pts = rng.uniform(-1, 1, size=(n, 3)).astype(np.float32)
        got, want = ctx.voxel_downsample(pts, leaf), restate.voxel_downsample(pts, leaf)
        assert got.shape == want.shape and np.array_equal(got, want)
    pn = rng.uniform(-1, 1, size=(3000, 6)).astype(np.float32)        # stride-6 input (PointNormal)
    assert np.array_equal(ctx.voxel_downsample(pn, 0.07), restate.voxel_downsample(pn, 0.07))
    with pytest.raises(RuntimeError):
        ctx.voxel_downsample(pn[:0], 0.07)                             # DownSamplePointCloud returns -1
    with pytest.raises(RuntimeError):
        ctx.voxel_downsample(pn, 0.0)


def test_voxel_golden_reference_clouds(ctx, restate, poly_pair, poly_stages):
    """Same voxels as the reference's pcl::VoxelGrid; centroids equal up to the order of the float
    additions inside a voxel (the reference's std::sort is unstable): <= 2 ulp of the coordinate."""
    g = poly_stages
    leaf = float(g["downsample_distance"][0])
    for name, cloud in (("tgt_ds", poly_pair["tgt"]), ("src_ds", poly_pair["src"])):
        got = ctx.voxel_downsample(cloud, leaf)
        want = g[name].reshape(-1, 3)
        assert got.shape == want.shape
        assert np.array_equal(np.floor(got / leaf), np.floor(want / leaf))
        assert np.max(np.abs(got - want)) <= 4 * np.finfo(np.float32).eps * 2.0
        assert np.array_equal(got, restate.voxel_downsample(cloud, leaf))


def test_average_spacing(ctx, restate, poly_pair, poly_stages):
    got = ctx.average_spacing(poly_pair["src"])
    assert got == float(poly_stages["average_space"][0])               # bit-exact vs the reference (FLANN kNN)
    rng = np.random.default_rng(5)
    for n in (5, 6, 100, 25000):
        pts = rng.uniform(0, 1, size=(n, 6)).astype(np.float32)
        assert ctx.average_spacing(pts) == restate.average_spacing(pts)


def test_bounding_box_vs_reference(ctx, poly_stages, ref):
    pts = poly_stages["src_ds"].reshape(-1, 3)
    rc, c, whd, corners = ctx.bounding_box(pts)
    rc2, c2, whd2, corners2 = ref.bounding_box(pts)
    assert rc == rc2 == 0
    assert np.allclose(c, c2, atol=2e-5) and np.allclose(np.sort(whd), np.sort(whd2), atol=2e-5)


# ---------------------------------------------------------------------------------------------- K1
def test_score_planes_bit_exact(ctx, restate):
    tgt, _, _ = make_pair(n_points=50000, n_planes=12, seed=3)
    rng = np.random.default_rng(6)
    planes = []
    for _ in range(24):
        i = rng.integers(0, len(tgt), size=3)
        p = tgt[i, :3].astype(np.float64)
        n = np.cross(p[1] - p[0], p[2] - p[1])
        if np.linalg.norm(n) < 1e-9:
            continue
        n /= np.linalg.norm(n)
        planes.append([*n, float(n @ p[0])])
    planes.append([0, 0, 1, float(tgt[:, 2].min())])
    planes = np.array(planes, np.float32)
    assigned = np.where(rng.random(len(tgt)) < 0.3, 0, -1).astype(np.int32)
    for eps in (0.004, 0.02):
        got, gmask = ctx.score_planes(tgt, planes, eps, 0.8, assigned, want_mask=True)
        want, wmask = restate.score_planes(tgt, planes, eps, 0.8, assigned, want_mask=True)
        assert np.array_equal(got, want) and np.array_equal(gmask, wmask)
    assert np.array_equal(ctx.score_planes(tgt, planes, 0.01, 0.8), restate.score_planes(tgt, planes, 0.01, 0.8))


def _match_planes(ours, theirs_par, theirs_sizes, eps):
    """For each reference plane (largest first) find our plane with the same normal (up to sign) and offset."""
    res = []
    for k in np.argsort(-theirs_sizes):
        n, d = theirs_par[k, :3], theirs_par[k, 3]
        best = None
        for j in range(len(ours)):
            m, e = ours.params[j, :3], ours.params[j, 3]
            s = np.sign(n @ m)
            ang = np.degrees(np.arccos(np.clip(abs(n @ m), 0, 1)))
            if ang < 1.0 and abs(d - s * e) < eps:
                best = (j, ang, abs(d - s * e))
                break
        res.append((k, best))
    return res


def test_ransac_planes_vs_reference_golden(ctx, poly_pair, poly_stages):
    """Statistical parity: the big planes the reference finds are found with the same
    parameters (normal < 0.5 deg, |d| < eps) and support within 5 %."""
    g = poly_stages
    cloud = poly_pair["tgt"]
    scale = max(np.ptp(cloud[:, 0]), np.ptp(cloud[:, 1]))
    eps = 0.005 * scale
    ours = ctx.detect_planes(cloud, 1000)
    sizes = np.diff(g["t_off"])
    big = [k for k in range(len(sizes)) if sizes[k] >= 2000]
    matched = _match_planes(ours, g["t_par"], sizes, eps)
    ours_sizes = ours.sizes()
    for k, best in matched:
        if k not in big:
            continue
        assert best is not None, "reference plane %d (support %d) not found" % (k, sizes[k])
        j, ang, dd = best
        assert ang < 0.5 and dd < eps
        assert abs(int(ours_sizes[j]) - int(sizes[k])) <= 0.05 * sizes[k] + 50
    # every returned plane respects min_support, indices are unique and in range
    assert ours_sizes.min() >= 1000
    assert len(np.unique(ours.indices)) == len(ours.indices) and ours.indices.max() < len(cloud)


def test_ransac_synthetic_ground_truth(ctx):
    tgt, src, gt, It, Is = make_pair(n_points=300000, n_planes=20, seed=9, return_ids=True)
    planes = ctx.extract_planes(tgt, 10000)
    assert 10 <= len(planes) <= 40
    # every member lies within the global-score band 3*eps of its plane (n.x + d = 0), and each plane is
    # dominated by one generated rectangle (two near-coplanar rectangles may legitimately merge)
    scale = max(np.ptp(tgt[:, 0]), np.ptp(tgt[:, 1]))
    for k in range(len(planes)):
        members = planes.indices[planes.offsets[k]:planes.offsets[k + 1]]
        dist = np.abs(tgt[members, :3].astype(np.float64) @ planes.params[k, :3] + planes.params[k, 3])
        assert dist.max() <= 3 * 0.005 * scale * 1.001
        assert np.bincount(It[members]).max() >= 0.6 * len(members)
    # deterministic for a fixed seed
    again = ctx.extract_planes(tgt, 10000)
    assert np.array_equal(planes.offsets, again.offsets) and np.array_equal(planes.params, again.params)


# ---------------------------------------------------------------------------------------------- K3c
def test_match_descriptors_bit_exact(ctx, restate, poly_stages):
    g = poly_stages
    db = g["tgt_db_desc"].reshape(-1, 8)
    rng = np.random.default_rng(7)
    q = db[rng.permutation(len(db))[:700]] + rng.normal(0, 0.012, size=(700, 8)).astype(np.float32)
    q = np.concatenate([q, db[:50]])                                     # exact duplicates: distance 0
    off, idx, d2 = ctx.match_descriptors(db, q, 0.04)
    woff, widx, wd2 = restate.match_descriptors(db, q, 0.04)
    assert np.array_equal(off, woff) and np.array_equal(idx, widx) and np.array_equal(d2, wd2)
    assert len(idx) > 700
    # empty / no-match cases
    off, idx, d2 = ctx.match_descriptors(db, q[:0], 0.04)
    assert len(idx) == 0 and np.array_equal(off, [0])
    off, idx, d2 = ctx.match_descriptors(db, q + 10.0, 0.04)
    assert len(idx) == 0 and np.all(off == 0)


def test_match_descriptors_vs_ann(ctx, ref, poly_stages):
    """Against the reference's own ANN kd-tree: same neighbour sets, same ascending distances."""
    db = poly_stages["tgt_db_desc"].reshape(-1, 8)
    rng = np.random.default_rng(8)
    q = db[rng.permutation(len(db))[:300]] + rng.normal(0, 0.01, size=(300, 8)).astype(np.float32)
    off, idx, d2 = ctx.match_descriptors(db, q, 0.04)
    roff, ridx, rd = ref.match_descriptors(db, q, 0.04)
    assert np.array_equal(off, roff)
    for a in range(len(q)):
        s = slice(off[a], off[a + 1])
        assert set(idx[s]) == set(ridx[s])
        assert np.array_equal(d2[s].astype(np.float32), rd[s])           # ANN returns float(double dist)


# ---------------------------------------------------------------------------------------------- K4
def test_nearest_points_two_lines_bit_exact_vs_cv_solve(ctx, ref, poly_stages):
    """K3a: closest points of line pairs.  The reference solves a 9x9 system with cv::solve(DECOMP_SVD) in
    float (PLADE/util.cpp:1167-1229); the GPU kernel restates OpenCV's Jacobi SVD + back-substitution
    operation for operation, so points and length must be IDENTICAL to the compiled reference's."""
    rng = np.random.default_rng(5)
    n = 300
    v1, v2 = rng.normal(size=(n, 3)), rng.normal(size=(n, 3))
    v2[:20] = v1[:20] + 1e-3 * rng.normal(size=(20, 3))          # nearly parallel: ill-conditioned solve
    p1, p2 = rng.uniform(-5, 5, size=(n, 3)), rng.uniform(-5, 5, size=(n, 3))
    p2[20:40] = p1[20:40] + v1[20:40] * 0.7 - v2[20:40] * 1.3       # intersecting lines
    rows = [np.concatenate([v1[i], p1[i], v2[i], p2[i]]) for i in range(n)]
    L = poly_stages["tgt_lines"].reshape(-1, 6)                     # the reference's own intersection lines
    for i in range(len(L)):
        for j in range(i + 1, len(L)):
            rows.append(np.concatenate([L[i], L[j]]))
    rows.append(np.concatenate([[0, 0, 2], [1, 2, 3], [0, 0, 5], [4, 5, 6]]))   # identical directions -> -1
    lines12 = np.asarray(rows, dtype=np.float32)
    pts, length = ctx.nearest_points_two_lines(lines12)
    worst = 0.0
    for i, r in enumerate(lines12):
        rc, q1, q2, ln = ref.nearest_points_two_lines(r[0:3], r[3:6], r[6:9], r[9:12])
        if rc != 0:
            assert length[i] == -1
            continue
        worst = max(worst, float(np.abs(pts[i, 0] - q1).max()), float(np.abs(pts[i, 1] - q2).max()))
        assert np.array_equal(pts[i, 0], q1) and np.array_equal(pts[i, 1], q2), (i, pts[i], q1, q2)
        assert length[i] == ln
    assert length[-1] == -1 and worst == 0.0


def test_transforms_from_matches_vs_eigen_umeyama(ctx, ref):
    """K4a runs a float restatement of Eigen's umeyama + JacobiSVD<Matrix3f> (umeyama.h): R and T must be
    IDENTICAL to the compiled reference's ComputeTransformationUsingTwoVecAndOnePoint (PLADE/util.cpp:604-624),
    including near-parallel direction pairs and noisy (non-rigid) correspondences."""
    rng = np.random.default_rng(9)
    n = 4000
    v1 = rng.normal(size=(n, 3)); v2 = rng.normal(size=(n, 3))
    v1 /= np.linalg.norm(v1, axis=1, keepdims=True); v2 /= np.linalg.norm(v2, axis=1, keepdims=True)
    v2[:200] = v1[:200] + 0.05 * rng.normal(size=(200, 3))                 # nearly parallel pairs
    v1 *= rng.uniform(0.3, 1, size=(n, 1)); v2 *= rng.uniform(0.3, 1, size=(n, 1))
    R, _ = _rand_rigid(rng, n, rot_deg=179)
    w1 = np.einsum("nij,nj->ni", R, v1) + rng.normal(0, 1e-3, size=(n, 3))
    w2 = np.einsum("nij,nj->ni", R, v2) + rng.normal(0, 1e-3, size=(n, 3))
    w1[200:400] = rng.normal(size=(200, 3)); w2[200:400] = rng.normal(size=(200, 3))   # no rigid relation at all
    sp, tp = rng.uniform(-1, 1, size=(n, 3)), rng.uniform(-1, 1, size=(n, 3))
    inp = np.concatenate([v1, v2, w1, w2, sp, tp], axis=1).astype(np.float32)
    Rg, Tg = ctx.transforms_from_matches(inp)
    Rr, Tr = ref.transform_from_two_vecs(inp)
    assert np.array_equal(Rg, Rr) and np.array_equal(Tg, Tr)
    ok = slice(400, None)
    assert np.allclose(np.einsum("nij,nkj->nik", Rg[ok], Rg[ok]), np.eye(3), atol=1e-5) and np.all(np.linalg.det(Rg[ok]) > 0.999)


def test_cluster_transforms(ctx, restate, ref):
    rng = np.random.default_rng(10)
    centres_R, centres_T = _rand_rigid(rng, 12, rot_deg=120, trans=1.0)
    Rs, Ts = [], []
    for k in range(12):
        m = int(rng.integers(1, 60))
        dR, dT = _rand_rigid(rng, m, rot_deg=0.6, trans=0.004)
        Rs.append(np.einsum("mij,jk->mik", dR, centres_R[k])); Ts.append(centres_T[k] + dT)
    R = np.concatenate(Rs).astype(np.float32); T = np.concatenate(Ts).astype(np.float32)
    perm = rng.permutation(len(R)); R, T = R[perm], T[perm]
    for tol, ang in ((0.01, 0.0436), (0.004, 0.0005)):
        got = ctx.cluster_transforms(R, T, tol, ang)
        assert np.array_equal(got, restate.cluster_transforms(R, T, tol, ang))
        nc, rlab = ref.cluster_transformations(R, T, tol, ang)              # the reference's CEC
        # same partition; our label = smallest member index = first element the reference emits
        assert len(np.unique(got)) == nc
        for c in range(nc):
            members = np.where(rlab == c)[0]
            assert len(np.unique(got[members])) == 1 and got[members[0]] == members.min()
    assert len(ctx.cluster_transforms(R[:0], T[:0], 0.01, 0.04)) == 0


# ---------------------------------------------------------------------------------------------- K4d
@pytest.mark.parametrize("which", ["poly", "synth"])
def test_penetration_filter_vs_reference(ctx, ref, poly_pair, poly_stages, synth_stages, which):
    """K4d against the reference's own loop + AreTwoPlanesPenetrable (oracle/_ref: ref_penetration_filter) on
    the same planes, rectangles, per-plane points and candidate transforms: identical drop flags.  The
    candidates are the reference's surviving hypotheses plus perturbed copies (most of which penetrate)."""
    if which == "poly":
        g, tgt, src = poly_stages, poly_pair["tgt"], poly_pair["src"]
    else:
        g = synth_stages
        tgt, src, _ = make_pair(n_points=200000, n_planes=20, seed=11)
    leaf = float(g["downsample_distance"][0])

    def side(prefix, cloud, p):
        off, idx = g[p + "_off"], g[p + "_idx"]
        pts, o = [], [0]
        for i in range(len(off) - 1):
            ds = ctx.voxel_downsample(cloud[idx[off[i]:off[i + 1]], :3], leaf)
            pts.append(ds)
            o.append(o[-1] + len(ds))
        assert np.array_equal(np.asarray(o), g[prefix + "plane_ds_offsets"])       # same voxels as the reference
        return dict(planes=g[prefix + "planes"].reshape(-1, 4), corners4=g[prefix + "plane_corners4"], center=g[prefix + "plane_center"],
                    pts=np.concatenate(pts), off=np.asarray(o, np.int32))

    s_side, t_side = side("src_", src, "s"), side("tgt_", tgt, "t")
    R, T = g["mr_R"].reshape(-1, 3, 3), g["mr_T"].reshape(-1, 3)
    rng = np.random.default_rng(12)
    hyps = [np.concatenate([R[i].ravel(), T[i]]) for i in range(len(R))]
    for i in range(len(R)):
        for _ in range(3):
            dR, dT = _rand_rigid(rng, 1, rot_deg=8, trans=0.08)
            hyps.append(np.concatenate([(dR[0] @ R[i]).ravel(), T[i] + dT[0]]))
    hyps = np.asarray(hyps, np.float32)
    mp = g["match_params"]
    got = ctx.penetration_filter(s_side, t_side, hyps, float(mp[0]), float(mp[1]))
    want = ref.penetration_filter(s_side, t_side, hyps, float(mp[0]), float(mp[1]))
    assert 0 < int(want.sum()) < len(want)
    assert np.array_equal(got, want), np.where(got != want)[0]
    assert not got[:len(R)].any()            # the reference kept these


# ------------------------------------------------------------------------------------- whole back-end
def _planes(g, p):
    return Planes(g[p + "_off"], g[p + "_idx"], g[p + "_par"])


def _check_registration_with_reference_planes(ctx, poly_pair, poly_stages, synth_stages, which):
    """registration(T, tgt, src, planes, planes) on the REFERENCE's planes: same hypothesis list, same
    winner.  Tolerance (north_star): rotation <= 0.1 deg, translation <= 1e-3 of the scene diagonal."""
    if which == "poly":
        g, tgt, src = poly_stages, poly_pair["tgt"], poly_pair["src"]
    else:
        g = synth_stages
        tgt, src, _ = make_pair(n_points=200000, n_planes=20, seed=11)
    ctx.set_debug(True)
    ok, T = ctx.register_with_planes(tgt, src, _planes(g, "t"), _planes(g, "s"))
    ctx.set_debug(False)
    assert ok == bool(g["ok"][0])
    diag = float(np.linalg.norm(np.ptp(tgt[:, :3], axis=0)))
    rot, tr = transform_error(T, g["T"], diag)
    assert rot <= 0.1 and tr <= 1e-3
    # stage-level agreement with the reference's dumps
    assert ctx.blob("average_space", np.float32)[0] == g["average_space"][0]
    for side in ("src", "tgt"):
        assert len(ctx.blob(side + "_ds", np.float32)) == len(g[side + "_ds"])
        assert np.allclose(ctx.blob(side + "_center", np.float32), g[side + "_center"], atol=1e-5)
        assert np.array_equal(ctx.blob(side + "_line_planes", np.int32), g[side + "_line_planes"])
        assert np.allclose(ctx.blob(side + "_lines", np.float32), g[side + "_lines"], atol=2e-5)
    assert np.array_equal(ctx.blob("tgt_db_pair", np.int32), g["tgt_db_pair"])
    # all 8 descriptor components are bit-identical to the reference's: component 0 (line-pair distance / scale)
    # comes from the restated 9x9 float SVD solve (K3a, svdsolve.cu), 1..7 are Eigen-ordered dot products
    ours_d, ref_d = ctx.blob("tgt_db_desc", np.float32).reshape(-1, 8), g["tgt_db_desc"].reshape(-1, 8)
    assert np.array_equal(ours_d, ref_d)
    assert np.array_equal(ctx.blob("lines_to_match", np.int32), g["lines_to_match"])
    # hypothesis list after clustering, plane consistency, candidate budget and the penetration filter: the
    # reference's, in its order.  K4a / K4b / K4c / K4d restate the reference's float arithmetic, so on the
    # polyhedron the transforms are bit-identical; on the synthetic scene the list may differ by the few
    # candidates whose penetration test sits on a sample-count threshold and sees plane points that differ in
    # the last bit (voxel centroids: the reference's unstable sort fixes the float summation order)
    R, Tt = ctx.blob("mr_R", np.float32).reshape(-1, 9), ctx.blob("mr_T", np.float32).reshape(-1, 3)
    Rr, Tr = g["mr_R"].reshape(-1, 9), g["mr_T"].reshape(-1, 3)
    if which == "poly":
        assert np.array_equal(R, Rr) and np.array_equal(Tt, Tr)
        assert np.array_equal(ctx.blob("mr_nplanes", np.int32), g["mr_nplanes"])
        assert np.array_equal(ctx.blob("ver_score", np.float32), g["ver_score"])
        assert np.array_equal(T, g["T"])                      # the returned 4x4, bit for bit
    else:
        ours = {(R[i].tobytes(), Tt[i].tobytes()) for i in range(len(R))}
        theirs = {(Rr[i].tobytes(), Tr[i].tobytes()) for i in range(len(Rr))}
        assert len(ours ^ theirs) <= 2 and len(ours & theirs) >= len(theirs) - 2
    sc, scr = ctx.blob("ver_score", np.float32), g["ver_score"]
    assert float(sc.max()) == float(scr.max())
    b, br = int(np.argmax(sc)), int(np.argmax(scr))
    assert np.array_equal(R[b], Rr[br]) and np.array_equal(Tt[b], Tr[br])


@pytest.mark.parametrize("which", ["poly", "synth"])
def test_registration_with_reference_planes(ctx, poly_pair, poly_stages, synth_stages, which):
    """(bit-level comparison with the reference's hypothesis list: run with the reference's literal budget of 200 candidates,
    PLADE/plade.cpp:54 -- the library's default of 1000 is a documented deviation)"""
    ctx.set_param("max_candidates", 200)
    try:
        _check_registration_with_reference_planes(ctx, poly_pair, poly_stages, synth_stages, which)
    finally:
        ctx.set_param("max_candidates", 1000)


def test_registration_end_to_end_polyhedron(ctx, poly_pair):
    """Config 1 (correctness gate): full pipeline incl. GPU RANSAC.  The reference itself only reaches
    the ground truth for some RANSAC seeds; ours must reach it with its fixed seed."""
    ok, T = ctx.register_clouds(poly_pair["tgt"], poly_pair["src"])
    assert ok
    diag = float(np.linalg.norm(np.ptp(poly_pair["tgt"][:, :3], axis=0)))
    rot, tr = transform_error(T, poly_pair["gt"], diag)
    assert rot <= 0.5 and tr <= 5e-3
    rot_p, tr_p = transform_error(T, poly_pair["published"], diag)
    assert rot_p <= 0.5 and tr_p <= 5e-3


def test_registration_end_to_end_synthetic(ctx):
    tgt, src, gt = make_pair(n_points=300000, n_planes=20, seed=11)
    ok, T = ctx.register_clouds(tgt, src)
    assert ok
    diag = float(np.linalg.norm(np.ptp(tgt[:, :3], axis=0)))
    rot, tr = transform_error(T, gt, diag)
    assert rot <= 0.5 and tr <= 5e-3


def test_sharded_verification_matches_single(ctx, poly_pair, poly_stages):
    """world_size-2 hypothesis sharding inside one process: both shards + max-reduce pick the same winner."""
    import plade_b200
    g = poly_stages
    tp, sp = _planes(g, "t"), _planes(g, "s")
    ok, T1 = ctx.register_with_planes(poly_pair["tgt"], poly_pair["src"], tp, sp)
    ctxs = [plade_b200.Context() for _ in range(2)]
    try:
        # a max-allreduce over two ranks that run one after the other: pass 1 records what each rank contributes to every
        # collective of the call (the packed key, then the hash of the hypothesis list and its complement), pass 2 answers
        # each collective with the element-wise maximum over the ranks
        sent = [[], []]
        for r, c in enumerate(ctxs):
            c.set_shard(r, 2, lambda k, r=r: (sent[r].append(k), k)[1])
            c.register_with_planes(poly_pair["tgt"], poly_pair["src"], tp, sp)
        assert len(sent[0]) == len(sent[1]) == 3
        assert sent[0][1:] == sent[1][1:]                     # same hypothesis list on both ranks
        reduced = [max(a, b) for a, b in zip(*sent)]
        for r, c in enumerate(ctxs):
            it = iter(reduced)
            c.set_shard(r, 2, lambda k, it=it: next(it))
            ok2, T2 = c.register_with_planes(poly_pair["tgt"], poly_pair["src"], tp, sp)
            assert ok2 and np.array_equal(T1, T2)
        # ranks that disagree on the hypothesis list must fail loudly, not pick a transform from different lists
        c = ctxs[0]
        bad = [reduced[0], reduced[1] ^ 1, reduced[2]]
        it = iter(bad)
        c.set_shard(0, 2, lambda k, it=it: next(it))
        ok3, _ = c.register_with_planes(poly_pair["tgt"], poly_pair["src"], tp, sp)
        assert not ok3 and "different hypothesis lists" in c.last_error()
        # the same through NCCL inside the library (a one-rank communicator on this GPU: device-side key + ncclAllReduce)
        c.shard_init_nccl(plade_b200.nccl_unique_id(), 0, 1)
        ok4, T4 = c.register_with_planes(poly_pair["tgt"], poly_pair["src"], tp, sp)
        c.shard_finalize()
        assert ok4 and np.array_equal(T1, T4)
    finally:
        for c in ctxs:
            c.close()


def test_config4_hypotheses_vs_reference_compute_overlap(ctx, ref):
    """BASELINE config 4 in the shape the oracle finishes in a minute: a 24-plane synthetic scene (1 M points; bench.py runs the
    5 M-point one), the true transform + 255 perturbations in the fixed shuffled order (seed 7).  Every inlier count of K5
    must equal the reference's ComputeOverlap (PLADE/util.h:612-647, FLANN radius searches) and plade_verify_sharded must
    return the reference's argmax (ties: lowest index)."""
    tgt, src, gt = make_pair(n_points=1000000, n_planes=24, seed=20240611)
    leaf = 4 * ctx.average_spacing(src)
    ds_t, ds_s = ctx.voxel_downsample(tgt[:, :3], leaf), ctx.voxel_downsample(src[:, :3], leaf)
    R, T, true_idx = perturbed_hypotheses(gt, 256, seed=7)
    rc, c_src, whd, _ = ctx.bounding_box(ds_s)
    cen = (np.einsum("hij,j->hi", R, c_src) + T).astype(np.float32)
    ball = float(max(whd) / 2)
    counts = ctx.verify_hypotheses(ds_s, ds_t, R, T, cen, ball, leaf)
    _, want = ref.compute_overlap(ds_s, ds_t, R, T, cen, ball, leaf)
    assert np.array_equal(counts.astype(np.int64), want.astype(np.int64))
    ctx.verify_upload(ds_s, ds_t, leaf)
    bi, bc, _ = ctx.verify_sharded(R, T, cen, ball, leaf)
    assert bi == int(np.argmax(want)) == true_idx and bc == int(want.max())


def test_verify_sharded_nccl_matches_plain_verification(ctx, restate):
    """plade_verify_sharded (config 4: device-side argmax of the shard + ncclAllReduce(u64, max) in the C++ library) against
    the plain verification + host argmax, and against the oracle's counts; on >= 2 GPUs also across two ranks in this process."""
    import plade_b200
    import threading
    rng = np.random.default_rng(9)
    tgt = rng.uniform(0, 1, size=(20000, 3)).astype(np.float32)
    src = (tgt[rng.permutation(20000)[:12000]] + rng.normal(0, 0.002, size=(12000, 3))).astype(np.float32)
    R, T = _rand_rigid(rng, 301, rot_deg=4, trans=0.03)
    R[137], T[137] = np.eye(3), 0
    cen = (np.einsum("hij,j->hi", R, src.mean(0)) + T).astype(np.float32)
    ball, inl = 0.9, 0.01
    want = restate.verify_counts(src, tgt, R, T, cen, ball, inl)
    counts = ctx.verify_hypotheses(src, tgt, R, T, cen, ball, inl)
    assert np.array_equal(counts, want)
    ctx.verify_upload(src, tgt, inl)
    bi, bc, ms = ctx.verify_sharded(R, T, cen, ball, inl)               # no communicator: world 1
    assert bi == int(np.argmax(want)) == 137 and bc == int(want.max()) and ms > 0
    ctx.shard_init_nccl(plade_b200.nccl_unique_id(), 0, 1)              # one-rank communicator: the collective really runs
    try:
        bi, bc, ms = ctx.verify_sharded(R, T, cen, ball, inl)
        assert bi == 137 and bc == int(want.max())
    finally:
        ctx.shard_finalize()
    if plade_b200.device_count() >= 2:
        cs = [plade_b200.Context(d) for d in (0, 1)]
        try:
            plade_b200.shard_init_nccl_all(cs)
            for c in cs:
                c.verify_upload(src, tgt, inl)
            out = [None, None]

            def work(k):
                out[k] = cs[k].verify_sharded(R, T, cen, ball, inl)
            th = [threading.Thread(target=work, args=(k,)) for k in range(2)]
            [t.start() for t in th]
            [t.join() for t in th]
            assert out[0][:2] == out[1][:2] == (137, int(want.max()))
        finally:
            for c in cs:
                c.shard_finalize()
                c.close()


def test_file_overload_and_errors(ctx, poly_pair, tmp_path):
    from tests.plyio import write_ply
    t, s = str(tmp_path / "t.ply"), str(tmp_path / "s.ply")
    write_ply(t, poly_pair["tgt"]); write_ply(s, poly_pair["src"])
    ok, T = ctx.register_files(t, s)
    ok2, T2 = ctx.register_clouds(poly_pair["tgt"], poly_pair["src"])
    assert ok and ok2 and np.array_equal(T, T2)
    ok, T = ctx.register_files(str(tmp_path / "t.xyz"), s)              # only PLY format is accepted
    assert not ok and np.array_equal(T, np.eye(4))
    ok, T = ctx.register_files(str(tmp_path / "missing.ply"), s)
    assert not ok and np.array_equal(T, np.eye(4))
    # too few planes -> false + identity (PLADE/plade.cpp:646-657)
    rng = np.random.default_rng(11)
    blob = rng.normal(size=(20000, 6)).astype(np.float32)
    ok, T = ctx.register_clouds(blob, blob)
    assert not ok and np.array_equal(T, np.eye(4))
    # swap rule: source >= 1.2 x target -> swapped internally, inverse returned (PLADE/plade.cpp:689-704).
    # (a synthetic pair: the polyhedron is symmetric enough that the reference itself flips on it)
    st, ss, gt = make_pair(n_points=200000, n_planes=20, seed=11)
    st = st[::2]                                                        # target half as dense: |src| >= 1.2 |tgt|
    assert len(ss) >= 1.2 * len(st)
    write_ply(t, st); write_ply(s, ss)
    ok, T3 = ctx.register_files(t, s)
    assert ok
    diag = float(np.linalg.norm(np.ptp(st[:, :3], axis=0)))
    rot, tr = transform_error(T3, gt, diag)
    assert rot <= 0.5 and tr <= 5e-3


def test_batch_mode_matches_single_registrations(ctx, poly_pair, tmp_path):
    """plade_register_batch (reference CLI batch mode, PLADE/main.cpp:97-159, spread over worker threads) gives, pair
    by pair, exactly what plade_register_files gives; bad entries fail alone and leave the identity."""
    import plade_b200
    from tests.plyio import write_ply
    pairs, want = [], []
    for k, seed in enumerate([11, 12, 13]):
        st, ss, gt = make_pair(n_points=120000, n_planes=20, seed=seed)
        t, s = str(tmp_path / ("t%d.ply" % k)), str(tmp_path / ("s%d.ply" % k))
        write_ply(t, st); write_ply(s, ss)
        pairs.append((t, s))
        want.append(ctx.register_files(t, s))
    pairs.insert(1, (str(tmp_path / "missing.ply"), pairs[0][1]))
    want.insert(1, (False, np.eye(4, dtype=np.float32)))
    # two workers on the same device: exercises the per-thread contexts and the loader threads
    ok, T = plade_b200.register_batch(pairs, devices=[0, 0])
    if plade_b200.device_count() >= 2:       # workers on different GPUs of one process (the per-device kernel attributes)
        okm, Tm = plade_b200.register_batch(pairs, devices=[0, 1])
        assert okm.tolist() == ok.tolist() and np.array_equal(Tm, T)
    assert ok.tolist() == [bool(w[0]) for w in want]
    for k in range(len(pairs)):
        assert np.array_equal(T[k], want[k][1]), k
    ok0, T0 = plade_b200.register_batch([], devices=[0])
    assert len(ok0) == 0 and T0.shape == (0, 4, 4)
    # the worker contexts are pooled across calls: a second batch reuses them and gives the same answers; release frees them
    ok2, T2 = plade_b200.register_batch(pairs, devices=[0, 0])
    assert ok2.tolist() == ok.tolist() and np.array_equal(T2, T)
    plade_b200.load_library().plade_batch_release()


def test_cli_result_file_format(poly_pair, tmp_path):
    """plade_b200_cli keeps both usages of PLADE/main.cpp and its result-file text format (main.cpp:84-91,139-146)."""
    import subprocess
    from tests.plyio import write_ply
    cli = os.path.join(ROOT, "plade_b200", "plade_b200_cli")
    st, ss, gt = make_pair(n_points=120000, n_planes=20, seed=11)
    t, s = str(tmp_path / "t.ply"), str(tmp_path / "s.ply")
    write_ply(t, st); write_ply(s, ss)
    res = str(tmp_path / "result.txt")
    r = subprocess.run([cli, t, s, res], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    lines = open(res).read().split("\n")
    assert lines[0] == "target: " + t and lines[1] == "source: " + s and lines[2] == "transformation:"
    M = np.array([[float(x) for x in l.split()] for l in lines[3:7]])
    diag = float(np.linalg.norm(np.ptp(st[:, :3], axis=0)))
    rot, tr = transform_error(M, gt, diag)
    assert rot <= 0.5 and tr <= 5e-3
    # options on top of usage 1: the headless report and the .vg plane dumps
    import json
    rep, pre = str(tmp_path / "report.json"), str(tmp_path / "planes")
    r = subprocess.run([cli, "--report", rep, "--dump-planes", pre, t, s, res], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    doc = json.load(open(rep))
    assert doc["target_points"] == len(st) and doc["source_points"] == len(ss) and doc["target_planes"] >= 10 and doc["source_planes"] >= 10
    assert 0 < doc["overlap_ratio"] <= 1 and doc["winner_inliers"] > 0 and len(doc["winner_matched_planes"]) >= 2
    for side, cloud in (("target", st), ("source", ss)):
        vg = open(pre + "_%s.vg" % side).read().split("\n")
        assert vg[0] == "num_points: %d" % len(cloud) and vg[5].startswith("num_groups: ")
    # usage 2: pair list -> blocks separated by a blank line; a name that cannot be opened is skipped with a message
    # and the pairing continues with the next existing name (main.cpp:117-134); a failed registration records the
    # identity; exit code is failure only when every pair failed (main.cpp:150-158)
    rng = np.random.default_rng(3)
    blob = str(tmp_path / "blob.ply")
    write_ply(blob, rng.normal(size=(20000, 6)).astype(np.float32))       # no planes -> registration fails
    lst = str(tmp_path / "file_pairs.txt")
    open(lst, "w").write("%s\n%s\n\n%s\n%s\n%s\n" % (t, str(tmp_path / "nope.ply"), s, blob, blob))
    res2 = str(tmp_path / "results.txt")
    r = subprocess.run([cli, lst, res2], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert "file doesn't exist: " + str(tmp_path / "nope.ply") in r.stderr
    assert "registration of 1 (out of 2) pairs failed" in r.stderr
    assert "the registration result has been written into file: " + res2 in r.stdout
    blocks = open(res2).read().strip("\n").split("\n\n")
    assert len(blocks) == 2
    b0, b1 = blocks[0].split("\n"), blocks[1].split("\n")
    assert b0[:3] == ["target: " + t, "source: " + s, "transformation:"]
    assert b1[:3] == ["target: " + blob, "source: " + blob, "registration failed, an identity matrix is recorded:"]
    assert [l.split() for l in b1[3:7]] == [["1", "0", "0", "0"], ["0", "1", "0", "0"], ["0", "0", "1", "0"], ["0", "0", "0", "1"]]
    M2 = np.array([[float(x) for x in l.split()] for l in b0[3:7]])
    assert np.array_equal(M, M2)


def test_kernel_times_and_cluster_path_parity(ctx):
    """plade_kernel_times reports the CUDA-event clock of the last registration, and the one-launch cluster
    refinement (refine_cluster_kernel) extracts exactly the planes of the multi-kernel path it replaces."""
    import subprocess, sys, json
    tgt, src, gt = make_pair(n_points=150000, n_planes=20, seed=21)
    ctx.set_param("kernel_clock", 1)             # (off by default: two more driver calls per timed launch)
    try:
        ok, T = ctx.register_clouds(tgt, src)
    finally:
        ctx.set_param("kernel_clock", 0)
    assert ok
    k1, k5 = ctx.kernel_times("score_candidates"), ctx.kernel_times("verify")
    assert k1["launches"] >= 2 and k1["ms"] > 0 and k1["algorithmic_bytes"] > 0
    assert k5["launches"] == 1 and k5["ms"] > 0
    kr = ctx.kernel_times("refine_cluster")
    # accept_loop_kernel: one launch per scoring round that had an eligible candidate; 20 B per point for every band pass
    # + 28 B per band point per evaluation
    assert kr["launches"] >= 2 and kr["ms"] > 0 and kr["algorithmic_bytes"] >= 20.0 * 150000 * 20
    with pytest.raises(KeyError):
        ctx.kernel_times("no_such_kernel")
    planes = ctx.extract_planes(tgt, 10000)
    # the same extraction in a fresh process with the cluster kernel disabled (the switch is read once per process)
    code = ("import sys, json, numpy as np; sys.path.insert(0, %r); import plade_b200; from plade_b200.synth import make_pair;"
            "t, s, g = make_pair(n_points=150000, n_planes=20, seed=21); c = plade_b200.Context();"
            "p = c.extract_planes(t, 10000); ok, T = c.register_clouds(t, s);"
            "print(json.dumps({'params': p.params.tolist(), 'sizes': p.sizes().tolist(), 'T': T.tolist()}))" % ROOT)
    env = dict(os.environ, PLADE_NO_CLUSTER_REFINE="1")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    other = json.loads(r.stdout.strip().split("\n")[-1])
    assert planes.sizes().tolist() == other["sizes"]
    assert np.array_equal(planes.params, np.asarray(other["params"], dtype=np.float32))
    assert np.array_equal(T, np.asarray(other["T"], dtype=np.float32))


def _largest_component_numpy(bmp):
    """cross closing without wrapping (out-of-range = 0 for the dilation, 1 for the erosion), 8-connected labels,
    largest by pixel count, first in raster order on ties (R/Bitmap.cpp:154,459,633; R/BitmapPrimitiveShape.cpp:170-173)"""
    from scipy import ndimage
    b = bmp.astype(bool)
    pad = np.pad(b, 1, constant_values=False)
    dil = pad[1:-1, 1:-1] | pad[1:-1, :-2] | pad[1:-1, 2:] | pad[:-2, 1:-1] | pad[2:, 1:-1]
    pad = np.pad(dil, 1, constant_values=True)
    ero = pad[1:-1, 1:-1] & pad[1:-1, :-2] & pad[1:-1, 2:] & pad[:-2, 1:-1] & pad[2:, 1:-1]
    lab, n = ndimage.label(ero, structure=np.ones((3, 3), dtype=int))     # labels in raster order of first pixel
    if n == 0:
        return np.zeros_like(bmp, dtype=np.uint8)
    sizes = np.bincount(lab.ravel(), minlength=n + 1)[1:]
    return (lab == (1 + int(np.argmax(sizes)))).astype(np.uint8)          # argmax: first maximum


def test_largest_component_bit_exact(ctx):
    """K1c: the device labelling (union-find, cc_kernel) against an independent numpy/scipy restatement of the
    reference's closing + Components + largest-label rule, on sparse, dense, tied and degenerate bitmaps."""
    rng = np.random.default_rng(5)
    cases = []
    for (ve, ue, dens) in [(2, 2, 0.5), (7, 13, 0.3), (40, 55, 0.45), (100, 100, 0.55), (64, 200, 0.62), (300, 257, 0.5), (1000, 1000, 0.58), (1, 500, 0.7), (500, 1, 0.7)]:
        cases.append((rng.uniform(size=(ve, ue)) < dens).astype(np.uint8))
    cases.append(np.zeros((20, 30), np.uint8))
    cases.append(np.ones((33, 65), np.uint8))
    tie = np.zeros((20, 40), np.uint8); tie[2:8, 2:10] = 1; tie[12:18, 20:28] = 1          # two equal components: first wins
    cases.append(tie)
    spiral = np.zeros((101, 101), np.uint8)                                                 # long thin component (deep union chains)
    for k in range(0, 50, 4):
        spiral[k, k:101 - k] = 1; spiral[k:101 - k, 100 - k] = 1; spiral[100 - k, k + 4:101 - k] = 1; spiral[k + 4:101 - k, k + 4] = 1
    cases.append(spiral)
    for b in cases:
        got, want = ctx.largest_component(b), _largest_component_numpy(b)
        assert np.array_equal(got, want), (b.shape, int(got.sum()), int(want.sum()))


def test_room_pair_real_scan_swap_path(ctx, tmp_path):
    """BASELINE config 2 on real scanned data (tests/golden/room_decimated.npz: the reference's room pair, source decimated
    8x so that it can travel): the file overload swaps the clouds (source >= 1.2 x target, PLADE/plade.cpp:689-704) and
    returns the inverse.  Bars: <= 2 deg / 3 % of the diagonal from the authors' ground truth -- the
    reference's own results on the same decimated pair are 1.0-1.5 deg / 1.6-1.9 % -- and within 2 deg of the reference."""
    from tests.plyio import write_ply
    g = np.load(os.path.join(ROOT, "tests", "golden", "room_decimated.npz"))
    tgt, src, gt = g["tgt"], g["src"], g["gt"]
    assert len(src) >= 1.2 * len(tgt)
    diag = float(np.linalg.norm(np.ptp(tgt[:, :3], axis=0)))
    for k in range(len(g["ref_T"])):                      # the fixture's own reference results meet the bar
        rot, tr = transform_error(g["ref_T"][k], gt, diag)
        assert rot <= 2.0 and tr <= 0.03
    t, s = str(tmp_path / "room_target.ply"), str(tmp_path / "room_source.ply")
    write_ply(t, tgt); write_ply(s, src)
    ok, T = ctx.register_files(t, s)
    assert ok
    rot, tr = transform_error(T, gt, diag)
    assert rot <= 2.0 and tr <= 0.03, (rot, tr)
    rot_r, tr_r = transform_error(T, g["ref_T"][0].astype(np.float64), diag)
    assert rot_r <= 2.0 and tr_r <= 0.03, (rot_r, tr_r)
    # the array overload on pre-swapped clouds gives the inverse of the same transform, bit for bit up to the inversion
    ok2, T2 = ctx.register_clouds(src, tgt)
    assert ok2 and np.allclose(np.linalg.inv(T2.astype(np.float64)), T, atol=1e-5)


# ---------------------------------------------------------------------------------- end to end vs the reference over seeds
_SWEEP_BARS = {"room_decimated": (2.0, 0.03), "room_full": (2.0, 0.03)}      # SURVEY.md 8(d) config 2; the rest: config 1's bar


def _sweep_case(name):
    g = np.load(os.path.join(ROOT, "tests", "golden", "room_decimated.npz"))
    p = np.load(os.path.join(ROOT, "tests", "golden", "polyhedron_pair.npz"))
    if name == "room_decimated":
        return g["tgt"], g["src"], g["gt"], True
    if name == "polyhedron":
        return p["tgt"], p["src"], p["gt"], False
    if name == "synth_150k":
        return make_pair(n_points=150000, n_planes=20, seed=11) + (False,)
    if name == "synth_1m":
        return make_pair(n_points=1000000, n_planes=20, seed=5) + (False,)
    f = np.load(os.path.join(ROOT, "tests", "golden", "_local", "room_full.npz"))
    return f["tgt"], f["src"], f["gt"], True


@pytest.mark.parametrize("name", ["room_decimated", "polyhedron", "synth_150k", "synth_1m", "room_full"])
def test_seed_sweep_success_rate_vs_reference(ctx, name):
    """The GPU RANSAC draws its candidates from a seeded counter-based generator, the reference seeds rand() from time():
    end-to-end parity is therefore a statement over seeds.  For RANSAC seeds 1..8 the GPU path must (a) land within the bar of
    the ground truth at least as often as the reference does over its own eight seeds (tests/golden/seed_sweep_ref.json, made
    by tests/golden/make_golden_seed_sweep.py from oracle/_ref), and (b) on every seed that does not end in a symmetric flip
    (the reference flips too: 5 of 8 seeds on the full-resolution room pair), be no further from the ground truth than the
    reference's own worst non-flipped run + (0.1 deg, 1e-3 of the diagonal) -- north_star's tolerance.
    room_full is BASELINE config 2 at full resolution (2.3 M-point source, swap path); its 58 MB fixture is local-only."""
    import json
    ref = json.load(open(os.path.join(ROOT, "tests", "golden", "seed_sweep_ref.json")))
    if name == "room_full" and not os.path.exists(os.path.join(ROOT, "tests", "golden", "_local", "room_full.npz")):
        pytest.skip("tests/golden/_local/room_full.npz not present (make_golden_seed_sweep.py --full-room)")
    tgt, src, gt, swapped = _sweep_case(name)
    diag = float(np.linalg.norm(np.ptp(tgt[:, :3], axis=0)))
    bar = _SWEEP_BARS.get(name, (0.5, 5e-3))
    ref_rows = ref["cases"][name]["errors"]
    ref_land = [(r, t) for (r, t, ok) in ref_rows if ok and r <= bar[0] and t <= bar[1]]
    assert ref_land, "the reference never lands on this case: not a parity case"
    # the reference's own error envelope: its runs that did not end in a symmetric flip (rotation error <= 10 deg)
    ref_near = [(r, t) for (r, t, ok) in ref_rows if ok and r <= 10.0]
    worst_rot, worst_tr = max(r for r, _ in ref_near), max(t for _, t in ref_near)
    landed = 0
    try:
        for seed in ref["seeds"]:
            ctx.set_param("seed", seed)
            ok, T = ctx.register_clouds(src, tgt) if swapped else ctx.register_clouds(tgt, src)
            if not ok:
                continue
            Tm = np.linalg.inv(T.astype(np.float64)) if swapped else T
            rot, tr = transform_error(Tm, gt, diag)
            if rot <= bar[0] and tr <= bar[1]:
                landed += 1
            if rot <= 10.0:          # not a symmetric flip: inside the reference's envelope + north_star's tolerance
                assert rot <= worst_rot + 0.1 and tr <= worst_tr + 1e-3, (seed, rot, tr, worst_rot, worst_tr)
    finally:
        ctx.set_param("seed", 2)
    assert landed >= len(ref_land), (landed, len(ref_land))


# ------------------------------------------------------------------ RANSAC building blocks vs the reference's own classes
def _cloud_thresholds(cloud):
    mn, mx = cloud[:, :3].min(0), cloud[:, :3].max(0)
    scale = np.float32(max(mx[0] - mn[0], mx[1] - mn[1]))          # (z ignored: PLADE/plane_extraction.cpp:71-80)
    return np.float32(0.005) * scale, np.float32(0.02) * scale


def _candidate_cases(cloud, planes, rng, n_random=3):
    """(normal, position) candidates: every reference plane as it stands, plus 3-point planes drawn from its members"""
    off, idx, par = planes
    par = par.reshape(-1, 4)
    out = []
    for k in range(len(off) - 1):
        n0 = par[k, :3].astype(np.float32)
        out.append((n0, (-par[k, 3] * n0).astype(np.float32)))
        mem = idx[off[k]:off[k + 1]]
        for _ in range(n_random):
            a, b, c = cloud[rng.choice(mem, 3, replace=False), :3].astype(np.float32)
            nn = np.cross(b - a, c - b).astype(np.float32)
            if float(nn @ nn) < 1e-6:
                continue
            out.append(((nn / np.float32(np.sqrt(nn @ nn))).astype(np.float32), a))
    return out


@pytest.mark.parametrize("which", ["poly", "synth"])
def test_connected_component_vs_reference_bitmap_code(ctx, ref, restate, poly_pair, poly_stages, synth_stages, which):
    """K1c against the reference's OWN BitmapPrimitiveShape::ConnectedComponent (R/BitmapPrimitiveShape.cpp:155-265 with
    BuildBitmap, BitmapExtent / InBitmap, DilateCross / ErodeCross, Components of R/Bitmap.cpp) -- not a restatement: the
    inliers of a plane at 3 eps are parametrised and rasterised by the product's planefit.h functions (bit-exact against
    PlanePrimitiveShape::Parameters, tests/test_oracle_cpu.py), labelled by cc_kernel on the device, and the member set must
    equal the reference's, point for point."""
    cloud, g = (poly_pair["tgt"], poly_stages) if which == "poly" else (synth_small_tgt(), synth_stages)
    import plade_b200
    eps, beps = _cloud_thresholds(cloud)
    rng = np.random.default_rng(12)
    n_checked = 0
    for nrm, pos in _candidate_cases(cloud, (g["t_off"], g["t_idx"], g["t_par"]), rng, n_random=2)[:24]:
        pl = np.array([nrm[0], nrm[1], nrm[2], np.float32(pos @ nrm)], np.float32)
        _, mask = restate.score_planes(cloud, pl[None], 3 * eps, 0.8, want_mask=True)
        inl = np.nonzero(mask)[0].astype(np.int32)
        if len(inl) < 50:
            continue
        want = ref.connected_component(cloud, nrm, pos, inl, beps, True)
        uv, _, _ = plade_b200.plane_parameters(nrm, pos, cloud[inl, :3])
        (ue, ve), pix = plade_b200.bitmap_layout(uv, beps)
        if ue * ve > (1 << 20):
            continue
        bmp = np.zeros(ue * ve, np.uint8)
        bmp[pix] = 1
        comp = ctx.largest_component(bmp.reshape(ve, ue)).reshape(-1)
        got = inl[comp[pix] != 0]
        assert np.array_equal(got, want), (len(got), len(want))
        n_checked += 1
    assert n_checked >= 8


@pytest.mark.parametrize("which", ["poly", "synth"])
def test_refine_candidate_chain_vs_reference(ctx, ref, poly_pair, poly_stages, synth_stages, which):
    """Rows a5-a8 in one: the acceptance chain of a candidate (GlobalScore at 3 eps, ConnectedComponent, WeightedScore,
    LeastSquaresFit x <= 3, RansacShapeDetector.cpp:613-655) on the device (band_compact_kernel + refine_cluster_kernel through
    plade_refine_candidate) against the same chain driven through the reference's own Candidate / PlanePrimitiveShape / Plane /
    octree classes (oracle ref_refine_candidate).  With min_support above the cloud size no refit can be accepted: the result
    is the candidate's own largest component and the member sets must be identical.  With the real min_support the refits
    take part: the device sums the members in double where the reference sums sequentially in float (DESIGN.md deviations),
    so the fitted planes may differ in the last bits -- bars: normal within 0.02 deg, support within 0.2 %, member sets
    differ by <= 0.2 %, weighted score within 1e-3 relative; no sign flip of the normal."""
    cloud, g = (poly_pair["tgt"], poly_stages) if which == "poly" else (synth_small_tgt(), synth_stages)
    eps, beps = _cloud_thresholds(cloud)
    rng = np.random.default_rng(3)
    cases = _candidate_cases(cloud, (g["t_off"], g["t_idx"], g["t_par"]), rng, n_random=2)[:18]
    exact = refit = 0
    for nrm, pos in cases:
        # (a) no refit can be accepted
        size_r, n_r, p_r, mem_r, _ = ref.refine_candidate(cloud, nrm, pos, eps, 0.8, beps, len(cloud) + 1)
        size_g, n_g, p_g, mask_g, _, _ = ctx.refine_candidate(cloud, nrm, pos, len(cloud) + 1)
        assert size_g == size_r and np.array_equal(np.nonzero(mask_g)[0], mem_r)
        assert np.array_equal(n_g, n_r)
        exact += 1
        # (b) the real chain
        size_r, n_r, p_r, mem_r, tr = ref.refine_candidate(cloud, nrm, pos, eps, 0.8, beps, 1000)
        size_g, n_g, p_g, mask_g, evals, score_g = ctx.refine_candidate(cloud, nrm, pos, 1000)
        if size_r < 200:
            assert size_g < 400
            continue
        mem_g = np.nonzero(mask_g)[0]
        diff = len(np.setxor1d(mem_g, mem_r))
        dn = float(np.dot(n_g / np.linalg.norm(n_g), n_r / np.linalg.norm(n_r)))
        assert dn > 0, "normal sign flipped"
        assert np.degrees(np.arccos(min(1.0, dn))) <= 0.02, (np.degrees(np.arccos(min(1.0, dn))), size_g, size_r)
        assert abs(size_g - size_r) <= max(2, 0.002 * size_r) and diff <= max(4, 0.002 * size_r), (size_g, size_r, diff)
        score_r = ref.weighted_score(cloud, n_r, p_r, mem_r, 3 * eps)
        assert abs(score_g - score_r) <= 1e-3 * score_r, (score_g, score_r)
        refit += 1
    assert exact >= 10 and refit >= 8


def test_accept_loop_on_the_device_extracts_the_same_planes(ctx, poly_pair):
    """accept_loop_kernel (the pool walk of a scoring round on the device, one launch per round) against the host-driven
    walk it replaces (ransac_batch = 1: one candidate per host round trip): same planes, same supports, same parameters."""
    clouds = [poly_pair["tgt"], make_pair(n_points=200000, n_planes=20, seed=31)[0]]
    try:
        for cloud in clouds:
            ctx.set_param("ransac_batch", 64)
            a = ctx.detect_planes(cloud, 1250)
            ctx.set_param("ransac_batch", 1)
            b = ctx.detect_planes(cloud, 1250)
            assert len(a) == len(b) and len(a) >= 7
            assert a.sizes().tolist() == b.sizes().tolist()
            assert np.array_equal(a.params, b.params)
            assert np.array_equal(a.indices, b.indices)
    finally:
        ctx.set_param("ransac_batch", 64)


def test_scoring_on_the_live_slots_extracts_the_same_planes(ctx, poly_pair):
    """A scoring round scores only the candidate slots that hold a plane (score_live >= 1: slot list from
    gen_candidates_kernel's validity bits, one candidate per thread, predicate without the early out) and runs stage 2 with the
    same candidate-per-thread kernel (score_live = 2, the default) -- the counts must land where scoring every slot with the
    point-per-thread stage 2 (score_live = 0) puts them, so the detection is the same plane by plane and point by point."""
    clouds = [poly_pair["tgt"], poly_pair["src"], make_pair(n_points=200000, n_planes=20, seed=31)[0]]
    try:
        for cloud in clouds:
            for seed in (2, 11):
                ctx.set_param("seed", seed)
                ctx.set_param("score_live", 0)
                b = ctx.detect_planes(cloud, 1250)
                assert len(b) >= 6
                for mode in (2, 1):
                    ctx.set_param("score_live", mode)
                    a = ctx.detect_planes(cloud, 1250)
                    assert len(a) == len(b)
                    assert a.sizes().tolist() == b.sizes().tolist()
                    assert np.array_equal(a.params, b.params)
                    assert np.array_equal(a.indices, b.indices)
    finally:
        ctx.set_param("score_live", 2)
        ctx.set_param("seed", 2)
