"""Synthetic planar indoor scenes: the workload generator of bench.py and
of the size-independent parity tests.  numpy only; no reference code involved.

Scene: an axis-aligned 6 x 4 x 3 room scaled so that its bounding-box diagonal is ~2 (the unit of the
reference's sample data): 6 shell planes + furniture rectangles with normals from {x, y, z} plus a few
tilted by 15-40 degrees.  Points are area-proportional and uniform in-plane, with Gaussian noise along
the normal; normals = plane normal oriented towards the room centre + jitter, renormalised.
A pair = two independent samplings of the scene, each cropped to a half-space so that they share
`overlap` of the scene, the source moved by a random rigid transform whose inverse is the ground truth.
"""
import numpy as np


def _rot_from_axis_angle(axis, angle):
    axis = axis / np.linalg.norm(axis)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + np.sin(angle) * K + (1 - np.cos(angle)) * (K @ K)


def random_rotation(rng):
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def make_scene(n_planes=20, seed=20240611):
    """Returns a list of rectangles: (origin[3], edge_u[3], edge_v[3], normal[3])."""
    rng = np.random.default_rng(seed)
    dims = np.array([6.0, 4.0, 3.0])
    dims = dims * (2.0 / np.linalg.norm(dims))
    rects = []
    # shell: the 6 faces of the room, normals pointing inwards.  A plain box is symmetric under a
    # half turn (two cropped views then align better the wrong way round than the right way), so
    # three faces are skewed by fixed, different angles: no two shell planes stay parallel.
    skew = {(0, 0): (2, 12.0), (1, 1): (2, -8.0), (2, 1): (0, 6.0)}     # (axis, side) -> (rotation axis, degrees)
    for ax in range(3):
        u, v = [a for a in range(3) if a != ax]
        for side in (0, 1):
            o = np.zeros(3)
            o[ax] = side * dims[ax]
            eu, ev = np.zeros(3), np.zeros(3)
            eu[u], ev[v] = dims[u], dims[v]
            n = np.zeros(3)
            n[ax] = 1.0 if side == 0 else -1.0
            if (ax, side) in skew:
                rax, deg = skew[(ax, side)]
                a = np.zeros(3)
                a[rax] = 1.0
                R = _rot_from_axis_angle(a, np.deg2rad(deg))
                c = o + 0.5 * (eu + ev)
                eu, ev, n = R @ eu, R @ ev, R @ n
                o = c - 0.5 * (eu + ev)
            rects.append((o, eu, ev, n))
    n_furn = max(0, n_planes - 6)
    n_tilt = min(4, n_furn // 3)
    for k in range(n_furn):
        ax = int(rng.integers(0, 3))
        u, v = [a for a in range(3) if a != ax]
        su, sv = rng.uniform(0.1, 0.5, size=2)
        eu, ev = np.zeros(3), np.zeros(3)
        eu[u], ev[v] = su * dims[u], sv * dims[v]
        o = np.zeros(3)
        o[u] = rng.uniform(0.05, 0.95 - su) * dims[u]
        o[v] = rng.uniform(0.05, 0.95 - sv) * dims[v]
        o[ax] = rng.uniform(0.15, 0.85) * dims[ax]
        n = np.zeros(3)
        n[ax] = 1.0
        if k < n_tilt:
            R = _rot_from_axis_angle(rng.normal(size=3), np.deg2rad(rng.uniform(15, 40)))
            c = o + 0.5 * (eu + ev)
            eu, ev, n = R @ eu, R @ ev, R @ n
            o = c - 0.5 * (eu + ev)
        rects.append((o, eu, ev, n))
    return rects, dims


def sample_scene(rects, dims, n_points, rng, noise=0.001, normal_jitter=0.02):
    areas = np.array([np.linalg.norm(np.cross(eu, ev)) for (_, eu, ev, _) in rects])
    counts = np.floor(areas / areas.sum() * n_points).astype(np.int64)
    counts[0] += n_points - counts.sum()
    centre = 0.5 * dims
    pts, nrms, ids = [], [], []
    for k, ((o, eu, ev, n), c) in enumerate(zip(rects, counts)):
        a, b = rng.random(c), rng.random(c)
        p = o + a[:, None] * eu + b[:, None] * ev + rng.normal(0.0, noise, size=c)[:, None] * n
        nn = n if np.dot(centre - (o + 0.5 * (eu + ev)), n) >= 0 else -n
        m = nn + rng.normal(0.0, normal_jitter, size=(c, 3))
        m /= np.linalg.norm(m, axis=1, keepdims=True)
        pts.append(p)
        nrms.append(m)
        ids.append(np.full(c, k, dtype=np.int32))
    P, N, I = np.concatenate(pts), np.concatenate(nrms), np.concatenate(ids)
    perm = rng.permutation(len(P))       # scan order is not plane order
    return P[perm], N[perm], I[perm]


def make_pair(n_points=2_000_000, n_planes=20, seed=20240611, overlap=0.4, noise=0.001, return_ids=False, rotation="moderate"):
    """(target[n,6], source[m,6], gt[4,4]) float32; gt maps source onto target."""
    rects, dims = make_scene(n_planes, seed)
    rng_t = np.random.default_rng([seed, 1])
    rng_s = np.random.default_rng([seed, 2])
    rng_x = np.random.default_rng([seed, 3])
    # target = the whole scene; source = an independent, denser sampling cropped to the half-space
    # x >= (1 - overlap) * Lx, so that `overlap` of the scene is shared and both clouds have ~n_points
    Pt, Nt, It = sample_scene(rects, dims, n_points, rng_t, noise)
    Ps, Ns, Is = sample_scene(rects, dims, int(round(n_points / overlap)), rng_s, noise)
    ms = Ps[:, 0] >= (1.0 - overlap) * dims[0]
    Ps, Ns, Is = Ps[ms], Ns[ms], Is[ms]
    # move the source: p_src = R p + t ; ground truth = inverse.
    # "moderate" (default): yaw U(-35, 35) deg + a tilt <= 10 deg.  PLADE's plane normals keep whatever
    # sign the Jacobi eigen-solver produces (positive along the dominant axis; correct_normal never
    # flips, PLADE/plane_extraction.cpp:43-58) and its descriptors are signed, so under a rotation
    # drawn uniformly from SO(3) ("so3") roughly half of the normals change sign between the two clouds
    # and the REFERENCE itself fails on 20-plane scenes; the moderate motion keeps the pair solvable.
    if rotation == "so3":
        R = random_rotation(rng_x)
    else:
        yaw = np.deg2rad(rng_x.uniform(-35.0, 35.0))
        tilt_axis = np.array([np.cos(rng_x.uniform(0, 2 * np.pi)), np.sin(rng_x.uniform(0, 2 * np.pi)), 0.0])
        R = _rot_from_axis_angle(tilt_axis, np.deg2rad(rng_x.uniform(0.0, 10.0))) @ _rot_from_axis_angle(np.array([0, 0, 1.0]), yaw)
    t = rng_x.uniform(-1.0, 1.0, size=3)
    Ps2 = Ps @ R.T + t
    Ns2 = Ns @ R.T
    gt = np.eye(4)
    gt[:3, :3] = R.T
    gt[:3, 3] = -R.T @ t
    tgt = np.concatenate([Pt, Nt], axis=1).astype(np.float32)
    src = np.concatenate([Ps2, Ns2], axis=1).astype(np.float32)
    if return_ids:
        return tgt, src, gt, It, Is
    return tgt, src, gt


def transform_error(T, G, diag=None):
    """(rotation error in degrees, translation error [relative to diag if given])."""
    T, G = np.asarray(T, dtype=np.float64), np.asarray(G, dtype=np.float64)
    Rd = T[:3, :3] @ G[:3, :3].T
    ang = np.degrees(np.arccos(np.clip((np.trace(Rd) - 1) / 2, -1, 1)))
    te = np.linalg.norm(T[:3, 3] - G[:3, 3])
    return ang, (te / diag if diag else te)


def perturbed_hypotheses(gt, n, seed=7, rot_sigma_deg=5.0, trans_sigma=0.05):
    """Config 4: the true transform + n-1 perturbations in a fixed shuffled order. Returns R[n,3,3], T[n,3]."""
    rng = np.random.default_rng(seed)
    Rs, Ts = [gt[:3, :3]], [gt[:3, 3]]
    for _ in range(n - 1):
        aa = rng.normal(0.0, np.deg2rad(rot_sigma_deg), size=3)
        ang = np.linalg.norm(aa)
        dR = _rot_from_axis_angle(aa if ang > 0 else np.array([1.0, 0, 0]), ang)
        Rs.append(dR @ gt[:3, :3])
        Ts.append(gt[:3, 3] + rng.normal(0.0, trans_sigma, size=3))
    perm = rng.permutation(n)
    return np.array(Rs, dtype=np.float32)[perm], np.array(Ts, dtype=np.float32)[perm], int(np.where(perm == 0)[0][0])
