"""CPU tests that pin the oracle (oracle/) — the reference ships no tests or golden vectors
(SURVEY.md §4), so the oracle is checked against independent numpy / scipy computations,
hand-built known-answer lines and the committed golden fixtures in tests/golden/."""
import json
import os

import numpy as np
import pytest
from scipy.optimize import least_squares
from scipy.spatial import cKDTree

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


# ---------------------------------------------------------------- A1 known-answer lines
def _ring(fn, n=1800, fov=2 * np.pi):
    a = np.arange(n) * (fov / n)
    r = np.array([fn(t) for t in a])
    x = np.zeros((n, 4), np.float32)
    x[:, 0] = r * np.cos(a)
    x[:, 1] = r * np.sin(a)
    x[:, 2] = 0.2
    x[:, 3] = 10.0
    return x, a


def test_a1_degenerate_sizes(orc):
    for n in (0, 1, 5, 10):
        s, f = orc.detect_feature_points(np.ones((n, 4), np.float32))
        assert len(s) == 0 and len(f) == 0


def test_a1_flat_wall_gives_flat_points_only(orc):
    # a wall at x = 5 seen over +-40 degrees: every part yields exactly one flat point, no corner
    x, _ = _ring(lambda t: 5.0 / np.cos(t - 0.7), n=600, fov=1.4)
    s, f = orc.detect_feature_points(x)
    assert len(s) == 0
    assert len(f) == 50  # thPartNum parts x thNumFlat, FE.cpp:355-356
    flags = orc.detect_feature_flags(x)
    assert set(np.unique(flags)) <= {0, 1, 2, 3}


def test_a1_square_room_corners(orc):
    # square room seen from the centre: the four 90-degree corners are flagged 150 (two-plane test)
    x, a = _ring(lambda t: 4.0 / max(abs(np.cos(t)), abs(np.sin(t))))
    s, f = orc.detect_feature_points(x)
    # the walk strides by 4 over flat stretches (count_num, FE.cpp:603), so a corner is only tested
    # when the stride lands on it: every detection is a true corner, and at least two are found
    true_corners = np.array([np.pi / 4, 3 * np.pi / 4, 5 * np.pi / 4, 7 * np.pi / 4])
    assert 2 <= len(s) <= 12
    for t in a[s]:
        assert np.abs(true_corners - t).min() < np.deg2rad(1.0)
    flags = orc.detect_feature_flags(x)
    assert (flags[s] == 150).all()


def test_a1_occlusion_jump_is_break_point(orc):
    # a near object in front of a far wall: the near edge of each depth jump is a break point (100)
    def rng(t):
        return 3.0 / np.cos(t - 0.5) if 0.3 < t < 0.7 else 8.0 / np.cos(t - 0.5)
    x, a = _ring(rng, n=800, fov=1.0)
    s, f = orc.detect_feature_points(x)
    flags = orc.detect_feature_flags(x)
    assert (flags == 100).sum() >= 2
    for i in np.nonzero(flags == 100)[0]:
        d = np.linalg.norm(x[i, :3])
        assert d < 4.5  # on the near object, not on the far wall


def test_a1_far_points_use_small_window(orc):
    # beyond 50 m the curvature window is 2 and flat points are always accepted (FE.cpp:424-428, 524)
    x, _ = _ring(lambda t: 60.0 / np.cos(t - 0.35), n=700, fov=0.7)
    s, f = orc.detect_feature_points(x)
    assert len(f) > 50  # far flat points bypass the one-per-part limit


def test_a1_labels_glue(orc, synth):
    T = synth.make_T(np.eye(3), np.array([1.0, 0.5, 0.0]))
    x, ring, _ = synth.vlp16_scan(T, seed=3, n_az=600)
    lab = orc.extract_scan(x, ring, 16)
    for r in range(16):
        idx = np.nonzero(ring == r)[0]
        s, f = orc.detect_feature_points(x[idx])
        assert np.array_equal(np.sort(idx[s]), np.nonzero((lab == 1) & (ring == r))[0])
        assert np.array_equal(np.sort(idx[f]), np.nonzero((lab == 2) & (ring == r))[0])
    assert np.array_equal(lab, orc.extract_scan(x, ring, 16, threads=4))


# ---------------------------------------------------------------- A2 / A3 / A4 / A5 / A6
def test_a2_ring_and_time(orc, synth):
    T = synth.make_T(np.eye(3), np.zeros(3))
    x, ring, s = synth.vlp16_scan(T, seed=5)
    line, rt = orc.velo_ring_time(x)
    assert np.array_equal(line, ring.astype(np.int16))
    assert np.abs(rt - s).max() < 2e-3 and rt.min() >= 0 and rt.max() <= 1.0


def test_a3_hori_filter(orc):
    off = np.array([0, 10, 20, 30, 40], np.uint32)
    xyz = np.array([[1, 0, 0], [0.005, 0, 0], [2, 0, 0], [3, 0, 0], [4, 0, 0]], np.float32)
    line = np.array([0, 1, 6, 5, 2], np.uint8)
    keep, rt = orc.hori_filter(off, xyz, line)
    assert keep.tolist() == [1, 0, 0, 1, 1]
    assert np.allclose(rt, [0, 0, 0, 0.75, 1.0])


def test_a4_undistort_matches_numpy(orc, synth):
    rng = np.random.default_rng(0)
    x = rng.uniform(-10, 10, (500, 4)).astype(np.float32)
    s = rng.uniform(0, 1, 500).astype(np.float32)
    phi = np.array([0.02, -0.01, 0.05])
    dR, dt = synth.rotvec_to_R(phi), np.array([0.1, -0.05, 0.02])
    out = orc.undistort(x, s, dR, dt)
    for i in range(0, 500, 50):
        Rs = synth.rotvec_to_R(phi * float(s[i]))  # slerp(I, q, s) = Exp(s * phi)
        p = dR.T @ (Rs @ x[i, :3].astype(np.float64) + float(s[i]) * dt - dt)
        assert np.abs(out[i, :3] - p).max() < 2e-6
    # s = 1 maps a point onto itself: dR^T (dR p + dt - dt) = p
    one = orc.undistort(x, np.ones(500, np.float32), dR, dt)
    assert np.abs(one[:, :3] - x[:, :3]).max() < 2e-6


def test_a5_cube_rule(orc):
    assert orc.cube_index([0, 0, 0]) == 10 + 21 * 10 + 21 * 21 * 5
    assert orc.cube_index([24.9, 0, 0]) == orc.cube_index([0, 0, 0])
    assert orc.cube_index([25.1, 0, 0]) == orc.cube_index([0, 0, 0]) + 1
    assert orc.cube_index([-25.1, 0, 0]) == orc.cube_index([0, 0, 0]) - 1
    assert orc.cube_index([0, 0, 600.0]) == 5000
    T = np.eye(4)
    T[:3, 3] = [1, 2, 3]
    assert np.allclose(orc.point_to_map([1, 1, 1], T), [2, 3, 4])


def test_a6_voxel_properties(orc):
    rng = np.random.default_rng(1)
    x = rng.uniform(-5, 5, (3000, 4)).astype(np.float32)
    out = orc.voxel_downsample(x, 0.5)
    # every centroid lies in a distinct voxel; voxel order ascending; mass is preserved per voxel
    inv = np.float32(1.0) / np.float32(0.5)
    key = lambda p: tuple(np.floor(p[:, :3] * inv).astype(int).T)
    kx, ko = np.stack(key(x), 1), np.stack(key(out), 1)
    assert len({tuple(k) for k in ko}) == out.shape[0] == len({tuple(k) for k in kx})
    mn = kx.min(0)
    dim = kx.max(0) - mn + 1
    lin = (ko - mn) @ np.array([1, dim[0], dim[0] * dim[1]])
    assert (np.diff(lin) > 0).all()
    # idempotent up to float rounding, and a single point is returned unchanged
    assert orc.voxel_downsample(out, 0.5).shape == out.shape
    assert np.array_equal(orc.voxel_downsample(x[:1], 0.5), x[:1])


# ---------------------------------------------------------------- k-NN and fits
def test_knn_kdtree_equals_brute_and_scipy(orc):
    rng = np.random.default_rng(2)
    cloud = rng.uniform(-20, 20, (5000, 4)).astype(np.float32)
    q = rng.uniform(-20, 20, (200, 4)).astype(np.float32)
    idx, d2 = orc.knn5_kdtree(cloud, q)
    tree = cKDTree(cloud[:, :3].astype(np.float64))
    _, sidx = tree.query(q[:, :3].astype(np.float64), k=5)
    for i in range(200):
        bi, bd = orc.knn5_brute(cloud, q[i, :3])
        assert np.array_equal(idx[i], bi) and np.array_equal(d2[i], bd)
        assert set(idx[i]) == set(sidx[i])
        assert (np.diff(d2[i]) >= 0).all()


def test_knn_ties_broken_by_index(orc):
    cloud = np.zeros((12, 4), np.float32)
    cloud[:, 0] = [1, -1, 1, -1, 1, -1, 1, -1, 5, 5, 5, 5]  # eight points at distance 1
    idx, d2 = orc.knn5_brute(cloud, np.zeros(3, np.float32))
    assert idx.tolist() == [0, 1, 2, 3, 4]
    kidx, _ = orc.knn5_kdtree(cloud, np.zeros((1, 4), np.float32))
    assert kidx[0].tolist() == [0, 1, 2, 3, 4]


def test_line_and_plane_fit_against_numpy(orc, synth):
    ms, mc = synth.feature_map(20000, 3000, seed=4)
    m = orc.Map()
    m.set(orc.SURF_LOCAL, ms)
    m.set(orc.CORNER_LOCAL, mc)
    T = np.eye(4)
    q = ms[::200].copy()
    q[:, :3] += 0.02
    feat, nf, M, nn = m.associate_plane(q, T, 1.0)
    assert nf > 50 and nn == nf
    tree = cKDTree(ms[:, :3].astype(np.float64))
    for i in np.nonzero(feat[:, 10] >= 0)[0][:40]:
        _, nb = tree.query(q[i, :3].astype(np.float64), k=5)
        A = ms[nb, :3].astype(np.float64)
        xs, *_ = np.linalg.lstsq(A, -np.ones(5), rcond=None)
        n = xs / np.linalg.norm(xs)
        assert np.abs(np.abs(feat[i, 6:9] @ n) - 1) < 1e-5
        # projection lies on the plane and the error is the point-plane distance
        assert abs(feat[i, 3:6] @ n + 1 / np.linalg.norm(xs)) < 1e-4
        assert abs(feat[i, 9] - abs(q[i, :3].astype(np.float64) @ n + 1 / np.linalg.norm(xs))) < 1e-4
    qc = mc[::30].copy()
    qc[:, :3] += 0.03
    lf, nl = m.associate_line(qc, T, 1.0)
    assert nl > 20
    tc = cKDTree(mc[:, :3].astype(np.float64))
    for i in np.nonzero(lf[:, 10] >= 0)[0][:40]:
        _, nb = tc.query(qc[i, :3].astype(np.float64), k=5)
        P = mc[nb, :3].astype(np.float64)
        w, V = np.linalg.eigh(np.cov(P.T, bias=True))
        assert w[2] > 3 * w[1]
        d = (lf[i, 3:6] - lf[i, 6:9]) / 0.2
        assert abs(abs(d @ V[:, 2]) - 1) < 1e-4


def test_localizability(orc):
    rng = np.random.default_rng(5)
    n = rng.normal(size=(200, 3))
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    M = n.T @ n
    sv = np.linalg.svd(n, compute_uv=False)
    assert abs(orc.localizability(M, 200) - sv[2]) < 1e-9
    assert orc.localizability(M, 10) == -1.0


# ---------------------------------------------------------------- residuals and solver
def _rand_pose(rng):
    return np.concatenate([rng.uniform(-2, 2, 3), rng.uniform(-0.4, 0.4, 3)])


@pytest.mark.parametrize("kind,wt", [(0, 0.0), (1, 0.0), (1, 0.0003)])
def test_jacobian_matches_finite_differences(orc, kind, wt):
    rng = np.random.default_rng(6 + kind)
    Tbl = np.eye(4)
    Tbl[:3, 3] = [0.05, -0.02, 0.1]
    for _ in range(20):
        x = _rand_pose(rng)
        p = rng.uniform(-8, 8, 3)
        if kind == 0:
            a = rng.uniform(-8, 8, 3)
            f = np.concatenate([p, a, a + 0.2 * rng.normal(size=3), [0, 1, 0]])
        else:
            n = rng.normal(size=3)
            n /= np.linalg.norm(n)
            f = np.concatenate([p, rng.uniform(-8, 8, 3), n, [0, 1, 0]])
        r, J = orc.residual(kind, f, x, Tbl, wt)
        Jn = np.zeros_like(J)
        for k in range(6):
            h = 1e-6
            xp, xm = x.copy(), x.copy()
            xp[k] += h
            xm[k] -= h
            Jn[:, k] = (orc.residual(kind, f, xp, Tbl, wt)[0] - orc.residual(kind, f, xm, Tbl, wt)[0]) / (2 * h)
        assert np.abs(J - Jn).max() <= 1e-5 * max(1.0, np.abs(Jn).max())


def test_so3_exp_log_roundtrip(orc, synth):
    rng = np.random.default_rng(7)
    for _ in range(50):
        phi = rng.uniform(-1.5, 1.5, 3)
        q, R = orc.so3_exp(phi)
        assert np.allclose(R, synth.rotvec_to_R(phi), atol=1e-12)
        assert np.allclose(orc.so3_log(q), phi, atol=1e-12)
    q, R = orc.so3_exp(np.zeros(3))
    assert np.allclose(q, [1, 0, 0, 0]) and np.allclose(R, np.eye(3))


def test_accumulate_is_sum_of_rows_with_huber(orc):
    rng = np.random.default_rng(8)
    Tbl = np.eye(4)
    x = _rand_pose(rng)
    lf = np.zeros((30, 12))
    pf = np.zeros((40, 12))
    for f in lf:
        a = rng.uniform(-5, 5, 3)
        f[:] = np.concatenate([rng.uniform(-5, 5, 3), a, a + 0.2 * rng.normal(size=3), [0, 1, 0]])
    for f in pf:
        n = rng.normal(size=3)
        f[:] = np.concatenate([rng.uniform(-5, 5, 3), rng.uniform(-5, 5, 3), n / np.linalg.norm(n), [0, 1, 0]])
    pf[3, 10] = 0  # |error| <= 1e-5 features are skipped (EST.cpp:1313)
    lf[5, 10] = -1
    a_h = 0.1 / 1.5e-3
    H, g, c = orc.accumulate(lf, pf, x, Tbl, 0.0, a_h)
    Hn, gn, cn = np.zeros((6, 6)), np.zeros(6), 0.0
    for kind, F in ((0, lf), (1, pf)):
        for f in F:
            if f[10] != 1:
                continue
            r, J = orc.residual(kind, f, x, Tbl, 0.0)
            s = float(r @ r)
            k1, rho = (np.sqrt(a_h / np.sqrt(s)), 2 * a_h * np.sqrt(s) - a_h**2) if s > a_h**2 else (1.0, s)
            Hn += (k1 * J).T @ (k1 * J)
            gn += (k1 * J).T @ (k1 * r)
            cn += 0.5 * rho
    assert np.allclose(H, Hn, rtol=1e-12) and np.allclose(g, gn, rtol=1e-12) and abs(c - cn) <= 1e-12 * cn
    H4, g4, c4 = orc.accumulate(lf, pf, x, Tbl, 0.0, a_h, threads=4)
    assert np.allclose(H, H4, rtol=1e-12) and abs(c - c4) <= 1e-12 * c


def test_estimate_converges_and_matches_scipy(orc, synth):
    """The restated dogleg (Ceres is absent) against scipy's trust-region solver with the same
    Huber loss on fixed correspondences, and against ground truth."""
    T_true = synth.make_T(synth.rot_z(0.3), np.array([-3.0, -1.0, 0.2]))
    x, ring, _ = synth.vlp16_scan(T_true, seed=1001)
    lab = orc.extract_scan(x, ring, 16)
    corner = orc.voxel_downsample(x[lab == 1], 0.4)
    surf = orc.voxel_downsample(x[lab == 2], 0.2)
    ms, mc = synth.feature_map(100_000, 5_000, seed=1002)
    m = orc.Map()
    m.set(orc.SURF_LOCAL, ms)
    m.set(orc.CORNER_LOCAL, mc)
    T0 = T_true @ synth.s1_offset_pose()
    q0, _ = orc.so3_exp(synth.R_to_rotvec(T0[:3, :3]))
    P, q, st = m.estimate(corner, surf, np.eye(4), T0[:3, 3], q0)
    assert np.abs(P - T_true[:3, 3]).max() < 5e-3 and st[0] <= 5
    # one outer iteration with frozen correspondences vs scipy
    lf, _ = m.associate_line(corner, T0, 25.0)
    pf, _, _, _ = m.associate_plane(surf, T0, 25.0)
    a_h = 0.1 / 1.5e-3
    x0 = np.concatenate([T0[:3, 3], synth.R_to_rotvec(T0[:3, :3])])

    def res(xx):
        out = []
        for kind, F in ((0, lf), (1, pf)):
            for f in F:
                if f[10] == 1:
                    out.append(orc.residual(kind, f, xx, np.eye(4), 0.0, jac=False)[0][:1])
        return np.concatenate(out)

    def cost(xx):
        return orc.accumulate(lf, pf, xx, np.eye(4), 0.0, a_h)[2]

    sol = least_squares(res, x0, loss="huber", f_scale=a_h, xtol=1e-12, ftol=1e-12, gtol=1e-12)
    p1 = orc.est_params(max_outer=1, max_inner=50)
    P1, q1, st1 = m.estimate(corner, surf, np.eye(4), T0[:3, 3], q0, p1)
    x1 = np.concatenate([P1, orc.so3_log(q1)])
    # same minimum: costs agree to 1e-6 relative, poses to well below the 1e-4 parity tolerance
    assert cost(x1) <= cost(sol.x) * (1 + 1e-5)
    assert np.abs(x1 - sol.x).max() < 5e-5


# ---------------------------------------------------------------- golden fixtures
def test_golden_fixtures(orc):
    """tests/golden/*.npz were produced by tests/golden/make_golden.py from this oracle on seeded
    inputs; they freeze today's behaviour so later oracle edits cannot drift silently."""
    meta = json.load(open(os.path.join(GOLDEN, "golden.json")))
    g = np.load(os.path.join(GOLDEN, "golden.npz"))
    assert np.array_equal(orc.extract_scan(g["scan_xyzi"], g["scan_line"], int(meta["n_lines"])), g["label"])
    assert np.array_equal(orc.voxel_downsample(g["scan_xyzi"][g["label"] == 2], 0.2), g["surf_ds"])
    m = orc.Map()
    m.set(orc.SURF_LOCAL, g["map_surf"])
    m.set(orc.CORNER_LOCAL, g["map_corner"])
    pf, nf, M, nn = m.associate_plane(g["surf_ds"], g["T_wl"], 10.0)
    assert nf == int(meta["n_plane"]) and np.array_equal(pf[:, 10], g["plane_valid"])
    P, q, st = m.estimate(g["corner_ds"], g["surf_ds"], np.eye(4), g["P0"], g["q0"])
    assert np.abs(P - g["P_est"]).max() < 1e-9 and np.abs(q - g["q_est"]).max() < 1e-9


def test_message_oracle_round_trip(synth):
    """oracle/msgs.py: CustomPoint records serialise to 19 bytes, the filter keeps message order, the near / far cuts
    are float32 squared ranges."""
    from oracle import msgs
    T = synth.make_T(synth.rot_z(0.1), np.array([-3.0, -1.0, 0.2]))
    hx, hl, hs = synth.horizon_scan(T, 2400, seed=5)
    off, xyz, refl, line = synth.horizon_custom_msg(hx, hl, hs)
    raw = msgs.pack_custom_points(off, xyz, refl, line)
    assert raw.size == 19 * len(off)
    x, l, s = msgs.unpack_custom_points(raw, 6)
    keep = (line <= 5) & (xyz[:, 0] >= 0.01)
    assert np.array_equal(x[:, :3], xyz[keep]) and np.array_equal(l, line[keep].astype(np.uint16))
    assert np.all(np.diff(s) >= 0) and abs(float(s[-1]) - 1.0) < 1e-6
    full, corner, surf = msgs.pack_union_clouds(hx, hs, hl, np.where(np.arange(len(hx)) % 7 == 0, 1, 2), 3.0, 9.0, 3.0, 0.0)
    r2 = (full[:, 0] * full[:, 0] + full[:, 1] * full[:, 1]) + full[:, 2] * full[:, 2]
    assert np.all(r2 >= np.float32(9.0)) and np.all(r2 <= np.float32(81.0)) and np.all(full[:, 3] == 1.0)
    assert np.all(corner[:, 6] == 1) and np.all(surf[:, 6] == 2)
