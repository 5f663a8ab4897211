"""The oracle against the REFERENCE TEXT (oracle/_ref: /root/reference's hot-path sources extracted verbatim and
compiled against stand-in Eigen / PCL / Ceres / Sophus / ROS headers, see oracle/ref/).

This is what pins the oracle restatement (and through it the CUDA path) to the reference's own code: control flow,
thresholds, float32/float64 types and operand order are the reference's; only third-party numerics (kd-tree, voxel
grid, eigen/QR/SVD, the trust-region loop) are stand-ins shared with or cross-checked by the oracle.
The library is built where /root/reference exists and travels as a built artefact; without it these tests skip.
"""
import numpy as np
import pytest

from oracle import ref

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built (needs /root/reference at build time)")

T_TRUE = None


def _scene(synth):
    return synth.make_T(synth.rot_z(0.3), np.array([-3.0, -1.0, 0.2]))


# ------------------------------------------------------------------------------------------- A1 (+A2, A3, glue)
def _cmp_labels(orc, xyzi, line, n_lines):
    a = orc.extract_scan(xyzi, line, n_lines)
    b = ref.extract_scan(xyzi, line, n_lines)
    assert np.array_equal(a, b), f"{int((a != b).sum())} labels differ from the reference text"
    return a


def test_labels_s1_s2_s3(orc, synth):
    T = _scene(synth)
    x, r, _ = synth.vlp16_scan(T, seed=1001)
    lab = _cmp_labels(orc, x, r, 16)
    assert (lab == 1).sum() > 50 and (lab == 2).sum() > 500
    x, l, _ = synth.horizon_scan(T, 24000, seed=1002)
    _cmp_labels(orc, x, l, 6)
    Ts = synth.trajectory(5)
    for k in range(1, 5):
        xv, rv, _ = synth.vlp16_scan(Ts[k], seed=2000 + k, T_ws_start=Ts[k - 1])
        xh, lh, _ = synth.horizon_scan(Ts[k], 24000, seed=3000 + k, T_ws_start=Ts[k - 1])
        _cmp_labels(orc, np.concatenate([xv, xh]), np.concatenate([rv, lh + 16]), 22)


def test_labels_s2_240k(orc, synth):
    """BASELINE config 2 size: 6 lines x 40 000 points (beyond the reference's 20 000-entry arrays: edit 1 of
    oracle/ref/extract.sh is what makes the reference text defined here)."""
    x, l, _ = synth.horizon_scan(_scene(synth), 240000, seed=1002)
    lab = _cmp_labels(orc, x, l, 6)
    assert (lab == 2).sum() > 10000


@pytest.mark.parametrize("seed", range(8))
def test_labels_random_scenes(orc, synth, seed):
    rng = np.random.default_rng(100 + seed)
    T = synth.make_T(synth.rot_z(rng.uniform(-3, 3)), np.array([rng.uniform(-6, 6), rng.uniform(-4, 4), rng.uniform(-0.5, 1.5)]))
    noise = [0.0, 0.005, 0.01, 0.03][seed % 4]
    x, r, _ = synth.vlp16_scan(T, seed=500 + seed, noise=noise, n_az=[1800, 900, 450][seed % 3])
    _cmp_labels(orc, x, r, 16)
    x, l, _ = synth.horizon_scan(T, [24000, 9000, 48000][seed % 3], seed=600 + seed, noise=noise)
    _cmp_labels(orc, x, l, 6)


def test_labels_short_and_degenerate_lines(orc):
    rng = np.random.default_rng(3)
    for n in (0, 1, 5, 10, 11, 12, 17, 60, 61, 111):
        x = np.zeros((n, 4), np.float32)
        x[:, 0] = 5.0 + rng.normal(0, 0.01, n)
        x[:, 1] = np.linspace(-1, 1, n)
        x[:, 3] = rng.integers(0, 255, n)
        s1, f1 = orc.detect_feature_points(x)
        s2, f2 = ref.detect_feature_points(x)
        assert np.array_equal(s1, s2) and np.array_equal(f1, f2), n


def test_velodyne_pipeline_verbatim(orc, synth):
    """FE.cpp:1135-1240 as one piece (ring, relative time, split, detector, label write-back) against the oracle's
    A2 + glue + A1."""
    x, _, _ = synth.vlp16_scan(_scene(synth), seed=1001)
    c = ref.velo_extract(x)
    ring, rt = orc.velo_ring_time(x)
    keep = ring >= 0
    assert c.shape[0] == int(keep.sum())
    assert np.array_equal(c[:, 5].astype(np.int16), ring[keep])
    assert np.array_equal(c[:, 4], rt[keep])
    lab = orc.extract_scan(x[keep], ring[keep].astype(np.uint16), 16)
    assert np.array_equal(lab, c[:, 6].astype(np.uint8))


def test_horizon_pipeline_verbatim(orc, synth):
    """getHoriFeatureExtract (FE.cpp:952-1035, 6 detector threads) against the oracle's A3 + glue + A1."""
    xh, lh, sh = synth.horizon_scan(_scene(synth), 24000, seed=1002)
    off, xyz, refl, line = synth.horizon_custom_msg(xh, lh, sh)
    line = line.copy()
    line[::97] = 7          # lines the filter rejects
    xyz = xyz.copy()
    xyz[5::131, 0] = 0.005  # x < 0.01
    c = ref.hori_extract(off, xyz, refl, line)
    keep, rt = orc.hori_filter(off, xyz, line)
    k = keep.astype(bool)
    assert c.shape[0] == int(k.sum())
    assert np.array_equal(c[:, 4], rt[k])
    xi = np.concatenate([xyz, refl[:, None].astype(np.float32)], 1)[k]
    lab = orc.extract_scan(xi, line[k].astype(np.uint16), 6)
    assert np.array_equal(lab, c[:, 6].astype(np.uint8))


# ------------------------------------------------------------------------------------------- A4, A5
def test_undistort(orc, synth):
    x, _, _ = synth.vlp16_scan(_scene(synth), seed=7)
    s = np.random.default_rng(1).random(len(x)).astype(np.float32)
    s[:10] = [0, 1, 0.5, 1e-9, 1 - 1e-7, 0.25, 0.75, 0.1, 0.9, 0.999]
    for rv, dt in (([0.01, -0.02, 0.03], [0.05, 0.01, -0.02]), ([0, 0, 0], [0.1, 0, 0]), ([0.3, 0.2, -0.4], [1.0, -2.0, 0.5])):
        dR = synth.rotvec_to_R(np.array(rv, float))
        a = ref.undistort(x, s, dR, np.array(dt))
        b = orc.undistort(x, s, dR, np.array(dt))
        assert np.array_equal(a, b)


def test_point_to_map_and_cube_rule(orc, synth):
    E = ref.Estimator()
    try:
        rng = np.random.default_rng(0)
        T = synth.make_T(synth.rotvec_to_R(np.array([0.1, -0.2, 0.7])), np.array([3.0, -40.0, 1.0]))
        P = rng.normal(0, 200, (3000, 3)).astype(np.float32)
        P[:50] = np.round(P[:50] / 25.0) * 25.0 + rng.choice([-1e-4, 0, 1e-4], (50, 3))  # on and next to cube faces
        for p in P:
            assert np.array_equal(ref.point_to_map(p, T), orc.point_to_map(p, T))
            assert E.cube_index(p, 0) == orc.cube_index(p, (10, 5, 10))
            assert E.cube_index(p, 1) == orc.cube_index(p, (10, 5, 10))
    finally:
        E.close()


# ------------------------------------------------------------------------------------------- A7 - A12
@pytest.fixture(scope="module")
def matched(orc, synth):
    """Reference Estimator with a global map built by its own MapIncrement, the same map in the oracle, and one
    labelled VLP-16 scan with an S1 offset start pose."""
    E = ref.Estimator(0.4, 0.2)
    ms, mc = synth.feature_map(100_000, 5_000, seed=12)
    E.map_increment(mc, ms)
    E.map_increment(mc[:300] + np.float32(0.001), ms[:300] + np.float32(0.001))  # second pass publishes the cubes
    gc, cen = E.global_map(0)
    gs, _ = E.global_map(1)
    om = orc.Map()
    om.set(orc.CORNER_GLOBAL, gc, cen)
    om.set(orc.SURF_GLOBAL, gs, cen)
    T_true = _scene(synth)
    yield dict(E=E, om=om, cen=cen, T_true=T_true, T_init=T_true @ synth.s1_offset_pose())
    E.close()


def _queries(orc, synth, T_true, seed):
    x, ring, _ = synth.vlp16_scan(T_true, seed=seed)
    lab = orc.extract_scan(x, ring, 16)
    return orc.voxel_downsample(x[lab == 1], 0.4), orc.voxel_downsample(x[lab == 2], 0.2), x, ring


def test_map_increment_moves_the_cube_centre(matched):
    # MapMove's loops (MM.cpp:491-579) leave CenHeight at 2 for a sensor near z = 0: the cube rule must take the
    # centre as a parameter, not as the constant (10, 5, 10)
    assert matched["cen"] == (10, 2, 10)


@pytest.mark.parametrize("thres", [25.0, 10.0, 1.0])
def test_association_features(orc, synth, matched, thres):
    E, om, T = matched["E"], matched["om"], matched["T_init"]
    corner, surf, _, _ = _queries(orc, synth, matched["T_true"], 11)
    fr = E.associate_line(corner, T, thres)
    fo, _ = om.associate_line(corner, T, thres)
    fo = fo[fo[:, 10] >= 0]
    assert len(fr) == len(fo) > 20
    assert np.array_equal(fr[:, :3], fo[:, :3])                 # same accepted queries, same order
    same = np.abs(fr[:, 3:9] - fo[:, 3:9]).max(1) < 1e-6          # eigenvector sign is arbitrary: a/b may swap
    swap = np.abs(fr[:, 3:9] - fo[:, [6, 7, 8, 3, 4, 5]]).max(1) < 1e-6
    assert (same | swap).all()
    assert np.allclose(fr[:, 9], fo[:, 9], rtol=0, atol=1e-9)
    pr, deg, fail = E.associate_plane(surf, T, thres)
    po, _, M, nn = om.associate_plane(surf, T, thres)
    po = po[po[:, 10] >= 0]
    assert len(pr) == len(po) > 300
    assert np.array_equal(pr[:, :3], po[:, :3])
    assert np.array_equal(pr[:, 3:6], po[:, 3:6])               # p_proj bit-equal
    n_ref = pr[:, 6:9] * 1.5e-3                                  # first row of sqrt_info = n / |n| / lidar_m
    n_orc = po[:, 6:9] / np.linalg.norm(po[:, 6:9], axis=1, keepdims=True)
    assert np.abs(n_ref - n_orc).max() < 1e-12
    assert np.abs(pr[:, 9:15]).max() == 0.0                      # plan_weight_tan = 0: tangential rows vanish
    assert np.allclose(pr[:, 15], po[:, 9], rtol=0, atol=1e-12)
    assert not deg and not fail
    assert abs(E.localizability(po[:, 6:9]) - orc.localizability(M, nn)) < 1e-9


def test_cost_functors_autodiff_vs_analytic(orc, synth, matched):
    """The reference's functors (CF.h:412-440, 533-555) under dual-number autodiff against the oracle's closed
    forms, with a non-trivial extrinsic, window-5 tangential weight included."""
    E, T = matched["E"], matched["T_init"]
    corner, surf, _, _ = _queries(orc, synth, matched["T_true"], 12)
    ex = np.eye(4)
    ex[:3, :3] = synth.rotvec_to_R(np.array([0.02, -0.03, 0.05]))
    ex[:3, 3] = [0.05, -0.02, 0.03]
    Tbl = np.linalg.inv(ex)
    rng = np.random.default_rng(0)
    fr = E.associate_line(corner, T, 1.0)
    for f in fr:
        x6 = np.concatenate([T[:3, 3] + rng.normal(0, 0.05, 3), synth.R_to_rotvec(T[:3, :3]) + rng.normal(0, 0.01, 3)])
        r1, J1 = ref.residual(0, f[:9], x6, Tbl)
        f12 = np.zeros(12)
        f12[:9] = f[:9]
        r2, J2 = orc.residual(0, f12, x6, Tbl)
        assert np.allclose(r1, r2, rtol=1e-11, atol=1e-9) and np.allclose(J1, J2, rtol=1e-10, atol=1e-7)
    for wt in (0.0, 0.0003):
        pr, _, _ = E.associate_plane(surf, T, 1.0, plan_weight_tan=wt)
        for f in pr[::5]:
            x6 = np.concatenate([T[:3, 3] + rng.normal(0, 0.05, 3), synth.R_to_rotvec(T[:3, :3]) + rng.normal(0, 0.01, 3)])
            r1, J1 = ref.residual(1, f[:15], x6, Tbl)
            f12 = np.zeros(12)
            f12[:6] = f[:6]
            f12[6:9] = f[6:9] * 1.5e-3
            r2, J2 = orc.residual(1, f12, x6, Tbl, plan_weight_tan=wt)
            # the basis of the tangent plane is arbitrary: compare what the solver sees
            assert abs(r1 @ r1 - r2 @ r2) <= 1e-11 * max(1.0, r1 @ r1)
            assert np.allclose(J1.T @ r1, J2.T @ r2, rtol=1e-10, atol=1e-6)
            assert np.allclose(J1.T @ J1, J2.T @ J2, rtol=1e-10, atol=1e-4)


@pytest.mark.parametrize("seed", [11, 12, 13, 14])
def test_estimate_lidar_pose_window1(orc, synth, matched, seed):
    """Estimator::EstimateLidarPose (EST.cpp:967-1141) end to end on the reference text — label split, VoxelGrid,
    association, Ceres problem with Huber loss, convergence test — against the oracle's orc_estimate."""
    E, om = matched["E"], matched["om"]
    T_true = matched["T_true"] @ synth.make_T(synth.rot_z(0.01 * (seed - 11)), np.array([0.02 * (seed - 11), 0, 0]))
    T_init = T_true @ synth.s1_offset_pose()
    x, _, _ = synth.vlp16_scan(T_true, seed=seed)
    c7 = ref.velo_extract(x)
    lab = c7[:, 6].astype(int)
    corner = orc.voxel_downsample(c7[lab == 1][:, :4], 0.4)
    surf = orc.voxel_downsample(c7[lab == 2][:, :4], 0.2)
    q0, _ = orc.so3_exp(synth.R_to_rotvec(T_init[:3, :3]))
    # the oracle sees the local map the reference holds at this moment
    om.set(orc.CORNER_LOCAL, E.local_map(0))
    om.set(orc.SURF_LOCAL, E.local_map(1))
    P1, q1, fail = E.estimate_lidar_pose(c7, T_init[:3, 3], q0, np.eye(4), 2)
    Po, qo, st = om.estimate(corner, surf, np.eye(4), T_init[:3, 3], q0)
    assert not fail
    assert np.abs(P1 - Po).max() < 1e-7 and 2 * np.abs(q1 - qo).max() < 1e-7
    assert np.abs(Po - T_true[:3, 3]).max() < 0.05


# ------------------------------------------------------------------------------------------- F1
def test_map_increment_local_matches_oracle(orc, synth):
    from oracle import map_maintenance as mmo

    E = ref.Estimator(0.4, 0.2)
    try:
        lm = mmo.LocalMap(0.4, 0.2)
        Ts = synth.trajectory(8, v=4.0)
        for k in range(1, 8):
            x, ring, _ = synth.vlp16_scan(Ts[k], seed=40 + k)
            lab = orc.extract_scan(x, ring, 16)
            corner = orc.voxel_downsample(x[lab == 1], 0.4)
            surf = orc.voxel_downsample(x[lab == 2], 0.2)
            E.map_increment_local(corner, surf, Ts[k])
            c, s = lm.increment(corner, surf, Ts[k])
            assert np.array_equal(E.local_map(0), c)
            assert np.array_equal(E.local_map(1), s)
    finally:
        E.close()


def test_map_increment_matches_oracle(orc, synth):
    """MAP_MANAGER::MapIncrement incl. MapMove (MM.cpp:125-581) against oracle/map_maintenance.CubeMap."""
    from oracle import map_maintenance as mmo

    E = ref.Estimator(0.4, 0.2)
    try:
        cm = mmo.CubeMap()
        so, co = synth.tiled_feature_map(60_000, 6_000, tiles=(2, 2, 1), seed=5)
        rng = np.random.default_rng(2)
        for k in range(4):
            sel_s = rng.choice(len(so), 20_000, replace=False)
            sel_c = rng.choice(len(co), 2_000, replace=False)
            T = synth.make_T(np.eye(3), np.array([110.0 * k, -160.0 * (k % 2), 0.3 + 30.0 * (k == 3)]))  # crosses MapMove bounds
            E.map_increment(co[sel_c], so[sel_s], T)
            cm.increment(co[sel_c], so[sel_s], T)
            gc, cen = E.global_map(2)
            gs, _ = E.global_map(3)
            assert cen == cm.cen
            assert np.array_equal(gc, cm.cloud(0))
            assert np.array_equal(gs, cm.cloud(1))
    finally:
        E.close()


# ------------------------------------------------------------------------------------------- sliding window (IMU)
def test_imu_preintegration_and_factor(orc, synth):
    """IMUIntegrator::PreIntegration (IMU.cpp:105-166) and Cost_NavState_PRV_Bias under autodiff (CF.h:321-393) with
    the reference's sqrt_information (EST.cpp:1238-1242) against the oracle's restatement."""
    imu, stamps = synth.imu_stream(4)
    Ts = synth.trajectory(4)
    rng = np.random.default_rng(0)
    grav = np.array([0, 0, -9.805])

    def x6(T):
        return np.concatenate([T[:3, 3], synth.R_to_rotvec(T[:3, :3])])

    for k in (1, 2, 3):
        bg = rng.normal(0, 2e-3, 3)
        ba = rng.normal(0, 2e-2, 3)
        t, gy, ac = imu[k]
        dq, dp, dv, dt, cov, jac = ref.imu_preintegrate(t, gy, ac, stamps[k - 1], bg, ba)
        P = orc.Preint(t, gy, ac, stamps[k - 1], bg, ba)
        assert np.abs(dq - P.dq).max() < 1e-14 and np.abs(dp - P.dp).max() < 1e-14 and np.abs(dv - P.dv).max() < 1e-13
        assert dt == P.dt
        assert np.abs(cov - P.cov).max() <= 1e-12 * np.abs(cov).max() and np.abs(jac - P.jac).max() < 1e-13
        for _ in range(3):
            pri = x6(Ts[k - 1]) + rng.normal(0, 0.01, 6)
            prj = x6(Ts[k]) + rng.normal(0, 0.01, 6)
            vbi = np.concatenate([synth.body_velocity_world(Ts[k - 1]) + rng.normal(0, 0.02, 3), bg + rng.normal(0, 1e-3, 3),
                                  ba + rng.normal(0, 1e-2, 3)])
            vbj = np.concatenate([synth.body_velocity_world(Ts[k]) + rng.normal(0, 0.02, 3), bg + rng.normal(0, 1e-3, 3),
                                  ba + rng.normal(0, 1e-2, 3)])
            r1, J1 = ref.imu_factor(t, gy, ac, stamps[k - 1], bg, ba, grav, pri, vbi, prj, vbj)
            r2, J2 = orc.imu_factor(P, grav, pri, vbi, prj, vbj)
            assert np.abs(r1 - r2).max() <= 1e-10 * np.abs(r1).max()
            assert np.abs(J1 - J2).max() <= 1e-10 * np.abs(J1).max()


@pytest.mark.parametrize("W", [2, 3, 4])
def test_estimate_window_with_imu(orc, synth, matched, W):
    """Estimator::EstimateLidarPose on a list of W frames with IMU factors (EST.cpp:967-1141 -> Estimate 1143-1581,
    IMU blocks 1235-1254), reference text against orc_estimate_window. BASELINE config 3 is W = 3."""
    E, om = matched["E"], matched["om"]
    imu, stamps = synth.imu_stream(8)
    Ts = synth.trajectory(8)
    rng = np.random.default_rng(W)
    base = 1
    clouds, corners, surfs, pre = [], [], [], [None]
    states = np.zeros((W, 16))
    for f in range(W):
        k = base + f
        x, _, _ = synth.vlp16_scan(Ts[k], seed=50 + k)
        c7 = ref.velo_extract(x)
        clouds.append(c7)
        lab = c7[:, 6].astype(int)
        corners.append(orc.voxel_downsample(c7[lab == 1][:, :4], 0.4))
        surfs.append(orc.voxel_downsample(c7[lab == 2][:, :4], 0.2))
        Tn = Ts[k] @ synth.make_T(synth.rot_z(rng.normal(0, 0.004)), rng.normal(0, 0.03, 3))
        q, _ = orc.so3_exp(synth.R_to_rotvec(Tn[:3, :3]))
        states[f, :3] = Tn[:3, 3]
        states[f, 3:7] = q
        states[f, 7:10] = synth.body_velocity_world(Ts[k]) + rng.normal(0, 0.02, 3)
        if f >= 1:
            pre.append(orc.Preint(*imu[k], stamps[k - 1], states[f - 1, 10:13], states[f - 1, 13:16]))
    om.set(orc.CORNER_LOCAL, E.local_map(0))
    om.set(orc.SURF_LOCAL, E.local_map(1))
    s1, fail = E.estimate_window(clouds, states, stamps[base:base + W], [imu[base + f] for f in range(W)])
    s2, st = orc.estimate_window(om, corners, surfs, np.eye(4), states, pre)
    assert not fail and st[0] >= 1
    assert np.abs(s1[:, :3] - s2[:, :3]).max() < 1e-7          # positions
    assert 2 * np.abs(s1[:, 3:7] - s2[:, 3:7]).max() < 1e-7    # rotations
    assert np.abs(s1[:, 7:10] - s2[:, 7:10]).max() < 1e-6      # velocities
    assert np.abs(s1[:, 10:] - s2[:, 10:]).max() < 1e-5        # biases
    for f in range(W):
        assert np.abs(s2[f, :3] - Ts[base + f][:3, 3]).max() < 0.05
