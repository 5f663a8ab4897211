"""Sliding-window rows (window sizes 2-4, BASELINE config 3 = window 3): host-side IMU code of the product against the
oracle (CPU), and the device path against the oracle (GPU)."""
import numpy as np
import pytest


def _x6(synth, T):
    return np.concatenate([T[:3, 3], synth.R_to_rotvec(T[:3, :3])])


def _state(orc, synth, T, v=0.5):
    st = np.zeros(16)
    st[:3] = T[:3, 3]
    st[3:7] = orc.so3_exp(synth.R_to_rotvec(T[:3, :3]))[0]
    st[7:10] = synth.body_velocity_world(T, v)
    return st


def test_host_imu_code_matches_oracle(mm, orc, synth):
    """mml_imu_preintegrate / mml_imu_factor / mml_imu_predict are host code: checked here without a GPU."""
    imu, stamps = synth.imu_stream(4)
    Ts = synth.trajectory(4)
    rng = np.random.default_rng(1)
    for k in (1, 2, 3):
        bg, ba = rng.normal(0, 2e-3, 3), rng.normal(0, 2e-2, 3)
        P = mm.imu_preintegrate(*imu[k], stamps[k - 1], bg, ba)
        O = orc.Preint(*imu[k], stamps[k - 1], bg, ba)
        assert np.abs(np.array(P.dq) - O.dq).max() < 1e-15 and np.abs(np.array(P.dp) - O.dp).max() < 1e-15
        Pm = mm.imu_preintegrate_mean(*imu[k], stamps[k - 1], bg, ba)  # the loop predicts from the mean alone: same bits
        assert list(Pm.dq) == list(P.dq) and list(Pm.dp) == list(P.dp) and list(Pm.dv) == list(P.dv) and Pm.dt == P.dt
        assert np.abs(np.array(P.cov).reshape(15, 15) - O.cov).max() <= 1e-14 * np.abs(O.cov).max()
        assert np.abs(np.array(P.sqrt_info).reshape(15, 15) - O.sqrt_info).max() <= 1e-9 * np.abs(O.sqrt_info).max()
        pri = _x6(synth, Ts[k - 1]) + rng.normal(0, 0.01, 6)
        prj = _x6(synth, Ts[k]) + rng.normal(0, 0.01, 6)
        vbi = np.concatenate([synth.body_velocity_world(Ts[k - 1]), bg + 1e-3, ba - 1e-2])
        vbj = np.concatenate([synth.body_velocity_world(Ts[k]), bg, ba])
        r1, J1 = mm.imu_factor(P, [0, 0, -9.805], pri, vbi, prj, vbj)
        r2, J2 = orc.imu_factor(O, [0, 0, -9.805], pri, vbi, prj, vbj)
        assert np.abs(r1 - r2).max() <= 1e-9 * np.abs(r2).max() and np.abs(J1 - J2).max() <= 1e-9 * np.abs(J2).max()
        s0 = _state(orc, synth, Ts[k - 1])
        assert np.abs(mm.imu_predict(s0, P) - orc.imu_predict(s0, O)).max() < 1e-14


def test_imu_factor_jacobian_finite_differences(mm, synth, orc):
    imu, stamps = synth.imu_stream(2)
    Ts = synth.trajectory(2)
    P = mm.imu_preintegrate(*imu[1], stamps[0])
    rng = np.random.default_rng(3)
    x = np.concatenate([_x6(synth, Ts[0]), synth.body_velocity_world(Ts[0]), np.zeros(6), _x6(synth, Ts[1]),
                        synth.body_velocity_world(Ts[1]), np.zeros(6)]) + rng.normal(0, 1e-3, 30)
    g = [0, 0, -9.805]

    def f(v):
        return mm.imu_factor(P, g, v[:6], v[6:15], v[15:21], v[21:30])[0]

    _, J = mm.imu_factor(P, g, x[:6], x[6:15], x[15:21], x[21:30])
    for c in range(30):
        h = 1e-6
        e = np.zeros(30)
        e[c] = h
        fd = (f(x + e) - f(x - e)) / (2 * h)
        assert np.abs(fd - J[:, c]).max() <= 1e-5 * max(1.0, np.abs(J[:, c]).max())


def _window_case(orc, synth, W, base=1, seed=0):
    imu, stamps = synth.imu_stream(8)
    Ts = synth.trajectory(8)
    rng = np.random.default_rng(seed)
    corners, surfs, states, pre_o = [], [], np.zeros((W, 16)), [None]
    for f in range(W):
        k = base + f
        xv, rv, _ = synth.vlp16_scan(Ts[k], seed=50 + k)
        xh, lh, _ = synth.horizon_scan(Ts[k], 24000, seed=80 + k)
        x = np.concatenate([xv, xh])
        lab = orc.extract_scan(x, np.concatenate([rv, lh + 16]), 22)
        corners.append(orc.voxel_downsample(x[lab == 1], 0.4))
        surfs.append(orc.voxel_downsample(x[lab == 2], 0.2))
        Tn = Ts[k] @ synth.make_T(synth.rot_z(rng.normal(0, 0.004)), rng.normal(0, 0.03, 3))
        states[f] = _state(orc, synth, Tn)
        states[f, 7:10] += rng.normal(0, 0.02, 3)
        if f >= 1:
            pre_o.append(orc.Preint(*imu[k], stamps[k - 1], states[f - 1, 10:13], states[f - 1, 13:16]))
    return imu, stamps, Ts, corners, surfs, states, pre_o


@pytest.mark.gpu
@pytest.mark.parametrize("W", [1, 2, 3, 4])
def test_estimate_window_matches_oracle(ctx, mm, orc, synth, scene, W):
    imu, stamps, Ts, corners, surfs, states, pre_o = _window_case(orc, synth, W, seed=W)
    ctx.map_set(mm.MAP_SURF_LOCAL, scene["map_surf"])
    ctx.map_set(mm.MAP_CORNER_LOCAL, scene["map_corner"])
    ctx.map_set(mm.MAP_SURF_GLOBAL, np.zeros((0, 4), np.float32))
    ctx.map_set(mm.MAP_CORNER_GLOBAL, np.zeros((0, 4), np.float32))
    om = orc.Map()
    om.set(orc.SURF_LOCAL, scene["map_surf"])
    om.set(orc.CORNER_LOCAL, scene["map_corner"])
    ctx.window_reset()
    pre_g = [None]
    for f in range(W):
        ctx.window_push_frame(corners[f], surfs[f], max_frames=W)
        if f >= 1:
            pre_g.append(mm.imu_preintegrate(*imu[1 + f], stamps[f], states[f - 1, 10:13], states[f - 1, 13:16]))
    assert ctx.window_size() == W
    ex = np.eye(4)
    ex[:3, :3] = synth.rotvec_to_R(np.array([0.01, -0.02, 0.015]))
    ex[:3, 3] = [0.04, -0.02, 0.03]
    s_g, st_g = ctx.estimate_window(states, pre_g, ex)
    s_o, st_o = orc.estimate_window(om, corners, surfs, ex, states, pre_o)
    assert st_g[0] == st_o[0] and st_g[2] == st_o[2] and st_g[3] == st_o[3]      # outer iterations, feature counts
    assert np.abs(s_g[:, :3] - s_o[:, :3]).max() < 1e-6                           # tolerance of north_star: 1e-4 m
    assert 2 * np.abs(s_g[:, 3:7] - s_o[:, 3:7]).max() < 1e-6                     # 1e-4 rad
    assert np.abs(s_g[:, 7:10] - s_o[:, 7:10]).max() < 1e-5
    assert np.abs(s_g[:, 10:] - s_o[:, 10:]).max() < 1e-4
    assert abs(st_g[5] - st_o[5]) < 1e-6 and st_g[6] == st_o[6]


@pytest.mark.gpu
@pytest.mark.parametrize("W", [3, 2])
def test_odom_run_window_matches_oracle_loop(ctx, mm, orc, synth, scene, W):
    """BASELINE config 3: merged VLP-16 + Horizon scans with motion distortion, IMU pre-integration, window 3."""
    from oracle import window_loop

    n = 7
    Ts = synth.trajectory(n)
    imu, stamps = synth.imu_stream(n)
    scans = []
    for k in range(1, n + 1):
        xv, rv, sv = synth.vlp16_scan(Ts[k], seed=2000 + k, T_ws_start=Ts[k - 1])
        xh, lh, sh = synth.horizon_scan(Ts[k], 24000, seed=3000 + k, T_ws_start=Ts[k - 1])
        scans.append((np.concatenate([xv, xh]), np.concatenate([rv, lh + 16]).astype(np.uint16), np.concatenate([sv, sh])))
    ctx.map_set(mm.MAP_SURF_LOCAL, scene["map_surf"])
    ctx.map_set(mm.MAP_CORNER_LOCAL, scene["map_corner"])
    ctx.map_set(mm.MAP_SURF_GLOBAL, np.zeros((0, 4), np.float32))
    ctx.map_set(mm.MAP_CORNER_GLOBAL, np.zeros((0, 4), np.float32))
    om = orc.Map()
    om.set(orc.SURF_LOCAL, scene["map_surf"])
    om.set(orc.CORNER_LOCAL, scene["map_corner"])
    state0 = _state(orc, synth, Ts[0])
    res_o = window_loop.run(om, scans, 22, W, stamps[1:], stamps[0], imu[1:], state0)
    host = [(np.ascontiguousarray(x, np.float32), np.ascontiguousarray(l, np.uint16), np.ascontiguousarray(s, np.float32), len(x))
            for x, l, s in scans]
    res_g = ctx.odom_run_window(host, 22, W, stamps[1:], stamps[0], imu[1:], state0, host_buffers=True)
    dpos = np.abs(res_g["poses_newest"][:, :3, 3] - res_o["poses_newest"][:, :3, 3]).max()
    drot = np.abs(res_g["poses_newest"][:, :3, :3] - res_o["poses_newest"][:, :3, :3]).max()
    assert dpos < 1e-4 and drot < 1e-4, (dpos, drot)
    assert np.abs(res_g["poses_front"][:, :3, 3] - res_o["poses_front"][:, :3, 3]).max() < 1e-4
    assert np.array_equal(res_g["stats"][:, 0], res_o["stats"][:, 0])
    err = np.abs(res_g["poses_newest"][:, :3, 3] - np.array([Ts[k][:3, 3] for k in range(1, n + 1)])).max()
    assert err < 0.05


@pytest.mark.gpu
def test_odom_run_window_with_map_updates_matches_oracle_loop(ctx, mm, orc, synth, scene):
    """The loop with the map update of EstimateLidarPose switched on (mml_est_params.map_update): a fast trajectory
    (0.4 m per scan) crosses the sqrt(0.5) m gate every other scan; poses, update count and the final local maps
    (bit for bit) against the oracle loop driving map_maintenance.LocalMap."""
    from oracle import map_maintenance as mmt
    from oracle import window_loop

    n, W, v = 8, 3, 4.0
    Ts = synth.trajectory(n, v=v)
    imu, stamps = synth.imu_stream(n, v=v)
    scans = []
    for k in range(1, n + 1):
        xv, rv, sv = synth.vlp16_scan(Ts[k], seed=4000 + k, T_ws_start=Ts[k - 1])
        xh, lh, sh = synth.horizon_scan(Ts[k], 24000, seed=5000 + k, T_ws_start=Ts[k - 1])
        scans.append((np.concatenate([xv, xh]), np.concatenate([rv, lh + 16]).astype(np.uint16), np.concatenate([sv, sh])))
    ctx.local_map_reset()
    ctx.map_set(mm.MAP_SURF_GLOBAL, np.zeros((0, 4), np.float32))
    ctx.map_set(mm.MAP_CORNER_GLOBAL, np.zeros((0, 4), np.float32))
    ctx.map_set(mm.MAP_SURF_LOCAL, scene["map_surf"])
    ctx.map_set(mm.MAP_CORNER_LOCAL, scene["map_corner"])
    ctx.local_map_seed(0, 49, scene["map_corner"])   # the map "earlier frames" built
    ctx.local_map_seed(1, 49, scene["map_surf"])
    om = orc.Map()
    om.set(orc.SURF_LOCAL, scene["map_surf"])
    om.set(orc.CORNER_LOCAL, scene["map_corner"])
    lm = mmt.LocalMap(0.4, 0.2)
    lm.ring[0][49] = np.ascontiguousarray(scene["map_corner"], np.float32)
    lm.ring[1][49] = np.ascontiguousarray(scene["map_surf"], np.float32)
    state0 = _state(orc, synth, Ts[0], v)
    res_o = window_loop.run(om, scans, 22, W, stamps[1:], stamps[0], imu[1:], state0, local_map=lm)
    assert res_o["map_updates"] >= 3
    host = [(np.ascontiguousarray(x, np.float32), np.ascontiguousarray(l, np.uint16), np.ascontiguousarray(s, np.float32), len(x))
            for x, l, s in scans]
    res_g = ctx.odom_run_window(host, 22, W, stamps[1:], stamps[0], imu[1:], state0, host_buffers=True,
                                params=mm.est_params(map_update=1))
    dpos = np.abs(res_g["poses_newest"][:, :3, 3] - res_o["poses_newest"][:, :3, 3]).max()
    drot = np.abs(res_g["poses_newest"][:, :3, :3] - res_o["poses_newest"][:, :3, :3]).max()
    assert dpos < 1e-4 and drot < 1e-4, (dpos, drot)
    assert np.array_equal(res_g["stats"][:, 0], res_o["stats"][:, 0])
    assert np.array_equal(ctx.local_map_get(0), lm.from_local[0]) and np.array_equal(ctx.local_map_get(1), lm.from_local[1])
    ctx.local_map_reset()

