"""Seeded synthetic inputs for the mm-loam hot path (SURVEY.md §8 d, configs S1-S5).

Pure numpy; no reference code, no oracle. The scene is an axis-aligned box room with four
square pillars, ray-cast by a VLP-16 model (16 rings, 1800 azimuth steps, azimuth-major
point order like the real driver) and a Livox-Horizon-like model (6 interleaved lines, each
tracing a Lissajous figure so that consecutive points of a line are spatial neighbours, as
`detectFeaturePoints` assumes, unionFeatureExtract.cpp:407-451).
"""
from __future__ import annotations

import numpy as np

ROOM = np.array([[-10.0, 10.0], [-7.0, 7.0], [-1.5, 4.5]])  # 20 x 14 x 6 m
PILLARS = [(-4.3, -3.1), (3.7, -2.3), (-3.4, 3.2), (4.6, 2.9)]  # centres, 0.6 m square
PILLAR_HALF = 0.3
HALL = np.array([[-24.0, 24.0], [-24.0, 24.0], [-4.0, 16.0]])  # S4: 48 x 48 x 20 m


def rot_z(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])


def rotvec_to_R(phi):
    phi = np.asarray(phi, dtype=np.float64)
    th = np.linalg.norm(phi)
    K = np.array([[0, -phi[2], phi[1]], [phi[2], 0, -phi[0]], [-phi[1], phi[0], 0]])
    if th < 1e-12:
        return np.eye(3) + K
    return np.eye(3) + np.sin(th) / th * K + (1 - np.cos(th)) / th**2 * K @ K


def R_to_rotvec(R):
    c = np.clip((np.trace(R) - 1) / 2, -1, 1)
    th = np.arccos(c)
    if th < 1e-12:
        return np.zeros(3)
    w = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]]) / (2 * np.sin(th))
    return w * th


def make_T(R, t):
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = t
    return T


def raycast(origins, dirs, room=ROOM, pillars=PILLARS):
    """Distance and surface id of the first hit of each ray (origin inside the room)."""
    o = np.asarray(origins, dtype=np.float64)
    d = np.asarray(dirs, dtype=np.float64)
    n = d.shape[0]
    if o.ndim == 1:
        o = np.broadcast_to(o, (n, 3))
    with np.errstate(divide="ignore", invalid="ignore"):
        t_ax = np.where(d > 0, (room[:, 1] - o) / d, np.where(d < 0, (room[:, 0] - o) / d, np.inf))
    ax = np.argmin(t_ax, axis=1)
    t = t_ax[np.arange(n), ax]
    surf = 2 * ax + (d[np.arange(n), ax] > 0)
    for k, (cx, cy) in enumerate(pillars):
        lo = np.array([cx - PILLAR_HALF, cy - PILLAR_HALF])
        hi = np.array([cx + PILLAR_HALF, cy + PILLAR_HALF])
        with np.errstate(divide="ignore", invalid="ignore"):
            t0 = (lo - o[:, :2]) / d[:, :2]
            t1 = (hi - o[:, :2]) / d[:, :2]
        tn = np.minimum(t0, t1)
        tf = np.maximum(t0, t1)
        tn = np.where(np.isnan(tn), -np.inf, tn)
        tf = np.where(np.isnan(tf), np.inf, tf)
        face = np.argmax(tn, axis=1)
        te = np.max(tn, axis=1)
        tx = np.min(tf, axis=1)
        hit = (te < tx) & (te > 0) & (te < t)
        t = np.where(hit, te, t)
        surf = np.where(hit, 6 + 4 * k + 2 * face + (d[np.arange(n), face] > 0), surf)
    return t, surf


def _finish(ranges, surf, dirs_sensor, rng, noise, intens_seed):
    irng = np.random.default_rng(intens_seed)
    table = irng.uniform(0.0, 255.0, size=64).astype(np.float32)
    r = ranges + (rng.normal(0.0, noise, size=ranges.shape) if noise > 0 else 0.0)
    pts = dirs_sensor * r[:, None]
    out = np.empty((pts.shape[0], 4), dtype=np.float32)
    out[:, :3] = pts.astype(np.float32)
    out[:, 3] = table[surf]
    return out


def _poses_along(T0, T1, s):
    """Per-ray sensor pose: rotation slerp + linear translation between T0 (s=0) and T1 (s=1)."""
    R0, R1 = T0[:3, :3], T1[:3, :3]
    phi = R_to_rotvec(R0.T @ R1)
    th = np.linalg.norm(phi)
    n = s.shape[0]
    if th < 1e-12:
        Rs = np.broadcast_to(R0, (n, 3, 3))
    else:
        k = phi / th
        K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
        a = (s * th)[:, None, None]
        Rs = R0 @ (np.eye(3) + np.sin(a) * K + (1 - np.cos(a)) * (K @ K))
    ts = T0[:3, 3] + s[:, None] * (T1[:3, 3] - T0[:3, 3])
    return Rs, ts


def vlp16_scan(T_ws, seed=1001, noise=0.01, T_ws_start=None, n_az=1800, intens_seed=7):
    """VLP-16: 16 rings (-15..+15 deg, 2 deg step) x n_az azimuth steps, azimuth-major order.

    Returns (xyzi float32 [n,4], ring uint16 [n], s float32 [n]) with s the true sweep
    fraction of each point. If T_ws_start is given the sensor moves from it (s=0) to T_ws
    (s=1) during the sweep (motion distortion); points stay in the instantaneous frame.
    """
    rng = np.random.default_rng(seed)
    az = (np.arange(n_az) + 0.37) * (2 * np.pi / n_az)  # clockwise, never exactly on +-pi
    el = np.deg2rad(-15.0 + 2.0 * np.arange(16))
    A, E = np.meshgrid(az, el, indexing="ij")  # azimuth-major
    A, E = A.ravel(), E.ravel()
    ring = np.tile(np.arange(16, dtype=np.uint16), n_az)
    d = np.stack([np.cos(E) * np.cos(A), -np.cos(E) * np.sin(A), np.sin(E)], axis=1)
    s = (np.repeat(np.arange(n_az), 16) / n_az).astype(np.float64)
    if T_ws_start is None:
        Rs, ts = np.broadcast_to(T_ws[:3, :3], (d.shape[0], 3, 3)), np.broadcast_to(T_ws[:3, 3], (d.shape[0], 3))
    else:
        Rs, ts = _poses_along(T_ws_start, T_ws, s)
    dw = np.einsum("nij,nj->ni", Rs, d)
    t, surf = raycast(ts, dw)
    return _finish(t, surf, d, rng, noise, intens_seed), ring, s.astype(np.float32)


def horizon_scan(T_ws, n_points=24000, seed=1002, noise=0.01, T_ws_start=None, intens_seed=7):
    """Livox-Horizon-like frame: 6 interleaved lines, FoV 81.7 x 25.1 deg, Lissajous sweep.

    Returns (xyzi float32 [n,4], line uint16 [n], s float32 [n]); point k belongs to line k%6.
    """
    rng = np.random.default_rng(seed)
    n_line = n_points // 6
    k = np.arange(n_line, dtype=np.float64)
    u = k / n_line
    pts_d, lines, ss = [], [], []
    ph = rng.uniform(0, 2 * np.pi, size=6)
    for l in range(6):
        az = np.deg2rad(81.7 / 2) * np.sin(2 * np.pi * 7 * u + ph[l]) * 0.98
        el = np.deg2rad(25.1 / 2) * (0.8 * np.sin(2 * np.pi * 3 * u + 0.7 * l) + 0.15 * (l - 2.5) / 2.5)
        pts_d.append(np.stack([np.cos(el) * np.cos(az), np.cos(el) * np.sin(az), np.sin(el)], axis=1))
        lines.append(np.full(n_line, l, dtype=np.uint16))
        ss.append((k * 6 + l) / (6 * n_line))
    d = np.stack(pts_d, axis=1).reshape(-1, 3)  # interleave: point index = k*6 + l
    line = np.stack(lines, axis=1).reshape(-1)
    s = np.stack(ss, axis=1).reshape(-1)
    if T_ws_start is None:
        Rs, ts = np.broadcast_to(T_ws[:3, :3], (d.shape[0], 3, 3)), np.broadcast_to(T_ws[:3, 3], (d.shape[0], 3))
    else:
        Rs, ts = _poses_along(T_ws_start, T_ws, s)
    dw = np.einsum("nij,nj->ni", Rs, d)
    t, surf = raycast(ts, dw)
    return _finish(t, surf, d, rng, noise, intens_seed), line, s.astype(np.float32)


def horizon_custom_msg(xyzi, line, s, frame_ns=100_000_000):
    """CustomPoint-style arrays for A3: offset_time u32 (ns), xyz f32[n,3], reflectivity u8, line u8."""
    off = np.round(s.astype(np.float64) * frame_ns).astype(np.uint32)
    off[-1] = max(int(off[-1]), 1)
    return off, np.ascontiguousarray(xyzi[:, :3]), xyzi[:, 3].astype(np.uint8), line.astype(np.uint8)


def _sample_planes(rng, box, pillars, n):
    """Points on the walls/floor/ceiling of `box` and on pillar faces, area-proportional."""
    ext = box[:, 1] - box[:, 0]
    faces = []
    for ax in range(3):
        o = [a for a in range(3) if a != ax]
        area = ext[o[0]] * ext[o[1]]
        for side in range(2):
            faces.append(("box", ax, side, area))
    for k, _ in enumerate(pillars):
        for ax in range(2):
            for side in range(2):
                faces.append(("pil", k, (ax, side), 2 * PILLAR_HALF * ext[2]))
    areas = np.array([f[3] for f in faces])
    cnt = rng.multinomial(n, areas / areas.sum())
    out = []
    for f, c in zip(faces, cnt):
        if c == 0:
            continue
        if f[0] == "box":
            ax, side = f[1], f[2]
            p = rng.uniform(box[:, 0], box[:, 1], size=(c, 3))
            p[:, ax] = box[ax, side]
        else:
            cx, cy = pillars[f[1]]
            ax, side = f[2]
            lo = np.array([cx - PILLAR_HALF, cy - PILLAR_HALF, box[2, 0]])
            hi = np.array([cx + PILLAR_HALF, cy + PILLAR_HALF, box[2, 1]])
            p = rng.uniform(lo, hi, size=(c, 3))
            p[:, ax] = (lo if side == 0 else hi)[ax]
        out.append(p)
    return np.concatenate(out, axis=0)


def _sample_edges(rng, box, pillars, n):
    """Points on the 12 box edges and the 4 vertical edges of every pillar, length-proportional."""
    segs = []
    b = box
    for ax in range(3):
        o = [a for a in range(3) if a != ax]
        for s0 in range(2):
            for s1 in range(2):
                a0 = np.zeros(3)
                a1 = np.zeros(3)
                a0[ax], a1[ax] = b[ax, 0], b[ax, 1]
                a0[o[0]] = a1[o[0]] = b[o[0], s0]
                a0[o[1]] = a1[o[1]] = b[o[1], s1]
                segs.append((a0, a1))
    for cx, cy in pillars:
        for sx in (-1, 1):
            for sy in (-1, 1):
                x, y = cx + sx * PILLAR_HALF, cy + sy * PILLAR_HALF
                segs.append((np.array([x, y, b[2, 0]]), np.array([x, y, b[2, 1]])))
    L = np.array([np.linalg.norm(s[1] - s[0]) for s in segs])
    cnt = rng.multinomial(n, L / L.sum())
    out = []
    for (a0, a1), c in zip(segs, cnt):
        if c:
            u = rng.uniform(0, 1, size=(c, 1))
            out.append(a0 + u * (a1 - a0))
    return np.concatenate(out, axis=0)


def feature_map(n_surf, n_corner, seed=1002, box=ROOM, pillars=PILLARS, noise=0.005):
    """World-frame feature map: (surf xyzi float32 [n_surf,4], corner xyzi float32 [n_corner,4])."""
    rng = np.random.default_rng(seed)
    s = _sample_planes(rng, box, pillars, n_surf) + rng.normal(0, noise, size=(n_surf, 3))
    c = _sample_edges(rng, box, pillars, n_corner) + rng.normal(0, noise, size=(n_corner, 3))
    so = np.zeros((n_surf, 4), np.float32)
    co = np.zeros((n_corner, 4), np.float32)
    so[:, :3] = s
    co[:, :3] = c
    return so, co


def tiled_feature_map(n_surf, n_corner, tiles=(4, 4, 1), seed=1005, noise=0.005):
    """S5: one 48 m hall per 50 m cube over tiles[0] x tiles[1] x tiles[2] cubes."""
    rng = np.random.default_rng(seed)
    nt = tiles[0] * tiles[1] * tiles[2]
    hall = np.array([[-24.0, 24.0], [-24.0, 24.0], [-10.0, 10.0]])
    surf, corner = [], []
    for ix in range(tiles[0]):
        for iy in range(tiles[1]):
            for iz in range(tiles[2]):
                off = np.array([50.0 * ix, 50.0 * iy, 50.0 * iz])
                s = _sample_planes(rng, hall, [], n_surf // nt) + off
                c = _sample_edges(rng, hall, [], n_corner // nt) + off
                surf.append(s + rng.normal(0, noise, size=s.shape))
                corner.append(c + rng.normal(0, noise, size=c.shape))
    s = np.concatenate(surf)
    c = np.concatenate(corner)
    so = np.zeros((s.shape[0], 4), np.float32)
    co = np.zeros((c.shape[0], 4), np.float32)
    so[:, :3] = s
    co[:, :3] = c
    return so, co


def queries_from_map(map_xyzi, n_q, T_wl, seed=1004, noise=0.01):
    """S4: queries = map points moved into the LiDAR frame of pose T_wl, plus noise."""
    rng = np.random.default_rng(seed)
    idx = rng.integers(0, map_xyzi.shape[0], size=n_q)
    pw = map_xyzi[idx, :3].astype(np.float64) + rng.normal(0, noise, size=(n_q, 3))
    pl = (pw - T_wl[:3, 3]) @ T_wl[:3, :3]  # R^T (p - t)
    q = np.zeros((n_q, 4), np.float32)
    q[:, :3] = pl
    return q


def s1_offset_pose():
    """SURVEY S1 target pose: (0.10, 0.05, 0.02) m, yaw 1 deg."""
    return make_T(rot_z(np.deg2rad(1.0)), np.array([0.10, 0.05, 0.02]))


def trajectory(n_frames, v=0.5, yaw_rate=0.2, dt=0.1, start=(-3.0, -1.0, 0.0)):
    """S3: constant-twist planar trajectory; returns list of T_ws (4x4) at frame ends."""
    Ts = []
    x, y, z = start
    yaw = 0.0
    for _ in range(n_frames + 1):
        Ts.append(make_T(rot_z(yaw), np.array([x, y, z])))
        # integrate unicycle exactly over dt
        if abs(yaw_rate) < 1e-12:
            x += v * dt * np.cos(yaw)
            y += v * dt * np.sin(yaw)
        else:
            x += v / yaw_rate * (np.sin(yaw + yaw_rate * dt) - np.sin(yaw))
            y += -v / yaw_rate * (np.cos(yaw + yaw_rate * dt) - np.cos(yaw))
        yaw += yaw_rate * dt
    return Ts


def imu_stream(n_frames, v=0.5, yaw_rate=0.2, dt=0.1, rate_hz=200.0, seed=1003, gyr_sigma=0.004, acc_sigma=0.08,
               gnorm=9.805):
    """S3: 200 Hz IMU samples for the constant-twist `trajectory` (body frame = IMU = LiDAR frame when the extrinsic
    is the identity). Returns per frame k = 1..n_frames a tuple (t [m], gyr [m,3] rad/s, acc [m,3] in units of g) holding
    the samples of (t_{k-1}, t_k], plus the frame stamps t_k (frame 0 at t = 0). Specific force of the unicycle:
    (0, v * yaw_rate, g) in the body frame; white noise with the densities of IMUIntegrator.h:79-82."""
    rng = np.random.default_rng(seed)
    per = int(round(rate_hz * dt))
    stamps = np.arange(n_frames + 1) * dt
    out = [(np.zeros(0), np.zeros((0, 3)), np.zeros((0, 3)))]
    for k in range(1, n_frames + 1):
        t = stamps[k - 1] + (np.arange(per) + 1) * (dt / per)
        gyr = np.tile(np.array([0.0, 0.0, yaw_rate]), (per, 1)) + rng.normal(0, gyr_sigma, (per, 3))
        acc = (np.tile(np.array([0.0, v * yaw_rate, gnorm]), (per, 1)) + rng.normal(0, acc_sigma, (per, 3))) / gnorm
        out.append((t, gyr, acc))
    return out, stamps


def body_velocity_world(T_wb, v=0.5):
    """World-frame velocity of the unicycle at pose T_wb."""
    return T_wb[:3, :3] @ np.array([v, 0.0, 0.0])
