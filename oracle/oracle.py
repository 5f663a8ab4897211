"""ctypes wrapper of the CPU ORACLE (oracle/oracle.h).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs. The product never imports this module.
Parity unpinned (the reference has no tests or golden vectors and cannot be built here).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libmmloam_oracle.so")
_lib = None


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".cpp", ".h"))]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.orc_map_create.restype = C.c_void_p
        _lib.orc_map_destroy.argtypes = [C.c_void_p]
        _lib.orc_localizability.restype = C.c_double
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class EstParams(C.Structure):
    _fields_ = [("max_outer", C.c_int), ("max_inner", C.c_int), ("lidar_m", C.c_double),
                ("plan_weight_tan", C.c_double), ("thres0", C.c_double), ("thres1", C.c_double),
                ("thres2", C.c_double), ("use_huber", C.c_int), ("threads", C.c_int)]


def est_params(**kw):
    p = EstParams()
    lib().orc_est_params_default(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def detect_feature_points(xyzi):
    xyzi = _f32(xyzi)
    n = xyzi.shape[0]
    sharp = np.zeros(max(n, 1), np.int32)
    flat = np.zeros(max(n, 1), np.int32)
    ns, nf = C.c_int(0), C.c_int(0)
    lib().orc_detect_feature_points(_p(xyzi), n, _p(sharp), C.byref(ns), _p(flat), C.byref(nf))
    return sharp[: ns.value].copy(), flat[: nf.value].copy()


def detect_feature_flags(xyzi):
    xyzi = _f32(xyzi)
    n = xyzi.shape[0]
    flags = np.zeros(max(n, 1), np.int32)
    m = lib().orc_detect_feature_flags(_p(xyzi), n, _p(flags))
    return flags[:m].copy()


def velo_ring_time(xyzi):
    xyzi = _f32(xyzi)
    n = xyzi.shape[0]
    line = np.zeros(n, np.int16)
    rt = np.zeros(n, np.float32)
    lib().orc_velo_ring_time(_p(xyzi), n, _p(line), _p(rt))
    return line, rt


def hori_filter(offset_time, xyz, line):
    offset_time = np.ascontiguousarray(offset_time, np.uint32)
    xyz = _f32(xyz)
    line = np.ascontiguousarray(line, np.uint8)
    n = xyz.shape[0]
    keep = np.zeros(n, np.uint8)
    rt = np.zeros(n, np.float32)
    lib().orc_hori_filter(_p(offset_time), _p(xyz), _p(line), n, _p(keep), _p(rt))
    return keep, rt


def extract_scan(xyzi, line_id, n_lines, threads=1):
    xyzi = _f32(xyzi)
    line_id = np.ascontiguousarray(line_id, np.uint16)
    n = xyzi.shape[0]
    label = np.zeros(max(n, 1), np.uint8)
    lib().orc_extract_scan(_p(xyzi), _p(line_id), n, n_lines, _p(label), threads)
    return label[:n]


def undistort(xyzi, s, dR, dt):
    out = _f32(xyzi).copy()
    s = _f32(s)
    dR = _f64(dR).reshape(9)
    dt = _f64(dt).reshape(3)
    lib().orc_undistort(_p(out), _p(s), out.shape[0], _p(dR), _p(dt))
    return out


def voxel_downsample(xyzi, leaf):
    xyzi = _f32(xyzi)
    n = xyzi.shape[0]
    out = np.zeros((max(n, 1), 4), np.float32)
    m = C.c_int(0)
    lib().orc_voxel_downsample(_p(xyzi), n, C.c_float(leaf), _p(out), C.byref(m))
    return out[: m.value].copy()


def point_to_map(p3, T):
    p3 = _f32(p3)
    T = _f64(T).reshape(16)
    out = np.zeros(3, np.float32)
    lib().orc_point_to_map(_p(p3), _p(T), _p(out))
    return out


def cube_index(p3, cen=(10, 5, 10)):
    p3 = _f32(p3)
    return lib().orc_cube_index(_p(p3), cen[0], cen[1], cen[2])


def knn5_brute(cloud, q3):
    cloud = _f32(cloud)
    q3 = _f32(q3)
    idx = np.zeros(5, np.int32)
    d2 = np.zeros(5, np.float32)
    lib().orc_knn5_brute(_p(cloud), cloud.shape[0], _p(q3), _p(idx), _p(d2))
    return idx, d2


def knn5_kdtree(cloud, q):
    cloud = _f32(cloud)
    q = _f32(q)
    nq = q.shape[0]
    idx = np.zeros((nq, 5), np.int32)
    d2 = np.zeros((nq, 5), np.float32)
    lib().orc_knn5_kdtree(_p(cloud), cloud.shape[0], _p(q), nq, _p(idx), _p(d2))
    return idx, d2


CORNER_GLOBAL, SURF_GLOBAL, CORNER_LOCAL, SURF_LOCAL = 0, 1, 2, 3


class Map:
    def __init__(self):
        self.h = C.c_void_p(lib().orc_map_create())

    def __del__(self):
        try:
            if self.h:
                lib().orc_map_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def set(self, kind, xyzi, cen=(10, 5, 10)):
        xyzi = _f32(xyzi).reshape(-1, 4)
        cen = np.asarray(cen, np.int32)
        return lib().orc_map_set(self.h, kind, _p(xyzi), xyzi.shape[0], _p(cen))

    def associate_line(self, q, T_wl, thres):
        q = _f32(q)
        T = _f64(T_wl).reshape(16)
        feat = np.zeros((max(q.shape[0], 1), 12), np.float64)
        nf = C.c_int(0)
        lib().orc_associate_line(self.h, _p(q), q.shape[0], _p(T), C.c_double(thres), _p(feat), C.byref(nf))
        return feat[: q.shape[0]], nf.value

    def associate_plane(self, q, T_wl, thres):
        q = _f32(q)
        T = _f64(T_wl).reshape(16)
        feat = np.zeros((max(q.shape[0], 1), 12), np.float64)
        nf, nn = C.c_int(0), C.c_int(0)
        M = np.zeros(9, np.float64)
        lib().orc_associate_plane(self.h, _p(q), q.shape[0], _p(T), C.c_double(thres), _p(feat), C.byref(nf), _p(M),
                                  C.byref(nn))
        return feat[: q.shape[0]], nf.value, M.reshape(3, 3), nn.value

    def estimate(self, corner, surf, exTlb, P, q_wxyz, params=None):
        corner = _f32(corner).reshape(-1, 4)
        surf = _f32(surf).reshape(-1, 4)
        ex = _f64(exTlb).reshape(16)
        P = _f64(P).copy()
        q = _f64(q_wxyz).copy()
        stats = np.zeros(16, np.float64)
        prm = params if params is not None else est_params()
        lib().orc_estimate(self.h, _p(corner), corner.shape[0], _p(surf), surf.shape[0], _p(ex), _p(P), _p(q),
                           C.byref(prm), _p(stats))
        return P, q, stats


def localizability(M, n):
    M = _f64(M).reshape(9)
    return lib().orc_localizability(_p(M), int(n))


def accumulate(line_feat, plane_feat, x6, T_bl, plan_weight_tan=0.0, huber_a=0.1 / 1.5e-3, threads=1):
    lf = _f64(line_feat).reshape(-1, 12)
    pf = _f64(plane_feat).reshape(-1, 12)
    x6 = _f64(x6)
    T = _f64(T_bl).reshape(16)
    H = np.zeros(36, np.float64)
    g = np.zeros(6, np.float64)
    cost = C.c_double(0)
    lib().orc_accumulate(_p(lf), lf.shape[0], _p(pf), pf.shape[0], _p(x6), _p(T), C.c_double(plan_weight_tan),
                         C.c_double(huber_a), _p(H), _p(g), C.byref(cost), threads)
    return H.reshape(6, 6), g, cost.value


def residual(kind, feat12, x6, T_bl, plan_weight_tan=0.0, jac=True):
    f = _f64(feat12)
    x6 = _f64(x6)
    T = _f64(T_bl).reshape(16)
    r = np.zeros(3, np.float64)
    J = np.zeros(18, np.float64)
    lib().orc_residual(kind, _p(f), _p(x6), _p(T), C.c_double(plan_weight_tan), _p(r), _p(J) if jac else None)
    n = 1 if kind == 0 else 3
    return r[:n].copy(), J.reshape(3, 6)[:n].copy()


def so3_exp(phi):
    phi = _f64(phi)
    q = np.zeros(4)
    R = np.zeros(9)
    lib().orc_so3_exp(_p(phi), _p(q), _p(R))
    return q, R.reshape(3, 3)


def so3_log(q):
    q = _f64(q)
    phi = np.zeros(3)
    lib().orc_so3_log(_p(q), _p(phi))
    return phi


# ---- sliding window (IMU factors) -------------------------------------------------------------------
class Preint:
    """IMUIntegrator::PreIntegration result (opaque block + the fields tests look at)."""

    def __init__(self, t, gyr, acc, last_time, bg=(0, 0, 0), ba=(0, 0, 0)):
        t = _f64(t)
        gyr = _f64(gyr).reshape(-1, 3)
        acc = _f64(acc).reshape(-1, 3)
        self.buf = np.zeros(lib().orc_preint_size() // 8, np.float64)
        lib().orc_imu_preintegrate(_p(t), _p(gyr), _p(acc), t.shape[0], C.c_double(last_time), _p(_f64(bg)), _p(_f64(ba)),
                                   _p(self.buf))

    dq = property(lambda s: s.buf[0:4])
    dp = property(lambda s: s.buf[4:7])
    dv = property(lambda s: s.buf[7:10])
    dt = property(lambda s: float(s.buf[10]))
    cov = property(lambda s: s.buf[17:17 + 225].reshape(15, 15))
    jac = property(lambda s: s.buf[17 + 225:17 + 450].reshape(15, 15))
    sqrt_info = property(lambda s: s.buf[17 + 450:17 + 675].reshape(15, 15))


def imu_factor(pre, gravity, pri, vbi, prj, vbj, jac=True):
    r = np.zeros(15)
    J = np.zeros((15, 30))
    lib().orc_imu_factor(_p(pre.buf), _p(_f64(gravity)), _p(_f64(pri)), _p(_f64(vbi)), _p(_f64(prj)), _p(_f64(vbj)), _p(r),
                         _p(J) if jac else None)
    return r, J


def imu_predict(prev16, pre):
    out = np.zeros(16)
    lib().orc_imu_predict(_p(_f64(prev16)), _p(pre.buf), _p(out))
    return out


def estimate_window(omap, corners, surfs, exTlb, states, preints, gravity=(0, 0, -9.805), params=None):
    """Estimator::Estimate for a window of len(corners) frames. states: [W, 16] (P, q_wxyz, V, bg, ba), returned
    updated. preints[f] links frame f-1 to f (preints[0] is ignored)."""
    W = len(corners)
    cs = [_f32(c).reshape(-1, 4) for c in corners]
    ss = [_f32(s).reshape(-1, 4) for s in surfs]
    cp = (C.c_void_p * W)(*[c.ctypes.data for c in cs])
    sp = (C.c_void_p * W)(*[s.ctypes.data for s in ss])
    nc = np.array([c.shape[0] for c in cs], np.int32)
    ns = np.array([s.shape[0] for s in ss], np.int32)
    pp = (C.c_void_p * W)(*[(p.buf.ctypes.data if p is not None else None) for p in preints])
    st = _f64(states).reshape(W, 16).copy()
    stats = np.zeros(16)
    prm = params if params is not None else est_params()
    rc = lib().orc_estimate_window(omap.h, W, cp, _p(nc), sp, _p(ns), _p(_f64(exTlb).reshape(16)), _p(st), pp,
                                   _p(_f64(gravity)), C.byref(prm), _p(stats))
    assert rc == 0
    return st, stats
