"""ctypes wrapper of the REFERENCE PIN (oracle/_ref/libmmloam_ref.so).

TEST INFRASTRUCTURE ONLY. The library is the reference's own hot-path text, extracted verbatim from
/root/reference by oracle/ref/extract.sh and compiled against stand-in Eigen / PCL / Ceres / Sophus /
ROS headers (oracle/ref/shim/). It is built in the authoring container by `make -C oracle ref`
(needs /root/reference) and travels to the GPU box as a built artefact; `available()` is False where
it was never built, and the tests that need it skip.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libmmloam_ref.so")
_lib = None


def available() -> bool:
    return os.path.exists(_SO)


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(_SO)
        _lib.ref_est_create.restype = C.c_void_p
        _lib.ref_est_create.argtypes = [C.c_float, C.c_float]
        _lib.ref_describe.restype = C.c_char_p
        _lib.ref_localizability.restype = C.c_double
        for name in ("ref_est_destroy", "ref_est_map_thread_step"):
            getattr(_lib, name).argtypes = [C.c_void_p]
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def detect_feature_points(xyzi):
    """feature_extraction::detectFeaturePoints (FE.cpp:341-844) on one scan line."""
    xyzi = _f32(xyzi).reshape(-1, 4)
    n = xyzi.shape[0]
    sharp = np.zeros(max(n, 1), np.int32)
    flat = np.zeros(max(n, 1), np.int32)
    ns, nf = C.c_int(0), C.c_int(0)
    lib().ref_detect_feature_points(_p(xyzi), n, _p(sharp), C.byref(ns), _p(flat), C.byref(nf))
    return sharp[: ns.value].copy(), flat[: nf.value].copy()


def hori_extract(offset_time, xyz, refl, line, used_line=6):
    """getHoriFeatureExtract (FE.cpp:952-1035). Returns the kept cloud, rows
    [x y z intensity rel_time line label]."""
    offset_time = np.ascontiguousarray(offset_time, np.uint32)
    xyz = _f32(xyz).reshape(-1, 3)
    refl = np.ascontiguousarray(refl, np.uint8)
    line = np.ascontiguousarray(line, np.uint8)
    n = xyz.shape[0]
    out = np.zeros((max(n, 1), 7), np.float32)
    m, nc, ns = C.c_int(0), C.c_int(0), C.c_int(0)
    lib().ref_hori_extract(_p(offset_time), _p(xyz), _p(refl), _p(line), n, used_line, _p(out), C.byref(m), C.byref(nc),
                           C.byref(ns))
    return out[: m.value].copy()


def velo_extract(xyzi):
    """Body of getVeloFeature (FE.cpp:1135-1240): ring, relative time, split, detector, labels."""
    xyzi = _f32(xyzi).reshape(-1, 4)
    n = xyzi.shape[0]
    out = np.zeros((max(n, 1), 7), np.float32)
    m = C.c_int(0)
    lib().ref_velo_extract(_p(xyzi), n, _p(out), C.byref(m))
    return out[: m.value].copy()


def extract_scan(xyzi, line_id, n_lines):
    """Label a scan whose line ids are given: the reference's split / detector / label glue
    (FE.cpp:1001-1023) restated around the verbatim detector, for scans that do not come from a
    CustomMsg or a PointCloud2."""
    xyzi = _f32(xyzi).reshape(-1, 4)
    line_id = np.asarray(line_id)
    label = np.zeros(xyzi.shape[0], np.uint8)
    for l in range(n_lines):
        src = np.nonzero(line_id == l)[0]
        sharp, flat = detect_feature_points(xyzi[src])
        label[src[sharp]] = 1
        label[src[flat]] = 2
    return label


def undistort(xyzi, s, dR, dt):
    out = _f32(xyzi).copy()
    s = _f32(s)
    dR = _f64(dR).reshape(9)
    dt = _f64(dt).reshape(3)
    lib().ref_undistort(_p(out), _p(s), out.shape[0], _p(dR), _p(dt))
    return out


def point_to_map(p3, T):
    p3 = _f32(p3)
    T = _f64(T).reshape(16)
    out = np.zeros(3, np.float32)
    lib().ref_point_to_map(_p(p3), _p(T), _p(out))
    return out


def residual(kind, feat, x6, T_bl, lidar_m=1.5e-3):
    """Reference cost functor + dual-number autodiff. kind 0: feat = pointOri, lineP1, lineP2 (9);
    kind 1: feat = pointOri, pointProj, sqrt_info row-major (15)."""
    f = _f64(feat)
    x6 = _f64(x6)
    T = _f64(T_bl).reshape(16)
    r = np.zeros(3, np.float64)
    J = np.zeros(18, np.float64)
    rc = lib().ref_residual(kind, _p(f), _p(x6), _p(T), C.c_double(lidar_m), _p(r), _p(J))
    assert rc == 0
    n = 1 if kind == 0 else 3
    return r[:n].copy(), J[: 6 * n].reshape(n, 6).copy()


def imu_preintegrate(t, gyr, acc, last_time, bg=(0, 0, 0), ba=(0, 0, 0)):
    """IMUIntegrator::PreIntegration (IMU.cpp:105-166). Returns dq(wxyz), dp, dv, dt, cov, jac."""
    t = _f64(t)
    gyr = _f64(gyr).reshape(-1, 3)
    acc = _f64(acc).reshape(-1, 3)
    out = np.zeros(11 + 450)
    lib().ref_imu_preintegrate(_p(t), _p(gyr), _p(acc), t.shape[0], C.c_double(last_time), _p(_f64(bg)), _p(_f64(ba)), _p(out))
    return out[0:4].copy(), out[4:7].copy(), out[7:10].copy(), float(out[10]), out[11:236].reshape(15, 15).copy(), \
        out[236:461].reshape(15, 15).copy()


def imu_factor(t, gyr, acc, last_time, bg, ba, gravity, pri, vbi, prj, vbj):
    """Cost_NavState_PRV_Bias (CF.h:321-393) with the reference's sqrt_information (EST.cpp:1238-1242), residuals and
    Jacobian (15 x 30, columns [PR_i | VBias_i | PR_j | VBias_j]) by dual-number autodiff."""
    t = _f64(t)
    gyr = _f64(gyr).reshape(-1, 3)
    acc = _f64(acc).reshape(-1, 3)
    r = np.zeros(15)
    J = np.zeros((15, 30))
    rc = lib().ref_imu_factor(_p(t), _p(gyr), _p(acc), t.shape[0], C.c_double(last_time), _p(_f64(bg)), _p(_f64(ba)),
                              _p(_f64(gravity)), _p(_f64(pri)), _p(_f64(vbi)), _p(_f64(prj)), _p(_f64(vbj)), _p(r), _p(J))
    assert rc == 0
    return r, J


class Estimator:
    """The reference's Estimator object (EST.h / EST.cpp verbatim) with its MAP_MANAGER."""

    def __init__(self, filter_corner=0.4, filter_surf=0.2):
        self.h = C.c_void_p(lib().ref_est_create(filter_corner, filter_surf))

    def close(self):
        if self.h:
            lib().ref_est_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def map_thread_step(self):
        lib().ref_est_map_thread_step(self.h)

    def cube_index(self, p3, kind=0):
        p3 = _f32(p3)
        return lib().ref_cube_index(self.h, _p(p3), kind)

    def map_increment(self, corner, surf, T=np.eye(4)):
        corner = _f32(corner).reshape(-1, 4)
        surf = _f32(surf).reshape(-1, 4)
        T = _f64(T).reshape(16)
        lib().ref_est_map_increment(self.h, _p(corner), corner.shape[0], _p(surf), surf.shape[0], _p(T))

    def global_map(self, kind):
        m = C.c_int(0)
        cen = np.zeros(3, np.int32)
        lib().ref_est_get_global_map(self.h, kind, None, 0, C.byref(m), _p(cen))
        out = np.zeros((max(m.value, 1), 4), np.float32)
        lib().ref_est_get_global_map(self.h, kind, _p(out), m.value, C.byref(m), _p(cen))
        return out[: m.value].copy(), tuple(int(c) for c in cen)

    def map_increment_local(self, corner, surf, T):
        corner = _f32(corner).reshape(-1, 4)
        surf = _f32(surf).reshape(-1, 4)
        T = _f64(T).reshape(16)
        lib().ref_est_map_increment_local(self.h, _p(corner), corner.shape[0], _p(surf), surf.shape[0], _p(T))

    def set_local_map(self, kind, xyzi):
        xyzi = _f32(xyzi).reshape(-1, 4)
        lib().ref_est_set_local_map(self.h, kind, _p(xyzi), xyzi.shape[0])

    def local_map(self, kind):
        m = C.c_int(0)
        lib().ref_est_get_local_map(self.h, kind, None, 0, C.byref(m))
        out = np.zeros((max(m.value, 1), 4), np.float32)
        lib().ref_est_get_local_map(self.h, kind, _p(out), m.value, C.byref(m))
        return out[: m.value].copy()

    def associate_line(self, q, T_wl, thres, exTlb=np.eye(4)):
        q = _f32(q).reshape(-1, 4)
        T = _f64(T_wl).reshape(16)
        ex = _f64(exTlb).reshape(16)
        feat = np.zeros((max(q.shape[0], 1), 12), np.float64)
        nf = C.c_int(0)
        lib().ref_est_associate_line(self.h, _p(q), q.shape[0], _p(T), _p(ex), C.c_double(thres), _p(feat), C.byref(nf))
        return feat[: nf.value].copy()

    def associate_plane(self, q, T_wl, thres, exTlb=np.eye(4), plan_weight_tan=0.0):
        q = _f32(q).reshape(-1, 4)
        T = _f64(T_wl).reshape(16)
        ex = _f64(exTlb).reshape(16)
        feat = np.zeros((max(q.shape[0], 1), 18), np.float64)
        nf, deg, fail = C.c_int(0), C.c_int(0), C.c_int(0)
        lib().ref_est_associate_plane(self.h, _p(q), q.shape[0], _p(T), _p(ex), C.c_double(thres),
                                      C.c_double(plan_weight_tan), _p(feat), C.byref(nf), C.byref(deg), C.byref(fail))
        return feat[: nf.value].copy(), bool(deg.value), bool(fail.value)

    def localizability(self, normals):
        nrm = _f64(normals).reshape(-1, 3)
        return lib().ref_localizability(self.h, _p(nrm), nrm.shape[0])

    def estimate_lidar_pose(self, cloud7, P, q_wxyz, exTlb=np.eye(4), lidar_mode=2):
        cloud7 = _f32(cloud7).reshape(-1, 7)
        P = _f64(P).copy()
        q = _f64(q_wxyz).copy()
        ex = _f64(exTlb).reshape(16)
        fail = C.c_int(0)
        lib().ref_est_estimate_lidar_pose(self.h, _p(cloud7), cloud7.shape[0], _p(P), _p(q), _p(ex), lidar_mode,
                                          C.byref(fail))
        return P, q, bool(fail.value)

    def estimate_window(self, clouds7, states, stamps, imu, exTlb=np.eye(4), gravity=(0, 0, -9.805), lidar_mode=2):
        """EstimateLidarPose on W frames. clouds7: list of [n,7]; states [W,16]; stamps [W]; imu: list of
        (t, gyr, acc) per frame (entry 0 unused)."""
        W = len(clouds7)
        cl = _f32(np.concatenate([np.asarray(c, np.float32).reshape(-1, 7) for c in clouds7]))
        npts = np.array([len(c) for c in clouds7], np.int32)
        st = _f64(states).reshape(W, 16).copy()
        it = _f64(np.concatenate([np.asarray(i[0], float).ravel() for i in imu]))
        ig = _f64(np.concatenate([np.asarray(i[1], float).reshape(-1, 3) for i in imu]))
        ia = _f64(np.concatenate([np.asarray(i[2], float).reshape(-1, 3) for i in imu]))
        inn = np.array([len(np.asarray(i[0]).ravel()) for i in imu], np.int32)
        fail = C.c_int(0)
        lib().ref_est_estimate_window(self.h, W, _p(cl), _p(npts), _p(st), _p(_f64(stamps)), _p(it), _p(ig), _p(ia), _p(inn),
                                      _p(_f64(exTlb).reshape(16)), _p(_f64(gravity)), lidar_mode, C.byref(fail))
        return st, bool(fail.value)
