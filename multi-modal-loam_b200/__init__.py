"""mmloam_b200 — B200-native scan-matching hot path of TIERS/multi-modal-loam.

Python side of the C-ABI in include/mmloam_b200.h (ctypes; no torch types cross the
boundary). The compute lives in libmmloam_b200.so (hand-written sm_100a CUDA, csrc/).
There is no CPU fallback: importing works anywhere, creating a Context needs a GPU and
the built library, and fails loudly otherwise.

Host-side mirror of the reference call surface (names follow the reference):
  feature_extraction.detectFeaturePoints   mm-loam/src/unionFeatureExtract.cpp:341
  LidarFeatureExtractor.detectFeaturePoint  (LIO-Livox spelling of the same method)
  Estimator.processPointToLine / processPointToPlanVec / EstimateLidarPose
                                            mm-loam/include/Estimator/Estimator.h:159-214
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MMLOAM_B200_LIB") or os.path.join(_HERE, "libmmloam_b200.so")
_lib = None

MAP_CORNER_GLOBAL, MAP_SURF_GLOBAL, MAP_CORNER_LOCAL, MAP_SURF_LOCAL = 0, 1, 2, 3

ERRORS = {0: "ok", -1: "invalid argument", -2: "no CUDA device", -3: "CUDA error", -4: "capacity exceeded",
          -5: "bad call order"}

# every symbol include/mmloam_b200.h declares (checked by tests/test_abi.py)
EXPORTS = [
    "mml_version", "mml_ctx_create", "mml_ctx_destroy", "mml_last_error", "mml_launch_count", "mml_sync",
    "mml_extract_features", "mml_extract_features_batch", "mml_velo_ring_time", "mml_hori_filter", "mml_undistort",
    "mml_voxel_downsample", "mml_map_set", "mml_map_set_dev", "mml_associate", "mml_accumulate", "mml_est_params_default",
    "mml_estimate", "mml_scan_to_pose", "mml_scan_to_pose_dev", "mml_frame_set", "mml_frame_associate",
    "mml_frame_accumulate", "mml_frame_associate_async", "mml_frame_associate_kind_async", "mml_frame_accumulate_async",
    "mml_odom_run", "mml_local_map_push", "mml_local_map_push_dev", "mml_local_map_seed", "mml_global_map_push", "mml_global_map_get", "mml_global_map_reset", "mml_unpack_custom_points", "mml_unpack_pointcloud2", "mml_pack_union_clouds", "mml_local_map_get", "mml_local_map_reset", "mml_timer_start",
    "mml_timer_stop_ms", "mml_frame_accumulate_partial_dev", "mml_stream_handle",
    "mml_imu_preintegrate", "mml_imu_preintegrate_mean", "mml_imu_factor", "mml_imu_predict", "mml_window_reset", "mml_window_size",
    "mml_window_push_frame", "mml_window_push_scan_dev", "mml_window_get_frame", "mml_estimate_window",
    "mml_odom_run_window", "mml_shard_init", "mml_shard_local_ptr", "mml_shard_connect_ipc", "mml_shard_connect_ptrs",
    "mml_shard_close", "mml_estimate_sharded",
]


class MmlError(RuntimeError):
    pass


class Preint(C.Structure):
    """mml_preint: IMUIntegrator after PreIntegration (IMU.cpp:105-166)."""
    _fields_ = [("dq", C.c_double * 4), ("dp", C.c_double * 3), ("dv", C.c_double * 3), ("dt", C.c_double),
                ("bg", C.c_double * 3), ("ba", C.c_double * 3), ("cov", C.c_double * 225), ("jac", C.c_double * 225),
                ("sqrt_info", C.c_double * 225)]


def imu_preintegrate(t, gyr, acc, last_time, bg=(0, 0, 0), ba=(0, 0, 0)):
    """IMUIntegrator::PreIntegration on the samples of (last_time, t_frame] (host code, no GPU needed)."""
    t = _f64(t)
    gyr = _f64(gyr).reshape(-1, 3)
    acc = _f64(acc).reshape(-1, 3)
    out = Preint()
    rc = load_library().mml_imu_preintegrate(_p(t), _p(gyr), _p(acc), int(t.shape[0]), C.c_double(last_time), _p(_f64(bg)),
                                             _p(_f64(ba)), C.byref(out))
    if rc != 0:
        raise MmlError(f"mml_imu_preintegrate: {ERRORS.get(rc, rc)}")
    return out


def imu_preintegrate_mean(t, gyr, acc, last_time, bg=(0, 0, 0), ba=(0, 0, 0)):
    """The mean of the pre-integration alone (dq, dp, dv, dt): what the pose prediction needs (host code)."""
    t = _f64(t)
    gyr = _f64(gyr).reshape(-1, 3)
    acc = _f64(acc).reshape(-1, 3)
    out = Preint()
    rc = load_library().mml_imu_preintegrate_mean(_p(t), _p(gyr), _p(acc), int(t.shape[0]), C.c_double(last_time), _p(_f64(bg)),
                                                  _p(_f64(ba)), C.byref(out))
    if rc != 0:
        raise MmlError(f"mml_imu_preintegrate_mean: {ERRORS.get(rc, rc)}")
    return out


def imu_factor(pre, gravity, pri, vbi, prj, vbj):
    """Cost_NavState_PRV_Bias weighted by sqrt_info: (r15, J[15, 30])."""
    r = np.zeros(15)
    J = np.zeros((15, 30))
    rc = load_library().mml_imu_factor(C.byref(pre), _p(_f64(gravity)), _p(_f64(pri)), _p(_f64(vbi)), _p(_f64(prj)),
                                       _p(_f64(vbj)), _p(r), _p(J))
    if rc != 0:
        raise MmlError(f"mml_imu_factor: {ERRORS.get(rc, rc)}")
    return r, J


def imu_predict(prev16, pre):
    """Pose prediction of process() (PE.cpp:812-829): state = P, q_wxyz, V, bg, ba."""
    out = np.zeros(16)
    rc = load_library().mml_imu_predict(_p(_f64(prev16)), C.byref(pre), _p(out))
    if rc != 0:
        raise MmlError(f"mml_imu_predict: {ERRORS.get(rc, rc)}")
    return out


class EstParams(C.Structure):
    _fields_ = [("max_outer", C.c_int), ("max_inner", C.c_int), ("lidar_m", C.c_double),
                ("plan_weight_tan", C.c_double), ("thres0", C.c_double), ("thres1", C.c_double),
                ("thres2", C.c_double), ("use_huber", C.c_int), ("map_update", C.c_int)]


def load_library():
    """dlopen libmmloam_b200.so. Raises MmlError if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MmlError(f"{LIB_PATH} is missing: run __graft_entry__.build() (nvcc, sm_100a). "
                           "mmloam_b200 has no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        lib.mml_last_error.restype = C.c_char_p
        lib.mml_last_error.argtypes = [C.c_void_p]
        lib.mml_launch_count.restype = C.c_longlong
        lib.mml_launch_count.argtypes = [C.c_void_p]
        lib.mml_stream_handle.restype = C.c_void_p
        lib.mml_stream_handle.argtypes = [C.c_void_p]
        _lib = lib
    return _lib


def est_params(**kw):
    p = EstParams()
    load_library().mml_est_params_default(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Context:
    """Owns one mml_ctx (streams, scratch, resident maps). One per host thread."""

    def __init__(self, device=0, streams=1):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.mml_ctx_create(int(device), int(streams), C.byref(h))
        if rc != 0:
            raise MmlError(f"mml_ctx_create failed: {ERRORS.get(rc, rc)} (a CUDA device is required; no CPU fallback)")
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.mml_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            msg = self.lib.mml_last_error(self.h)
            raise MmlError(f"{ERRORS.get(rc, rc)}: {msg.decode() if msg else ''}")

    @property
    def launches(self):
        return int(self.lib.mml_launch_count(self.h))

    def sync(self):
        self._ck(self.lib.mml_sync(self.h))

    # ---- A1
    def extract_features(self, xyzi, line_id, n_lines):
        xyzi = _f32(xyzi).reshape(-1, 4)
        line_id = np.ascontiguousarray(line_id, np.uint16)
        n = xyzi.shape[0]
        label = np.zeros(max(n, 1), np.uint8)
        ns, nf = C.c_int(0), C.c_int(0)
        self._ck(self.lib.mml_extract_features(self.h, _p(xyzi), _p(line_id), n, int(n_lines), _p(label), C.byref(ns),
                                               C.byref(nf)))
        return label[:n], ns.value, nf.value

    def extract_features_batch(self, xyzi, line_id, scan_offsets, n_lines):
        xyzi = _f32(xyzi).reshape(-1, 4)
        line_id = np.ascontiguousarray(line_id, np.uint16)
        off = np.ascontiguousarray(scan_offsets, np.int32)
        ns = off.shape[0] - 1
        n = xyzi.shape[0]
        label = np.zeros(max(n, 1), np.uint8)
        n_sharp = np.zeros(max(ns, 1), np.int32)
        n_flat = np.zeros(max(ns, 1), np.int32)
        self._ck(self.lib.mml_extract_features_batch(self.h, _p(xyzi), _p(line_id), _p(off), ns, int(n_lines), _p(label),
                                                     _p(n_sharp), _p(n_flat)))
        return label[:n], n_sharp[:ns], n_flat[:ns]

    def extract_batch_dev_async(self, xyzi_dev, line_dev, scan_offsets, n_lines, label_dev):
        off = np.ascontiguousarray(scan_offsets, np.int32)
        self._ck(self.lib.mml_extract_features_batch_dev(self.h, xyzi_dev, line_dev, _p(off), off.shape[0] - 1, int(n_lines),
                                                         label_dev))

    # ---- A2 / A3
    def velo_ring_time(self, xyzi):
        xyzi = _f32(xyzi).reshape(-1, 4)
        n = xyzi.shape[0]
        line = np.zeros(max(n, 1), np.int16)
        rt = np.zeros(max(n, 1), np.float32)
        self._ck(self.lib.mml_velo_ring_time(self.h, _p(xyzi), n, _p(line), _p(rt)))
        return line[:n], rt[:n]

    def hori_filter(self, offset_time, xyz, line):
        offset_time = np.ascontiguousarray(offset_time, np.uint32)
        xyz = _f32(xyz).reshape(-1, 3)
        line = np.ascontiguousarray(line, np.uint8)
        n = xyz.shape[0]
        keep = np.zeros(max(n, 1), np.uint8)
        rt = np.zeros(max(n, 1), np.float32)
        self._ck(self.lib.mml_hori_filter(self.h, _p(offset_time), _p(xyz), _p(line), n, _p(keep), _p(rt)))
        return keep[:n], rt[:n]

    # ---- A4
    def undistort(self, xyzi, s, dR, dt):
        out = _f32(xyzi).reshape(-1, 4).copy()
        s = _f32(s)
        dR = _f64(dR).reshape(9)
        dt = _f64(dt).reshape(3)
        self._ck(self.lib.mml_undistort(self.h, _p(out), _p(s), out.shape[0], _p(dR), _p(dt)))
        return out

    # ---- A6
    def voxel_downsample(self, xyzi, leaf):
        xyzi = _f32(xyzi).reshape(-1, 4)
        n = xyzi.shape[0]
        out = np.zeros((max(n, 1), 4), np.float32)
        m = C.c_int(0)
        self._ck(self.lib.mml_voxel_downsample(self.h, _p(xyzi), n, C.c_float(leaf), _p(out), C.byref(m)))
        return out[: m.value].copy()

    # ---- maps
    def map_set(self, kind, xyzi, cen=(10, 5, 10), cell=0.0):
        xyzi = _f32(xyzi).reshape(-1, 4)
        cen = np.asarray(cen, np.int32)
        self._ck(self.lib.mml_map_set_ex(self.h, int(kind), _p(xyzi), xyzi.shape[0], _p(cen), C.c_float(cell)))

    def map_info(self, kind):
        info = np.zeros(8, np.float64)
        self._ck(self.lib.mml_map_info(self.h, int(kind), _p(info)))
        return dict(valid=bool(info[0]), m=int(info[1]), cell=float(info[2]), dim=tuple(int(v) for v in info[3:6]),
                    ncell=int(info[6]), k_per_cube=int(info[7]))

    # ---- A7 / A8
    def associate(self, kind, q, T_wl, thres):
        q = _f32(q).reshape(-1, 4)
        T = _f64(T_wl).reshape(16)
        nq = q.shape[0]
        feat = np.zeros((max(nq, 1), 12), np.float64)
        nf, nn = C.c_int(0), C.c_int(0)
        M = np.zeros(9, np.float64)
        self._ck(self.lib.mml_associate(self.h, int(kind), _p(q), nq, _p(T), C.c_double(thres), _p(feat), C.byref(nf), _p(M),
                                        C.byref(nn)))
        return feat[:nq], nf.value, M.reshape(3, 3), nn.value

    # ---- A9-A11
    def accumulate(self, line_feat, plane_feat, x6, T_bl, plan_weight_tan=0.0, huber_a=0.1 / 1.5e-3):
        lf = _f64(line_feat).reshape(-1, 12)
        pf = _f64(plane_feat).reshape(-1, 12)
        x6 = _f64(x6)
        T = _f64(T_bl).reshape(16)
        H = np.zeros(36)
        g = np.zeros(6)
        cost = C.c_double(0)
        self._ck(self.lib.mml_accumulate(self.h, _p(lf), lf.shape[0], _p(pf), pf.shape[0], _p(x6), _p(T),
                                         C.c_double(plan_weight_tan), C.c_double(huber_a), _p(H), _p(g), C.byref(cost)))
        return H.reshape(6, 6), g, cost.value

    # ---- frame slot (device resident)
    def frame_set(self, corner, surf):
        corner = _f32(corner).reshape(-1, 4)
        surf = _f32(surf).reshape(-1, 4)
        self._ck(self.lib.mml_frame_set(self.h, _p(corner), corner.shape[0], _p(surf), surf.shape[0]))
        self._nq = (corner.shape[0], surf.shape[0])

    def frame_associate(self, T_wl, thres):
        T = _f64(T_wl).reshape(16)
        nl, np_, nn = C.c_int(0), C.c_int(0), C.c_int(0)
        M = np.zeros(9)
        self._ck(self.lib.mml_frame_associate(self.h, _p(T), C.c_double(thres), C.byref(nl), C.byref(np_), _p(M), C.byref(nn)))
        return nl.value, np_.value, M.reshape(3, 3), nn.value

    def frame_get_features(self, kind):
        nq = self._nq[kind]
        out = np.zeros((max(nq, 1), 12), np.float64)
        self._ck(self.lib.mml_frame_get_features(self.h, int(kind), _p(out)))
        return out[:nq]

    def frame_accumulate(self, x6, T_bl, plan_weight_tan=0.0, huber_a=0.1 / 1.5e-3):
        x6 = _f64(x6)
        T = _f64(T_bl).reshape(16)
        H = np.zeros(36)
        g = np.zeros(6)
        cost = C.c_double(0)
        self._ck(self.lib.mml_frame_accumulate(self.h, _p(x6), _p(T), C.c_double(plan_weight_tan), C.c_double(huber_a),
                                               _p(H), _p(g), C.byref(cost)))
        return H.reshape(6, 6), g, cost.value

    def frame_associate_async(self, T_wl, thres, repeat=1):
        T = _f64(T_wl).reshape(16)
        self._ck(self.lib.mml_frame_associate_async(self.h, _p(T), C.c_double(thres), int(repeat)))

    def frame_associate_kind_async(self, kind, T_wl, thres, repeat=1):
        """one association kernel only (0 line / 1 plane): used to time a single kernel"""
        T = _f64(T_wl).reshape(16)
        self._ck(self.lib.mml_frame_associate_kind_async(self.h, int(kind), _p(T), C.c_double(thres), int(repeat)))

    def frame_accumulate_async(self, x6, T_bl, plan_weight_tan=0.0, huber_a=0.1 / 1.5e-3, repeat=1):
        x6 = _f64(x6)
        T = _f64(T_bl).reshape(16)
        self._ck(self.lib.mml_frame_accumulate_async(self.h, _p(x6), _p(T), C.c_double(plan_weight_tan),
                                                     C.c_double(huber_a), int(repeat)))

    def profile_enable(self, on=True):
        self._ck(self.lib.mml_profile_enable(self.h, 1 if on else 0))

    def profile_read(self):
        """(ms per stage [extract, undistort+split+voxel, estimate] summed over scans, n_scans)"""
        ms = np.zeros(3)
        n = C.c_longlong(0)
        self._ck(self.lib.mml_profile_read(self.h, _p(ms), C.byref(n)))
        return ms, n.value

    def timer_start(self):
        self._ck(self.lib.mml_timer_start(self.h))

    def timer_stop_ms(self):
        ms = C.c_float(0)
        self._ck(self.lib.mml_timer_stop_ms(self.h, C.byref(ms)))
        return ms.value

    # ---- A12
    def estimate(self, corner, surf, exTlb, P, q_wxyz, params=None):
        if corner is None and surf is None:   # the frame frame_set() left in HBM
            ex = _f64(exTlb).reshape(16)
            P = _f64(P).copy()
            q = _f64(q_wxyz).copy()
            stats = np.zeros(16)
            prm = params if params is not None else est_params()
            self._ck(self.lib.mml_estimate(self.h, None, -1, None, -1, _p(ex), _p(P), _p(q), C.byref(prm), _p(stats)))
            return P, q, stats
        corner = _f32(corner).reshape(-1, 4)
        surf = _f32(surf).reshape(-1, 4)
        ex = _f64(exTlb).reshape(16)
        P = _f64(P).copy()
        q = _f64(q_wxyz).copy()
        stats = np.zeros(16)
        prm = params if params is not None else est_params()
        self._ck(self.lib.mml_estimate(self.h, _p(corner), corner.shape[0], _p(surf), surf.shape[0], _p(ex), _p(P), _p(q),
                                       C.byref(prm), _p(stats)))
        return P, q, stats

    # ---- (e) cube-sharded global map over several GPUs: exchange through peer memory inside the kernels
    def shard_init(self, rank, world):
        """Allocate this rank's exchange buffer; returns its 64-byte cudaIpcMemHandle (bytes) for an all-gather."""
        h = (C.c_ubyte * 64)()
        self._ck(self.lib.mml_shard_init(self.h, int(rank), int(world), h))
        return bytes(h)

    def shard_local_ptr(self):
        p = C.c_void_p()
        self._ck(self.lib.mml_shard_local_ptr(self.h, C.byref(p)))
        return p.value

    def shard_connect_ipc(self, handles):
        """handles: list of the ranks' 64-byte handles in rank order (one process per GPU)."""
        raw = b"".join(handles)
        self._ck(self.lib.mml_shard_connect_ipc(self.h, raw))

    def shard_connect_ptrs(self, ptrs, devices=None):
        """ptrs: exchange-buffer pointers of all ranks living in this process (shard_local_ptr of each context)."""
        arr = (C.c_void_p * len(ptrs))(*ptrs)
        dev = (C.c_int * len(ptrs))(*devices) if devices is not None else None
        self._ck(self.lib.mml_shard_connect_ptrs(self.h, arr, dev))

    def shard_close(self):
        self._ck(self.lib.mml_shard_close(self.h))

    def estimate_sharded(self, corner, surf, exTlb, P, q_wxyz, params=None):
        """mml_estimate_sharded: a collective over the ranks of shard_init (same queries and start pose on each).
        corner = surf = None: solve the frame frame_set() left in HBM."""
        if corner is None and surf is None:
            ex = _f64(exTlb).reshape(16)
            P = _f64(P).copy()
            q = _f64(q_wxyz).copy()
            stats = np.zeros(16)
            prm = params if params is not None else est_params()
            self._ck(self.lib.mml_estimate_sharded(self.h, None, -1, None, -1, _p(ex), _p(P), _p(q), C.byref(prm), _p(stats)))
            return P, q, stats
        corner = _f32(corner).reshape(-1, 4)
        surf = _f32(surf).reshape(-1, 4)
        ex = _f64(exTlb).reshape(16)
        P = _f64(P).copy()
        q = _f64(q_wxyz).copy()
        stats = np.zeros(16)
        prm = params if params is not None else est_params()
        self._ck(self.lib.mml_estimate_sharded(self.h, _p(corner), corner.shape[0], _p(surf), surf.shape[0], _p(ex), _p(P), _p(q),
                                               C.byref(prm), _p(stats)))
        return P, q, stats

    # ---- whole per-scan path
    def scan_to_pose(self, xyzi, line_id, s, n_lines, dR, dt, exTlb, P, q_wxyz, leaf_corner=0.4, leaf_surf=0.2,
                     params=None):
        xyzi = _f32(xyzi).reshape(-1, 4)
        line_id = np.ascontiguousarray(line_id, np.uint16)
        s = _f32(s) if s is not None else None
        dR = _f64(dR).reshape(9) if dR is not None else None
        dt = _f64(dt).reshape(3) if dt is not None else None
        ex = _f64(exTlb).reshape(16)
        P = _f64(P).copy()
        q = _f64(q_wxyz).copy()
        stats = np.zeros(16)
        counts = np.zeros(4, np.int32)
        prm = params if params is not None else est_params()
        self._ck(self.lib.mml_scan_to_pose(self.h, _p(xyzi), _p(line_id), _p(s), xyzi.shape[0], int(n_lines), _p(dR), _p(dt),
                                           C.c_float(leaf_corner), C.c_float(leaf_surf), _p(ex), _p(P), _p(q),
                                           C.byref(prm), _p(stats), _p(counts)))
        return P, q, stats, counts

    # ---- native odometry loop (pipelined extraction, constant-velocity prediction)
    def odom_run(self, scans, n_lines, T_init, T_prev, exTlb, host_buffers=False, leaf_corner=0.4, leaf_surf=0.2,
                 params=None):
        """scans: list of (xyzi, line, s, n). Device pointers (ctypes.c_void_p or int) when host_buffers is False,
        numpy arrays or raw host addresses (ideally pinned memory) otherwise. Returns (poses [k,4,4], total_ms, counts [k,4])."""
        k = len(scans)
        PtrArr = C.c_void_p * max(k, 1)

        def ptr(v):
            if isinstance(v, C.c_void_p):
                return v.value
            if isinstance(v, np.ndarray):
                return v.ctypes.data
            return int(v) if v is not None else None  # a raw address (device, or pinned host memory with host_buffers)

        xs = PtrArr(*[ptr(sc[0]) for sc in scans])
        ls = PtrArr(*[ptr(sc[1]) for sc in scans])
        ss = PtrArr(*[ptr(sc[2]) for sc in scans])
        ns = np.ascontiguousarray([int(sc[3]) for sc in scans], np.int32)
        poses = np.zeros((max(k, 1), 16), np.float64)
        counts = np.zeros((max(k, 1), 4), np.int32)
        ms = C.c_float(0)
        Ti = _f64(T_init).reshape(16)
        Tp = _f64(T_prev).reshape(16)
        ex = _f64(exTlb).reshape(16)
        prm = params if params is not None else est_params()
        self._ck(self.lib.mml_odom_run(self.h, xs, ls, ss, _p(ns), k, int(n_lines), 1 if host_buffers else 0, _p(Ti), _p(Tp),
                                       _p(ex), C.c_float(leaf_corner), C.c_float(leaf_surf), C.byref(prm), _p(poses),
                                       C.byref(ms), _p(counts)))
        return poses[:k].reshape(k, 4, 4), ms.value, counts[:k]

    # ---- sliding window (sizes 2-4, IMU factors) --------------------------------------------------------------
    def window_reset(self):
        self._ck(self.lib.mml_window_reset(self.h))

    def window_size(self):
        return int(self.lib.mml_window_size(self.h))

    def window_push_frame(self, corner, surf, max_frames=3):
        corner = _f32(corner).reshape(-1, 4)
        surf = _f32(surf).reshape(-1, 4)
        self._ck(self.lib.mml_window_push_frame(self.h, _p(corner), corner.shape[0], _p(surf), surf.shape[0], int(max_frames)))

    def window_push_scan_dev(self, xyzi_dev, line_dev, s_dev, n, n_lines, dR, dt, max_frames=3, leaf_corner=0.4,
                             leaf_surf=0.2):
        xyzi_dev, line_dev, s_dev = [v.value if isinstance(v, C.c_void_p) else v for v in (xyzi_dev, line_dev, s_dev)]
        dRa = _f64(dR).reshape(9) if dR is not None else None
        dta = _f64(dt).reshape(3) if dt is not None else None
        counts = np.zeros(4, np.int32)
        self._ck(self.lib.mml_window_push_scan_dev(self.h, C.c_void_p(xyzi_dev), C.c_void_p(line_dev),
                                                   C.c_void_p(s_dev) if s_dev else None, int(n), int(n_lines), _p(dRa),
                                                   _p(dta), C.c_float(leaf_corner), C.c_float(leaf_surf), int(max_frames),
                                                   _p(counts)))
        return counts

    def window_get_frame(self, f, kind):
        n = C.c_int(0)
        self._ck(self.lib.mml_window_get_frame(self.h, int(f), int(kind), None, 0, C.byref(n)))
        out = np.zeros((max(n.value, 1), 4), np.float32)
        self._ck(self.lib.mml_window_get_frame(self.h, int(f), int(kind), _p(out), n.value, C.byref(n)))
        return out[: n.value].copy()

    def estimate_window(self, states, preints, exTlb=np.eye(4), gravity=(0, 0, -9.805), params=None):
        """Estimator::Estimate on the frames in the window. states [W, 16] = P, q_wxyz, V, bg, ba per frame;
        preints[f] (f >= 1) links frame f-1 to f. Returns (states, stats)."""
        W = self.window_size()
        st = _f64(states).reshape(W, 16).copy()
        pp = (C.POINTER(Preint) * W)()
        for f in range(W):
            if preints[f] is not None:
                pp[f] = C.pointer(preints[f])
        stats = np.zeros(16)
        prm = params if params is not None else est_params()
        self._ck(self.lib.mml_estimate_window(self.h, _p(st), pp, _p(_f64(exTlb).reshape(16)), _p(_f64(gravity)),
                                              C.byref(prm), _p(stats)))
        return st, stats

    def odom_run_window(self, scans, n_lines, window, stamps, stamp0, imu, state0, exTlb=np.eye(4), gravity=(0, 0, -9.805),
                        host_buffers=False, leaf_corner=0.4, leaf_surf=0.2, params=None):
        """mml_odom_run_window. scans: list of (xyzi, line, s, n) with device pointers (ints) or, with
        host_buffers=True, numpy arrays. imu: list of (t, gyr, acc) per scan. Returns dict with poses_front,
        poses_newest, states, stats, total_ms."""
        n = len(scans)
        keep = []

        def ptr(v):
            if isinstance(v, C.c_void_p):
                return v.value
            if isinstance(v, np.ndarray):
                keep.append(v)
                return v.__array_interface__["data"][0]  # (.ctypes.data builds an object per access: this call is timed end to end)
            return int(v) if v is not None else None

        def f64(v, cols):
            if isinstance(v, np.ndarray) and v.dtype == np.float64:
                return v.reshape(-1) if cols == 1 else v.reshape(-1, 3)
            v = np.asarray(v, float)
            return v.ravel() if cols == 1 else v.reshape(-1, 3)

        xs = (C.c_void_p * n)(*[ptr(sc[0]) for sc in scans])
        ls = (C.c_void_p * n)(*[ptr(sc[1]) for sc in scans])
        ss = (C.c_void_p * n)(*[ptr(sc[2]) for sc in scans])
        npts = np.fromiter((sc[3] for sc in scans), np.int32, n)
        ts = [f64(i[0], 1) for i in imu]
        it = _f64(np.concatenate(ts))
        ig = _f64(np.concatenate([f64(i[1], 3) for i in imu]))
        ia = _f64(np.concatenate([f64(i[2], 3) for i in imu]))
        inn = np.fromiter((t.shape[0] for t in ts), np.int32, len(ts))
        pf = np.zeros((n, 16))
        pn = np.zeros((n, 16))
        so = np.zeros((n, 16))
        st = np.zeros((n, 8))
        ms = C.c_float(0)
        prm = params if params is not None else est_params()
        self._ck(self.lib.mml_odom_run_window(self.h, xs, ls, ss, _p(npts), n, int(n_lines), 1 if host_buffers else 0,
                                              int(window), _p(_f64(stamps)), C.c_double(stamp0), _p(it), _p(ig), _p(ia),
                                              _p(inn), _p(_f64(state0)), _p(_f64(exTlb).reshape(16)), _p(_f64(gravity)),
                                              C.c_float(leaf_corner), C.c_float(leaf_surf), C.byref(prm), _p(pf), _p(pn),
                                              _p(so), _p(st), C.byref(ms)))
        return dict(poses_front=pf.reshape(n, 4, 4), poses_newest=pn.reshape(n, 4, 4), states=so, stats=st,
                    total_ms=float(ms.value))

    # ---- global cube map on the device (MAP_MANAGER::MapIncrement + MapMove, MM.cpp:125-281, 288-581)
    def global_map_push(self, corner_w, surf_w, T_wl=None):
        corner_w = np.ascontiguousarray(corner_w, np.float32).reshape(-1, 4)
        surf_w = np.ascontiguousarray(surf_w, np.float32).reshape(-1, 4)
        T = _f64(T_wl).reshape(16) if T_wl is not None else None
        n2 = np.zeros(2, np.int32)
        self._ck(self.lib.mml_global_map_push(self.h, _p(corner_w), int(corner_w.shape[0]), _p(surf_w), int(surf_w.shape[0]),
                                              _p(T) if T is not None else None, _p(n2)))
        return int(n2[0]), int(n2[1])

    def global_map_get(self, kind, which=0):
        """which: 0 all cubes, 1 the matcher's snapshot, 2 laserCloud*FromMap. Returns (cloud [n,4], cen)."""
        n = C.c_int(0)
        cen = np.zeros(3, np.int32)
        self._ck(self.lib.mml_global_map_get(self.h, int(kind), int(which), None, 0, C.byref(n), _p(cen)))
        out = np.zeros((max(n.value, 1), 4), np.float32)
        self._ck(self.lib.mml_global_map_get(self.h, int(kind), int(which), _p(out), int(out.shape[0]), C.byref(n), _p(cen)))
        return out[:n.value], tuple(int(v) for v in cen)

    def global_map_reset(self):
        self._ck(self.lib.mml_global_map_reset(self.h))

    # ---- F2: message unpack / pack on the device (csrc/msgpack.cu)
    def unpack_custom_points(self, points19, used_line=6):
        """CustomMsg points (19-byte records) -> (xyzi [m,4] f32, line [m] u16, s [m] f32), FE.cpp:985-998."""
        raw = np.ascontiguousarray(points19, np.uint8).reshape(-1)
        n = raw.size // 19
        xyzi = np.zeros((max(n, 1), 4), np.float32)
        line = np.zeros(max(n, 1), np.uint16)
        s = np.zeros(max(n, 1), np.float32)
        m = C.c_int(0)
        self._ck(self.lib.mml_unpack_custom_points(self.h, _p(raw), n, int(used_line), _p(xyzi), _p(line), _p(s), C.byref(m)))
        return xyzi[:m.value], line[:m.value], s[:m.value]

    def unpack_pointcloud2(self, data, point_step, off_x=0, off_y=4, off_z=8, off_intensity=12):
        raw = np.ascontiguousarray(data, np.uint8).reshape(-1)
        n = raw.size // int(point_step)
        xyzi = np.zeros((max(n, 1), 4), np.float32)
        m = C.c_int(0)
        self._ck(self.lib.mml_unpack_pointcloud2(self.h, _p(raw), n, int(point_step), int(off_x), int(off_y), int(off_z),
                                                 int(off_intensity), _p(xyzi), C.byref(m)))
        return xyzi[:m.value]

    def pack_union_clouds(self, xyzi, s, line, label, near_full, far_full, near_feat, far_feat, zero_full_intensity=False):
        """-> (full, corner, surf) as [m, 12] float32 views of pcl::PointXYZINormal records."""
        xyzi = np.ascontiguousarray(xyzi, np.float32).reshape(-1, 4)
        n = xyzi.shape[0]
        s = np.ascontiguousarray(s, np.float32)
        line = np.ascontiguousarray(line, np.uint16)
        label = np.ascontiguousarray(label, np.uint8)
        outs = [np.zeros((max(n, 1), 12), np.float32) for _ in range(3)]
        cnt = np.zeros(3, np.int32)
        self._ck(self.lib.mml_pack_union_clouds(self.h, _p(xyzi), _p(s), _p(line), _p(label), n, C.c_float(near_full), C.c_float(far_full),
                                                C.c_float(near_feat), C.c_float(far_feat), 1 if zero_full_intensity else 0,
                                                _p(outs[0]), _p(outs[1]), _p(outs[2]), _p(cnt)))
        return tuple(o[:int(c)] for o, c in zip(outs, cnt))

    # ---- local feature map on the device (Estimator::MapIncrementLocal, EST.cpp:1585-1643)
    def local_map_push(self, corner, surf, T_wl, leaf_corner=0.4, leaf_surf=0.2):
        """One map update from a frame's corner / surf clouds (LiDAR frame) and its pose. Returns the sizes of the new
        local corner / surf maps; the association searches them from now on (map kinds 2 / 3)."""
        corner = np.ascontiguousarray(corner, np.float32).reshape(-1, 4)
        surf = np.ascontiguousarray(surf, np.float32).reshape(-1, 4)
        T = _f64(T_wl).reshape(16)
        nc, ns = C.c_int(0), C.c_int(0)
        self._ck(self.lib.mml_local_map_push(self.h, _p(corner), int(corner.shape[0]), _p(surf), int(surf.shape[0]), _p(T),
                                             C.c_float(leaf_corner), C.c_float(leaf_surf), C.byref(nc), C.byref(ns)))
        return nc.value, ns.value

    def local_map_seed(self, kind, slot, xyzi_world):
        """Place a world-frame cloud into ring entry `slot` of one kind (0 corner / 1 surf)."""
        x = np.ascontiguousarray(xyzi_world, np.float32).reshape(-1, 4)
        self._ck(self.lib.mml_local_map_seed(self.h, int(kind), int(slot), _p(x), int(x.shape[0])))

    def local_map_get(self, kind):
        n = C.c_int(0)
        self._ck(self.lib.mml_local_map_get(self.h, int(kind), None, 0, C.byref(n)))
        out = np.zeros((max(n.value, 1), 4), np.float32)
        self._ck(self.lib.mml_local_map_get(self.h, int(kind), _p(out), int(out.shape[0]), C.byref(n)))
        return out[: n.value].copy()

    def local_map_reset(self):
        self._ck(self.lib.mml_local_map_reset(self.h))

    # ---- device-resident scans (bench.py)
    def map_set_dev(self, kind, xyzi_dev, m, cen=None):
        """mml_map_set with the points already in HBM (pointer from dev_upload)."""
        cen_a = np.asarray(cen, np.int32) if cen is not None else None
        self._ck(self.lib.mml_map_set_dev(self.h, int(kind), xyzi_dev, int(m), _p(cen_a) if cen_a is not None else None))

    def dev_upload(self, arr):
        arr = np.ascontiguousarray(arr)
        ptr = C.c_void_p()
        self._ck(self.lib.mml_dev_alloc(self.h, C.c_size_t(arr.nbytes), C.byref(ptr)))
        self._ck(self.lib.mml_dev_upload(self.h, ptr, _p(arr), C.c_size_t(arr.nbytes)))
        return ptr

    def dev_free(self, ptr):
        self._ck(self.lib.mml_dev_free(self.h, ptr))

    def scan_to_pose_dev(self, xyzi_dev, line_dev, s_dev, n, n_lines, dR, dt, exTlb, P, q_wxyz, leaf_corner=0.4,
                         leaf_surf=0.2, params=None):
        dR = _f64(dR).reshape(9) if dR is not None else None
        dt = _f64(dt).reshape(3) if dt is not None else None
        ex = _f64(exTlb).reshape(16)
        P = _f64(P).copy()
        q = _f64(q_wxyz).copy()
        stats = np.zeros(16)
        counts = np.zeros(4, np.int32)
        prm = params if params is not None else est_params()
        self._ck(self.lib.mml_scan_to_pose_dev(self.h, xyzi_dev, line_dev, s_dev, int(n), int(n_lines), _p(dR), _p(dt),
                                               C.c_float(leaf_corner), C.c_float(leaf_surf), _p(ex), _p(P), _p(q),
                                               C.byref(prm), _p(stats), _p(counts)))
        return P, q, stats, counts


# ---------------------------------------------------------------------------------------
# Host-side mirror of the reference's call surface
# ---------------------------------------------------------------------------------------
class feature_extraction:
    """Mirror of `class feature_extraction` (mm-loam/src/unionFeatureExtract.cpp:143)."""

    def __init__(self, ctx: Context | None = None):
        self.ctx = ctx or Context()

    def detectFeaturePoints(self, cloud_xyzi):
        """FE.cpp:341-343: one scan line in, (pointsLessSharp, pointsLessFlat) line-local indices out."""
        cloud = _f32(cloud_xyzi).reshape(-1, 4)
        n = cloud.shape[0]
        label, _, _ = self.ctx.extract_features(cloud, np.zeros(n, np.uint16), 1)
        return np.nonzero(label == 1)[0].astype(np.int32), np.nonzero(label == 2)[0].astype(np.int32)

    def getFeatures(self, cloud_xyzi, line_id, n_lines):
        """Split / detect / label glue of getHoriFeatureExtract (FE.cpp:952-1035) and getVeloFeature
        (FE.cpp:1113-1317): returns normal_z-style labels (0 / 1 corner / 2 surf) in input order."""
        return self.ctx.extract_features(cloud_xyzi, line_id, n_lines)[0]


class LidarFeatureExtractor(feature_extraction):
    """LIO-Livox spelling of the same class (SURVEY.md §0 naming caveat)."""

    def detectFeaturePoint(self, cloud_xyzi):
        return self.detectFeaturePoints(cloud_xyzi)


class Estimator:
    """Mirror of the hot-path members of `class Estimator` (include/Estimator/Estimator.h:26)."""

    SLIDEWINDOWSIZE = 5  # EST.h:30

    def __init__(self, filter_corner=0.4, filter_surf=0.2, ctx: Context | None = None):
        self.ctx = ctx or Context()
        self.filter_corner = float(filter_corner)
        self.filter_surf = float(filter_surf)
        self.thres_dist = 1.0  # EST.h:336
        self._fail_detected = False

    def setLocalMap(self, corner_xyzi, surf_xyzi):
        """laserCloudCornerFromLocal / SurfFromLocal + kd-tree rebuild (EST.cpp:1159-1167)."""
        self.ctx.map_set(MAP_CORNER_LOCAL, corner_xyzi)
        self.ctx.map_set(MAP_SURF_LOCAL, surf_xyzi)

    def setGlobalMap(self, corner_xyzi, surf_xyzi, cen=(10, 5, 10)):
        """GlobalCornerMap / GlobalSurfMap copies (EST.cpp:1171-1182)."""
        self.ctx.map_set(MAP_CORNER_GLOBAL, corner_xyzi, cen)
        self.ctx.map_set(MAP_SURF_GLOBAL, surf_xyzi, cen)

    def MapIncrementLocal(self, laserCloudCornerStack, laserCloudSurfStack, transformTobeMapped):
        """EST.cpp:1585-1643 (corner and surf clouds): the frame joins the 50-frame ring, the previous filtered map and
        the ring are voxel-filtered again and become the local maps the association searches. Returns their sizes."""
        return self.ctx.local_map_push(laserCloudCornerStack, laserCloudSurfStack, transformTobeMapped, self.filter_corner,
                                       self.filter_surf)

    def processPointToLine(self, laserCloudCorner, m4d):
        """EST.h:159-165. Returns the FeatureLine records (12 doubles each, see mmloam_b200.h)."""
        feat, n, _, _ = self.ctx.associate(0, laserCloudCorner, m4d, self.thres_dist)
        return feat[feat[:, 10] >= 0]

    def processPointToPlanVec(self, laserCloudSurf, m4d):
        """EST.h:179-186. Returns (FeaturePlanVec records, is_degenerate)."""
        feat, n, M, nn = self.ctx.associate(1, laserCloudSurf, m4d, self.thres_dist)
        sv = -1.0
        if nn > 10:  # checkLocalizability, EST.cpp:536-565
            sv = float(np.sqrt(max(np.linalg.eigvalsh(M)[0], 0.0)))
        if sv < 2.0:
            self._fail_detected = True
        return feat[feat[:, 10] >= 0], sv < 3.0

    def EstimateLidarPose(self, cloud_xyzi_labelled, label, exTlb, P, q_wxyz, params=None):
        """EST.h:211-214 for a one-frame list: label split + voxel filter (EST.cpp:992-1026), Estimate
        (EST.cpp:1143-1581). Returns (P, q, stats)."""
        cloud = _f32(cloud_xyzi_labelled).reshape(-1, 4)
        corner = self.ctx.voxel_downsample(cloud[label == 1], self.filter_corner)
        surf = self.ctx.voxel_downsample(cloud[label == 2], self.filter_surf)
        P, q, stats = self.ctx.estimate(corner, surf, exTlb, P, q_wxyz, params)
        self._fail_detected = bool(stats[6])
        return P, q, stats

    def failureDetected(self):
        return self._fail_detected
