"""CPU ORACLE (test infrastructure only): the odometry loop with an IMU-initialised sliding window, i.e. the per-scan
body of process() in its LidarIMUInited branch (mm-loam/src/unionPoseEstimation.cpp:796-891) on top of the C++
oracle's pieces. Used by tests/ and by bench.py's cpu_baseline / --impl reference legs for BASELINE config 3."""
from __future__ import annotations

import numpy as np

from . import oracle as orc


def _pose(st):
    _, R = orc.so3_exp(orc.so3_log(st[3:7]))
    T = np.eye(4)
    # rotation straight from the quaternion (Eigen toRotationMatrix), not through log/exp
    w, x, y, z = st[3:7]
    T[:3, :3] = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                          [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                          [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])
    T[:3, 3] = st[:3]
    return T


def _twl(st, T_bl):
    return _pose(st) @ T_bl


def run(omap, scans, n_lines, window, stamps, stamp0, imu, state0, exTlb=np.eye(4), gravity=(0, 0, -9.805),
        leaf_corner=0.4, leaf_surf=0.2, params=None, threads=1, labels=None, local_map=None):
    """scans: list of (xyzi, line, s). imu: list of (t, gyr, acc) per scan. Returns dict(poses_front, poses_newest,
    states, stats). labels (optional): precomputed label arrays (the extraction node's output).
    local_map (optional, map_maintenance.LocalMap): the map update of EstimateLidarPose runs after every solve
    (EST.cpp:1041-1135, lidarMode 2): when the solve is not degenerate and the sensor moved by >= sqrt(0.5) m since the
    last update, the OLDEST frame's clouds go through MapIncrementLocal (FromLocal cleared first) and the local maps
    of `omap` are replaced."""
    ex = np.asarray(exTlb, float).reshape(4, 4)
    T_bl = np.linalg.inv(ex)            # exRbl = R^T, exPbl = -R^T t (PE.cpp:1456-1459)
    prev = np.asarray(state0, float).copy()
    t_prev = float(stamp0)
    win_states, win_pre, win_c, win_s = [], [], [], []
    out_f, out_n, out_s, out_st = [], [], [], []
    prm = params if params is not None else orc.est_params(threads=threads)
    last_update = np.array([-1.0, -1.0, -1.0])   # last_velo_update_pose, Estimator.h:339
    win_sharp = []
    updates = 0
    for k, (xyzi, line, s) in enumerate(scans):
        pre = orc.Preint(*imu[k], t_prev, prev[10:13], prev[13:16])          # PE.cpp:807-809
        nxt = orc.imu_predict(prev, pre)                                      # PE.cpp:811-820
        dT = np.linalg.inv(_pose(prev) @ T_bl) @ (_pose(nxt) @ T_bl)          # PE.cpp:822-829
        lab = labels[k] if labels is not None else orc.extract_scan(xyzi, line, n_lines, threads=threads)
        und = orc.undistort(xyzi, s, dT[:3, :3], dT[:3, 3])                   # PE.cpp:862
        corner = orc.voxel_downsample(und[lab == 1], leaf_corner)            # EST.cpp:992-1026
        surf = orc.voxel_downsample(und[lab == 2], leaf_surf)
        if len(win_states) >= window:                                         # PE.cpp:830-832
            win_states.pop(0); win_pre.pop(0); win_c.pop(0); win_s.pop(0); win_sharp.pop(0)
        win_states.append(nxt); win_pre.append(pre); win_c.append(corner); win_s.append(surf)
        win_sharp.append(int((lab == 1).sum()))
        pres = [None] + win_pre[1:]
        st, stats = orc.estimate_window(omap, win_c, win_s, ex, np.array(win_states), pres, gravity, prm)
        win_states = [st[f].copy() for f in range(len(win_states))]
        prev = win_states[-1].copy()
        t_prev = float(stamps[k])
        out_f.append(_pose(win_states[0])); out_n.append(_pose(prev)); out_s.append(prev.copy()); out_st.append(stats[:8].copy())
        if local_map is not None:
            if sum(win_sharp) > 50:                                           # EST.cpp:1048-1052
                T_map = _twl(win_states[0], T_bl)
            else:                                                             # EST.cpp:1053-1065
                T_map = _twl(nxt, T_bl)
                T_map[0, 3], T_map[1, 3] = win_states[0][0], win_states[0][1]
            if stats[6] == 0:                                                 # EST.cpp:1069
                d = last_update - T_map[:3, 3]
                if np.float32(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]) >= np.float32(0.5):   # EST.cpp:1123-1125
                    mc, ms = local_map.increment(win_c[0], win_s[0], T_map, clear_first=True)
                    omap.set(orc.CORNER_LOCAL, mc)
                    omap.set(orc.SURF_LOCAL, ms)
                    last_update = T_map[:3, 3].copy()
                    updates += 1
    return dict(map_updates=updates, poses_front=np.array(out_f), poses_newest=np.array(out_n), states=np.array(out_s), stats=np.array(out_st))
