"""CPU ORACLE (test infrastructure only) for SURVEY.md §8 (f) F2, the message stages either side of the extractor:

  getHoriFeatureExtract's CustomMsg loop      mm-loam/src/unionFeatureExtract.cpp:985-998
  pcl::fromROSMsg + removeNaNFromPointCloud   FE.cpp:1129-1133
  removeNearPointCloud / removeNearFarPoints  mm-loam/include/lidars_extrinsic_cali.h:424-477
  the clouds of union_cloud.msg               FE.cpp:916-937 (Horizon), 1263-1297 (VLP-16)

CustomPoint (livox_ros_driver/msg/CustomPoint.msg:3-9) serialises to 19 packed little-endian bytes."""
from __future__ import annotations

import numpy as np

CUSTOM_POINT = np.dtype([("offset_time", "<u4"), ("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("reflectivity", "u1"), ("tag", "u1"),
                         ("line", "u1")])
assert CUSTOM_POINT.itemsize == 19


def pack_custom_points(offset_time, xyz, reflectivity, line, tag=0):
    """The serialised `points` array of a CustomMsg (test input)."""
    rec = np.zeros(len(offset_time), CUSTOM_POINT)
    rec["offset_time"] = offset_time
    rec["x"], rec["y"], rec["z"] = xyz[:, 0], xyz[:, 1], xyz[:, 2]
    rec["reflectivity"] = reflectivity
    rec["tag"] = tag
    rec["line"] = line
    return rec.view(np.uint8).reshape(-1)


def _to_sec(t):
    t = np.asarray(t, np.uint64)
    return (t // np.uint64(1_000_000_000)).astype(np.float64) + 1e-9 * (t % np.uint64(1_000_000_000)).astype(np.float64)


def unpack_custom_points(raw, used_line=6):
    rec = np.frombuffer(np.ascontiguousarray(raw, np.uint8).tobytes(), CUSTOM_POINT)
    if rec.size == 0:
        return np.zeros((0, 4), np.float32), np.zeros(0, np.uint16), np.zeros(0, np.float32)
    span = _to_sec(rec["offset_time"][-1])                                    # FE.cpp:985
    keep = ~(rec["line"].astype(np.int64) > used_line - 1) & ~(rec["x"].astype(np.float64) < 0.01)   # FE.cpp:988-989
    k = rec[keep]
    xyzi = np.stack([k["x"], k["y"], k["z"], k["reflectivity"].astype(np.float32)], 1).astype(np.float32)
    s = (_to_sec(k["offset_time"]) / span).astype(np.float32)                 # FE.cpp:994
    return xyzi, k["line"].astype(np.uint16), s


def unpack_pointcloud2(raw, point_step, off_x=0, off_y=4, off_z=8, off_intensity=12):
    b = np.ascontiguousarray(raw, np.uint8).reshape(-1, point_step)

    def f(off):
        return b[:, off:off + 4].copy().view("<f4").reshape(-1)

    x, y, z = f(off_x), f(off_y), f(off_z)
    i = f(off_intensity) if off_intensity >= 0 else np.zeros_like(x)
    ok = np.isfinite(x) & np.isfinite(y) & np.isfinite(z)
    return np.stack([x, y, z, i], 1)[ok].astype(np.float32)


def _records(xyzi, s, line, label, intensity):
    out = np.zeros((xyzi.shape[0], 12), np.float32)
    out[:, 0:3] = xyzi[:, :3]
    out[:, 3] = 1.0
    out[:, 4] = s
    out[:, 5] = line
    out[:, 6] = label
    out[:, 8] = intensity
    return out


def pack_union_clouds(xyzi, s, line, label, near_full, far_full, near_feat, far_feat, zero_full_intensity=False):
    xyzi = np.ascontiguousarray(xyzi, np.float32)
    x, y, z = xyzi[:, 0], xyzi[:, 1], xyzi[:, 2]
    dis = (x * x + y * y) + z * z                                             # float32, lidars_extrinsic_cali.h:463-465

    def cut(near, far):
        drop = dis < np.float32(near) * np.float32(near)
        if far > 0:
            drop |= dis > np.float32(far) * np.float32(far)
        return ~drop

    rec = _records(xyzi, np.asarray(s, np.float32), np.asarray(line, np.float32), np.asarray(label, np.float32), xyzi[:, 3])
    full = rec[cut(near_full, far_full)].copy()
    if zero_full_intensity:
        full[:, 8] = 0.0
    feat = cut(near_feat, far_feat)
    lab = np.asarray(label)
    return full, rec[feat & (lab == 1)], rec[feat & (lab == 2)]
