"""CPU ORACLE for the map-maintenance rows SURVEY.md §8 (f) F1 marks "next" (TEST INFRASTRUCTURE ONLY).

Restates, on top of the C++ oracle's point transform (A5) and voxel filter (A6):

  Estimator::MapIncrementLocal   mm-loam/src/lio/Estimator.cpp:1585-1643
  MAP_MANAGER::MapIncrement      mm-loam/src/lio/Map_Manager.cpp:125-281  with MapMove, MM.cpp:288-581

Checked against the reference text itself (oracle/_ref, tests/test_ref_pin.py) and by tests/test_map_maintenance.py.
"""
from __future__ import annotations

import numpy as np

from . import oracle as orc

LOCAL_WINDOW = 50          # localMapWindowSize, include/Estimator/Estimator.h:326
CUBE_DOWNSAMPLE_OVER = 300  # MM.cpp:222: a touched cube is voxel-filtered when it holds more than 300 points
CUBE_W, CUBE_H, CUBE_D = 21, 11, 21  # laserCloudWidth / Height / Depth, include/MapManager/Map_Manager.h:117-119


def point_to_map(xyz, T):
    """MAP_MANAGER::pointAssociateToMap (MM.cpp:75-89) on every row: float64 products summed left to right, result
    rounded to float32 (vectorised form of oracle.point_to_map, checked against it in the tests)."""
    p = np.ascontiguousarray(xyz, np.float32).astype(np.float64)
    T = np.asarray(T, np.float64).reshape(4, 4)
    out = np.empty((p.shape[0], 3), np.float32)
    for r in range(3):
        out[:, r] = (((T[r, 0] * p[:, 0] + T[r, 1] * p[:, 1]) + T[r, 2] * p[:, 2]) + T[r, 3]).astype(np.float32)
    return out


def cube_index(xyz, cen=(10, 5, 10)):
    """Cube of every row (MM.cpp:161-173 = 583-605): int() truncates toward zero, negatives are corrected, cubes
    outside the 21 x 11 x 21 grid map to 5000 (vectorised form of oracle.cube_index)."""
    p = np.ascontiguousarray(xyz, np.float32).astype(np.float64)
    cen_w, cen_h, cen_d = cen
    ci = np.trunc((p[:, 0] + 25.0) / 50.0).astype(np.int64) + cen_d
    cj = np.trunc((p[:, 1] + 25.0) / 50.0).astype(np.int64) + cen_w
    ck = np.trunc((p[:, 2] + 25.0) / 50.0).astype(np.int64) + cen_h
    ci -= (p[:, 0] + 25.0 < 0)
    cj -= (p[:, 1] + 25.0 < 0)
    ck -= (p[:, 2] + 25.0 < 0)
    ok = (ci >= 0) & (ci < CUBE_D) & (cj >= 0) & (cj < CUBE_W) & (ck >= 0) & (ck < CUBE_H)
    idx = ci + CUBE_D * cj + CUBE_D * CUBE_W * ck  # MAP_MANAGER::ToIndex
    return np.where(ok, idx, 5000).astype(np.int64)


def _transform(xyzi, T):
    xyzi = np.ascontiguousarray(xyzi, np.float32)
    out = xyzi.copy()
    if xyzi.shape[0]:
        out[:, :3] = point_to_map(xyzi[:, :3], T)
    return out


class LocalMap:
    """Estimator::MapIncrementLocal for the corner and surf clouds (the non-feature cloud follows the same steps
    without the final filter, EST.cpp:1636-1639, and is not used by the window-1 path).

    MapIncrementLocal itself ADDS the 50 ring entries to the cloud it finds in laserCloud*FromLocal
    (EST.cpp:1620-1624). Its only caller, EstimateLidarPose, clears those clouds first (EST.cpp:1085-1087,
    1127-1129), so in the odometry loop the local map is voxel(concatenation of the ring): `clear_first=True`."""

    def __init__(self, leaf_corner=0.4, leaf_surf=0.2):
        self.leaf = (leaf_corner, leaf_surf)
        self.ring = [[np.zeros((0, 4), np.float32) for _ in range(LOCAL_WINDOW)] for _ in range(2)]
        self.from_local = [np.zeros((0, 4), np.float32), np.zeros((0, 4), np.float32)]
        self.local_map_id = 0

    def increment(self, corner_stack, surf_stack, T_wl, clear_first=False):
        if clear_first:                                              # EST.cpp:1085-1087 / 1127-1129
            self.from_local = [np.zeros((0, 4), np.float32), np.zeros((0, 4), np.float32)]
        slot = self.local_map_id % LOCAL_WINDOW                      # EST.cpp:1597
        for kind, stack in enumerate((corner_stack, surf_stack)):
            self.ring[kind][slot] = _transform(stack, T_wl)          # EST.cpp:1600-1612
            cloud = np.concatenate([self.from_local[kind]] + self.ring[kind])  # EST.cpp:1620-1624
            self.from_local[kind] = orc.voxel_downsample(cloud, self.leaf[kind])  # EST.cpp:1630-1635
        self.local_map_id += 1                                       # EST.cpp:1640
        return self.from_local[0], self.from_local[1]


class CubeMap:
    """MAP_MANAGER::MapIncrement for the corner and surf clouds, including MapMove. Input points are already in the
    world frame (MM.cpp:159: the stack is copied, not transformed); `T_wl` only feeds MapMove (MM.cpp:288-581).

    Both voxel filters of MAP_MANAGER have leaf 0.4 whatever the constructor is given (MM.cpp:56-58).
    `for_match` / `cen_last` are the snapshots Estimate() matches against: they are taken at the START of an update
    (MM.cpp:133-146), i.e. they lag the cubes by one MapIncrement.

    MapMove's loop bounds (`while (centerCube < 8)` ... `while (centerCube >= size - 8)`, MM.cpp:307-579) are restated
    literally: along the height axis (11 cubes) the two loops overlap, so every update shifts the cubes up until
    centre cube 8 and back down to 2 (CenHeight ends at 2 for a sensor near z = 0) and whatever that pushes over
    the top is dropped."""

    LEAF = 0.4

    def __init__(self, leaf_corner=0.4, leaf_surf=0.2, cen=(10, 5, 10)):
        self.cen = tuple(cen)              # (CenWidth, CenHeight, CenDepth)
        self.cen_last = tuple(cen)
        self.cubes = [dict(), dict()]      # (i, j, k) -> float32 [m, 4]
        self.for_match = [dict(), dict()]
        self.from_map = [np.zeros((0, 4), np.float32), np.zeros((0, 4), np.float32)]

    @staticmethod
    def _to_index(i, j, k):
        return i + CUBE_D * j + CUBE_D * CUBE_W * k       # MAP_MANAGER::ToIndex, MM.cpp:65-67

    def _shift(self, axis, step):
        """One pass of a MapMove while-loop body: every cube moves by `step` along `axis`, the layer that wraps
        around arrives empty."""
        size = (CUBE_D, CUBE_W, CUBE_H)[axis]
        for kind in range(2):
            moved = {}
            for key, pts in self.cubes[kind].items():
                k2 = list(key)
                k2[axis] += step
                if 0 <= k2[axis] < size and pts.shape[0]:
                    moved[tuple(k2)] = pts
            self.cubes[kind] = moved

    def map_move(self, T_wl):
        t = np.asarray(T_wl, np.float64).reshape(4, 4)[:3, 3]
        cen_w, cen_h, cen_d = self.cen
        c = [int(np.trunc((t[0] + 25.0) / 50.0)) + cen_d, int(np.trunc((t[1] + 25.0) / 50.0)) + cen_w,
             int(np.trunc((t[2] + 25.0) / 50.0)) + cen_h]                       # MM.cpp:299-301
        for a in range(3):
            if t[a] + 25.0 < 0:
                c[a] -= 1                                                        # MM.cpp:303-305
        cen = [cen_d, cen_w, cen_h]
        for a, size in enumerate((CUBE_D, CUBE_W, CUBE_H)):                      # I (depth), J (width), K (height)
            while c[a] < 8:
                self._shift(a, +1); c[a] += 1; cen[a] += 1
            while c[a] >= size - 8:
                self._shift(a, -1); c[a] -= 1; cen[a] -= 1
        self.cen = (cen[1], cen[2], cen[0])

    def increment(self, corner_w, surf_w, T_wl=None):
        # MM.cpp:133-146: snapshot for the matcher, before anything moves
        self.for_match = [dict(self.cubes[0]), dict(self.cubes[1])]
        self.cen_last = self.cen
        if T_wl is not None:
            self.map_move(T_wl)                                                  # MM.cpp:149
        cen_w, cen_h, cen_d = self.cen
        for kind, stack in enumerate((corner_w, surf_w)):
            stack = np.ascontiguousarray(stack, np.float32)
            touched = []
            if stack.shape[0]:
                idx = cube_index(stack[:, :3], self.cen)             # MM.cpp:161-173, 5000 = outside the grid
                for ci in np.unique(idx):
                    if ci == 5000:
                        continue
                    key = (int(ci) % CUBE_D, (int(ci) // CUBE_D) % CUBE_W, int(ci) // (CUBE_D * CUBE_W))
                    pts = stack[idx == ci]                           # input order inside a cube (push_back)
                    old = self.cubes[kind].get(key, np.zeros((0, 4), np.float32))
                    self.cubes[kind][key] = np.concatenate([old, pts])
                    touched.append(key)
            out = []
            for key in sorted(touched, key=lambda q: self._to_index(*q)):       # MM.cpp:219: cubes in index order
                if self.cubes[kind][key].shape[0] > CUBE_DOWNSAMPLE_OVER:        # MM.cpp:222
                    self.cubes[kind][key] = orc.voxel_downsample(self.cubes[kind][key], self.LEAF)
                out.append(self.cubes[kind][key])
            # MM.cpp:215-217, 233: only the cubes touched by this update form laserCloud*FromMap
            self.from_map[kind] = np.concatenate(out) if out else np.zeros((0, 4), np.float32)
        return self.from_map[0], self.from_map[1]

    def cloud(self, kind, matched=False):
        """All cubes of one kind in cube-index order: what the per-cube kd-trees jointly hold (the k-NN target).
        matched=True: the snapshot Estimate() sees (one update behind)."""
        src = self.for_match[kind] if matched else self.cubes[kind]
        keys = sorted(src, key=lambda q: self._to_index(*q))
        return np.concatenate([src[k] for k in keys]) if keys else np.zeros((0, 4), np.float32)
