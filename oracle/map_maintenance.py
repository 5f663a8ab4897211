"""CPU ORACLE for the map-maintenance rows SURVEY.md §8 (f) F1 marks "next" (TEST INFRASTRUCTURE ONLY).

Restates, on top of the C++ oracle's point transform (A5) and voxel filter (A6):

  Estimator::MapIncrementLocal   mm-loam/src/lio/Estimator.cpp:1585-1643
  MAP_MANAGER::MapIncrement      mm-loam/src/lio/Map_Manager.cpp:125-281  (without MapMove, MM.cpp:307-579:
                                 the cube centre stays at (CenWidth, CenHeight, CenDepth) = (10, 5, 10))

No product code implements these rows yet: this module and tests/test_map_maintenance.py pin the behaviour the
device-side version will have to reproduce. Parity unpinned like the rest of the oracle (no reference tests).
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

    Note the reference's accumulation (EST.cpp:1620-1624): the 50 ring entries are ADDED to the previous filtered map
    (laserCloud*FromLocal is never cleared), so a point leaves the local map only when the voxel filter merges it."""

    def __init__(self, leaf_corner=0.4, leaf_surf=0.2):
        self.leaf = (leaf_corner, leaf_surf)
        self.ring = [[np.zeros((0, 4), np.float32) for _ in range(LOCAL_WINDOW)] for _ in range(2)]
        self.from_local = [np.zeros((0, 4), np.float32), np.zeros((0, 4), np.float32)]
        self.local_map_id = 0

    def increment(self, corner_stack, surf_stack, T_wl):
        slot = self.local_map_id % LOCAL_WINDOW                      # EST.cpp:1597
        for kind, stack in enumerate((corner_stack, surf_stack)):
            self.ring[kind][slot] = _transform(stack, T_wl)          # EST.cpp:1600-1612
            cloud = np.concatenate([self.from_local[kind]] + self.ring[kind])  # EST.cpp:1620-1624
            self.from_local[kind] = orc.voxel_downsample(cloud, self.leaf[kind])  # EST.cpp:1630-1635
        self.local_map_id += 1                                       # EST.cpp:1640
        return self.from_local[0], self.from_local[1]


class CubeMap:
    """MAP_MANAGER::MapIncrement for the corner and surf clouds. Input points are already in the world frame
    (MM.cpp:159: the stack is copied, not transformed)."""

    def __init__(self, leaf_corner=0.4, leaf_surf=0.2, cen=(10, 5, 10)):
        self.leaf = (leaf_corner, leaf_surf)
        self.cen = tuple(cen)
        self.cubes = [dict(), dict()]      # cube index -> float32 [m, 4]
        self.from_map = [np.zeros((0, 4), np.float32), np.zeros((0, 4), np.float32)]

    def increment(self, corner_w, surf_w):
        for kind, stack in enumerate((corner_w, surf_w)):
            stack = np.ascontiguousarray(stack, np.float32)
            touched = []
            if stack.shape[0]:
                idx = cube_index(stack[:, :3], self.cen)             # MM.cpp:161-173, 5000 = outside the grid
                for ci in np.unique(idx):
                    if ci == 5000:
                        continue
                    pts = stack[idx == ci]                           # input order inside a cube (push_back)
                    old = self.cubes[kind].get(int(ci), np.zeros((0, 4), np.float32))
                    self.cubes[kind][int(ci)] = np.concatenate([old, pts])
                    touched.append(int(ci))
            out = []
            for ci in sorted(touched):                               # MM.cpp:219: cubes in index order
                if self.cubes[kind][ci].shape[0] > CUBE_DOWNSAMPLE_OVER:   # MM.cpp:222
                    self.cubes[kind][ci] = orc.voxel_downsample(self.cubes[kind][ci], self.leaf[kind])
                out.append(self.cubes[kind][ci])
            # MM.cpp:215-217, 233: only the cubes touched by this update form laserCloud*FromMap
            self.from_map[kind] = np.concatenate(out) if out else np.zeros((0, 4), np.float32)
        return self.from_map[0], self.from_map[1]

    def cloud(self, kind):
        """All cubes of one kind in cube-index order: what the per-cube kd-trees jointly hold (the k-NN target)."""
        keys = sorted(self.cubes[kind])
        return np.concatenate([self.cubes[kind][k] for k in keys]) if keys else np.zeros((0, 4), np.float32)
