"""CPU tests of the map-maintenance oracle (SURVEY.md §8 f, row F1: the "next" row after the hot path).
No product code implements F1 yet; these pin the behaviour its device-side version must reproduce."""
import numpy as np
import pytest

from oracle import map_maintenance as mmt


def _cloud(rng, n, lo, hi):
    p = np.zeros((n, 4), np.float32)
    p[:, :3] = rng.uniform(lo, hi, size=(n, 3)).astype(np.float32)
    p[:, 3] = rng.uniform(0, 255, size=n).astype(np.float32)
    return p


def _pose(synth, yaw, t):
    return synth.make_T(synth.rot_z(yaw), np.asarray(t, np.float64))


def test_vectorised_helpers_equal_the_scalar_oracle(orc):
    """point_to_map / cube_index: the numpy forms used here against the C++ oracle's scalar functions, including
    negative coordinates, cube faces and points outside the 21 x 11 x 21 grid."""
    rng = np.random.default_rng(1)
    pts = rng.uniform(-700, 700, size=(300, 3)).astype(np.float32)
    pts[:8] = [[-25.0, 0, 0], [25.0, 0, 0], [-25.000002, 0, 0], [24.999998, -75.0, 275.0], [0, 0, -275.0001], [524.9, 0, 0],
               [525.0, 0, 0], [-525.1, 0, 0]]
    idx = mmt.cube_index(pts)
    for i in range(pts.shape[0]):
        assert idx[i] == orc.cube_index(pts[i]), pts[i]
    assert (idx == 5000).any() and (idx != 5000).any()
    T = np.eye(4)
    T[:3, :3] = np.array([[0.36, 0.48, -0.8], [-0.8, 0.6, 0.0], [0.48, 0.64, 0.6]])
    T[:3, 3] = [1.5, -2.25, 0.125]
    w = mmt.point_to_map(pts, T)
    for i in range(0, pts.shape[0], 7):
        assert np.array_equal(w[i], orc.point_to_map(pts[i], T))


def test_local_map_first_increment_is_the_filtered_transformed_stack(orc, synth):
    rng = np.random.default_rng(2)
    corner, surf = _cloud(rng, 300, -5, 5), _cloud(rng, 3000, -8, 8)
    T = _pose(synth, 0.4, [2.0, -1.0, 0.3])
    lm = mmt.LocalMap()
    c, s = lm.increment(corner, surf, T)
    cw, sw = corner.copy(), surf.copy()
    cw[:, :3] = mmt.point_to_map(corner[:, :3], T)
    sw[:, :3] = mmt.point_to_map(surf[:, :3], T)
    assert np.array_equal(c, orc.voxel_downsample(cw, 0.4))
    assert np.array_equal(s, orc.voxel_downsample(sw, 0.2))
    assert lm.local_map_id == 1


def test_local_map_accumulates_the_previous_filtered_map(orc, synth):
    """EST.cpp:1620-1624 adds the ring to the previous result instead of rebuilding it: after two increments the map
    is voxel(voxel(A) + A + B), which differs from voxel(A + B) in the centroids of the cells A occupies."""
    rng = np.random.default_rng(3)
    A, B = _cloud(rng, 2000, -4, 4), _cloud(rng, 2000, -4, 4)
    T = np.eye(4)
    lm = mmt.LocalMap()
    lm.increment(A[:10], A, T)
    _, s2 = lm.increment(B[:10], B, T)
    expect = orc.voxel_downsample(np.concatenate([orc.voxel_downsample(A, 0.2), A, B]), 0.2)
    assert np.array_equal(s2, expect)
    rebuilt = orc.voxel_downsample(np.concatenate([A, B]), 0.2)
    assert rebuilt.shape == s2.shape  # same occupied voxels ...
    assert not np.array_equal(rebuilt, s2)  # ... but the accumulated centroids weigh the first scan's cells differently


def test_local_map_ring_evicts_after_fifty_frames(synth):
    """The 51st frame overwrites ring slot 0 (EST.cpp:1597-1602); evicted points stay only as filtered centroids."""
    lm = mmt.LocalMap()
    T = np.eye(4)
    for k in range(mmt.LOCAL_WINDOW + 1):
        pts = np.zeros((3, 4), np.float32)
        pts[:, 0] = 10.0 * k + np.arange(3)  # every frame in its own voxels
        lm.increment(pts, pts, T)
    assert lm.local_map_id == mmt.LOCAL_WINDOW + 1
    assert np.array_equal(lm.ring[1][0][:, 0], 10.0 * mmt.LOCAL_WINDOW + np.arange(3))  # slot 0 now holds frame 50
    xs = np.sort(lm.from_local[1][:, 0])
    assert xs.shape[0] == 3 * (mmt.LOCAL_WINDOW + 1) and xs[0] == 0.0  # frame 0 survives as centroids


def test_cube_map_bins_filters_and_reports_touched_cubes(orc):
    rng = np.random.default_rng(4)
    cm = mmt.CubeMap()
    near = _cloud(rng, 250, -20, 20)          # centre cube only, below the 300-point threshold
    c, s = cm.increment(near[:20], near)
    centre = (10, 10, 5)                      # (i, j, k) of the cube around the origin with the initial centre
    assert mmt.CubeMap._to_index(*centre) == int(mmt.cube_index(np.zeros((1, 3), np.float32))[0])
    assert set(cm.cubes[1]) == {centre} and np.array_equal(s, near)  # not filtered yet: order of arrival kept
    more = _cloud(rng, 200, -20, 20)
    far = _cloud(rng, 50, 30, 70)              # cube (+1, +1, +1)
    far[:, 2] = rng.uniform(30, 70, 50).astype(np.float32)
    out = _cloud(rng, 5, 600, 700)             # outside the grid: dropped (MM.cpp:168-175)
    _, s = cm.increment(near[:0], np.concatenate([more, far, out]))
    assert len(cm.cubes[1]) == 2
    filtered = orc.voxel_downsample(np.concatenate([near, more]), 0.4)   # MM.cpp:56-58: leaf 0.4 for every kind
    assert np.array_equal(cm.cubes[1][centre], filtered)            # 450 > 300: filtered in place
    other = [k for k in cm.cubes[1] if k != centre][0]
    assert np.array_equal(cm.cubes[1][other], far)                  # 50 points: kept as they arrived
    assert s.shape[0] == filtered.shape[0] + 50                     # FromMap = the cubes touched by this update
    _, s = cm.increment(near[:0], far[:3])
    assert s.shape[0] == 53 and cm.cloud(1).shape[0] == filtered.shape[0] + 53  # only the far cube was touched
