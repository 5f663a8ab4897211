#!/usr/bin/env python
"""Generate tests/golden/golden.npz + golden.json.

The reference (ROS + PCL + Eigen + Ceres, C++) cannot be built or imported in this image and
ships no golden vectors (SURVEY.md §4, §8 c), so these fixtures are produced by the CPU oracle on
seeded synthetic inputs. They pin the oracle against silent drift and give the GPU tests a
committed input/output pair that does not depend on regenerating the scene.
Run from the repo root:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

synth = ge.load_synth()
orc = ge.load_oracle()
orc.build()
here = os.path.dirname(os.path.abspath(__file__))

T_true = synth.make_T(synth.rot_z(-0.4), np.array([2.0, 1.5, 0.1]))
vx, vr, _ = synth.vlp16_scan(T_true, seed=2001, n_az=450)
hx, hl, _ = synth.horizon_scan(T_true, 6000, seed=2002)
x = np.concatenate([vx, hx])
line = np.concatenate([vr, hl + 16]).astype(np.uint16)
label = orc.extract_scan(x, line, 22)
corner_ds = orc.voxel_downsample(x[label == 1], 0.4)
surf_ds = orc.voxel_downsample(x[label == 2], 0.2)
ms, mc = synth.feature_map(20000, 2000, seed=2003)
m = orc.Map()
m.set(orc.SURF_LOCAL, ms)
m.set(orc.CORNER_LOCAL, mc)
T_wl = T_true @ synth.s1_offset_pose()
pf, nf, M, nn = m.associate_plane(surf_ds, T_wl, 10.0)
q0, _ = orc.so3_exp(synth.R_to_rotvec(T_wl[:3, :3]))
P, q, st = m.estimate(corner_ds, surf_ds, np.eye(4), T_wl[:3, 3], q0)
np.savez_compressed(os.path.join(here, "golden.npz"), scan_xyzi=x, scan_line=line, label=label, corner_ds=corner_ds,
                    surf_ds=surf_ds, map_surf=ms, map_corner=mc, T_wl=T_wl, plane_valid=pf[:, 10], P0=T_wl[:3, 3], q0=q0,
                    P_est=P, q_est=q)
json.dump({"n_lines": 22, "n_plane": int(nf), "n_sharp": int((label == 1).sum()), "n_flat": int((label == 2).sum()),
           "outer_iters": int(st[0]), "generator": "tests/golden/make_golden.py (CPU oracle; reference unbuildable here)"},
          open(os.path.join(here, "golden.json"), "w"), indent=1)
print("golden written", x.shape, int(nf), st[:7])
