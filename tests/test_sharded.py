"""Multi-GPU host logic on CPU: world_size-2 gloo run of the cube-sharded Estimate loop.

The compute backend here is the oracle (tests may use it); the code under test is the product's
sharding + all-reduce + host dogleg (multi-modal-loam_b200/sharded.py, mml_solver_*)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _scene():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    synth, orc = ge.load_synth(), ge.load_oracle()
    shift = np.array([27.0, 0.0, 0.0])  # the room straddles the cube boundary at x = 25 m
    T_true = synth.make_T(synth.rot_z(0.3), np.array([-3.0, -1.0, 0.2]) + shift)
    T0 = synth.make_T(synth.rot_z(0.3), np.array([-3.0, -1.0, 0.2]))
    x, ring, _ = synth.vlp16_scan(T0, seed=1001)
    lab = orc.extract_scan(x, ring, 16)
    corner = orc.voxel_downsample(x[lab == 1], 0.4)
    surf = orc.voxel_downsample(x[lab == 2], 0.2)
    ms, mc = synth.feature_map(60_000, 4_000, seed=1002)
    ms[:, :3] += shift.astype(np.float32)
    mc[:, :3] += shift.astype(np.float32)
    T_init = T_true @ synth.s1_offset_pose()
    x6 = np.concatenate([T_init[:3, 3], synth.R_to_rotvec(T_init[:3, :3])])
    return ge, synth, orc, corner, surf, ms, mc, x6, T_true


class OracleShardBackend:
    def __init__(self, orc, corner, surf, ms, mc):
        self.orc, self.corner, self.surf = orc, corner, surf
        self.map = orc.Map()
        self.map.set(orc.SURF_GLOBAL, ms)
        self.map.set(orc.CORNER_GLOBAL, mc)
        self.lf = self.pf = None

    def associate(self, T, thres):
        self.lf, nl = self.map.associate_line(self.corner, T, thres)
        self.pf, npl, M, nn = self.map.associate_plane(self.surf, T, thres)
        return nl, npl, M, nn

    def accumulate_partial(self, x6):
        from mmloam_b200.sharded import pack28
        H, g, c = self.orc.accumulate(self.lf, self.pf, x6, np.eye(4))
        return pack28(H, g, c)


def _worker(rank, world, port, out_path):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ge, synth, orc, corner, surf, ms, mc, x6, T_true = _scene()
    ge.load_package()
    from mmloam_b200 import sharded
    ms_r, owner = sharded.shard_points(ms, rank, world)
    mc_r, _ = sharded.shard_points(mc, rank, world, owner=sharded.cube_owner(sharded.cube_index(ms), world))
    backend = OracleShardBackend(orc, corner, surf, ms_r, mc_r)
    est = sharded.ShardedEstimator(backend, lambda t: dist.all_reduce(t))
    x, stats = est.estimate(x6)
    np.savez(out_path + f".{rank}.npz", x=x, n_line=stats["n_line"], n_plane=stats["n_plane"], n_own=ms_r.shape[0],
             n_allreduce=est.n_allreduce)
    dist.barrier()
    dist.destroy_process_group()


def test_cube_index_matches_oracle():
    ge, synth, orc, *_ = _scene()
    ge.load_package()
    from mmloam_b200 import sharded
    rng = np.random.default_rng(0)
    p = rng.uniform(-600, 600, (2000, 3)).astype(np.float32)
    p[:50] = np.array([[25.0, -25.0, 24.999]], np.float32) + rng.normal(0, 1e-4, (50, 3)).astype(np.float32)
    ids = sharded.cube_index(p)
    for i in range(0, 2000, 7):
        assert ids[i] == orc.cube_index(p[i])
    owner = sharded.cube_owner(ids, 4)
    assert set(owner.values()) <= {0, 1, 2, 3} and 5000 not in owner


def test_host_solver_matches_oracle_dogleg():
    """mml_solver_* (the code the device loop runs, compiled for the host) vs the oracle's dogleg."""
    ge, synth, orc, corner, surf, ms, mc, x6, T_true = _scene()
    ge.load_package()
    from mmloam_b200 import sharded
    m = orc.Map()
    m.set(orc.SURF_GLOBAL, ms)
    m.set(orc.CORNER_GLOBAL, mc)
    T = synth.make_T(synth.rotvec_to_R(x6[3:]), x6[:3])
    lf, _ = m.associate_line(corner, T, 25.0)
    pf, _, _, _ = m.associate_plane(surf, T, 25.0)
    s = sharded.HostSolver()
    s.begin(x6, 10)
    xe, done, n_eval = x6.copy(), False, 0
    while not done:
        H, g, c = orc.accumulate(lf, pf, xe, np.eye(4))
        xe, done = s.feed(sharded.pack28(H, g, c))
        n_eval += 1
    xs, cost, iters = s.result()
    q0, _ = orc.so3_exp(x6[3:])
    P, q, st = m.estimate(corner, surf, np.eye(4), x6[:3], q0, orc.est_params(max_outer=1))
    assert np.abs(xs[:3] - P).max() < 1e-9 and np.abs(xs[3:] - orc.so3_log(q)).max() < 1e-9
    assert iters == int(st[1]) and n_eval <= 11


def test_sharded_estimate_world2_gloo(tmp_path):
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() % 2000)
    out = str(tmp_path / "res")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    r0, r1 = np.load(out + ".0.npz"), np.load(out + ".1.npz")
    # every rank takes the same steps: bit-identical poses
    assert np.array_equal(r0["x"], r1["x"])
    assert r0["n_own"] > 0 and r1["n_own"] > 0 and r0["n_own"] + r1["n_own"] == 60_000
    assert int(r0["n_allreduce"]) == int(r1["n_allreduce"]) > 2
    # and equal to the unsharded solve on the whole map
    ge, synth, orc, corner, surf, ms, mc, x6, T_true = _scene()
    m = orc.Map()
    m.set(orc.SURF_GLOBAL, ms)
    m.set(orc.CORNER_GLOBAL, mc)
    q0, _ = orc.so3_exp(x6[3:])
    P, q, st = m.estimate(corner, surf, np.eye(4), x6[:3], q0)
    assert np.abs(r0["x"][:3] - P).max() < 1e-7 and np.abs(r0["x"][3:] - orc.so3_log(q)).max() < 1e-7
    assert int(r0["n_line"]) == int(st[2]) and int(r0["n_plane"]) == int(st[3])
    assert np.abs(r0["x"][:3] - T_true[:3, 3]).max() < 0.01
