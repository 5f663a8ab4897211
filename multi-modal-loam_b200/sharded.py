"""Multi-GPU scan matching: the global feature map sharded by 50 m cube (SURVEY.md §8 e).

The reference already cuts the global map into 50 m cubes and lets a query see only its own cube
(mm-loam/src/lio/Map_Manager.cpp:583-605, Estimator.cpp:192-199). That is the shard boundary:
  * rank r owns the cubes whose index (in sorted order of the populated cubes) is r modulo N and
    uploads only their points;
  * the (small) query set is replicated; a query whose cube lives elsewhere simply finds no map
    on this rank, so every query is matched on exactly one rank — no all-to-all;
  * the only exchange is the sum of the per-rank partial normal equations: 28 doubles
    [cost, g(6), upper H(21)] per evaluation (plus counts / normal moments per association),
    one small all-reduce over NCCL / NVLink, after which every rank takes the same dogleg step
    with the host-side solver (mml_solver_*), keeping poses bit-identical across ranks.

Two drivers of the same partition:
  * `connect_ranks` + `Context.estimate_sharded` (C-ABI mml_estimate_sharded): the exchange goes through PEER MEMORY from
    inside the kernels (the last CTA of an evaluation stores its sums into every rank's buffer over NVLink, waits for the
    others and sums in rank order) and the dogleg step stays on the device - no NCCL or host call between evaluations;
    `torch.distributed` only carries the 64-byte IPC handles once at set-up;
  * `ShardedEstimator`: the same loop driven from the host with a `torch.distributed` all-reduce per evaluation (NCCL on
    GPUs, gloo in the CPU tests), compute backend injected so the host logic can be tested without a GPU.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import load_library

CUBE_W, CUBE_H, CUBE_D = 21, 11, 21  # Map_Manager.h:117-119


def cube_index(xyz, cen=(10, 5, 10)):
    """Vectorised MAP_MANAGER::FindUsed*Map (Map_Manager.cpp:583-605); 5000 = outside the grid."""
    p = np.asarray(xyz, dtype=np.float32)[:, :3].astype(np.float64)
    c = np.trunc((p + 25.0) / 50.0).astype(np.int64)
    c -= (p + 25.0 < 0)
    cI, cJ, cK = c[:, 0] + cen[2], c[:, 1] + cen[0], c[:, 2] + cen[1]
    ok = (cI >= 0) & (cI < CUBE_D) & (cJ >= 0) & (cJ < CUBE_W) & (cK >= 0) & (cK < CUBE_H)
    return np.where(ok, cI + CUBE_D * cJ + CUBE_D * CUBE_W * cK, 5000)


def cube_owner(populated_cubes, world):
    """cube id -> owning rank: round-robin over the sorted populated cubes."""
    cubes = np.unique(np.asarray(populated_cubes))
    cubes = cubes[cubes != 5000]
    return {int(c): i % world for i, c in enumerate(cubes)}


def shard_points(xyzi, rank, world, owner=None, cen=(10, 5, 10)):
    """The points of `xyzi` that rank `rank` holds, and the cube->owner table used."""
    ids = cube_index(xyzi, cen)
    if owner is None:
        owner = cube_owner(ids, world)
    own = np.array([owner.get(int(c), -1) for c in ids]) == rank
    return np.ascontiguousarray(xyzi[own]), owner


class HostSolver:
    """ctypes view of mml_solver_* (the dogleg state machine, host side, no GPU needed)."""

    def __init__(self):
        self.lib = load_library()
        self.h = C.c_void_p()
        if self.lib.mml_solver_create(C.byref(self.h)) != 0:
            raise RuntimeError("mml_solver_create failed")

    def __del__(self):
        try:
            if self.h:
                self.lib.mml_solver_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def begin(self, x6, max_inner=10):
        x6 = np.ascontiguousarray(x6, np.float64)
        self.lib.mml_solver_begin(self.h, x6.ctypes.data_as(C.c_void_p), int(max_inner))

    def feed(self, out28):
        out28 = np.ascontiguousarray(out28, np.float64)
        nxt = np.zeros(6)
        done = C.c_int(0)
        self.lib.mml_solver_feed(self.h, out28.ctypes.data_as(C.c_void_p), nxt.ctypes.data_as(C.c_void_p), C.byref(done))
        return nxt, bool(done.value)

    def result(self):
        x = np.zeros(6)
        cost = C.c_double(0)
        it = C.c_int(0)
        self.lib.mml_solver_result(self.h, x.ctypes.data_as(C.c_void_p), C.byref(cost), C.byref(it))
        return x, cost.value, it.value


def pack28(H, g, cost):
    out = np.zeros(28)
    out[0] = cost
    out[1:7] = g
    k = 7
    for i in range(6):
        for j in range(i, 6):
            out[k] = H[i, j]
            k += 1
    return out


def _so3_exp_R(phi):
    th = np.linalg.norm(phi)
    K = np.array([[0, -phi[2], phi[1]], [phi[2], 0, -phi[0]], [-phi[1], phi[0], 0]])
    if th < 1e-12:
        return np.eye(3) + K
    return np.eye(3) + np.sin(th) / th * K + (1 - np.cos(th)) / th**2 * K @ K


class ShardedEstimator:
    """Estimator::Estimate (window size 1) over a cube-sharded global map.

    backend must provide, for THIS rank's shard:
      associate(T_wl, thres) -> (n_line, n_plane, moment 3x3, n_normals)
      accumulate_partial(x6) -> 28 doubles [cost, g, upper H]   (numpy array or torch tensor)
    allreduce(t) sums a torch tensor over the ranks in place (torch.distributed.all_reduce).
    """

    def __init__(self, backend, allreduce, max_outer=5, max_inner=10, thres=(25.0, 10.0, 1.0)):
        self.backend = backend
        self.allreduce = allreduce
        self.max_outer, self.max_inner, self.thres = max_outer, max_inner, thres
        self.solver = HostSolver()
        self.n_allreduce = 0

    def _reduce(self, arr):
        import torch

        t = arr if isinstance(arr, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(arr, np.float64))
        self.allreduce(t)
        self.n_allreduce += 1
        return t.detach().cpu().numpy().astype(np.float64)

    def estimate(self, x6):
        """x6 = [t_wb, phi_wb] (extrinsic = identity). Returns (x6, stats)."""
        x = np.asarray(x6, np.float64).copy()
        stats = {}
        for it in range(self.max_outer):
            T = np.eye(4)
            T[:3, :3] = _so3_exp_R(x[3:])
            T[:3, 3] = x[:3]
            nl, npl, M, nn = self.backend.associate(T, self.thres[min(it, 2)])
            red = self._reduce(np.concatenate([[nl, npl, nn], np.asarray(M, np.float64).reshape(9)]))
            stats = {"n_line": int(round(red[0])), "n_plane": int(round(red[1])), "outer": it + 1}
            x_before = x.copy()
            self.solver.begin(x, self.max_inner)
            x_eval, done = x.copy(), False
            while not done:
                out28 = self._reduce(self.backend.accumulate_partial(x_eval))
                x_eval, done = self.solver.feed(out28)
            x, cost, _ = self.solver.result()
            stats["cost"] = cost
            dR = _so3_exp_R(x_before[3:]).T @ _so3_exp_R(x[3:])
            ang = np.degrees(np.arccos(np.clip((np.trace(dR) - 1) / 2, -1, 1)))
            if (ang < 0.05 and np.linalg.norm(x[:3] - x_before[:3]) < 0.05) or it + 1 == self.max_outer:
                break
        return x, stats


class GpuShardBackend:
    """This rank's shard on its GPU: resident map + replicated queries in a mmloam Context."""

    def __init__(self, ctx, device_index):
        self.ctx = ctx
        self.dev = device_index

    def associate(self, T_wl, thres):
        return self.ctx.frame_associate(T_wl, thres)

    def accumulate_partial(self, x6):
        """Launch the evaluation and expose the 28 device doubles as a torch tensor (zero copy),
        so the NCCL all-reduce reads them straight from HBM on the context's stream."""
        import torch

        ptr = C.c_void_p()
        x6 = np.ascontiguousarray(x6, np.float64)
        T = np.eye(4).reshape(16)
        self.ctx._ck(self.ctx.lib.mml_frame_accumulate_partial_dev(self.ctx.h, x6.ctypes.data_as(C.c_void_p),
                                                                   T.ctypes.data_as(C.c_void_p), C.c_double(0.0),
                                                                   C.c_double(0.1 / 1.5e-3), C.byref(ptr)))

        class _Dev:
            __cuda_array_interface__ = {"shape": (28,), "typestr": "<f8", "data": (ptr.value, False), "version": 2}

        return torch.as_tensor(_Dev(), device=f"cuda:{self.dev}")


def nccl_allreduce_on(ctx):
    """all-reduce ordered after the context's kernels: torch's current stream := the context's stream."""
    import torch
    import torch.distributed as dist

    ext = torch.cuda.ExternalStream(ctx.lib.mml_stream_handle(ctx.h))

    def fn(t):
        if t.is_cuda:
            with torch.cuda.stream(ext):
                dist.all_reduce(t)
            ext.synchronize()
        else:
            tc = t.cuda()
            dist.all_reduce(tc)
            t.copy_(tc.cpu())

    return fn



def cube_owner_union(clouds, world, cen=(10, 5, 10)):
    """cube id -> owning rank over the cubes populated by ANY of the clouds (corner and surf maps share the table: a
    corner cube without surf points must still have an owner)."""
    ids = np.concatenate([cube_index(c, cen) for c in clouds if len(c)]) if any(len(c) for c in clouds) else np.zeros(0, np.int64)
    return cube_owner(ids, world)


def connect_ranks(ctx, rank, world):
    """One process per GPU: all-gather the exchange buffers' IPC handles over torch.distributed and open the peers."""
    import torch.distributed as dist

    mine = ctx.shard_init(rank, world)
    if world == 1:
        return
    handles = [None] * world
    dist.all_gather_object(handles, mine)
    ctx.shard_connect_ipc(handles)
    dist.barrier()
