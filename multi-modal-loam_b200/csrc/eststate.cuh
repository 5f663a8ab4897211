// Device-side solver / odometry state shared by the Estimate kernels (accumulate.cu) and the fused
// split + undistort + voxel kernel (splitvoxel.cu), which starts a scan's solve in the chained odometry loop.
#pragma once
#include "smallmath.cuh"

namespace mml {

struct EstState {
  // pose of the body frame (EST.h:33-56) and extrinsics
  double P[3], Q[4];
  double Rbl[9], Pbl[3];
  double T_wl[16];
  float thres;
  int n_line, n_plane;
  // control
  int done_outer, done_inner, outer_it, inner_it, first, total_inner, is_degenerate, outer_next;
  int max_outer, max_inner;
  double lidar_m, w_tan, huber_a, thres_sched[3];
  // trust-region state (Ceres 2.1 TrustRegionMinimizer + DoglegStrategy)
  double x[6], x_cand[6], x_best[6];
  double cost, min_cost, H[36], g[6], scale[6];
  double radius, mu, alpha, dogleg_norm, model_change, step_norm, x_norm;
  double diag[6], idiag[6], grad[6], gn[6];
  int reuse, num_invalid;
  double q_before[4], t_before[3];
  double min_sv, final_cost;
};

// Parameters of one solve (kernel argument of k_est_init / the chained k_split_voxel).
struct EstInit {
  double P[3], Q[4], Rbl[9], Pbl[3];
  double lidar_m, w_tan, huber_a, thres_sched[3];
  int max_outer, max_inner;
};

// Cube-sharded global map over several GPUs (SURVEY.md 8 e): every rank evaluates the queries that fall into its cubes and
// the ranks exchange their partial sums THROUGH PEER MEMORY from inside the kernel that formed them: the last CTA of an
// evaluation stores its 28 sums into every rank's exchange buffer (NVLink / NVSwitch peer stores), publishes a sequence
// word behind a system-scope fence, waits for the other ranks' words in its own buffer, and sums the contributions in
// rank order - every rank gets the same bits, takes the same dogleg step on the device, and no host or library call
// sits between two evaluations.
constexpr int kShardMaxWorld = 8;
constexpr int kShardSlot = 40;   // doubles per message
constexpr int kShardBufDoubles = 2 * kShardMaxWorld * kShardSlot + 2 * kShardMaxWorld;  // slots[2][8][40] + sequence words [2][8]
struct ShardDev {
  int rank, world;
  unsigned seq;                   // exchanges completed (the same number on every rank)
  int pad;                        // set to 1 when a wait for a peer timed out
  double* peer[kShardMaxWorld];   // exchange buffers of all ranks, peer[rank] = this rank's own
};

// Executed by one full warp. local / total: `count` (<= kShardSlot) doubles in shared or global memory.
__device__ inline void shard_allreduce(ShardDev* sh, const double* local, int count, double* total) {
  const int lane = threadIdx.x & 31;
  const int rank = sh->rank, world = sh->world;
  const unsigned seq = sh->seq + 1u;
  const int par = (int)(seq & 1u);
  __syncwarp();
  for (int p = 0; p < world; p++) {
    double* dst = sh->peer[p] + (size_t)(par * kShardMaxWorld + rank) * kShardSlot;
    for (int k = lane; k < count; k += 32) dst[k] = local[k];
  }
  __threadfence_system();
  __syncwarp();
  if (lane < world) {
    volatile unsigned long long* flag =
        reinterpret_cast<volatile unsigned long long*>(sh->peer[lane] + 2 * kShardMaxWorld * kShardSlot) + par * kShardMaxWorld + rank;
    *flag = (unsigned long long)seq;
  }
  double* own = sh->peer[rank];
  if (lane < world) {
    volatile unsigned long long* flag =
        reinterpret_cast<volatile unsigned long long*>(own + 2 * kShardMaxWorld * kShardSlot) + par * kShardMaxWorld + lane;
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while (*flag != (unsigned long long)seq) {
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > 2000000000ull) { sh->pad = 1; break; }  // 2 s: a peer never arrived (it failed) - give up instead of hanging the device
    }
  }
  __threadfence_system();
  __syncwarp();
  for (int k = lane; k < count; k += 32) {
    double v = 0;
    for (int r = 0; r < world; r++) v += reinterpret_cast<volatile double*>(own)[(size_t)(par * kShardMaxWorld + r) * kShardSlot + k];
    total[k] = v;
  }
  __syncwarp();
  if (lane == 0) sh->seq = seq;
}

// Chained odometry loop (odometry.cu): the last two poses live on the device, so that the constant-velocity
// prediction of scan k+1 (PE.cpp:847-852, 882-890) needs no host round trip after scan k.
struct OdomDev {
  double T_last[16], T_before[16];
  int scan;     // index of the scan being matched
  int pad[3];
};
// per-scan outputs of the chained loop
struct ChainOut {
  double* poses;  // [n][16] row-major T_wb
  double* stats;  // [n][8]  outer_iters, inner_iters, n_line, n_plane, final_cost, min_sv, is_degenerate, 0
  int* counts;    // [n][8]  n_sharp, n_flat, n_corner_ds, n_surf_ds, extract overflow, split overflow, 0, 0
};

// chained-loop arguments of the fused split + undistort + voxel launch (splitvoxel.cu)
struct SvChain {
  const OdomDev* od;
  EstState* est;
  EstInit I;
  const int* fe_counters;
  int* counts_out;
  const int* pre_idx[2];  // labelled indices compacted on the extraction stream (k_label_compact), or null
  const int* pre_cnt;
};

__host__ __device__ inline void mat4_mul(const double* A, const double* B, double* C) {
#pragma unroll
  for (int r = 0; r < 4; r++)
#pragma unroll
    for (int c = 0; c < 4; c++) {
      double s = 0;
#pragma unroll
      for (int k = 0; k < 4; k++) s += A[4 * r + k] * B[4 * k + c];
      C[4 * r + c] = s;
    }
}
__host__ __device__ inline void rigid_inv(const double* T, double* Ti) {
#pragma unroll
  for (int r = 0; r < 3; r++) {
#pragma unroll
    for (int c = 0; c < 3; c++) Ti[4 * r + c] = T[4 * c + r];
    Ti[4 * r + 3] = -(T[0 * 4 + r] * T[3] + T[1 * 4 + r] * T[7] + T[2 * 4 + r] * T[11]);
  }
  Ti[12] = Ti[13] = Ti[14] = 0;
  Ti[15] = 1;
}

// EST.cpp:1268-1270 T_wl + thres_dist schedule (EST.cpp:1207, 1377-1381): what the association needs
__device__ inline void est_begin_assoc(EstState* S) {
  const int it = S->outer_next;
  S->outer_it = it;
  const Quat Q = {S->Q[0], S->Q[1], S->Q[2], S->Q[3]};
  double Rq[9];
  quat_to_R(Q, Rq);
  // exRbl = Rbl, exPbl = Pbl
  for (int r = 0; r < 3; r++) {
    for (int c = 0; c < 3; c++)
      S->T_wl[4 * r + c] = Rq[3 * r] * S->Rbl[c] + Rq[3 * r + 1] * S->Rbl[3 + c] + Rq[3 * r + 2] * S->Rbl[6 + c];
    S->T_wl[4 * r + 3] = Rq[3 * r] * S->Pbl[0] + Rq[3 * r + 1] * S->Pbl[1] + Rq[3 * r + 2] * S->Pbl[2] + S->P[r];
  }
  S->T_wl[12] = 0; S->T_wl[13] = 0; S->T_wl[14] = 0; S->T_wl[15] = 1;
  S->thres = (float)S->thres_sched[it < 2 ? it : 2];
}
// EST.cpp:1212 vector2double: what the solve needs (idempotent)
__device__ inline void est_begin_solve(EstState* S) {
  const Quat Q = {S->Q[0], S->Q[1], S->Q[2], S->Q[3]};
  S->x[0] = S->P[0]; S->x[1] = S->P[1]; S->x[2] = S->P[2];
  so3_log(Q, S->x + 3);
  for (int i = 0; i < 4; i++) S->q_before[i] = S->Q[i];
  for (int i = 0; i < 3; i++) S->t_before[i] = S->P[i];
  S->first = 1;
  S->done_inner = 0;
}


// Fill a zeroed EstState for a new solve starting at body pose (P, Q): what k_est_init's thread 0 does.
__device__ inline void est_fill(EstState* S, const EstInit& I, const double* P, const double* Q) {
  for (int i = 0; i < 3; i++) { S->P[i] = P[i]; S->Pbl[i] = I.Pbl[i]; S->thres_sched[i] = I.thres_sched[i]; }
  for (int i = 0; i < 4; i++) S->Q[i] = Q[i];
  for (int i = 0; i < 9; i++) S->Rbl[i] = I.Rbl[i];
  S->max_outer = I.max_outer; S->max_inner = I.max_inner;
  S->lidar_m = I.lidar_m; S->w_tan = I.w_tan; S->huber_a = I.huber_a;
  est_begin_assoc(S);
  est_begin_solve(S);
}

}  // namespace mml
