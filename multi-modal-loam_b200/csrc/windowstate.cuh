// Sliding-window state shared by window.cu (frame slots, IMU host code, the odometry loop) and windowsolve.cu (the
// device-side solve of the window: Estimator::Estimate for 1 <= windowSize < SLIDEWINDOWSIZE, EST.cpp:1143-1581).
#pragma once
#include "common.cuh"
#include "../../include/mmloam_b200.h"

namespace mml {

constexpr int kMaxWindow = 4;            // frames (window sizes 1-4: no marginalisation, EST.cpp:1450)
constexpr int kWinN = 15 * kMaxWindow;   // unknowns: [P, log Q] x W, then [V, bg, ba] x W (EST.cpp:937-950)

// Device-resident window: everything the association and solve kernels read. The head (up to `upload_end`) is
// written by the host (per-call API) or shifted on the device (odometry loop); the rest belongs to the kernels.
struct WinDev {
  int W, n;                               // frames in the window, unknowns (6 for a lone frame, else 15 W)
  int slot_of[kMaxWindow];                // frame f (0 = oldest) -> physical frame slot
  int max_outer, max_inner;
  double states[kMaxWindow][16];          // P(3) q_wxyz(4) V(3) bg(3) ba(3), updated in place by every outer iteration
  mml_preint pre[kMaxWindow];             // pre[f] links frame f-1 -> f (f >= 1)
  double gravity[3], Rbl_raw[9], Rbl_q[9], Pbl[3];
  double lidar_m, w_tan, huber_a, thres_sched[3];
  unsigned seq;                           // sequence number of the solve, published to the host with the result
  int upload_end;
  // association inputs, by PHYSICAL slot (the association nodes of the graph are bound to physical slots)
  double T_wl[kMaxWindow][16];
  float thres;
  int done_outer, outer_it;
  // statistics of the solve
  int total_inner, evals, is_degenerate, n_line_last, n_plane_last;
  double final_cost, min_sv;
};

// one scan's hand-over to the device-side window in the odometry loop (k_win_push)
struct WinPush {
  double state[16];
  mml_preint pre;
  int slot;        // physical slot the new frame's clouds were written to
  int window;      // configured window size
  unsigned seq;
  int pad;
};

struct WinSlot {
  DevBuf q_corner, q_surf, f_line, f_plane;
  DevBuf assoc_stats, assoc_part[2];  // per-frame association statistics: frames are associated concurrently
  DevBuf cnt;                         // device counts [n_corner, n_surf, ...] (8 ints)
  int n_corner = 0, n_surf = 0;       // host copies (-1: only known on the device)
  int cap = 0;                        // capacity the slot's buffers are reserved for
};

struct WindowState {
  WinSlot slot[kMaxWindow];           // physical slots
  int order[kMaxWindow] = {0, 1, 2, 3};  // frame f -> physical slot
  int n_slots = 0;
  DevBuf dev;             // WinDev
  DevBuf push_dev;        // WinPush staging on the device (odometry loop)
  // zero-copy hand-over to the host: mapped pinned memory written by the kernels
  //   [0, 16 W) states | [64, 80) statistics | sequence word at kMapDoubles
  void* mapped = nullptr;
  double* mapped_dev = nullptr;
  void* pin_up = nullptr;             // pinned upload staging (WinDev head / ring of WinPush records)
  unsigned seq = 0;
  cudaStream_t fstream[kMaxWindow][2] = {};
  cudaEvent_t fev[kMaxWindow][2] = {};
  cudaEvent_t fork = nullptr;
  bool streams_ok = false;
  cudaGraphExec_t graph = nullptr;    // device solve of the window: 2 outer iterations as plain nodes + WHILE
  long long graph_key = 0;
  long long graph_launches = 0;
  // odometry loop: the next scan is copied (host buffers) and labelled on its own stream while the window is solved
  cudaStream_t xstream = nullptr;
  cudaEvent_t xev[2] = {nullptr, nullptr}, xfree[2] = {nullptr, nullptr};
  DevBuf x_xyzi[2], x_line[2], x_s[2], x_label[2], x_counters[2];
  DevBuf x_idx[2];        // labelled indices compacted behind the extraction: [idx label 1 | idx label 2 | counts(2)]
  WinSlot& frame(int f) { return slot[order[f]]; }
};

constexpr int kMapDoubles = (28 + 20) * kMaxWindow;
constexpr size_t kMapBytes = sizeof(double) * kMapDoubles + 64;
constexpr size_t kPinUpBytes = sizeof(WinDev) + 8 * sizeof(WinPush);

}  // namespace mml

// windowsolve.cu
int mml_window_solve_graph(mml_ctx* c, mml::WindowState* w, int cap);
int mml_window_solve_launch(mml_ctx* c, mml::WindowState* w, int cap, int max_outer);
int mml_window_begin_launch(mml_ctx* c, mml::WindowState* w);
int mml_window_push_launch(mml_ctx* c, mml::WindowState* w, const mml::WinPush* push_dev);
