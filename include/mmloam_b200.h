/*
 * mmloam_b200 — C-ABI of the B200-native mm-loam scan-matching hot path.
 *
 * Drop-in boundary for TIERS/multi-modal-loam (reference paths relative to
 * /root/reference/mm-loam). The reference has no FFI layer: the path sits behind the
 * C++ member functions listed at each entry point below; the header-only adapters in
 * multi-modal-loam_b200/host/ turn those calls into these.
 *
 * Conventions: every function returns 0 on success or a negative MML_ERR_* code and
 * never throws. Pointers are caller-owned HOST buffers unless the name ends in `_dev`.
 * Buffers are little-endian and tightly packed. A context is not re-entrant: use one
 * context per host thread (the reference calls detectFeaturePoints from 6 threads,
 * src/unionFeatureExtract.cpp:1008-1015 — the batched extractor takes all lines of a scan
 * in one call instead). There is NO CPU fallback: without a CUDA device
 * mml_ctx_create fails with MML_ERR_NO_DEVICE.
 *
 * Point layout: float4 (x, y, z, intensity) replaces pcl::PointXYZINormal (48 B);
 * side arrays carry line id (u16), sweep fraction (f32) and label (u8).
 */
#ifndef MMLOAM_B200_H
#define MMLOAM_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define MML_OK 0
#define MML_ERR_INVALID (-1)    /* bad argument */
#define MML_ERR_NO_DEVICE (-2)  /* no usable CUDA device */
#define MML_ERR_CUDA (-3)       /* CUDA runtime error, see mml_last_error */
#define MML_ERR_CAPACITY (-4)   /* input exceeds a documented limit */
#define MML_ERR_STATE (-5)      /* call order (e.g. associate before map_set) */

typedef struct mml_ctx mml_ctx;

int mml_version(void);
/* Opaque handle owning streams, scratch and the resident feature maps. */
int mml_ctx_create(int device, int stream_count, mml_ctx** out);
int mml_ctx_destroy(mml_ctx* ctx);
const char* mml_last_error(const mml_ctx* ctx);
/* Number of kernels this context has launched so far (bench.py's gpu_launches). */
long long mml_launch_count(const mml_ctx* ctx);
int mml_sync(mml_ctx* ctx);

/* ---- A1 (+ glue): feature_extraction::detectFeaturePoints, FE.cpp:341-844, called per
 * line at FE.cpp:1010 / 1229; label write-back FE.cpp:1016-1023, 1233-1240.
 * One call labels every line of a scan: label 0 none / 1 corner ("sharp") / 2 surf
 * ("flat") in input order = the reference's normal_z. line_id[i] < n_lines.           */
int mml_extract_features(mml_ctx* ctx, const float* xyzi, const uint16_t* line_id, int n, int n_lines,
                         uint8_t* out_label, int* out_n_sharp, int* out_n_flat);
/* Batched form: scans are concatenated, scan s = [scan_offsets[s], scan_offsets[s+1]). */
int mml_extract_features_batch(mml_ctx* ctx, const float* xyzi, const uint16_t* line_id,
                               const int* scan_offsets, int n_scans, int n_lines, uint8_t* out_label,
                               int* out_n_sharp, int* out_n_flat);

/* Device-resident, asynchronous form (kernels only; labels stay in HBM): used for batched throughput runs. */
int mml_extract_features_batch_dev(mml_ctx* ctx, const void* xyzi_dev, const void* line_id_dev, const int* scan_offsets,
                                   int n_scans, int n_lines, void* label_dev);

/* ---- A2: getVeloFeature ring + relative time, FE.cpp:1136-1195. line_out = -1 rejected. */
int mml_velo_ring_time(mml_ctx* ctx, const float* xyzi, int n, int16_t* line_out, float* reltime_out);
/* ---- A3: getHoriFeatureExtract filter, FE.cpp:985-998 (CustomPoint fields as arrays). */
int mml_hori_filter(mml_ctx* ctx, const uint32_t* offset_time, const float* xyz3, const uint8_t* line, int n,
                    uint8_t* keep, float* reltime_out);

/* ---- global feature map kept on the device: MAP_MANAGER::MapIncrement with MapMove, src/lio/Map_Manager.cpp:125-281,
 * 288-581 (SURVEY.md 8 f, row F1, second half). The 21 x 11 x 21 grid of 50 m cubes lives in HBM, every populated cube
 * in insertion order. One call = one update: the cubes as they are BEFORE the update become the global maps (kinds 0 / 1)
 * the association searches (MM.cpp:133-146: the matcher lags the map by one update), MapMove re-centres the grid on
 * T_wl16 (may be NULL), the new world-frame points are appended to their cubes and touched cubes above 300 points are
 * voxel-filtered with leaf 0.4. n_from_map2 (may be NULL): sizes of laserCloudCornerFromMap / SurfFromMap.            */
int mml_global_map_push(mml_ctx* ctx, const float* corner_world_xyzi, int n_corner, const float* surf_world_xyzi,
                        int n_surf, const double* T_wl16, int* n_from_map2);
/* which: 0 all cubes now, 1 the snapshot the matcher sees, 2 laserCloud*FromMap; cen3_out: (CenWidth, CenHeight, CenDepth) */
int mml_global_map_get(mml_ctx* ctx, int kind, int which, float* out_xyzi, int cap, int* n_out, int* cen3_out);
int mml_global_map_reset(mml_ctx* ctx);

/* ---- F2: the message stages either side of the extractor, on the device (csrc/msgpack.cu).
 * livox_ros_driver/CustomMsg points (CustomPoint.msg:3-9 serialised: 19 packed bytes per point) -> xyzi, line and sweep
 * fraction behind the filter of getHoriFeatureExtract (FE.cpp:985-998: line <= used_line - 1, x >= 0.01), in message
 * order. Outputs may be NULL (the unpacked scan also stays in the context's staging buffers); *m_out = points kept.     */
int mml_unpack_custom_points(mml_ctx* ctx, const void* points19, int n, int used_line, float* xyzi_out,
                             uint16_t* line_out, float* s_out, int* m_out);
/* sensor_msgs/PointCloud2 data (point_step bytes per point, float32 fields at the given byte offsets, off_intensity < 0:
 * none) -> xyzi with pcl::removeNaNFromPointCloud (FE.cpp:1129-1133).                                                   */
int mml_unpack_pointcloud2(mml_ctx* ctx, const void* data, int n, int point_step, int off_x, int off_y, int off_z,
                           int off_intensity, float* xyzi_out, int* m_out);
/* labelled scan -> the three clouds of union_cloud.msg as pcl::PointXYZINormal records (48 B: x y z 1 | normal_x = s,
 * normal_y = line, normal_z = label, 0 | intensity, curvature = 0, 0, 0: the bytes pcl::toROSMsg copies).
 * full: every point inside [near_full, far_full]; corner / surf: label 1 / 2 inside [near_feat, far_feat]; a far
 * threshold <= 0 switches that cut off (removeNearPointCloud vs removeNearFarPoints, lidars_extrinsic_cali.h:424-477;
 * FE.cpp:916-937 Horizon, 1278-1297 VLP-16). zero_full_intensity: the VLP-16 branch zeroes the full cloud's intensity
 * (FE.cpp:1263-1265). counts3 = points written to full / corner / surf; the outputs hold n records each.              */
int mml_pack_union_clouds(mml_ctx* ctx, const float* xyzi, const float* s, const uint16_t* line, const uint8_t* label,
                          int n, float near_full, float far_full, float near_feat, float far_feat,
                          int zero_full_intensity, void* full_out, void* corner_out, void* surf_out, int* counts3);

/* ---- A4: RemoveLidarDistortion, src/unionPoseEstimation.cpp:402-421. In place.      */
int mml_undistort(mml_ctx* ctx, float* xyzi, const float* s, int n, const double* dR9, const double* dt3);

/* ---- A6: pcl::VoxelGrid::filter as called at src/lio/Estimator.cpp:1015-1024.
 * out has capacity n points; output ordered by voxel index.                           */
int mml_voxel_downsample(mml_ctx* ctx, const float* xyzi, int n, float leaf, float* out, int* m_out);

/* ---- map upload: replaces the kd-tree (re)builds at EST.cpp:1159-1179 and the cube
 * binning of src/lio/Map_Manager.cpp:159-175. Builds the device spatial hash.
 * kind: 0 corner-global, 1 surf-global, 2 corner-local, 3 surf-local.
 * cube_centre3 = (laserCloudCenWidth, CenHeight, CenDepth), NULL = (10, 5, 10).       */
#define MML_MAP_CORNER_GLOBAL 0
#define MML_MAP_SURF_GLOBAL 1
#define MML_MAP_CORNER_LOCAL 2
#define MML_MAP_SURF_LOCAL 3
int mml_map_set(mml_ctx* ctx, int kind, const float* xyzi, int m, const int* cube_centre3);
/* the same with the points already resident in HBM (float4 xyzi)                       */
int mml_map_set_dev(mml_ctx* ctx, int kind, const void* xyzi_dev, int m, const int* cube_centre3);
/* Same with an explicit hash-cell edge in metres (0 = choose automatically).           */
int mml_map_set_ex(mml_ctx* ctx, int kind, const float* xyzi, int m, const int* cube_centre3, float cell);
/* info8 = [valid, points, cell edge, dim x, dim y, dim z, cells, cells per 50 m cube]  */
int mml_map_info(mml_ctx* ctx, int kind, double* info8);
/* Test view of a built map level (0 fine, 1 coarse): cell-sorted points (xyz + original index bits) and the
 * cell table; mml_map_dims: dims7 = [dim xyz, coarse dim xyz, coarse factor], org3 = grid origin.     */
int mml_map_dump(mml_ctx* ctx, int kind, int level, float* pts_out, int* cell_start_out);
int mml_map_dims(mml_ctx* ctx, int kind, int* dims7, double* org3);

/* ---- A7 / A8: Estimator::processPointToLine EST.cpp:148-365 and
 * Estimator::processPointToPlanVec EST.cpp:573-777 (incl. A5 pointAssociateToMap +
 * cube rule, MM.cpp:75-89, 583-629).
 * kind 0: line features  [pointOri(3) lineP1(3) lineP2(3) error valid src]
 * kind 1: plane features [pointOri(3) pointProj(3) normal(3) error valid src]
 * out_feat: nq x 12 doubles, slot i belongs to query i; valid = -1 no feature,
 * 0 feature with |error| <= 1e-5 (dropped by EST.cpp:1313/1385), 1 used.
 * normal_moment9 / n_normals (kind 1, may be NULL): sum n n^T and count of accepted planes,
 * the input of checkLocalizability (EST.cpp:536-565).                                  */
int mml_associate(mml_ctx* ctx, int kind, const float* q_xyzi, int nq, const double* T_wl16, double thres_dist,
                  double* out_feat, int* n_feat, double* normal_moment9, int* n_normals);

/* ---- A9-A11: Cost_NavState_IMU_Line / _Plan_Vec evaluation + Huber + per-pose
 * accumulation, include/utils/ceresfunc.h:412-440, 533-555, 33-63 (what ceres::Solve at
 * EST.cpp:1425-1432 evaluates). x6 = [t_wb, phi_wb]; T_bl16 = exTlb^-1 row-major.
 * Outputs H = sum J^T J (6x6 row-major), g = sum J^T r, cost = 1/2 sum rho.
 * huber_a <= 0 disables the loss (window size 5, EST.cpp:1219).                        */
int mml_accumulate(mml_ctx* ctx, const double* line_feat, int n_line, const double* plane_feat, int n_plane,
                   const double* x6, const double* T_bl16, double plan_weight_tan, double huber_a, double* H36,
                   double* g6, double* cost);

/* ---- A12: Estimator::Estimate outer loop for one frame, EST.cpp:1143-1581 (window size
 * 1, the branch the shipped launch file runs): <= max_outer x { associate ; dogleg solve
 * <= max_inner iterations } entirely on the device (one CUDA-graph launch, no host round
 * trip per iteration). P3 / q_wxyz4 = body pose, updated in place.                     */
typedef struct {
  int max_outer;          /* 5    EST.cpp:1210 */
  int max_inner;          /* 10   EST.cpp:1428 */
  double lidar_m;         /* 1.5e-3 include/IMUIntegrator/IMUIntegrator.h:83 */
  double plan_weight_tan; /* 0.0  EST.cpp:1206 */
  double thres0, thres1, thres2; /* 25, 10, 1  EST.cpp:1207, 1377-1381 */
  int use_huber;          /* 1    EST.cpp:1221 */
  int map_update;         /* 0    mml_odom_run_window only. 0: the context's maps are used as they are. 1: after a
                           *      solve the loop runs the map update of EstimateLidarPose (EST.cpp:1073-1135, the
                           *      lidarMode 2 branch PE.cpp:709, 872 takes): when the solve is not degenerate and the
                           *      sensor has moved by >= sqrt(0.5) m since the last update, the OLDEST frame's clouds
                           *      go through Estimator::MapIncrementLocal (mml_local_map_push_dev, clear_first) */
} mml_est_params;
void mml_est_params_default(mml_est_params* p);
/* stats (may be NULL, 16 doubles): [outer_iters, inner_iters, n_line, n_plane, final_cost,
 * min_singular_value, is_degenerate, ...]. corner_xyzi = surf_xyzi = NULL with n_corner = n_surf = -1: solve the frame
 * mml_frame_set left in HBM (the same holds for mml_estimate_sharded).                    */
int mml_estimate(mml_ctx* ctx, const float* corner_xyzi, int n_corner, const float* surf_xyzi, int n_surf,
                 const double* exTlb16, double* P3, double* q_wxyz4, const mml_est_params* prm, double* stats);

/* ---- whole per-scan path, device resident between stages:
 * extract (A1) -> undistort (A4) -> label split + voxel filter (A6) -> estimate (A12).
 * This is the loop body of process(), src/unionPoseEstimation.cpp:862-872, for one scan.
 * s = per-point sweep fraction (normal_x). dR9/dt3 = predicted motion for undistortion.
 * out_counts (may be NULL, 4 ints): n_sharp, n_flat, n_corner_ds, n_surf_ds.           */
int mml_scan_to_pose(mml_ctx* ctx, const float* xyzi, const uint16_t* line_id, const float* s, int n, int n_lines,
                     const double* dR9, const double* dt3, float leaf_corner, float leaf_surf,
                     const double* exTlb16, double* P3, double* q_wxyz4, const mml_est_params* prm,
                     double* stats, int* out_counts);
/* Same with inputs already resident in device memory (bench.py's kernel-only `value`). */
int mml_scan_to_pose_dev(mml_ctx* ctx, const void* xyzi_dev, const void* line_id_dev, const void* s_dev, int n,
                         int n_lines, const double* dR9, const double* dt3, float leaf_corner, float leaf_surf,
                         const double* exTlb16, double* P3, double* q_wxyz4, const mml_est_params* prm,
                         double* stats, int* out_counts);

/* ---- native odometry loop: the per-scan body of process() (src/unionPoseEstimation.cpp:650-906) over a
 * sequence of scans, keeping the reference's node pipeline (the copy of scan k+3 and the feature extraction of
 * scans k+1, k+2 overlap the matching of scan k). The loop is chained on the device: the pose history and the
 * constant-velocity prediction live in device memory and a scan's solve is one graph launch, so the call enqueues
 * the whole sequence and reads the poses back once. Scans that exceed the fused kernels' capacities (a scan line
 * beyond the selection kernel's tier, more than 16384 labelled points of a kind) are re-run on the general path;
 * results are the same either way. MML_ODOM_CLASSIC=1 selects the host-driven driver (one wait per outer iteration). xyzi / line / s: arrays of per-scan pointers, device pointers when
 * host_buffers == 0, host (ideally pinned) pointers otherwise. T_init16 / T_prev16: the two poses before the
 * first scan (constant-velocity seed, PE.cpp:847-852). poses_out: n_scans x 16 row-major T_wb.
 * total_ms (may be NULL): CUDA-event time of the run. counts_out (may be NULL): n_scans x 4.            */
int mml_odom_run(mml_ctx* ctx, const void* const* xyzi, const void* const* line, const void* const* s,
                 const int* n_pts, int n_scans, int n_lines, int host_buffers, const double* T_init16,
                 const double* T_prev16, const double* exTlb16, float leaf_corner, float leaf_surf,
                 const mml_est_params* prm, double* poses_out, float* total_ms, int* counts_out);

/* ---- local feature map kept on the device: Estimator::MapIncrementLocal, src/lio/Estimator.cpp:1585-1643
 * (SURVEY.md 8 f, row F1). One call = one map update: the frame's corner / surf clouds (LiDAR frame) are moved to
 * the world frame with T_wl16 (MAP_MANAGER::pointAssociateToMap, MM.cpp:75-89), stored in the 50-frame ring, the
 * previous filtered map and the ring are concatenated (EST.cpp:1620-1624) and voxel-filtered (EST.cpp:1630-1635),
 * and the spatial hash of the local map kinds (2, 3) is rebuilt from the result - no host copy of the map exists.  */
int mml_local_map_push(mml_ctx* ctx, const float* corner_xyzi, int n_corner, const float* surf_xyzi, int n_surf,
                       const double* T_wl16, float leaf_corner, float leaf_surf, int* n_corner_map, int* n_surf_map);
/* current local map of one kind (0 corner / 1 surf); out_xyzi may be NULL to query the size only                  */
/* the same with the clouds resident in HBM (float4 xyzi). clear_first != 0: laserCloud*FromLocal were cleared by the
 * caller (EST.cpp:1085-1087, 1127-1129, what EstimateLidarPose does), so the map becomes the filtered ring alone.
 * An update is transactional: ring, map sizes and the association's search structures change only when every step
 * of both kinds succeeded.                                                                                        */
int mml_local_map_push_dev(mml_ctx* ctx, const void* corner_dev, int n_corner, const void* surf_dev, int n_surf,
                           const double* T_wl16, float leaf_corner, float leaf_surf, int clear_first,
                           int* n_corner_map, int* n_surf_map);
/* place a world-frame cloud into ring entry `slot` (0..49) of one kind: the state after earlier frames were pushed */
int mml_local_map_seed(mml_ctx* ctx, int kind, int slot, const float* xyzi_world, int n);
int mml_local_map_get(mml_ctx* ctx, int kind, float* out_xyzi, int cap, int* n_out);
/* forget the ring and the filtered maps; map kinds 2 / 3 become invalid (nothing is matched against them)          */
int mml_local_map_reset(mml_ctx* ctx);

/* ---- sliding window, sizes 2-4 (IMU factors, no marginalisation; BASELINE config 3 is window 3):
 * Estimator::Estimate for windowSize < SLIDEWINDOWSIZE, src/lio/Estimator.cpp:1143-1581.
 * mml_preint = IMUIntegrator after PreIntegration (src/lio/IMUIntegrator.cpp:105-166): delta rotation (w, x, y, z),
 * position, velocity, time, the biases it was linearised at, covariance and Jacobian (15 x 15 row-major, order
 * P R V BG BA, include/IMUIntegrator/IMUIntegrator.h:86-93) and sqrt_info = LLT(cov^-1).matrixL()^T (EST.cpp:1240-1242). */
typedef struct {
  double dq[4];
  double dp[3], dv[3], dt;
  double bg[3], ba[3];
  double cov[225], jac[225], sqrt_info[225];
} mml_preint;
/* t / gyr / acc: n samples in (last_time, t_frame]; acc in units of g (scaled by gnorm = 9.805, IMU.h:84). Host code. */
int mml_imu_preintegrate(const double* t, const double* gyr, const double* acc, int n, double last_time,
                         const double* bg3, const double* ba3, mml_preint* out);
/* The mean alone (dq, dp, dv, dt, bg, ba; cov / jac / sqrt_info are left untouched): the same arithmetic as
 * mml_imu_preintegrate, bit for bit, at a fraction of its time. The odometry loop predicts the new frame's pose from it
 * (PE.cpp:812-829) and forms the full pre-integration while the device already works on the scan. Host code.       */
int mml_imu_preintegrate_mean(const double* t, const double* gyr, const double* acc, int n, double last_time,
                              const double* bg3, const double* ba3, mml_preint* out);
/* Cost_NavState_PRV_Bias (include/utils/ceresfunc.h:321-393) weighted by sqrt_info: r15 and, if J450 != NULL, the
 * Jacobian (15 x 30 row-major, columns [PR_i 6 | VBias_i 9 | PR_j 6 | VBias_j 9]) by forward-mode differentiation. */
int mml_imu_factor(const mml_preint* pre, const double* gravity3, const double* pri6, const double* vbi9,
                   const double* prj6, const double* vbj9, double* r15, double* J450);
/* Pose prediction of process(), src/unionPoseEstimation.cpp:812-829. state = P(3) q_wxyz(4) V(3) bg(3) ba(3).      */
int mml_imu_predict(const double* prev16, const mml_preint* pre, double* next16);
/* Window frames live in HBM as their downsampled corner / surf clouds. A push drops the oldest frame when the
 * window already holds max_frames (PE.cpp:830-832).                                                               */
int mml_window_reset(mml_ctx* ctx);
int mml_window_size(mml_ctx* ctx);
int mml_window_push_frame(mml_ctx* ctx, const float* corner_xyzi, int n_corner, const float* surf_xyzi, int n_surf,
                          int max_frames);
/* raw scan resident in HBM: extraction -> undistortion + label split + voxel filter on the device               */
int mml_window_push_scan_dev(mml_ctx* ctx, const void* xyzi_dev, const void* line_id_dev, const void* s_dev, int n,
                             int n_lines, const double* dR9, const double* dt3, float leaf_corner, float leaf_surf,
                             int max_frames, int* out_counts);
int mml_window_get_frame(mml_ctx* ctx, int frame, int kind, float* out_xyzi, int cap, int* n_out);
/* Estimate over the frames in the window. states: W x 16 doubles in place; preints[f] (f >= 1) links frame f-1 -> f.
 * The whole solve runs on the device (csrc/windowsolve.cu): association of every frame, lidar terms, the W-1 IMU
 * factors and the (15 W)-dimensional dogleg step, one graph launch and one host wait per call.
 * stats (may be NULL, 16 doubles): [outer, inner, n_line, n_plane (newest frame), final_cost, min_sv, degenerate, evaluations]. */
int mml_estimate_window(mml_ctx* ctx, double* states, const mml_preint* const* preints, const double* exTlb16,
                        const double* gravity3, const mml_est_params* prm, double* stats);

/* Odometry loop with an IMU-initialised sliding window of `window` frames: the per-scan body of process() in its
 * LidarIMUInited branch (src/unionPoseEstimation.cpp:796-891): pre-integration of the IMU samples of
 * (t_{k-1}, t_k], state prediction, undistortion with the predicted LiDAR motion, window push, EstimateLidarPose.
 * xyzi / line / s: per-scan pointers (device pointers when host_buffers == 0). imu_*: all samples concatenated, scan
 * k owns imu_n[k] of them (acc in units of g). state0 / stamp0: state and time of the frame before the first scan.
 * poses_front: the reference's odometry output, the OLDEST frame of the window (EST.cpp:1043-1049); poses_newest:
 * the newest frame; states_out: its full state; stats_out: n_scans x 8 (see mml_estimate_window). Any may be NULL.
 * prm->map_update selects whether the local feature maps follow the trajectory (see mml_est_params).               */
int mml_odom_run_window(mml_ctx* ctx, const void* const* xyzi, const void* const* line, const void* const* s,
                        const int* n_pts, int n_scans, int n_lines, int host_buffers, int window, const double* stamps,
                        double stamp0, const double* imu_t, const double* imu_gyr, const double* imu_acc, const int* imu_n,
                        const double* state0, const double* exTlb16, const double* gravity3, float leaf_corner,
                        float leaf_surf, const mml_est_params* prm, double* poses_front, double* poses_newest,
                        double* states_out, double* stats_out, float* total_ms);

/* ---- (e) multi-GPU: the global map sharded by 50 m cube over up to 8 GPUs (MM.cpp:583-605 is the shard boundary; replaces
 * the `_allreduce(ncclComm_t)` variant SURVEY.md 8 b sketches). Rank r uploads only its cubes (mml_map_set on the global
 * kinds); queries are replicated; each query is matched on exactly one rank; the only exchange - 28 doubles per
 * evaluation, 18 per association - goes through PEER MEMORY from inside the kernels: the last CTA of an evaluation stores
 * its sums into every rank's exchange buffer over NVLink, publishes a sequence word, waits for the others and sums in
 * rank order, then takes the dogleg step on the device. No NCCL or host call between two evaluations.
 *   mml_shard_init         allocate this rank's exchange buffer; ipc_handle64_out (may be NULL) receives its 64-byte
 *                          cudaIpcMemHandle for the caller to all-gather (MPI, torch.distributed, a file ...)
 *   mml_shard_connect_ipc  handles of all ranks in rank order (world x 64 bytes), one process per GPU
 *   mml_shard_connect_ptrs exchange-buffer pointers of all ranks when they live in one process (mml_shard_local_ptr),
 *                          devices[r] = CUDA device of rank r (peer access is enabled here)
 *   mml_estimate_sharded   collective: same queries and start pose on every rank; every rank returns the same pose.  */
int mml_shard_init(mml_ctx* ctx, int rank, int world, void* ipc_handle64_out);
int mml_shard_local_ptr(mml_ctx* ctx, void** buf_out);
int mml_shard_connect_ipc(mml_ctx* ctx, const void* handles);
int mml_shard_connect_ptrs(mml_ctx* ctx, void* const* bufs, const int* devices);
int mml_shard_close(mml_ctx* ctx);
int mml_estimate_sharded(mml_ctx* ctx, const float* corner_xyzi, int n_corner, const float* surf_xyzi, int n_surf,
                         const double* exTlb16, double* P3, double* q_wxyz4, const mml_est_params* prm, double* stats);

/* ---- device-resident building blocks used by bench.py's roofline sweep (S4):
 * queries and maps stay in HBM; one call = one association or one evaluation.          */
int mml_frame_set(mml_ctx* ctx, const float* corner_xyzi, int n_corner, const float* surf_xyzi, int n_surf);
int mml_frame_associate(mml_ctx* ctx, const double* T_wl16, double thres_dist, int* n_line, int* n_plane,
                        double* normal_moment9, int* n_normals);
int mml_frame_accumulate(mml_ctx* ctx, const double* x6, const double* T_bl16, double plan_weight_tan,
                         double huber_a, double* H36, double* g6, double* cost);
/* Features of the frame slot as host records (kind 0 line / 1 plane, n_query x 12).    */
int mml_frame_get_features(mml_ctx* ctx, int kind, double* out_feat);
/* Raw device buffers for callers that keep scans resident in HBM (bench.py).           */
int mml_dev_alloc(mml_ctx* ctx, size_t bytes, void** out);
int mml_dev_free(mml_ctx* ctx, void* p);
int mml_dev_upload(mml_ctx* ctx, void* dst_dev, const void* src, size_t bytes);
/* Per-stage CUDA-event timing of mml_scan_to_pose[_dev], summed over calls:
 * stage_ms3 = [extract, undistort + split + voxel, estimate].                          */
int mml_profile_enable(mml_ctx* ctx, int on);
int mml_profile_read(mml_ctx* ctx, double* stage_ms3, long long* n_scans);
/* asynchronous variants for timing: enqueue `repeat` launches, no host read-back.      */
int mml_frame_associate_async(mml_ctx* ctx, const double* T_wl16, double thres_dist, int repeat);
int mml_frame_accumulate_async(mml_ctx* ctx, const double* x6, const double* T_bl16, double plan_weight_tan,
                               double huber_a, int repeat);
/* one association kind only (0 line / 1 plane): times a single kernel                  */
int mml_frame_associate_kind_async(mml_ctx* ctx, int kind, const double* T_wl16, double thres_dist, int repeat);
/* CUDA-event timing on the context's stream (bench.py cannot see it from torch).       */
int mml_timer_start(mml_ctx* ctx);
int mml_timer_stop_ms(mml_ctx* ctx, float* ms);
/* Partial normal equations of this rank's map shard, left on the device for an NCCL
 * all-reduce: returns the device pointer to 28 doubles [cost, g 6, H upper triangle 21]. */
int mml_frame_accumulate_partial_dev(mml_ctx* ctx, const double* x6, const double* T_bl16, double plan_weight_tan,
                                     double huber_a, void** partial28_dev);
void* mml_stream_handle(mml_ctx* ctx);

/* ---- host-side trust-region state machine (the same code the device loop runs) for callers
 * that reduce the normal equations themselves, e.g. after an NCCL all-reduce of per-shard
 * partial sums (SURVEY.md §8 e). Restates ceres::Solve's DOGLEG iteration (EST.cpp:1425-1432)
 * on [cost, g(6), upper-triangular H(21)]. Needs no GPU.                                  */
typedef struct mml_solver mml_solver;
int mml_solver_create(mml_solver** out);
int mml_solver_destroy(mml_solver* s);
int mml_solver_begin(mml_solver* s, const double* x6, int max_inner);
int mml_solver_feed(mml_solver* s, const double* out28, double* x_next6, int* done);
int mml_solver_result(mml_solver* s, double* x_best6, double* min_cost, int* iterations);

#ifdef __cplusplus
}
#endif
#endif
