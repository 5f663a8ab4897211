// mmloam_b200: the message stages either side of the extractor, on the device (SURVEY.md §8 f, F2).
//   unpack  livox_ros_driver/CustomMsg points -> xyzi float4 + line + sweep fraction, with the filter of
//           getHoriFeatureExtract                      src/unionFeatureExtract.cpp:985-998
//           (CustomPoint.msg:3-9 serialises to 19 packed little-endian bytes: u32 offset_time, f32 x y z,
//            u8 reflectivity, u8 tag, u8 line)
//   unpack  sensor_msgs/PointCloud2 -> xyzi float4 with pcl::removeNaNFromPointCloud   FE.cpp:1129-1133
//   pack    labelled cloud -> the three clouds of union_cloud.msg (combine / corner / surface) as
//           pcl::PointXYZINormal records (48 B, what pcl::toROSMsg copies) behind removeNearFarPoints /
//           removeNearPointCloud                        FE.cpp:916-937, 1278-1297; include/lidars_extrinsic_cali.h:424-477
// Every selection is a stable compaction (the reference push_backs in input order): flags -> exclusive scan -> scatter.
// Compiled with -fmad=false: the squared range is ((x x) + (y y)) + (z z) in float32 like the reference's.
#include "common.cuh"
#include "sort.cuh"
#include <math.h>

namespace mml {

__device__ __forceinline__ uint32_t ld_u32(const uint8_t* p) {  // records are not 4-byte aligned
  return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}
__device__ __forceinline__ float ld_f32(const uint8_t* p) { return __uint_as_float(ld_u32(p)); }
// ros::Time().fromNSec(t).toSec(): seconds and nanoseconds are split before the conversion
__device__ __forceinline__ double nsec_to_sec(uint32_t t) { return (double)(t / 1000000000u) + 1e-9 * (double)(t % 1000000000u); }

constexpr int kCustomPointBytes = 19;

__global__ void __launch_bounds__(256) k_custom_flags(const uint8_t* __restrict__ rec, int n, int used_line, int* __restrict__ flag) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const uint8_t* r = rec + (size_t)kCustomPointBytes * i;
  const int line = (int)r[18];
  const float x = ld_f32(r + 4);
  flag[i] = (!(line > used_line - 1) && !((double)x < 0.01)) ? 1 : 0;  // FE.cpp:988-989
}
__global__ void __launch_bounds__(256) k_custom_scatter(const uint8_t* __restrict__ rec, int n, const int* __restrict__ pos,
                                                        float4* __restrict__ xyzi, uint16_t* __restrict__ line_out,
                                                        float* __restrict__ s_out) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const int o = pos[i];
  if (pos[i + 1] == o) return;  // pos has n + 1 entries: pos[n] = number kept
  const uint8_t* r = rec + (size_t)kCustomPointBytes * i;
  const double time_span = nsec_to_sec(ld_u32(rec + (size_t)kCustomPointBytes * (n - 1)));  // FE.cpp:985
  xyzi[o] = make_float4(ld_f32(r + 4), ld_f32(r + 8), ld_f32(r + 12), (float)r[16]);
  line_out[o] = (uint16_t)r[18];
  s_out[o] = (float)(nsec_to_sec(ld_u32(r)) / time_span);  // FE.cpp:994
}

__global__ void __launch_bounds__(256) k_pc2_flags(const uint8_t* __restrict__ data, int n, int step, int ox, int oy, int oz,
                                                   int* __restrict__ flag) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const uint8_t* r = data + (size_t)step * i;
  const float x = ld_f32(r + ox), y = ld_f32(r + oy), z = ld_f32(r + oz);
  flag[i] = (isfinite(x) && isfinite(y) && isfinite(z)) ? 1 : 0;  // pcl::removeNaNFromPointCloud
}
__global__ void __launch_bounds__(256) k_pc2_scatter(const uint8_t* __restrict__ data, int n, int step, int ox, int oy, int oz,
                                                     int oi, const int* __restrict__ pos, float4* __restrict__ xyzi) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const int o = pos[i];
  if (pos[i + 1] == o) return;
  const uint8_t* r = data + (size_t)step * i;
  xyzi[o] = make_float4(ld_f32(r + ox), ld_f32(r + oy), ld_f32(r + oz), oi >= 0 ? ld_f32(r + oi) : 0.f);
}

struct PackCuts { float near_full2, far_full2, near_feat2, far_feat2; int far_full_on, far_feat_on; };

__global__ void __launch_bounds__(256) k_pack_flags(const float4* __restrict__ xyzi, const uint8_t* __restrict__ label, int n, PackCuts C,
                                                    int* __restrict__ f_full, int* __restrict__ f_corner, int* __restrict__ f_surf) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const float4 p = xyzi[i];
  const float dis = (p.x * p.x + p.y * p.y) + p.z * p.z;  // lidars_extrinsic_cali.h:463-465
  const bool full = !(dis < C.near_full2 || (C.far_full_on && dis > C.far_full2));
  const bool feat = !(dis < C.near_feat2 || (C.far_feat_on && dis > C.far_feat2));
  const uint8_t l = label[i];
  f_full[i] = full ? 1 : 0;
  f_corner[i] = (feat && l == 1) ? 1 : 0;
  f_surf[i] = (feat && l == 2) ? 1 : 0;
}
// pcl::PointXYZINormal: data[4] = x y z 1 | data_n[4] = normal_x normal_y normal_z 0 | intensity curvature 0 0
__device__ __forceinline__ void store_xyzin(float4* out, int o, float4 p, float nx, float ny, float nz, float intensity) {
  out[3 * (size_t)o] = make_float4(p.x, p.y, p.z, 1.f);
  out[3 * (size_t)o + 1] = make_float4(nx, ny, nz, 0.f);
  out[3 * (size_t)o + 2] = make_float4(intensity, 0.f, 0.f, 0.f);
}
__global__ void __launch_bounds__(256) k_pack_scatter(const float4* __restrict__ xyzi, const float* __restrict__ s,
                                                      const uint16_t* __restrict__ line, const uint8_t* __restrict__ label, int n,
                                                      const int* __restrict__ p_full, const int* __restrict__ p_corner,
                                                      const int* __restrict__ p_surf, int zero_full_intensity,
                                                      float4* __restrict__ full, float4* __restrict__ corner, float4* __restrict__ surf) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const float4 p = xyzi[i];
  const float nx = s ? s[i] : 0.f, ny = (float)line[i], nz = (float)label[i];
  if (p_full[i + 1] != p_full[i]) store_xyzin(full, p_full[i], p, nx, ny, nz, zero_full_intensity ? 0.f : p.w);  // FE.cpp:1263-1265
  if (p_corner[i + 1] != p_corner[i]) store_xyzin(corner, p_corner[i], p, nx, ny, nz, p.w);
  if (p_surf[i + 1] != p_surf[i]) store_xyzin(surf, p_surf[i], p, nx, ny, nz, p.w);
}

}  // namespace mml

using namespace mml;

// flags[0..n) (0/1) -> positions, flags[n] = number kept (the scan runs over n + 1 entries, the last one zero)
static int scan_flags(mml_ctx* c, int* flags, int n) {
  MML_CUDA(c, cudaMemsetAsync(flags + n, 0, sizeof(int), c->stream));
  return exclusive_scan_device(c, flags, nullptr, n + 1, nullptr);
}
static int fetch_int(mml_ctx* c, const int* dev, int* host) {
  MML_CUDA(c, cudaMemcpyAsync(host, dev, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  MML_CUDA(c, cudaStreamSynchronize(c->stream));
  return MML_OK;
}
static int upload_bytes(mml_ctx* c, mml::DevBuf& dst, const void* src, size_t bytes) {
  MML_CUDA(c, dst.reserve(bytes ? bytes : 16));
  if (bytes) MML_CUDA(c, cudaMemcpyAsync(dst.p, src, bytes, cudaMemcpyHostToDevice, c->stream));
  return MML_OK;
}

extern "C" {

int mml_unpack_custom_points(mml_ctx* c, const void* points19, int n, int used_line, float* xyzi_out, uint16_t* line_out,
                             float* s_out, int* m_out) {
  if (!c || n < 0 || (n && !points19) || !m_out) return MML_ERR_INVALID;
  *m_out = 0;
  if (n == 0) return MML_OK;
  cudaSetDevice(c->device);
  cudaStream_t st = c->stream;
  MML_CHECK(upload_bytes(c, c->msg_raw, points19, (size_t)kCustomPointBytes * (size_t)n));
  MML_CUDA(c, c->tmp_a.reserve(sizeof(int) * ((size_t)n + 1)));
  MML_CUDA(c, c->in_xyzi.reserve(sizeof(float4) * (size_t)n));
  MML_CUDA(c, c->in_line.reserve(sizeof(uint16_t) * (size_t)n));
  MML_CUDA(c, c->in_s.reserve(sizeof(float) * (size_t)n));
  int* flag = c->tmp_a.as<int>();
  k_custom_flags<<<div_up(n, 256), 256, 0, st>>>(c->msg_raw.as<uint8_t>(), n, used_line, flag);
  MML_LAUNCHED(c);
  MML_CHECK(scan_flags(c, flag, n));
  k_custom_scatter<<<div_up(n, 256), 256, 0, st>>>(c->msg_raw.as<uint8_t>(), n, flag, c->in_xyzi.as<float4>(), c->in_line.as<uint16_t>(),
                                                   c->in_s.as<float>());
  MML_LAUNCHED(c);
  MML_CUDA(c, cudaGetLastError());
  int m = 0;
  MML_CHECK(fetch_int(c, flag + n, &m));
  *m_out = m;
  if (m && xyzi_out) MML_CUDA(c, cudaMemcpyAsync(xyzi_out, c->in_xyzi.p, sizeof(float4) * (size_t)m, cudaMemcpyDeviceToHost, st));
  if (m && line_out) MML_CUDA(c, cudaMemcpyAsync(line_out, c->in_line.p, sizeof(uint16_t) * (size_t)m, cudaMemcpyDeviceToHost, st));
  if (m && s_out) MML_CUDA(c, cudaMemcpyAsync(s_out, c->in_s.p, sizeof(float) * (size_t)m, cudaMemcpyDeviceToHost, st));
  MML_CUDA(c, cudaStreamSynchronize(st));
  return MML_OK;
}

int mml_unpack_pointcloud2(mml_ctx* c, const void* data, int n, int point_step, int off_x, int off_y, int off_z, int off_intensity,
                           float* xyzi_out, int* m_out) {
  if (!c || n < 0 || (n && !data) || !m_out || point_step < 12 || off_x < 0 || off_y < 0 || off_z < 0 || off_x + 4 > point_step ||
      off_y + 4 > point_step || off_z + 4 > point_step || off_intensity + 4 > point_step)
    return MML_ERR_INVALID;
  *m_out = 0;
  if (n == 0) return MML_OK;
  cudaSetDevice(c->device);
  cudaStream_t st = c->stream;
  MML_CHECK(upload_bytes(c, c->msg_raw, data, (size_t)point_step * (size_t)n));
  MML_CUDA(c, c->tmp_a.reserve(sizeof(int) * ((size_t)n + 1)));
  MML_CUDA(c, c->in_xyzi.reserve(sizeof(float4) * (size_t)n));
  int* flag = c->tmp_a.as<int>();
  k_pc2_flags<<<div_up(n, 256), 256, 0, st>>>(c->msg_raw.as<uint8_t>(), n, point_step, off_x, off_y, off_z, flag);
  MML_LAUNCHED(c);
  MML_CHECK(scan_flags(c, flag, n));
  k_pc2_scatter<<<div_up(n, 256), 256, 0, st>>>(c->msg_raw.as<uint8_t>(), n, point_step, off_x, off_y, off_z, off_intensity, flag,
                                                c->in_xyzi.as<float4>());
  MML_LAUNCHED(c);
  MML_CUDA(c, cudaGetLastError());
  int m = 0;
  MML_CHECK(fetch_int(c, flag + n, &m));
  *m_out = m;
  if (m && xyzi_out) MML_CUDA(c, cudaMemcpyAsync(xyzi_out, c->in_xyzi.p, sizeof(float4) * (size_t)m, cudaMemcpyDeviceToHost, st));
  MML_CUDA(c, cudaStreamSynchronize(st));
  return MML_OK;
}

int mml_pack_union_clouds(mml_ctx* c, const float* xyzi, const float* s, const uint16_t* line, const uint8_t* label, int n,
                          float near_full, float far_full, float near_feat, float far_feat, int zero_full_intensity, void* full_out,
                          void* corner_out, void* surf_out, int* counts3) {
  if (!c || n < 0 || (n && (!xyzi || !line || !label)) || !counts3) return MML_ERR_INVALID;
  counts3[0] = counts3[1] = counts3[2] = 0;
  if (n == 0) return MML_OK;
  cudaSetDevice(c->device);
  cudaStream_t st = c->stream;
  MML_CHECK(upload_bytes(c, c->in_xyzi, xyzi, sizeof(float4) * (size_t)n));
  MML_CHECK(upload_bytes(c, c->in_line, line, sizeof(uint16_t) * (size_t)n));
  MML_CHECK(upload_bytes(c, c->in_label, label, (size_t)n));
  if (s) MML_CHECK(upload_bytes(c, c->in_s, s, sizeof(float) * (size_t)n));
  MML_CUDA(c, c->tmp_a.reserve(sizeof(int) * ((size_t)n + 1)));
  MML_CUDA(c, c->tmp_b.reserve(sizeof(int) * ((size_t)n + 1)));
  MML_CUDA(c, c->tmp_c.reserve(sizeof(int) * ((size_t)n + 1)));
  MML_CUDA(c, c->msg_raw.reserve(sizeof(float4) * 3 * 3 * (size_t)n));  // three clouds of at most n records of 48 bytes
  int *ff = c->tmp_a.as<int>(), *fc = c->tmp_b.as<int>(), *fs = c->tmp_c.as<int>();
  PackCuts C;
  C.near_full2 = near_full * near_full; C.far_full2 = far_full * far_full; C.far_full_on = far_full > 0.f;
  C.near_feat2 = near_feat * near_feat; C.far_feat2 = far_feat * far_feat; C.far_feat_on = far_feat > 0.f;
  k_pack_flags<<<div_up(n, 256), 256, 0, st>>>(c->in_xyzi.as<float4>(), c->in_label.as<uint8_t>(), n, C, ff, fc, fs);
  MML_LAUNCHED(c);
  MML_CHECK(scan_flags(c, ff, n));
  MML_CHECK(scan_flags(c, fc, n));
  MML_CHECK(scan_flags(c, fs, n));
  float4* full = c->msg_raw.as<float4>();
  float4* corner = full + 3 * (size_t)n;
  float4* surf = corner + 3 * (size_t)n;
  k_pack_scatter<<<div_up(n, 256), 256, 0, st>>>(c->in_xyzi.as<float4>(), s ? c->in_s.as<float>() : nullptr, c->in_line.as<uint16_t>(),
                                                 c->in_label.as<uint8_t>(), n, ff, fc, fs, zero_full_intensity, full, corner, surf);
  MML_LAUNCHED(c);
  MML_CUDA(c, cudaGetLastError());
  int* hc = counts3;
  MML_CUDA(c, cudaMemcpyAsync(hc, ff + n, sizeof(int), cudaMemcpyDeviceToHost, st));
  MML_CUDA(c, cudaMemcpyAsync(hc + 1, fc + n, sizeof(int), cudaMemcpyDeviceToHost, st));
  MML_CUDA(c, cudaMemcpyAsync(hc + 2, fs + n, sizeof(int), cudaMemcpyDeviceToHost, st));
  MML_CUDA(c, cudaStreamSynchronize(st));
  if (full_out && hc[0]) MML_CUDA(c, cudaMemcpyAsync(full_out, full, 48 * (size_t)hc[0], cudaMemcpyDeviceToHost, st));
  if (corner_out && hc[1]) MML_CUDA(c, cudaMemcpyAsync(corner_out, corner, 48 * (size_t)hc[1], cudaMemcpyDeviceToHost, st));
  if (surf_out && hc[2]) MML_CUDA(c, cudaMemcpyAsync(surf_out, surf, 48 * (size_t)hc[2], cudaMemcpyDeviceToHost, st));
  MML_CUDA(c, cudaStreamSynchronize(st));
  return MML_OK;
}

}  // extern "C"
