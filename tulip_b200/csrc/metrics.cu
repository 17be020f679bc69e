// Evaluation metrics on the GPU: range image -> point cloud, voxel IoU / precision / recall / F1 and Chamfer
// distance.  Reference: tulip/util/evaluation.py:52-116 (img_to_pcd_kitti / img_to_pcd_carla), :125-134 (chamfer_distance),
// :148-175 (voxelize_point_cloud, calculate_metrics), driven by engine_upsampling.py:223-276.  The reference does all of this on
// the host with numpy (dense boolean grids of ~1600 x 1600 x 300 voxels per KITTI frame) plus a third-party Chamfer extension.
#include "common.cuh"
#include "kernels.h"

namespace {

// ---- projection: x = (sin_h[w] cos_v[h]) r, y = (cos_h[w] cos_v[h]) r, z = sin_v[h] r, r = img * max_range -----------------------
// float32 multiplies only, in the reference's order, on angle tables the host computed with the reference's own numpy expressions:
// bit-exact with evaluation.py:75-84 / :106-114.
__global__ void __launch_bounds__(256) range_to_points_kernel(const float* __restrict__ img, const float* __restrict__ sin_h,
                                                              const float* __restrict__ cos_h, const float* __restrict__ sin_v,
                                                              const float* __restrict__ cos_v, float max_range, float* __restrict__ pts,
                                                              long n, int H, int W) {
  pdl_sync();
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const int w = (int)(i % W), h = (int)((i / W) % H);
    const float r = __fmul_rn(img[i], max_range);
    const float cv = cos_v[h];
    pts[3 * i] = __fmul_rn(__fmul_rn(sin_h[w], cv), r);
    pts[3 * i + 1] = __fmul_rn(__fmul_rn(cos_h[w], cv), r);
    pts[3 * i + 2] = __fmul_rn(sin_v[h], r);
  }
}

// DurLAR (Ouster OS1-128) projection, evaluation.py:19-50: r = img * max_range and r - origin_offset in float32, everything after in
// float64 on host tables of cos / sin(encoder + azimuth), cos / sin(encoder) per column and cos / sin(elevation) per row, in the
// reference's multiplication order, no contraction; the point of pixel (row, col) lands at row * W + (col + W - offset_lut[row]) % W.
__global__ void __launch_bounds__(256) range_to_points_durlar_kernel(const float* __restrict__ img, const double* __restrict__ ca,
                                                                     const double* __restrict__ sa, const double* __restrict__ ce,
                                                                     const double* __restrict__ se, const double* __restrict__ cel,
                                                                     const double* __restrict__ sel, const int* __restrict__ offset_lut,
                                                                     float max_range, float origin_offset_f, double origin_offset,
                                                                     double z_offset, double* __restrict__ pts, long n, int H, int W) {
  pdl_sync();
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const int w = (int)(i % W), h = (int)((i / W) % H);
    const long b = i / ((long)W * H);
    const double rr = (double)__fsub_rn(__fmul_rn(img[i], max_range), origin_offset_f);
    const double x = __dadd_rn(__dmul_rn(__dmul_rn(rr, ca[w]), cel[h]), __dmul_rn(origin_offset, ce[w]));
    const double y = __dadd_rn(__dmul_rn(__dmul_rn(rr, sa[w]), cel[h]), __dmul_rn(origin_offset, se[w]));
    const double z = __dmul_rn(rr, sel[h]);
    const long o = (b * H + h) * W + (w + W - offset_lut[h]) % W;
    pts[3 * o] = -x;
    pts[3 * o + 1] = -y;
    pts[3 * o + 2] = __dadd_rn(z, z_offset);
  }
}

// ---- voxel metrics ---------------------------------------------------------------------------------------------------------------
// workspace layout (8-byte words): [0..5] min xyz, max xyz as floats (two per word is avoided: one float per word, low half),
// [8..10] counters |A|, |B|, |A & B|, then two open-addressing key tables of `cap` words each.
struct VoxelWs {
  float* mm;                   // 6 floats: min x,y,z, max x,y,z
  unsigned long long* cnt;     // 3 counters
  long long* ta;               // keys of the predicted cloud
  long long* tb;               // keys of the ground-truth cloud
  int cap;
};

// monotone maps of the coordinates to unsigned integers for atomicMin / atomicMax (float: 32-bit image in a 64-bit slot)
__device__ __forceinline__ unsigned long long to_ordered(float f) {
  const unsigned int u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ unsigned long long to_ordered(double f) {
  const unsigned long long u = (unsigned long long)__double_as_longlong(f);
  return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ void from_ordered(unsigned long long u, float* f) {
  const unsigned int w = (unsigned int)u;
  *f = __uint_as_float((w & 0x80000000u) ? (w & 0x7fffffffu) : ~w);
}
__device__ __forceinline__ void from_ordered(unsigned long long u, double* f) {
  *f = __longlong_as_double((long long)((u >> 63) ? (u & 0x7fffffffffffffffull) : ~u));
}

__global__ void voxel_init_kernel(unsigned long long* mm_ord, unsigned long long* cnt, long long* tables, long words, int is_f32) {
  pdl_sync();
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 3) { mm_ord[i] = is_f32 ? 0xffffffffull : ~0ull; mm_ord[3 + i] = 0ull; cnt[i] = 0ull; }
  for (long k = i; k < words; k += (long)gridDim.x * blockDim.x) tables[k] = -1ll;
}

template <class T>
__global__ void __launch_bounds__(256) voxel_minmax_kernel(const T* __restrict__ a, const T* __restrict__ b, int n,
                                                           unsigned long long* mm_ord) {
  pdl_sync();
  T mn[3] = {(T)INFINITY, (T)INFINITY, (T)INFINITY}, mx[3] = {(T)-INFINITY, (T)-INFINITY, (T)-INFINITY};
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < 2l * n; i += (long)gridDim.x * blockDim.x) {
    const T* p = (i < n) ? a + 3 * i : b + 3 * (i - n);
#pragma unroll
    for (int k = 0; k < 3; ++k) { mn[k] = p[k] < mn[k] ? p[k] : mn[k]; mx[k] = p[k] > mx[k] ? p[k] : mx[k]; }
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    for (int o = 16; o > 0; o >>= 1) {
      const T omn = __shfl_xor_sync(0xffffffffu, mn[k], o), omx = __shfl_xor_sync(0xffffffffu, mx[k], o);
      mn[k] = omn < mn[k] ? omn : mn[k];
      mx[k] = omx > mx[k] ? omx : mx[k];
    }
    if ((threadIdx.x & 31) == 0) { atomicMin(mm_ord + k, to_ordered(mn[k])); atomicMax(mm_ord + 3 + k, to_ordered(mx[k])); }
  }
}

// voxel key exactly as voxelize_point_cloud (evaluation.py:148-159): dims = int((max - min) / g) + 1, idx = int((p - min) / g), in
// the precision of the clouds (float32 for kitti / carla, float64 for durlar) with IEEE division; the dense-grid position becomes a
// linear int64 key
__device__ __forceinline__ float vsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ double vsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ float vdiv(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ double vdiv(double a, double b) { return __ddiv_rn(a, b); }
template <class T>
__device__ __forceinline__ long long voxel_key(const T* p, const T* mn, const T* mx, T g) {
  const int dy = (int)vdiv(vsub(mx[1], mn[1]), g) + 1, dz = (int)vdiv(vsub(mx[2], mn[2]), g) + 1;
  const int ix = (int)vdiv(vsub(p[0], mn[0]), g), iy = (int)vdiv(vsub(p[1], mn[1]), g), iz = (int)vdiv(vsub(p[2], mn[2]), g);
  return ((long long)ix * dy + iy) * dz + iz;
}
__device__ __forceinline__ unsigned int hash64(unsigned long long k) {
  k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
  return (unsigned int)k;
}

template <class T>
__global__ void __launch_bounds__(256) voxel_insert_kernel(const T* __restrict__ a, const T* __restrict__ b, int n, T g,
                                                           const unsigned long long* mm_ord, long long* ta, long long* tb, int cap,
                                                           unsigned long long* cnt) {
  pdl_sync();
  T mn[3], mx[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) { from_ordered(mm_ord[k], &mn[k]); from_ordered(mm_ord[3 + k], &mx[k]); }
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < 2l * n; i += (long)gridDim.x * blockDim.x) {
    const bool first = i < n;
    const T* p = first ? a + 3 * i : b + 3 * (i - n);
    long long* tab = first ? ta : tb;
    const long long key = voxel_key(p, mn, mx, g);
    unsigned int slot = hash64((unsigned long long)key) & (cap - 1);
    while (true) {
      const long long prev = (long long)atomicCAS(reinterpret_cast<unsigned long long*>(tab + slot), (unsigned long long)-1ll,
                                                  (unsigned long long)key);
      if (prev == -1ll) { atomicAdd(cnt + (first ? 0 : 1), 1ull); break; }      // new occupied voxel of this cloud
      if (prev == key) break;
      slot = (slot + 1) & (cap - 1);
    }
  }
}

__global__ void __launch_bounds__(256) voxel_intersect_kernel(const long long* ta, const long long* tb, int cap, unsigned long long* cnt) {
  pdl_sync();
  unsigned int local = 0;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += (long)gridDim.x * blockDim.x) {
    const long long key = ta[i];
    if (key == -1ll) continue;
    unsigned int slot = hash64((unsigned long long)key) & (cap - 1);
    while (true) {
      const long long other = tb[slot];
      if (other == key) { ++local; break; }
      if (other == -1ll) break;
      slot = (slot + 1) & (cap - 1);
    }
  }
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  if ((threadIdx.x & 31) == 0 && local) atomicAdd(cnt + 2, (unsigned long long)local);
}

__global__ void voxel_finalize_kernel(const unsigned long long* cnt, double* out4) {    // calculate_metrics, evaluation.py:161-175
  pdl_sync();
  const double pa = (double)cnt[0], gb = (double)cnt[1], inter = (double)cnt[2];
  const double iou = inter / (pa + gb - inter), precision = inter / pa, recall = inter / gb;
  out4[0] = iou; out4[1] = precision; out4[2] = recall;
  out4[3] = 2.0 * (precision * recall) / (precision + recall);                           // engine_upsampling.py:277
}

// ---- Chamfer distance: nearest-neighbour squared distance, brute force through shared-memory tiles -------------------------------
constexpr int CD_TILE = 1024;
__global__ void __launch_bounds__(256) chamfer_nn_kernel(const float* __restrict__ q, const float* __restrict__ t, int nq, int nt,
                                                         float* __restrict__ dist) {
  pdl_sync();
  __shared__ float st[3 * CD_TILE];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool ok = i < nq;
  const float qx = ok ? q[3 * i] : 0.f, qy = ok ? q[3 * i + 1] : 0.f, qz = ok ? q[3 * i + 2] : 0.f;
  float best = INFINITY;
  for (int t0 = 0; t0 < nt; t0 += CD_TILE) {
    const int m = min(CD_TILE, nt - t0);
    __syncthreads();
    for (int k = threadIdx.x; k < 3 * m; k += blockDim.x) st[k] = t[3 * (long)t0 + k];
    __syncthreads();
#pragma unroll 4
    for (int k = 0; k < m; ++k) {
      const float dx = __fsub_rn(qx, st[3 * k]), dy = __fsub_rn(qy, st[3 * k + 1]), dz = __fsub_rn(qz, st[3 * k + 2]);
      const float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));    // no FMA contraction
      best = fminf(best, d);
    }
  }
  if (ok) dist[i] = best;
}

__global__ void __launch_bounds__(256) mean_pair_kernel(const float* a, int na, const float* b, int nb, float* out3) {
  pdl_sync();
  __shared__ double red[2][8];
  double sa = 0.0, sb = 0.0;
  for (int i = threadIdx.x; i < na; i += blockDim.x) sa += a[i];
  for (int i = threadIdx.x; i < nb; i += blockDim.x) sb += b[i];
  for (int o = 16; o > 0; o >>= 1) { sa += __shfl_xor_sync(0xffffffffu, sa, o); sb += __shfl_xor_sync(0xffffffffu, sb, o); }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = sa; red[1][threadIdx.x >> 5] = sb; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ta = 0.0, tb = 0.0;
    for (int w = 0; w < 8; ++w) { ta += red[0][w]; tb += red[1][w]; }
    const float ma = (float)(ta / na), mb = (float)(tb / nb);
    out3[0] = ma + mb; out3[1] = ma; out3[2] = mb;                                       // evaluation.py:132
  }
}

int pow2_at_least(long v) { int p = 1024; while (p < v) p <<= 1; return p; }

}  // namespace

int range_to_points(const float* img, const float* sin_h, const float* cos_h, const float* sin_v, const float* cos_v, float max_range,
                    float* points, int B, int H, int W, cudaStream_t st) {
  TULIP_REQUIRE(B > 0 && H > 0 && W > 0, "range_to_points: empty image");
  const long n = (long)B * H * W;
  const int grid = (int)((n + 255) / 256 < 4l * tulip_num_sms() ? (n + 255) / 256 : 4l * tulip_num_sms());
  tulip_launch(range_to_points_kernel, grid, 256, 0, st, img, sin_h, cos_h, sin_v, cos_v, max_range, points, n, H, W);
  TULIP_CHECK_LAUNCH();
  return TULIP_OK;
}

long voxel_metrics_workspace_bytes(int n) { return 128 + 2l * pow2_at_least(4l * n) * 8; }

template <class T>
int voxel_metrics_t(const T* pts_pred, const T* pts_gt, int n, T grid_size, void* workspace, double* out4, cudaStream_t st) {
  TULIP_REQUIRE(n > 0 && grid_size > (T)0, "voxel_metrics: empty cloud or non-positive grid size");
  const int cap = pow2_at_least(4l * n);
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  unsigned long long* mm_ord = reinterpret_cast<unsigned long long*>(ws);
  unsigned long long* cnt = reinterpret_cast<unsigned long long*>(ws + 64);
  long long* ta = reinterpret_cast<long long*>(ws + 128);
  long long* tb = ta + cap;
  const int grid = 2 * tulip_num_sms();
  tulip_launch(voxel_init_kernel, grid, 256, 0, st, mm_ord, cnt, ta, 2l * cap, (int)(sizeof(T) == 4));
  tulip_launch(voxel_minmax_kernel<T>, grid, 256, 0, st, pts_pred, pts_gt, n, mm_ord);
  tulip_launch(voxel_insert_kernel<T>, grid, 256, 0, st, pts_pred, pts_gt, n, grid_size, (const unsigned long long*)mm_ord, ta, tb, cap, cnt);
  tulip_launch(voxel_intersect_kernel, grid, 256, 0, st, (const long long*)ta, (const long long*)tb, cap, cnt);
  tulip_launch(voxel_finalize_kernel, 1, 1, 0, st, (const unsigned long long*)cnt, out4);
  TULIP_CHECK_LAUNCH();
  return TULIP_OK;
}

int voxel_metrics(const float* pts_pred, const float* pts_gt, int n, float grid_size, void* workspace, double* out4, cudaStream_t st) {
  return voxel_metrics_t<float>(pts_pred, pts_gt, n, grid_size, workspace, out4, st);
}
int voxel_metrics_f64(const double* pts_pred, const double* pts_gt, int n, double grid_size, void* workspace, double* out4, cudaStream_t st) {
  return voxel_metrics_t<double>(pts_pred, pts_gt, n, grid_size, workspace, out4, st);
}

int range_to_points_durlar(const float* img, const double* ca, const double* sa, const double* ce, const double* se, const double* cel,
                           const double* sel, const int* offset_lut, float max_range, double origin_offset, double z_offset,
                           double* points, int B, int H, int W, cudaStream_t st) {
  TULIP_REQUIRE(B > 0 && H > 0 && W > 0, "range_to_points_durlar: empty image");
  const long n = (long)B * H * W;
  const int grid = (int)((n + 255) / 256 < 4l * tulip_num_sms() ? (n + 255) / 256 : 4l * tulip_num_sms());
  tulip_launch(range_to_points_durlar_kernel, grid, 256, 0, st, img, ca, sa, ce, se, cel, sel, offset_lut, max_range, (float)origin_offset,
               origin_offset, z_offset, points, n, H, W);
  TULIP_CHECK_LAUNCH();
  return TULIP_OK;
}

int chamfer_distance(const float* a, const float* b, int na, int nb, float* dist_a, float* dist_b, float* out3, cudaStream_t st) {
  TULIP_REQUIRE(na > 0 && nb > 0, "chamfer_distance: empty cloud");
  tulip_launch(chamfer_nn_kernel, ceil_div(na, 256), 256, 0, st, a, b, na, nb, dist_a);
  tulip_launch(chamfer_nn_kernel, ceil_div(nb, 256), 256, 0, st, b, a, nb, na, dist_b);
  tulip_launch(mean_pair_kernel, 1, 256, 0, st, (const float*)dist_a, na, (const float*)dist_b, nb, out3);
  TULIP_CHECK_LAUNCH();
  return TULIP_OK;
}
