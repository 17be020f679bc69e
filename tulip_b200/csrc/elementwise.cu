// HBM-bound kernels of the TULIP path: LayerNorm (+ PatchMerging gather), PatchEmbed, weight repack,
// loss, small element-wise helpers and the stand-alone index ops.  All of them move each byte once,
// with 16-byte vector accesses and grids sized in multiples of the SM count.
#include <cuda_fp16.h>

#include "common.cuh"
#include "kernels.h"
#include "window_index.cuh"

namespace {

// ------------------------------------------------------------------------------------------------
// LayerNorm.  A group of LANES lanes owns one row; each lane holds NCH 16-byte chunks (8 bf16) of it.
// Reference: nn.LayerNorm(eps=1e-6) at tulip.py:330,334 (block norms), :80,104 (PatchMerging, on the
// 2x2-gathered 4C vector), :569,720 (norm_up).  Statistics in fp32, two-pass in registers.

template <int LANES>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = LANES / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// source address of 16-byte chunk `ch` of row `row` (plain or PatchMerging gather, tulip.py:94-98)
__device__ __forceinline__ long ln_src_offset(const LnArgs& a, int row, int ch) {
  if (!a.gather) return (long)row * a.C + ch * 8;
  const int Cs = a.C >> 2;                       // source channels
  const int cps = Cs >> 3;                       // chunks per source token
  const int qd = ch / cps, cc = ch % cps;
  const int w2 = row % a.W2;
  const int bh = row / a.W2;
  const int h2 = bh % a.H2;
  const int b = bh / a.H2;
  const int hs = 2 * h2 + (qd & 1), ws = 2 * w2 + (qd >> 1);
  return ((long)(b * 2 * a.H2 + hs) * (2 * a.W2) + ws) * Cs + cc * 8;
}

template <int LANES, int NCH>
__global__ void __launch_bounds__(256) layernorm_fwd_kernel(const LnArgs a) {
  pdl_sync();
  const int lane = threadIdx.x % LANES;
  const int group = (blockIdx.x * blockDim.x + threadIdx.x) / LANES;
  const int ngroups = (gridDim.x * blockDim.x) / LANES;
  const int nch = a.C >> 3;
  const float invC = 1.0f / (float)a.C;
  for (int row = group; row < a.rows; row += ngroups) {
    uint4 v[NCH];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int ch = lane + i * LANES;
      if (ch < nch) {
        v[i] = *reinterpret_cast<const uint4*>(a.x + ln_src_offset(a, row, ch));
        const float2 p0 = unpack_bf16(v[i].x), p1 = unpack_bf16(v[i].y), p2 = unpack_bf16(v[i].z), p3 = unpack_bf16(v[i].w);
        sum += (p0.x + p0.y) + (p1.x + p1.y) + (p2.x + p2.y) + (p3.x + p3.y);
      }
    }
    const float mean = group_sum<LANES>(sum) * invC;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int ch = lane + i * LANES;
      if (ch < nch) {
        const uint32_t u[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 p = unpack_bf16(u[q]);
          sq += (p.x - mean) * (p.x - mean) + (p.y - mean) * (p.y - mean);
        }
      }
    }
    const float rstd = rsqrtf(group_sum<LANES>(sq) * invC + a.eps);
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int ch = lane + i * LANES;
      if (ch < nch) {
        const uint32_t u[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
        uint32_t o[4];
        const float4 w0 = *reinterpret_cast<const float4*>(a.w + ch * 8), w1 = *reinterpret_cast<const float4*>(a.w + ch * 8 + 4);
        const float4 b0 = *reinterpret_cast<const float4*>(a.b + ch * 8), b1 = *reinterpret_cast<const float4*>(a.b + ch * 8 + 4);
        const float ww[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 p = unpack_bf16(u[q]);
          o[q] = pack_bf16((p.x - mean) * rstd * ww[2 * q] + bb[2 * q], (p.y - mean) * rstd * ww[2 * q + 1] + bb[2 * q + 1]);
        }
        *reinterpret_cast<uint4*>(a.y + (long)row * a.C + ch * 8) = make_uint4(o[0], o[1], o[2], o[3]);
      }
    }
    if (lane == 0 && a.stats) *reinterpret_cast<float2*>(a.stats + 2 * (long)row) = make_float2(mean, rstd);
  }
}

// dx = [dres +] rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * w;  dw += dy * xhat; db += dy
template <int LANES, int NCH, bool REGACC>
__global__ void __launch_bounds__(256, (NCH <= 3) ? 2 : 1) layernorm_bwd_kernel(const LnArgs a) {
  pdl_sync();
  const int lane = threadIdx.x % LANES;
  const int group = (blockIdx.x * blockDim.x + threadIdx.x) / LANES;
  const int ngroups = (gridDim.x * blockDim.x) / LANES;
  const int nch = a.C >> 3;
  const float invC = 1.0f / (float)a.C;
  const int copy_off = a.dcopies > 1 ? (int)(blockIdx.x % a.dcopies) * a.dstride : 0;
  float* const gdw = a.dw + copy_off;
  float* const gdb = a.db + copy_off;
  float accw[REGACC ? NCH : 1][8], accb[REGACC ? NCH : 1][8];
  if (REGACC) {
#pragma unroll
    for (int i = 0; i < NCH; ++i)
#pragma unroll
      for (int q = 0; q < 8; ++q) accw[i][q] = accb[i][q] = 0.f;
  }
  for (int row = group; row < a.rows; row += ngroups) {
    const float2 st = *reinterpret_cast<const float2*>(a.stats + 2 * (long)row);
    const float mean = st.x, rstd = st.y;
    uint4 xv[NCH], gv[NCH], rv[NCH];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {                       // issue every load of the row before the first use
      const int ch = lane + i * LANES;
      rv[i] = make_uint4(0, 0, 0, 0);
      if (ch < nch) {
        const long off = ln_src_offset(a, row, ch);
        xv[i] = *reinterpret_cast<const uint4*>(a.x + off);
        gv[i] = *reinterpret_cast<const uint4*>(a.dy + (long)row * a.C + ch * 8);
        if (a.dres) rv[i] = *reinterpret_cast<const uint4*>(a.dres + off);
      }
    }
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int ch = lane + i * LANES;
      if (ch < nch) {
        const uint32_t xu[4] = {xv[i].x, xv[i].y, xv[i].z, xv[i].w};
        const uint32_t gu[4] = {gv[i].x, gv[i].y, gv[i].z, gv[i].w};
        const float4 w0 = *reinterpret_cast<const float4*>(a.w + ch * 8), w1 = *reinterpret_cast<const float4*>(a.w + ch * 8 + 4);
        const float ww[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 x = unpack_bf16(xu[q]), g = unpack_bf16(gu[q]);
          const float h0 = (x.x - mean) * rstd, h1 = (x.y - mean) * rstd;
          const float g0 = g.x * ww[2 * q], g1 = g.y * ww[2 * q + 1];
          s1 += g0 + g1;
          s2 += g0 * h0 + g1 * h1;
          if (REGACC) {
            accw[i][2 * q] += g.x * h0; accw[i][2 * q + 1] += g.y * h1;
            accb[i][2 * q] += g.x;      accb[i][2 * q + 1] += g.y;
          } else {
            atomicAdd(gdw + ch * 8 + 2 * q, g.x * h0); atomicAdd(gdw + ch * 8 + 2 * q + 1, g.y * h1);
            atomicAdd(gdb + ch * 8 + 2 * q, g.x);      atomicAdd(gdb + ch * 8 + 2 * q + 1, g.y);
          }
        }
      }
    }
    const float m1 = group_sum<LANES>(s1) * invC, m2 = group_sum<LANES>(s2) * invC;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int ch = lane + i * LANES;
      if (ch < nch) {
        const uint32_t xu[4] = {xv[i].x, xv[i].y, xv[i].z, xv[i].w};
        const uint32_t gu[4] = {gv[i].x, gv[i].y, gv[i].z, gv[i].w};
        const float4 w0 = *reinterpret_cast<const float4*>(a.w + ch * 8), w1 = *reinterpret_cast<const float4*>(a.w + ch * 8 + 4);
        const float ww[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
        const long off = ln_src_offset(a, row, ch);        // dx has the layout of the LN input (scatter for the merge)
        const uint32_t ru[4] = {rv[i].x, rv[i].y, rv[i].z, rv[i].w};
        uint32_t o[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 x = unpack_bf16(xu[q]), g = unpack_bf16(gu[q]), r = unpack_bf16(ru[q]);
          const float h0 = (x.x - mean) * rstd, h1 = (x.y - mean) * rstd;
          const float d0 = rstd * (g.x * ww[2 * q] - m1 - h0 * m2), d1 = rstd * (g.y * ww[2 * q + 1] - m1 - h1 * m2);
          o[q] = pack_bf16(r.x + d0, r.y + d1);
        }
        *reinterpret_cast<uint4*>(a.dx + off) = make_uint4(o[0], o[1], o[2], o[3]);
        if (a.dxs) {                                       // same values, scaled per sample (DropPath backward)
          const float sc = a.row_scale[row / a.rows_per_sample];
          uint32_t os[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float2 f = unpack_bf16(o[q]);
            os[q] = pack_bf16(f.x * sc, f.y * sc);
          }
          *reinterpret_cast<uint4*>(a.dxs + off) = make_uint4(os[0], os[1], os[2], os[3]);
        }
      }
    }
  }
  if (REGACC) {
    // column sums: shuffle-reduce over the row groups that share a warp, one smem slab per warp, then one global
    // atomic per column per CTA
    extern __shared__ float s_acc[];                      // [8 warps][2][C]
#pragma unroll
    for (int i = 0; i < NCH; ++i)
#pragma unroll
      for (int q = 0; q < 8; ++q) {
#pragma unroll
        for (int o = LANES; o < 32; o <<= 1) {
          accw[i][q] += __shfl_xor_sync(0xffffffffu, accw[i][q], o);
          accb[i][q] += __shfl_xor_sync(0xffffffffu, accb[i][q], o);
        }
      }
    const int wl = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (wl < LANES) {
      float* slab = s_acc + wid * 2 * a.C;
#pragma unroll
      for (int i = 0; i < NCH; ++i) {
        const int ch = wl + i * LANES;
        if (ch < nch) {
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            slab[ch * 8 + q] = accw[i][q];
            slab[a.C + ch * 8 + q] = accb[i][q];
          }
        }
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * a.C; i += blockDim.x) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += s_acc[w * 2 * a.C + i];
      atomicAdd((i < a.C ? gdw : gdb - a.C) + i, t);
    }
  }
}

struct LnCfg { int lanes, nch; };
inline bool ln_config(int C, LnCfg* cfg) {
  if (C % 8) return false;
  const int nch = C / 8;
  // C = 96 * 2^k (every LayerNorm on the TULIP path): 3 chunks per lane, 4 * 2^k lanes per row -> no idle lanes
  if (C % 96 == 0) {
    const int m = C / 96;
    if (m == 1) { *cfg = {4, 3}; return true; }
    if (m == 2) { *cfg = {8, 3}; return true; }
    if (m == 4) { *cfg = {16, 3}; return true; }
    if (m == 8) { *cfg = {32, 3}; return true; }
    if (m == 16) { *cfg = {32, 6}; return true; }
    if (m == 32) { *cfg = {32, 12}; return true; }
  }
  if (nch <= 16) *cfg = {16, 1};
  else if (nch <= 32) *cfg = {16, 2};
  else if (nch <= 48) *cfg = {16, 3};
  else if (nch <= 96) *cfg = {32, 3};
  else if (nch <= 192) *cfg = {32, 6};
  else if (nch <= 384) *cfg = {32, 12};
  else return false;
  return true;
}

// resident CTAs per SM for a grid-stride kernel (so the grid is exactly one full wave)
template <class K>
int wave_grid(K kernel, int threads, int smem, long max_ctas) {
  int per_sm = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
  const long g = (long)per_sm * tulip_num_sms();
  return (int)(max_ctas < g ? (max_ctas > 0 ? max_ctas : 1) : g);
}

// ------------------------------------------------------------------------------------------------
// PatchEmbedding: circular pad W by (2,2) -> Conv2d(1 -> E, k=(ph,8), s=(ph,4)) -> NHWC -> LayerNorm.
// Reference tulip.py:59-73.  One warp per output token, lane l owns channels l, l+32, ...
template <int EPL>   // channels per lane = E / 32
__global__ void __launch_bounds__(256) patch_embed_fwd_kernel(const EmbedArgs a) {
  pdl_sync();
  extern __shared__ float s_w[];                          // [ph*8][E] conv weight (transposed: lanes read consecutive channels)
  const int KW = a.ph * 8;
  for (int i = threadIdx.x; i < a.E * KW; i += blockDim.x) s_w[(i % KW) * a.E + i / KW] = a.w[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int Ho = a.Himg / a.ph, Wo = a.Wimg / 4;
  const int tokens = a.B * Ho * Wo;
  float cb[EPL], lw[EPL], lb[EPL];
#pragma unroll
  for (int q = 0; q < EPL; ++q) { cb[q] = a.b[lane + 32 * q]; lw[q] = a.ln_w[lane + 32 * q]; lb[q] = a.ln_b[lane + 32 * q]; }
  const float invE = 1.0f / (float)a.E;
  for (int t = warp; t < tokens; t += nwarps) {
    const int wo = t % Wo;
    const int bh = t / Wo;
    const int ho = bh % Ho, b = bh / Ho;
    float u[EPL];
#pragma unroll
    for (int q = 0; q < EPL; ++q) u[q] = cb[q];
    for (int dh = 0; dh < a.ph; ++dh) {
      const float* row = a.x + ((long)b * a.Himg + ho * a.ph + dh) * a.Wimg;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        int wi = 4 * wo + j - 2;
        wi = wi < 0 ? wi + a.Wimg : (wi >= a.Wimg ? wi - a.Wimg : wi);
        const float xv = __ldg(row + wi);
#pragma unroll
        for (int q = 0; q < EPL; ++q) u[q] = fmaf(s_w[(dh * 8 + j) * a.E + lane + 32 * q], xv, u[q]);
      }
    }
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < EPL; ++q) s += u[q];
    const float mean = warp_sum(s) * invE;
    float sq = 0.f;
#pragma unroll
    for (int q = 0; q < EPL; ++q) sq += (u[q] - mean) * (u[q] - mean);
    const float rstd = rsqrtf(warp_sum(sq) * invE + a.eps);
#pragma unroll
    for (int q = 0; q < EPL; ++q)
      a.y[(long)t * a.E + lane + 32 * q] = __float2bfloat16_rn((u[q] - mean) * rstd * lw[q] + lb[q]);
  }
}

// Backward: recompute conv + LN statistics from the input image, then LN backward and the conv
// weight / bias gradients (no input gradient: the input is data).  ph == 1 only (all shipped configs).
template <int EPL>
__global__ void __launch_bounds__(256) patch_embed_bwd_kernel(const EmbedArgs a) {
  pdl_sync();
  extern __shared__ float s_w[];                          // [8][E] weights (transposed), then [E][12] accumulators
  float* s_acc = s_w + a.E * 8;
  for (int i = threadIdx.x; i < a.E * 8; i += blockDim.x) s_w[(i % 8) * a.E + i / 8] = a.w[i];
  for (int i = threadIdx.x; i < a.E * 12; i += blockDim.x) s_acc[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int Wo = a.Wimg / 4;
  const int tokens = a.B * a.Himg * Wo;
  float cb[EPL], lw[EPL];
  float gw[EPL][8], gb[EPL], glw[EPL], glb[EPL];
#pragma unroll
  for (int q = 0; q < EPL; ++q) {
    cb[q] = a.b[lane + 32 * q]; lw[q] = a.ln_w[lane + 32 * q];
    gb[q] = glw[q] = glb[q] = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) gw[q][j] = 0.f;
  }
  const float invE = 1.0f / (float)a.E;
  for (int t = warp; t < tokens; t += nwarps) {
    const int wo = t % Wo;
    const int bh = t / Wo;
    const float* row = a.x + (long)bh * a.Wimg;
    float xv[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int wi = 4 * wo + j - 2;
      wi = wi < 0 ? wi + a.Wimg : (wi >= a.Wimg ? wi - a.Wimg : wi);
      xv[j] = __ldg(row + wi);
    }
    float u[EPL], dy[EPL];
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < EPL; ++q) {
      u[q] = cb[q];
#pragma unroll
      for (int j = 0; j < 8; ++j) u[q] = fmaf(s_w[j * a.E + lane + 32 * q], xv[j], u[q]);
      s += u[q];
      dy[q] = __bfloat162float(a.dy[(long)t * a.E + lane + 32 * q]);
    }
    const float mean = warp_sum(s) * invE;
    float sq = 0.f;
#pragma unroll
    for (int q = 0; q < EPL; ++q) sq += (u[q] - mean) * (u[q] - mean);
    const float rstd = rsqrtf(warp_sum(sq) * invE + a.eps);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int q = 0; q < EPL; ++q) {
      const float h = (u[q] - mean) * rstd, g = dy[q] * lw[q];
      s1 += g; s2 += g * h;
      glw[q] += dy[q] * h; glb[q] += dy[q];
    }
    const float m1 = warp_sum(s1) * invE, m2 = warp_sum(s2) * invE;
#pragma unroll
    for (int q = 0; q < EPL; ++q) {
      const float h = (u[q] - mean) * rstd;
      const float du = rstd * (dy[q] * lw[q] - m1 - h * m2);
      gb[q] += du;
#pragma unroll
      for (int j = 0; j < 8; ++j) gw[q][j] = fmaf(du, xv[j], gw[q][j]);
    }
  }
#pragma unroll
  for (int q = 0; q < EPL; ++q) {
    const int c = lane + 32 * q;
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(&s_acc[c * 12 + j], gw[q][j]);
    atomicAdd(&s_acc[c * 12 + 8], gb[q]);
    atomicAdd(&s_acc[c * 12 + 9], glw[q]);
    atomicAdd(&s_acc[c * 12 + 10], glb[q]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < a.E * 12; i += blockDim.x) {
    const int c = i / 12, k = i % 12;
    const float v = s_acc[i];
    const int co = a.dcopies > 1 ? (int)(blockIdx.x % a.dcopies) * a.dstride : 0;
    if (k < 8) atomicAdd(a.dw + co + c * 8 + k, v);
    else if (k == 8) atomicAdd(a.db + co + c, v);
    else if (k == 9) atomicAdd(a.dln_w + co + c, v);
    else if (k == 10) atomicAdd(a.dln_b + co + c, v);
  }
}

// ------------------------------------------------------------------------------------------------
// PatchEmbedding, patch height 1, E = 12 * LPT (96 or 192): a group of LPT lanes owns a token and each lane 12 consecutive
// channels with its 12 x 8 conv weights in registers, so a warp finishes 32 / LPT tokens per iteration with two short
// group reductions.  The one-warp-per-token kernels above spend ~100 (forward) / ~190 (backward) instructions per token
// and were issue-bound (60 / 74 us for 27 MB of traffic); this layout needs about half.
template <int LPT>
__global__ void __launch_bounds__(128) patch_embed_fwd12_kernel(const EmbedArgs a) {
  pdl_sync();
  constexpr int TPW = 32 / LPT;
  const int lane = threadIdx.x & 31, sub = lane % LPT, slot = lane / LPT;
  const int c0 = sub * 12;
  float w[12][8], cb[12], lw[12], lb[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) {
    const float4 w0 = *reinterpret_cast<const float4*>(a.w + (c0 + i) * 8), w1 = *reinterpret_cast<const float4*>(a.w + (c0 + i) * 8 + 4);
    w[i][0] = w0.x; w[i][1] = w0.y; w[i][2] = w0.z; w[i][3] = w0.w; w[i][4] = w1.x; w[i][5] = w1.y; w[i][6] = w1.z; w[i][7] = w1.w;
    cb[i] = a.b[c0 + i]; lw[i] = a.ln_w[c0 + i]; lb[i] = a.ln_b[c0 + i];
  }
  const int Wo = a.Wimg / 4;
  const int tokens = a.B * a.Himg * Wo;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const float invE = 1.0f / (float)a.E;
  auto load_x = [&](int t, float (&xv)[8]) {
    const int tt = t < tokens ? t : tokens - 1;
    const int wo = tt % Wo;
    const float* row = a.x + (long)(tt / Wo) * a.Wimg;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int wi = 4 * wo + j - 2;
      wi = wi < 0 ? wi + a.Wimg : (wi >= a.Wimg ? wi - a.Wimg : wi);
      xv[j] = __ldg(row + wi);
    }
  };
  float xn[8];
  int t = warp * TPW + slot;
  load_x(t, xn);
  for (; t - slot < tokens; t += nwarps * TPW) {
    float xv[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) xv[j] = xn[j];
    load_x(t + nwarps * TPW, xn);                          // next iteration's pixels are in flight during this one's math
    float u[12], s = 0.f;
#pragma unroll
    for (int i = 0; i < 12; ++i) {
      float acc = cb[i];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc = fmaf(w[i][j], xv[j], acc);
      u[i] = acc; s += acc;
    }
    const float mean = group_sum<LPT>(s) * invE;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < 12; ++i) sq += (u[i] - mean) * (u[i] - mean);
    const float rstd = rsqrtf(group_sum<LPT>(sq) * invE + a.eps);
    if (t < tokens) {
      uint32_t o[6];
#pragma unroll
      for (int i = 0; i < 6; ++i)
        o[i] = pack_bf16((u[2 * i] - mean) * rstd * lw[2 * i] + lb[2 * i], (u[2 * i + 1] - mean) * rstd * lw[2 * i + 1] + lb[2 * i + 1]);
      uint2* dst = reinterpret_cast<uint2*>(a.y + (long)t * a.E + c0);
      dst[0] = make_uint2(o[0], o[1]); dst[1] = make_uint2(o[2], o[3]); dst[2] = make_uint2(o[4], o[5]);
    }
  }
}

// Backward, E = 96: 16 lanes per token, 6 channels per lane (weights and the 6 x 8 weight-gradient accumulators in registers).
__global__ void __launch_bounds__(128) patch_embed_bwd6_kernel(const EmbedArgs a) {
  pdl_sync();
  extern __shared__ float s_acc[];                        // [E][13] per-CTA accumulators: 8 dW taps, db, dgamma, dbeta (+2 pad)
  for (int i = threadIdx.x; i < a.E * 13; i += blockDim.x) s_acc[i] = 0.f;
  __syncthreads();
  constexpr int LPT = 16, TPW = 2;
  const int lane = threadIdx.x & 31, sub = lane % LPT, slot = lane / LPT;
  const int c0 = sub * 6;
  float w[6][8], cb[6], lw[6], gw[6][8], gb[6], glw[6], glb[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const float4 w0 = *reinterpret_cast<const float4*>(a.w + (c0 + i) * 8), w1 = *reinterpret_cast<const float4*>(a.w + (c0 + i) * 8 + 4);
    w[i][0] = w0.x; w[i][1] = w0.y; w[i][2] = w0.z; w[i][3] = w0.w; w[i][4] = w1.x; w[i][5] = w1.y; w[i][6] = w1.z; w[i][7] = w1.w;
    cb[i] = a.b[c0 + i]; lw[i] = a.ln_w[c0 + i];
    gb[i] = glw[i] = glb[i] = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) gw[i][j] = 0.f;
  }
  const int Wo = a.Wimg / 4;
  const int tokens = a.B * a.Himg * Wo;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const float invE = 1.0f / (float)a.E;
  // a warp's iterations are serialised on its own loads unless the next token's pixels and gradient row are already in flight
  auto load_tok = [&](int t, float (&xv)[8], uint32_t (&dv)[3]) {
    const int tt = t < tokens ? t : tokens - 1;
    const int wo = tt % Wo;
    const float* row = a.x + (long)(tt / Wo) * a.Wimg;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int wi = 4 * wo + j - 2;
      wi = wi < 0 ? wi + a.Wimg : (wi >= a.Wimg ? wi - a.Wimg : wi);
      xv[j] = __ldg(row + wi);
    }
    const uint32_t* dyp = reinterpret_cast<const uint32_t*>(a.dy + (long)tt * a.E + c0);
    dv[0] = dyp[0]; dv[1] = dyp[1]; dv[2] = dyp[2];
  };
  float xn[8]; uint32_t dn[3];
  load_tok(warp * TPW + slot, xn, dn);
  for (int t = warp * TPW + slot; t - slot < tokens; t += nwarps * TPW) {
    const bool valid = t < tokens;
    float xv[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) xv[j] = xn[j];
    const uint32_t d01 = dn[0], d23 = dn[1], d45 = dn[2];
    load_tok(t + nwarps * TPW, xn, dn);
    float dy[6];
    { const float2 f0 = unpack_bf16(d01), f1 = unpack_bf16(d23), f2 = unpack_bf16(d45);
      dy[0] = f0.x; dy[1] = f0.y; dy[2] = f1.x; dy[3] = f1.y; dy[4] = f2.x; dy[5] = f2.y; }
    if (!valid) {
#pragma unroll
      for (int i = 0; i < 6; ++i) dy[i] = 0.f;             // a padded slot contributes nothing
    }
    float u[6], s = 0.f;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      float acc = cb[i];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc = fmaf(w[i][j], xv[j], acc);
      u[i] = acc; s += acc;
    }
    const float mean = group_sum<LPT>(s) * invE;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < 6; ++i) sq += (u[i] - mean) * (u[i] - mean);
    const float rstd = rsqrtf(group_sum<LPT>(sq) * invE + a.eps);
    float s1 = 0.f, s2 = 0.f, h[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      h[i] = (u[i] - mean) * rstd;
      const float g = dy[i] * lw[i];
      s1 += g; s2 += g * h[i];
      glw[i] += dy[i] * h[i]; glb[i] += dy[i];
    }
    const float m1 = group_sum<LPT>(s1) * invE, m2 = group_sum<LPT>(s2) * invE;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const float du = valid ? rstd * (dy[i] * lw[i] - m1 - h[i] * m2) : 0.f;
      gb[i] += du;
#pragma unroll
      for (int j = 0; j < 8; ++j) gw[i][j] = fmaf(du, xv[j], gw[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 6; ++i) {                             // fold the warp's two token slots, then lanes 0..15 add to the CTA
#pragma unroll
    for (int j = 0; j < 8; ++j) gw[i][j] += __shfl_xor_sync(0xffffffffu, gw[i][j], 16);
    gb[i] += __shfl_xor_sync(0xffffffffu, gb[i], 16);
    glw[i] += __shfl_xor_sync(0xffffffffu, glw[i], 16);
    glb[i] += __shfl_xor_sync(0xffffffffu, glb[i], 16);
  }
  if (slot == 0) {
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const int c = c0 + i;                                 // rows of 13 floats: the 16 lanes hit 16 different banks
#pragma unroll
      for (int j = 0; j < 8; ++j) atomicAdd(&s_acc[c * 13 + j], gw[i][j]);
      atomicAdd(&s_acc[c * 13 + 8], gb[i]);
      atomicAdd(&s_acc[c * 13 + 9], glw[i]);
      atomicAdd(&s_acc[c * 13 + 10], glb[i]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < a.E * 11; i += blockDim.x) {
    const int c = i / 11, k = i % 11;
    const float v = s_acc[c * 13 + k];
    const int co = a.dcopies > 1 ? (int)(blockIdx.x % a.dcopies) * a.dstride : 0;
    if (k < 8) atomicAdd(a.dw + co + c * 8 + k, v);
    else if (k == 8) atomicAdd(a.db + co + c, v);
    else if (k == 9) atomicAdd(a.dln_w + co + c, v);
    else if (k == 10) atomicAdd(a.dln_b + co + c, v);
  }
}

// ------------------------------------------------------------------------------------------------
// fp32 -> bf16 weight repack, 32x32 tiles, optional row permutation and transposed copy.
__global__ void __launch_bounds__(256) pack_weights_kernel(const float* __restrict__ flat, bf16* __restrict__ arena,
                                                           const PackItem* __restrict__ items, int n_items) {
  pdl_sync();
  __shared__ float tile[32][33];
  __shared__ int s_item;
  if (threadIdx.x == 0) {
    int lo = 0, hi = n_items - 1;                          // last item with tile_begin <= blockIdx.x
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (items[mid].tile_begin <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
    }
    s_item = lo;
  }
  __syncthreads();
  const PackItem it = items[s_item];
  const int tiles_c = (it.cols + 31) / 32;
  const int tl = blockIdx.x - it.tile_begin;
  const int r0 = (tl / tiles_c) * 32, c0 = (tl % tiles_c) * 32;      // destination (permuted) row block
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = r0 + ty + 8 * i, c = c0 + tx;
    float v = 0.f;
    if (r < it.rows && c < it.cols) {
      const int sr = it.perm_R2 > 1 ? (r % it.perm_Cc) * it.perm_R2 + r / it.perm_Cc : r;
      v = flat[it.src_off + (long)sr * it.cols + c];
      arena[it.dst_off + (long)r * it.cols + c] = __float2bfloat16_rn(v);
    }
    tile[ty + 8 * i][tx] = v;
  }
  if (it.dstT_off < 0) return;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = c0 + ty + 8 * i, r = r0 + tx;
    if (r < it.rows && c < it.cols) arena[it.dstT_off + (long)c * it.rows + r] = __float2bfloat16_rn(tile[tx][ty + 8 * i]);
  }
}

__global__ void permute_bias_kernel(const float* __restrict__ src, float* __restrict__ dst, int n, int R2, int Cc) {
  pdl_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[(i % Cc) * R2 + i / Cc];
}

// ------------------------------------------------------------------------------------------------
__global__ void add_inplace_kernel(uint4* __restrict__ dst, const uint4* __restrict__ src, long n16) {
  pdl_sync();
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (long)gridDim.x * blockDim.x) {
    const uint4 a = dst[i], b = src[i];
    const uint32_t au[4] = {a.x, a.y, a.z, a.w}, bu[4] = {b.x, b.y, b.z, b.w};
    uint32_t o[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float2 x = unpack_bf16(au[q]), y = unpack_bf16(bu[q]);
      o[q] = pack_bf16(x.x + y.x, x.y + y.y);
    }
    dst[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

__global__ void scale_rows_kernel(uint4* __restrict__ dst, const uint4* __restrict__ src, const float* __restrict__ row_scale,
                                  long n16, int chunks_per_row, int rows_per_sample) {
  pdl_sync();
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (long)gridDim.x * blockDim.x) {
    const float s = row_scale[(i / chunks_per_row) / rows_per_sample];
    const uint4 a = src[i];
    const uint32_t au[4] = {a.x, a.y, a.z, a.w};
    uint32_t o[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float2 x = unpack_bf16(au[q]);
      o[q] = pack_bf16(x.x * s, x.y * s);
    }
    dst[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// mean |pred - y| and mean |expm1(pred) - expm1(y)| (tulip.py:690-700); acc2 must be zero on entry
__global__ void __launch_bounds__(256) l1_loss_kernel(const float* __restrict__ pred, const float* __restrict__ target, long n,
                                                      int log_transform, float* acc2) {
  pdl_sync();
  float s0 = 0.f, s1 = 0.f;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const float p = pred[i], y = target[i];
    s0 += fabsf(p - y);
    if (log_transform) s1 += fabsf(expm1f(p) - expm1f(y));
  }
  s0 = warp_sum(s0); s1 = warp_sum(s1);
  __shared__ float sh[2][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { sh[0][warp] = s0; sh[1][warp] = s1; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b = 0.f;
    for (int w = 0; w < 8; ++w) { a += sh[0][w]; b += sh[1][w]; }
    atomicAdd(acc2, a);
    atomicAdd(acc2 + 1, b);
  }
}
__global__ void l1_loss_finalize_kernel(const float* acc2, float* out2, float inv_n, int log_transform) {
  pdl_sync();
  out2[0] = acc2[0] * inv_n;
  out2[1] = log_transform ? acc2[1] * inv_n : acc2[0] * inv_n;
}

// ------------------------------------------------------------------------------------------------
// stand-alone index ops (16-byte chunks; C % 8 == 0)
__global__ void window_copy_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int B, WinGeom g, int cpt, bool scatter) {
  pdl_sync();
  const int L = g.Mh * g.Mw;
  const int nWh = g.H / g.Mh, nWw = g.W / g.Mw;
  const long total = (long)B * g.H * g.W * cpt;
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int c = idx % cpt;
    long wt = idx / cpt;                                   // windowed token index (B Nh Nw) L
    const int i = wt % L; wt /= L;
    const int ww = wt % nWw; wt /= nWw;
    const int wh = wt % nWh;
    const int b = wt / nWh;
    const long nat = (long)win_token_index(g, b, wh, ww, i) * cpt + c;
    if (scatter) dst[nat] = src[idx]; else dst[idx] = src[nat];
  }
}

__global__ void shift_mask_kernel(float* out, WinGeom g) {
  pdl_sync();
  const int L = g.Mh * g.Mw;
  const int nWw = g.W / g.Mw;
  const int total = (g.H / g.Mh) * nWw * L * L;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int j = idx % L, i = (idx / L) % L, win = idx / (L * L);
    const int wh = win / nWw, ww = win % nWw;
    out[idx] = win_region_id(g, wh, ww, i) != win_region_id(g, wh, ww, j) ? -100.0f : 0.0f;
  }
}

__global__ void rel_bias_gather_kernel(const float* table, float* out, int heads, int Mh, int Mw) {
  pdl_sync();
  const int L = Mh * Mw;
  const int total = heads * L * L;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int j = idx % L, i = (idx / L) % L, h = idx / (L * L);
    out[idx] = table[rel_bias_index(Mh, Mw, i, j) * heads + h];
  }
}

__global__ void merge_gather_kernel(const uint4* __restrict__ x, uint4* __restrict__ out, LnArgs a, long total) {
  pdl_sync();
  const int nch = a.C >> 3;
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int ch = idx % nch;
    const int row = idx / nch;
    out[idx] = x[ln_src_offset(a, row, ch) >> 3];
  }
}

__global__ void pixel_shuffle_kernel(const bf16* __restrict__ x, bf16* __restrict__ out, int B, int H, int W, int Cout, int r) {
  pdl_sync();
  // out[b, h*r+i, w*r+j, c] = x[b, h, w, c*r*r + i*r + j]
  const long total = (long)B * H * W * Cout * r * r;
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int c = idx % Cout;
    long p = idx / Cout;
    const int wo = p % (W * r); p /= (W * r);
    const int ho = p % (H * r);
    const int b = p / (H * r);
    const int h = ho / r, i = ho % r, w = wo / r, j = wo % r;
    out[idx] = x[(((long)b * H + h) * W + w) * (Cout * r * r) + c * r * r + i * r + j];
  }
}

inline int ew_grid(long n, int threads) {
  const long want = (n + threads - 1) / threads;
  const long cap = (long)tulip_num_sms() * 8;
  return (int)(want < cap ? (want > 0 ? want : 1) : cap);
}

}  // namespace

#define LN_CASE(L, N) (cfg.lanes == L && cfg.nch == N)

int layernorm_fwd(const LnArgs& a, cudaStream_t st) {
  LnCfg cfg;
  TULIP_REQUIRE(ln_config(a.C, &cfg), "layernorm: C must be a multiple of 8 and <= 3072");
  TULIP_REQUIRE(a.rows > 0, "layernorm: empty input");
  const long want = ceil_div(a.rows, 256 / cfg.lanes);
#define LN_FWD(L, N) { const int grid = wave_grid(layernorm_fwd_kernel<L, N>, 256, 0, want); tulip_launch(layernorm_fwd_kernel<L, N>, grid, 256, 0, st, a); }
  if LN_CASE(4, 3) LN_FWD(4, 3)
  else if LN_CASE(8, 3) LN_FWD(8, 3)
  else if LN_CASE(16, 1) LN_FWD(16, 1)
  else if LN_CASE(16, 2) LN_FWD(16, 2)
  else if LN_CASE(16, 3) LN_FWD(16, 3)
  else if LN_CASE(32, 3) LN_FWD(32, 3)
  else if LN_CASE(32, 6) LN_FWD(32, 6)
  else LN_FWD(32, 12)
#undef LN_FWD
  TULIP_CHECK_LAUNCH();
  return TULIP_OK;
}

int layernorm_bwd(const LnArgs& a, cudaStream_t st) {
  LnCfg cfg;
  TULIP_REQUIRE(ln_config(a.C, &cfg), "layernorm: C must be a multiple of 8 and <= 3072");
  TULIP_REQUIRE(a.rows > 0, "layernorm: empty input");
  const long want = ceil_div(a.rows, 256 / cfg.lanes);
  const int smem = 8 * 2 * a.C * (int)sizeof(float);
#define LN_BWD(L, N, R) { const int sm = (R) ? smem : 0; \
    if (sm > 48 * 1024) TULIP_CUDA(cudaFuncSetAttribute(layernorm_bwd_kernel<L, N, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm)); \
    const int grid = wave_grid(layernorm_bwd_kernel<L, N, R>, 256, sm, want); \
    tulip_launch(layernorm_bwd_kernel<L, N, R>, grid, 256, sm, st, a); }
  if LN_CASE(4, 3) LN_BWD(4, 3, true)
  else if LN_CASE(8, 3) LN_BWD(8, 3, true)
  else if LN_CASE(16, 1) LN_BWD(16, 1, true)
  else if LN_CASE(16, 2) LN_BWD(16, 2, true)
  else if LN_CASE(16, 3) LN_BWD(16, 3, true)
  else if LN_CASE(32, 3) LN_BWD(32, 3, true)
  else if LN_CASE(32, 6) LN_BWD(32, 6, true)
  else LN_BWD(32, 12, false)
#undef LN_BWD
  TULIP_CHECK_LAUNCH();
  return TULIP_OK;
}

int patch_embed_fwd(const EmbedArgs& a, cudaStream_t st) {
  TULIP_REQUIRE(a.E % 32 == 0 && a.E <= 192, "patch_embed: embed_dim must be a multiple of 32, <= 192");
  TULIP_REQUIRE(a.Wimg % 4 == 0 && a.Himg % a.ph == 0 && a.Wimg >= 4, "patch_embed: image not divisible by the patch");
  const int tokens = a.B * (a.Himg / a.ph) * (a.Wimg / 4);
  if (a.ph == 1 && (a.E == 96 || a.E == 192)) {
    const int tpw = a.E == 96 ? 4 : 2;
    const int want = ceil_div(tokens, 4 * tpw);
    if (a.E == 96) { const int grid = wave_grid(patch_embed_fwd12_kernel<8>, 128, 0, want); tulip_launch(patch_embed_fwd12_kernel<8>, grid, 128, 0, st, a); }
    else { const int grid = wave_grid(patch_embed_fwd12_kernel<16>, 128, 0, want); tulip_launch(patch_embed_fwd12_kernel<16>, grid, 128, 0, st, a); }
    TULIP_CHECK_LAUNCH();
    return TULIP_OK;
  }
  const int smem = a.E * a.ph * 8 * (int)sizeof(float);
  const int grid = min(ceil_div(tokens, 8 * 4), tulip_num_sms() * 8);
  switch (a.E / 32) {
    case 1: tulip_launch(patch_embed_fwd_kernel<1>, grid, 256, smem, st, a); break;
    case 2: tulip_launch(patch_embed_fwd_kernel<2>, grid, 256, smem, st, a); break;
    case 3: tulip_launch(patch_embed_fwd_kernel<3>, grid, 256, smem, st, a); break;
    case 4: tulip_launch(patch_embed_fwd_kernel<4>, grid, 256, smem, st, a); break;
    case 5: tulip_launch(patch_embed_fwd_kernel<5>, grid, 256, smem, st, a); break;
    default: tulip_launch(patch_embed_fwd_kernel<6>, grid, 256, smem, st, a); break;
  }
  TULIP_CHECK_LAUNCH();
  return TULIP_OK;
}

int patch_embed_bwd(const EmbedArgs& a, cudaStream_t st) {
  TULIP_REQUIRE(a.E % 32 == 0 && a.E <= 192, "patch_embed: embed_dim must be a multiple of 32, <= 192");
  TULIP_REQUIRE(a.ph == 1, "patch_embed backward: patch height must be 1");
  const int tokens = a.B * a.Himg * (a.Wimg / 4);
  if (a.E == 96) {
    const int smem6 = a.E * 13 * (int)sizeof(float);
    const int grid6 = wave_grid(patch_embed_bwd6_kernel, 128, smem6, ceil_div(tokens, 8));
    tulip_launch(patch_embed_bwd6_kernel, grid6, 128, smem6, st, a);
    TULIP_CHECK_LAUNCH();
    return TULIP_OK;
  }
  const int grid = min(ceil_div(tokens, 8 * 8), tulip_num_sms() * 6);
  const int smem = a.E * 20 * (int)sizeof(float);
  switch (a.E / 32) {
    case 1: tulip_launch(patch_embed_bwd_kernel<1>, grid, 256, smem, st, a); break;
    case 2: tulip_launch(patch_embed_bwd_kernel<2>, grid, 256, smem, st, a); break;
    case 3: tulip_launch(patch_embed_bwd_kernel<3>, grid, 256, smem, st, a); break;
    case 4: tulip_launch(patch_embed_bwd_kernel<4>, grid, 256, smem, st, a); break;
    case 5: tulip_launch(patch_embed_bwd_kernel<5>, grid, 256, smem, st, a); break;
    default: tulip_launch(patch_embed_bwd_kernel<6>, grid, 256, smem, st, a); break;
  }
  TULIP_CHECK_LAUNCH();
  return TULIP_OK;
}

int pack_weights(const float* flat, bf16* arena, const PackItem* items_dev, int n_items, int n_tiles, cudaStream_t st) {
  if (n_tiles <= 0) return TULIP_OK;
  tulip_launch(pack_weights_kernel, n_tiles, 256, 0, st, flat, arena, items_dev, n_items);
  TULIP_CHECK_LAUNCH();
  return TULIP_OK;
}

int permute_bias(const float* src, float* dst, int n, int R2, int Cc, cudaStream_t st) {
  tulip_launch(permute_bias_kernel, ceil_div(n, 256), 256, 0, st, src, dst, n, R2, Cc);
  TULIP_CHECK_LAUNCH();
  return TULIP_OK;
}

namespace {
// block (item, y): 8 warps each sum every 8th copy of 32 consecutive elements (independent loads, all in flight together),
// then warp 0 folds the 8 partial sums in a fixed order.  blockIdx.y strides over the element chunks of long items.
__global__ void __launch_bounds__(256) sum_copies_kernel(const __grid_constant__ SumCopiesArgs a) {
  pdl_sync();
  __shared__ float part[8][32];
  const SumCopiesItem it = a.item[blockIdx.x];
  const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
  for (int base = blockIdx.y * 32; base < it.n; base += gridDim.y * 32) {
    const int i = base + lane;
    float s = 0.f;
    if (i < it.n) {
#pragma unroll 4
      for (int c = grp; c < a.copies; c += 8) s += it.src[(long)c * it.stride + i];
    }
    part[grp][lane] = s;
    __syncthreads();
    if (grp == 0 && i < it.n) {
      float t = part[0][lane];
#pragma unroll
      for (int g2 = 1; g2 < 8; ++g2) t += part[g2][lane];
      it.dst[i] += t;
    }
    __syncthreads();
  }
}
}  // namespace

int sum_copies(const SumCopiesArgs& a, cudaStream_t st) {
  if (a.count <= 0) return TULIP_OK;
  int nmax = 0;
  for (int i = 0; i < a.count; ++i) nmax = a.item[i].n > nmax ? a.item[i].n : nmax;
  const int chunks = nmax > 256 ? 8 : (nmax + 31) / 32;
  tulip_launch(sum_copies_kernel, dim3(a.count, chunks > 0 ? chunks : 1), 256, 0, st, a);
  TULIP_CHECK_LAUNCH();
  return TULIP_OK;
}

int add_inplace_bf16(bf16* dst, const bf16* src, long n, cudaStream_t st) {
  TULIP_REQUIRE(n % 8 == 0, "add_inplace: length must be a multiple of 8");
  tulip_launch(add_inplace_kernel, ew_grid(n / 8, 256), 256, 0, st, reinterpret_cast<uint4*>(dst), reinterpret_cast<const uint4*>(src), n / 8);
  TULIP_CHECK_LAUNCH();
  return TULIP_OK;
}

int scale_rows_bf16(bf16* dst, const bf16* src, const float* row_scale, int rows, int C, int rows_per_sample, cudaStream_t st) {
  TULIP_REQUIRE(C % 8 == 0, "scale_rows: C must be a multiple of 8");
  const long n16 = (long)rows * (C / 8);
  tulip_launch(scale_rows_kernel, ew_grid(n16, 256), 256, 0, st, reinterpret_cast<uint4*>(dst), reinterpret_cast<const uint4*>(src), row_scale,
                                                       n16, C / 8, rows_per_sample);
  TULIP_CHECK_LAUNCH();
  return TULIP_OK;
}

// Inputs of a step into the module's persistent buffers (stable addresses let the step replay as a CUDA graph): low-res
// frames, targets and DropPath scales in ONE launch instead of three copy launches on the host's critical path.
namespace {
struct Stage3 { const float* src[3]; float* dst[3]; long n[3]; };
__global__ void stage_inputs_kernel(const Stage3 s) {
  pdl_sync();
  const long stride = (long)gridDim.x * blockDim.x;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const long n4 = s.n[k] >> 2;
    const float4* a = reinterpret_cast<const float4*>(s.src[k]);
    float4* b = reinterpret_cast<float4*>(s.dst[k]);
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) b[i] = a[i];
    for (long i = (n4 << 2) + (long)blockIdx.x * blockDim.x + threadIdx.x; i < s.n[k]; i += stride) s.dst[k][i] = s.src[k][i];
  }
}
}  // namespace

int stage_inputs(const float* const* src, float* const* dst, const long* n, cudaStream_t st) {
  Stage3 s;
  long total = 0;
  for (int k = 0; k < 3; ++k) {
    s.src[k] = src[k]; s.dst[k] = dst[k]; s.n[k] = (src[k] && dst[k]) ? n[k] : 0;
    TULIP_REQUIRE(s.n[k] == 0 || (((reinterpret_cast<uintptr_t>(src[k]) | reinterpret_cast<uintptr_t>(dst[k])) & 15) == 0),
                  "stage_inputs: 16-byte aligned buffers");
    total += s.n[k];
  }
  if (total == 0) return TULIP_OK;
  tulip_launch(stage_inputs_kernel, ew_grid((total + 3) / 4, 256), 256, 0, st, s);
  TULIP_CHECK_LAUNCH();
  return TULIP_OK;
}

int l1_loss(const float* pred, const float* target, long n, int log_transform, float* acc2, float* out2, cudaStream_t st) {
  TULIP_CUDA(cudaMemsetAsync(acc2, 0, 2 * sizeof(float), st));
  tulip_launch(l1_loss_kernel, ew_grid(n, 256), 256, 0, st, pred, target, n, log_transform, acc2);
  tulip_launch(l1_loss_finalize_kernel, 1, 1, 0, st, acc2, out2, 1.0f / (float)n, log_transform);
  TULIP_CHECK_LAUNCH();
  return TULIP_OK;
}

// ------------------------------------------------------------------------------------------------
// Evaluation post-processing of one batch of predictions (reference engine_upsampling.py:174-244, evaluate()):
//   expm1 of pred / target / input when the model was trained in log space (:177-180), range clip of the prediction to
//   [clip_lo, 1] else 0 (:183-188), pixel loss = mean |pred - target| per frame (:192-193), and for the rows the sensor
//   measured (every `factor`-th row, :214-221, :236-244) the loss against the input and the overwrite pred[row] = input.
// One CTA-strided pass: each byte read once, pred written once; per-frame sums via block reduction + 2 atomics per CTA.
__global__ void __launch_bounds__(256) eval_postprocess_kernel(const float* __restrict__ pred, const float* __restrict__ lo,
                                                               const float* __restrict__ hi, float* __restrict__ out,
                                                               float* __restrict__ sums, int H, int W, int factor, int log_transform,
                                                               float clip_lo, int keep_low_res, int chunks_per_frame) {
  pdl_sync();
  const int b = blockIdx.x / chunks_per_frame, chunk = blockIdx.x % chunks_per_frame;
  const long frame = (long)H * W;
  const int h_lo = H / factor;
  float s_pix = 0.f, s_low = 0.f;
  for (long i = (long)chunk * blockDim.x + threadIdx.x; i < frame / 4; i += (long)chunks_per_frame * blockDim.x) {
    const long px = 4 * i;
    const int row = (int)(px / W), col = (int)(px % W);
    const float4 p4 = *reinterpret_cast<const float4*>(pred + b * frame + px);
    const float4 t4 = *reinterpret_cast<const float4*>(hi + b * frame + px);
    const bool sensor_row = keep_low_res && (row % factor == 0);
    float4 l4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (sensor_row) l4 = *reinterpret_cast<const float4*>(lo + ((long)b * h_lo + row / factor) * W + col);
    float p[4] = {p4.x, p4.y, p4.z, p4.w}, t[4] = {t4.x, t4.y, t4.z, t4.w}, l[4] = {l4.x, l4.y, l4.z, l4.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float v = log_transform ? expm1f(p[k]) : p[k];
      const float tv = log_transform ? expm1f(t[k]) : t[k];
      v = (v >= clip_lo && v <= 1.0f) ? v : 0.f;
      s_pix += fabsf(v - tv);
      if (sensor_row) {
        const float lv = log_transform ? expm1f(l[k]) : l[k];
        s_low += fabsf(v - lv);
        v = lv;
      }
      p[k] = v;
    }
    *reinterpret_cast<float4*>(out + b * frame + px) = make_float4(p[0], p[1], p[2], p[3]);
  }
  __shared__ float red[2][8];
  s_pix = warp_sum(s_pix); s_low = warp_sum(s_low);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s_pix; red[1][threadIdx.x >> 5] = s_low; }
  __syncthreads();
  if (threadIdx.x < 2) {
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) a += red[threadIdx.x][w];
    atomicAdd(sums + 2 * b + threadIdx.x, a);
  }
}

__global__ void eval_postprocess_finalize_kernel(const float* sums, float* losses, int B, float inv_pix, float inv_low) {
  pdl_sync();
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    losses[2 * b] = sums[2 * b] * inv_pix;
    losses[2 * b + 1] = sums[2 * b + 1] * inv_low;
  }
}

// Monte-Carlo-dropout aggregation (engine_upsampling.py:423-427): mean and unbiased std over the N stochastic passes of a pixel,
// pixels whose std exceeds threshold * mean are zeroed.  One thread per pixel, two passes over its N values (N <= a few dozen).
__global__ void __launch_bounds__(256) mc_aggregate_kernel(const float* __restrict__ preds, float* __restrict__ out, float* __restrict__ std_out,
                                                           int n, long npix, float threshold) {
  pdl_sync();
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += (long)gridDim.x * blockDim.x) {
    float sum = 0.f;
    for (int k = 0; k < n; ++k) sum += preds[(long)k * npix + i];
    const float mean = sum / (float)n;
    float sq = 0.f;
    for (int k = 0; k < n; ++k) {
      const float d = preds[(long)k * npix + i] - mean;
      sq = fmaf(d, d, sq);
    }
    const float sd = sqrtf(sq / (float)(n - 1));              // torch.std: Bessel's correction
    if (std_out) std_out[i] = sd;
    out[i] = (sd > threshold * mean) ? 0.f : mean;
  }
}

int mc_aggregate(const float* preds, float* out, float* std_out, int n, long npix, float threshold, cudaStream_t st) {
  TULIP_REQUIRE(n >= 2 && npix > 0, "mc_aggregate: needs at least two passes");
  const int grid = (int)min((npix + 255) / 256, (long)tulip_num_sms() * 8);
  tulip_launch(mc_aggregate_kernel, grid, 256, 0, st, preds, out, std_out, n, npix, threshold);
  TULIP_CHECK_LAUNCH();
  return TULIP_OK;
}

int eval_postprocess(const float* pred, const float* lo, const float* hi, float* out, float* losses, float* scratch, int B, int H, int W,
                     int h_lo, int log_transform, float clip_lo, int keep_low_res, cudaStream_t st) {
  TULIP_REQUIRE(B > 0 && H > 0 && W > 0 && h_lo > 0 && H % h_lo == 0, "eval_postprocess: H must be a multiple of the input height");
  TULIP_REQUIRE(W % 4 == 0, "eval_postprocess: width must be a multiple of 4");
  const int factor = H / h_lo;
  const long quads = (long)H * W / 4;
  int chunks = (int)((quads + 255) / 256);
  const int want = max(1, 4 * tulip_num_sms() / B);
  if (chunks > want) chunks = want;
  TULIP_CUDA(cudaMemsetAsync(scratch, 0, 2 * B * sizeof(float), st));
  tulip_launch(eval_postprocess_kernel, B * chunks, 256, 0, st, pred, lo, hi, out, scratch, H, W, factor, log_transform, clip_lo,
               keep_low_res, chunks);
  tulip_launch(eval_postprocess_finalize_kernel, 1, 128, 0, st, scratch, losses, B, 1.0f / ((float)H * W),
               keep_low_res ? 1.0f / ((float)h_lo * W) : 0.f);
  TULIP_CHECK_LAUNCH();
  return TULIP_OK;
}

// ------------------------------------------------------------------------------------------------
// Input pipeline of one batch of raw range frames (reference tulip/util/datasets.py): npy_loader's channel pick (:187-191),
// ScaleTensor (:140-144), FilterInvalidPixels (:146-154, durlar / carla only), DownsampleTensor rows (:120-128) and
// DownsampleTensorWidth (:130-138) for the low-resolution input, LogTransform (:73-75) -- the transform chains of
// build_{kitti,durlar,carla}_upsampling_dataset (:244-369) -- as one pass that writes both model inputs.
__global__ void __launch_bounds__(256) preprocess_range_kernel(const float* __restrict__ raw, int channels, float scale, int filter,
                                                               float min_range, float max_range, int row_factor, int col_factor,
                                                               int log_transform, float* __restrict__ hi, float* __restrict__ lo, long n,
                                                               int H, int W) {
  pdl_sync();
  const int h_lo = H / row_factor, w_lo = W / col_factor;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const int w = (int)(i % W), h = (int)((i / W) % H);
    const long b = i / ((long)W * H);
    float v = __fmul_rn(raw[i * channels], scale);
    if (filter) v = (v >= min_range && v <= max_range) ? v : 0.f;
    if (log_transform) v = log1pf(v);
    hi[i] = v;
    if (h % row_factor == 0 && w % col_factor == 0) lo[(b * h_lo + h / row_factor) * w_lo + w / col_factor] = v;
  }
}

int preprocess_range(const float* raw, int channels, float scale, int filter, float min_range, float max_range, int row_factor,
                     int col_factor, int log_transform, float* hi, float* lo, int B, int H, int W, cudaStream_t st) {
  TULIP_REQUIRE(B > 0 && H > 0 && W > 0 && channels >= 1, "preprocess_range: empty input");
  TULIP_REQUIRE(row_factor >= 1 && col_factor >= 1 && H % row_factor == 0 && W % col_factor == 0,
                "preprocess_range: image size is not a multiple of the downsampling factors");
  const long n = (long)B * H * W;
  tulip_launch(preprocess_range_kernel, ew_grid(n, 256), 256, 0, st, raw, channels, scale, filter, min_range, max_range, row_factor,
               col_factor, log_transform, hi, lo, n, H, W);
  TULIP_CHECK_LAUNCH();
  return TULIP_OK;
}

// ------------------------------------------------------------------------------------------------
// CARLA `.rimg` frames (reference rimg_loader, tulip/util/datasets.py:181-193): the file stores size[1] rows of size[0] float16
// values; the loader returns flip(transpose(rows)) as float32, i.e. frame[i][j] = rows[size1 - 1 - j][size0 - 1 - i].
// 32 x 32 tiles through shared memory so both the fp16 reads and the fp32 writes are coalesced.
__global__ void __launch_bounds__(256) rimg_decode_kernel(const __half* __restrict__ rows, float* __restrict__ frame, int s0, int s1) {
  pdl_sync();
  __shared__ float tile[32][33];
  const long fb = (long)blockIdx.z * s0 * s1;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
#pragma unroll
  for (int k = 0; k < 32; k += 8) {
    const int r = r0 + threadIdx.y + k, c = c0 + threadIdx.x;
    if (r < s1 && c < s0) tile[threadIdx.y + k][threadIdx.x] = __half2float(rows[fb + (long)r * s0 + c]);
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 32; k += 8) {
    const int c = c0 + threadIdx.y + k, r = r0 + threadIdx.x;
    if (r < s1 && c < s0) frame[fb + (long)(s0 - 1 - c) * s1 + (s1 - 1 - r)] = tile[threadIdx.x][threadIdx.y + k];
  }
}

int rimg_decode(const void* rows_f16, float* frame, int B, int size0, int size1, cudaStream_t st) {
  TULIP_REQUIRE(B > 0 && size0 > 0 && size1 > 0, "rimg_decode: empty input");
  TULIP_REQUIRE(B <= 65535 && ceil_div(size1, 32) <= 65535, "rimg_decode: batch / frame width beyond the launch grid");
  tulip_launch(rimg_decode_kernel, dim3(ceil_div(size0, 32), ceil_div(size1, 32), B), dim3(32, 8), 0, st,
               reinterpret_cast<const __half*>(rows_f16), frame, size0, size1);
  TULIP_CHECK_LAUNCH();
  return TULIP_OK;
}

// ------------------------------------------------------------------------------------------------
// AdamW over the flat parameter / gradient buffers (reference: torch.optim.AdamW over param_groups_layer_decay groups,
// main_lidar_upsampling.py:281-283; the update is torch's _single_tensor_adamw, decoupled weight decay, no amsgrad) and the global
// gradient L2 norm (util/misc.py:317-329 get_grad_norm_).  The 212 parameters are segments of one buffer, so the step is ONE
// launch (p, g, m, v read once, p, m, v written once) instead of a multi-tensor apply over 212 tensors.
__device__ __forceinline__ int adamw_find_segment(const AdamwSegment* __restrict__ segs, int n, long i) {
  int lo = 0, hi = n - 1;                                  // last segment with offset <= i
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (segs[mid].offset <= i) lo = mid; else hi = mid - 1;
  }
  return lo;
}

__global__ void __launch_bounds__(256) adamw_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                         float* __restrict__ v, const AdamwSegment* __restrict__ segs, int n_segs,
                                                         long span4, const __grid_constant__ AdamwHyper hp) {
  pdl_sync();
  for (long q = (long)blockIdx.x * blockDim.x + threadIdx.x; q < span4; q += (long)gridDim.x * blockDim.x) {
    const long i = 4 * q;
    const AdamwSegment sg = segs[adamw_find_segment(segs, n_segs, i)];
    if (i < sg.offset || i >= sg.offset + sg.numel) continue;             // alignment padding between parameters
    const float lr = hp.lr[sg.group], wd = hp.weight_decay[sg.group];
    const float step_size = lr / hp.bias_correction1;
    float4 p4 = *reinterpret_cast<float4*>(p + i), m4 = *reinterpret_cast<float4*>(m + i), v4 = *reinterpret_cast<float4*>(v + i);
    const float4 g4 = *reinterpret_cast<const float4*>(g + i);
    float pp[4] = {p4.x, p4.y, p4.z, p4.w}, mm[4] = {m4.x, m4.y, m4.z, m4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w};
    const float gg[4] = {g4.x * hp.grad_scale, g4.y * hp.grad_scale, g4.z * hp.grad_scale, g4.w * hp.grad_scale};
    const int valid = (int)min(4l, sg.offset + sg.numel - i);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (k >= valid) break;
      pp[k] = pp[k] * (1.0f - lr * wd);                                    // param.mul_(1 - lr * weight_decay)
      mm[k] = mm[k] + (gg[k] - mm[k]) * (1.0f - hp.beta1);                 // exp_avg.lerp_(grad, 1 - beta1)
      vv[k] = vv[k] * hp.beta2 + (1.0f - hp.beta2) * gg[k] * gg[k];        // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
      const float denom = sqrtf(vv[k]) / hp.bias_correction2_sqrt + hp.eps;
      pp[k] = pp[k] - step_size * (mm[k] / denom);                         // param.addcdiv_(exp_avg, denom, value=-step_size)
    }
    *reinterpret_cast<float4*>(p + i) = make_float4(pp[0], pp[1], pp[2], pp[3]);
    *reinterpret_cast<float4*>(m + i) = make_float4(mm[0], mm[1], mm[2], mm[3]);
    *reinterpret_cast<float4*>(v + i) = make_float4(vv[0], vv[1], vv[2], vv[3]);
  }
}

__global__ void __launch_bounds__(256) grad_sumsq_kernel(const float* __restrict__ g, const AdamwSegment* __restrict__ segs, int n_segs,
                                                         long span4, double* __restrict__ acc) {
  pdl_sync();
  double s = 0.0;
  for (long q = (long)blockIdx.x * blockDim.x + threadIdx.x; q < span4; q += (long)gridDim.x * blockDim.x) {
    const long i = 4 * q;
    const AdamwSegment sg = segs[adamw_find_segment(segs, n_segs, i)];
    if (i < sg.offset || i >= sg.offset + sg.numel) continue;
    const float4 g4 = *reinterpret_cast<const float4*>(g + i);
    const float gg[4] = {g4.x, g4.y, g4.z, g4.w};
    const int valid = (int)min(4l, sg.offset + sg.numel - i);
    for (int k = 0; k < valid; ++k) s += (double)gg[k] * gg[k];
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  __shared__ double red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[w];
    atomicAdd(acc, t);
  }
}
__global__ void grad_norm_finalize_kernel(const double* acc, float* out) { pdl_sync(); out[0] = (float)sqrt(acc[0]); }

int adamw_step(float* p, const float* g, float* m, float* v, const AdamwSegment* segs_dev, int n_segs, long span, const AdamwHyper& hp,
               cudaStream_t st) {
  TULIP_REQUIRE(n_segs > 0 && span > 0 && span % 4 == 0, "adamw_step: empty or unaligned parameter span");
  tulip_launch(adamw_step_kernel, ew_grid(span / 4, 256), 256, 0, st, p, g, m, v, segs_dev, n_segs, span / 4, hp);
  TULIP_CHECK_LAUNCH();
  return TULIP_OK;
}

int grad_norm(const float* g, const AdamwSegment* segs_dev, int n_segs, long span, double* scratch, float* out, cudaStream_t st) {
  TULIP_REQUIRE(n_segs > 0 && span > 0 && span % 4 == 0, "grad_norm: empty or unaligned parameter span");
  TULIP_CUDA(cudaMemsetAsync(scratch, 0, sizeof(double), st));
  tulip_launch(grad_sumsq_kernel, ew_grid(span / 4, 256), 256, 0, st, g, segs_dev, n_segs, span / 4, scratch);
  tulip_launch(grad_norm_finalize_kernel, 1, 1, 0, st, (const double*)scratch, out);
  TULIP_CHECK_LAUNCH();
  return TULIP_OK;
}

int window_gather(const bf16* x, bf16* out, int B, int H, int W, int C, int Mh, int Mw, int sh, int sw, cudaStream_t st) {
  TULIP_REQUIRE(C % 8 == 0 && H % Mh == 0 && W % Mw == 0, "H or W is not divisible by window_size");
  const long n = (long)B * H * W * (C / 8);
  tulip_launch(window_copy_kernel, ew_grid(n, 256), 256, 0, st, reinterpret_cast<const uint4*>(x), reinterpret_cast<uint4*>(out), B,
                                                      WinGeom{H, W, Mh, Mw, sh, sw}, C / 8, false);
  TULIP_CHECK_LAUNCH();
  return TULIP_OK;
}

int window_scatter(const bf16* xw, bf16* out, int B, int H, int W, int C, int Mh, int Mw, int sh, int sw, cudaStream_t st) {
  TULIP_REQUIRE(C % 8 == 0 && H % Mh == 0 && W % Mw == 0, "H or W is not divisible by window_size");
  const long n = (long)B * H * W * (C / 8);
  tulip_launch(window_copy_kernel, ew_grid(n, 256), 256, 0, st, reinterpret_cast<const uint4*>(xw), reinterpret_cast<uint4*>(out), B,
                                                      WinGeom{H, W, Mh, Mw, sh, sw}, C / 8, true);
  TULIP_CHECK_LAUNCH();
  return TULIP_OK;
}

int shift_mask(float* out, int H, int W, int Mh, int Mw, int sh, int sw, cudaStream_t st) {
  TULIP_REQUIRE(H % Mh == 0 && W % Mw == 0, "H or W is not divisible by window_size");
  const int L = Mh * Mw;
  const long n = (long)(H / Mh) * (W / Mw) * L * L;
  tulip_launch(shift_mask_kernel, ew_grid(n, 256), 256, 0, st, out, WinGeom{H, W, Mh, Mw, sh, sw});
  TULIP_CHECK_LAUNCH();
  return TULIP_OK;
}

int rel_bias_gather(const float* table, float* out, int heads, int Mh, int Mw, cudaStream_t st) {
  const int L = Mh * Mw;
  tulip_launch(rel_bias_gather_kernel, ew_grid((long)heads * L * L, 256), 256, 0, st, table, out, heads, Mh, Mw);
  TULIP_CHECK_LAUNCH();
  return TULIP_OK;
}

int merge_gather(const bf16* x, bf16* out, int B, int H, int W, int C, cudaStream_t st) {
  TULIP_REQUIRE(C % 8 == 0 && H % 2 == 0 && W % 2 == 0, "merge: H, W must be even and C a multiple of 8");
  LnArgs a = {};
  a.C = 4 * C; a.gather = 1; a.H2 = H / 2; a.W2 = W / 2;
  const long total = (long)B * (H / 2) * (W / 2) * (4 * C / 8);
  tulip_launch(merge_gather_kernel, ew_grid(total, 256), 256, 0, st, reinterpret_cast<const uint4*>(x), reinterpret_cast<uint4*>(out), a, total);
  TULIP_CHECK_LAUNCH();
  return TULIP_OK;
}

int pixel_shuffle_nhwc(const bf16* x, bf16* out, int B, int H, int W, int Cout, int r, cudaStream_t st) {
  const long total = (long)B * H * W * Cout * r * r;
  tulip_launch(pixel_shuffle_kernel, ew_grid(total, 256), 256, 0, st, x, out, B, H, W, Cout, r);
  TULIP_CHECK_LAUNCH();
  return TULIP_OK;
}
