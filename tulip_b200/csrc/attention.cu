// Window attention core (W-MSA / SW-MSA) forward and backward.
//
// Reference: tulip/model/tulip.py:282-324 (WindowAttention.forward) minus the two Linear layers,
// which are per-token and therefore run as plain GEMMs on the un-partitioned, un-shifted tensor.
// The cyclic shift (torch.roll, :290/:323), window partition / reverse (:295/:320), the shift mask
// (create_mask, :254-280, closed form of SURVEY.md App. D) and the relative-position-bias gather
// (:304-308) are index arithmetic inside these kernels -- no data-movement kernels, no mask tensor.
//
// A window is 16 tokens and head_dim is 32 at every stage, so one (window, head) problem is
// S = Q K^T (16x16x32) and O = P V (16x32x16): far below a tcgen05 tile (M >= 64).  One warp owns
// one (window, head) and keeps S / P in registers with m16n8k16 warp MMAs; a CTA of 3 warps stages
// the q|k|v rows of one window and three heads in shared memory with 16-byte cp.async.
#include "common.cuh"
#include "kernels.h"
#include "window_index.cuh"

namespace {

constexpr int L = 16;          // tokens per window
constexpr int HD = 32;         // head dim
constexpr int HG = 3;          // heads per CTA
constexpr int QLD = 3 * HG * HD + 8;   // 296-element rows (592 B): conflict-free ldmatrix
constexpr int OLD = HG * HD + 8;       // 104-element rows for dO
constexpr int PLD = 24;                // 16x16 scratch rows (48 B)
constexpr int BWD_NB = 3;              // backward: staging buffers per CTA (windows in flight + 1)

__device__ __forceinline__ WinGeom geom(const AttnArgs& a) { return WinGeom{a.H, a.W, a.Mh, a.Mw, a.sh, a.sw}; }
__device__ __forceinline__ int token_index(const AttnArgs& a, int b, int wh, int ww, int i) {
  return win_token_index(geom(a), b, wh, ww, i);
}
__device__ __forceinline__ int region_id(const AttnArgs& a, int wh, int ww, int i) { return win_region_id(geom(a), wh, ww, i); }
__device__ __forceinline__ int bias_index(const AttnArgs& a, int i, int j) { return rel_bias_index(a.bMh, a.bMw, i, j); }

// this thread's 8 relative-position-bias values (fixed (i,j) fragment positions, fixed head for the whole kernel)
__device__ __forceinline__ void load_bias(const AttnArgs& a, int head, int lane, float (&bias)[2][4]) {
  const int gq = lane >> 2, tq = lane & 3;
#pragma unroll
  for (int nt = 0; nt < 2; ++nt)
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int i = gq + (q >> 1) * 8, j = nt * 8 + 2 * tq + (q & 1);
      bias[nt][q] = a.bias_table[bias_index(a, i, j) * a.heads + head];
    }
}

// S (two m16n8 accumulators = 16x16) for this warp's head: scale * Q K^T + bias (+ mask)
__device__ __forceinline__ void scores(const AttnArgs& a, const bf16* sq, int hl, uint32_t maskbits,
                                       const float (&bias)[2][4], float (&s)[2][4], int lane) {
#pragma unroll
  for (int nt = 0; nt < 2; ++nt)
#pragma unroll
    for (int q = 0; q < 4; ++q) s[nt][q] = 0.f;
  const int mat = lane >> 3;
#pragma unroll
  for (int ks = 0; ks < 2; ++ks) {
    uint32_t af[4], bfr[4];
    ldmatrix_x4(af, sq + (lane & 15) * QLD + hl * HD + ks * 16 + (lane >> 4) * 8);
    ldmatrix_x4(bfr, sq + ((mat >> 1) * 8 + (lane & 7)) * QLD + HG * HD + hl * HD + ks * 16 + (mat & 1) * 8);
    mma_bf16_16816(s[0], af, bfr[0], bfr[1]);
    mma_bf16_16816(s[1], af, bfr[2], bfr[3]);
  }
#pragma unroll
  for (int nt = 0; nt < 2; ++nt)
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float v = s[nt][q] * a.scale + bias[nt][q];
      if ((maskbits >> (nt * 4 + q)) & 1u) v += -100.0f;
      s[nt][q] = v;
    }
}

// Per-thread window geometry, computed once per kernel: every runtime division of the partition / shift / mask index
// arithmetic lives here, so the per-window work is a handful of adds and compares (the kernels were issue-bound on it).
struct WinThread {
  int H, W, Mh, Mw, sh, sw, nWh, nWw;
  int lr[2], lc[2];        // window (row, col) of the two tokens this thread stages (li, li + 8)
  int orow[2], ocol[2];    // ... and of the two fragment rows it writes (gq, gq + 8)
  uint32_t diffH, diffW;   // bit nt*4+q: tokens i, j of that score fragment lie on different sides of the shift seam
};
__device__ __forceinline__ WinThread win_thread(const AttnArgs& a, int li, int lane) {
  WinThread t;
  t.H = a.H; t.W = a.W; t.Mh = a.Mh; t.Mw = a.Mw; t.sh = a.sh; t.sw = a.sw; t.nWh = a.H / a.Mh; t.nWw = a.W / a.Mw;
  const int gq = lane >> 2, tq = lane & 3;
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    t.lr[k] = (li + 8 * k) / a.Mw; t.lc[k] = (li + 8 * k) % a.Mw;
    t.orow[k] = (gq + 8 * k) / a.Mw; t.ocol[k] = (gq + 8 * k) % a.Mw;
  }
  t.diffH = 0; t.diffW = 0;
#pragma unroll
  for (int nt = 0; nt < 2; ++nt)
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int i = gq + (q >> 1) * 8, j = nt * 8 + 2 * tq + (q & 1);
      // create_mask (tulip.py:261-271): inside the last window row the rows >= Mh - sh form their own region, same for columns
      if (((i / a.Mw) >= a.Mh - a.sh) != ((j / a.Mw) >= a.Mh - a.sh)) t.diffH |= 1u << (nt * 4 + q);
      if (((i % a.Mw) >= a.Mw - a.sw) != ((j % a.Mw) >= a.Mw - a.sw)) t.diffW |= 1u << (nt * 4 + q);
    }
  return t;
}
// flat NHWC token index of window-local (r, c) in window (b, wh, ww): win_token_index without the divisions
__device__ __forceinline__ int win_tok(const WinThread& t, int b, int wh, int ww, int r, int c) {
  int hs = wh * t.Mh + r + t.sh;
  int ws = ww * t.Mw + c + t.sw;
  if (hs >= t.H) hs -= t.H;
  if (ws >= t.W) ws -= t.W;
  return (b * t.H + hs) * t.W + ws;
}
// shift mask of one window as fragment bits (win_region_id differs between i and j)
__device__ __forceinline__ uint32_t win_maskbits(const AttnArgs& a, const WinThread& t, int wh, int ww) {
  if (!a.masked) return 0u;
  uint32_t m = 0;
  if (t.sh > 0 && wh == t.nWh - 1) m |= t.diffH;
  if (t.sw > 0 && ww == t.nWw - 1) m |= t.diffW;
  return m;
}

__device__ __forceinline__ void softmax_rows(float (&s)[2][4]) {
#pragma unroll
  for (int hf = 0; hf < 2; ++hf) {
    float m = fmaxf(fmaxf(s[0][hf * 2], s[0][hf * 2 + 1]), fmaxf(s[1][hf * 2], s[1][hf * 2 + 1]));
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
    float e[4], sum = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      e[q] = __expf(s[q >> 1][hf * 2 + (q & 1)] - m);
      sum += e[q];
    }
    sum += __shfl_xor_sync(0xffffffffu, sum, 1);
    sum += __shfl_xor_sync(0xffffffffu, sum, 2);
    const float inv = 1.0f / sum;
#pragma unroll
    for (int q = 0; q < 4; ++q) s[q >> 1][hf * 2 + (q & 1)] = e[q] * inv;
  }
}

// stage the q|k|v rows of one window for this CTA's head group: thread -> tokens (li, li+8), 16-byte chunk lc8 of each segment
__device__ __forceinline__ void load_window_rows(const AttnArgs& a, bf16* sq, const bf16* __restrict__ qkv, const int (&tok)[2],
                                                 int li, int lc8, int hg) {
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const bf16* src = qkv + (long)tok[k] * 3 * a.C + hg * HG * HD + lc8;
    bf16* dst = sq + (li + 8 * k) * QLD + lc8;
#pragma unroll
    for (int seg = 0; seg < 3; ++seg) {
      if (a.hint) cp_async16_pol(dst + seg * HG * HD, src + seg * a.C, l2_evict_first_policy());
      else cp_async16(dst + seg * HG * HD, src + seg * a.C, 16);
    }
  }
}

__global__ void __launch_bounds__(96) win_attn_fwd_kernel(const AttnArgs a) {
  pdl_sync();
  __shared__ __align__(16) bf16 sq2[2][L * QLD];       // double buffer: the next window's rows land while this one is computed
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int li = tid / 12, lc8 = (tid % 12) * 8;
  const WinThread wt = win_thread(a, li, lane);
  const int hgn = a.heads / HG;
  const int nwin = a.B * wt.nWh * wt.nWw;
  // a CTA keeps one head group for its whole life (grid is a multiple of hgn), so bias values live in registers
  const int hg = blockIdx.x % hgn;
  const int head = hg * HG + warp;
  const int wstep = gridDim.x / hgn;
  float bias[2][4];
  load_bias(a, head, lane, bias);

  auto decode = [&](int widx, int& b, int& wh, int& ww) {
    ww = widx % wt.nWw;
    const int t = widx / wt.nWw;
    wh = t % wt.nWh;
    b = t / wt.nWh;
  };
  auto prefetch = [&](int b, int wh, int ww, int buf) {
    const int tok[2] = {win_tok(wt, b, wh, ww, wt.lr[0], wt.lc[0]), win_tok(wt, b, wh, ww, wt.lr[1], wt.lc[1])};
    load_window_rows(a, sq2[buf], a.qkv, tok, li, lc8, hg);
  };

  int widx = blockIdx.x / hgn;
  int buf = 0;
  int b = 0, wh = 0, ww = 0;
  if (widx < nwin) { decode(widx, b, wh, ww); prefetch(b, wh, ww, 0); }
  cp_async_commit();
  for (; widx < nwin; widx += wstep, buf ^= 1) {
    const int next = widx + wstep;
    int nb = 0, nwh = 0, nww = 0;
    if (next < nwin) { decode(next, nb, nwh, nww); prefetch(nb, nwh, nww, buf ^ 1); }
    cp_async_commit();
    cp_async_wait<1>();                                 // everything but the newest group (the prefetch) has landed
    __syncthreads();
    const bf16* sq = sq2[buf];

    float s[2][4];
    scores(a, sq, warp, win_maskbits(a, wt, wh, ww), bias, s, lane);
    softmax_rows(s);
    // P enters the tensor core as a two-term bf16 split (hi + lo, 16 significand bits): its rounding would otherwise be the
    // largest error of the kernel (2^-9 per element); the kernel is HBM-bound, the four extra MMAs are free
    uint32_t pf[4], pl[4];
    split_bf16(s[0][0], s[0][1], pf[0], pl[0]); split_bf16(s[0][2], s[0][3], pf[1], pl[1]);
    split_bf16(s[1][0], s[1][1], pf[2], pl[2]); split_bf16(s[1][2], s[1][3], pf[3], pl[3]);
    float o[4][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int q = 0; q < 4; ++q) o[nt][q] = 0.f;
    const int mat = lane >> 3;
#pragma unroll
    for (int np = 0; np < 2; ++np) {
      uint32_t bfr[4];
      ldmatrix_x4_trans(bfr, sq + ((mat & 1) * 8 + (lane & 7)) * QLD + 2 * HG * HD + warp * HD + np * 16 + (mat >> 1) * 8);
      mma_bf16_16816(o[2 * np], pf, bfr[0], bfr[1]);
      mma_bf16_16816(o[2 * np + 1], pf, bfr[2], bfr[3]);
      mma_bf16_16816(o[2 * np], pl, bfr[0], bfr[1]);
      mma_bf16_16816(o[2 * np + 1], pl, bfr[2], bfr[3]);
    }
    const int tq = lane & 3;
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      const int t = win_tok(wt, b, wh, ww, wt.orow[hf], wt.ocol[hf]);
      bf16* dst = a.out + (long)t * a.C + head * HD + 2 * tq;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) *reinterpret_cast<uint32_t*>(dst + nt * 8) = pack_bf16(o[nt][hf * 2], o[nt][hf * 2 + 1]);
    }
    b = nb; wh = nwh; ww = nww;
    __syncthreads();                                    // this buffer is overwritten by the prefetch of the next iteration
  }
}

// Backward of the attention core for one (window, head) per warp.  Recomputes S and P from the saved qkv.
//   dV = P^T dO, dP = dO V^T, dS = P o (dP - rowsum(dP o P)), dQ = scale * dS K, dK = scale * dS^T Q,
//   d bias_table[idx(i,j), head] += dS[i,j]     
__global__ void __launch_bounds__(96) win_attn_bwd_kernel(const AttnArgs a) {
  pdl_sync();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  bf16* sbuf = reinterpret_cast<bf16*>(smem_raw);             // NB x { [16][QLD] q|k|v, [16][OLD] dO }: two windows in flight
  constexpr int BUF = L * QLD + L * OLD;
  bf16* sp = sbuf + BWD_NB * BUF;                             // [3 warps][4][16][PLD]  P and dS scratch (hi and lo parts)
  float* s_dtab = reinterpret_cast<float*>(sp + 3 * 4 * L * PLD);   // [HG][nbias]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int li = tid / 12, lc8 = (tid % 12) * 8;
  const WinThread wt = win_thread(a, li, lane);
  const int hgn = a.heads / HG;
  const int nwin = a.B * wt.nWh * wt.nWw;
  const int hg = blockIdx.x % hgn;                              // fixed head group per CTA (grid is a multiple of hgn)
  const int head = hg * HG + warp;
  float bias[2][4], dsacc[2][4];                                // bias values and dS sums at this thread's fixed (i,j) positions
  load_bias(a, head, lane, bias);
#pragma unroll
  for (int nt = 0; nt < 2; ++nt)
#pragma unroll
    for (int q = 0; q < 4; ++q) dsacc[nt][q] = 0.f;
  for (int i = tid; i < HG * a.nbias; i += 96) s_dtab[i] = 0.f;

  const int wstep = gridDim.x / hgn;
  auto decode = [&](int widx, int& b, int& wh, int& ww) {
    ww = widx % wt.nWw;
    const int t = widx / wt.nWw;
    wh = t % wt.nWh;
    b = t / wt.nWh;
  };
  auto prefetch = [&](int b, int wh, int ww, int buf) {
    bf16* q = sbuf + buf * BUF;
    bf16* d = q + L * QLD;
    const int tok[2] = {win_tok(wt, b, wh, ww, wt.lr[0], wt.lc[0]), win_tok(wt, b, wh, ww, wt.lr[1], wt.lc[1])};
    load_window_rows(a, q, a.qkv, tok, li, lc8, hg);
#pragma unroll
    for (int k = 0; k < 2; ++k)
      if (a.hint) cp_async16_pol(d + (li + 8 * k) * OLD + lc8, a.dout + (long)tok[k] * a.C + hg * HG * HD + lc8, l2_evict_first_policy());
      else cp_async16(d + (li + 8 * k) * OLD + lc8, a.dout + (long)tok[k] * a.C + hg * HG * HD + lc8, 16);
  };
  // The kernel is bound by loads in flight (5 CTAs of ~120 registers per SM): each CTA keeps the rows of the next TWO
  // windows on their way (cp.async groups) while it works on the current one.
  int widx0 = blockIdx.x / hgn;
  int buf = 0;
  int b = 0, wh = 0, ww = 0;
#pragma unroll
  for (int k = 0; k < BWD_NB - 1; ++k) {
    const int w = widx0 + k * wstep;
    if (w < nwin) { decode(w, b, wh, ww); prefetch(b, wh, ww, k); }
    cp_async_commit();
  }
  for (int widx = widx0; widx < nwin; widx += wstep, buf = (buf + 1 == BWD_NB ? 0 : buf + 1)) {
    const int ahead = widx + (BWD_NB - 1) * wstep;
    if (ahead < nwin) {
      int nb, nwh, nww;
      decode(ahead, nb, nwh, nww);
      prefetch(nb, nwh, nww, buf + BWD_NB - 1 >= BWD_NB ? buf - 1 : buf + BWD_NB - 1);
    }
    cp_async_commit();
    cp_async_wait<BWD_NB - 1>();                         // everything but the newest groups (the windows ahead) has landed
    __syncthreads();
    decode(widx, b, wh, ww);
    const bf16* sq = sbuf + buf * BUF;
    const bf16* sdo = sq + L * QLD;

    const int mat = lane >> 3, gq = lane >> 2, tq = lane & 3;
    float p[2][4];
    scores(a, sq, warp, win_maskbits(a, wt, wh, ww), bias, p, lane);
    softmax_rows(p);

    // dP = dO V^T
    float dp[2][4];
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int q = 0; q < 4; ++q) dp[nt][q] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      uint32_t af[4], bfr[4];
      ldmatrix_x4(af, sdo + (lane & 15) * OLD + warp * HD + ks * 16 + (lane >> 4) * 8);
      ldmatrix_x4(bfr, sq + ((mat >> 1) * 8 + (lane & 7)) * QLD + 2 * HG * HD + warp * HD + ks * 16 + (mat & 1) * 8);
      mma_bf16_16816(dp[0], af, bfr[0], bfr[1]);
      mma_bf16_16816(dp[1], af, bfr[2], bfr[3]);
    }
    // dS = P o (dP - rowsum(dP o P))
    float ds[2][4];
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      float r = 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q) r += dp[q >> 1][hf * 2 + (q & 1)] * p[q >> 1][hf * 2 + (q & 1)];
      r += __shfl_xor_sync(0xffffffffu, r, 1);
      r += __shfl_xor_sync(0xffffffffu, r, 2);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int nt = q >> 1, e = hf * 2 + (q & 1);
        ds[nt][e] = p[nt][e] * (dp[nt][e] - r);
      }
    }
    // relative-position-bias gradient: summed in registers across this CTA's windows, scattered to bins once at the end
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int q = 0; q < 4; ++q) dsacc[nt][q] += ds[nt][q];
    // P and dS as two-term bf16 splits (hi + lo): stash both parts ([i][j]) for the transposed operands
    bf16* wp = sp + warp * 4 * L * PLD;
    bf16* wds = wp + L * PLD;
    bf16* wpl = wds + L * PLD;
    bf16* wdsl = wpl + L * PLD;
    uint32_t dsf[4], dsl[4];
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const int i = gq + hf * 8, j = nt * 8 + 2 * tq;
        uint32_t ph, pl_;
        split_bf16(p[nt][hf * 2], p[nt][hf * 2 + 1], ph, pl_);
        split_bf16(ds[nt][hf * 2], ds[nt][hf * 2 + 1], dsf[nt * 2 + hf], dsl[nt * 2 + hf]);
        *reinterpret_cast<uint32_t*>(wp + i * PLD + j) = ph;
        *reinterpret_cast<uint32_t*>(wpl + i * PLD + j) = pl_;
        *reinterpret_cast<uint32_t*>(wds + i * PLD + j) = dsf[nt * 2 + hf];
        *reinterpret_cast<uint32_t*>(wdsl + i * PLD + j) = dsl[nt * 2 + hf];
      }
    __syncwarp();

    float dq[4][4], dk[4][4], dv[4][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int q = 0; q < 4; ++q) dq[nt][q] = dk[nt][q] = dv[nt][q] = 0.f;
    uint32_t ptf[4], dstf[4], ptl[4], dstl[4];                  // P^T and dS^T (hi, lo) as A operands (rows j, contraction i)
    ldmatrix_x4_trans(ptf, wp + ((mat >> 1) * 8 + (lane & 7)) * PLD + (mat & 1) * 8);
    ldmatrix_x4_trans(dstf, wds + ((mat >> 1) * 8 + (lane & 7)) * PLD + (mat & 1) * 8);
    ldmatrix_x4_trans(ptl, wpl + ((mat >> 1) * 8 + (lane & 7)) * PLD + (mat & 1) * 8);
    ldmatrix_x4_trans(dstl, wdsl + ((mat >> 1) * 8 + (lane & 7)) * PLD + (mat & 1) * 8);
#pragma unroll
    for (int np = 0; np < 2; ++np) {
      uint32_t bdo[4], bk[4], bq[4];
      const int roff = ((mat & 1) * 8 + (lane & 7));
      const int coff = warp * HD + np * 16 + (mat >> 1) * 8;
      ldmatrix_x4_trans(bdo, sdo + roff * OLD + coff);                         // B[k=i][n=d] = dO[i][d]
      ldmatrix_x4_trans(bk, sq + roff * QLD + HG * HD + coff);                 // B[k=j][n=d] = K[j][d]
      ldmatrix_x4_trans(bq, sq + roff * QLD + coff);                           // B[k=i][n=d] = Q[i][d]
      mma_bf16_16816(dv[2 * np], ptf, bdo[0], bdo[1]);
      mma_bf16_16816(dv[2 * np + 1], ptf, bdo[2], bdo[3]);
      mma_bf16_16816(dq[2 * np], dsf, bk[0], bk[1]);
      mma_bf16_16816(dq[2 * np + 1], dsf, bk[2], bk[3]);
      mma_bf16_16816(dk[2 * np], dstf, bq[0], bq[1]);
      mma_bf16_16816(dk[2 * np + 1], dstf, bq[2], bq[3]);
      mma_bf16_16816(dv[2 * np], ptl, bdo[0], bdo[1]);
      mma_bf16_16816(dv[2 * np + 1], ptl, bdo[2], bdo[3]);
      mma_bf16_16816(dq[2 * np], dsl, bk[0], bk[1]);
      mma_bf16_16816(dq[2 * np + 1], dsl, bk[2], bk[3]);
      mma_bf16_16816(dk[2 * np], dstl, bq[0], bq[1]);
      mma_bf16_16816(dk[2 * np + 1], dstl, bq[2], bq[3]);
    }
    // dq|dk|dv replace this head's q|k|v columns in the staging buffer (only this warp reads or writes them), then the CTA
    // writes the window back with the same coalesced 16-byte pattern it was loaded with: 6 stores per thread instead of 24
    // 4-byte ones (the fragment layout gives each store instruction only 16 contiguous bytes per row)
    __syncwarp();
    {
      bf16* sw = const_cast<bf16*>(sq);
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        bf16* dst = sw + (gq + hf * 8) * QLD + warp * HD + 2 * tq;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          *reinterpret_cast<uint32_t*>(dst + nt * 8) = pack_bf16(dq[nt][hf * 2] * a.scale, dq[nt][hf * 2 + 1] * a.scale);
          *reinterpret_cast<uint32_t*>(dst + HG * HD + nt * 8) = pack_bf16(dk[nt][hf * 2] * a.scale, dk[nt][hf * 2 + 1] * a.scale);
          *reinterpret_cast<uint32_t*>(dst + 2 * HG * HD + nt * 8) = pack_bf16(dv[nt][hf * 2], dv[nt][hf * 2 + 1]);
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int t = win_tok(wt, b, wh, ww, wt.lr[k], wt.lc[k]);
      const bf16* src = sq + (li + 8 * k) * QLD + lc8;
      bf16* dst = a.dqkv + (long)t * 3 * a.C + hg * HG * HD + lc8;
#pragma unroll
      for (int seg = 0; seg < 3; ++seg)
        *reinterpret_cast<uint4*>(dst + seg * a.C) = *reinterpret_cast<const uint4*>(src + seg * HG * HD);
    }
    __syncthreads();                                    // buffers and the P/dS scratch are reused by the next iterations
  }
  __syncthreads();
  {
    const int gq = lane >> 2, tq = lane & 3;
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int i = gq + (q >> 1) * 8, j = nt * 8 + 2 * tq + (q & 1);
        atomicAdd(&s_dtab[warp * a.nbias + bias_index(a, i, j)], dsacc[nt][q]);
      }
  }
  __syncthreads();
  if (a.dbias_table == nullptr) return;
  // every CTA adds into the same few hundred addresses at about the same time: same-address atomics serialise in L2 (7-11 us
  // per launch, measured), so the CTAs spread over `dbias_copies` copies of the table that the caller sums afterwards
  float* dtab = a.dbias_table + (long)((blockIdx.x / hgn) % a.dbias_copies) * a.nbias * a.heads;
  for (int i = tid; i < HG * a.nbias; i += 96) {
    const float v = s_dtab[i];
    if (v != 0.f) atomicAdd(dtab + (i % a.nbias) * a.heads + hg * HG + i / a.nbias, v);
  }
}

int check_attn(const AttnArgs& a) {
  TULIP_REQUIRE(a.Mh * a.Mw == L, "window attention: windows must hold 16 tokens");
  TULIP_REQUIRE(a.bMh * a.bMw == L, "window attention: bias window must hold 16 tokens");
  TULIP_REQUIRE(a.C == a.heads * HD, "window attention: head_dim must be 32");
  TULIP_REQUIRE(a.heads % HG == 0, "window attention: heads must be a multiple of 3");
  TULIP_REQUIRE(a.H % a.Mh == 0 && a.W % a.Mw == 0, "H or W is not divisible by window_size");
  TULIP_REQUIRE(a.nbias == (2 * a.bMh - 1) * (2 * a.bMw - 1), "window attention: bias table size");
  return TULIP_OK;
}

}  // namespace

int win_attn_fwd(const AttnArgs& a, cudaStream_t st) {
  int rc = check_attn(a);
  if (rc) return rc;
  const int hgn = a.heads / HG, nwin = a.B * (a.H / a.Mh) * (a.W / a.Mw);
  static int per_sm = 0;
  if (!per_sm && (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, win_attn_fwd_kernel, 96, 0) != cudaSuccess || per_sm < 1)) per_sm = 8;
  const int slots = max(1, min(nwin, per_sm * tulip_num_sms() / hgn));      // one full wave; grid stays a multiple of hgn
  AttnArgs ah = a;
  ah.hint = (tulip_hints() & 32) ? 1 : 0;
  tulip_launch(win_attn_fwd_kernel, slots * hgn, 96, 0, st, ah);
  TULIP_CHECK_LAUNCH();
  return TULIP_OK;
}

int win_attn_bwd(const AttnArgs& a_in, cudaStream_t st) {
  AttnArgs a = a_in;
  if (a.dbias_copies < 1) a.dbias_copies = 1;
  int rc = check_attn(a);
  if (rc) return rc;
  const int hgn = a.heads / HG, nwin = a.B * (a.H / a.Mh) * (a.W / a.Mw);
  const int smem = (BWD_NB * (L * QLD + L * OLD) + 3 * 4 * L * PLD) * 2 + HG * a.nbias * 4;
  static int per_sm = 0;
  if (!per_sm && (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, win_attn_bwd_kernel, 96, smem) != cudaSuccess || per_sm < 1)) per_sm = 4;
  const int slots = max(1, min(nwin, per_sm * tulip_num_sms() / hgn));
  a.hint = (tulip_hints() & 32) ? 1 : 0;
  tulip_launch(win_attn_bwd_kernel, slots * hgn, 96, smem, st, a);
  TULIP_CHECK_LAUNCH();
  return TULIP_OK;
}
