// GEMM argument blocks + epilogue math of the tcgen05 GEMMs (gemm_tc05.cu, gemm_tn_group.cu).
//
//   NT : out[M,N]  = A[M,K] . B[N,K]^T  (+bias, epilogue)      forward linears, dX = dY . Wt^T
//   TN : dW[N,K]  += dY[M,N]^T . X[M,K], db[N] += colsum(dY)   weight gradients (split over M, fp32 atomics)
#pragma once
#include "common.cuh"

enum GemmEpi : int {
  EPI_STORE = 0,      // out = acc + bias
  EPI_GELU = 1,       // out2 = pre (acc+bias), out = gelu(pre)            (Mlp.fc1, tulip.py:195-196)
  EPI_RESID = 2,      // out = aux + row_scale[sample] * (acc + bias)      (residual + DropPath, tulip.py:343-344,350-351)
  EPI_PIXSHUF = 3,    // NHWC PixelShuffle(2) scatter of (acc + bias)      (PatchUnmerging, tulip.py:117-123)
  EPI_SPLIT2 = 4,     // cols < split_col -> out, others -> out2            (skip-Linear backward: d[x | skip])
  EPI_DGELU = 5,      // out = acc * gelu'(aux)                             (backward through GELU)
  EPI_HEAD = 6,       // pred[m, ij] (+)= sum_c wd[c] * leaky(acc + bias)   (PixelShuffleHead + decoder_pred, tulip.py:174-178,731)
                      // hd_ln: pred[m, ij] = sum_c wd[c] * LayerNorm(acc)[c]      (FinalPatchExpanding, tulip.py:152-159)
  EPI_HEAD_BWD = 7,   // recompute pre; out = dh (bf16), dwd += colsum(dpred * leaky(pre)); hd_ln: LayerNorm backward instead
  EPI_ROWSCALE = 8,   // out = row_scale[sample] * acc                      (DropPath backward on a branch dX)
  EPI_DGELU2 = 9,     // out = (A . B^T) * gelu'(A2 . B2^T + bias): the GELU pre-activation is RECOMPUTED by a second
                      // accumulator instead of being saved by the forward pass (tcgen05 path only)
  EPI_LNBWD = 10,     // the output rows (N = 96 or 192: one tile holds whole rows) are dL/dy of a LayerNorm whose input rows are
                      // `aux`:  out = [aux2 +] rstd * (g - mean(g) - xhat * mean(g * xhat)), g = acc * ln_w;  out2 = row_scale * out;
                      // ln_dw += colsum(acc * xhat), ln_db += colsum(acc)   (tcgen05 path only; replaces GEMM + layernorm_bwd)
  EPI_STORE_LN = 11,  // EPI_STORE / EPI_RESID whose output rows (N = 96 or 192) feed a LayerNorm next: also writes
  EPI_RESID_LN = 12,  // ln_y = LayerNorm(out) (from the rounded rows, as layernorm_fwd reads them) and its (mean, rstd) rows
};

enum GemmAMode : int {
  A_PLAIN = 0,
  A_UNSHUFFLE = 1,    // A[m=(b,h,w), k=ij*Cc+c] = src[(b,2h+i,2w+j), c]   (PixelShuffle(2) backward gather)
};

struct GemmArgs {
  // operands
  const bf16* A; long lda;
  const bf16* A2; long lda2; int K1;     // k >= K1 is read from A2 at column k-K1 (two-source concat); K1 == K when unused
  const bf16* B; long ldb;
  const bf16* B2; long ldb2; int K2;     // EPI_DGELU2: second product A2[M,K2] . B2[N,K2]^T
  int M, N, K;
  int a_mode; int g_H, g_W, g_Cc;        // gather geometry (input grid H x W, Cc channels per shuffled pixel)
  const float* bias;
  // epilogue
  bf16* out; long ldo;
  bf16* out2; long ldo2;
  const bf16* aux; long ldaux;
  const float* row_scale; int rows_per_sample;
  int split_col;
  // head
  const float* wd; const float* target; float* pred; const float* gscale;
  float* dwd; int dwd_copies;              // EPI_HEAD_BWD: CTA b adds into copy b % dwd_copies (stride hd_E); 0/1 = in place
  int hd_H, hd_W, hd_r, hd_E; float hd_inv_npix;
  // hd_ln = 1: FinalPatchExpanding head (tulip.py:144-159, 727-731) instead of PixelShuffleHead: no bias, no LeakyReLU;
  // the hd_E channels of a pixel go through LayerNorm(ln_w, ln_b, ln_eps) before the decoder_pred dot product.
  // EPI_HEAD also writes the pixel's (mean, rstd) to ln_ystats [pixels, 2]; EPI_HEAD_BWD reads them from ln_stats and
  // accumulates d(gamma) -> ln_dw, d(beta) -> ln_db, d(decoder_pred.weight) -> dwd, all with ln_copies / ln_stride.
  int hd_ln;
  // EPI_LNBWD
  const bf16* aux2; long ldaux2;           // residual-path gradient added to dx, or null
  const float* ln_w; const float* ln_stats;    // gamma [N], (mean, rstd) per row [M, 2]
  float* ln_dw; float* ln_db; int ln_copies, ln_stride;   // CTA b adds into copy b % ln_copies (ln_stride floats apart)
  // EPI_STORE_LN / EPI_RESID_LN: ln_w = gamma, ln_b = beta, ln_y [M, N] bf16, ln_ystats [M, 2]
  const float* ln_b; bf16* ln_y; float* ln_ystats; float ln_eps;
};

struct GemmTNArgs {
  const bf16* dY; long ldy;              // [M, N]
  const bf16* X; long ldx;               // [M, K1]
  const bf16* X2; long ldx2; int K1;     // columns >= K1 of the (virtual) X come from X2
  int M, N, K;
  int y_mode; int g_H, g_W, g_Cc;        // A_UNSHUFFLE gather on dY
  float* dW; long lddw;                  // [N, K] fp32, accumulated with atomics
  float* db;                             // [N] or null
  int perm_R2, perm_Cc;                  // row n' = ij*Cc + c is written to row c*R2 + ij (R2 == 1: identity)
  int splits;
  int hint;                              // set by the launcher: operand loads carry the L2 evict_first priority (tulip_hints() & 1)
};

int gemm_nt_tc05(const GemmArgs& g, int epi, cudaStream_t st);   // returns TULIP_ERR_UNSUPPORTED for shapes it does not take
int gemm_tn_tc05(const GemmTNArgs& g, cudaStream_t st);
int gemm_nt_tc05_plan(int M, int N, int K, int epi, int save_pre, int* out10);   // host-side tiling decision (tests, tooling)
int gemm_nt_pairs_mode(int mode);                                // CTA-pair schedule: 0 off, 1 every eligible launch, 2 K >= 384; returns the previous mode
int gemm_nt(const GemmArgs& g, int epi, cudaStream_t st);        // the tcgen05 GEMM; an unsupported shape is an error (no second backend)
int gemm_tn(const GemmTNArgs& g, cudaStream_t st);
// Several independent weight gradients in ONE persistent launch (gemm_tn_group.cu).  Problems must be "plain" (no concat, no
// PixelShuffle gather): gemm_tn_groupable() says so; per_out / items_out of the plan are for tests and tooling.
constexpr int TN_GROUP_MAX = 4;
bool gemm_tn_groupable(const GemmTNArgs& g);
int gemm_tn_group(const GemmTNArgs* gs, int n, cudaStream_t st);
int gemm_tn_group_plan(const int* M, const int* N, const int* K, int n, int sms, int* per_out, int* items_out);
bool gemm_nt_lnbwd_supported(int M, int N, int K);               // EPI_LNBWD takes this shape (else: EPI_STORE + layernorm_bwd)
bool gemm_nt_lnfwd_supported(int M, int N, int K);               // EPI_STORE_LN / EPI_RESID_LN take this shape

// ---- epilogue math on a run of NV consecutive columns of one output row (shared by both GEMMs) ----

__device__ __forceinline__ void store_bf16_run2(bf16* p, float a, float b) {
  *reinterpret_cast<uint32_t*>(p) = pack_bf16(a, b);
}

// pixel-shuffle destination row for source row m and shuffle slot ij (r = 2)
__device__ __forceinline__ long pixshuf_row(int m, int ij, int H, int W) {
  const int w = m % W;
  const int bh = m / W;
  const int h = bh % H;
  const int b = bh / H;
  return ((long)(b * 2 * H + 2 * h + (ij >> 1)) * (2 * W) + 2 * w + (ij & 1));
}

// Handles two adjacent columns (n, n+1) of row m. n is even, so both columns share every group boundary used here.
template <int EPI>
__device__ __forceinline__ void epi_pair(const GemmArgs& g, int m, int n, float v0, float v1) {
  if (EPI == EPI_STORE) {
    if (g.bias) { v0 += g.bias[n]; v1 += g.bias[n + 1]; }
    store_bf16_run2(g.out + (long)m * g.ldo + n, v0, v1);
  } else if (EPI == EPI_GELU) {
    v0 += g.bias[n]; v1 += g.bias[n + 1];
    if (g.out2) store_bf16_run2(g.out2 + (long)m * g.ldo2 + n, v0, v1);
    store_bf16_run2(g.out + (long)m * g.ldo + n, gelu_erf(v0), gelu_erf(v1));
  } else if (EPI == EPI_RESID) {
    if (g.bias) { v0 += g.bias[n]; v1 += g.bias[n + 1]; }
    const float s = g.row_scale ? g.row_scale[m / g.rows_per_sample] : 1.0f;
    const float2 r = unpack_bf16(*reinterpret_cast<const uint32_t*>(g.aux + (long)m * g.ldaux + n));
    store_bf16_run2(g.out + (long)m * g.ldo + n, r.x + s * v0, r.y + s * v1);
  } else if (EPI == EPI_PIXSHUF) {
    if (g.bias) { v0 += g.bias[n]; v1 += g.bias[n + 1]; }
    const int ij = n / g.g_Cc, c = n % g.g_Cc;
    store_bf16_run2(g.out + pixshuf_row(m, ij, g.g_H, g.g_W) * g.ldo + c, v0, v1);
  } else if (EPI == EPI_SPLIT2) {
    if (n < g.split_col) store_bf16_run2(g.out + (long)m * g.ldo + n, v0, v1);
    else store_bf16_run2(g.out2 + (long)m * g.ldo2 + (n - g.split_col), v0, v1);
  } else if (EPI == EPI_DGELU) {
    const float2 p = unpack_bf16(*reinterpret_cast<const uint32_t*>(g.aux + (long)m * g.ldaux + n));
    store_bf16_run2(g.out + (long)m * g.ldo + n, v0 * gelu_erf_grad(p.x), v1 * gelu_erf_grad(p.y));
  } else if (EPI == EPI_ROWSCALE) {
    const float s = g.row_scale ? g.row_scale[m / g.rows_per_sample] : 1.0f;
    store_bf16_run2(g.out + (long)m * g.ldo + n, s * v0, s * v1);
  }
}

// head geometry: row m = (b, h, w) of the low-res grid, slot ij = i*r + j -> pixel (b, h*r+i, w*r+j)
__device__ __forceinline__ long head_pixel(const GemmArgs& g, int m, int ij) {
  const int w = m % g.hd_W;
  const int bh = m / g.hd_W;
  const int h = bh % g.hd_H;
  const int b = bh / g.hd_H;
  const int r = g.hd_r;
  return ((long)(b * g.hd_H * r + h * r + ij / r) * (g.hd_W * r) + w * r + (ij % r));
}

__device__ __forceinline__ float leaky(float x) { return x > 0.f ? x : 0.01f * x; }
