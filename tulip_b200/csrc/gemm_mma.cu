// Legacy warp-level tensor-core GEMMs (mma.sync m16n8k16 bf16, cp.async 3-stage pipeline).
// Used for shapes the tcgen05 kernels do not take (tiny M) and as the cross-check
// implementation (env TULIP_B200_GEMM=mma).  Every N and K on the TULIP path is a multiple
// of 96 and 32 respectively (C = 96 * 2^s), which fixes the tile shape: 128 x 96 x 32.
#include "gemm.cuh"

namespace {

constexpr int BM = 128, BN = 96, BK = 32, STAGES = 3, PAD = 8;
constexpr int LDS = BK + PAD;                               // smem row stride in elements (80 B: conflict-free ldmatrix)
constexpr int NT_SMEM = STAGES * (BM + BN) * LDS * 2;       // 53,760 B

template <int AMODE>
__device__ __forceinline__ const bf16* a_src(const GemmArgs& g, int m, int k) {
  if (AMODE == A_UNSHUFFLE) {
    const int ij = k / g.g_Cc, c = k % g.g_Cc;
    return g.A + pixshuf_row(m, ij, g.g_H, g.g_W) * g.lda + c;
  }
  if (k >= g.K1) return g.A2 + (long)m * g.lda2 + (k - g.K1);
  return g.A + (long)m * g.lda + k;
}

template <int EPI, int AMODE>
__global__ void __launch_bounds__(128) gemm_nt_mma_kernel(const GemmArgs g) {
  pdl_sync();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  bf16* sA = reinterpret_cast<bf16*>(smem_raw);
  bf16* sB = sA + STAGES * BM * LDS;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int KT = g.K / BK;

  auto load_stage = [&](int stage, int kt) {
    const int k0 = kt * BK;
    bf16* a = sA + stage * BM * LDS;
    bf16* b = sB + stage * BN * LDS;
#pragma unroll
    for (int i = 0; i < (BM * 4) / 128; ++i) {
      const int ch = tid + i * 128;
      const int r = ch >> 2, c = (ch & 3) * 8;
      const int m = m0 + r;
      const bool ok = m < g.M;
      cp_async16(a + r * LDS + c, a_src<AMODE>(g, ok ? m : 0, k0 + c), ok ? 16 : 0);
    }
#pragma unroll
    for (int i = 0; i < (BN * 4) / 128; ++i) {
      const int ch = tid + i * 128;
      const int r = ch >> 2, c = (ch & 3) * 8;
      cp_async16(b + r * LDS + c, g.B + (long)(n0 + r) * g.ldb + k0 + c, 16);
    }
  };

  float acc[2][12][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 12; ++j)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[i][j][q] = 0.f;

#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < KT) load_stage(s, s);
    cp_async_commit();
  }
  for (int kt = 0; kt < KT; ++kt) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    if (kt + STAGES - 1 < KT) load_stage((kt + STAGES - 1) % STAGES, kt + STAGES - 1);
    cp_async_commit();
    const bf16* a = sA + (kt % STAGES) * BM * LDS + (warp * 32) * LDS;
    const bf16* b = sB + (kt % STAGES) * BN * LDS;
#pragma unroll
    for (int ks = 0; ks < BK / 16; ++ks) {
      uint32_t af[2][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
        ldmatrix_x4(af[mt], a + (mt * 16 + (lane & 15)) * LDS + ks * 16 + (lane >> 4) * 8);
#pragma unroll
      for (int np = 0; np < 6; ++np) {
        uint32_t bfr[4];
        const int mat = lane >> 3;
        ldmatrix_x4(bfr, b + (np * 16 + (mat >> 1) * 8 + (lane & 7)) * LDS + ks * 16 + (mat & 1) * 8);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          mma_bf16_16816(acc[mt][2 * np], af[mt], bfr[0], bfr[1]);
          mma_bf16_16816(acc[mt][2 * np + 1], af[mt], bfr[2], bfr[3]);
        }
      }
    }
  }
  cp_async_wait<0>();

  const int gq = lane >> 2, tq = lane & 3;
  if (EPI == EPI_HEAD || EPI == EPI_HEAD_BWD) {
    // One CTA tile = 96 consecutive expanded channels n' = ij*E + c of 128 low-res pixels.
    const int ij = n0 / g.hd_E;
    const int c0 = n0 % g.hd_E;
    float wd[12][2], bs[12][2];
#pragma unroll
    for (int nt = 0; nt < 12; ++nt)
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        wd[nt][q] = g.wd[c0 + nt * 8 + 2 * tq + q];
        bs[nt][q] = g.bias[n0 + nt * 8 + 2 * tq + q];
      }
    if (EPI == EPI_HEAD) {
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          float s = 0.f;
#pragma unroll
          for (int nt = 0; nt < 12; ++nt)
#pragma unroll
            for (int q = 0; q < 2; ++q) s += wd[nt][q] * leaky(acc[mt][nt][hf * 2 + q] + bs[nt][q]);
          s += __shfl_xor_sync(0xffffffffu, s, 1);
          s += __shfl_xor_sync(0xffffffffu, s, 2);
          const int m = m0 + warp * 32 + mt * 16 + hf * 8 + gq;
          if (tq == 0 && m < g.M) {
            float* p = g.pred + head_pixel(g, m, ij);
            if (g.hd_E == BN) *p = s; else atomicAdd(p, s);           // E > 96: pred is zeroed by the caller
          }
        }
    } else {
      float cw[12][2];                                              // column partial sums for dwd
#pragma unroll
      for (int nt = 0; nt < 12; ++nt) { cw[nt][0] = cw[nt][1] = 0.f; }
      const float gs = g.gscale[0] * g.hd_inv_npix;
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          const int m = m0 + warp * 32 + mt * 16 + hf * 8 + gq;
          float dp = 0.f;
          if (m < g.M) {
            const long px = head_pixel(g, m, ij);
            const float d = g.pred[px] - g.target[px];
            dp = (d > 0.f ? gs : (d < 0.f ? -gs : 0.f));
          }
#pragma unroll
          for (int nt = 0; nt < 12; ++nt) {
            float dh[2];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              const float pre = acc[mt][nt][hf * 2 + q] + bs[nt][q];
              cw[nt][q] += dp * leaky(pre);
              dh[q] = dp * wd[nt][q] * (pre > 0.f ? 1.f : 0.01f);
            }
            if (m < g.M) store_bf16_run2(g.out + (long)m * g.ldo + n0 + nt * 8 + 2 * tq, dh[0], dh[1]);
          }
        }
#pragma unroll
      for (int nt = 0; nt < 12; ++nt)
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          float a = cw[nt][q];
#pragma unroll
          for (int o = 4; o < 32; o <<= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
          if (gq == 0) atomicAdd(g.dwd + c0 + nt * 8 + 2 * tq + q, a);
        }
    }
    return;
  }
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      const int m = m0 + warp * 32 + mt * 16 + hf * 8 + gq;
      if (m >= g.M) continue;
#pragma unroll
      for (int nt = 0; nt < 12; ++nt)
        epi_pair<EPI>(g, m, n0 + nt * 8 + 2 * tq, acc[mt][nt][hf * 2], acc[mt][nt][hf * 2 + 1]);
    }
}

template <int EPI, int AMODE>
int launch_nt(const GemmArgs& g, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    TULIP_CUDA(cudaFuncSetAttribute(gemm_nt_mma_kernel<EPI, AMODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, NT_SMEM));
    configured = true;
  }
  dim3 grid(ceil_div(g.M, BM), g.N / BN);
  tulip_launch(gemm_nt_mma_kernel<EPI, AMODE>, grid, 128, NT_SMEM, st, g);
  TULIP_CHECK_LAUNCH();
  return TULIP_OK;
}

// ------------------------------------------------------------------------------------------------
// TN: dW[N,K] += dY[M,N]^T X[M,K] over a slice of M per CTA; fp32 atomics into the gradient buffer.
constexpr int TBN = 96, TBK = 96, TBM = 32, TLD = 96 + 8;     // token-chunk rows are 104 elements (208 B)
constexpr int TN_SMEM = STAGES * 2 * TBM * TLD * 2;           // 39,936 B

template <int YMODE>
__global__ void __launch_bounds__(128) gemm_tn_mma_kernel(const GemmTNArgs g) {
  pdl_sync();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  bf16* sY = reinterpret_cast<bf16*>(smem_raw);
  bf16* sX = sY + STAGES * TBM * TLD;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n0 = blockIdx.x * TBN, k0 = blockIdx.y * TBK;
  const int chunks_total = ceil_div(g.M, TBM);
  const int per = ceil_div(chunks_total, g.splits);
  const int c_begin = blockIdx.z * per;
  const int c_end = min(chunks_total, c_begin + per);
  const int NC = c_end - c_begin;
  if (NC <= 0) return;
  const bool do_db = (g.db != nullptr) && blockIdx.y == 0;
  float dbsum = 0.f;

  auto load_stage = [&](int stage, int ci) {
    const int mbase = (c_begin + ci) * TBM;
    bf16* y = sY + stage * TBM * TLD;
    bf16* x = sX + stage * TBM * TLD;
#pragma unroll
    for (int i = 0; i < (TBM * 12) / 128; ++i) {
      const int ch = tid + i * 128;
      const int r = ch / 12, c = (ch % 12) * 8;
      const int m = mbase + r;
      const bool ok = m < g.M;
      const int mm = ok ? m : 0;
      const bf16* ysrc;
      if (YMODE == A_UNSHUFFLE) {
        const int n = n0 + c;
        ysrc = g.dY + pixshuf_row(mm, n / g.g_Cc, g.g_H, g.g_W) * g.ldy + (n % g.g_Cc);
      } else {
        ysrc = g.dY + (long)mm * g.ldy + n0 + c;
      }
      cp_async16(y + r * TLD + c, ysrc, ok ? 16 : 0);
      const int k = k0 + c;
      const bf16* xsrc = (k >= g.K1) ? g.X2 + (long)mm * g.ldx2 + (k - g.K1) : g.X + (long)mm * g.ldx + k;
      cp_async16(x + r * TLD + c, xsrc, ok ? 16 : 0);
    }
  };

  const int wn = (warp >> 1) * 48, wk = (warp & 1) * 48;        // warp tile: 48 (n) x 48 (k)
  float acc[3][6][4];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 6; ++j)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[i][j][q] = 0.f;

#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < NC) load_stage(s, s);
    cp_async_commit();
  }
  for (int ci = 0; ci < NC; ++ci) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    if (ci + STAGES - 1 < NC) load_stage((ci + STAGES - 1) % STAGES, ci + STAGES - 1);
    cp_async_commit();
    const bf16* y = sY + (ci % STAGES) * TBM * TLD;
    const bf16* x = sX + (ci % STAGES) * TBM * TLD;
    if (do_db && tid < TBN) {
#pragma unroll 8
      for (int r = 0; r < TBM; ++r) dbsum += __bfloat162float(y[r * TLD + tid]);
    }
#pragma unroll
    for (int ks = 0; ks < TBM / 16; ++ks) {
      // A fragment = dY^T: rows are output rows n, contraction index is the token (stored as smem rows)
      uint32_t af[3][4];
      const int mat = lane >> 3;
#pragma unroll
      for (int mt = 0; mt < 3; ++mt)
        ldmatrix_x4_trans(af[mt], y + (ks * 16 + (mat >> 1) * 8 + (lane & 7)) * TLD + wn + mt * 16 + (mat & 1) * 8);
#pragma unroll
      for (int np = 0; np < 3; ++np) {
        uint32_t bfr[4];
        ldmatrix_x4_trans(bfr, x + (ks * 16 + (mat & 1) * 8 + (lane & 7)) * TLD + wk + np * 16 + (mat >> 1) * 8);
#pragma unroll
        for (int mt = 0; mt < 3; ++mt) {
          mma_bf16_16816(acc[mt][2 * np], af[mt], bfr[0], bfr[1]);
          mma_bf16_16816(acc[mt][2 * np + 1], af[mt], bfr[2], bfr[3]);
        }
      }
    }
  }
  cp_async_wait<0>();

  const int gq = lane >> 2, tq = lane & 3;
#pragma unroll
  for (int mt = 0; mt < 3; ++mt)
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      const int n = n0 + wn + mt * 16 + hf * 8 + gq;
      const int row = (g.perm_R2 > 1) ? (n % g.perm_Cc) * g.perm_R2 + n / g.perm_Cc : n;
      float* dst = g.dW + (long)row * g.lddw + k0 + wk + 2 * tq;
#pragma unroll
      for (int nt = 0; nt < 6; ++nt) {
        atomicAdd(dst + nt * 8, acc[mt][nt][hf * 2]);
        atomicAdd(dst + nt * 8 + 1, acc[mt][nt][hf * 2 + 1]);
      }
    }
  if (do_db && tid < TBN) {
    const int n = n0 + tid;
    const int row = (g.perm_R2 > 1) ? (n % g.perm_Cc) * g.perm_R2 + n / g.perm_Cc : n;
    atomicAdd(g.db + row, dbsum);
  }
}

}  // namespace

int gemm_nt_mma(const GemmArgs& g, int epi, cudaStream_t st) {
  TULIP_REQUIRE(g.N % BN == 0 && g.K % BK == 0 && g.K1 % BK == 0, "gemm_nt: N must be a multiple of 96 and K of 32");
  TULIP_REQUIRE(g.M > 0, "gemm_nt: empty M");
  if (g.a_mode == A_UNSHUFFLE) {
    TULIP_REQUIRE(epi == EPI_STORE, "gemm_nt: unshuffle gather only with the plain epilogue");
    return launch_nt<EPI_STORE, A_UNSHUFFLE>(g, st);
  }
  switch (epi) {
    case EPI_STORE: return launch_nt<EPI_STORE, A_PLAIN>(g, st);
    case EPI_GELU: return launch_nt<EPI_GELU, A_PLAIN>(g, st);
    case EPI_RESID: return launch_nt<EPI_RESID, A_PLAIN>(g, st);
    case EPI_PIXSHUF: return launch_nt<EPI_PIXSHUF, A_PLAIN>(g, st);
    case EPI_SPLIT2: return launch_nt<EPI_SPLIT2, A_PLAIN>(g, st);
    case EPI_DGELU: return launch_nt<EPI_DGELU, A_PLAIN>(g, st);
    case EPI_HEAD: return launch_nt<EPI_HEAD, A_PLAIN>(g, st);
    case EPI_HEAD_BWD: return launch_nt<EPI_HEAD_BWD, A_PLAIN>(g, st);
    case EPI_ROWSCALE: return launch_nt<EPI_ROWSCALE, A_PLAIN>(g, st);
  }
  if (epi == EPI_DGELU2 || epi == EPI_LNBWD || epi == EPI_STORE_LN || epi == EPI_RESID_LN) {
    tulip_set_error("gemm_nt: EPI_DGELU2 and the fused-LayerNorm epilogues exist on the tcgen05 path only (the latter for N = 96 or 192)");
    return TULIP_ERR_UNSUPPORTED;
  }
  tulip_set_error("gemm_nt: unknown epilogue");
  return TULIP_ERR_ARG;
}

int gemm_tn_mma(const GemmTNArgs& g, cudaStream_t st) {
  TULIP_REQUIRE(g.N % TBN == 0 && g.K % TBK == 0 && g.K1 % 8 == 0, "gemm_tn: N and K must be multiples of 96");
  TULIP_REQUIRE(g.K1 % TBK == 0 || g.K1 == g.K, "gemm_tn: concat split must fall on a 96-column tile edge");
  TULIP_REQUIRE(g.M > 0 && g.splits > 0, "gemm_tn: empty problem");
  static bool configured = false;
  if (!configured) {
    TULIP_CUDA(cudaFuncSetAttribute(gemm_tn_mma_kernel<A_PLAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, TN_SMEM));
    TULIP_CUDA(cudaFuncSetAttribute(gemm_tn_mma_kernel<A_UNSHUFFLE>, cudaFuncAttributeMaxDynamicSharedMemorySize, TN_SMEM));
    configured = true;
  }
  dim3 grid(g.N / TBN, g.K / TBK, g.splits);
  if (g.y_mode == A_UNSHUFFLE) tulip_launch(gemm_tn_mma_kernel<A_UNSHUFFLE>, grid, 128, TN_SMEM, st, g);
  else tulip_launch(gemm_tn_mma_kernel<A_PLAIN>, grid, 128, TN_SMEM, st, g);
  TULIP_CHECK_LAUNCH();
  return TULIP_OK;
}
