// Fused W-MSA / SW-MSA half-block, forward:   y = x + s_b * proj(attn(qkv(LN1(x))))
//
// Reference: tulip/model/tulip.py:338-346 (SwinTransformerBlock.forward, attention half) with WindowAttention.forward
// (:282-324) inlined: LayerNorm(eps 1e-6) -> roll -> window_partition -> qkv Linear -> per (window, head)
// softmax(q k^T * scale + rel-pos bias (+ shift mask)) v -> head concat -> proj Linear -> window reverse -> roll back ->
// DropPath scale -> residual.  ONE launch; x is read once and y written once: 4 T C bytes of HBM traffic instead of the
// 26 T C bytes of the LayerNorm / GEMM / attention / GEMM chain.
//
// A persistent CTA walks 128-token tiles = 8 windows; tile row r is token r % 16 of window r / 16, so roll, partition and
// reverse are coordinates: a window is ONE 4-D TMA box {C, Mw, Mh, 1} of the NHWC tensor at its rolled position (windows that
// straddle the cyclic seam take one box per half row).  Five warpgroups; registers are re-balanced with setmaxnreg:
//   warp 0        tcgen05 issuer: loads the bf16 weights once by TMA (resident in shared memory for the CTA's life) and
//                 issues QKV = LN(x) . Wqkv^T (M 128, N 3 x 96, K 96) and the projection (M 128, N 96, K 96), fp32 in TMEM
//   warp 1        loader: TMA window boxes of x into a ring of three raw tiles (mbarrier transaction bytes)
//   warp 2        storer: the finished y tile (written in place over its raw tile) back to global memory as the same boxes
//   warps 4..7    LayerNorm: one thread per tile row, raw row -> K-major, 64B-swizzled A operand tile (double-buffered)
//   warps 8..19   attention: warp (q, head) owns TMEM lane quarter q = tile windows 2q, 2q+1 and one head.  It reads q, k, v of
//                 a (window, head) unit as MMA fragments DIRECTLY from the accumulator (tcgen05.ld.16x256b hands out the
//                 m16n8 C-fragment layout), so S = q k^T, softmax, O = P v run in registers with m16n8k16 warp MMAs
//                 (a 16-token x 32-dim problem cannot feed a tcgen05 tile); v is turned into the
//                 B operand of P v with movmatrix; O goes to shared memory as the A operand of the projection.  The same
//                 warps run the epilogue of the previous tile (projection accumulator + bias, DropPath scale, residual
//                 from the raw tile, store), 32 rows x 32 columns each.
// The phases of consecutive tiles overlap through mbarriers: QKV of tile i+1 is issued as soon as the attention warps
// hold tile i's fragments, the projection / epilogue of tile i run while tile i+1 is in attention, raw tiles arrive two
// steps ahead of their LayerNorm.
#include "fused.cuh"
#include "kernels.h"
#include "window_index.cuh"

#include <cstdlib>
#include <cstring>

namespace {
using namespace fused;

constexpr int WC = 96;                      // channels handled by this build of the kernel (stage 0 of both factories)
constexpr int WH = WC / 32;                 // heads
constexpr int WM_WARPS = 20, WM_THREADS = 32 * WM_WARPS;     // warps are allocated to a CTA four at a time
constexpr int AT_WARPS = 12, IO_WARPS = 4, IO_WARP0 = 4, AT_WARP0 = 8;
constexpr int REGS_ISSUER = 24, REGS_IO = 72, REGS_AT = 128;    // 128 x (24 + 72) + 384 x 128 = 61440 = the CTA's pool (640 x 96)
constexpr int KBLK = 32;                    // K block: 32 bf16 = 64 B rows, SWIZZLE_64B
constexpr int NKB = WC / KBLK;              // 3
constexpr int A_BLK = 128 * 64;             // one K block of a 128-row operand tile: 8 KB
constexpr int W_BLK = 96 * 64;              // one K block of a 96-row weight chunk: 6 KB
constexpr int ROWB = WC * 2;                // bytes of a raw row
constexpr int RAW_TILE = 128 * ROWB;        // 24 KB, row-major (what the TMA window boxes produce)
constexpr int RAW_RING = 4;
constexpr int OFF_WQ = 0;                                   // [3 n-chunks][3 K blocks][96 x 32]
constexpr int OFF_WP = OFF_WQ + 3 * NKB * W_BLK;            // [3 K blocks][96 x 32]
constexpr int OFF_A = OFF_WP + NKB * W_BLK;                 // [3 K blocks][128 x 32]: LayerNorm output, A operand of QKV
constexpr int OFF_O = OFF_A + NKB * A_BLK;                  // [3 K blocks][128 x 32]: attention output, A operand of the projection
constexpr int OFF_RAW = OFF_O + NKB * A_BLK;                // ring of raw x tiles: LayerNorm source, residual, then the y tile
constexpr int OFF_PAR = OFF_RAW + RAW_RING * RAW_TILE;      // fp32: qkv bias [288] | proj bias [96] | gamma [96] | beta [96]
constexpr int OFF_BAR = OFF_PAR + (3 * WC + 3 * WC) * 4;
constexpr int WM_SMEM = OFF_BAR + 256 + 1024;
constexpr int TM_QKV = 0, TM_P = 3 * WC;                    // TMEM columns: q|k|v accumulators, then two projection buffers
static_assert(TM_P + 2 * WC <= 512, "TMEM budget");
static_assert(WM_SMEM <= 227 * 1024, "shared memory budget");

struct WmsaArgs {
  const float* ln_w; const float* ln_b; const float* bqkv; const float* bproj; const float* bias_table;
  const float* row_scale;                   // [B] DropPath scales or null
  int B, H, W;
  int Mh, Mw, sh, sw, masked, bMh, bMw;
  int nWh, nWw; uint32_t mul_nWw, mul_nWh;  // windows per column / row and their division magics (fast_div)
  float eps, scale;
  int ntiles, nwin;                         // 128-row tiles (the last one may hold fewer than 8 windows) and windows
  long long* trace;                         // bring-up: clock64() stamps of CTA 0 (tulip_debug_wmsa_trace), null in production
};
#ifdef TULIP_WMSA_TRACE
#define WM_TRACE(role, it, k)                                                                      \
  do {                                                                                             \
    if (a.trace && blockIdx.x == 0 && (it) < 8 && lane == 0) a.trace[(((role) * 8 + (it)) * 8) + (k)] = clock64(); \
  } while (0)
#else
#define WM_TRACE(role, it, k) do { } while (0)
#endif

// window widx (row-major over (b, wh, ww)) -> sample and window coordinates
__device__ __forceinline__ void decode_window(const WmsaArgs& a, int widx, int& b, int& wh, int& ww) {
  const int q = (int)fast_div((uint32_t)widx, a.mul_nWw);
  ww = widx - q * a.nWw;
  b = (int)fast_div((uint32_t)q, a.mul_nWh);
  wh = q - b * a.nWh;
}

__global__ void __launch_bounds__(WM_THREADS, 1)
wmsa_block_fwd_kernel(const __grid_constant__ CUtensorMap mapWq, const __grid_constant__ CUtensorMap mapWp,
                      const __grid_constant__ CUtensorMap mapWin, const __grid_constant__ CUtensorMap mapSeg,
                      const __grid_constant__ CUtensorMap mapWinY, const __grid_constant__ CUtensorMap mapSegY,
                      const __grid_constant__ WmsaArgs a) {
  extern __shared__ unsigned char smem_raw[];
  pdl_trigger();
#ifdef TULIP_WMSA_TRACE
  if (a.trace && blockIdx.x == 0 && threadIdx.x == 0) a.trace[7 * 8 + 0] = clock64();      // kernel entry
#endif
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  float* spar = reinterpret_cast<float*>(smem + OFF_PAR);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* w_full = bars;            // weights landed (TMA bytes)
  uint64_t* a_full = bars + 1;        // A tile written by the LayerNorm warps
  uint64_t* a_empty = bars + 2;       // QKV MMAs have read the A tile
  uint64_t* qkv_full = bars + 3;      // q|k|v accumulators complete
  uint64_t* qkv_empty = bars + 4;     // attention warps hold their fragments
  uint64_t* o_full = bars + 5;        // O tile written by the attention warps
  uint64_t* o_empty = bars + 6;       // projection MMAs have read the O tile
  uint64_t* p_full = bars + 7;        // [2] projection accumulator complete
  uint64_t* p_empty = bars + 9;       // [2] drained by the epilogue
  uint64_t* raw_full = bars + 11;     // [RAW_RING] raw tile landed (TMA bytes)
  uint64_t* raw_empty = bars + 15;    // [RAW_RING] the y tile that replaced the raw tile in place has been stored
  uint64_t* y_full = bars + 19;       // [RAW_RING] epilogue has turned the raw tile into the y tile
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 23);

  const int warp = tc::warp_idx_sync(), lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    tc::mbar_init(w_full, 1);
    tc::mbar_init(a_full, IO_WARPS); tc::mbar_init(a_empty, 1);
    for (int b = 0; b < 2; ++b) { tc::mbar_init(p_full + b, 1); tc::mbar_init(p_empty + b, AT_WARPS); }
    for (int b = 0; b < RAW_RING; ++b) { tc::mbar_init(raw_full + b, 1); tc::mbar_init(raw_empty + b, 1); tc::mbar_init(y_full + b, AT_WARPS); }
    tc::mbar_init(qkv_full, 1); tc::mbar_init(qkv_empty, AT_WARPS);
    tc::mbar_init(o_full, AT_WARPS); tc::mbar_init(o_empty, 1);
    tc::fence_barrier_init();
  }
  if (warp == 0) tc::tmem_alloc<512>(tmem_slot);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
#ifdef TULIP_WMSA_TRACE
  if (a.trace && blockIdx.x == 0 && threadIdx.x == 0) a.trace[7 * 8 + 1] = clock64();      // barriers + TMEM ready
#endif
  const int my_tiles = (a.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp < IO_WARP0) {
    reg_dec<REGS_ISSUER>();
  }
  if (warp == 0) {
    // ------------------------------------------------------------------ TMA (weights, once) + tcgen05 issuer
    if (tc::elect_one_sync()) {
      tc::prefetch_tensormap(&mapWq);
      tc::prefetch_tensormap(&mapWp);
      tc::mbar_expect_tx(w_full, (3 * NKB + NKB) * W_BLK);
      for (int nc = 0; nc < 3; ++nc)
        for (int kb = 0; kb < NKB; ++kb)
          tc::tma_load_2d(smem + OFF_WQ + (nc * NKB + kb) * W_BLK, &mapWq, w_full, kb * KBLK, nc * WC);
      for (int kb = 0; kb < NKB; ++kb) tc::tma_load_2d(smem + OFF_WP + kb * W_BLK, &mapWp, w_full, kb * KBLK, 0);
    }
    __syncwarp();
    tc::mbar_wait(w_full, 0);
    constexpr uint32_t idesc = tc::make_idesc(128, WC, 0, 0);
    const uint64_t dWq = desc_k_sw64(smem + OFF_WQ), dWp = desc_k_sw64(smem + OFF_WP);
    const uint64_t dA0 = desc_k_sw64(smem + OFF_A), dO = desc_k_sw64(smem + OFF_O);
    auto issue_proj = [&](int j) {
      const int pb = j & 1;
      tc::mbar_wait(o_full, j & 1);
      WM_TRACE(0, j, 3);
      tc::mbar_wait(p_empty + pb, ((j >> 1) & 1) ^ 1);
      tc::fence_after_sync();
      if (tc::elect_one_sync()) {
#pragma unroll
        for (int kb = 0; kb < NKB; ++kb)
#pragma unroll
          for (int ks = 0; ks < 2; ++ks)
            tc::umma_bf16(tmem_base + TM_P + pb * WC, dO + (uint64_t)((kb * A_BLK) >> 4) + 2 * ks,
                          dWp + (uint64_t)((kb * W_BLK) >> 4) + 2 * ks, idesc, (kb | ks) ? 1u : 0u);
        tc::umma_commit(o_empty);
        tc::umma_commit(p_full + pb);
      }
      __syncwarp();
      WM_TRACE(0, j, 4);
    };
    for (int it = 0; it < my_tiles; ++it) {
      WM_TRACE(0, it, 0);
      tc::mbar_wait(a_full, it & 1);
      WM_TRACE(0, it, 1);
      if (it > 0) tc::mbar_wait(qkv_empty, (it - 1) & 1);
      tc::fence_after_sync();
      WM_TRACE(0, it, 2);
      if (tc::elect_one_sync()) {
        const uint64_t dA = dA0;
#pragma unroll
        for (int nc = 0; nc < 3; ++nc)
#pragma unroll
          for (int kb = 0; kb < NKB; ++kb)
#pragma unroll
            for (int ks = 0; ks < 2; ++ks)
              tc::umma_bf16(tmem_base + TM_QKV + nc * WC, dA + (uint64_t)((kb * A_BLK) >> 4) + 2 * ks,
                            dWq + (uint64_t)(((nc * NKB + kb) * W_BLK) >> 4) + 2 * ks, idesc, (kb | ks) ? 1u : 0u);
        tc::umma_commit(a_empty);
        tc::umma_commit(qkv_full);
      }
      __syncwarp();
      if (it > 0) issue_proj(it - 1);
    }
    if (my_tiles > 0) issue_proj(my_tiles - 1);
  } else if (warp == 1) {
    // ------------------------------------------------------------------ loader: window boxes of x -> raw tile ring
    if (tc::elect_one_sync()) {
      tc::prefetch_tensormap(&mapWin);
      tc::prefetch_tensormap(&mapSeg);
    }
    pdl_wait();                                               // x is produced by the preceding kernel
    const int hw = a.Mw >> 1;                                 // half a window row: the seam of the cyclic shift cuts there
    for (int it = 0; it < my_tiles; ++it) {
      const int tile = blockIdx.x + it * gridDim.x, slot = it % RAW_RING;
      if (it >= RAW_RING) tc::mbar_wait(raw_empty + slot, ((it / RAW_RING) - 1) & 1);
      const int nv = min(8, a.nwin - tile * 8);               // windows of this tile (8 except in a partial last tile)
      if (tc::elect_one_sync()) {
        unsigned char* dst = smem + OFF_RAW + slot * RAW_TILE;
        tc::mbar_expect_tx(raw_full + slot, nv * 16 * ROWB);
        for (int w8 = 0; w8 < nv; ++w8) {
          int b, wh, ww;
          decode_window(a, tile * 8 + w8, b, wh, ww);
          const int h0 = wh * a.Mh + a.sh, w0 = ww * a.Mw + a.sw;      // rolled[h] = x[(h + sh) % H]   (tulip.py:290)
          if (h0 + a.Mh <= a.H && w0 + a.Mw <= a.W) {
            tc::tma_load_4d(dst + w8 * 16 * ROWB, &mapWin, raw_full + slot, 0, w0, h0, b);
          } else {
            for (int ri = 0; ri < a.Mh; ++ri) {
              int hh = h0 + ri;
              if (hh >= a.H) hh -= a.H;
              for (int half = 0; half < 2; ++half) {
                int wc = w0 + half * hw;
                if (wc >= a.W) wc -= a.W;
                tc::tma_load_4d(dst + (w8 * 16 + ri * a.Mw + half * hw) * ROWB, &mapSeg, raw_full + slot, 0, wc, hh, b);
              }
            }
          }
        }
      }
      __syncwarp();
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------------ storer: y tile (written in place over the raw tile) -> y
    if (tc::elect_one_sync()) {
      tc::prefetch_tensormap(&mapWinY);
      tc::prefetch_tensormap(&mapSegY);
    }
    const int hw = a.Mw >> 1;
    for (int it = 0; it < my_tiles; ++it) {
      const int tile = blockIdx.x + it * gridDim.x, slot = it % RAW_RING;
      tc::mbar_wait(y_full + slot, (it / RAW_RING) & 1);
      const int nv = min(8, a.nwin - tile * 8);
      if (tc::elect_one_sync()) {
        const unsigned char* src = smem + OFF_RAW + slot * RAW_TILE;
        for (int w8 = 0; w8 < nv; ++w8) {
          int b, wh, ww;
          decode_window(a, tile * 8 + w8, b, wh, ww);
          const int h0 = wh * a.Mh + a.sh, w0 = ww * a.Mw + a.sw;      // window reverse + roll back = the same box (tulip.py:320-323)
          if (h0 + a.Mh <= a.H && w0 + a.Mw <= a.W) {
            tc::tma_store_4d(&mapWinY, src + w8 * 16 * ROWB, 0, w0, h0, b);
          } else {
            for (int ri = 0; ri < a.Mh; ++ri) {
              int hh = h0 + ri;
              if (hh >= a.H) hh -= a.H;
              for (int half = 0; half < 2; ++half) {
                int wc = w0 + half * hw;
                if (wc >= a.W) wc -= a.W;
                tc::tma_store_4d(&mapSegY, src + (w8 * 16 + ri * a.Mw + half * hw) * ROWB, 0, wc, hh, b);
              }
            }
          }
        }
        tc::tma_store_commit();
        tc::tma_store_wait_read<0>();                           // shared memory read out: the slot may take the rows of step it + 3
        tc::mbar_arrive(raw_empty + slot);
      }
      __syncwarp();
    }
    if (tc::elect_one_sync()) tc::tma_store_wait<0>();          // global writes complete before the CTA retires
    __syncwarp();
  } else if (warp >= IO_WARP0 && warp < AT_WARP0) {
    // ------------------------------------------------------------------ IO warps: LayerNorm prologue + residual epilogue
    reg_dec<REGS_IO>();            // below the 96 registers every thread of a 640-thread CTA starts with
    const int q = warp & 3;
    const int r = q * 32 + lane;
    // parameters used by every tile: shared copies (fp32)
    for (int i = threadIdx.x - 32 * IO_WARP0; i < 6 * WC; i += 32 * IO_WARPS) {
      float v;
      if (i < 3 * WC) v = a.bqkv[i];
      else if (i < 4 * WC) v = a.bproj[i - 3 * WC];
      else if (i < 5 * WC) v = a.ln_w[i - 4 * WC];
      else v = a.ln_b[i - 5 * WC];
      spar[i] = v;
    }
    tc::named_bar_sync(1, 32 * (IO_WARPS + AT_WARPS));        // parameters visible to the IO and attention warps
    const float* s_g = spar + 4 * WC;
    const float* s_b = spar + 5 * WC;
    const int rot = (r >> 1) % 12;                            // chunk rotation: 8 consecutive rows hit 8 distinct bank groups

    auto normalise_tile = [&](int it) {
      const int slot = it % RAW_RING;
      const unsigned char* src = smem + OFF_RAW + slot * RAW_TILE + r * ROWB;
      unsigned char* dst = smem + OFF_A;
      if (warp == IO_WARP0) WM_TRACE(1, it, 3);
      tc::mbar_wait(raw_full + slot, (it / RAW_RING) & 1);
      if (warp == IO_WARP0) WM_TRACE(1, it, 4);
      // one pass over the row, shifted by its first element: sum (x - x0), sum (x - x0)^2
      const float x0 = unpack_bf16(*reinterpret_cast<const uint32_t*>(src)).x;
      float s1 = 0.f, s2 = 0.f;
      int k = rot;
#pragma unroll 4
      for (int c = 0; c < WC / 8; ++c) {
        const uint4 v = *reinterpret_cast<const uint4*>(src + k * 16);
        const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = unpack_bf16(w4[e]);
          const float d0 = f.x - x0, d1 = f.y - x0;
          s1 += d0 + d1;
          s2 = fmaf(d0, d0, s2);
          s2 = fmaf(d1, d1, s2);
        }
        k = (k == 11) ? 0 : k + 1;
      }
      const float dm = s1 * (1.0f / WC);
      const float mean = x0 + dm;
      const float rstd = rsqrtf(fmaxf(s2 * (1.0f / WC) - dm * dm, 0.f) + a.eps);
      if (it >= 1) tc::mbar_wait(a_empty, (it - 1) & 1);       // QKV of the previous tile has read the A tile
      if (warp == IO_WARP0) WM_TRACE(1, it, 5);
      asm volatile("" ::: "memory");                          // second pass re-reads the row from shared memory (no 48 live registers)
      k = rot;
#pragma unroll 2
      for (int c = 0; c < WC / 8; ++c) {
        const uint4 v = *reinterpret_cast<const uint4*>(src + k * 16);
        const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
        const float4 g0 = *reinterpret_cast<const float4*>(s_g + k * 8), g1 = *reinterpret_cast<const float4*>(s_g + k * 8 + 4);
        const float4 b0 = *reinterpret_cast<const float4*>(s_b + k * 8), b1 = *reinterpret_cast<const float4*>(s_b + k * 8 + 4);
        const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        uint32_t o4[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = unpack_bf16(w4[e]);
          o4[e] = pack_bf16(fmaf((f.x - mean) * rstd, gg[2 * e], bb[2 * e]), fmaf((f.y - mean) * rstd, gg[2 * e + 1], bb[2 * e + 1]));
        }
        *reinterpret_cast<uint4*>(dst + (k >> 2) * A_BLK + sw64_off(r, k & 3)) = make_uint4(o4[0], o4[1], o4[2], o4[3]);
        k = (k == 11) ? 0 : k + 1;
      }
      tc::fence_proxy_async();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(a_full);
      if (warp == IO_WARP0) WM_TRACE(1, it, 6);
    };
    for (int it = 0; it < my_tiles; ++it) normalise_tile(it);
  } else if (warp >= AT_WARP0) {
    // ------------------------------------------------------------------ attention warps
    reg_inc<REGS_AT>();
    const int q = warp & 3;                                   // TMEM lane quarter: tile windows 2q, 2q + 1
    const int head = (warp - AT_WARP0) >> 2;                  // 0..2
    const int g = lane >> 2, t = lane & 3;
    tc::named_bar_sync(1, 32 * (IO_WARPS + AT_WARPS));
    // relative-position bias of this head at this thread's fragment positions, and the seam masks (attention.cu)
    float bias[2][4];
    uint32_t diffH = 0, diffW = 0;
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int i = g + (e >> 1) * 8, j = nt * 8 + 2 * t + (e & 1);
        bias[nt][e] = a.bias_table[rel_bias_index(a.bMh, a.bMw, i, j) * WH + head] * 1.4426950408889634f;
        if (((i / a.Mw) >= a.Mh - a.sh) != ((j / a.Mw) >= a.Mh - a.sh)) diffH |= 1u << (nt * 4 + e);
        if (((i % a.Mw) >= a.Mw - a.sw) != ((j % a.Mw) >= a.Mw - a.sw)) diffW |= 1u << (nt * 4 + e);
      }
    // q|k|v bias at this thread's fragment columns (8 per tensor: cols 8j + 2t, +1 of the head's 32): re-read from shared
    // memory every tile instead of living in 24 registers
    const float* s_qb = spar + head * 32 + 2 * t;
    const float scale_l2e = a.scale * 1.4426950408889634f;
    unsigned char* sO = smem + OFF_O + head * A_BLK;           // K block `head` of the O tile: this head's 32 columns
    // Epilogue of a finished tile, spread over the 12 attention warps (one accumulator row per lane, this head's 32 columns):
    // projection accumulator + bias, DropPath scale and the residual, written IN PLACE over the raw tile; the storer warp sends
    // the finished tile to y as TMA window boxes (a lane-per-row store would touch 32 lines per instruction).
    const float* s_bp = spar + 3 * WC + head * 32;
    auto epilogue = [&](int it) {
      const int tile = blockIdx.x + it * gridDim.x, pb = it & 1, slot = it % RAW_RING;
      float rs = 1.0f;
      if (a.row_scale) {
        const int widx = min(tile * 8 + 2 * q + (lane >> 4), a.nwin - 1);
        int b, wh, ww;
        decode_window(a, widx, b, wh, ww);
        rs = a.row_scale[b];
      }
      unsigned char* xr = smem + OFF_RAW + slot * RAW_TILE + (q * 32 + lane) * ROWB + head * 64;   // residual in, y out (in place)
      if (warp == AT_WARP0) WM_TRACE(3, it, 0);
      tc::mbar_wait(p_full + pb, (it >> 1) & 1);
      tc::fence_after_sync();
      if (warp == AT_WARP0) WM_TRACE(3, it, 1);
      // four 16-byte chunks per lane, visited in an order rotated by (lane >> 1): 8 consecutive rows (192 B apart) then touch
      // 8 different bank groups (as in the LayerNorm passes).  TMEM addresses must be warp-uniform, so all 32 columns are
      // loaded and the chunk of each step is picked with selects.
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + TM_P + pb * WC + head * 32;
      const int rot = (lane >> 1) & 3;
      float v0[8], v1[8], v2[8], v3[8];
      tmem_ld_row8(taddr, v0);
      tmem_ld_row8(taddr + 8, v1);
      tmem_ld_row8(taddr + 16, v2);
      tmem_ld_row8(taddr + 24, v3);
      tmem_ld_wait();
      tmem_ld_use8(v0); tmem_ld_use8(v1); tmem_ld_use8(v2); tmem_ld_use8(v3);
      if (warp == AT_WARP0) WM_TRACE(3, it, 2);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int k = (c + rot) & 3;
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = sel4(k, v0[e], v1[e], v2[e], v3[e]);
        const uint4 xx = *reinterpret_cast<const uint4*>(xr + k * 16);
        const uint32_t w4[4] = {xx.x, xx.y, xx.z, xx.w};
        const float4 p0 = *reinterpret_cast<const float4*>(s_bp + k * 8), p1 = *reinterpret_cast<const float4*>(s_bp + k * 8 + 4);
        const float pp[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
        uint32_t o4[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = unpack_bf16(w4[e]);
          o4[e] = pack_bf16(fmaf(rs, v[2 * e] + pp[2 * e], f.x), fmaf(rs, v[2 * e + 1] + pp[2 * e + 1], f.y));
        }
        *reinterpret_cast<uint4*>(xr + k * 16) = make_uint4(o4[0], o4[1], o4[2], o4[3]);
      }
      if (warp == AT_WARP0) WM_TRACE(3, it, 3);
      tc::fence_before_sync();
      tc::fence_proxy_async();                                // the y tile is read by TMA stores
      __syncwarp();
      if (lane == 0) { tc::mbar_arrive(p_empty + pb); tc::mbar_arrive(y_full + slot); }
      if (warp == AT_WARP0) WM_TRACE(3, it, 4);
    };
    for (int it = 0; it < my_tiles; ++it) {
      const int tile = blockIdx.x + it * gridDim.x;
      if (warp == AT_WARP0) WM_TRACE(2, it, 0);
      tc::mbar_wait(qkv_full, it & 1);
      tc::fence_after_sync();
      if (warp == AT_WARP0) WM_TRACE(2, it, 1);
      // fragments of both units (windows 2q, 2q+1) up front, so the accumulators are released for the next tile's QKV
      uint32_t qf[2][8], kf[2][8], vf[2][8];
      float qb[8], kb_[8], vb[8];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f0 = *reinterpret_cast<const float2*>(s_qb + 8 * j);
        const float2 f1 = *reinterpret_cast<const float2*>(s_qb + WC + 8 * j);
        const float2 f2 = *reinterpret_cast<const float2*>(s_qb + 2 * WC + 8 * j);
        qb[2 * j] = f0.x; qb[2 * j + 1] = f0.y; kb_[2 * j] = f1.x; kb_[2 * j + 1] = f1.y; vb[2 * j] = f2.x; vb[2 * j + 1] = f2.y;
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32 + u * 16) << 16) + TM_QKV + head * 32;
        float vq[16], vk[16], vv[16];
        tmem_ld_frag32(lane_addr, vq);
        tmem_ld_frag32(lane_addr + WC, vk);
        tmem_ld_frag32(lane_addr + 2 * WC, vv);
        tmem_ld_wait();
        tmem_ld_use(vq); tmem_ld_use(vk); tmem_ld_use(vv);
        // A operand of S = q k^T (m16k16 per 16 dims): {a0,a1,a2,a3} = {(g, k lo), (g+8, k lo), (g, k hi), (g+8, k hi)}
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          qf[u][4 * ks + 0] = pack_bf16(vq[8 * ks + 0] + qb[4 * ks + 0], vq[8 * ks + 1] + qb[4 * ks + 1]);
          qf[u][4 * ks + 1] = pack_bf16(vq[8 * ks + 2] + qb[4 * ks + 0], vq[8 * ks + 3] + qb[4 * ks + 1]);
          qf[u][4 * ks + 2] = pack_bf16(vq[8 * ks + 4] + qb[4 * ks + 2], vq[8 * ks + 5] + qb[4 * ks + 3]);
          qf[u][4 * ks + 3] = pack_bf16(vq[8 * ks + 6] + qb[4 * ks + 2], vq[8 * ks + 7] + qb[4 * ks + 3]);
        }
        // k as B operand (k16 x n8, n = key): for key tile nt: b0 = k[key g + 8 nt][16 ks + 2t, +1], b1 = ... [16 ks + 8 + 2t, +1]
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          kf[u][4 * ks + 0] = pack_bf16(vk[8 * ks + 0] + kb_[4 * ks + 0], vk[8 * ks + 1] + kb_[4 * ks + 1]);   // nt 0, b0
          kf[u][4 * ks + 1] = pack_bf16(vk[8 * ks + 4] + kb_[4 * ks + 2], vk[8 * ks + 5] + kb_[4 * ks + 3]);   // nt 0, b1
          kf[u][4 * ks + 2] = pack_bf16(vk[8 * ks + 2] + kb_[4 * ks + 0], vk[8 * ks + 3] + kb_[4 * ks + 1]);   // nt 1, b0
          kf[u][4 * ks + 3] = pack_bf16(vk[8 * ks + 6] + kb_[4 * ks + 2], vk[8 * ks + 7] + kb_[4 * ks + 3]);   // nt 1, b1
        }
        // v in C layout (token rows); movmatrix turns each 8x8 block into the B operand of O = P v (k = token, n = dim)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          vf[u][2 * j + 0] = movm_trans(pack_bf16(vv[4 * j + 0] + vb[2 * j], vv[4 * j + 1] + vb[2 * j + 1]));   // tokens 0..7
          vf[u][2 * j + 1] = movm_trans(pack_bf16(vv[4 * j + 2] + vb[2 * j], vv[4 * j + 3] + vb[2 * j + 1]));   // tokens 8..15
        }
      }
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(qkv_empty);
      if (warp == AT_WARP0) WM_TRACE(2, it, 2);
      if (it > 0) epilogue(it - 1);                           // its projection was issued while the fragments above were loading
      if (warp == AT_WARP0) WM_TRACE(2, it, 6);

      // both units advance in lock step: two independent dependency chains per warp
      uint32_t maskbits[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        int b, wh, ww;
        decode_window(a, tile * 8 + 2 * q + u, b, wh, ww);
        maskbits[u] = 0;
        if (a.masked) {
          if (a.sh > 0 && wh == a.nWh - 1) maskbits[u] |= diffH;
          if (a.sw > 0 && ww == a.nWw - 1) maskbits[u] |= diffW;
        }
      }
      float s[2][2][4];
#pragma unroll
      for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
          for (int e = 0; e < 4; ++e) s[u][nt][e] = 0.f;
#pragma unroll
      for (int ks = 0; ks < 2; ++ks)
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const uint32_t af[4] = {qf[u][4 * ks], qf[u][4 * ks + 1], qf[u][4 * ks + 2], qf[u][4 * ks + 3]};
          mma_bf16_16816(s[u][0], af, kf[u][4 * ks + 0], kf[u][4 * ks + 1]);
          mma_bf16_16816(s[u][1], af, kf[u][4 * ks + 2], kf[u][4 * ks + 3]);
        }
#pragma unroll
      for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float v = fmaf(s[u][nt][e], scale_l2e, bias[nt][e]);          // scores in log2 units: exp(x) = 2^(x log2 e)
            if ((maskbits[u] >> (nt * 4 + e)) & 1u) v += -100.0f * 1.4426950408889634f;
            s[u][nt][e] = v;
          }
      if (warp == AT_WARP0) WM_TRACE(3, it, 5);
      // softmax over the 16 keys of each row (rows g and g + 8; a row lives in the 4 lanes of a quad): 4 rows in flight
      float mx[2][2], sm[2][2];
#pragma unroll
      for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int hf = 0; hf < 2; ++hf)
          mx[u][hf] = fmaxf(fmaxf(s[u][0][hf * 2], s[u][0][hf * 2 + 1]), fmaxf(s[u][1][hf * 2], s[u][1][hf * 2 + 1]));
#pragma unroll
      for (int o = 1; o <= 2; o <<= 1)
#pragma unroll
        for (int u = 0; u < 2; ++u)
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) mx[u][hf] = fmaxf(mx[u][hf], __shfl_xor_sync(0xffffffffu, mx[u][hf], o));
#pragma unroll
      for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          sm[u][hf] = 0.f;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float e = fast_ex2(s[u][k >> 1][hf * 2 + (k & 1)] - mx[u][hf]);
            s[u][k >> 1][hf * 2 + (k & 1)] = e;
            sm[u][hf] += e;
          }
        }
#pragma unroll
      for (int o = 1; o <= 2; o <<= 1)
#pragma unroll
        for (int u = 0; u < 2; ++u)
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) sm[u][hf] += __shfl_xor_sync(0xffffffffu, sm[u][hf], o);
      // P as a two-term bf16 split (hi + lo): its rounding would otherwise be the largest error of the attention core
      uint32_t pf[2][4], pl[2][4];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const float i0 = 1.0f / sm[u][0], i1 = 1.0f / sm[u][1];
        split_bf16(s[u][0][0] * i0, s[u][0][1] * i0, pf[u][0], pl[u][0]);
        split_bf16(s[u][0][2] * i1, s[u][0][3] * i1, pf[u][1], pl[u][1]);
        split_bf16(s[u][1][0] * i0, s[u][1][1] * i0, pf[u][2], pl[u][2]);
        split_bf16(s[u][1][2] * i1, s[u][1][3] * i1, pf[u][3], pl[u][3]);
      }
      if (warp == AT_WARP0) WM_TRACE(3, it, 6);
      float o[2][4][4];
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int u = 0; u < 2; ++u) {
#pragma unroll
          for (int e = 0; e < 4; ++e) o[u][j][e] = 0.f;
          mma_bf16_16816(o[u][j], pf[u], vf[u][2 * j], vf[u][2 * j + 1]);
          mma_bf16_16816(o[u][j], pl[u], vf[u][2 * j], vf[u][2 * j + 1]);
        }
      // O tile: rows 32q + 16u + g (+8), this head's K block, 16-byte chunk j, bytes 4t..4t+3
      if (warp == AT_WARP0) WM_TRACE(2, it, 3);
      if (it > 0) tc::mbar_wait(o_empty, (it - 1) & 1);        // the projection of the previous tile has read the O tile
      if (warp == AT_WARP0) WM_TRACE(2, it, 4);
#pragma unroll
      for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int r0 = q * 32 + u * 16 + g;
          *reinterpret_cast<uint32_t*>(sO + sw64_off(r0, j) + 4 * t) = pack_bf16(o[u][j][0], o[u][j][1]);
          *reinterpret_cast<uint32_t*>(sO + sw64_off(r0 + 8, j) + 4 * t) = pack_bf16(o[u][j][2], o[u][j][3]);
        }
      tc::fence_proxy_async();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(o_full);
      if (warp == AT_WARP0) WM_TRACE(2, it, 5);
    }
    if (my_tiles > 0) epilogue(my_tiles - 1);
  }
  tc::fence_before_sync();
  __syncthreads();
#ifdef TULIP_WMSA_TRACE
  if (a.trace && blockIdx.x == 0 && threadIdx.x == 0) a.trace[7 * 8 + 2] = clock64();      // all roles done
#endif
  if (warp == 0) tc::tmem_dealloc<512>(tmem_base);
}

long long* g_wmsa_trace = nullptr;

bool wmsa_disabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("TULIP_B200_NO_FUSED_WMSA");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

uint32_t div_magic(uint32_t d) { return d <= 1 ? 0u : (uint32_t)(((1ull << 32) + d - 1) / d); }

}  // namespace

bool wmsa_block_supported(int B, int H, int W, int C, int heads, int Mh, int Mw) {
  if (wmsa_disabled()) return false;
  if (C != WC || heads != WH || Mh * Mw != 16 || (Mw & 1) || H % Mh || W % Mw) return false;
  const long nwin = (long)B * (H / Mh) * (W / Mw);
  return nwin >= 1 && nwin * max(H / Mh, W / Mw) < (1l << 31);          // fast_div range
}

int wmsa_block_fwd(const WmsaBlockArgs& w, cudaStream_t st) {
  TULIP_REQUIRE(wmsa_block_supported(w.B, w.H, w.W, w.C, w.heads, w.Mh, w.Mw),
                "fused W-MSA block: needs C = 96 (3 heads of 32) and 16-token windows that tile the grid");
  TULIP_REQUIRE(w.bMh * w.bMw == 16, "fused W-MSA block: bias window must hold 16 tokens");
  TULIP_REQUIRE(!(reinterpret_cast<uintptr_t>(w.x) & 15) && !(reinterpret_cast<uintptr_t>(w.y) & 15) &&
                !(reinterpret_cast<uintptr_t>(w.wqkv) & 15) && !(reinterpret_cast<uintptr_t>(w.wproj) & 15),
                "fused W-MSA block: 16-byte aligned activations and weights");
  TULIP_REQUIRE((w.sh == 0 || 2 * w.sh == w.Mh) && (w.sw == 0 || 2 * w.sw == w.Mw),
                "fused W-MSA block: the cyclic shift must be half a window (tulip.py:220)");
  CUtensorMap mq, mp, mwin, mseg, mwiny, msegy;
  {
    const uint64_t dims[2] = {(uint64_t)WC, (uint64_t)(3 * WC)};
    const uint64_t str[1] = {(uint64_t)WC * 2};
    const uint32_t box[2] = {KBLK, WC};
    int rc = tulip_make_tmap(&mq, w.wqkv, 2, dims, str, box, 64);
    if (rc) return rc;
    const uint64_t dimsp[2] = {(uint64_t)WC, (uint64_t)WC};
    rc = tulip_make_tmap(&mp, w.wproj, 2, dimsp, str, box, 64);
    if (rc) return rc;
    // x as (c, w, h, b): a window is one box {C, Mw, Mh, 1}; a half window row {C, Mw/2, 1, 1} where the shift seam cuts it
    const uint64_t dx[4] = {(uint64_t)WC, (uint64_t)w.W, (uint64_t)w.H, (uint64_t)w.B};
    const uint64_t sx[3] = {(uint64_t)WC * 2, (uint64_t)w.W * WC * 2, (uint64_t)w.H * w.W * WC * 2};
    const uint32_t bwin[4] = {WC, (uint32_t)w.Mw, (uint32_t)w.Mh, 1};
    const uint32_t bseg[4] = {WC, (uint32_t)(w.Mw / 2), 1, 1};
    rc = tulip_make_tmap(&mwin, w.x, 4, dx, sx, bwin, 0);
    if (rc) return rc;
    rc = tulip_make_tmap(&mseg, w.x, 4, dx, sx, bseg, 0);
    if (rc) return rc;
    rc = tulip_make_tmap(&mwiny, w.y, 4, dx, sx, bwin, 0);
    if (rc) return rc;
    rc = tulip_make_tmap(&msegy, w.y, 4, dx, sx, bseg, 0);
    if (rc) return rc;
  }
  WmsaArgs a;
  memset(&a, 0, sizeof a);
  a.ln_w = w.ln_w; a.ln_b = w.ln_b; a.bqkv = w.bqkv; a.bproj = w.bproj; a.bias_table = w.bias_table;
  a.row_scale = w.row_scale;
  a.B = w.B; a.H = w.H; a.W = w.W; a.Mh = w.Mh; a.Mw = w.Mw; a.sh = w.sh; a.sw = w.sw; a.masked = w.masked; a.bMh = w.bMh; a.bMw = w.bMw;
  a.nWh = w.H / w.Mh; a.nWw = w.W / w.Mw; a.mul_nWw = div_magic(a.nWw); a.mul_nWh = div_magic(a.nWh);
  a.eps = w.eps; a.scale = 1.0f / sqrtf(32.0f);
  a.nwin = w.B * a.nWh * a.nWw;
  a.ntiles = (a.nwin + 7) / 8;
  a.trace = g_wmsa_trace;
  static bool configured = false;
  if (!configured) {
    TULIP_CUDA(cudaFuncSetAttribute(wmsa_block_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WM_SMEM));
    configured = true;
  }
  const int grid = min(a.ntiles, tulip_num_sms());
  tulip_launch(wmsa_block_fwd_kernel, grid, WM_THREADS, WM_SMEM, st, mq, mp, mwin, mseg, mwiny, msegy, a);
  TULIP_CHECK_LAUNCH();
  return TULIP_OK;
}

// bring-up hook (not part of the public header): clock64() stamps of CTA 0, [3 roles][8 tiles][8 points] long long
extern "C" void tulip_debug_wmsa_trace(long long* device_buf) { g_wmsa_trace = device_buf; }
