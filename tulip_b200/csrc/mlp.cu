// Fused MLP half-block, forward:   y = x + s_b * fc2(gelu(fc1(LN2(x))))
//
// Reference: tulip/model/tulip.py:347-352 (SwinTransformerBlock.forward, MLP half) with Mlp.forward (:194-200) inlined:
// LayerNorm(eps 1e-6) -> fc1 (C -> 4C) + bias -> exact-erf GELU -> fc2 (4C -> C) + bias -> DropPath scale -> residual.
// ONE launch; the [T, 4C] hidden tensor never reaches HBM in inference.  In training the same launch also emits what the
// backward pass consumes -- the LayerNorm output, its (mean, rstd) and the activated hidden tensor -- straight from the
// operand tiles it has in shared memory (TMA stores), which replaces LayerNorm + GEMM + GEMM (26 T C bytes) by 14 T C bytes.
//
// A persistent CTA walks 128-token tiles (tokens are contiguous rows: no window geometry here).  The hidden dimension is
// processed in four chunks of 96 units: fc1 chunk -> TMEM -> GELU warps -> bf16 H tile in shared memory -> fc2 partial
// product accumulating in TMEM.  The two weight matrices (2 x 72 KB at C = 96) do not fit next to the tiles, so their
// eight [96 x 96] chunks stream through a 4-slot TMA ring from L2 in exactly the order the MMAs consume them:
//     f1(0), f1(1), f2(0), f1(2), f2(1), ...        (f2 of a chunk is issued one fc1 later: its GELU runs meanwhile)
// Five warpgroups (registers re-balanced with setmaxnreg):
//   warp 0        tcgen05 issuer        warp 1  x-tile loader (TMA)     warp 2  storer (y tile, by-products)
//   warp 3        weight-chunk loader (TMA ring)
//   warps 4..7    LayerNorm: one thread per row, raw row -> K-major 64B-swizzled A tile (+ statistics in training)
//   warps 8..19   GELU + epilogue: warp (q, j) owns rows 32q.. and columns 32j.. of every 96-column chunk / of the output
//
// MODE 1 of the same kernel is the HEAD BACKWARD (PixelShuffleHead + decoder_pred + L1, tulip.py:161-178, 727-731, 692-693, at
// embed_dim 96): the head is an MLP with 16 "hidden" chunks -- chunk ij = the 96 channels of shuffle slot ij.  fc1 chunk =
// recomputed pre-activation (xn_up . We'[96 ij.., :]^T), the GELU step becomes dh = dpred[m, ij] * wd[c] * LeakyReLU'(pre)
// (+ the column sums for d(decoder_pred.weight)), fc2 accumulates d(xn_up) += dh_chunk . We'[chunk] over the 16 chunks, and the
// dh chunks leave through the storer for the weight-gradient GEMM.  It replaces the recompute GEMM (epilogue 7) AND the dX GEMM
// that re-read the 402 MB dh tensor.  The LayerNorm warps only copy the (already normalised) input rows into the A operand.
#include "fused.cuh"
#include "kernels.h"

#include <cstdlib>
#include <cstring>

namespace {
using namespace fused;

constexpr int MC = 96;                      // channels
constexpr int MAX_NCH = 16;                 // hidden chunks per tile: 4 (MLP, hidden 4C) or 16 (head backward, r^2 = 16 shuffle slots)
constexpr int ML_WARPS = 20, ML_THREADS = 32 * ML_WARPS;
constexpr int EP_WARPS = 12, LN_WARPS = 4, LN_WARP0 = 4, EP_WARP0 = 8;
constexpr int REGS_ISSUER = 24, REGS_LN = 72, REGS_EP = 128;
constexpr int KBLK = 32, NKB = MC / KBLK;
constexpr int A_BLK = 128 * 64, W_BLK = 96 * 64;
constexpr int ROWB = MC * 2, RAW_TILE = 128 * ROWB, RAW_RING = 3;
constexpr int W_SLOT = NKB * W_BLK, W_RING = 4;
constexpr int OFF_WR = 0;                                   // weight ring: 4 x [3 K blocks][96 x 32]
constexpr int OFF_A = OFF_WR + W_RING * W_SLOT;             // LayerNorm output, A operand of fc1
constexpr int OFF_H = OFF_A + NKB * A_BLK;                  // 2 x activated hidden chunk, A operand of fc2
constexpr int OFF_RAW = OFF_H + 2 * NKB * A_BLK;            // ring of raw x tiles: LayerNorm source, residual, then the y tile
constexpr int OFF_PAR = OFF_RAW + RAW_RING * RAW_TILE;      // fp32: b1 [96 NCH] | b2 [96] (head: wd) | gamma [96] | beta [96]
template <int NCH_>
struct Lay {
  static constexpr int NCH = NCH_, MHID = NCH_ * MC;
  static constexpr int OFF_BAR = OFF_PAR + (MHID + 3 * MC) * 4;
  static constexpr int SMEM = OFF_BAR + 512 + 1024;
  static_assert(SMEM <= 227 * 1024, "shared memory budget");
};
constexpr int TM_ACC1 = 0, TM_ACC2 = 2 * MC;                // TMEM columns: two fc1 chunk buffers, two fc2 tile buffers

struct MlpArgs {
  const float* ln_w; const float* ln_b; const float* b1; const float* b2;
  const float* row_scale; int rows_per_sample;
  float* stats;                             // [T, 2] (mean, rstd) or null
  int T, ntiles;
  int save;                                 // training: LayerNorm output and activated hidden tensor are stored too
  float eps;
  // MODE 1 (head backward): b1 = permuted conv_expand bias [1536], b2 = decoder_pred.weight [96]; dpred is formed from pred / target
  const float* pred; const float* target; const float* gscale; float inv_npix;
  float* dwd; int dwd_copies;               // CTA b adds its column sums into copy b % dwd_copies (96 floats apart)
  int hd_H, hd_W, hd_r;
  int hints;                                // tulip_hints(): 16 input tile + by-product stores evict_first, 64 weight chunks evict_last
};

template <int MODE>
__global__ void __launch_bounds__(ML_THREADS, 1)
mlp_block_fwd_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapY,
                     const __grid_constant__ CUtensorMap mapW1, const __grid_constant__ CUtensorMap mapW2,
                     const __grid_constant__ CUtensorMap mapXn, const __grid_constant__ CUtensorMap mapHact,
                     const __grid_constant__ MlpArgs a) {
  using LY = Lay<MODE == 0 ? 4 : MAX_NCH>;
  constexpr int NCH = LY::NCH, MHID = LY::MHID;
  extern __shared__ unsigned char smem_raw[];
  pdl_trigger();
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  float* spar = reinterpret_cast<float*>(smem + OFF_PAR);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + LY::OFF_BAR);
  uint64_t* w_full = bars;              // [4] weight chunk landed
  uint64_t* w_empty = bars + 4;         // [4] its MMAs have read it
  uint64_t* raw_full = bars + 8;        // [3]
  uint64_t* raw_empty = bars + 11;      // [3] y tile stored
  uint64_t* y_full = bars + 14;         // [3] epilogue wrote the y tile in place
  uint64_t* a_full = bars + 17;         // A tile written
  uint64_t* a_empty = bars + 18;        // fc1 of the tile has read it (and, in training, the storer)
  uint64_t* acc1_full = bars + 19;      // [2] fc1 chunk accumulator complete
  uint64_t* acc1_empty = bars + 21;     // [2] drained by the GELU warps
  uint64_t* h_full = bars + 23;         // [2] H chunk written
  uint64_t* h_empty = bars + 25;        // [2] fc2 partial product has read it (and, in training, the storer)
  uint64_t* acc2_full = bars + 27;      // [2] fc2 tile accumulator complete
  uint64_t* acc2_empty = bars + 29;     // [2] drained by the epilogue
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 31);

  const int warp = tc::warp_idx_sync(), lane = threadIdx.x & 31;
  const bool save_xn = (MODE == 0) && a.save, save_h = a.save != 0;   // head backward: only the dh chunks are stored
  const int a_readers = save_xn ? 2 : 1, h_readers = save_h ? 2 : 1;  // MMA commit (+ storer)
  if (threadIdx.x == 0) {
    for (int b = 0; b < W_RING; ++b) { tc::mbar_init(w_full + b, 1); tc::mbar_init(w_empty + b, 1); }
    for (int b = 0; b < RAW_RING; ++b) { tc::mbar_init(raw_full + b, 1); tc::mbar_init(raw_empty + b, 1); tc::mbar_init(y_full + b, EP_WARPS); }
    tc::mbar_init(a_full, LN_WARPS); tc::mbar_init(a_empty, a_readers);
    for (int b = 0; b < 2; ++b) {
      tc::mbar_init(acc1_full + b, 1); tc::mbar_init(acc1_empty + b, EP_WARPS);
      tc::mbar_init(h_full + b, EP_WARPS); tc::mbar_init(h_empty + b, h_readers);
      tc::mbar_init(acc2_full + b, 1); tc::mbar_init(acc2_empty + b, EP_WARPS);
    }
    tc::fence_barrier_init();
  }
  if (warp == 0) tc::tmem_alloc<512>(tmem_slot);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const int my_tiles = (a.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int NG = NCH * my_tiles;                              // hidden chunks this CTA processes, in order

  if (warp < LN_WARP0) reg_dec<REGS_ISSUER>();
  if (warp == 0) {
    // ------------------------------------------------------------------ tcgen05 issuer
    constexpr uint32_t idesc = tc::make_idesc(128, MC, 0, 0);
    const uint64_t dW0 = desc_k_sw64(smem + OFF_WR), dA = desc_k_sw64(smem + OFF_A), dH0 = desc_k_sw64(smem + OFF_H);
    int wn = 0;                                               // weight chunks consumed so far (ring position)
    auto issue = [&](uint32_t tmem_d, uint64_t dOp, bool accumulate) {
      const int slot = wn % W_RING;
      tc::mbar_wait(w_full + slot, (wn / W_RING) & 1);
      tc::fence_after_sync();
      if (tc::elect_one_sync()) {
        const uint64_t dW = dW0 + (uint64_t)((slot * W_SLOT) >> 4);
#pragma unroll
        for (int kb = 0; kb < NKB; ++kb)
#pragma unroll
          for (int ks = 0; ks < 2; ++ks)
            tc::umma_bf16(tmem_d, dOp + (uint64_t)((kb * A_BLK) >> 4) + 2 * ks, dW + (uint64_t)((kb * W_BLK) >> 4) + 2 * ks, idesc,
                          (accumulate || kb || ks) ? 1u : 0u);
        tc::umma_commit(w_empty + slot);
      }
      __syncwarp();
      ++wn;
    };
    auto f1 = [&](int G) {
      const int it = G / NCH, c = G - it * NCH, buf = G & 1;
      if (c == 0) tc::mbar_wait(a_full, it & 1);
      tc::mbar_wait(acc1_empty + buf, ((G >> 1) & 1) ^ 1);
      issue(tmem_base + TM_ACC1 + buf * MC, dA, false);
      if (tc::elect_one_sync()) {
        tc::umma_commit(acc1_full + buf);
        if (c == NCH - 1) tc::umma_commit(a_empty);
      }
      __syncwarp();
    };
    auto f2 = [&](int G) {
      const int it = G / NCH, c = G - it * NCH, buf = G & 1, tb = it & 1;
      tc::mbar_wait(h_full + buf, (G >> 1) & 1);
      if (c == 0) tc::mbar_wait(acc2_empty + tb, ((it >> 1) & 1) ^ 1);
      issue(tmem_base + TM_ACC2 + tb * MC, dH0 + (uint64_t)((buf * NKB * A_BLK) >> 4), c > 0);
      if (tc::elect_one_sync()) {
        tc::umma_commit(h_empty + buf);
        if (c == NCH - 1) tc::umma_commit(acc2_full + tb);
      }
      __syncwarp();
    };
    for (int G = 0; G <= NG; ++G) {
      if (G < NG) f1(G);
      if (G >= 1) f2(G - 1);
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ x-tile loader
    if (tc::elect_one_sync()) tc::prefetch_tensormap(&mapX);
    pdl_wait();                                               // x is produced by the preceding kernel
    for (int it = 0; it < my_tiles; ++it) {
      const int tile = blockIdx.x + it * gridDim.x, slot = it % RAW_RING;
      if (it >= RAW_RING) tc::mbar_wait(raw_empty + slot, ((it / RAW_RING) - 1) & 1);
      if (tc::elect_one_sync()) {
        tc::mbar_expect_tx(raw_full + slot, RAW_TILE);        // rows past T are zero-filled by the TMA unit and still count
        tc::tma_load_2d_pol(smem + OFF_RAW + slot * RAW_TILE, &mapX, raw_full + slot, 0, tile * 128, tc::l2_policy_evict_first(),
                            (a.hints & 16) != 0);
      }
      __syncwarp();
    }
  } else if (warp == 3) {
    // ------------------------------------------------------------------ weight-chunk loader (order = MMA order)
    if (tc::elect_one_sync()) { tc::prefetch_tensormap(&mapW1); tc::prefetch_tensormap(&mapW2); }
    int wn = 0;
    auto load = [&](const CUtensorMap* map, int col0, int row0) {
      const int slot = wn % W_RING;
      if (wn >= W_RING) tc::mbar_wait(w_empty + slot, ((wn / W_RING) - 1) & 1);
      if (tc::elect_one_sync()) {
        tc::mbar_expect_tx(w_full + slot, W_SLOT);
        for (int kb = 0; kb < NKB; ++kb)
          tc::tma_load_2d_pol(smem + OFF_WR + slot * W_SLOT + kb * W_BLK, map, w_full + slot, col0 + kb * KBLK, row0,
                              tc::l2_policy_evict_last(), (a.hints & 64) != 0);
      }
      __syncwarp();
      ++wn;
    };
    for (int G = 0; G <= NG; ++G) {
      if (G < NG) load(&mapW1, 0, (G % NCH) * MC);                           // fc1: rows 96c.. of W1 [4C, C]
      if (G >= 1) load(&mapW2, ((G - 1) % NCH) * MC, 0);                     // fc2: columns 96c.. of W2 [C, 4C]
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------------ storer: by-products (training) and the y tile
    if (tc::elect_one_sync()) { tc::prefetch_tensormap(&mapY); tc::prefetch_tensormap(&mapXn); tc::prefetch_tensormap(&mapHact); }
    for (int it = 0; it < my_tiles; ++it) {
      const int tile = blockIdx.x + it * gridDim.x;
      if (save_xn) {
        tc::mbar_wait(a_full, it & 1);
        if (tc::elect_one_sync()) {
          for (int kb = 0; kb < NKB; ++kb)
            tc::tma_store_2d_pol(&mapXn, smem + OFF_A + kb * A_BLK, kb * KBLK, tile * 128, tc::l2_policy_evict_first(), (a.hints & 16) != 0);
          tc::tma_store_commit();
          tc::tma_store_wait_read<0>();
          tc::mbar_arrive(a_empty);
        }
        __syncwarp();
      }
      if (save_h) {
        for (int c = 0; c < NCH; ++c) {
          const int G = it * NCH + c, buf = G & 1;
          tc::mbar_wait(h_full + buf, (G >> 1) & 1);
          if (tc::elect_one_sync()) {
            for (int kb = 0; kb < NKB; ++kb)
              tc::tma_store_2d_pol(&mapHact, smem + OFF_H + buf * NKB * A_BLK + kb * A_BLK, c * MC + kb * KBLK, tile * 128,
                                   tc::l2_policy_evict_first(), (a.hints & 16) != 0);
            tc::tma_store_commit();
            tc::tma_store_wait_read<0>();
            tc::mbar_arrive(h_empty + buf);
          }
          __syncwarp();
        }
      }
      if (it > 0) {
        const int j = it - 1, slot = j % RAW_RING;
        tc::mbar_wait(y_full + slot, (j / RAW_RING) & 1);
        if (tc::elect_one_sync()) {
          tc::tma_store_2d(&mapY, smem + OFF_RAW + slot * RAW_TILE, 0, (blockIdx.x + j * gridDim.x) * 128);
          tc::tma_store_commit();
          tc::tma_store_wait_read<0>();
          tc::mbar_arrive(raw_empty + slot);
        }
        __syncwarp();
      }
    }
    if (my_tiles > 0) {
      const int j = my_tiles - 1, slot = j % RAW_RING;
      tc::mbar_wait(y_full + slot, (j / RAW_RING) & 1);
      if (tc::elect_one_sync()) {
        tc::tma_store_2d(&mapY, smem + OFF_RAW + slot * RAW_TILE, 0, (blockIdx.x + j * gridDim.x) * 128);
        tc::tma_store_commit();
      }
      __syncwarp();
    }
    if (tc::elect_one_sync()) tc::tma_store_wait<0>();
    __syncwarp();
  } else if (warp >= LN_WARP0 && warp < EP_WARP0) {
    // ------------------------------------------------------------------ LayerNorm warps
    reg_dec<REGS_LN>();
    const int q = warp & 3;
    const int r = q * 32 + lane;
    for (int i = threadIdx.x - 32 * LN_WARP0; i < MHID + (MODE == 0 ? 3 : 1) * MC; i += 32 * LN_WARPS) {
      float v;
      if (i < MHID) v = a.b1[i];
      else if (i < MHID + MC) v = a.b2[i - MHID];
      else if (i < MHID + 2 * MC) v = a.ln_w[i - MHID - MC];
      else v = a.ln_b[i - MHID - 2 * MC];
      spar[i] = v;
    }
    tc::named_bar_sync(1, 32 * (LN_WARPS + EP_WARPS));
    const float* s_g = spar + MHID + MC;
    const float* s_b = spar + MHID + 2 * MC;
    const int rot = (r >> 1) % 12;
    for (int it = 0; it < my_tiles; ++it) {
      const int tile = blockIdx.x + it * gridDim.x, slot = it % RAW_RING;
      const unsigned char* src = smem + OFF_RAW + slot * RAW_TILE + r * ROWB;
      unsigned char* dst = smem + OFF_A;
      tc::mbar_wait(raw_full + slot, (it / RAW_RING) & 1);
      if (MODE == 1) {
        // head backward: the rows are LayerNorm output already; they only move into the swizzled A operand
        if (it >= 1) tc::mbar_wait(a_empty, (it - 1) & 1);
        int kk = rot;
#pragma unroll 4
        for (int c = 0; c < MC / 8; ++c) {
          const uint4 v = *reinterpret_cast<const uint4*>(src + kk * 16);
          *reinterpret_cast<uint4*>(dst + (kk >> 2) * A_BLK + sw64_off(r, kk & 3)) = v;
          kk = (kk == 11) ? 0 : kk + 1;
        }
        tc::fence_proxy_async();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(a_full);
        continue;
      }
      const float x0 = unpack_bf16(*reinterpret_cast<const uint32_t*>(src)).x;
      float s1 = 0.f, s2 = 0.f;
      int k = rot;
#pragma unroll 4
      for (int c = 0; c < MC / 8; ++c) {
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
      const float dm = s1 * (1.0f / MC);
      const float mean = x0 + dm;
      const float rstd = rsqrtf(fmaxf(s2 * (1.0f / MC) - dm * dm, 0.f) + a.eps);
      if (a.stats && tile * 128 + r < a.T) *reinterpret_cast<float2*>(a.stats + 2 * (long)(tile * 128 + r)) = make_float2(mean, rstd);
      if (it >= 1) tc::mbar_wait(a_empty, (it - 1) & 1);       // fc1 of the previous tile (and the storer) have read the A tile
      asm volatile("" ::: "memory");
      k = rot;
#pragma unroll 2
      for (int c = 0; c < MC / 8; ++c) {
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
    }
  } else if (warp >= EP_WARP0) {
    // ------------------------------------------------------------------ GELU + epilogue warps
    reg_inc<REGS_EP>();
    const int q = warp & 3, jg = (warp - EP_WARP0) >> 2;      // rows 32q + lane, columns 32 jg .. of a chunk / of the output
    const int r = q * 32 + lane;
    tc::named_bar_sync(1, 32 * (LN_WARPS + EP_WARPS));
    const float* s_b2 = spar + MHID + jg * 32;
    auto epilogue = [&](int it) {
      const int tile = blockIdx.x + it * gridDim.x, tb = it & 1, slot = it % RAW_RING;
      float rs = 1.0f;
      if (a.row_scale) rs = a.row_scale[min(tile * 128 + r, a.T - 1) / a.rows_per_sample];
      unsigned char* xr = smem + OFF_RAW + slot * RAW_TILE + r * ROWB + jg * 64;      // residual in, y out (in place)
      tc::mbar_wait(acc2_full + tb, (it >> 1) & 1);
      tc::fence_after_sync();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + TM_ACC2 + tb * MC + jg * 32;
      const int rot = (lane >> 1) & 3;
      float v0[8], v1[8], v2[8], v3[8];
      tmem_ld_row8(taddr, v0);
      tmem_ld_row8(taddr + 8, v1);
      tmem_ld_row8(taddr + 16, v2);
      tmem_ld_row8(taddr + 24, v3);
      tmem_ld_wait();
      tmem_ld_use8(v0); tmem_ld_use8(v1); tmem_ld_use8(v2); tmem_ld_use8(v3);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int k = (c + rot) & 3;
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = sel4(k, v0[e], v1[e], v2[e], v3[e]);
        const uint4 xx = *reinterpret_cast<const uint4*>(xr + k * 16);
        const uint32_t w4[4] = {xx.x, xx.y, xx.z, xx.w};
        const float4 p0 = *reinterpret_cast<const float4*>(s_b2 + k * 8), p1 = *reinterpret_cast<const float4*>(s_b2 + k * 8 + 4);
        const float pp[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
        uint32_t o4[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = unpack_bf16(w4[e]);
          if (MODE == 1) o4[e] = pack_bf16(v[2 * e], v[2 * e + 1]);             // d(xn_up): no bias, no residual
          else o4[e] = pack_bf16(fmaf(rs, v[2 * e] + pp[2 * e], f.x), fmaf(rs, v[2 * e + 1] + pp[2 * e + 1], f.y));
        }
        *reinterpret_cast<uint4*>(xr + k * 16) = make_uint4(o4[0], o4[1], o4[2], o4[3]);
      }
      tc::fence_before_sync();
      tc::fence_proxy_async();
      __syncwarp();
      if (lane == 0) { tc::mbar_arrive(acc2_empty + tb); tc::mbar_arrive(y_full + slot); }
    };
    // head backward: dpred of (row, shuffle slot) = sign(pred - target) * gscale / npix, fetched one chunk ahead;
    // cw[i] = this thread's part of the column sums sum_rows dpred * LeakyReLU(pre) for d(decoder_pred.weight)
    float cw[32];
    float hd_gs = 0.f, pn = 0.f, tn = 0.f;
    long hd_base = -1;                                       // pixel (b, 4h, 4w) of this thread's row in the current tile, -1: row past T
    const int hd_Wr = a.hd_W * 4;                            // r = 4 in this mode (16 shuffle slots)
    auto hd_fetch = [&](int G_) {                            // pred / target of this thread's row for chunk G_ (slot ij = G_ mod 16)
      const int ij = G_ & (NCH - 1);
      if (ij == 0) {                                         // first slot of a tile: the row's base pixel (the only divisions)
        const int m = (blockIdx.x + (G_ / NCH) * gridDim.x) * 128 + r;
        hd_base = -1;
        if (m < a.T) {
          const int w = m % a.hd_W, bh = m / a.hd_W, hh = bh % a.hd_H, b_ = bh / a.hd_H;
          hd_base = (long)(b_ * a.hd_H + hh) * 4 * hd_Wr + w * 4;
        }
      }
      pn = tn = 0.f;
      if (hd_base >= 0) {
        const long px = hd_base + (ij >> 2) * hd_Wr + (ij & 3);
        pn = a.pred[px]; tn = a.target[px];
      }
    };
    if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < 32; ++i) cw[i] = 0.f;
      hd_gs = a.gscale[0] * a.inv_npix;
      if (NG > 0) hd_fetch(0);
    }
    for (int G = 0; G < NG; ++G) {
      const int it = G / NCH, c = G - it * NCH, buf = G & 1;
      float dp = 0.f;
      if (MODE == 1) {
        const float d = pn - tn;
        dp = d > 0.f ? hd_gs : (d < 0.f ? -hd_gs : 0.f);
        if (G + 1 < NG) hd_fetch(G + 1);
      }
      tc::mbar_wait(acc1_full + buf, (G >> 1) & 1);
      tc::fence_after_sync();
      const float* bb = spar + c * MC + jg * 32;
      uint32_t h[16];
      if constexpr (MODE == 1) {
        // two 16-column halves: next to the 32 column-sum accumulators a whole 32-column load does not fit the register budget
        const float* wdp = spar + MHID + jg * 32;             // decoder_pred.weight sits in the b2 slot
        const float dp01 = 0.01f * dp;
        const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + TM_ACC1 + buf * MC + jg * 32;
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          float v16[16];
          tc::tmem_ld16(ta + 16 * hf, v16);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 b4 = *reinterpret_cast<const float4*>(bb + 16 * hf + 4 * i);
            const float4 w4 = *reinterpret_cast<const float4*>(wdp + 16 * hf + 4 * i);
            const float pre[4] = {v16[4 * i] + b4.x, v16[4 * i + 1] + b4.y, v16[4 * i + 2] + b4.z, v16[4 * i + 3] + b4.w};
            const float ww[4] = {w4.x, w4.y, w4.z, w4.w};
            float o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float t = pre[e] > 0.f ? dp : dp01;       // dpred * LeakyReLU'(pre)
              cw[16 * hf + 4 * i + e] = fmaf(t, pre[e], cw[16 * hf + 4 * i + e]);   // dpred * LeakyReLU(pre)
              o[e] = t * ww[e];
            }
            h[8 * hf + 2 * i] = pack_bf16(o[0], o[1]);
            h[8 * hf + 2 * i + 1] = pack_bf16(o[2], o[3]);
          }
        }
        tc::fence_before_sync();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(acc1_empty + buf);     // the next-but-one fc1 chunk may overwrite the accumulator
      } else {
      float v[32];
      tc::tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + TM_ACC1 + buf * MC + jg * 32, v);
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(acc1_empty + buf);       // the next-but-one fc1 chunk may overwrite the accumulator
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 b4 = *reinterpret_cast<const float4*>(bb + 4 * i);
        h[2 * i] = pack_bf16(gelu_erf(v[4 * i] + b4.x), gelu_erf(v[4 * i + 1] + b4.y));
        h[2 * i + 1] = pack_bf16(gelu_erf(v[4 * i + 2] + b4.z), gelu_erf(v[4 * i + 3] + b4.w));
      }
      }
      tc::mbar_wait(h_empty + buf, ((G >> 1) & 1) ^ 1);        // fc2 of the chunk two back (and the storer) have read this H buffer
      unsigned char* hd = smem + OFF_H + buf * NKB * A_BLK + jg * A_BLK;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        *reinterpret_cast<uint4*>(hd + sw64_off(r, j)) = make_uint4(h[4 * j], h[4 * j + 1], h[4 * j + 2], h[4 * j + 3]);
      tc::fence_proxy_async();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(h_full + buf);
      if (c == 1 && it > 0) epilogue(it - 1);                 // its last fc2 partial product was issued two chunks ago
    }
    if (my_tiles > 0) epilogue(my_tiles - 1);
    if constexpr (MODE == 1) {
      // column sums over the 32 rows of the warp (fixed tree), one atomic per column into this CTA's gradient copy
      float* dwd = a.dwd + (a.dwd_copies > 1 ? (int)(blockIdx.x % a.dwd_copies) * MC : 0);
      const float sum = warp_colsum32(cw, lane);
      atomicAdd(dwd + jg * 32 + lane, sum);
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<512>(tmem_base);
}

bool mlp_disabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("TULIP_B200_NO_FUSED_MLP");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

}  // namespace

bool mlp_block_supported(int T, int C) { return !mlp_disabled() && C == MC && T >= 1; }

int mlp_block_fwd(const MlpBlockArgs& m, cudaStream_t st) {
  TULIP_REQUIRE(mlp_block_supported(m.T, m.C), "fused MLP block: needs C = 96 (hidden 384)");
  auto aligned = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  TULIP_REQUIRE(aligned(m.x) && aligned(m.y) && aligned(m.w1) && aligned(m.w2) && aligned(m.xn) && aligned(m.hact),
                "fused MLP block: 16-byte aligned activations and weights");
  const bool save = m.xn != nullptr;
  TULIP_REQUIRE(!save || (m.hact != nullptr && m.stats != nullptr), "fused MLP block: training mode needs xn, hact and stats together");
  CUtensorMap mx, my, mw1, mw2, mxn, mh;
  {
    const uint64_t dx[2] = {(uint64_t)MC, (uint64_t)m.T};
    const uint64_t sx[1] = {(uint64_t)MC * 2};
    const uint32_t brow[2] = {MC, 128};                     // a whole 128-row tile, linear (rows are contiguous tokens)
    int rc = tulip_make_tmap(&mx, m.x, 2, dx, sx, brow, 0);
    if (rc) return rc;
    rc = tulip_make_tmap(&my, m.y, 2, dx, sx, brow, 0);
    if (rc) return rc;
    const uint64_t d1[2] = {(uint64_t)MC, (uint64_t)Lay<4>::MHID};
    const uint32_t bw[2] = {KBLK, MC};
    rc = tulip_make_tmap(&mw1, m.w1, 2, d1, sx, bw, 64);
    if (rc) return rc;
    const uint64_t d2[2] = {(uint64_t)Lay<4>::MHID, (uint64_t)MC};
    const uint64_t s2[1] = {(uint64_t)Lay<4>::MHID * 2};
    rc = tulip_make_tmap(&mw2, m.w2, 2, d2, s2, bw, 64);
    if (rc) return rc;
    mxn = mx; mh = mx;
    if (save) {
      const uint32_t bk[2] = {KBLK, 128};                   // one 64B-swizzled K block of an operand tile
      rc = tulip_make_tmap(&mxn, m.xn, 2, dx, sx, bk, 64);
      if (rc) return rc;
      const uint64_t dh[2] = {(uint64_t)Lay<4>::MHID, (uint64_t)m.T};
      rc = tulip_make_tmap(&mh, m.hact, 2, dh, s2, bk, 64);
      if (rc) return rc;
    }
  }
  MlpArgs a;
  memset(&a, 0, sizeof a);
  a.ln_w = m.ln_w; a.ln_b = m.ln_b; a.b1 = m.b1; a.b2 = m.b2; a.row_scale = m.row_scale;
  a.rows_per_sample = m.rows_per_sample > 0 ? m.rows_per_sample : 1;
  a.stats = save ? m.stats : nullptr;
  a.T = m.T; a.ntiles = (m.T + 127) / 128; a.save = save ? 1 : 0; a.eps = m.eps;
  a.hints = tulip_hints();
  static bool configured = false;
  if (!configured) {
    TULIP_CUDA(cudaFuncSetAttribute(mlp_block_fwd_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, Lay<4>::SMEM));
    configured = true;
  }
  const int grid = min(a.ntiles, tulip_num_sms());
  tulip_launch(mlp_block_fwd_kernel<0>, grid, ML_THREADS, Lay<4>::SMEM, st, mx, my, mw1, mw2, mxn, mh, a);
  TULIP_CHECK_LAUNCH();
  return TULIP_OK;
}

bool head_bwd_fused_supported(int E, int r) {
  static int off = -1;
  if (off < 0) {
    const char* e = getenv("TULIP_B200_NO_FUSED_HEAD_BWD");
    off = (e && e[0] == '1') ? 1 : 0;
  }
  return !off && E == MC && r * r == MAX_NCH;
}

int head_bwd_fused(const HeadBwdArgs& m, cudaStream_t st) {
  TULIP_REQUIRE(head_bwd_fused_supported(m.E, m.r) && m.T >= 1, "fused head backward: needs embed_dim 96 and upscale factor 4");
  auto aligned = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  TULIP_REQUIRE(aligned(m.xn) && aligned(m.dxn) && aligned(m.we) && aligned(m.wet) && aligned(m.dh),
                "fused head backward: 16-byte aligned activations and weights");
  TULIP_REQUIRE(m.bias && m.wd && m.pred && m.target && m.gscale && m.dwd && m.dh, "fused head backward: null argument");
  constexpr int HID = MAX_NCH * MC;
  CUtensorMap mx, my, mw1, mw2, mh;
  {
    const uint64_t dx[2] = {(uint64_t)MC, (uint64_t)m.T};
    const uint64_t sx[1] = {(uint64_t)MC * 2};
    const uint32_t brow[2] = {MC, 128};
    int rc = tulip_make_tmap(&mx, m.xn, 2, dx, sx, brow, 0);
    if (rc) return rc;
    rc = tulip_make_tmap(&my, m.dxn, 2, dx, sx, brow, 0);
    if (rc) return rc;
    const uint64_t d1[2] = {(uint64_t)MC, (uint64_t)HID};       // We' [E r^2, E]: rows 96 ij .. = shuffle slot ij
    const uint32_t bw[2] = {KBLK, MC};
    rc = tulip_make_tmap(&mw1, m.we, 2, d1, sx, bw, 64);
    if (rc) return rc;
    const uint64_t d2[2] = {(uint64_t)HID, (uint64_t)MC};       // We'^T [E, E r^2]: columns 96 ij ..
    const uint64_t s2[1] = {(uint64_t)HID * 2};
    rc = tulip_make_tmap(&mw2, m.wet, 2, d2, s2, bw, 64);
    if (rc) return rc;
    const uint32_t bk[2] = {KBLK, 128};
    const uint64_t dh[2] = {(uint64_t)HID, (uint64_t)m.T};
    rc = tulip_make_tmap(&mh, m.dh, 2, dh, s2, bk, 64);
    if (rc) return rc;
  }
  MlpArgs a;
  memset(&a, 0, sizeof a);
  a.b1 = m.bias; a.b2 = m.wd;
  a.rows_per_sample = 1;
  a.T = m.T; a.ntiles = (m.T + 127) / 128; a.save = 1;
  a.pred = m.pred; a.target = m.target; a.gscale = m.gscale; a.inv_npix = 1.0f / ((float)m.T * m.r * m.r);
  a.dwd = m.dwd; a.dwd_copies = m.dwd_copies; a.hd_H = m.H; a.hd_W = m.W; a.hd_r = m.r;
  a.hints = tulip_hints();
  static bool configured = false;
  if (!configured) {
    TULIP_CUDA(cudaFuncSetAttribute(mlp_block_fwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Lay<MAX_NCH>::SMEM));
    configured = true;
  }
  const int grid = min(a.ntiles, tulip_num_sms());
  tulip_launch(mlp_block_fwd_kernel<1>, grid, ML_THREADS, Lay<MAX_NCH>::SMEM, st, mx, my, mw1, mw2, mx, mh, a);
  TULIP_CHECK_LAUNCH();
  return TULIP_OK;
}
