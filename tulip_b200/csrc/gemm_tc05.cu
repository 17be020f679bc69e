// tcgen05 / TMEM / TMA GEMMs for sm_100a.
//
// NT: out[M,N] = A[M,K] . B[N,K]^T with fused epilogues (gemm.cuh).  Persistent, warp-specialised CTA of 14 warps:
//   warp 0      TMA producer: A (128 x 64) and B (BN x 64) bf16 blocks, 128B-swizzled, through an mbarrier ring -- or, in
//               the B-stationary schedule (struct Sched), the weight rows of the CTA's n-chunk once and then only A blocks
//   warp 1      MMA issuer: tcgen05.mma cta_group::1 M=128 N=BN K=16, fp32 accumulators in TMEM (2 buffers, 3 for the head)
//   warps 2..13 epilogue, each warp autonomous: tcgen05.ld (one accumulator row per thread, one 32-column box per warp) ->
//               bias / GELU / residual / ... -> its own 2 KB slices of the bf16 staging tile (64B-swizzled) -> its own
//               32 x 32 TMA stores.  The residual / saved pre-activation tile is TMA-loaded into those slices one tile ahead
//               and transformed in place.  Scatter epilogues (PixelShuffle, split, head) store directly.
// Producer and issuer run as whole warps and predicate only the issuing instruction on elect.sync (see tc05.cuh).
// K is walked in 64-column blocks per *segment*; TMA's out-of-bounds zero fill pads K = 96 / 288 to the block size and
// makes the two-source concat (skip Linear) and the PixelShuffle-backward gather (5-D tensor map) plain coordinates.
#include "gemm.cuh"
#include "tc05.cuh"

#include <cstdlib>
#include <cstring>

namespace {

constexpr int BM = 128, BK = 64;
constexpr int EPI_WARPS = 12, EPI_GROUPS = 3, THREADS = 64 + 32 * EPI_WARPS;   // 3 warps per TMEM lane quarter, one column box each
constexpr int BOXC = 32;                                   // output / aux staging box: 32 bf16 columns (64 B rows, SWIZZLE_64B)
constexpr int BOX_BYTES = BM * BOXC * 2;                   // 8 KB

struct Segments {
  int n;                 // K segments (1 plain, 2 concat, 4 unshuffle)
  int len[4];            // columns of A in the segment
  int amap[4];           // 0: mapA, 1: mapA2
  int bcol[4];           // first column of B for the segment
  int bmap[4];           // 0: mapB, 1: mapB2
  int acc[4];            // accumulator the segment feeds (EPI_DGELU2 uses two)
  int sj[4], si[4];      // (j, i) of the shuffle slot for the 5-D gather map
  int a5d;               // A map is the 5-D PixelShuffle-backward view
  int gW;                // 5-D: grid width W
};

struct Maps {
  CUtensorMap A, A2, B, B2, out, out2, aux, aux2;
};

// How the 128 x BN output tiles are dealt to the persistent CTAs.
//   panel == 0: tile-major round robin, A and B blocks of every tile streamed through one ring (any shape).
//   panel == 1: B-stationary row panels.  CTA (chunk, worker) keeps the weight rows of its n-chunk (npc tiles x kb
//               64-column blocks) resident in shared memory for its whole life and walks the 128-row panels worker,
//               worker + nworkers, ...; the A blocks of a panel are loaded once and reused by the npc tiles of the chunk.
//               The SM<->L2 link (~64 GB/s per SM, measured) is what bounds the wide-stage GEMMs: this cuts the bytes a
//               CTA pulls per output tile from A + B to A / npc.
struct Sched {
  int tiles_m, tiles_n;
  int panel;
  int n_chunks, npc, nworkers;
  int kb;            // 64-column K blocks per tile
  int klast;         // 16-column MMA steps in the last K block (1..4)
  int nsa;           // A ring stages (panel mode)
  int bres_bytes;    // resident B region (panel mode), the A ring follows it
  int hints;         // tulip_hints(): L2 eviction priorities of the operand / auxiliary loads
  int cg;            // 2: CTA pairs (cta_group::2) -- pair p walks 256-row x BN tiles, CTA rank r owns rows 128 r .. 128 r + 127
};
__device__ __forceinline__ bool tile_at(const Sched& sc, int it, int& mt, int& nt) {
  if (sc.cg == 2) {
    const int pt = (int)(blockIdx.x >> 1) + it * (int)(gridDim.x >> 1);
    if (pt >= ((sc.tiles_m + 1) >> 1) * sc.tiles_n) return false;
    const int mg = pt / sc.tiles_n;
    nt = pt - mg * sc.tiles_n;
    mt = 2 * mg + (int)(blockIdx.x & 1);                     // an odd tile count leaves one phantom tile: loaded as zeros, never stored
    return true;
  }
  if (!sc.panel) {
    const int tile = blockIdx.x + it * gridDim.x;
    if (tile >= sc.tiles_m * sc.tiles_n) return false;
    mt = tile / sc.tiles_n; nt = tile - mt * sc.tiles_n;
    return true;
  }
  const int worker = blockIdx.x / sc.n_chunks, chunk = blockIdx.x - worker * sc.n_chunks;
  const int mi = it / sc.npc, ni = it - mi * sc.npc;
  mt = worker + mi * sc.nworkers;
  nt = chunk * sc.npc + ni;
  return mt < sc.tiles_m;
}
constexpr int NSA_MAX = 8;

__host__ __device__ constexpr bool epi_tma_out(int epi) {
  return epi == EPI_STORE || epi == EPI_GELU || epi == EPI_RESID || epi == EPI_DGELU || epi == EPI_HEAD_BWD || epi == EPI_DGELU2 ||
         epi == EPI_LNBWD || epi == EPI_STORE_LN || epi == EPI_RESID_LN;
}
__host__ __device__ constexpr bool epi_ln_fwd(int epi) { return epi == EPI_STORE_LN || epi == EPI_RESID_LN; }
// (for the staging-buffer count: the fused-LayerNorm epilogues always use two tiles, input/output and second output)
__host__ __device__ constexpr bool epi_has_aux(int epi) {
  return epi == EPI_RESID || epi == EPI_DGELU || epi == EPI_LNBWD || epi_ln_fwd(epi);
}
// EPI_LNBWD: per-row partial sums exchanged by the three warps of a TMEM lane quarter, [2 parities][4][3][32] float2
__host__ __device__ constexpr int cfg_red_bytes(int epi) { return (epi == EPI_LNBWD || epi_ln_fwd(epi)) ? 2 * 4 * EPI_GROUPS * 32 * 8 : 0; }
// Output staging buffers (each one [128, BN] bf16 tile).  Two let the stores of tile i overlap the epilogue of tile i+1 and,
// for the aux epilogues, hold the in-place aux tile of the next tile.  The wide tile keeps one where it can: its GEMMs are
// the deep-K ones, paced by the depth of the operand ring, and a CTA only sees a handful of tiles.
__host__ __device__ constexpr int cfg_out_bufs(int bn, int epi) {
  return !epi_tma_out(epi) ? 0 : ((bn > 96 && !epi_has_aux(epi)) ? 1 : 2);
}
// operand ring stages: whatever fits next to the staging buffers (227 KB - barriers - alignment slack), at most 8
__host__ __device__ constexpr int cfg_stages(int bn, int epi) {
  const int stage = (128 + bn) * 64 * 2;
  const int n = (227 * 1024 - 512 - 1024 - cfg_out_bufs(bn, epi) * 128 * bn * 2 - cfg_red_bytes(epi)) / stage;
  return n > 8 ? 8 : n;
}

template <int BN, int EPI>
struct Cfg {
  static constexpr bool TMA_OUT = epi_tma_out(EPI);
  static constexpr bool HAS_AUX = epi_has_aux(EPI);
  static constexpr int NACC = (EPI == EPI_DGELU2) ? 2 : 1;                  // accumulators per tile
  static constexpr int NBUF = (EPI == EPI_HEAD) ? 3 : 2;                    // accumulator buffers in TMEM (HEAD: one per epilogue group)
  static constexpr int NBOX = BN / BOXC;
  static constexpr int TILE_BYTES = NBOX * BOX_BYTES;                       // bf16 [128, BN]
  static constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int OUT_BUFS = cfg_out_bufs(BN, EPI);                    // GELU with a saved pre-activation uses both per tile
  // the auxiliary tile (residual / saved pre-activation) is TMA-loaded straight into the output staging slices by the
  // epilogue warps and transformed in place, so it costs no shared memory of its own
  static constexpr int STAGES = cfg_stages(BN, EPI);
  static constexpr int OUT_OFF = STAGES * STAGE_BYTES;
  static constexpr int RED_OFF = OUT_OFF + OUT_BUFS * TILE_BYTES;
  static constexpr int BAR_OFF = RED_OFF + cfg_red_bytes(EPI);
  static constexpr int TOTAL = BAR_OFF + 512 + 1024;                        // barriers + tmem slot, +1024 manual alignment
  static_assert(TOTAL <= 227 * 1024, "shared memory budget");
};

// byte offset of logical 16-byte chunk c (0..3) of row r inside a [128 x 32] bf16 SWIZZLE_64B box
__device__ __forceinline__ int box_off(int r, int c) { return r * 64 + ((c ^ ((r >> 1) & 3)) << 4); }

__device__ __forceinline__ void store_box_row(unsigned char* box, int r, const float (&v)[32]) {
#pragma unroll
  for (int c = 0; c < 4; ++c)
    *reinterpret_cast<uint4*>(box + box_off(r, c)) =
        make_uint4(pack_bf16(v[8 * c], v[8 * c + 1]), pack_bf16(v[8 * c + 2], v[8 * c + 3]), pack_bf16(v[8 * c + 4], v[8 * c + 5]),
                   pack_bf16(v[8 * c + 6], v[8 * c + 7]));
}
__device__ __forceinline__ void load_box_row(const unsigned char* box, int r, float (&v)[32]) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const uint4 u = *reinterpret_cast<const uint4*>(box + box_off(r, c));
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float2 f = unpack_bf16(w[q]);
      v[8 * c + 2 * q] = f.x;
      v[8 * c + 2 * q + 1] = f.y;
    }
  }
}
__device__ __forceinline__ void add_bias32(const float* __restrict__ bias, int n, float (&v)[32]) {
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float4 b = *reinterpret_cast<const float4*>(bias + n + 4 * q);
    v[4 * q] += b.x; v[4 * q + 1] += b.y; v[4 * q + 2] += b.z; v[4 * q + 3] += b.w;
  }
}

// direct-store epilogues (scatter destinations): 16 consecutive columns of one row
template <int EPI>
__device__ __forceinline__ void epi_direct16(const GemmArgs& g, int m, int n, float (&v)[16]) {
  bf16* dst = nullptr;
  if (EPI == EPI_PIXSHUF) {
    if (g.bias) {
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] += g.bias[n + i];
    }
    const int ij = n / g.g_Cc, c = n % g.g_Cc;
    dst = g.out + pixshuf_row(m, ij, g.g_H, g.g_W) * g.ldo + c;
  } else if (EPI == EPI_SPLIT2) {
    dst = n < g.split_col ? g.out + (long)m * g.ldo + n : g.out2 + (long)m * g.ldo2 + (n - g.split_col);
  } else {  // EPI_ROWSCALE
    const float s = g.row_scale ? g.row_scale[m / g.rows_per_sample] : 1.0f;
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] *= s;
    dst = g.out + (long)m * g.ldo + n;
  }
  uint4* p = reinterpret_cast<uint4*>(dst);
  p[0] = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
  p[1] = make_uint4(pack_bf16(v[8], v[9]), pack_bf16(v[10], v[11]), pack_bf16(v[12], v[13]), pack_bf16(v[14], v[15]));
}

// VAR = 1: head backward with two 96-channel groups (hd_E = 192): a second set of d(decoder_pred.weight) accumulators
// VAR = 2: the head epilogues in their FinalPatchExpanding form (GemmArgs::hd_ln, hd_E = 96): per-pixel LayerNorm over the
//          tile's 96 columns instead of bias + LeakyReLU
// CG = 2: CTA pairs (tc05.cuh): tile-major schedule only; the even CTA issues M = 256 MMAs for both, each CTA loads its own A
// rows and half of the B rows, so the deep-K GEMMs of stages 2-3 pull A + B/2 per K block over their SM's L2 port.
template <int BN, int EPI, int VAR = 0, int CG = 1>
__global__ void __launch_bounds__(THREADS, 1)
gemm_nt_tc05_kernel(const __grid_constant__ Maps maps, const __grid_constant__ GemmArgs g, const __grid_constant__ Segments sg,
                    const __grid_constant__ Sched sc) {
  using CF = Cfg<BN, EPI>;
  constexpr int STAGES = CF::STAGES;
  pdl_trigger();                                            // successor may start its own set-up right away
  constexpr int NBUF = CF::NBUF;
  constexpr int TMEM_COLS = (NBUF * CF::NACC * BN <= 128) ? 128 : (NBUF * CF::NACC * BN <= 256 ? 256 : 512);
  static_assert(NBUF * CF::NACC * BN <= 512, "TMEM budget");
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + CF::BAR_OFF);
  uint64_t* empty = full + NSA_MAX;
  uint64_t* tfull = empty + NSA_MAX;
  uint64_t* tempty = tfull + 3;
  uint64_t* bfull = tempty + 3;                             // panel mode: resident B region has landed
  uint64_t* auxbar = bfull + 1;                             // [EPI_WARPS][2]: per-warp, per-staging-buffer aux arrival
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(auxbar + 2 * EPI_WARPS);

  const int warp = tc::warp_idx_sync(), lane = threadIdx.x & 31;
  const uint32_t rank = (CG == 2) ? (blockIdx.x & 1u) : 0u;   // cluster dims (2,1,1): rank in the pair
  if (threadIdx.x == 0) {
    for (int s = 0; s < NSA_MAX; ++s) { tc::mbar_init(full + s, 1); tc::mbar_init(empty + s, 1); }
    tc::mbar_init(bfull, 1);
    for (int b = 0; b < 3; ++b) {                             // HEAD: one epilogue group (4 warps) drains a buffer
      tc::mbar_init(tfull + b, 1); tc::mbar_init(tempty + b, EPI == EPI_HEAD ? 4 : CG * EPI_WARPS);   // pairs: both CTAs' warps
    }
    for (int b = 0; b < 2 * EPI_WARPS; ++b) tc::mbar_init(auxbar + b, 1);
    tc::fence_barrier_init();
  }
  if (CG == 2) {
    if (warp == 1) tc::tmem_alloc_cg2<TMEM_COLS>(tmem_slot);
    tc::fence_before_sync();
    tc::cluster_sync_all();                                   // the peer's barriers exist before anything arrives on them
  } else {
    if (warp == 1) tc::tmem_alloc<TMEM_COLS>(tmem_slot);
    tc::fence_before_sync();
    __syncthreads();
  }
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                                               // set-up above overlapped the previous kernel's tail

  const int tiles_n = sc.tiles_n;

  if (warp == 0 && sc.panel) {
    // ---- TMA producer, B-stationary panels ----
    const int worker = blockIdx.x / sc.n_chunks, chunk = blockIdx.x - worker * sc.n_chunks;
    // one instruction form on the issue path whatever the hint mask says: the policy is chosen once, outside the loops
    const uint64_t polA = (sc.hints & 2) ? tc::l2_policy_evict_first() : tc::l2_policy_evict_normal();
    const uint64_t polB = (sc.hints & 64) ? tc::l2_policy_evict_last() : tc::l2_policy_evict_normal();
    if (tc::elect_one_sync()) {
      tc::prefetch_tensormap(&maps.A);
      tc::prefetch_tensormap(&maps.B);
      tc::mbar_expect_tx(bfull, sc.bres_bytes);
      for (int ni = 0; ni < sc.npc; ++ni)
        for (int kbi = 0; kbi < sc.kb; ++kbi)
          tc::tma_load_2d_hint(smem + (ni * sc.kb + kbi) * CF::B_BYTES, &maps.B, bfull, kbi * BK, (chunk * sc.npc + ni) * BN, polB);
    }
    __syncwarp();
    int stage = 0; uint32_t phase = 0;
    for (int mt = worker; mt < sc.tiles_m; mt += sc.nworkers) {
      for (int kbi = 0; kbi < sc.kb; ++kbi) {
        tc::mbar_wait(empty + stage, phase ^ 1);
        if (tc::elect_one_sync()) {
          tc::mbar_expect_tx(full + stage, CF::A_BYTES);
          tc::tma_load_2d_hint(smem + sc.bres_bytes + stage * CF::A_BYTES, &maps.A, full + stage, kbi * BK, mt * BM, polA);
        }
        __syncwarp();
        if (++stage == sc.nsa) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1 && sc.panel) {
    // ---- MMA issuer, B-stationary panels: the A blocks of a panel are waited for by its first tile and released by its last
    // Descriptors are advanced incrementally (the start-address field is the only part that changes and never carries out
    // of its 14 bits for shared-memory addresses): the issue chain of this warp -- uniform-datapath instructions with long
    // dependent latencies -- is what paces short-K tiles, so it is kept to a wait, four UTCHMMAs and a commit per K block.
    constexpr uint32_t idesc = tc::make_idesc(BM, BN, 0, 0);
    constexpr uint32_t A16 = CF::A_BYTES >> 4, B16 = CF::B_BYTES >> 4;
    const int worker = blockIdx.x / sc.n_chunks;
    const uint64_t dA0 = tc::make_desc_kmajor_sw128(smem + sc.bres_bytes);
    const uint64_t dB0 = tc::make_desc_kmajor_sw128(smem);
    tc::mbar_wait(bfull, 0);
    int stage = 0; uint32_t phase = 0;
    uint64_t da_stage = dA0;
    int buf = 0; uint32_t tphase = 0;
    const int kb = sc.kb, npc = sc.npc, nsa = sc.nsa, klast = sc.klast;
    for (int mt = worker; mt < sc.tiles_m; mt += sc.nworkers) {
      int st = stage; uint32_t ph = phase; uint64_t da = da_stage;
      uint64_t db = dB0;
      for (int ni = 0; ni < npc; ++ni) {
        tc::mbar_wait(tempty + buf, tphase ^ 1);
        tc::fence_after_sync();
        const uint32_t tmem_d = tmem_base + buf * CF::NACC * BN;
        st = stage; ph = phase; da = da_stage;
        for (int kbi = 0; kbi < kb; ++kbi) {
          if (ni == 0) { tc::mbar_wait(full + st, ph); tc::fence_after_sync(); }
          if (tc::elect_one_sync()) {
            if (kbi < kb - 1 || klast == 4) {
#pragma unroll
              for (int k = 0; k < BK / 16; ++k) tc::umma_bf16(tmem_d, da + 2 * k, db + 2 * k, idesc, (kbi | k) ? 1u : 0u);
            } else {                                          // zero-padded K steps of the last block are skipped
#pragma unroll
              for (int k = 0; k < BK / 16 - 1; ++k)
                if (k < klast) tc::umma_bf16(tmem_d, da + 2 * k, db + 2 * k, idesc, (kbi | k) ? 1u : 0u);
            }
            if (ni == npc - 1) tc::umma_commit(empty + st);
          }
          __syncwarp();
          db += B16;
          if (++st == nsa) { st = 0; ph ^= 1; da = dA0; } else { da += A16; }
        }
        if (tc::elect_one_sync()) tc::umma_commit(tfull + buf);
        __syncwarp();
        if (++buf == NBUF) { buf = 0; tphase ^= 1; }
      }
      stage = st; phase = ph; da_stage = da;
    }
  } else if (warp == 0) {
    // TMA producer: the whole warp walks the loop (uniform control flow), one elected lane issues
    if (tc::elect_one_sync()) {
      tc::prefetch_tensormap(&maps.A);
      tc::prefetch_tensormap(&maps.B);
    }
    const uint64_t polA = (sc.hints & 4) ? tc::l2_policy_evict_first() : tc::l2_policy_evict_normal();
    const uint64_t polB = (sc.hints & 64) ? tc::l2_policy_evict_last() : tc::l2_policy_evict_normal();
    int stage = 0; uint32_t phase = 0;
    int mt, nt;
    for (int it = 0; tile_at(sc, it, mt, nt); ++it) {
      const int m0 = mt * BM, n0 = nt * BN;
      for (int s = 0; s < sg.n; ++s) {
        const int nb = (sg.len[s] + BK - 1) / BK;
        for (int kb = 0; kb < nb; ++kb) {
          tc::mbar_wait(empty + stage, phase ^ 1);
          if (CG == 2) {
            // both CTAs load (own A rows, own half of the B rows); the bytes of both are counted on the leader's barrier
            if (tc::elect_one_sync()) {
              unsigned char* a = smem + stage * CF::STAGE_BYTES;
              unsigned char* b = a + CF::A_BYTES;
              if (rank == 0) tc::mbar_expect_tx(full + stage, 2 * CF::A_BYTES + CF::B_BYTES);
              tc::tma_load_2d_cg2(a, sg.amap[s] ? &maps.A2 : &maps.A, full + stage, kb * BK, m0);
              tc::tma_load_2d_cg2(b, sg.bmap[s] ? &maps.B2 : &maps.B, full + stage, sg.bcol[s] + kb * BK, n0 + (int)rank * (BN / 2));
            }
          } else
          if (tc::elect_one_sync()) {
            unsigned char* a = smem + stage * CF::STAGE_BYTES;
            unsigned char* b = a + CF::A_BYTES;
            tc::mbar_expect_tx(full + stage, CF::STAGE_BYTES);
            if (sg.a5d) {
              const int bh0 = m0 / sg.gW, w0 = m0 % sg.gW;
              tc::tma_load_5d(a, &maps.A, full + stage, kb * BK, sg.sj[s], w0, sg.si[s], bh0);
            } else {
              tc::tma_load_2d_hint(a, sg.amap[s] ? &maps.A2 : &maps.A, full + stage, kb * BK, m0, polA);
            }
            tc::tma_load_2d_hint(b, sg.bmap[s] ? &maps.B2 : &maps.B, full + stage, sg.bcol[s] + kb * BK, n0, polB);
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1 && CG == 2 && rank != 0) {
    // the peer CTA of a pair issues nothing: its operands are read and its accumulator rows written by the leader's MMAs
  } else if (warp == 1) {
    // MMA issuer: same scheme; descriptors stay in uniform registers
    constexpr uint32_t idesc = tc::make_idesc(CG * BM, BN, 0, 0);
    constexpr uint32_t S16 = CF::STAGE_BYTES >> 4, A16 = CF::A_BYTES >> 4;
    const uint64_t d0 = tc::make_desc_kmajor_sw128(smem);
    uint64_t da = d0;                                       // A descriptor of `stage`; its B block follows A16 further
    int stage = 0; uint32_t phase = 0;
    int buf = 0; uint32_t tphase = 0;
    int mt, nt;
    for (int it = 0; tile_at(sc, it, mt, nt); ++it) {
      tc::mbar_wait(tempty + buf, tphase ^ 1);             // epilogue has drained this accumulator buffer
      tc::fence_after_sync();
      const uint32_t tmem_d = tmem_base + buf * CF::NACC * BN;
      uint32_t started = 0;                                 // bit a: accumulator a has received its first MMA
      for (int s = 0; s < sg.n; ++s) {
        const int nb = (sg.len[s] + BK - 1) / BK;
        const int ac = CF::NACC > 1 ? sg.acc[s] : 0;
        const uint32_t tmem_acc = tmem_d + ac * BN;
        for (int kb = 0; kb < nb; ++kb) {
          tc::mbar_wait(full + stage, phase);
          tc::fence_after_sync();
          if (tc::elect_one_sync()) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {             // +32 bytes per K=16 step inside the 128B swizzle atom
              if (CG == 2) tc::umma_bf16_cg2(tmem_acc, da + 2 * k, da + A16 + 2 * k, idesc, (k > 0) ? 1u : ((started >> ac) & 1u));
              else tc::umma_bf16(tmem_acc, da + 2 * k, da + A16 + 2 * k, idesc, (k > 0) ? 1u : ((started >> ac) & 1u));
            }
            if (CG == 2) tc::umma_commit_cg2(empty + stage);  // the slot is free in BOTH CTAs once these MMAs have read it
            else tc::umma_commit(empty + stage);            // smem slot free once these MMAs have read it
          }
          __syncwarp();
          started |= 1u << ac;
          if (++stage == STAGES) { stage = 0; phase ^= 1; da = d0; } else { da += S16; }
        }
      }
      if (tc::elect_one_sync()) {                           // accumulator complete (pairs: in both CTAs)
        if (CG == 2) tc::umma_commit_cg2(tfull + buf); else tc::umma_commit(tfull + buf);
      }
      __syncwarp();
      if (++buf == NBUF) { buf = 0; tphase ^= 1; }
    }
  } else {
    const uint64_t aux_pol = (sc.hints & 8) ? tc::l2_policy_evict_first() : tc::l2_policy_evict_normal();   // auxiliary tiles are read once here
    const int q = warp & 3;                                   // TMEM lane quarter this warp may access
    const int jgrp = (warp - 2) >> 2;                         // which column boxes of the tile this warp handles
    const int r = q * 32 + lane;                              // row inside the tile
    // HEAD_BWD: per-thread column sums for d(decoder_pred.weight), one set per 96-channel group of the head (hd_E = 96 or 192)
    float cwacc[(CF::NBOX + EPI_GROUPS - 1) / EPI_GROUPS][32], cwacc_hi[VAR == 1 ? (CF::NBOX + EPI_GROUPS - 1) / EPI_GROUPS : 1][32];
    if (EPI == EPI_HEAD_BWD) {
#pragma unroll
      for (int a = 0; a < (CF::NBOX + EPI_GROUPS - 1) / EPI_GROUPS; ++a)
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          cwacc[a][i] = 0.f;
          if (VAR == 1) cwacc_hi[a][i] = 0.f;
        }
    }
    // Every epilogue warp is autonomous: it owns rows [32q, 32q+32) of its column boxes, stages them in its own 2 KB
    // slices of the output buffers and issues its own 32 x 32 TMA stores.  No CTA-wide barrier sits on the per-tile path;
    // warps drift apart by up to one tile (two accumulator buffers), which overlaps their latency chains.
    // EPI_HEAD instead hands whole tiles to groups of 4 warps (group jgrp <-> accumulator buffer jgrp of 3).
    int buf = (EPI == EPI_HEAD) ? jgrp : 0; uint32_t tphase = 0;
    int obuf = 0;
    uint32_t auxphase = 0;                                    // bit b: parity of this warp's aux barrier for staging buffer b
    uint64_t* mybar = auxbar + 2 * (warp - 2);
    constexpr int it_step = (EPI == EPI_HEAD) ? EPI_GROUPS : 1;
    // aux tile slices of this warp for tile (mt_a, nt_a) -> staging buffer ob (rows 32q.. of its column boxes), one elected lane
    auto issue_aux = [&](int mt_a, int nt_a, int ob_a) {
      const int m0a = mt_a * BM + q * 32, n0a = nt_a * BN;
      if (m0a >= g.M) return;                                 // nothing of this slice is inside the matrix: no load, no wait
      constexpr int MYBOX = (CF::NBOX + EPI_GROUPS - 1) / EPI_GROUPS;
      int nmine = 0;
#pragma unroll
      for (int jj = 0; jj < MYBOX; ++jj) nmine += (jgrp + jj * EPI_GROUPS < CF::NBOX) ? 1 : 0;
      tc::mbar_expect_tx(mybar + ob_a, nmine * 2048);
#pragma unroll
      for (int jj = 0; jj < MYBOX; ++jj) {
        const int j = jgrp + jj * EPI_GROUPS;
        if (j < CF::NBOX)
          tc::tma_load_2d_hint(smem + CF::OUT_OFF + ob_a * CF::TILE_BYTES + j * BOX_BYTES + q * 2048, &maps.aux, mybar + ob_a, n0a + j * BOXC, m0a,
                               aux_pol);
      }
    };
    int mt, nt;
    // FinalPatchExpanding head: bw = sum_c beta[c] wd[c] and mean_c(gamma[c] wd[c]) (per-thread constants, fixed order)
    float hl_bw = 0.f, hl_mgw = 0.f, hl_dsum = 0.f;
    if ((EPI == EPI_HEAD || EPI == EPI_HEAD_BWD) && VAR == 2) {
      for (int cidx = 0; cidx < BN; ++cidx) {
        const float wdc = g.wd[cidx];
        hl_bw = fmaf(g.ln_b[cidx], wdc, hl_bw);
        hl_mgw = fmaf(g.ln_w[cidx], wdc, hl_mgw);
      }
      hl_mgw *= 1.0f / (float)BN;
    }
    if (EPI == EPI_LNBWD) {
      // ---- LayerNorm backward on whole output rows (tiles_n == 1).  Staging buffer 0: LN input rows x (-> scaled dx when a
      // second output is asked for), buffer 1: residual-path gradient (-> dx), both transformed in place.  Per row the three
      // warps of a lane quarter own one (BN 96) or two (BN 192) 32-column boxes each; their partial row sums meet in shared
      // memory behind a 96-thread named barrier.  The column sums d(gamma), d(beta) are reduced over the 32 rows of a warp by
      // recursive halving and accumulated in one register per box, so a CTA issues 2 N atomics in its whole life.
      constexpr int MYBOX = CF::NBOX / EPI_GROUPS;
      static_assert(EPI != EPI_LNBWD || CF::NBOX % EPI_GROUPS == 0, "whole boxes per warp");
      unsigned char* const bx = smem + CF::OUT_OFF;
      unsigned char* const br = smem + CF::OUT_OFF + CF::TILE_BYTES;
      float2* const red = reinterpret_cast<float2*>(smem + CF::RED_OFF);
      const bool has_res = g.aux2 != nullptr, has_dxs = g.out2 != nullptr;
      const float invC = 1.0f / (float)BN;
      float cg[MYBOX], cb[MYBOX];
#pragma unroll
      for (int jj = 0; jj < MYBOX; ++jj) cg[jj] = cb[jj] = 0.f;
      auto issue_in = [&](int mt_a) {
        const int m0a = mt_a * BM + q * 32;
        if (m0a >= g.M) return;
        tc::mbar_expect_tx(mybar, MYBOX * 2048 * (has_res ? 2 : 1));
#pragma unroll
        for (int jj = 0; jj < MYBOX; ++jj) {
          const int j = jgrp + jj * EPI_GROUPS;
          tc::tma_load_2d_hint(bx + j * BOX_BYTES + q * 2048, &maps.aux, mybar, j * BOXC, m0a, aux_pol);
          if (has_res) tc::tma_load_2d_hint(br + j * BOX_BYTES + q * 2048, &maps.aux2, mybar, j * BOXC, m0a, aux_pol);
        }
      };
      if (tile_at(sc, 0, mt, nt)) {
        if (tc::elect_one_sync()) issue_in(mt);
        __syncwarp();
      }
      uint32_t inphase = 0;
      for (int it = 0; tile_at(sc, it, mt, nt); ++it) {
        const int m0 = mt * BM, m = m0 + r;
        const bool valid = m < g.M;
        float mean = 0.f, rstd = 0.f, rsc = 1.f;
        if (valid) {
          const float2 st = *reinterpret_cast<const float2*>(g.ln_stats + 2 * (long)m);
          mean = st.x; rstd = st.y;
          if (has_dxs) rsc = g.row_scale[m / g.rows_per_sample];
        }
        tc::mbar_wait(tfull + buf, tphase);
        tc::fence_after_sync();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * BN;
        if (m0 + q * 32 < g.M) {
          tc::mbar_wait(mybar, inphase);
          inphase ^= 1u;
        }
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int jj = 0; jj < MYBOX; ++jj) {
          const int j = jgrp + jj * EPI_GROUPS;
          float v[32], h[32];
          tc::tmem_ld32(taddr + j * BOXC, v);
          load_box_row(bx + j * BOX_BYTES, r, h);
#pragma unroll
          for (int q4 = 0; q4 < 8; ++q4) {
            const float4 w = *reinterpret_cast<const float4*>(g.ln_w + j * BOXC + 4 * q4);
            const float ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int i = 4 * q4 + e;
              const float xh = valid ? (h[i] - mean) * rstd : 0.f;
              const float t = v[i] * ww[e];
              s1 += t;
              s2 = fmaf(t, xh, s2);
              h[i] = v[i] * xh;
            }
          }
          cg[jj] += warp_colsum32(h, lane);
          cb[jj] += warp_colsum32(v, lane);
        }
        float2* const myred = red + ((it & 1) * 4 + q) * (EPI_GROUPS * 32);
        myred[jgrp * 32 + lane] = make_float2(s1, s2);
        tc::named_bar_sync(1 + q, 32 * EPI_GROUPS);
        {
          const float2 p0 = myred[lane], p1 = myred[32 + lane], p2 = myred[64 + lane];
          s1 = (p0.x + p1.x) + p2.x;
          s2 = (p0.y + p1.y) + p2.y;
        }
        const float m1 = s1 * invC, m2 = s2 * invC;
#pragma unroll
        for (int jj = 0; jj < MYBOX; ++jj) {
          const int j = jgrp + jj * EPI_GROUPS;
          float v[32], h[32];
          tc::tmem_ld32(taddr + j * BOXC, v);
          load_box_row(bx + j * BOX_BYTES, r, h);
#pragma unroll
          for (int q4 = 0; q4 < 8; ++q4) {
            const float4 w = *reinterpret_cast<const float4*>(g.ln_w + j * BOXC + 4 * q4);
            const float ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int i = 4 * q4 + e;
              const float xh = valid ? (h[i] - mean) * rstd : 0.f;
              v[i] = rstd * (v[i] * ww[e] - m1 - xh * m2);
            }
          }
          if (has_res) {
            load_box_row(br + j * BOX_BYTES, r, h);
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] += h[i];
          }
          store_box_row(br + j * BOX_BYTES, r, v);
          if (has_dxs) {                                     // the scaled copy is formed from the rounded dx (as layernorm_bwd does)
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = rsc * bf16_round(v[i]);
            store_box_row(bx + j * BOX_BYTES, r, v);
          }
        }
        tc::fence_before_sync();
        tc::fence_proxy_async();
        __syncwarp();
        if (lane == 0) { if (CG == 2) tc::mbar_arrive_cluster(tempty + buf, 0); else tc::mbar_arrive(tempty + buf); }
        if (tc::elect_one_sync()) {
          if (m0 + q * 32 < g.M) {
#pragma unroll
            for (int jj = 0; jj < MYBOX; ++jj) {
              const int j = jgrp + jj * EPI_GROUPS;
              tc::tma_store_2d(&maps.out, br + j * BOX_BYTES + q * 2048, j * BOXC, m0 + q * 32);
              if (has_dxs) tc::tma_store_2d(&maps.out2, bx + j * BOX_BYTES + q * 2048, j * BOXC, m0 + q * 32);
            }
          }
          tc::tma_store_commit();
          int mt2, nt2;
          if (tile_at(sc, it + 1, mt2, nt2)) {               // the slices are free once the stores have read them
            tc::tma_store_wait_read<0>();
            issue_in(mt2);
          }
        }
        __syncwarp();
        if (++buf == NBUF) { buf = 0; tphase ^= 1; }
      }
      {
        const int copy_off = g.ln_copies > 1 ? (int)(blockIdx.x % g.ln_copies) * g.ln_stride : 0;
#pragma unroll
        for (int jj = 0; jj < MYBOX; ++jj) {
          const int n = (jgrp + jj * EPI_GROUPS) * BOXC + lane;
          atomicAdd(g.ln_dw + copy_off + n, cg[jj]);
          atomicAdd(g.ln_db + copy_off + n, cb[jj]);
        }
      }
      if (tc::elect_one_sync()) tc::tma_store_wait<0>();
    } else if (epi_ln_fwd(EPI)) {
      // ---- EPI_STORE / EPI_RESID on whole rows followed by the LayerNorm that reads them next.  Staging buffer 0: residual
      // tile -> out (in place), buffer 1: LayerNorm(out).  The statistics are taken from the bf16-rounded rows, in two passes
      // (mean, then centred squares) as layernorm_fwd does; the three warps of a lane quarter exchange their partial sums
      // through shared memory twice per tile.
      constexpr bool RES = (EPI == EPI_RESID_LN);
      constexpr int MYBOX = CF::NBOX / EPI_GROUPS;
      unsigned char* const b0 = smem + CF::OUT_OFF;
      unsigned char* const b1 = smem + CF::OUT_OFF + CF::TILE_BYTES;
      float2* const red = reinterpret_cast<float2*>(smem + CF::RED_OFF);
      const float invC = 1.0f / (float)BN;
      auto issue_in = [&](int mt_a) {
        const int m0a = mt_a * BM + q * 32;
        if (!RES || m0a >= g.M) return;
        tc::mbar_expect_tx(mybar, MYBOX * 2048);
#pragma unroll
        for (int jj = 0; jj < MYBOX; ++jj) {
          const int j = jgrp + jj * EPI_GROUPS;
          tc::tma_load_2d_hint(b0 + j * BOX_BYTES + q * 2048, &maps.aux, mybar, j * BOXC, m0a, aux_pol);
        }
      };
      if (RES && tile_at(sc, 0, mt, nt)) {
        if (tc::elect_one_sync()) issue_in(mt);
        __syncwarp();
      }
      uint32_t inphase = 0;
      for (int it = 0; tile_at(sc, it, mt, nt); ++it) {
        const int m0 = mt * BM, m = m0 + r;
        const float rs = (RES && g.row_scale) ? g.row_scale[(m < g.M ? m : g.M - 1) / g.rows_per_sample] : 1.0f;
        tc::mbar_wait(tfull + buf, tphase);
        tc::fence_after_sync();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * BN;
        if (RES && m0 + q * 32 < g.M) {
          tc::mbar_wait(mybar, inphase);
          inphase ^= 1u;
        }
        float s = 0.f;
#pragma unroll
        for (int jj = 0; jj < MYBOX; ++jj) {
          const int j = jgrp + jj * EPI_GROUPS;
          float v[32];
          tc::tmem_ld32(taddr + j * BOXC, v);
          if (g.bias) add_bias32(g.bias, j * BOXC, v);
          if (RES) {
            float a[32];
            load_box_row(b0 + j * BOX_BYTES, r, a);
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = a[i] + rs * v[i];
          }
#pragma unroll
          for (int i = 0; i < 32; ++i) { v[i] = bf16_round(v[i]); s += v[i]; }
          store_box_row(b0 + j * BOX_BYTES, r, v);
        }
        tc::fence_before_sync();                             // accumulator consumed
        __syncwarp();
        if (lane == 0) { if (CG == 2) tc::mbar_arrive_cluster(tempty + buf, 0); else tc::mbar_arrive(tempty + buf); }
        float2* const red0 = red + (0 * 4 + q) * (EPI_GROUPS * 32);
        float2* const red1 = red + (1 * 4 + q) * (EPI_GROUPS * 32);
        red0[jgrp * 32 + lane].x = s;
        tc::named_bar_sync(1 + q, 32 * EPI_GROUPS);
        const float mean = ((red0[lane].x + red0[32 + lane].x) + red0[64 + lane].x) * invC;
        float sq = 0.f;
#pragma unroll
        for (int jj = 0; jj < MYBOX; ++jj) {
          float v[32];
          load_box_row(b0 + (jgrp + jj * EPI_GROUPS) * BOX_BYTES, r, v);
#pragma unroll
          for (int i = 0; i < 32; ++i) sq = fmaf(v[i] - mean, v[i] - mean, sq);
        }
        red1[jgrp * 32 + lane].x = sq;
        tc::named_bar_sync(1 + q, 32 * EPI_GROUPS);
        const float rstd = rsqrtf(((red1[lane].x + red1[32 + lane].x) + red1[64 + lane].x) * invC + g.ln_eps);
#pragma unroll
        for (int jj = 0; jj < MYBOX; ++jj) {
          const int j = jgrp + jj * EPI_GROUPS;
          float v[32];
          load_box_row(b0 + j * BOX_BYTES, r, v);
#pragma unroll
          for (int q4 = 0; q4 < 8; ++q4) {
            const float4 w = *reinterpret_cast<const float4*>(g.ln_w + j * BOXC + 4 * q4);
            const float4 b = *reinterpret_cast<const float4*>(g.ln_b + j * BOXC + 4 * q4);
            v[4 * q4] = (v[4 * q4] - mean) * rstd * w.x + b.x;
            v[4 * q4 + 1] = (v[4 * q4 + 1] - mean) * rstd * w.y + b.y;
            v[4 * q4 + 2] = (v[4 * q4 + 2] - mean) * rstd * w.z + b.z;
            v[4 * q4 + 3] = (v[4 * q4 + 3] - mean) * rstd * w.w + b.w;
          }
          store_box_row(b1 + j * BOX_BYTES, r, v);
        }
        if (jgrp == 0 && m < g.M) *reinterpret_cast<float2*>(g.ln_ystats + 2 * (long)m) = make_float2(mean, rstd);
        tc::fence_proxy_async();
        __syncwarp();
        if (tc::elect_one_sync()) {
          if (m0 + q * 32 < g.M) {
#pragma unroll
            for (int jj = 0; jj < MYBOX; ++jj) {
              const int j = jgrp + jj * EPI_GROUPS;
              tc::tma_store_2d(&maps.out, b0 + j * BOX_BYTES + q * 2048, j * BOXC, m0 + q * 32);
              tc::tma_store_2d(&maps.out2, b1 + j * BOX_BYTES + q * 2048, j * BOXC, m0 + q * 32);
            }
          }
          tc::tma_store_commit();
          int mt2, nt2;
          if (tile_at(sc, it + 1, mt2, nt2)) {               // the slices are free once the stores have read them
            tc::tma_store_wait_read<0>();
            issue_in(mt2);
          }
        }
        __syncwarp();
        if (++buf == NBUF) { buf = 0; tphase ^= 1; }
      }
      if (tc::elect_one_sync()) tc::tma_store_wait<0>();
    } else {
    if (CF::HAS_AUX && tile_at(sc, 0, mt, nt)) {
      if (tc::elect_one_sync()) issue_aux(mt, nt, 0);
      __syncwarp();
    }
    for (int it = (EPI == EPI_HEAD) ? jgrp : 0; tile_at(sc, it, mt, nt); it += it_step) {
      const int m0 = mt * BM, n0 = nt * BN;
      const int m = m0 + r;
      // the first box's bias vector is fetched before the accumulator wait so its latency is off the critical path
      float bias0[32];
      if ((EPI == EPI_STORE || EPI == EPI_GELU || EPI == EPI_RESID) && g.bias) {
#pragma unroll
        for (int q8 = 0; q8 < 8; ++q8) {
          const float4 b = *reinterpret_cast<const float4*>(g.bias + n0 + jgrp * BOXC + 4 * q8);
          bias0[4 * q8] = b.x; bias0[4 * q8 + 1] = b.y; bias0[4 * q8 + 2] = b.z; bias0[4 * q8 + 3] = b.w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) bias0[i] = 0.f;
      }
      // per-row scalars of the epilogue come from global memory: fetch them before the accumulator wait as well
      float dp = 0.f;
      float hl_mean = 0.f, hl_rstd = 0.f, hl_s = 0.f;      // FinalPatchExpanding: the pixel's LayerNorm statistics, (pred - bw) / E
      int ij = 0, c0 = 0;
      if (EPI == EPI_HEAD_BWD) {
        ij = n0 / g.hd_E; c0 = n0 % g.hd_E;
        if (m < g.M) {
          const long px = head_pixel(g, m, ij);
          const float pr = g.pred[px];
          const float d = pr - g.target[px];
          const float gs = g.gscale[0] * g.hd_inv_npix;
          dp = d > 0.f ? gs : (d < 0.f ? -gs : 0.f);
          if (VAR == 2) {
            const float2 ms = *reinterpret_cast<const float2*>(g.ln_stats + 2 * px);
            hl_mean = ms.x; hl_rstd = ms.y;
            hl_s = (pr - hl_bw) * (1.0f / (float)BN);
            if (jgrp == 0) hl_dsum += dp;                    // one warp per lane quarter counts the row
          }
        }
      }
      const float rs = (EPI == EPI_RESID && g.row_scale) ? g.row_scale[(m < g.M ? m : g.M - 1) / g.rows_per_sample] : 1.0f;
      const bool two_out = (EPI == EPI_GELU) && g.out2 != nullptr;               // pre-activation is saved too: both buffers per tile
      if (CF::TMA_OUT) {
        // staging slices must have been read out by this warp's earlier TMA stores.  Bulk-group bookkeeping is per thread:
        // elect.sync picks the same lane for the same (full) mask every time.  One group per tile and warp: the slices of
        // buffer `obuf` were last read by the group committed two tiles ago.
        if (tc::elect_one_sync()) {
          if (CF::HAS_AUX) {
            // the other buffer's slices were stored by the previous tile: once that store has read them, the NEXT tile's
            // aux slices are loaded into them, one full tile ahead of their use
            tc::tma_store_wait_read<0>();
            int mt2, nt2;
            if (tile_at(sc, it + it_step, mt2, nt2)) issue_aux(mt2, nt2, obuf ^ 1);
          } else if (two_out || CF::OUT_BUFS == 1) {
            tc::tma_store_wait_read<0>();
          } else {
            tc::tma_store_wait_read<1>();
          }
        }
        __syncwarp();
      }
      tc::mbar_wait(tfull + buf, tphase);
      tc::fence_after_sync();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * CF::NACC * BN;

      if (CF::TMA_OUT) {
        unsigned char* ob = smem + CF::OUT_OFF + ((two_out || CF::OUT_BUFS == 1) ? 0 : obuf) * CF::TILE_BYTES;
        unsigned char* ob2 = smem + CF::OUT_OFF + (CF::OUT_BUFS - 1) * CF::TILE_BYTES; // GELU: activation tile (two_out only)
        const unsigned char* ab = ob;                                                 // aux slices are transformed in place
        if (CF::HAS_AUX && m0 + q * 32 < g.M) {
          tc::mbar_wait(mybar + obuf, (auxphase >> obuf) & 1u);
          auxphase ^= 1u << obuf;
        }
#pragma unroll
        for (int jj = 0; jj < (CF::NBOX + EPI_GROUPS - 1) / EPI_GROUPS; ++jj) {
          const int j = jgrp + jj * EPI_GROUPS;
          if (j >= CF::NBOX) break;
          float v[32];
          tc::tmem_ld32(taddr + j * BOXC, v);
          const int n = n0 + j * BOXC;
          if (EPI == EPI_STORE || EPI == EPI_GELU || EPI == EPI_RESID) {
            if (jj == 0) {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] += bias0[i];
            } else if (g.bias) {
              add_bias32(g.bias, n, v);
            }
          }
          if (EPI == EPI_STORE) {
          } else if (EPI == EPI_GELU) {
            if (two_out) store_box_row(ob + j * BOX_BYTES, r, v);
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = gelu_erf(v[i]);
          } else if (EPI == EPI_RESID) {
            float a[32];
            load_box_row(ab + j * BOX_BYTES, r, a);
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = a[i] + rs * v[i];
          } else if (EPI == EPI_DGELU) {
            float a[32];
            load_box_row(ab + j * BOX_BYTES, r, a);
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] *= gelu_erf_grad(a[i]);
          } else if (EPI == EPI_DGELU2) {
            float a[32];
            tc::tmem_ld32(taddr + BN + j * BOXC, a);                       // recomputed fc1 pre-activation (second accumulator)
            if (g.bias) add_bias32(g.bias, n, a);
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] *= gelu_erf_grad(a[i]);
          } else if (EPI == EPI_HEAD_BWD && VAR == 2) {
            // y = xhat gamma + beta, pred = sum_c wd[c] y[c]:  dv = rstd dp (gw[c] - mean(gw) - xhat (pred - bw) / E), gw = gamma wd;
            // cwacc[c] += dp xhat[c] feeds d(gamma) and d(decoder_pred.weight) at the end of the CTA
            float wdv[32], gam[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) { wdv[i] = 0.f; gam[i] = 0.f; }
            add_bias32(g.wd, j * BOXC, wdv);
            add_bias32(g.ln_w, j * BOXC, gam);
            const float rd = hl_rstd * dp;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float xh = (v[i] - hl_mean) * hl_rstd;
              cwacc[jj][i] = fmaf(dp, xh, cwacc[jj][i]);
              v[i] = rd * (fmaf(gam[i], wdv[i], -hl_mgw) - xh * hl_s);
            }
          } else if (EPI == EPI_HEAD_BWD) {
            // dh = dpred * wd[c] * leaky'(pre);  dwd[c] += sum_rows dpred * leaky(pre)  (summed per thread across tiles)
            add_bias32(g.bias, n, v);
            float wdv[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) wdv[i] = 0.f;
            add_bias32(g.wd, c0 + j * BOXC, wdv);                          // vector loads of decoder_pred.weight
            if (VAR != 1 || c0 == 0) {
#pragma unroll
              for (int i = 0; i < 32; ++i) cwacc[jj][i] += dp * leaky(v[i]);
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) cwacc_hi[jj][i] += dp * leaky(v[i]);
            }
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = dp * wdv[i] * (v[i] > 0.f ? 1.f : 0.01f);
          }
          store_box_row((two_out ? ob2 : ob) + j * BOX_BYTES, r, v);
        }
        // accumulator and aux tile consumed: release them before the (slower) store path
        tc::fence_before_sync();
        tc::fence_proxy_async();
        __syncwarp();
        if (lane == 0) { if (CG == 2) tc::mbar_arrive_cluster(tempty + buf, 0); else tc::mbar_arrive(tempty + buf); }
        if (tc::elect_one_sync()) {
          if (m0 + q * 32 < g.M) {                            // rows past M are clipped by the tensor map; skip all-out boxes
#pragma unroll
            for (int jj = 0; jj < (CF::NBOX + EPI_GROUPS - 1) / EPI_GROUPS; ++jj) {
              const int j = jgrp + jj * EPI_GROUPS;
              if (j >= CF::NBOX) break;
              if (two_out) tc::tma_store_2d(&maps.out2, ob + j * BOX_BYTES + q * 2048, n0 + j * BOXC, m0 + q * 32);
              tc::tma_store_2d(&maps.out, (two_out ? ob2 : ob) + j * BOX_BYTES + q * 2048, n0 + j * BOXC, m0 + q * 32);
            }
          }
          tc::tma_store_commit();
        }
        __syncwarp();
        obuf ^= 1;
      } else if (EPI == EPI_HEAD) {
        // tile = 128 low-res pixels x 96 expanded channels n' = ij*E + c of one shuffle slot (tulip.py:174-178, 731);
        // this thread sums the whole channel run of its pixel in a fixed order (bitwise reproducible, no cross-warp step)
        const int ij = n0 / g.hd_E, c0 = n0 % g.hd_E;
        float acc = 0.f;
        float2 ln_ms = make_float2(0.f, 0.f);
        if (VAR == 2) {
          // FinalPatchExpanding: the tile's 96 columns are the pixel's whole channel run.  Two passes over the accumulator
          // (mean, then centred squares and the gamma wd dot product), as layernorm_fwd does on rows.
          float s1 = 0.f;
#pragma unroll
          for (int j = 0; j < CF::NBOX; ++j) {
            float v[32];
            tc::tmem_ld32(taddr + j * BOXC, v);
#pragma unroll
            for (int i = 0; i < 32; ++i) s1 += v[i];
          }
          const float mean = s1 * (1.0f / (float)BN);
          float s2 = 0.f;
#pragma unroll
          for (int j = 0; j < CF::NBOX; ++j) {
            float v[32], wdv[32], gam[32];
            tc::tmem_ld32(taddr + j * BOXC, v);
#pragma unroll
            for (int i = 0; i < 32; ++i) { wdv[i] = 0.f; gam[i] = 0.f; }
            add_bias32(g.wd, j * BOXC, wdv);
            add_bias32(g.ln_w, j * BOXC, gam);
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float d = v[i] - mean;
              s2 = fmaf(d, d, s2);
              acc = fmaf(d, gam[i] * wdv[i], acc);
            }
          }
          const float rstd = rsqrtf(s2 * (1.0f / (float)BN) + g.ln_eps);
          acc = fmaf(rstd, acc, hl_bw);
          ln_ms = make_float2(mean, rstd);
        } else {
#pragma unroll
        for (int j = 0; j < CF::NBOX; ++j) {
          float v[32], wdv[32];
          tc::tmem_ld32(taddr + j * BOXC, v);
          add_bias32(g.bias, n0 + j * BOXC, v);
#pragma unroll
          for (int i = 0; i < 32; ++i) wdv[i] = 0.f;
          add_bias32(g.wd, c0 + j * BOXC, wdv);
#pragma unroll
          for (int i = 0; i < 32; ++i) acc = fmaf(wdv[i], leaky(v[i]), acc);
        }
        }
        tc::fence_before_sync();
        __syncwarp();
        if (lane == 0) { if (CG == 2) tc::mbar_arrive_cluster(tempty + buf, 0); else tc::mbar_arrive(tempty + buf); }
        if (m < g.M) {
          const long px = head_pixel(g, m, ij);
          float* p = g.pred + px;
          if (g.hd_E == BN) *p = acc; else atomicAdd(p, acc);
          if (VAR == 2) *reinterpret_cast<float2*>(g.ln_ystats + 2 * px) = ln_ms;
        }
        tphase ^= 1;                                          // this group's buffer completes one phase per tile it handles
        continue;
      } else {
#pragma unroll 1
        for (int ch = 2 * jgrp; ch < BN / 16; ch += 2 * EPI_GROUPS) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            float v[16];
            tc::tmem_ld16(taddr + (ch + h) * 16, v);
            if (m < g.M) epi_direct16<EPI>(g, m, n0 + (ch + h) * 16, v);
          }
        }
        tc::fence_before_sync();
        __syncwarp();
        if (lane == 0) { if (CG == 2) tc::mbar_arrive_cluster(tempty + buf, 0); else tc::mbar_arrive(tempty + buf); }
      }
      if (++buf == NBUF) { buf = 0; tphase ^= 1; }
    }
    if (EPI == EPI_HEAD_BWD && VAR == 2) {
      // d(gamma)[c] = wd[c] A[c],  d(beta)[c] = wd[c] D,  d(wd)[c] = gamma[c] A[c] + beta[c] D  with A[c] = sum dp xhat[c], D = sum dp
      const int copy_off = g.ln_copies > 1 ? (int)(blockIdx.x % g.ln_copies) * g.ln_stride : 0;
#pragma unroll
      for (int jj = 0; jj < (CF::NBOX + EPI_GROUPS - 1) / EPI_GROUPS; ++jj) {
        const int j = jgrp + jj * EPI_GROUPS;
        if (j < CF::NBOX) {
          const float A = warp_colsum32(cwacc[jj], lane);
          const int cc = j * BOXC + lane;
          atomicAdd(g.ln_dw + copy_off + cc, g.wd[cc] * A);
          atomicAdd(g.dwd + copy_off + cc, g.ln_w[cc] * A);
        }
      }
      if (jgrp == 0) {
        const float D = warp_sum(hl_dsum);
        for (int cc = lane; cc < BN; cc += 32) {
          atomicAdd(g.ln_db + copy_off + cc, g.wd[cc] * D);
          atomicAdd(g.dwd + copy_off + cc, g.ln_b[cc] * D);
        }
      }
    } else if (EPI == EPI_HEAD_BWD) {
      // hd_E is BN or 2 BN on this path (checked by the launcher): a tile covers channels c0 .. c0 + 95 with c0 = 0 or 96
      float* dwd = g.dwd + (g.dwd_copies > 1 ? (int)(blockIdx.x % g.dwd_copies) * g.hd_E : 0);
#pragma unroll
      for (int jj = 0; jj < (CF::NBOX + EPI_GROUPS - 1) / EPI_GROUPS; ++jj) {
        const int j = jgrp + jj * EPI_GROUPS;
        if (j < CF::NBOX) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float cw = warp_sum(cwacc[jj][i]);
            if (lane == i) atomicAdd(dwd + j * BOXC + i, cw);
          }
          if (VAR == 1) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float cw = warp_sum(cwacc_hi[jj][i]);
              if (lane == i) atomicAdd(dwd + BN + j * BOXC + i, cw);
            }
          }
        }
      }
    }
    if (CF::TMA_OUT && tc::elect_one_sync()) tc::tma_store_wait<0>();   // global writes complete before the CTA retires
    }
  }
  tc::fence_before_sync();
  if (CG == 2) {
    tc::cluster_sync_all();                                 // neither CTA of a pair retires while the other may still signal it
    if (warp == 1) tc::tmem_dealloc_cg2<TMEM_COLS>(tmem_base);
  } else {
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

bool panel_disabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("TULIP_B200_NO_PANEL");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

// Tile schedule for a BN-wide launch: B-stationary panels when the weight rows of an n-chunk fit next to a useful A ring
// and every worker gets enough row panels to amortise loading them; tile-major round robin otherwise.
Sched make_sched(int bn, int epi, const GemmArgs& g, const Segments& sg) {
  Sched sc;
  memset(&sc, 0, sizeof sc);
  sc.tiles_m = ceil_div(g.M, BM); sc.tiles_n = g.N / bn;
  sc.kb = ceil_div(g.K, BK);
  sc.klast = ceil_div(g.K - BK * (sc.kb - 1), 16);
  sc.n_chunks = 1; sc.npc = sc.tiles_n; sc.nworkers = 1;
  if (panel_disabled() || epi == EPI_DGELU2 || sg.n != 1 || sg.a5d) return sc;
  const int a_bytes = BM * BK * 2, b_bytes = bn * BK * 2;
  const int area = cfg_stages(bn, epi) * (a_bytes + b_bytes);          // Cfg::STAGES * Cfg::STAGE_BYTES
  const int sms = tulip_num_sms();
  for (int nc = 1; nc <= sc.tiles_n; ++nc) {
    if (sc.tiles_n % nc) continue;
    const int npc = sc.tiles_n / nc;
    const int bres = npc * sc.kb * b_bytes;
    if (bres + 3 * a_bytes > area) continue;
    int nsa = (area - bres) / a_bytes;
    if (nsa > NSA_MAX) nsa = NSA_MAX;
    if (npc > 1 && nsa < sc.kb + 1) continue;                            // a panel's A blocks stay resident across its tiles
    const int nworkers = min(sc.tiles_m, sms / nc);
    if (nworkers < 1 || sc.tiles_m < 4 * nworkers) break;               // too few row panels per worker: B loads would not amortise
    sc.panel = 1; sc.n_chunks = nc; sc.npc = npc; sc.nworkers = nworkers; sc.nsa = nsa; sc.bres_bytes = bres;
    break;
  }
  return sc;
}

// Tile width and schedule of one launch.  Wide tiles (BN = 192) where the problem has enough of them to fill the chip (deep
// K: tensor / operand-traffic bound); narrow tiles if only they allow the resident-B schedule (not for GELU: its epilogue wants
// the wide tile); GELU with a saved pre-activation needs both staging buffers, which only the narrow tile has.
Sched choose_tiling(const GemmArgs& g, int epi, const Segments& sg, int* bn_out) {
  const bool head = (epi == EPI_HEAD || epi == EPI_HEAD_BWD);
  int bn = (!head && epi != EPI_DGELU2 && g.N % 192 == 0 && (long)ceil_div(g.M, BM) * (g.N / 192) >= tulip_num_sms()) ? 192 : 96;
  if (epi == EPI_GELU && g.out2 != nullptr) bn = 96;
  if (epi == EPI_LNBWD || epi_ln_fwd(epi)) bn = g.N;      // one tile holds whole rows
  Sched sc = make_sched(bn, epi, g, sg);
  if (bn == 192 && !sc.panel && epi != EPI_GELU && epi != EPI_LNBWD && !epi_ln_fwd(epi)) {
    const Sched s96 = make_sched(96, epi, g, sg);
    if (s96.panel) { bn = 96; sc = s96; }
  }
  *bn_out = bn;
  return sc;
}

// CTA pairs (cta_group::2) for the tile-major launches with a deep K: the pair cuts the bytes an SM pulls per K block from
// A + B to A + B/2.  MEASURED (scripts/time_nt.py, graph replay, B = 32): every eligible launch of the step is 0.1-1.0 us
// SLOWER as a pair (fc1 stage 3: 13.1 -> 14.1 us, fc2 stage 2: 11.9 -> 12.0 us) -- at M <= 8192 rows these launches last
// 5-18 us and are paced by fill / drain and the cluster hand-shakes, not by operand bytes -- so pairs are OFF by default.
// TULIP_B200_CG2=2 enables them for K >= 384, =1 for every eligible launch (parity tests run them this way).
int g_pairs_mode = -1;
bool pairs_wanted(const GemmArgs& g, int epi, const Segments& sg, const Sched& sc, int bn) {
  int& mode = g_pairs_mode;
  if (mode < 0) {
    const char* e = getenv("TULIP_B200_CG2");
    mode = e ? atoi(e) : 0;
  }
  if (mode == 0 || sc.panel || sg.a5d || sg.n != 1) return false;
  if (epi != EPI_STORE && epi != EPI_GELU && epi != EPI_RESID) return false;
  if (epi == EPI_GELU && g.out2 != nullptr) return false;
  if (sc.tiles_m < 2 || (bn != 96 && bn != 192)) return false;
  return mode == 1 || g.K >= 384;
}

template <int BN, int EPI, int VAR = 0>
int launch(const Maps& maps, const GemmArgs& g, const Segments& sg, const Sched& sc, cudaStream_t st) {
  using CF = Cfg<BN, EPI>;
  static bool configured = false;
  if (!configured) {
    TULIP_CUDA(cudaFuncSetAttribute(gemm_nt_tc05_kernel<BN, EPI, VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, CF::TOTAL));
    configured = true;
  }
  const int grid = sc.panel ? sc.n_chunks * sc.nworkers : min(sc.tiles_m * sc.tiles_n, tulip_num_sms());
  tulip_launch(gemm_nt_tc05_kernel<BN, EPI, VAR>, grid, THREADS, CF::TOTAL, st, maps, g, sg, sc);
  TULIP_CHECK_LAUNCH();
  return TULIP_OK;
}

// CTA-pair launch: clusters of two CTAs (one TPC), one pair per 256-row x BN tile in flight
template <int BN, int EPI>
int launch_pairs(const Maps& maps, const GemmArgs& g, const Segments& sg, const Sched& sc, cudaStream_t st) {
  using CF = Cfg<BN, EPI>;
  auto kern = gemm_nt_tc05_kernel<BN, EPI, 0, 2>;
  static bool configured = false;
  if (!configured) {
    TULIP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, CF::TOTAL));
    configured = true;
  }
  const int pair_tiles = ((sc.tiles_m + 1) / 2) * sc.tiles_n;
  const int pairs = min(pair_tiles, tulip_num_sms() / 2);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * pairs); cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = CF::TOTAL; cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = tulip_pdl_enabled() ? 2 : 1;
  (void)cudaLaunchKernelEx(&cfg, kern, maps, g, sg, sc);
  TULIP_CHECK_LAUNCH();
  return TULIP_OK;
}

template <int EPI>
int launch_bn(int bn, const Maps& maps, const GemmArgs& g, const Segments& sg, const Sched& sc, cudaStream_t st) {
  if (sc.cg == 2) {
    if constexpr (EPI == EPI_STORE || EPI == EPI_GELU || EPI == EPI_RESID) {
      if (bn == 192) return launch_pairs<192, EPI>(maps, g, sg, sc, st);
      return launch_pairs<96, EPI>(maps, g, sg, sc, st);
    }
    return TULIP_ERR_UNSUPPORTED;
  }
  if (bn == 192) return launch<192, EPI>(maps, g, sg, sc, st);
  return launch<96, EPI>(maps, g, sg, sc, st);
}

int make_io_map(CUtensorMap* map, const void* base, long ld, int M, int N, int box_rows) {
  // [M, N] bf16 row-major viewed in 32-column boxes of box_rows rows, 64B swizzle (epilogue staging layout): 128-row boxes
  // for the producer's auxiliary-tile loads, 32-row boxes for the per-warp stores
  const uint64_t dims[2] = {(uint64_t)N, (uint64_t)M};
  const uint64_t str[1] = {(uint64_t)ld * 2};
  const uint32_t box[2] = {BOXC, (uint32_t)box_rows};
  return tulip_make_tmap(map, base, 2, dims, str, box, 64);
}

}  // namespace

bool gemm_nt_lnbwd_supported(int M, int N, int K) {
  static int off = -1;
  if (off < 0) {
    const char* e = getenv("TULIP_B200_NO_FUSED_LNBWD");
    off = (e && e[0] == '1') ? 1 : 0;
  }
  return !off && M > 0 && (N == 96 || N == 192) && K % 8 == 0;
}

bool gemm_nt_lnfwd_supported(int M, int N, int K) {
  static int off = -1;
  if (off < 0) {
    const char* e = getenv("TULIP_B200_NO_FUSED_LNFWD");
    off = (e && e[0] == '1') ? 1 : 0;
  }
  return !off && M > 0 && (N == 96 || N == 192) && K % 8 == 0;
}

tulip_tmap_encode_fn tulip_tmap_encoder() {
  static tulip_tmap_encode_fn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<tulip_tmap_encode_fn>(p);
  }
  return fn;
}

int tulip_make_tmap(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box, int swizzle_bytes) {
  tulip_tmap_encode_fn enc = tulip_tmap_encoder();
  TULIP_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t d[5], s[4];
  cuuint32_t b[5], e[5];
  for (int i = 0; i < rank; ++i) { d[i] = dims[i]; b[i] = box[i]; e[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) s[i] = strides_bytes[i];
  const CUtensorMapSwizzle sw = swizzle_bytes == 0 ? CU_TENSOR_MAP_SWIZZLE_NONE
                                : (swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B);
  const CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), d, s, b, e,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    tulip_set_error("cuTensorMapEncodeTiled failed");
    return TULIP_ERR_CUDA;
  }
  return TULIP_OK;
}

int gemm_nt_tc05(const GemmArgs& g, int epi, cudaStream_t st) {
  if (g.N % 96 || g.K % 8 || g.M <= 0) return TULIP_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(g.A) & 15) || (reinterpret_cast<uintptr_t>(g.B) & 15) || (g.lda % 8) || (g.ldb % 8))
    return TULIP_ERR_UNSUPPORTED;
  const bool head = (epi == EPI_HEAD || epi == EPI_HEAD_BWD);
  if (epi == EPI_HEAD_BWD && g.hd_E != 96 && g.hd_E != 192) return TULIP_ERR_UNSUPPORTED;   // dwd accumulators: one or two 96-channel groups
  if (head && g.hd_ln) {                                   // FinalPatchExpanding: a tile must hold the pixel's whole channel run
    if (g.hd_E != 96) return TULIP_ERR_UNSUPPORTED;
    TULIP_REQUIRE(g.ln_w && g.ln_b && g.wd && g.pred && (epi == EPI_HEAD ? g.ln_ystats != nullptr : (g.ln_stats && g.ln_dw && g.ln_db && g.dwd && g.target && g.gscale)),
                  "gemm_nt head (hd_ln): needs gamma, beta, decoder_pred.weight, pred and the statistics / gradient buffers");
  }
  if (epi == EPI_LNBWD) {
    if (!gemm_nt_lnbwd_supported(g.M, g.N, g.K) || g.a_mode != A_PLAIN || g.K1 < g.K) return TULIP_ERR_UNSUPPORTED;
    TULIP_REQUIRE(g.aux && g.ln_w && g.ln_stats && g.ln_dw && g.ln_db && (!g.out2 || g.row_scale),
                  "gemm_nt EPI_LNBWD: needs the LayerNorm input rows, gamma, row statistics and the two gradient accumulators");
  }
  if (epi_ln_fwd(epi)) {
    if (!gemm_nt_lnfwd_supported(g.M, g.N, g.K) || g.a_mode != A_PLAIN) return TULIP_ERR_UNSUPPORTED;
    TULIP_REQUIRE(g.ln_w && g.ln_b && g.ln_y && g.ln_ystats && (epi == EPI_STORE_LN || g.aux),
                  "gemm_nt EPI_STORE_LN / EPI_RESID_LN: needs gamma, beta, the LayerNorm output and its statistics rows");
  }
  int bn = 96;                                            // chosen with the schedule once the K segments are known

  Segments sg;
  memset(&sg, 0, sizeof sg);
  Maps maps;
  memset(&maps, 0, sizeof maps);
  int rc;
  if (g.a_mode == A_UNSHUFFLE) {
    // A[m=(b,h,w), k=ij*Cc+c] = src[b, 2h+i, 2w+j, c]  as a 5-D view (c, j, w, i, bh) of the [B,2H,2W,Cc] tensor
    const int W = g.g_W, Cc = g.g_Cc;
    if (!((W >= BM && W % BM == 0) || (W < BM && BM % W == 0)) || g.M % W || Cc % 8) return TULIP_ERR_UNSUPPORTED;
    const uint64_t dims[5] = {(uint64_t)Cc, 2, (uint64_t)W, 2, (uint64_t)(g.M / W)};
    const uint64_t str[4] = {(uint64_t)Cc * 2, (uint64_t)2 * Cc * 2, (uint64_t)2 * W * Cc * 2, (uint64_t)4 * W * Cc * 2};
    const uint32_t box[5] = {64, 1, (uint32_t)(W >= BM ? BM : W), 1, (uint32_t)(W >= BM ? 1 : BM / W)};
    rc = tulip_make_tmap(&maps.A, g.A, 5, dims, str, box);
    if (rc) return rc;
    maps.A2 = maps.A;
    sg.n = 4; sg.a5d = 1; sg.gW = W;
    for (int s = 0; s < 4; ++s) { sg.len[s] = Cc; sg.bcol[s] = s * Cc; sg.sj[s] = s & 1; sg.si[s] = s >> 1; }
  } else {
    const int K1 = g.K1 < g.K ? g.K1 : g.K;
    const uint64_t dims[2] = {(uint64_t)K1, (uint64_t)g.M};
    const uint64_t str[1] = {(uint64_t)g.lda * 2};
    const uint32_t box[2] = {64, BM};
    rc = tulip_make_tmap(&maps.A, g.A, 2, dims, str, box);
    if (rc) return rc;
    maps.A2 = maps.A;
    sg.n = 1; sg.len[0] = K1; sg.amap[0] = 0; sg.bcol[0] = 0;
    if (K1 < g.K) {
      if ((reinterpret_cast<uintptr_t>(g.A2) & 15) || (g.lda2 % 8)) return TULIP_ERR_UNSUPPORTED;
      const uint64_t dims2[2] = {(uint64_t)(g.K - K1), (uint64_t)g.M};
      const uint64_t str2[1] = {(uint64_t)g.lda2 * 2};
      rc = tulip_make_tmap(&maps.A2, g.A2, 2, dims2, str2, box);
      if (rc) return rc;
      sg.n = 2; sg.len[1] = g.K - K1; sg.amap[1] = 1; sg.bcol[1] = K1;
    }
  }
  Sched sc = choose_tiling(g, epi, sg, &bn);
  sc.cg = pairs_wanted(g, epi, sg, sc, bn) ? 2 : 1;
  sc.hints = tulip_hints();
  {
    const uint64_t dims[2] = {(uint64_t)g.K, (uint64_t)g.N};
    const uint64_t str[1] = {(uint64_t)g.ldb * 2};
    const uint32_t box[2] = {64, (uint32_t)(sc.cg == 2 ? bn / 2 : bn)};   // pairs: each CTA loads half of the B rows
    rc = tulip_make_tmap(&maps.B, g.B, 2, dims, str, box);
    if (rc) return rc;
  }
  maps.B2 = maps.out = maps.out2 = maps.aux = maps.aux2 = maps.B;
  if (epi == EPI_DGELU2) {
    // second product: pre = A2[M,K2] . B2[N,K2]^T feeds accumulator 1
    if (g.a_mode != A_PLAIN || g.K1 < g.K || !g.A2 || !g.B2 || (g.lda2 % 8) || (g.ldb2 % 8) || (g.K2 % 8) ||
        (reinterpret_cast<uintptr_t>(g.A2) & 15) || (reinterpret_cast<uintptr_t>(g.B2) & 15))
      return TULIP_ERR_UNSUPPORTED;
    const uint64_t dimsA[2] = {(uint64_t)g.K2, (uint64_t)g.M};
    const uint64_t strA[1] = {(uint64_t)g.lda2 * 2};
    const uint32_t boxA[2] = {64, BM};
    rc = tulip_make_tmap(&maps.A2, g.A2, 2, dimsA, strA, boxA);
    if (rc) return rc;
    const uint64_t dimsB[2] = {(uint64_t)g.K2, (uint64_t)g.N};
    const uint64_t strB[1] = {(uint64_t)g.ldb2 * 2};
    const uint32_t boxB[2] = {64, (uint32_t)bn};
    rc = tulip_make_tmap(&maps.B2, g.B2, 2, dimsB, strB, boxB);
    if (rc) return rc;
    sg.n = 2; sg.len[1] = g.K2; sg.amap[1] = 1; sg.bmap[1] = 1; sg.bcol[1] = 0; sg.acc[1] = 1;
  }
  if (epi_tma_out(epi)) {
    if ((reinterpret_cast<uintptr_t>(g.out) & 15) || (g.ldo % 8)) return TULIP_ERR_UNSUPPORTED;
    rc = make_io_map(&maps.out, g.out, g.ldo, g.M, g.N, 32);
    if (rc) return rc;
    if ((epi == EPI_GELU || epi == EPI_LNBWD) && g.out2) {
      if ((reinterpret_cast<uintptr_t>(g.out2) & 15) || (g.ldo2 % 8)) return TULIP_ERR_UNSUPPORTED;
      rc = make_io_map(&maps.out2, g.out2, g.ldo2, g.M, g.N, 32);
      if (rc) return rc;
    }
    if (epi_ln_fwd(epi)) {
      if (reinterpret_cast<uintptr_t>(g.ln_y) & 15) return TULIP_ERR_UNSUPPORTED;
      rc = make_io_map(&maps.out2, g.ln_y, g.N, g.M, g.N, 32);
      if (rc) return rc;
    }
    if (epi_has_aux(epi) && epi != EPI_STORE_LN) {
      if (!g.aux || (reinterpret_cast<uintptr_t>(g.aux) & 15) || (g.ldaux % 8)) return TULIP_ERR_UNSUPPORTED;
      rc = make_io_map(&maps.aux, g.aux, g.ldaux, g.M, g.N, 32);
      if (rc) return rc;
    }
    maps.aux2 = maps.aux;
    if (epi == EPI_LNBWD && g.aux2) {
      if ((reinterpret_cast<uintptr_t>(g.aux2) & 15) || (g.ldaux2 % 8)) return TULIP_ERR_UNSUPPORTED;
      rc = make_io_map(&maps.aux2, g.aux2, g.ldaux2, g.M, g.N, 32);
      if (rc) return rc;
    }
  }
  switch (epi) {
    case EPI_STORE: return launch_bn<EPI_STORE>(bn, maps, g, sg, sc, st);
    case EPI_GELU: return launch_bn<EPI_GELU>(bn, maps, g, sg, sc, st);
    case EPI_RESID: return launch_bn<EPI_RESID>(bn, maps, g, sg, sc, st);
    case EPI_PIXSHUF: return launch_bn<EPI_PIXSHUF>(bn, maps, g, sg, sc, st);
    case EPI_SPLIT2: return launch_bn<EPI_SPLIT2>(bn, maps, g, sg, sc, st);
    case EPI_DGELU: return launch_bn<EPI_DGELU>(bn, maps, g, sg, sc, st);
    case EPI_DGELU2: return launch<96, EPI_DGELU2>(maps, g, sg, sc, st);     // two accumulators x two buffers: BN = 96 only
    case EPI_ROWSCALE: return launch_bn<EPI_ROWSCALE>(bn, maps, g, sg, sc, st);
    case EPI_LNBWD: return launch_bn<EPI_LNBWD>(bn, maps, g, sg, sc, st);
    case EPI_STORE_LN: return launch_bn<EPI_STORE_LN>(bn, maps, g, sg, sc, st);
    case EPI_RESID_LN: return launch_bn<EPI_RESID_LN>(bn, maps, g, sg, sc, st);
    case EPI_HEAD: return g.hd_ln ? launch<96, EPI_HEAD, 2>(maps, g, sg, sc, st) : launch<96, EPI_HEAD>(maps, g, sg, sc, st);
    case EPI_HEAD_BWD:
      if (g.hd_ln) return launch<96, EPI_HEAD_BWD, 2>(maps, g, sg, sc, st);
      return g.hd_E == 96 ? launch<96, EPI_HEAD_BWD>(maps, g, sg, sc, st) : launch<96, EPI_HEAD_BWD, 1>(maps, g, sg, sc, st);
  }
  return TULIP_ERR_UNSUPPORTED;
}


int gemm_nt_pairs_mode(int mode) {
  const int prev = g_pairs_mode;
  if (mode >= 0 && mode <= 2) g_pairs_mode = mode;
  return prev;
}

int gemm_nt_tc05_plan(int M, int N, int K, int epi, int save_pre, int* out) {
  // host-side tiling decision for a plain (single-segment) launch, for tests and tooling:
  // out = {bn, schedule (0 tile-major, 1 B-stationary panels, 2 CTA pairs), n_chunks, npc, nworkers, nsa, kb, klast, grid, stages}
  GemmArgs g;
  memset(&g, 0, sizeof g);
  g.M = M; g.N = N; g.K = K; g.K1 = K;
  if (save_pre) g.out2 = reinterpret_cast<bf16*>(16);
  if (N % 96 || K % 8 || M <= 0) return TULIP_ERR_UNSUPPORTED;
  Segments sg;
  memset(&sg, 0, sizeof sg);
  sg.n = 1; sg.len[0] = K;
  int bn = 96;
  const Sched sc = choose_tiling(g, epi, sg, &bn);
  if (pairs_wanted(g, epi, sg, sc, bn)) {                  // schedule 2: CTA pairs on 256-row tiles
    out[0] = bn; out[1] = 2; out[2] = 1; out[3] = sc.tiles_n; out[4] = 1; out[5] = 0; out[6] = sc.kb; out[7] = sc.klast;
    out[8] = 2 * min(((sc.tiles_m + 1) / 2) * sc.tiles_n, tulip_num_sms() / 2);
    out[9] = cfg_stages(bn, epi);
    return TULIP_OK;
  }
  out[0] = bn; out[1] = sc.panel; out[2] = sc.n_chunks; out[3] = sc.npc; out[4] = sc.nworkers; out[5] = sc.nsa; out[6] = sc.kb;
  out[7] = sc.klast; out[8] = sc.panel ? sc.n_chunks * sc.nworkers : min(sc.tiles_m * sc.tiles_n, tulip_num_sms());
  out[9] = cfg_stages(bn, epi);
  return TULIP_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// TN: dW[N,K] += dY[M,N]^T . X[M,K]  (+ db[N] += colsum(dY)) -- weight gradients.
// Both operands are MN-major for the tensor core (the contraction runs over tokens, the strided dimension), which
// tcgen05 takes directly from 128B-swizzled TMA boxes {64 columns, 64 tokens}.  One CTA owns a 128-row x <=192-column
// block of dW in TMEM for a slice of the tokens; the bias gradient rides along as one extra B block whose first column
// is all ones.  Slices are combined with vectorised fp32 reductions (red.global.add.v4.f32) into the flat gradient.
namespace {

constexpr int TN_TOK = 64;                  // tokens per pipeline stage (4 UMMA K-steps)
constexpr int TN_STAGES = 5;                  // 5 x 40 KB ring (4 -> 5: 3-6 % on the HBM-bound launches; two 2-stage CTAs per SM lost 5 %)
constexpr int TN_BOX = 64 * TN_TOK * 2;     // 8 KB
constexpr int TN_STAGE_BYTES = (2 + 3) * TN_BOX;          // A: 2 boxes (128 dW rows), B: up to 3 boxes
constexpr int TN_ONES_OFF = TN_STAGES * TN_STAGE_BYTES;
constexpr int TN_BAR_OFF = TN_ONES_OFF + TN_BOX;
constexpr int TN_SMEM = TN_BAR_OFF + 256 + 1024;
constexpr int TN_THREADS = 256;

// dst[0 .. bytes/4) += src[0 .. bytes/4) as ONE bulk reduction (shared -> global, fp32 add performed at L2): a 32-column run
// of an accumulator row costs one of these instead of eight 16-byte red.global.add (6144 per CTA: the reduction, not the
// mainloop, was the larger part of a deep-stage launch)
__device__ __forceinline__ void bulk_reduce_add_f32(float* gdst, const void* ssrc, int bytes) {
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;"
               ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
}
constexpr int TN_STG_PITCH = 144;             // per-thread staging run: 32 floats, pitch keeps 16-byte stores conflict-free
static_assert(TN_THREADS * TN_STG_PITCH <= TN_STAGE_BYTES, "reduction staging aliases ring stage 0");

__global__ void __launch_bounds__(TN_THREADS, 1)
gemm_tn_tc05_kernel(const __grid_constant__ CUtensorMap mapY, const __grid_constant__ CUtensorMap mapX,
                    const __grid_constant__ CUtensorMap mapX2, const __grid_constant__ GemmTNArgs g, int kcols_per_tile, int tok_blocks_per_split,
                    int row_seg_len, int row_tiles_per_seg, int y5d) {
  extern __shared__ unsigned char smem_raw[];
  pdl_trigger();
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + TN_BAR_OFF);
  uint64_t* empty = full + TN_STAGES;
  uint64_t* done = empty + TN_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
  const int warp = tc::warp_idx_sync(), lane = threadIdx.x & 31;

  // dW rows come in segments of row_seg_len (one segment normally; the 4 shuffle slots of a PixelShuffle-backward
  // gather otherwise), each tiled by 128 rows with TMA zero fill past the segment end
  const int seg = blockIdx.x / row_tiles_per_seg;
  const int c0 = (blockIdx.x % row_tiles_per_seg) * 128;         // first row inside the segment
  const int n0 = seg * row_seg_len + c0;                         // first dW row (permuted order) of this CTA
  const int rows_valid = min(128, row_seg_len - c0);
  const int k0 = blockIdx.y * kcols_per_tile;                    // first dW column
  const bool x2 = k0 >= g.K1;                                    // this column tile reads the second X source (concat)
  const int kvalid = min(kcols_per_tile, g.K - k0);
  const int nbx = (kvalid + 63) / 64;                            // data boxes of X per stage
  const bool with_db = (g.db != nullptr) && blockIdx.y == 0;
  const int nb = nbx + (with_db ? 1 : 0);                        // B blocks per MMA (N_u = 64 * nb)
  const int tb_total = (g.M + TN_TOK - 1) / TN_TOK;
  const int tb_begin = blockIdx.z * tok_blocks_per_split;
  const int tb_end = min(tb_total, tb_begin + tok_blocks_per_split);
  const int ntb = tb_end - tb_begin;
  if (ntb <= 0) { pdl_wait(); return; }

  if (threadIdx.x == 0) {
    for (int s = 0; s < TN_STAGES; ++s) { tc::mbar_init(full + s, 1); tc::mbar_init(empty + s, 1); }
    tc::mbar_init(done, 1);
    tc::fence_barrier_init();
  }
  if (with_db) {
    // ones block: B[k = column 0 of this block, token] = 1, every other column 0 (MN-major, 128B-swizzled rows of 64 columns)
    uint4* ones = reinterpret_cast<uint4*>(smem + TN_ONES_OFF);
    for (int i = threadIdx.x; i < TN_BOX / 16; i += TN_THREADS) {
      const int row = i >> 3, chunk = i & 7;                     // physical 16B chunk `chunk` of token row `row`
      const int logical = chunk ^ (row & 7);
      ones[i] = make_uint4(logical == 0 ? 0x00003F80u : 0u, 0u, 0u, 0u);     // bf16(1.0) in element 0
    }
    tc::fence_proxy_async();
  }
  if (warp == 1) tc::tmem_alloc<256>(tmem_slot);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  if (warp == 0) {
    // TMA producer: whole warp walks the loop, one elected lane issues (see tc::elect_one_sync)
    int stage = 0; uint32_t phase = 0;
    const uint64_t pol_first = tc::l2_policy_evict_first();
    const bool hint = g.hint != 0;                               // operands are streamed once: L2 evict_first
    for (int tb = tb_begin; tb < tb_end; ++tb) {
      tc::mbar_wait(empty + stage, phase ^ 1);
      if (tc::elect_one_sync()) {
        unsigned char* a = smem + stage * TN_STAGE_BYTES;
        tc::mbar_expect_tx(full + stage, (2 + nbx) * TN_BOX);
        if (y5d) {
          // dY'[m=(bh,w), n'=slot*Cc+c] = dOut[bh, 2h+i, 2w+j, c]: 5-D view (c, j, w, i, bh); a 64-token box is a run of w
          // (W >= 64) or 64/W whole rows
          const int t0 = tb * TN_TOK;
          const int bh0 = t0 / g.g_W, w0 = t0 % g.g_W;
          tc::tma_load_5d(a, &mapY, full + stage, c0, seg & 1, w0, seg >> 1, bh0);
          tc::tma_load_5d(a + TN_BOX, &mapY, full + stage, c0 + 64, seg & 1, w0, seg >> 1, bh0);
        } else {
          tc::tma_load_2d_pol(a, &mapY, full + stage, n0, tb * TN_TOK, pol_first, hint);
          tc::tma_load_2d_pol(a + TN_BOX, &mapY, full + stage, n0 + 64, tb * TN_TOK, pol_first, hint);
        }
        for (int j = 0; j < nbx; ++j)
          tc::tma_load_2d_pol(a + (2 + j) * TN_BOX, x2 ? &mapX2 : &mapX, full + stage, (x2 ? k0 - g.K1 : k0) + 64 * j, tb * TN_TOK, pol_first, hint);
      }
      __syncwarp();
      if (++stage == TN_STAGES) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    const uint32_t idesc_x = tc::make_idesc(128, 64 * nbx, 1, 1);
    const uint32_t idesc_1 = tc::make_idesc(128, 64, 1, 1);
    const uint64_t d_ones = tc::make_desc_mnmajor_sw128(smem + TN_ONES_OFF, TN_BOX);
    const uint64_t d0 = tc::make_desc_mnmajor_sw128(smem, TN_BOX);
    constexpr uint32_t S16 = TN_STAGE_BYTES >> 4, B16 = (2 * TN_BOX) >> 4;
    uint64_t da = d0;                                         // dY boxes of `stage`; the X boxes follow B16 further
    int stage = 0; uint32_t phase = 0;
    for (int t = 0; t < ntb; ++t) {
      tc::mbar_wait(full + stage, phase);
      tc::fence_after_sync();
      if (tc::elect_one_sync()) {
#pragma unroll
        for (int k = 0; k < TN_TOK / 16; ++k) {                   // 16 tokens = 2 groups of 8 rows = 2048 B per K-step
          const uint32_t acc = (t | k) ? 1u : 0u;
          tc::umma_bf16(tmem_base, da + 128 * k, da + B16 + 128 * k, idesc_x, acc);
          if (with_db) tc::umma_bf16(tmem_base + 64 * nbx, da + 128 * k, d_ones + 128 * k, idesc_1, acc);
        }
        tc::umma_commit(empty + stage);
      }
      __syncwarp();
      if (++stage == TN_STAGES) { stage = 0; phase ^= 1; da = d0; } else { da += S16; }
    }
    if (tc::elect_one_sync()) tc::umma_commit(done);
    __syncwarp();
  }
  // epilogue: every warp drains its TMEM lane quarter; the two warps of a quarter split the columns
  __syncwarp();
  tc::mbar_wait(done, 0);
  tc::fence_after_sync();
  {
    const int q = warp & 3, half = warp >> 2;
    const int n = n0 + q * 32 + lane;
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    const bool row_ok = (q * 32 + lane) < rows_valid && n < g.N;
    const int row = (g.perm_R2 > 1) ? (n % g.perm_Cc) * g.perm_R2 + n / g.perm_Cc : n;
    float* drow = g.dW + (long)row * g.lddw + k0;
    const int nchunks = (kvalid + 31) / 32;
    // every MMA has completed (done barrier): the operand ring is free and stage 0 serves as the reduction staging area
    unsigned char* const stg = smem + threadIdx.x * TN_STG_PITCH;
    for (int ch = half; ch < nchunks; ch += 2) {
      float v[32];
      tc::tmem_ld32(taddr + ch * 32, v);
      if (row_ok) {
        tc::tma_store_wait_read<0>();                         // this thread's previous run has left its staging slot
#pragma unroll
        for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(stg + 4 * i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        tc::fence_proxy_async();
        bulk_reduce_add_f32(drow + ch * 32, stg, 4 * min(32, kvalid - ch * 32));
        tc::tma_store_commit();
      }
    }
    tc::tma_store_wait<0>();
    if (with_db && half == 1) {
      float v[16];
      tc::tmem_ld16(taddr + 64 * nbx, v);
      if (row_ok) atomicAdd(g.db + row, v[0]);
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc<256>(tmem_base);
}

}  // namespace

int gemm_tn_tc05(const GemmTNArgs& g, cudaStream_t st) {
  if (g.M <= 0) return TULIP_ERR_UNSUPPORTED;
  if (g.N % 8 || g.K % 8 || (g.ldy % 8) || (g.ldx % 8) || (g.lddw % 4) || (g.K % 4)) return TULIP_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(g.dY) & 15) || (reinterpret_cast<uintptr_t>(g.X) & 15) || (reinterpret_cast<uintptr_t>(g.dW) & 15))
    return TULIP_ERR_UNSUPPORTED;
  const bool concat = g.K1 < g.K;
  int kcols = 192;
  if (concat) {
    if (!g.X2 || (reinterpret_cast<uintptr_t>(g.X2) & 15) || (g.ldx2 % 8) || g.K1 % 8) return TULIP_ERR_UNSUPPORTED;
    if (g.K1 % 192) kcols = (g.K1 % 96 == 0) ? 96 : 0;        // column tiles must not straddle the two sources
    if (!kcols) return TULIP_ERR_UNSUPPORTED;
  }
  CUtensorMap mY, mX, mX2;
  int row_seg_len = g.N, y5d = 0;
  if (g.y_mode == A_UNSHUFFLE) {
    const int W = g.g_W, Cc = g.g_Cc;
    if (!((W >= TN_TOK && W % TN_TOK == 0) || (W < TN_TOK && TN_TOK % W == 0)) || g.M % W || Cc % 8 || g.N != 4 * Cc) return TULIP_ERR_UNSUPPORTED;
    const uint64_t dims[5] = {(uint64_t)Cc, 2, (uint64_t)W, 2, (uint64_t)(g.M / W)};
    const uint64_t str[4] = {(uint64_t)Cc * 2, (uint64_t)2 * Cc * 2, (uint64_t)2 * W * Cc * 2, (uint64_t)4 * W * Cc * 2};
    const uint32_t box[5] = {64, 1, (uint32_t)(W >= TN_TOK ? TN_TOK : W), 1, (uint32_t)(W >= TN_TOK ? 1 : TN_TOK / W)};
    int rc = tulip_make_tmap(&mY, g.dY, 5, dims, str, box);
    if (rc) return rc;
    row_seg_len = Cc; y5d = 1;
  } else {
    const uint64_t dims[2] = {(uint64_t)g.N, (uint64_t)g.M};
    const uint64_t str[1] = {(uint64_t)g.ldy * 2};
    const uint32_t box[2] = {64, TN_TOK};
    int rc = tulip_make_tmap(&mY, g.dY, 2, dims, str, box);
    if (rc) return rc;
  }
  {
    const uint64_t dims[2] = {(uint64_t)(concat ? g.K1 : g.K), (uint64_t)g.M};
    const uint64_t str[1] = {(uint64_t)g.ldx * 2};
    const uint32_t box[2] = {64, TN_TOK};
    int rc = tulip_make_tmap(&mX, g.X, 2, dims, str, box);
    if (rc) return rc;
    mX2 = mX;
    if (concat) {
      const uint64_t dims2[2] = {(uint64_t)(g.K - g.K1), (uint64_t)g.M};
      const uint64_t str2[1] = {(uint64_t)g.ldx2 * 2};
      rc = tulip_make_tmap(&mX2, g.X2, 2, dims2, str2, box);
      if (rc) return rc;
    }
  }
  static bool configured = false;
  if (!configured) {
    TULIP_CUDA(cudaFuncSetAttribute(gemm_tn_tc05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TN_SMEM));
    configured = true;
  }
  const int nseg = g.N / row_seg_len;
  const int row_tiles_per_seg = ceil_div(row_seg_len, 128);
  const int n_tiles = nseg * row_tiles_per_seg, k_tiles = ceil_div(g.K, kcols);
  const int tb_total = ceil_div(g.M, TN_TOK);
  // Token splits: every CTA streams its [tokens, 128 + kcols] operand slice over its own SM<->L2 link, so the launch takes
  // waves(s) / s of the one-split time.  Pick the s that minimises that (ties: fewer splits = fewer atomics); a split keeps at
  // least four token blocks so the ring fills.
  const int tiles = n_tiles * k_tiles, sms = tulip_num_sms();
  int splits = 1;
  {
    double best = 1e30;
    const int smax = max(1, min(tb_total / 4, 4 * sms / tiles + 1));
    for (int sp = 1; sp <= smax; ++sp) {
      const int per_ = ceil_div(tb_total, sp);
      const int real = ceil_div(tb_total, per_);
      const double cost = (double)ceil_div(tiles * real, sms) * per_ + 2.0 * ceil_div(tiles * real, sms);   // + a fixed cost per wave
      if (cost < best - 1e-9) { best = cost; splits = real; }
    }
  }
  const int per = ceil_div(tb_total, splits);
  splits = ceil_div(tb_total, per);
  dim3 grid(n_tiles, k_tiles, splits);
  GemmTNArgs gh = g;
  gh.hint = (tulip_hints() & 1) ? 1 : 0;
  tulip_launch(gemm_tn_tc05_kernel, grid, TN_THREADS, TN_SMEM, st, mY, mX, mX2, gh, kcols, per, row_seg_len, row_tiles_per_seg, y5d);
  TULIP_CHECK_LAUNCH();
  return TULIP_OK;
}
