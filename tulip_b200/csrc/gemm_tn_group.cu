// Grouped, persistent weight-gradient GEMM for sm_100a:  dW_p[N,K] += dY_p[M,N]^T . X_p[M,K]  (+ db_p += colsum(dY_p))
// for up to four independent problems p in ONE launch.
//
// The stand-alone kernel (gemm_tn_tc05_kernel) gives every (row tile, column tile, token split) its own CTA: set-up, ring fill,
// the fp32 reduction of the accumulator into the flat gradient and the tail are paid once per CTA and nothing overlaps them,
// which held the HBM-bound launches of stages 0-1 at 55-65 % of the copy peak and made the deep-stage launches (3-6 us of
// mainloop) mostly fixed cost.  Here one CTA per SM walks a list of work items -- (problem, 128-row tile of dW, <=192-column
// tile, token range) dealt round robin -- with
//   warp 0      TMA producer: dY boxes {64 cols, 64 tokens} x 2 and X boxes x <=3 per 64-token block, 4-stage ring that runs
//               straight through item boundaries (the ring never drains between items)
//   warp 1      tcgen05.mma issuer: both operands MN-major (the contraction runs over tokens), fp32 accumulator in one of TWO
//               256-column TMEM buffers; the bias gradient rides along as an extra B block whose first column is ones
//   warps 2..9  reduction of item i while item i+1 accumulates: tcgen05.ld -> the thread's own 128-byte staging run ->
//               cp.reduce.async.bulk (.add.f32) into the flat gradient, one bulk reduction per 32-column run of a row
// The weight-gradient GEMMs of one Swin half-block (fc2 + fc1, proj + qkv) are leaves of the backward graph and go out
// together (net.cu), so one launch carries 2x the work, token ranges are cut so the items of all problems cost the same, and
// the item count is chosen against the round-robin makespan over the SMs.
#include "gemm.cuh"
#include "tc05.cuh"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

namespace {

constexpr int G_TOK = 64;                        // tokens per pipeline stage (4 UMMA K steps)
constexpr int G_STAGES = 4;
constexpr int G_BOX = 64 * G_TOK * 2;            // 8 KB: {64 columns, 64 tokens} bf16
constexpr int G_STAGE_BYTES = (2 + 3) * G_BOX;   // dY: 2 boxes (128 dW rows), X: up to 3 boxes (192 dW columns)
constexpr int G_ONES_OFF = G_STAGES * G_STAGE_BYTES;
constexpr int G_EPI_WARPS = 8;
constexpr int G_THREADS = 64 + 32 * G_EPI_WARPS;
// reduction staging: one 32-float run per reduction thread (pitch 144 B: conflict-free 16-byte stores, 16-byte aligned rows)
constexpr int G_STG_PITCH = 144;
constexpr int G_STG_OFF = G_ONES_OFF + G_BOX;
constexpr int G_BAR_OFF = G_STG_OFF + 32 * G_EPI_WARPS * G_STG_PITCH;
constexpr int G_SMEM = G_BAR_OFF + 256 + 1024;
constexpr int G_KCOLS = 192;
constexpr int G_BUF_COLS = 256;                  // accumulator buffer: 192 data columns + 64 for the ones block
static_assert(G_SMEM <= 227 * 1024, "shared memory budget");

struct TNProb {
  float* dW; float* db;
  int N, K, lddw;
  int perm_R2, perm_Cc;
  int k_tiles, tiles;      // column tiles, row tiles x column tiles
  int per, tb_total;       // 64-token blocks per item, in total
  int item0;               // index of this problem's first work item
};
struct TNGroupArgs {
  int np, items;
  int x_hint, y_hint;      // L2 eviction priority of the X / dY loads: 0 none, 1 evict_first, 2 evict_last
  TNProb p[TN_GROUP_MAX];
};
struct TNGroupMaps {
  CUtensorMap Y[TN_GROUP_MAX], X[TN_GROUP_MAX];
};

struct Item {
  int p, n0, rows_valid, k0, kvalid, nbx, with_db, tb_begin, ntb;
};
// items of a problem: token range slowest, (row tile, column tile) fastest -- CTAs that run side by side read the same tokens
__device__ __forceinline__ Item decode_item(const TNGroupArgs& a, int idx) {
  Item it;
  int p = 0;
#pragma unroll
  for (int i = 1; i < TN_GROUP_MAX; ++i)
    if (i < a.np && idx >= a.p[i].item0) p = i;
  const TNProb& P = a.p[p];
  const int local = idx - P.item0;
  const int split = local / P.tiles, tile = local - split * P.tiles;
  const int nt = tile / P.k_tiles, kt = tile - nt * P.k_tiles;
  it.p = p;
  it.n0 = nt * 128;
  it.rows_valid = min(128, P.N - it.n0);
  it.k0 = kt * G_KCOLS;
  it.kvalid = min(G_KCOLS, P.K - it.k0);
  it.nbx = (it.kvalid + 63) >> 6;
  it.with_db = (P.db != nullptr && kt == 0) ? 1 : 0;
  it.tb_begin = split * P.per;
  it.ntb = min(P.tb_total, it.tb_begin + P.per) - it.tb_begin;
  return it;
}

// dst[0 .. bytes/4) += src[0 .. bytes/4) as ONE bulk reduction (shared -> global, fp32 add performed at L2); bulk-group
// completion like a TMA store.  A row run of the accumulator costs one of these instead of eight 16-byte red.global.add.
__device__ __forceinline__ void bulk_reduce_add_f32(float* gdst, const void* ssrc, int bytes) {
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;"
               ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
}

__global__ void __launch_bounds__(G_THREADS, 1)
gemm_tn_group_kernel(const __grid_constant__ TNGroupMaps maps, const __grid_constant__ TNGroupArgs a) {
  extern __shared__ unsigned char smem_raw[];
  pdl_trigger();
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + G_BAR_OFF);
  uint64_t* empty = full + G_STAGES;
  uint64_t* tfull = empty + G_STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  const int warp = tc::warp_idx_sync(), lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < G_STAGES; ++s) { tc::mbar_init(full + s, 1); tc::mbar_init(empty + s, 1); }
    for (int b = 0; b < 2; ++b) { tc::mbar_init(tfull + b, 1); tc::mbar_init(tempty + b, G_EPI_WARPS); }
    tc::fence_barrier_init();
  }
  {
    // ones block: B[k = column 0, token] = 1, every other column 0 (MN-major, 128B-swizzled rows of 64 columns)
    uint4* ones = reinterpret_cast<uint4*>(smem + G_ONES_OFF);
    for (int i = threadIdx.x; i < G_BOX / 16; i += G_THREADS) {
      const int row = i >> 3, chunk = i & 7;                     // physical 16B chunk `chunk` of token row `row`
      const int logical = chunk ^ (row & 7);
      ones[i] = make_uint4(logical == 0 ? 0x00003F80u : 0u, 0u, 0u, 0u);     // bf16(1.0) in element 0
    }
    tc::fence_proxy_async();
  }
  if (warp == 1) tc::tmem_alloc<512>(tmem_slot);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  if (warp == 0) {
    // ---- TMA producer: the whole warp walks the loops, one elected lane issues ----
    // Items are dealt statically (item i -> CTA i mod grid).  A global work counter was tried (atomic fetched one item ahead,
    // published to the other warps through a shared-memory ring): 2-10 % slower per launch and 1.5 % slower in the step.
    int stage = 0; uint32_t phase = 0;
    const uint64_t pol_first = tc::l2_policy_evict_first(), pol_last = tc::l2_policy_evict_last();
    const uint64_t polX = a.x_hint == 1 ? pol_first : pol_last, polY = a.y_hint == 1 ? pol_first : pol_last;
    for (int idx = blockIdx.x; idx < a.items; idx += gridDim.x) {
      const Item it = decode_item(a, idx);
      const CUtensorMap* mY = &maps.Y[it.p];
      const CUtensorMap* mX = &maps.X[it.p];
      for (int tb = it.tb_begin; tb < it.tb_begin + it.ntb; ++tb) {
        tc::mbar_wait(empty + stage, phase ^ 1);
        if (tc::elect_one_sync()) {
          unsigned char* s = smem + stage * G_STAGE_BYTES;
          tc::mbar_expect_tx(full + stage, (2 + it.nbx) * G_BOX);
          if (a.y_hint) {
            tc::tma_load_2d_hint(s, mY, full + stage, it.n0, tb * G_TOK, polY);
            tc::tma_load_2d_hint(s + G_BOX, mY, full + stage, it.n0 + 64, tb * G_TOK, polY);
          } else {
            tc::tma_load_2d(s, mY, full + stage, it.n0, tb * G_TOK);
            tc::tma_load_2d(s + G_BOX, mY, full + stage, it.n0 + 64, tb * G_TOK);
          }
          for (int j = 0; j < it.nbx; ++j) {
            if (a.x_hint) tc::tma_load_2d_hint(s + (2 + j) * G_BOX, mX, full + stage, it.k0 + 64 * j, tb * G_TOK, polX);
            else tc::tma_load_2d(s + (2 + j) * G_BOX, mX, full + stage, it.k0 + 64 * j, tb * G_TOK);
          }
        }
        __syncwarp();
        if (++stage == G_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer ----
    const uint32_t idesc_1 = tc::make_idesc(128, 64, 1, 1);
    const uint64_t d_ones = tc::make_desc_mnmajor_sw128(smem + G_ONES_OFF, G_BOX);
    const uint64_t d0 = tc::make_desc_mnmajor_sw128(smem, G_BOX);
    constexpr uint32_t S16 = G_STAGE_BYTES >> 4, B16 = (2 * G_BOX) >> 4;
    uint64_t da = d0;                                         // dY boxes of `stage`; its X boxes follow B16 further
    int stage = 0; uint32_t phase = 0;
    int buf = 0; uint32_t tphase = 0;
    for (int idx = blockIdx.x; idx < a.items; idx += gridDim.x) {
      const Item it = decode_item(a, idx);
      const uint32_t idesc_x = tc::make_idesc(128, 64 * it.nbx, 1, 1);
      const uint32_t tmem_d = tmem_base + buf * G_BUF_COLS;
      const uint32_t tmem_b = tmem_d + 64 * it.nbx;
      tc::mbar_wait(tempty + buf, tphase ^ 1);                // the reduction warps have drained this buffer
      tc::fence_after_sync();
      for (int t = 0; t < it.ntb; ++t) {
        tc::mbar_wait(full + stage, phase);
        tc::fence_after_sync();
        if (tc::elect_one_sync()) {
#pragma unroll
          for (int k = 0; k < G_TOK / 16; ++k) {                  // 16 tokens = 2 groups of 8 rows = 2048 B per K step
            const uint32_t acc = (t | k) ? 1u : 0u;
            tc::umma_bf16(tmem_d, da + 128 * k, da + B16 + 128 * k, idesc_x, acc);
            if (it.with_db) tc::umma_bf16(tmem_b, da + 128 * k, d_ones + 128 * k, idesc_1, acc);
          }
          tc::umma_commit(empty + stage);
        }
        __syncwarp();
        if (++stage == G_STAGES) { stage = 0; phase ^= 1; da = d0; } else { da += S16; }
      }
      if (tc::elect_one_sync()) tc::umma_commit(tfull + buf);
      __syncwarp();
      if (++buf == 2) { buf = 0; tphase ^= 1; }
    }
  } else {
    // ---- reduction warps: two per TMEM lane quarter, splitting the 32-column chunks ----
    const int q = warp & 3, half = (warp - 2) >> 2;
    unsigned char* const stg = smem + G_STG_OFF + (threadIdx.x - 64) * G_STG_PITCH;
    int buf = 0; uint32_t tphase = 0;
    for (int idx = blockIdx.x; idx < a.items; idx += gridDim.x) {
      const Item it = decode_item(a, idx);
      const TNProb& P = a.p[it.p];
      const int n = it.n0 + q * 32 + lane;
      const bool row_ok = (q * 32 + lane) < it.rows_valid;
      const int row = (P.perm_R2 > 1) ? (n % P.perm_Cc) * P.perm_R2 + n / P.perm_Cc : n;
      float* drow = P.dW + (long)row * P.lddw + it.k0;
      const int nchunks = (it.kvalid + 31) >> 5;
      tc::mbar_wait(tfull + buf, tphase);
      tc::fence_after_sync();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * G_BUF_COLS;
      for (int ch = half; ch < nchunks; ch += 2) {
        float v[32];
        tc::tmem_ld32(taddr + ch * 32, v);
        if (row_ok) {
          tc::tma_store_wait_read<0>();                       // this thread's previous run has left the staging slot
#pragma unroll
          for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(stg + 4 * i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
          tc::fence_proxy_async();
          bulk_reduce_add_f32(drow + ch * 32, stg, 4 * min(32, it.kvalid - ch * 32));
          tc::tma_store_commit();
        }
      }
      if (it.with_db && half == 1) {
        float v[16];
        tc::tmem_ld16(taddr + 64 * it.nbx, v);
        if (row_ok) atomicAdd(P.db + row, v[0]);
      }
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(tempty + buf);
      if (++buf == 2) { buf = 0; tphase ^= 1; }
    }
    tc::tma_store_wait<0>();                                  // every reduction of this thread has been performed
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc<512>(tmem_base);
}

bool group_disabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("TULIP_B200_NO_TN_GROUP");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

}  // namespace

bool gemm_tn_groupable(const GemmTNArgs& g) {
  if (group_disabled()) return false;
  if (g.y_mode != A_PLAIN || g.K1 < g.K || g.M <= 0) return false;
  if (g.N % 8 || g.K % 8 || (g.ldy % 8) || (g.ldx % 8) || (g.lddw % 4) || (g.K % 4)) return false;
  if ((reinterpret_cast<uintptr_t>(g.dY) & 15) || (reinterpret_cast<uintptr_t>(g.X) & 15) || (reinterpret_cast<uintptr_t>(g.dW) & 15))
    return false;
  return true;
}

// Token blocks per work item for every problem of a group.  Cost unit: one 8 KB operand box pulled into shared memory.
// For each candidate number of items per SM the token ranges are cut so that items of all problems cost about the same, the
// items are dealt round robin exactly as the kernel deals them, and the longest SM decides.
static int gemm_tn_group_plan_uncached(const int* M, const int* N, const int* K, int n, int sms, int* per_out, int* items_out);

int gemm_tn_group_plan(const int* M, const int* N, const int* K, int n, int sms, int* per_out, int* items_out) {
  if (n < 1 || n > TN_GROUP_MAX) return TULIP_ERR_ARG;
  // the step issues the same few dozen groups every call: remember their plans
  struct Entry { int key[3 * TN_GROUP_MAX + 2]; int per[TN_GROUP_MAX]; int items; };
  static std::vector<Entry> cache;
  static std::mutex mu;
  Entry e;
  memset(&e, 0, sizeof e);
  e.key[0] = n; e.key[1] = sms;
  for (int p = 0; p < n; ++p) { e.key[2 + 3 * p] = M[p]; e.key[3 + 3 * p] = N[p]; e.key[4 + 3 * p] = K[p]; }
  {
    std::lock_guard<std::mutex> lock(mu);
    for (const Entry& c : cache)
      if (memcmp(c.key, e.key, sizeof e.key) == 0) {
        for (int p = 0; p < n; ++p) per_out[p] = c.per[p];
        if (items_out) *items_out = c.items;
        return TULIP_OK;
      }
  }
  const int rc = gemm_tn_group_plan_uncached(M, N, K, n, sms, e.per, &e.items);
  if (rc) return rc;
  for (int p = 0; p < n; ++p) per_out[p] = e.per[p];
  if (items_out) *items_out = e.items;
  std::lock_guard<std::mutex> lock(mu);
  if (cache.size() < 4096) cache.push_back(e);
  return TULIP_OK;
}

static int gemm_tn_group_plan_uncached(const int* M, const int* N, const int* K, int n, int sms, int* per_out, int* items_out) {
  // per-item cost: barrier round trips + the part of the fp32 reduction (<= 96 KB of red.global.add per item, all CTAs into
  // the same few hundred KB of L2) that the next item's mainloop does not hide.  TULIP_B200_TN_ITEM_COST overrides (tuning).
  static double ITEM_COST = -1.0;
  if (ITEM_COST < 0.0) {
    const char* e = getenv("TULIP_B200_TN_ITEM_COST");
    ITEM_COST = (e && atof(e) > 0.0) ? atof(e) : 32.0;     // measured optimum of the bench step (8 / 32 / 64 / 128 / 256)
  }
  constexpr int MIN_PER = 4;                     // a token range keeps the ring busy
  int tbt[TN_GROUP_MAX], ntile[TN_GROUP_MAX], ktile[TN_GROUP_MAX];
  double boxes_sum[TN_GROUP_MAX], total = 0.0;
  for (int p = 0; p < n; ++p) {
    if (M[p] <= 0 || N[p] <= 0 || K[p] <= 0) return TULIP_ERR_ARG;
    tbt[p] = ceil_div(M[p], G_TOK);
    ntile[p] = ceil_div(N[p], 128);
    ktile[p] = ceil_div(K[p], G_KCOLS);
    boxes_sum[p] = 0.0;
    for (int kt = 0; kt < ktile[p]; ++kt) boxes_sum[p] += 2 + ceil_div(std::min(G_KCOLS, K[p] - kt * G_KCOLS), 64);
    boxes_sum[p] *= ntile[p];
    total += boxes_sum[p] * tbt[p];
  }
  double best = 1e300;
  std::vector<double> load(sms);
  for (int w4 = 2; w4 <= 64; ++w4) {                       // candidate items per SM: 0.5, 0.75, ... 16
    const double target = total / ((double)sms * 0.25 * w4);
    int per[TN_GROUP_MAX];
    long items = 0;
    for (int p = 0; p < n; ++p) {
      const double avg = boxes_sum[p] / (ntile[p] * ktile[p]);
      int pp = (int)(target / avg + 0.5);
      pp = std::max(std::min(MIN_PER, tbt[p]), std::min(pp, tbt[p]));
      const int splits = ceil_div(tbt[p], pp);
      per[p] = ceil_div(tbt[p], splits);
      items += (long)ceil_div(tbt[p], per[p]) * ntile[p] * ktile[p];
    }
    const int grid = (int)std::min<long>(items, sms);
    std::fill(load.begin(), load.end(), 0.0);
    long idx = 0;
    for (int p = 0; p < n; ++p) {
      const int splits = ceil_div(tbt[p], per[p]);
      for (int s = 0; s < splits; ++s) {
        const int ntb = std::min(tbt[p], (s + 1) * per[p]) - s * per[p];
        for (int nt = 0; nt < ntile[p]; ++nt)
          for (int kt = 0; kt < ktile[p]; ++kt, ++idx)
            load[idx % grid] += ntb * (2.0 + ceil_div(std::min(G_KCOLS, K[p] - kt * G_KCOLS), 64)) + ITEM_COST;
      }
    }
    double mk = 0.0;
    for (int i = 0; i < grid; ++i) mk = std::max(mk, load[i]);
    if (mk < best - 1e-9) {
      best = mk;
      for (int p = 0; p < n; ++p) per_out[p] = per[p];
      if (items_out) *items_out = (int)items;
    }
  }
  return TULIP_OK;
}

int gemm_tn_group(const GemmTNArgs* gs, int n, cudaStream_t st) {
  if (n < 1 || n > TN_GROUP_MAX) { tulip_set_error("gemm_tn_group: 1..4 problems per launch"); return TULIP_ERR_ARG; }
  TNGroupMaps maps;
  TNGroupArgs a;
  memset(&maps, 0, sizeof maps);
  memset(&a, 0, sizeof a);
  int Ms[TN_GROUP_MAX], Ns[TN_GROUP_MAX], Ks[TN_GROUP_MAX], per[TN_GROUP_MAX];
  for (int p = 0; p < n; ++p) {
    const GemmTNArgs& g = gs[p];
    if (!gemm_tn_groupable(g)) return TULIP_ERR_UNSUPPORTED;
    Ms[p] = g.M; Ns[p] = g.N; Ks[p] = g.K;
    {
      const uint64_t dims[2] = {(uint64_t)g.N, (uint64_t)g.M};
      const uint64_t str[1] = {(uint64_t)g.ldy * 2};
      const uint32_t box[2] = {64, G_TOK};
      const int rc = tulip_make_tmap(&maps.Y[p], g.dY, 2, dims, str, box);
      if (rc) return rc;
    }
    {
      const uint64_t dims[2] = {(uint64_t)g.K, (uint64_t)g.M};
      const uint64_t str[1] = {(uint64_t)g.ldx * 2};
      const uint32_t box[2] = {64, G_TOK};
      const int rc = tulip_make_tmap(&maps.X[p], g.X, 2, dims, str, box);
      if (rc) return rc;
    }
  }
  const int sms = tulip_num_sms();
  int rc = gemm_tn_group_plan(Ms, Ns, Ks, n, sms, per, nullptr);
  if (rc) return rc;
  a.np = n;
  int items = 0;
  for (int p = 0; p < n; ++p) {
    const GemmTNArgs& g = gs[p];
    TNProb& P = a.p[p];
    P.dW = g.dW; P.db = g.db; P.N = g.N; P.K = g.K; P.lddw = (int)g.lddw;
    P.perm_R2 = g.perm_R2 > 1 ? g.perm_R2 : 1; P.perm_Cc = g.perm_R2 > 1 ? g.perm_Cc : 1;
    P.k_tiles = ceil_div(g.K, G_KCOLS);
    P.tiles = ceil_div(g.N, 128) * P.k_tiles;
    P.tb_total = ceil_div(g.M, G_TOK);
    P.per = per[p];
    P.item0 = items;
    items += ceil_div(P.tb_total, P.per) * P.tiles;
  }
  a.items = items;
  // both operands are streamed once by this kernel: evict_first keeps them from pushing the dX chain's tensors out of L2
  // (measured: step 4.92 -> 4.84 ms; X first + dY last: no gain)
  a.x_hint = a.y_hint = (tulip_hints() & 1) ? 1 : 0;
  static bool configured = false;
  if (!configured) {
    TULIP_CUDA(cudaFuncSetAttribute(gemm_tn_group_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, G_SMEM));
    configured = true;
  }
  const int grid = std::min(items, sms);
  tulip_launch(gemm_tn_group_kernel, grid, G_THREADS, G_SMEM, st, maps, a);
  TULIP_CHECK_LAUNCH();
  return TULIP_OK;
}
