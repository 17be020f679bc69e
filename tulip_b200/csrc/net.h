// Whole-network executor: plans the activation arena, repacks the weights and launches every kernel
// of TULIP.forward (tulip/model/tulip.py:702-737) and of its autograd backward
// from C++, so a training step costs one host call per direction instead of ~11.6k ATen dispatches.
#pragma once
#include <string>
#include <vector>

#include "../../include/tulip_b200.h"
#include "gemm.cuh"
#include "kernels.h"

// kernel-function tags for the built-in per-launch profiler (bench.py roofline section)
enum KTag : int {
  K_MISC = 0, K_NT_STORE, K_NT_GELU, K_NT_RESID, K_NT_PIXSHUF, K_NT_SPLIT2, K_NT_DGELU, K_NT_HEAD, K_NT_HEAD_BWD, K_NT_ROWSCALE,
  K_NT_UNSHUFFLE, K_TN, K_TN_UNSHUFFLE, K_ATTN_FWD, K_ATTN_BWD, K_LN_FWD, K_LN_BWD, K_EMBED_FWD, K_EMBED_BWD, K_PACK,
  K_ELEMWISE, K_LOSS, K_WMSA_FWD, K_MLP_FWD, K_NT_LNBWD, K_NT_STORE_LN, K_NT_RESID_LN, K_COUNT
};
const char* ktag_name(int tag);
inline int nt_tag(int epi) {
  switch (epi) {
    case EPI_STORE: return K_NT_STORE; case EPI_GELU: return K_NT_GELU; case EPI_RESID: return K_NT_RESID;
    case EPI_PIXSHUF: return K_NT_PIXSHUF; case EPI_SPLIT2: return K_NT_SPLIT2; case EPI_DGELU: return K_NT_DGELU;
    case EPI_HEAD: return K_NT_HEAD; case EPI_HEAD_BWD: return K_NT_HEAD_BWD; case EPI_ROWSCALE: return K_NT_ROWSCALE;
    case EPI_DGELU2: return K_NT_DGELU; case EPI_LNBWD: return K_NT_LNBWD;
    case EPI_STORE_LN: return K_NT_STORE_LN; case EPI_RESID_LN: return K_NT_RESID_LN;
  }
  return K_MISC;
}

struct ProfRec { int tag; double flops, bytes; cudaEvent_t e0, e1; int where; };   // where = stage | part << 8 | backward << 16

struct ParamInfo {
  std::string name;
  int ndim;
  long shape[4];
  long numel;
};

struct Linear {            // one GEMM weight: bf16 copy W' [N,K] (rows optionally permuted) and its transpose
  int slot_w, slot_b;      // parameter slots (slot_b = -1: no bias)
  int N, K;
  int perm_R2, perm_Cc;    // row n' = ij*Cc + c holds source row c*R2 + ij (PixelShuffle-friendly order)
  long w_off, wt_off;      // element offsets into the bf16 weight arena
  long pbias_off;          // element offset into the fp32 aux arena of the permuted bias, -1 if not permuted
};

struct BlockDef {
  int n1w, n1b, table, n2w, n2b;
  int qkv, proj, fc1, fc2;     // indices into Net::linears
  int stage, shift, index;     // index = running block number (DropPath scale rows 2*index, 2*index+1)
};

struct BlockBuf { long xn1, st1, qkv, ao, xmid, xn2, st2, hpre, hact, xout; };

struct Plan {
  long pe_out;
  std::vector<BlockBuf> blocks;
  std::vector<long> xn_m, st_m, x_merged;          // per encoder stage (valid for s < L-1)
  long x_fpe;
  std::vector<long> x_skip, x_up;                  // per decoder stage
  long x_fpe_pre = -1, st_fpe = -1;                // PatchExpanding: rearranged rows before their LayerNorm, (mean, rstd) rows
  std::vector<long> x_up_pre, st_upx;
  long xn_up, st_up;
  long st_head = -1;                               // FinalPatchExpanding: (mean, rstd) of every output pixel
  long gA, gB, scr_gs, scr_gsm, scr_big, scr_dxn, scr_do, scr_dqkv;
  std::vector<long> g_save;
  long loss_acc;
  long gscr, gscr_bytes;                           // fp32 scratch for gradient copies (see GRAD_COPIES)
  long total;
};

// Small gradients that every CTA of a kernel accumulates (LayerNorm gamma/beta, bias tables, PatchEmbed, decoder_pred) go
// to GRAD_COPIES scratch copies (CTA b -> copy b % GRAD_COPIES) and are summed by one kernel at the end of the backward
// pass: same-address global atomics from a whole grid serialise in L2 (7-11 us per launch, measured).
constexpr int GRAD_COPIES = 32;

struct tulip_net {
  tulip_config cfg;
  int L;                                           // stages
  int H0, W0, r;                                   // token grid at stage 0, head upscale factor
  std::vector<ParamInfo> params;
  std::vector<Linear> linears;
  std::vector<BlockDef> blocks;                    // encoder blocks then decoder blocks, execution order
  std::vector<std::vector<int>> enc_blocks, dec_blocks;
  std::vector<int> merge_nw, merge_nb, merge_lin;  // per encoder stage
  std::vector<int> up_lin;                         // per decoder stage (-1: Identity)
  std::vector<int> up_nw, up_nb;                   // PatchExpanding only: LayerNorm(C/2) after the rearrange (-1: PatchUnmerging)
  int fpe_nw = -1, fpe_nb = -1;
  std::vector<int> skip_lin;
  int fpe_lin, head_lin;
  int slot_normup_w, slot_normup_b, slot_pe_w, slot_pe_b, slot_pe_nw, slot_pe_nb, slot_dec_w;
  int slot_fh_nw = -1, slot_fh_nb = -1;            // FinalPatchExpanding.norm (expanding_head only)
  // device-side persistent state
  bf16* warena = nullptr; long warena_elems = 0;
  float* faux = nullptr; long faux_elems = 0;
  PackItem* items_dev = nullptr; int n_items = 0, n_tiles = 0;
  std::vector<long> items_offsets_cache;           // parameter offsets the uploaded pack table was built for
  long kernel_launches = 0;
  // forward-only mode (no backward will follow): the fused half-block kernels run and nothing is saved for autograd
  bool inference = false;
  // parameter version announced by the caller (tulip_net_set_params_version): the bf16 weight arena is re-packed only when it
  // differs from the version (and buffer) the arena was packed from; -1 = unknown, always re-pack
  long long params_version = -1, packed_version = -1; const float* packed_from = nullptr;
  // per-launch profiler (off by default; adds two cudaEventRecord per launch when on)
  bool profiling = false;
  int cur_tag = K_MISC; double cur_flops = 0, cur_bytes = 0;
  int cur_stage = 0, cur_part = 0, cur_dir = 0;     // where the next launches belong: stage, 0 other / 1 attention half / 2 MLP half, 0 fwd / 1 bwd
  void at(int stage, int part) { cur_stage = stage; cur_part = part; }
  std::vector<ProfRec> recs;
  std::vector<cudaEvent_t> ev_pool; size_t ev_used = 0;
  // CUDA-graph replay of a whole direction (forward or backward): once the same call (same batch, same pointers) has been
  // seen twice in a row its ~150 launches are captured -- PDL edges, side-stream fork/joins and memsets included -- and later
  // calls are one cudaGraphLaunch.  TULIP_B200_GRAPHS=0 turns it off; the per-launch profiler bypasses it.
  struct GraphSlot {
    cudaGraphExec_t exec = nullptr;
    std::vector<uint64_t> key;          // key of `exec` (valid when exec != nullptr)
    std::vector<uint64_t> last;         // key of the previous call
    long launches = 0;                  // kernels per replay
  };
  GraphSlot graph_fwd, graph_bwd, graph_bwd_phase[3];
  // backward phases (gradient slices finish in this order, so their all-reduce can start under the rest of the pass):
  // 0 head + decoder + first_patch_expanding, 1 top encoder stage, 2 remaining encoder stages + PatchEmbed
  bool live = true;                                 // launches are issued only while the walk is inside the requested phases
  template <class F>
  int run_graphed(GraphSlot& slot, const std::vector<uint64_t>& key, cudaStream_t st, F&& body);
  // side stream for the weight-gradient GEMMs of the backward pass (they are off the dX critical path)
  cudaStream_t side = nullptr;
  std::vector<cudaEvent_t> sync_pool; size_t sync_used = 0;
  cudaEvent_t next_sync_event();
  void tag(int t, double flops, double bytes) { cur_tag = t; cur_flops = flops; cur_bytes = bytes; }
  void prof_begin(cudaStream_t st);
  void prof_end(cudaStream_t st);
  void prof_reset();

  int build();
  int ensure_device();
  Plan plan(int B) const;
  long grad_scratch_bytes() const;
  int upload_pack_table(const int64_t* offs, cudaStream_t st);
  int forward(int B, const float* params, const int64_t* offs, const float* x_lo, const float* target, const float* drop_scales,
              const int* win_mode, void* ws, float* pred, float* losses, cudaStream_t st);
  int backward(int B, const float* params, const int64_t* offs, float* grads, const float* x_lo, const float* target,
               const float* pred, const float* grad_loss, const float* drop_scales, const int* win_mode, void* ws,
               cudaStream_t st, int phase_lo = 0, int phase_hi = 2);
};
