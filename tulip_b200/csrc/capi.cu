// extern "C" surface of libtulip_b200.so (declared in include/tulip_b200.h).
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>

#include "net.h"
#include <vector>

namespace {
thread_local std::string g_error;
int g_num_sms = 0;
int g_sm_budget = 0;     // > 0: persistent kernels size their grids for this many SMs (tulip_set_sm_budget)
}  // namespace

void tulip_set_error(const char* msg) { g_error = msg ? msg : "unknown error"; }

int tulip_num_sms() {
  if (g_num_sms == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      g_num_sms = n;
    else
      g_num_sms = 148;
    const char* e = getenv("TULIP_B200_SM_BUDGET");      // measurement knob: the whole process under a budget
    if (e && atoi(e) > 0) g_sm_budget = atoi(e);
  }
  return (g_sm_budget > 0 && g_sm_budget < g_num_sms) ? g_sm_budget : g_num_sms;
}

int tulip_hints() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("TULIP_B200_HINTS");
    v = e ? atoi(e) : 127;
  }
  return v;
}

bool tulip_pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("TULIP_B200_NO_PDL");
    v = (e && e[0] == '1') ? 0 : 1;
  }
  return v == 1;
}

// One GEMM implementation (tcgen05, gemm_tc05.cu).  A shape it does not take is an error for the caller, never a detour
// through another kernel: every N on the TULIP path is a multiple of 96, every K a multiple of 8, operands 16-byte aligned.
int gemm_nt(const GemmArgs& g, int epi, cudaStream_t st) {
  const int rc = gemm_nt_tc05(g, epi, st);
  if (rc == TULIP_ERR_UNSUPPORTED)
    tulip_set_error("gemm_nt: shape / operand layout not handled by the tcgen05 GEMM (N % 96, K % 8, 16-byte alignment, epilogue limits)");
  return rc;
}

int gemm_tn(const GemmTNArgs& g, cudaStream_t st) {
  const int rc = gemm_tn_tc05(g, st);
  if (rc == TULIP_ERR_UNSUPPORTED)
    tulip_set_error("gemm_tn: shape / operand layout not handled by the tcgen05 GEMM (N % 8, K % 8, 16-byte alignment, gather geometry)");
  return rc;
}

static bool graphs_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("TULIP_B200_GRAPHS");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

static uint64_t hash_words(const void* p, size_t bytes) {          // FNV-1a over small host arrays (offsets, window modes)
  const unsigned char* b = static_cast<const unsigned char*>(p);
  uint64_t h = 1469598103934665603ull;
  for (size_t i = 0; i < bytes; ++i) { h ^= b[i]; h *= 1099511628211ull; }
  return h;
}

template <class F>
int tulip_net::run_graphed(GraphSlot& slot, const std::vector<uint64_t>& key, cudaStream_t st, F&& body) {
  if (!graphs_enabled() || profiling) return body();
  if (slot.exec && slot.key == key) {
    TULIP_CUDA(cudaGraphLaunch(slot.exec, st));
    kernel_launches += slot.launches;
    return TULIP_OK;
  }
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) return body();   // caller captures already
  if (slot.last != key) {                                  // first sighting of this call: run it eagerly (also warms every
    slot.last = key;                                       // one-time initialisation: attributes, occupancy queries, allocations)
    return body();
  }
  if (slot.exec) { cudaGraphExecDestroy(slot.exec); slot.exec = nullptr; }
  const long before = kernel_launches;
  if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); return body(); }
  const int rc = body();
  cudaGraph_t graph = nullptr;
  const cudaError_t ce = cudaStreamEndCapture(st, &graph);
  if (rc != TULIP_OK || ce != cudaSuccess || graph == nullptr) {
    if (graph) cudaGraphDestroy(graph);
    cudaGetLastError();
    kernel_launches = before;
    if (rc != TULIP_OK) return rc;
    slot.last.clear();
    return body();                                         // capture refused: fall back to plain launches
  }
  const cudaError_t ie = cudaGraphInstantiate(&slot.exec, graph, 0);
  cudaGraphDestroy(graph);
  if (ie != cudaSuccess) { slot.exec = nullptr; cudaGetLastError(); kernel_launches = before; slot.last.clear(); return body(); }
  slot.key = key;
  slot.launches = kernel_launches - before;
  kernel_launches = before;
  TULIP_CUDA(cudaGraphLaunch(slot.exec, st));
  kernel_launches += slot.launches;
  return TULIP_OK;
}


extern "C" {

const char* tulip_last_error(void) { return g_error.c_str(); }
int tulip_abi_version(void) { return 2; }   // 2: tulip_config.patch_expanding / expanding_head

int tulip_net_create(const tulip_config* cfg, tulip_net** out) {
  if (!cfg || !out) { tulip_set_error("tulip_net_create: null argument"); return TULIP_ERR_ARG; }
  tulip_net* n = new tulip_net();
  n->cfg = *cfg;
  const int rc = n->build();
  if (rc != TULIP_OK) { delete n; *out = nullptr; return rc; }
  *out = n;
  return TULIP_OK;
}

void tulip_net_destroy(tulip_net* n) {
  if (!n) return;
  if (n->warena) cudaFree(n->warena);
  if (n->faux) cudaFree(n->faux);
  if (n->items_dev) cudaFree(n->items_dev);
  for (cudaEvent_t e : n->sync_pool) cudaEventDestroy(e);
  for (cudaEvent_t e : n->ev_pool) cudaEventDestroy(e);
  if (n->side) cudaStreamDestroy(n->side);
  if (n->graph_fwd.exec) cudaGraphExecDestroy(n->graph_fwd.exec);
  if (n->graph_bwd.exec) cudaGraphExecDestroy(n->graph_bwd.exec);
  delete n;
}

int tulip_net_num_params(const tulip_net* n) { return (int)n->params.size(); }

int tulip_net_param_info(const tulip_net* n, int i, char* name, int name_cap, int64_t shape[4], int* ndim) {
  if (i < 0 || i >= (int)n->params.size()) { tulip_set_error("param index out of range"); return TULIP_ERR_ARG; }
  const ParamInfo& p = n->params[i];
  if (name && name_cap > 0) { strncpy(name, p.name.c_str(), name_cap - 1); name[name_cap - 1] = 0; }
  for (int k = 0; k < 4; ++k) shape[k] = p.shape[k];
  *ndim = p.ndim;
  return TULIP_OK;
}

int tulip_net_num_blocks(const tulip_net* n) { return (int)n->blocks.size(); }

int tulip_net_block_info(const tulip_net* n, int i, int* stage, int* shifted, int* H, int* W) {
  if (i < 0 || i >= (int)n->blocks.size()) { tulip_set_error("block index out of range"); return TULIP_ERR_ARG; }
  const BlockDef& b = n->blocks[i];
  *stage = b.stage; *shifted = b.shift; *H = n->H0 >> b.stage; *W = n->W0 >> b.stage;
  return TULIP_OK;
}

int64_t tulip_net_workspace_bytes(const tulip_net* n, int batch) { return batch > 0 ? n->plan(batch).total : 0; }
int64_t tulip_net_kernel_launches(const tulip_net* n) { return n->kernel_launches; }

int tulip_net_profile(tulip_net* n, int enable) {
  if (!n) { tulip_set_error("tulip_net_profile: null net"); return TULIP_ERR_ARG; }
  n->profiling = enable != 0;
  n->prof_reset();
  return TULIP_OK;
}

int tulip_net_profile_num_tags(void) { return K_COUNT; }

int tulip_net_profile_read(tulip_net* n, int tag_id, char* name, int name_cap, double* ms, double* flops, double* bytes,
                           int64_t* launches) {
  if (!n || tag_id < 0 || tag_id >= K_COUNT) { tulip_set_error("tulip_net_profile_read: bad argument"); return TULIP_ERR_ARG; }
  if (name && name_cap > 0) { strncpy(name, ktag_name(tag_id), name_cap - 1); name[name_cap - 1] = 0; }
  double t = 0, f = 0, b = 0;
  int64_t cnt = 0;
  for (const ProfRec& r : n->recs) {
    if (r.tag != tag_id) continue;
    if (cudaEventSynchronize(r.e1) != cudaSuccess) { tulip_set_error("profile: event sync failed"); return TULIP_ERR_CUDA; }
    float dt = 0.f;
    if (cudaEventElapsedTime(&dt, r.e0, r.e1) != cudaSuccess) { tulip_set_error("profile: elapsed failed"); return TULIP_ERR_CUDA; }
    t += dt; f += r.flops; b += r.bytes; ++cnt;
  }
  *ms = t; *flops = f; *bytes = b; *launches = cnt;
  return TULIP_OK;
}

int tulip_net_profile_record(tulip_net* n, int i, int* tag, double* ms, double* flops, double* bytes) {
  if (!n) { tulip_set_error("tulip_net_profile_record: null net"); return -1; }
  const int cnt = (int)n->recs.size();
  if (i < 0 || i >= cnt) return cnt;
  const ProfRec& r = n->recs[i];
  float dt = 0.f;
  if (cudaEventSynchronize(r.e1) != cudaSuccess || cudaEventElapsedTime(&dt, r.e0, r.e1) != cudaSuccess) {
    tulip_set_error("profile: event read failed");
    return -1;
  }
  if (tag) *tag = r.tag;
  if (ms) *ms = dt;
  if (flops) *flops = r.flops;
  if (bytes) *bytes = r.bytes;
  return cnt;
}

int tulip_net_profile_where(tulip_net* n, int i, int* stage, int* part, int* backward) {
  if (!n) { tulip_set_error("tulip_net_profile_where: null net"); return -1; }
  const int cnt = (int)n->recs.size();
  if (i < 0 || i >= cnt) return cnt;
  const int w = n->recs[i].where;
  if (stage) *stage = w & 0xff;
  if (part) *part = (w >> 8) & 0xff;
  if (backward) *backward = (w >> 16) & 1;
  return cnt;
}

int tulip_net_forward(tulip_net* n, int batch, const float* params, const int64_t* offs, const float* x_lo, const float* target,
                      const float* drop_scales, const int* win_mode, void* ws, float* pred, float* losses, void* stream) {
  if (!n || !params || !offs || !x_lo || !ws || !pred) { tulip_set_error("tulip_net_forward: null argument"); return TULIP_ERR_ARG; }
  if (target && !losses) { tulip_set_error("tulip_net_forward: losses is null"); return TULIP_ERR_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  // one-time device allocations and the (synchronising) pack-table upload must not happen inside a capture
  int rc = n->ensure_device();
  if (rc) return rc;
  rc = n->upload_pack_table(offs, st);
  if (rc) return rc;
  const std::vector<uint64_t> key = {
      (uint64_t)batch, (uint64_t)(uintptr_t)params, hash_words(offs, n->params.size() * sizeof(int64_t)), (uint64_t)(uintptr_t)x_lo,
      (uint64_t)(uintptr_t)target, (uint64_t)(uintptr_t)drop_scales, win_mode ? hash_words(win_mode, n->blocks.size() * sizeof(int)) : 0,
      (uint64_t)(uintptr_t)ws, (uint64_t)(uintptr_t)pred, (uint64_t)(uintptr_t)losses, (uint64_t)(uintptr_t)stream,
      (uint64_t)n->inference,
      (uint64_t)((n->params_version < 0 || n->params_version != n->packed_version || n->packed_from != params) ? 1 : 0)};
  return n->run_graphed(n->graph_fwd, key, st, [&]() {
    return n->forward(batch, params, offs, x_lo, target, drop_scales, win_mode, ws, pred, losses, st);
  });
}

int tulip_net_set_params_version(tulip_net* n, long long version) {
  if (!n) { tulip_set_error("tulip_net_set_params_version: null net"); return TULIP_ERR_ARG; }
  n->params_version = version;
  return TULIP_OK;
}

int tulip_net_set_inference(tulip_net* n, int forward_only) {
  if (!n) { tulip_set_error("tulip_net_set_inference: null net"); return TULIP_ERR_ARG; }
  n->inference = forward_only != 0;
  return TULIP_OK;
}

static int net_backward(tulip_net* n, int batch, const float* params, const int64_t* offs, float* grads, const float* x_lo,
                        const float* target, const float* pred, const float* grad_loss, const float* drop_scales,
                        const int* win_mode, void* ws, void* stream, int phase_lo, int phase_hi) {
  if (!n || !params || !offs || !grads || !x_lo || !ws) { tulip_set_error("tulip_net_backward: null argument"); return TULIP_ERR_ARG; }
  if (n->inference) { tulip_set_error("tulip_net_backward: the last forward ran in forward-only mode (tulip_net_set_inference)"); return TULIP_ERR_ARG; }
  if (phase_lo < 0 || phase_hi > 2 || phase_lo > phase_hi) { tulip_set_error("tulip_net_backward_phases: phases are 0..2"); return TULIP_ERR_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  const std::vector<uint64_t> key = {
      (uint64_t)batch, (uint64_t)(uintptr_t)params, hash_words(offs, n->params.size() * sizeof(int64_t)), (uint64_t)(uintptr_t)grads,
      (uint64_t)(uintptr_t)x_lo, (uint64_t)(uintptr_t)target, (uint64_t)(uintptr_t)pred, (uint64_t)(uintptr_t)grad_loss,
      (uint64_t)(uintptr_t)drop_scales, win_mode ? hash_words(win_mode, n->blocks.size() * sizeof(int)) : 0, (uint64_t)(uintptr_t)ws,
      (uint64_t)(uintptr_t)stream, (uint64_t)(phase_lo * 4 + phase_hi)};
  tulip_net::GraphSlot& slot = (phase_lo == 0 && phase_hi == 2) ? n->graph_bwd : n->graph_bwd_phase[phase_lo];
  return n->run_graphed(slot, key, st, [&]() {
    return n->backward(batch, params, offs, grads, x_lo, target, pred, grad_loss, drop_scales, win_mode, ws, st, phase_lo, phase_hi);
  });
}

int tulip_net_backward(tulip_net* n, int batch, const float* params, const int64_t* offs, float* grads, const float* x_lo,
                       const float* target, const float* pred, const float* grad_loss, const float* drop_scales,
                       const int* win_mode, void* ws, void* stream) {
  return net_backward(n, batch, params, offs, grads, x_lo, target, pred, grad_loss, drop_scales, win_mode, ws, stream, 0, 2);
}

int tulip_net_backward_phases(tulip_net* n, int batch, const float* params, const int64_t* offs, float* grads, const float* x_lo,
                              const float* target, const float* pred, const float* grad_loss, const float* drop_scales,
                              const int* win_mode, void* ws, void* stream, int phase_lo, int phase_hi) {
  return net_backward(n, batch, params, offs, grads, x_lo, target, pred, grad_loss, drop_scales, win_mode, ws, stream, phase_lo, phase_hi);
}

int tulip_gemm_nt(const void* A, const void* W, const float* bias, void* out, void* out2, const void* aux, const float* row_scale,
                  int rows_per_sample, int M, int N, int K, int epilogue, int impl, void* stream) {
  GemmArgs g;
  memset(&g, 0, sizeof g);
  g.A = (const bf16*)A; g.lda = K; g.K1 = K; g.B = (const bf16*)W; g.ldb = K; g.M = M; g.N = N; g.K = K; g.bias = bias;
  g.out = (bf16*)out; g.ldo = N; g.out2 = (bf16*)out2; g.ldo2 = N; g.aux = (const bf16*)aux; g.ldaux = N;
  g.row_scale = row_scale; g.rows_per_sample = rows_per_sample > 0 ? rows_per_sample : 1;
  if (epilogue != EPI_STORE && epilogue != EPI_GELU && epilogue != EPI_RESID && epilogue != EPI_DGELU) {
    tulip_set_error("tulip_gemm_nt: epilogue must be 0 (store), 1 (gelu), 2 (residual) or 5 (dgelu)");
    return TULIP_ERR_ARG;
  }
  if (impl != 0 && impl != 2) { tulip_set_error("tulip_gemm_nt: impl must be 0 (the warp-MMA backend, impl 1, was removed)"); return TULIP_ERR_ARG; }
  return gemm_nt(g, epilogue, (cudaStream_t)stream);
}

int tulip_gemm_nt_plan(int M, int N, int K, int epilogue, int save_pre, int* out10) {
  if (!out10) { tulip_set_error("tulip_gemm_nt_plan: null output"); return TULIP_ERR_ARG; }
  const int rc = gemm_nt_tc05_plan(M, N, K, epilogue, save_pre, out10);
  if (rc) tulip_set_error("tulip_gemm_nt_plan: shape not handled by the tcgen05 GEMM (N % 96, K % 8)");
  return rc;
}

int tulip_gemm_nt_pairs_mode(int mode) { return gemm_nt_pairs_mode(mode); }

int tulip_set_sm_budget(int n) {
  const int prev = g_sm_budget;
  g_sm_budget = n > 0 ? n : 0;
  return prev;
}

int tulip_gemm_tn(const void* dY, const void* X, float* dW, float* db, int M, int N, int K, int impl, void* stream) {
  GemmTNArgs g;
  memset(&g, 0, sizeof g);
  g.dY = (const bf16*)dY; g.ldy = N; g.X = (const bf16*)X; g.ldx = K; g.K1 = K; g.M = M; g.N = N; g.K = K;
  g.dW = dW; g.lddw = K; g.db = db; g.perm_R2 = 1; g.perm_Cc = 1;
  const int tiles = (N / 96) * (K / 96);
  int splits = tiles > 0 ? (2 * tulip_num_sms() + tiles - 1) / tiles : 1;
  const int max_splits = (M + 255) / 256;
  g.splits = splits > max_splits ? max_splits : (splits < 1 ? 1 : splits);
  if (impl != 0 && impl != 2) { tulip_set_error("tulip_gemm_tn: impl must be 0 (the warp-MMA backend, impl 1, was removed)"); return TULIP_ERR_ARG; }
  return gemm_tn(g, (cudaStream_t)stream);
}

int tulip_gemm_nt_ex(const tulip_gemm_desc* d, int epilogue, void* stream) {
  if (!d) { tulip_set_error("tulip_gemm_nt_ex: null descriptor"); return TULIP_ERR_ARG; }
  if (epilogue < EPI_STORE || epilogue > EPI_RESID_LN) { tulip_set_error("tulip_gemm_nt_ex: unknown epilogue"); return TULIP_ERR_ARG; }
  GemmArgs g;
  memset(&g, 0, sizeof g);
  g.A = (const bf16*)d->A; g.lda = d->lda; g.A2 = (const bf16*)d->A2; g.lda2 = d->lda2; g.K1 = d->K1 > 0 ? d->K1 : d->K;
  g.B = (const bf16*)d->B; g.ldb = d->ldb; g.B2 = (const bf16*)d->B2; g.ldb2 = d->ldb2; g.K2 = d->K2;
  g.M = d->M; g.N = d->N; g.K = d->K;
  g.a_mode = d->a_mode; g.g_H = d->g_H; g.g_W = d->g_W; g.g_Cc = d->g_Cc;
  g.bias = d->bias;
  g.out = (bf16*)d->out; g.ldo = d->ldo; g.out2 = (bf16*)d->out2; g.ldo2 = d->ldo2; g.aux = (const bf16*)d->aux; g.ldaux = d->ldaux;
  g.row_scale = d->row_scale; g.rows_per_sample = d->rows_per_sample > 0 ? d->rows_per_sample : 1;
  g.split_col = d->split_col;
  g.wd = d->wd; g.target = d->target; g.pred = d->pred; g.gscale = d->gscale; g.dwd = d->dwd; g.dwd_copies = 1;
  g.hd_H = d->hd_H; g.hd_W = d->hd_W; g.hd_r = d->hd_r; g.hd_E = d->hd_E;
  if (d->hd_r > 0) g.hd_inv_npix = 1.0f / ((float)d->M * d->hd_r * d->hd_r);
  g.aux2 = (const bf16*)d->aux2; g.ldaux2 = d->ldaux2;
  g.ln_w = d->ln_w; g.ln_stats = d->ln_stats; g.ln_dw = d->ln_dw; g.ln_db = d->ln_db; g.ln_copies = 1;
  g.ln_b = d->ln_b; g.ln_y = (bf16*)d->ln_y; g.ln_ystats = d->ln_ystats; g.ln_eps = d->ln_eps;
  g.hd_ln = d->hd_ln;
  return gemm_nt(g, epilogue, (cudaStream_t)stream);
}

int tulip_gemm_tn_ex(const tulip_gemm_tn_desc* d, void* stream) {
  if (!d) { tulip_set_error("tulip_gemm_tn_ex: null descriptor"); return TULIP_ERR_ARG; }
  GemmTNArgs g;
  memset(&g, 0, sizeof g);
  g.dY = (const bf16*)d->dY; g.ldy = d->ldy; g.X = (const bf16*)d->X; g.ldx = d->ldx; g.X2 = (const bf16*)d->X2; g.ldx2 = d->ldx2;
  g.K1 = d->K1 > 0 ? d->K1 : d->K;
  g.M = d->M; g.N = d->N; g.K = d->K;
  g.y_mode = d->y_mode; g.g_H = d->g_H; g.g_W = d->g_W; g.g_Cc = d->g_Cc;
  g.dW = d->dW; g.lddw = d->lddw; g.db = d->db;
  g.perm_R2 = d->perm_R2 > 1 ? d->perm_R2 : 1; g.perm_Cc = d->perm_R2 > 1 ? d->perm_Cc : 1;
  const int tiles = (g.N / 96) * (g.K / 96);
  int splits = tiles > 0 ? (2 * tulip_num_sms() + tiles - 1) / tiles : 1;
  const int max_splits = (g.M + 255) / 256;
  g.splits = splits > max_splits ? max_splits : (splits < 1 ? 1 : splits);
  return gemm_tn(g, (cudaStream_t)stream);
}

int tulip_gemm_tn_group(const tulip_gemm_tn_desc* d, int n, void* stream) {
  if (!d || n < 1 || n > TN_GROUP_MAX) { tulip_set_error("tulip_gemm_tn_group: 1..4 descriptors"); return TULIP_ERR_ARG; }
  GemmTNArgs gs[TN_GROUP_MAX];
  for (int i = 0; i < n; ++i) {
    GemmTNArgs& g = gs[i];
    memset(&g, 0, sizeof g);
    g.dY = (const bf16*)d[i].dY; g.ldy = d[i].ldy; g.X = (const bf16*)d[i].X; g.ldx = d[i].ldx;
    g.K1 = d[i].K; g.M = d[i].M; g.N = d[i].N; g.K = d[i].K;
    g.y_mode = d[i].y_mode;
    g.dW = d[i].dW; g.lddw = d[i].lddw; g.db = d[i].db;
    g.perm_R2 = d[i].perm_R2 > 1 ? d[i].perm_R2 : 1; g.perm_Cc = d[i].perm_R2 > 1 ? d[i].perm_Cc : 1;
    g.splits = 1;
    if (d[i].X2 || (d[i].K1 > 0 && d[i].K1 < d[i].K) || !gemm_tn_groupable(g)) {
      tulip_set_error("tulip_gemm_tn_group: plain operands only (no X2, y_mode 0, 16-byte aligned, N % 8 == K % 8 == 0)");
      return TULIP_ERR_UNSUPPORTED;
    }
  }
  return gemm_tn_group(gs, n, (cudaStream_t)stream);
}

int tulip_gemm_tn_group_plan(const int* M, const int* N, const int* K, int n, int sms, int* per4, int* items) {
  if (!M || !N || !K || !per4 || sms < 1) { tulip_set_error("tulip_gemm_tn_group_plan: null argument"); return TULIP_ERR_ARG; }
  const int rc = gemm_tn_group_plan(M, N, K, n, sms, per4, items);
  if (rc) tulip_set_error("tulip_gemm_tn_group_plan: 1..4 problems with positive sizes");
  return rc;
}

int tulip_head_bwd_fused_supported(int E, int r) { return head_bwd_fused_supported(E, r) ? 1 : 0; }

int tulip_head_bwd_fused(const void* xn, void* dxn, const void* we, const void* wet, const float* bias, const float* wd,
                         const float* pred, const float* target, const float* gscale, void* dh, float* dwd, int T, int E, int H,
                         int W, int r, void* stream) {
  HeadBwdArgs hb;
  memset(&hb, 0, sizeof hb);
  hb.xn = (const bf16*)xn; hb.dxn = (bf16*)dxn; hb.we = (const bf16*)we; hb.wet = (const bf16*)wet; hb.bias = bias; hb.wd = wd;
  hb.pred = pred; hb.target = target; hb.gscale = gscale; hb.dh = (bf16*)dh; hb.dwd = dwd; hb.dwd_copies = 1;
  hb.T = T; hb.E = E; hb.H = H; hb.W = W; hb.r = r;
  return head_bwd_fused(hb, (cudaStream_t)stream);
}

int tulip_wmsa_block_supported(int B, int H, int W, int C, int heads, int Mh, int Mw) {
  return wmsa_block_supported(B, H, W, C, heads, Mh, Mw) ? 1 : 0;
}

int tulip_wmsa_block_fwd(const void* x, void* y, const float* ln_w, const float* ln_b, const void* wqkv, const float* bqkv,
                         const void* wproj, const float* bproj, const float* bias_table, const float* row_scale, int B, int H, int W,
                         int C, int heads, int Mh, int Mw, int sh, int sw, int masked, int bias_Mh, int bias_Mw, float eps,
                         void* stream) {
  WmsaBlockArgs a;
  memset(&a, 0, sizeof a);
  a.x = (const bf16*)x; a.y = (bf16*)y; a.ln_w = ln_w; a.ln_b = ln_b; a.wqkv = (const bf16*)wqkv; a.bqkv = bqkv;
  a.wproj = (const bf16*)wproj; a.bproj = bproj; a.bias_table = bias_table; a.row_scale = row_scale;
  a.B = B; a.H = H; a.W = W; a.C = C; a.heads = heads; a.Mh = Mh; a.Mw = Mw; a.sh = sh; a.sw = sw; a.masked = masked;
  a.bMh = bias_Mh; a.bMw = bias_Mw; a.eps = eps;
  return wmsa_block_fwd(a, (cudaStream_t)stream);
}

int tulip_mlp_block_supported(int T, int C) { return mlp_block_supported(T, C) ? 1 : 0; }

int tulip_mlp_block_fwd(const void* x, void* y, const float* ln_w, const float* ln_b, const void* w1, const float* b1, const void* w2,
                        const float* b2, const float* row_scale, int rows_per_sample, void* xn, float* stats, void* hact, int T, int C,
                        float eps, void* stream) {
  MlpBlockArgs a;
  memset(&a, 0, sizeof a);
  a.x = (const bf16*)x; a.y = (bf16*)y; a.ln_w = ln_w; a.ln_b = ln_b; a.w1 = (const bf16*)w1; a.b1 = b1; a.w2 = (const bf16*)w2; a.b2 = b2;
  a.row_scale = row_scale; a.rows_per_sample = rows_per_sample; a.xn = (bf16*)xn; a.stats = stats; a.hact = (bf16*)hact;
  a.T = T; a.C = C; a.eps = eps;
  return mlp_block_fwd(a, (cudaStream_t)stream);
}

static AttnArgs make_attn(const void* qkv, const float* table, int B, int H, int W, int C, int heads, int Mh, int Mw, int sh, int sw,
                          int masked, int bMh, int bMw) {
  AttnArgs a;
  memset(&a, 0, sizeof a);
  a.qkv = (const bf16*)qkv; a.bias_table = table; a.B = B; a.H = H; a.W = W; a.C = C; a.heads = heads;
  a.Mh = Mh; a.Mw = Mw; a.sh = sh; a.sw = sw; a.masked = masked; a.bMh = bMh; a.bMw = bMw;
  a.nbias = (2 * bMh - 1) * (2 * bMw - 1);
  a.scale = 1.0f / sqrtf((float)(heads > 0 ? C / heads : 1));
  return a;
}

int tulip_window_attention_fwd(const void* qkv, const float* bias_table, void* out, int B, int H, int W, int C, int heads, int Mh,
                               int Mw, int sh, int sw, int masked, int bias_Mh, int bias_Mw, void* stream) {
  AttnArgs a = make_attn(qkv, bias_table, B, H, W, C, heads, Mh, Mw, sh, sw, masked, bias_Mh, bias_Mw);
  a.out = (bf16*)out;
  return win_attn_fwd(a, (cudaStream_t)stream);
}

int tulip_window_attention_bwd(const void* qkv, const float* bias_table, const void* dout, void* dqkv, float* dbias_table, int B,
                               int H, int W, int C, int heads, int Mh, int Mw, int sh, int sw, int masked, int bias_Mh, int bias_Mw,
                               void* stream) {
  AttnArgs a = make_attn(qkv, bias_table, B, H, W, C, heads, Mh, Mw, sh, sw, masked, bias_Mh, bias_Mw);
  a.dout = (const bf16*)dout; a.dqkv = (bf16*)dqkv; a.dbias_table = dbias_table;
  return win_attn_bwd(a, (cudaStream_t)stream);
}

int tulip_layernorm_fwd(const void* x, const float* w, const float* b, void* y, float* stats, int rows, int C, float eps,
                        int merge_gather_, int H2, int W2, void* stream) {
  LnArgs a;
  memset(&a, 0, sizeof a);
  a.x = (const bf16*)x; a.w = w; a.b = b; a.y = (bf16*)y; a.stats = stats; a.rows = rows; a.C = C; a.eps = eps;
  a.gather = merge_gather_; a.H2 = H2; a.W2 = W2;
  return layernorm_fwd(a, (cudaStream_t)stream);
}

int tulip_layernorm_bwd(const void* x, const float* w, const float* stats, const void* dy, const void* dres, void* dx, float* dw,
                        float* db, int rows, int C, int merge_gather_, int H2, int W2, void* stream) {
  LnArgs a;
  memset(&a, 0, sizeof a);
  a.x = (const bf16*)x; a.w = w; a.stats = const_cast<float*>(stats); a.dy = (const bf16*)dy; a.dres = (const bf16*)dres;
  a.dx = (bf16*)dx; a.dw = dw; a.db = db; a.rows = rows; a.C = C; a.gather = merge_gather_; a.H2 = H2; a.W2 = W2;
  return layernorm_bwd(a, (cudaStream_t)stream);
}

int tulip_patch_embed_fwd(const float* x, const float* w, const float* b, const float* ln_w, const float* ln_b, void* y, int B,
                          int Himg, int Wimg, int ph, int E, float eps, void* stream) {
  EmbedArgs e;
  memset(&e, 0, sizeof e);
  e.x = x; e.w = w; e.b = b; e.ln_w = ln_w; e.ln_b = ln_b; e.y = (bf16*)y; e.B = B; e.Himg = Himg; e.Wimg = Wimg; e.ph = ph; e.E = E;
  e.eps = eps;
  return patch_embed_fwd(e, (cudaStream_t)stream);
}

int tulip_patch_embed_bwd(const float* x, const float* w, const float* b, const float* ln_w, const void* dy, float* dw, float* db,
                          float* dln_w, float* dln_b, int B, int Himg, int Wimg, int ph, int E, float eps, void* stream) {
  EmbedArgs e;
  memset(&e, 0, sizeof e);
  e.x = x; e.w = w; e.b = b; e.ln_w = ln_w; e.B = B; e.Himg = Himg; e.Wimg = Wimg; e.ph = ph; e.E = E; e.eps = eps;
  e.dy = (const bf16*)dy; e.dw = dw; e.db = db; e.dln_w = dln_w; e.dln_b = dln_b;
  return patch_embed_bwd(e, (cudaStream_t)stream);
}

int tulip_eval_postprocess(const float* pred, const float* x_lo, const float* target, float* out, float* losses, float* scratch,
                            int B, int H, int W, int h_lo, int log_transform, float clip_lo, int keep_low_res, void* stream) {
  if (!pred || !x_lo || !target || !out || !losses || !scratch) { tulip_set_error("tulip_eval_postprocess: null argument"); return TULIP_ERR_ARG; }
  return eval_postprocess(pred, x_lo, target, out, losses, scratch, B, H, W, h_lo, log_transform, clip_lo, keep_low_res, (cudaStream_t)stream);
}

int tulip_mc_dropout_aggregate(const float* preds, float* out, float* std_out, int n_passes, int64_t npix, float threshold, void* stream) {
  if (!preds || !out) { tulip_set_error("tulip_mc_dropout_aggregate: null argument"); return TULIP_ERR_ARG; }
  return mc_aggregate(preds, out, std_out, n_passes, (long)npix, threshold, (cudaStream_t)stream);
}

int tulip_range_to_points(const float* img, const float* sin_h, const float* cos_h, const float* sin_v, const float* cos_v,
                          float max_range, float* points, int B, int H, int W, void* stream) {
  if (!img || !sin_h || !cos_h || !sin_v || !cos_v || !points) { tulip_set_error("tulip_range_to_points: null argument"); return TULIP_ERR_ARG; }
  return range_to_points(img, sin_h, cos_h, sin_v, cos_v, max_range, points, B, H, W, (cudaStream_t)stream);
}

int64_t tulip_voxel_metrics_workspace_bytes(int n_points) { return n_points > 0 ? (int64_t)voxel_metrics_workspace_bytes(n_points) : 0; }

int tulip_voxel_metrics(const float* pts_pred, const float* pts_gt, int n_points, float grid_size, void* workspace, double* out4,
                        void* stream) {
  if (!pts_pred || !pts_gt || !workspace || !out4) { tulip_set_error("tulip_voxel_metrics: null argument"); return TULIP_ERR_ARG; }
  return voxel_metrics(pts_pred, pts_gt, n_points, grid_size, workspace, out4, (cudaStream_t)stream);
}

int tulip_range_to_points_durlar(const float* img, const double* cos_ea, const double* sin_ea, const double* cos_e, const double* sin_e,
                                 const double* cos_el, const double* sin_el, const int* offset_lut, float max_range, double origin_offset,
                                 double z_offset, double* points, int B, int H, int W, void* stream) {
  if (!img || !cos_ea || !sin_ea || !cos_e || !sin_e || !cos_el || !sin_el || !offset_lut || !points) {
    tulip_set_error("tulip_range_to_points_durlar: null argument");
    return TULIP_ERR_ARG;
  }
  return range_to_points_durlar(img, cos_ea, sin_ea, cos_e, sin_e, cos_el, sin_el, offset_lut, max_range, origin_offset, z_offset, points, B,
                                H, W, (cudaStream_t)stream);
}

int tulip_voxel_metrics_f64(const double* pts_pred, const double* pts_gt, int n_points, double grid_size, void* workspace, double* out4,
                            void* stream) {
  if (!pts_pred || !pts_gt || !workspace || !out4) { tulip_set_error("tulip_voxel_metrics_f64: null argument"); return TULIP_ERR_ARG; }
  return voxel_metrics_f64(pts_pred, pts_gt, n_points, grid_size, workspace, out4, (cudaStream_t)stream);
}

int tulip_chamfer_distance(const float* a, const float* b, int na, int nb, float* dist_a, float* dist_b, float* out3, void* stream) {
  if (!a || !b || !dist_a || !dist_b || !out3) { tulip_set_error("tulip_chamfer_distance: null argument"); return TULIP_ERR_ARG; }
  return chamfer_distance(a, b, na, nb, dist_a, dist_b, out3, (cudaStream_t)stream);
}

int tulip_preprocess_range(const float* raw, int channels, float scale, int filter, float min_range, float max_range, int row_factor,
                           int col_factor, int log_transform, float* hi, float* lo, int B, int H, int W, void* stream) {
  if (!raw || !hi || !lo) { tulip_set_error("tulip_preprocess_range: null argument"); return TULIP_ERR_ARG; }
  return preprocess_range(raw, channels, scale, filter, min_range, max_range, row_factor, col_factor, log_transform, hi, lo, B, H, W,
                          (cudaStream_t)stream);
}

int tulip_rimg_decode(const void* rows_f16, float* frames, int B, int size0, int size1, void* stream) {
  if (!rows_f16 || !frames) { tulip_set_error("tulip_rimg_decode: null argument"); return TULIP_ERR_ARG; }
  return rimg_decode(rows_f16, frames, B, size0, size1, (cudaStream_t)stream);
}

int tulip_adamw_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, const tulip_adamw_segment* segments_dev,
                     int n_segments, int64_t span, const tulip_adamw_hyper* hyper_host, void* stream) {
  if (!params || !grads || !exp_avg || !exp_avg_sq || !segments_dev || !hyper_host) { tulip_set_error("tulip_adamw_step: null argument"); return TULIP_ERR_ARG; }
  static_assert(sizeof(tulip_adamw_segment) == sizeof(AdamwSegment) && sizeof(tulip_adamw_hyper) == sizeof(AdamwHyper), "ABI structs");
  AdamwHyper hp;
  memcpy(&hp, hyper_host, sizeof hp);
  return adamw_step(params, grads, exp_avg, exp_avg_sq, reinterpret_cast<const AdamwSegment*>(segments_dev), n_segments, (long)span, hp,
                    (cudaStream_t)stream);
}

int tulip_grad_norm(const float* grads, const tulip_adamw_segment* segments_dev, int n_segments, int64_t span, double* scratch, float* out,
                    void* stream) {
  if (!grads || !segments_dev || !scratch || !out) { tulip_set_error("tulip_grad_norm: null argument"); return TULIP_ERR_ARG; }
  return grad_norm(grads, reinterpret_cast<const AdamwSegment*>(segments_dev), n_segments, (long)span, scratch, out, (cudaStream_t)stream);
}

int tulip_stage_inputs(const float* lo, float* lo_dst, int64_t n_lo, const float* hi, float* hi_dst, int64_t n_hi, const float* drop,
                       float* drop_dst, int64_t n_drop, void* stream) {
  const float* src[3] = {lo, hi, drop};
  float* dst[3] = {lo_dst, hi_dst, drop_dst};
  const long n[3] = {(long)n_lo, (long)n_hi, (long)n_drop};
  return stage_inputs(src, dst, n, (cudaStream_t)stream);
}

int tulip_l1_loss(const float* pred, const float* target, int64_t n, int log_transform, float* scratch2, float* out2, void* stream) {
  return l1_loss(pred, target, (long)n, log_transform, scratch2, out2, (cudaStream_t)stream);
}

int tulip_window_partition(const void* x, void* out, int B, int H, int W, int C, int Mh, int Mw, int sh, int sw, void* stream) {
  return window_gather((const bf16*)x, (bf16*)out, B, H, W, C, Mh, Mw, sh, sw, (cudaStream_t)stream);
}
int tulip_window_reverse(const void* xw, void* out, int B, int H, int W, int C, int Mh, int Mw, int sh, int sw, void* stream) {
  return window_scatter((const bf16*)xw, (bf16*)out, B, H, W, C, Mh, Mw, sh, sw, (cudaStream_t)stream);
}
int tulip_shift_mask(float* out, int H, int W, int Mh, int Mw, int sh, int sw, void* stream) {
  return shift_mask(out, H, W, Mh, Mw, sh, sw, (cudaStream_t)stream);
}
int tulip_rel_bias_gather(const float* table, float* out, int heads, int Mh, int Mw, void* stream) {
  return rel_bias_gather(table, out, heads, Mh, Mw, (cudaStream_t)stream);
}
int tulip_merge_gather(const void* x, void* out, int B, int H, int W, int C, void* stream) {
  return merge_gather((const bf16*)x, (bf16*)out, B, H, W, C, (cudaStream_t)stream);
}
int tulip_pixel_shuffle(const void* x, void* out, int B, int H, int W, int Cout, int r, void* stream) {
  return pixel_shuffle_nhwc((const bf16*)x, (bf16*)out, B, H, W, Cout, r, (cudaStream_t)stream);
}

}  // extern "C"
