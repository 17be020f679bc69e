// Whole-network executor (see net.h).  Reference orchestration: tulip/model/tulip.py:702-737
// (TULIP.forward), :431-436 (BasicBlock), :477-481 (BasicBlockUp), :338-352 (SwinTransformerBlock).
#include "net.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <utility>

namespace {

long align_up(long v, long a) { return (v + a - 1) / a * a; }

struct Bump {
  long off = 0;
  long take(long bytes) {
    const long o = off;
    off = align_up(off + bytes, 256);
    return o;
  }
};

}  // namespace

const char* ktag_name(int t) {
  static const char* names[K_COUNT] = {
      "misc", "gemm_nt<store>", "gemm_nt<gelu>", "gemm_nt<resid>", "gemm_nt<pixshuf>", "gemm_nt<split2>", "gemm_nt<dgelu>",
      "gemm_nt<head>", "gemm_nt<head_bwd>", "gemm_nt<rowscale>", "gemm_nt<unshuffle>", "gemm_tn", "gemm_tn<unshuffle>",
      "win_attn_fwd", "win_attn_bwd", "layernorm_fwd", "layernorm_bwd", "patch_embed_fwd", "patch_embed_bwd", "pack_weights",
      "elementwise", "l1_loss", "wmsa_block_fwd", "mlp_block_fwd", "gemm_nt<ln_bwd>", "gemm_nt<store+ln>", "gemm_nt<resid+ln>"};
  return (t >= 0 && t < K_COUNT) ? names[t] : "?";
}

void tulip_net::prof_begin(cudaStream_t st) {
  if (!profiling) return;
  while (ev_pool.size() < ev_used + 2) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    ev_pool.push_back(e);
  }
  cudaEventRecord(ev_pool[ev_used], st);
}

void tulip_net::prof_end(cudaStream_t st) {
  if (profiling) {
    cudaEventRecord(ev_pool[ev_used + 1], st);
    recs.push_back(ProfRec{cur_tag, cur_flops, cur_bytes, ev_pool[ev_used], ev_pool[ev_used + 1], cur_stage | (cur_part << 8) | (cur_dir << 16)});
    ev_used += 2;
  }
  cur_tag = K_MISC; cur_flops = 0; cur_bytes = 0;
}

void tulip_net::prof_reset() {
  recs.clear();
  ev_used = 0;
}

cudaEvent_t tulip_net::next_sync_event() {
  if (sync_used == sync_pool.size()) {
    cudaEvent_t e;
    cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    sync_pool.push_back(e);
  }
  return sync_pool[sync_used++];
}

// problems per grouped weight-gradient launch (TULIP_B200_TN_GROUP_MAX = 1: every problem goes out where it is registered)
static int tn_group_limit() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("TULIP_B200_TN_GROUP_MAX");
    v = (e && atoi(e) >= 1 && atoi(e) <= TN_GROUP_MAX) ? atoi(e) : TN_GROUP_MAX;
  }
  return v;
}

// MEASUREMENT ONLY (TULIP_B200_DEBUG_SKIP_TN=1): the weight-gradient GEMMs are not launched, so the step that remains is the
// dX chain alone -- the difference to the real step is what the weight gradients cost after overlap.  Gradients are wrong.
static bool debug_skip_tn() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("TULIP_B200_DEBUG_SKIP_TN");
    v = (e && e[0] == '1') ? 1 : 0;
    if (v) fprintf(stderr, "tulip_b200: TULIP_B200_DEBUG_SKIP_TN=1 -- weight-gradient GEMMs are NOT launched, gradients are WRONG (timing only)\n");
  }
  return v == 1;
}

static bool side_stream_disabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("TULIP_B200_NO_SIDE");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

int tulip_net::build() {
  const tulip_config& c = cfg;
  TULIP_REQUIRE(c.num_layers >= 2 && c.num_layers <= TULIP_MAX_STAGES, "tulip: num_layers must be in [2, 8]");
  TULIP_REQUIRE(c.in_chans == 1, "tulip_b200: in_chans must be 1 (all shipped configurations)");
  TULIP_REQUIRE(c.patch_w == 4, "tulip_b200: patch width must be 4 (circular-padding conv k=(ph,8), s=(ph,4))");
  TULIP_REQUIRE(c.patch_h == 1, "tulip_b200: patch height must be 1");
  TULIP_REQUIRE(c.win_h * c.win_w == 16, "tulip_b200: windows must hold 16 tokens");
  TULIP_REQUIRE(c.embed_dim % 96 == 0 && c.embed_dim <= 192, "tulip_b200: embed_dim must be 96 or 192");
  TULIP_REQUIRE(c.mlp_ratio == 4, "tulip_b200: mlp_ratio must be 4");
  TULIP_REQUIRE(c.img_h % c.patch_h == 0 && c.img_w % c.patch_w == 0, "tulip: image not divisible by the patch size");
  TULIP_REQUIRE(c.expanding_head == 0 || c.embed_dim == 96,
                "tulip_b200: the FinalPatchExpanding head (pixel_shuffle=False, tulip.py:144-159) is built for embed_dim 96 only");
  L = c.num_layers;
  H0 = c.img_h / c.patch_h;
  W0 = c.img_w / c.patch_w;
  // upscale factor exactly as tulip.py:577
  {
    const double ratio = ((double)c.tgt_h * c.tgt_w) / ((double)c.img_h * c.img_w);
    const int a = (int)std::sqrt(ratio);
    const int b = (int)std::sqrt((double)((c.patch_h * c.patch_w) / 4));
    r = a * 2 * b;
  }
  TULIP_REQUIRE(r >= 1 && H0 * r == c.tgt_h && W0 * r == c.tgt_w,
                "tulip: upscale_factor (tulip.py:577) does not map the token grid onto target_img_size");
  for (int s = 0; s < L; ++s) {
    const int C = c.embed_dim << s;
    TULIP_REQUIRE(c.num_heads[s] * 32 == C, "tulip_b200: head_dim must be 32 at every stage");
    TULIP_REQUIRE(c.depths[s] >= 1, "tulip: depths must be positive");
    if (s < L - 1) TULIP_REQUIRE(((H0 >> s) % 2 == 0) && ((W0 >> s) % 2 == 0), "tulip_b200: odd grid before PatchMerging (zero-pad path not built)");
  }

  auto add_param = [&](const std::string& name, std::initializer_list<long> shape) {
    ParamInfo p;
    p.name = name;
    p.ndim = (int)shape.size();
    p.numel = 1;
    int i = 0;
    for (long d : shape) { p.shape[i++] = d; p.numel *= d; }
    for (; i < 4; ++i) p.shape[i] = 1;
    params.push_back(p);
    return (int)params.size() - 1;
  };
  auto add_linear = [&](int slot_w, int slot_b, int N, int K, int R2, int Cc) {
    Linear l{slot_w, slot_b, N, K, R2, Cc, 0, 0, -1};
    linears.push_back(l);
    return (int)linears.size() - 1;
  };
  const int nbias = (2 * c.win_h - 1) * (2 * c.win_w - 1);
  int running = 0;
  auto add_block = [&](const std::string& pre, int stage, int bidx) {
    const long C = c.embed_dim << stage;
    const long heads = c.num_heads[stage];
    BlockDef b;
    b.stage = stage;
    b.shift = bidx % 2;                                   // tulip.py:417,460
    b.index = running++;
    b.n1w = add_param(pre + ".norm1.weight", {C});
    b.n1b = add_param(pre + ".norm1.bias", {C});
    b.table = add_param(pre + ".attn.relative_position_bias_table", {nbias, heads});
    const int qw = add_param(pre + ".attn.qkv.weight", {3 * C, C});
    const int qb = add_param(pre + ".attn.qkv.bias", {3 * C});
    const int pw = add_param(pre + ".attn.proj.weight", {C, C});
    const int pb = add_param(pre + ".attn.proj.bias", {C});
    b.n2w = add_param(pre + ".norm2.weight", {C});
    b.n2b = add_param(pre + ".norm2.bias", {C});
    const int f1w = add_param(pre + ".mlp.fc1.weight", {4 * C, C});
    const int f1b = add_param(pre + ".mlp.fc1.bias", {4 * C});
    const int f2w = add_param(pre + ".mlp.fc2.weight", {C, 4 * C});
    const int f2b = add_param(pre + ".mlp.fc2.bias", {C});
    b.qkv = add_linear(qw, qb, 3 * C, C, 1, 0);
    b.proj = add_linear(pw, pb, C, C, 1, 0);
    b.fc1 = add_linear(f1w, f1b, 4 * C, C, 1, 0);
    b.fc2 = add_linear(f2w, f2b, C, 4 * C, 1, 0);
    blocks.push_back(b);
    return (int)blocks.size() - 1;
  };

  enc_blocks.assign(L, {});
  dec_blocks.assign(L - 1, {});
  merge_nw.assign(L, -1); merge_nb.assign(L, -1); merge_lin.assign(L, -1);
  up_lin.assign(L - 1, -1);
  up_nw.assign(L - 1, -1); up_nb.assign(L - 1, -1);
  skip_lin.assign(L - 1, -1);
  char buf[128];
  for (int s = 0; s < L; ++s) {                           // tulip.py:643-660
    const long C = c.embed_dim << s;
    for (int b = 0; b < c.depths[s]; ++b) {
      snprintf(buf, sizeof buf, "layers.%d.blocks.%d", s, b);
      enc_blocks[s].push_back(add_block(buf, s, b));
    }
    if (s < L - 1) {
      snprintf(buf, sizeof buf, "layers.%d.downsample", s);
      merge_nw[s] = add_param(std::string(buf) + ".norm.weight", {4 * C});
      merge_nb[s] = add_param(std::string(buf) + ".norm.bias", {4 * C});
      const int w = add_param(std::string(buf) + ".reduction.weight", {2 * C, 4 * C});
      merge_lin[s] = add_linear(w, -1, 2 * C, 4 * C, 1, 0);
    }
  }
  for (int u = 0; u < L - 1; ++u) {                       // tulip.py:662-680; stage index per :447
    const int s = L - u - 2;
    const long C = c.embed_dim << s;
    for (int b = 0; b < c.depths[s]; ++b) {
      snprintf(buf, sizeof buf, "layers_up.%d.blocks.%d", u, b);
      dec_blocks[u].push_back(add_block(buf, s, b));
    }
    if (u < L - 2 && !c.patch_expanding) {                 // PatchUnmerging, tulip.py:109-115
      snprintf(buf, sizeof buf, "layers_up.%d.upsample.expand", u);
      const int w = add_param(std::string(buf) + ".weight", {2 * C, C, 1, 1});
      const int bb = add_param(std::string(buf) + ".bias", {2 * C});
      up_lin[u] = add_linear(w, bb, 2 * C, C, 4, C / 2);
    } else if (u < L - 2) {                                // PatchExpanding, tulip.py:126-132: the Linear's columns are already
      snprintf(buf, sizeof buf, "layers_up.%d.upsample", u); // in shuffle-slot order (P1 P2 C): no row permutation, no bias
      const int w = add_param(std::string(buf) + ".expand.weight", {2 * C, C});
      up_nw[u] = add_param(std::string(buf) + ".norm.weight", {C / 2});
      up_nb[u] = add_param(std::string(buf) + ".norm.bias", {C / 2});
      up_lin[u] = add_linear(w, -1, 2 * C, C, 1, 0);
    }
  }
  {
    const long C = c.embed_dim << (L - 1);
    if (!c.patch_expanding) {
      const int w = add_param("first_patch_expanding.expand.weight", {2 * C, C, 1, 1});
      const int bb = add_param("first_patch_expanding.expand.bias", {2 * C});
      fpe_lin = add_linear(w, bb, 2 * C, C, 4, C / 2);
    } else {                                               // tulip.py:565
      const int w = add_param("first_patch_expanding.expand.weight", {2 * C, C});
      fpe_nw = add_param("first_patch_expanding.norm.weight", {C / 2});
      fpe_nb = add_param("first_patch_expanding.norm.bias", {C / 2});
      fpe_lin = add_linear(w, -1, 2 * C, C, 1, 0);
    }
  }
  for (int u = 0; u < L - 1; ++u) {                       // tulip.py:682-688
    const long C = c.embed_dim << (L - 2 - u);
    snprintf(buf, sizeof buf, "skip_connection_layers.%d", u);
    const int w = add_param(std::string(buf) + ".weight", {C, 2 * C});
    const int bb = add_param(std::string(buf) + ".bias", {C});
    skip_lin[u] = add_linear(w, bb, C, 2 * C, 1, 0);
  }
  const long E = c.embed_dim;
  slot_normup_w = add_param("norm_up.weight", {E});
  slot_normup_b = add_param("norm_up.bias", {E});
  slot_pe_w = add_param("patch_embed.proj.weight", {E, 1, c.patch_h, 8});
  slot_pe_b = add_param("patch_embed.proj.bias", {E});
  slot_pe_nw = add_param("patch_embed.norm.weight", {E});
  slot_pe_nb = add_param("patch_embed.norm.bias", {E});
  slot_dec_w = add_param("decoder_pred.weight", {1, E, 1, 1});
  if (!c.expanding_head) {
    const int w = add_param("ps_head.conv_expand.0.weight", {E * r * r, E, 1, 1});
    const int bb = add_param("ps_head.conv_expand.0.bias", {E * r * r});
    head_lin = add_linear(w, bb, E * r * r, E, r * r, E);
  } else {                                                 // tulip.py:582; columns already in (P1 P2 C) = shuffle-slot order
    const int w = add_param("final_patch_expanding.expand.weight", {E * r * r, E});
    slot_fh_nw = add_param("final_patch_expanding.norm.weight", {E});
    slot_fh_nb = add_param("final_patch_expanding.norm.bias", {E});
    head_lin = add_linear(w, -1, E * r * r, E, 1, 0);
  }

  // bf16 weight arena + fp32 aux arena (permuted biases)
  long woff = 0, foff = 0;
  for (Linear& l : linears) {
    l.w_off = woff; woff += align_up((long)l.N * l.K, 128);
    l.wt_off = woff; woff += align_up((long)l.N * l.K, 128);
    if (l.perm_R2 > 1 && l.slot_b >= 0) { l.pbias_off = foff; foff += align_up(l.N, 64); }
  }
  warena_elems = woff;
  faux_elems = foff > 0 ? foff : 64;
  return TULIP_OK;                                        // device buffers are allocated on the first forward
}

int tulip_net::ensure_device() {
  if (warena) return TULIP_OK;
  TULIP_CUDA(cudaMalloc(&warena, warena_elems * sizeof(bf16)));
  TULIP_CUDA(cudaMalloc(&faux, faux_elems * sizeof(float)));
  return TULIP_OK;
}

Plan tulip_net::plan(int B) const {
  Plan p;
  const bool save_pre = false;                 // the GELU pre-activation is recomputed in the backward pass (EPI_DGELU2), never stored
  Bump bump;
  const long E = cfg.embed_dim;
  const long T0 = (long)B * H0 * W0;
  auto act = [&](long rows, long cols) { return bump.take(rows * cols * 2); };
  p.pe_out = act(T0, E);
  p.blocks.resize(blocks.size());
  auto plan_block = [&](int bi) {
    const BlockDef& b = blocks[bi];
    const long T = (long)B * (H0 >> b.stage) * (W0 >> b.stage), C = E << b.stage;
    BlockBuf& bb = p.blocks[bi];
    bb.xn1 = act(T, C); bb.st1 = bump.take(T * 8);
    bb.qkv = act(T, 3 * C); bb.ao = act(T, C); bb.xmid = act(T, C);
    bb.xn2 = act(T, C); bb.st2 = bump.take(T * 8);
    bb.hpre = save_pre ? act(T, 4 * C) : -1; bb.hact = act(T, 4 * C); bb.xout = act(T, C);
  };
  p.xn_m.assign(L, -1); p.st_m.assign(L, -1); p.x_merged.assign(L, -1);
  for (int s = 0; s < L; ++s) {
    for (int bi : enc_blocks[s]) plan_block(bi);
    if (s < L - 1) {
      const long T = (long)B * (H0 >> s) * (W0 >> s), C = E << s;
      p.xn_m[s] = act(T / 4, 4 * C); p.st_m[s] = bump.take(T / 4 * 8); p.x_merged[s] = act(T / 4, 2 * C);
    }
  }
  {
    const long T = (long)B * (H0 >> (L - 1)) * (W0 >> (L - 1)), C = E << (L - 1);
    p.x_fpe = act(4 * T, C / 2);
    if (fpe_nw >= 0) { p.x_fpe_pre = act(4 * T, C / 2); p.st_fpe = bump.take(4 * T * 8); }
  }
  p.x_skip.assign(L - 1, -1); p.x_up.assign(L - 1, -1);
  p.x_up_pre.assign(L - 1, -1); p.st_upx.assign(L - 1, -1);
  for (int u = 0; u < L - 1; ++u) {
    const int s = L - u - 2;
    const long T = (long)B * (H0 >> s) * (W0 >> s), C = E << s;
    p.x_skip[u] = act(T, C);
    for (int bi : dec_blocks[u]) plan_block(bi);
    if (u < L - 2) p.x_up[u] = act(4 * T, C / 2);
    if (u < L - 2 && up_nw[u] >= 0) { p.x_up_pre[u] = act(4 * T, C / 2); p.st_upx[u] = bump.take(4 * T * 8); }
  }
  p.xn_up = act(T0, E); p.st_up = bump.take(T0 * 8);
  if (cfg.expanding_head) p.st_head = bump.take(T0 * r * r * 8);
  // backward scratch (T_s * C_s is largest at stage 0)
  p.gA = act(T0, E); p.gB = act(T0, E); p.scr_gs = act(T0, E); p.scr_gsm = act(T0, E);
  p.scr_dxn = act(T0, E); p.scr_do = act(T0, E); p.scr_dqkv = act(T0, 3 * E);
  const long big = (long)E * r * r > 4 * E ? (long)E * r * r : 4 * E;      // head dh [T0, E r^2] or block dh [T0, 4E]
  p.scr_big = act(T0, big);
  p.g_save.assign(L, -1);
  for (int s = 0; s < L - 1; ++s) p.g_save[s] = act((long)B * (H0 >> s) * (W0 >> s), E << s);
  p.loss_acc = bump.take(256);
  p.gscr_bytes = grad_scratch_bytes();
  p.gscr = bump.take(p.gscr_bytes);
  p.total = bump.off;
  return p;
}

// bytes of gradient-copy scratch the backward pass hands out (GRAD_COPIES copies of every small gradient, 64-float granules):
// mirrors the grad_scratch() calls of backward()
long tulip_net::grad_scratch_bytes() const {
  const long E = cfg.embed_dim;
  const int nbias = (2 * cfg.win_h - 1) * (2 * cfg.win_w - 1);
  auto take = [](long n) { return align_up(n * GRAD_COPIES, 64); };
  long fl = take(E) + take(2 * E) + take(11 * E);                  // decoder_pred.weight, norm_up, PatchEmbed
  if (cfg.expanding_head) fl += take(3 * E);                        // FinalPatchExpanding: [d(gamma) | d(beta) | d(decoder_pred.weight)]
  for (const BlockDef& b : blocks) {
    const long C = E << b.stage;
    fl += 2 * take(2 * C) + take((long)nbias * cfg.num_heads[b.stage]);
  }
  for (int s = 0; s + 1 < L; ++s) fl += take(2 * 4 * (E << s));    // PatchMerging norms
  if (cfg.patch_expanding)
    for (int s = 1; s < L; ++s) fl += take(2 * ((E << s) / 2));     // PatchExpanding norms (one per upsampling site, C/2 wide)
  return fl * 4 + 1024;
}

int tulip_net::upload_pack_table(const int64_t* offs, cudaStream_t st) {
  bool same = items_dev != nullptr && items_offsets_cache.size() == params.size();
  if (same)
    for (size_t i = 0; i < params.size(); ++i)
      if (items_offsets_cache[i] != offs[i]) { same = false; break; }
  if (same) return TULIP_OK;
  std::vector<PackItem> items;
  int tiles = 0;
  for (const Linear& l : linears) {
    PackItem it;
    it.src_off = offs[l.slot_w];
    it.dst_off = l.w_off;
    it.dstT_off = l.wt_off;
    it.rows = l.N; it.cols = l.K;
    it.perm_R2 = l.perm_R2; it.perm_Cc = l.perm_Cc;
    it.tile_begin = tiles;
    tiles += ((l.N + 31) / 32) * ((l.K + 31) / 32);
    items.push_back(it);
  }
  if (!items_dev) TULIP_CUDA(cudaMalloc(&items_dev, items.size() * sizeof(PackItem)));
  TULIP_CUDA(cudaMemcpyAsync(items_dev, items.data(), items.size() * sizeof(PackItem), cudaMemcpyHostToDevice, st));
  TULIP_CUDA(cudaStreamSynchronize(st));                  // `items` is a stack vector; happens once per parameter layout
  n_items = (int)items.size();
  n_tiles = tiles;
  items_offsets_cache.assign(offs, offs + params.size());
  return TULIP_OK;
}

#define RUN(call)                   \
  do {                              \
    if (live) {                     \
      prof_begin(st);               \
      int rc__ = (call);            \
      if (rc__ != TULIP_OK) return rc__; \
      prof_end(st);                 \
      ++kernel_launches;            \
    }                               \
  } while (0)

// tagged GEMM launches: the tag names the kernel function (template instantiation) that runs
#define RUN_NT(g, epi)                                                                         \
  do {                                                                                         \
    tag((g).a_mode == A_UNSHUFFLE ? K_NT_UNSHUFFLE : nt_tag(epi), 2.0 * (g).M * (g).N * (g).K, \
        2.0 * ((double)(g).M * (g).K + (double)(g).N * (g).K + (double)(g).M * (g).N));        \
    RUN(gemm_nt((g), (epi), st));                                                              \
  } while (0)
#define RUN_TN(g)                                                                              \
  do {                                                                                         \
    const GemmTNArgs& g__ = (g);                                                               \
    tag(g__.y_mode == A_UNSHUFFLE ? K_TN_UNSHUFFLE : K_TN, 2.0 * g__.M * g__.N * g__.K,        \
        2.0 * ((double)g__.M * g__.N + (double)g__.M * g__.K) + 4.0 * g__.N * g__.K);          \
    RUN(gemm_tn(g__, st));                                                                     \
  } while (0)

namespace {

struct Ctx {
  tulip_net* net;
  int B;
  const float* params; const int64_t* offs; float* grads;
  const float* drop; const int* win_mode;
  unsigned char* ws;
  cudaStream_t st;
  const float* P(int slot) const { return params + offs[slot]; }
  float* G(int slot) const { return grads + offs[slot]; }
  bf16* A(long off) const { return reinterpret_cast<bf16*>(ws + off); }
  float* F(long off) const { return reinterpret_cast<float*>(ws + off); }
  const bf16* W(const Linear& l) const { return net->warena + l.w_off; }
  const bf16* Wt(const Linear& l) const { return net->warena + l.wt_off; }
  const float* bias(const Linear& l) const {
    if (l.slot_b < 0) return nullptr;
    return l.pbias_off >= 0 ? net->faux + l.pbias_off : P(l.slot_b);
  }
};

GemmArgs nt_args(const bf16* A, long lda, const bf16* B, long ldb, int M, int N, int K, const float* bias, bf16* out, long ldo) {
  GemmArgs g;
  memset(&g, 0, sizeof g);
  g.A = A; g.lda = lda; g.K1 = K; g.B = B; g.ldb = ldb; g.M = M; g.N = N; g.K = K; g.bias = bias; g.out = out; g.ldo = ldo;
  return g;
}

GemmTNArgs tn_args(const bf16* dY, long ldy, const bf16* X, long ldx, int M, int N, int K, float* dW, float* db) {
  GemmTNArgs g;
  memset(&g, 0, sizeof g);
  g.dY = dY; g.ldy = ldy; g.X = X; g.ldx = ldx; g.K1 = K; g.M = M; g.N = N; g.K = K; g.dW = dW; g.lddw = K; g.db = db;
  g.perm_R2 = 1; g.perm_Cc = 1;
  const int tiles = (N / 96) * (K / 96);
  int splits = (2 * tulip_num_sms() + tiles - 1) / tiles;
  const int max_splits = (M + 255) / 256;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  g.splits = splits;
  return g;
}

void window_of(const tulip_net& n, const BlockDef& b, int mode, int* Mh, int* Mw, int* sh, int* sw) {
  const int L = n.cfg.win_h * n.cfg.win_w;
  if (mode == 0) { *Mh = n.cfg.win_h; *Mw = n.cfg.win_w; } else { *Mh = 1; *Mw = L; }
  if (!b.shift) { *sh = 0; *sw = 0; }
  else if (mode == 0) { *sh = n.cfg.win_h / 2; *sw = n.cfg.win_w / 2; }
  else { *sh = 0; *sw = L / 2; }
}

AttnArgs attn_args(const Ctx& c, const BlockDef& b, const bf16* qkv) {
  const tulip_net& n = *c.net;
  AttnArgs a;
  memset(&a, 0, sizeof a);
  a.qkv = qkv;
  a.bias_table = c.P(b.table);
  a.B = c.B; a.H = n.H0 >> b.stage; a.W = n.W0 >> b.stage; a.C = n.cfg.embed_dim << b.stage; a.heads = n.cfg.num_heads[b.stage];
  window_of(n, b, c.win_mode ? c.win_mode[b.index] : 0, &a.Mh, &a.Mw, &a.sh, &a.sw);
  a.masked = b.shift;
  a.bMh = n.cfg.win_h; a.bMw = n.cfg.win_w; a.nbias = (2 * a.bMh - 1) * (2 * a.bMw - 1);
  a.scale = 1.0f / sqrtf(32.0f);
  return a;
}

}  // namespace

int tulip_net::forward(int B, const float* params_, const int64_t* offs, const float* x_lo, const float* target,
                       const float* drop_scales, const int* win_mode, void* ws, float* pred, float* losses, cudaStream_t st) {
  TULIP_REQUIRE(B > 0, "tulip: empty batch");
  int rc = ensure_device();
  if (rc) return rc;
  rc = upload_pack_table(offs, st);
  if (rc) return rc;
  const Plan p = plan(B);
  Ctx c{this, B, params_, offs, nullptr, drop_scales, win_mode, reinterpret_cast<unsigned char*>(ws), st};
  const int E = cfg.embed_dim;
  cur_dir = 0; at(0, 0);

  // The weight repack is only needed by the first GEMM: it runs on the side stream next to PatchEmbed and the first LayerNorm.
  const bool use_side = !profiling && !side_stream_disabled();
  if (use_side && !side) TULIP_CUDA(cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking));
  sync_used = 0;
  const cudaStream_t pst = use_side ? side : st;
  if (use_side) {
    cudaEvent_t e = next_sync_event();
    cudaEventRecord(e, st);
    cudaStreamWaitEvent(side, e, 0);
  }
  const bool repack = params_version < 0 || params_version != packed_version || packed_from != params_;
  if (repack) {
    tag(K_PACK, 0, 8.0 * warena_elems / 2);
    RUN(pack_weights(params_, warena, items_dev, n_items, n_tiles, pst));
    for (const Linear& l : linears)
      if (l.pbias_off >= 0) RUN(permute_bias(c.P(l.slot_b), faux + l.pbias_off, l.N, l.perm_R2, l.perm_Cc, pst));
    packed_version = params_version; packed_from = params_;
  }
  bool pack_pending = use_side;
  auto join_pack = [&]() {
    if (!pack_pending) return;
    cudaEvent_t e = next_sync_event();
    cudaEventRecord(e, side);
    cudaStreamWaitEvent(st, e, 0);
    pack_pending = false;
  };

  {
    EmbedArgs e;
    memset(&e, 0, sizeof e);
    e.x = x_lo; e.w = c.P(slot_pe_w); e.b = c.P(slot_pe_b); e.ln_w = c.P(slot_pe_nw); e.ln_b = c.P(slot_pe_nb);
    e.y = c.A(p.pe_out); e.B = B; e.Himg = cfg.img_h; e.Wimg = cfg.img_w; e.ph = cfg.patch_h; e.E = E; e.eps = cfg.ln_eps;
    tag(K_EMBED_FWD, 0, 4.0 * B * cfg.img_h * cfg.img_w + 2.0 * B * H0 * W0 * E);
    RUN(patch_embed_fwd(e, st));
  }

  auto ln = [&](const bf16* x, int wslot, int bslot, bf16* y, float* stats, int rows, int C, int gather, int H2, int W2) {
    LnArgs a;
    memset(&a, 0, sizeof a);
    a.x = x; a.w = c.P(wslot); a.b = c.P(bslot); a.y = y; a.stats = stats; a.rows = rows; a.C = C; a.eps = cfg.ln_eps;
    a.gather = gather; a.H2 = H2; a.W2 = W2;
    tag(K_LN_FWD, 0, 4.0 * rows * C + 8.0 * rows);
    return layernorm_fwd(a, st);
  };

  // A LayerNorm that reads the output of a GEMM with whole rows in a tile (N = 96 / 192) runs in that GEMM's epilogue
  // (EPI_STORE_LN / EPI_RESID_LN): NextLn names the LayerNorm that follows, fuse_ln() switches the epilogue when the shape
  // qualifies and says whether it did, so the consumer skips its own launch.
  struct NextLn { int wslot = -1, bslot = -1; bf16* y = nullptr; float* stats = nullptr; };
  auto fuse_ln = [&](GemmArgs& g, int& epi, const NextLn& n) -> bool {
    // training only: the forward-only call at small batch is latency-bound and the longer epilogue costs more than the
    // stand-alone launch it removes (measured B = 1: 0.587 vs 0.564 ms); at B = 32 the step gains 0.2 %
    if (inference || !n.y || !gemm_nt_lnfwd_supported(g.M, g.N, g.K) || g.a_mode != A_PLAIN) return false;
    g.ln_w = c.P(n.wslot); g.ln_b = c.P(n.bslot); g.ln_y = n.y; g.ln_ystats = n.stats; g.ln_eps = cfg.ln_eps;
    epi = (epi == EPI_RESID) ? EPI_RESID_LN : EPI_STORE_LN;
    return true;
  };
  auto uses_wmsa = [&](int bi) {
    const BlockDef& b = blocks[bi];
    const AttnArgs a0 = attn_args(c, b, nullptr);
    return inference && wmsa_block_supported(B, H0 >> b.stage, W0 >> b.stage, E << b.stage, a0.heads, a0.Mh, a0.Mw);
  };
  auto ln1_of = [&](int bi) {                             // the LayerNorm block bi starts with, unless its fused kernel owns it
    NextLn n;
    if (bi >= 0 && !uses_wmsa(bi)) { n.wslot = blocks[bi].n1w; n.bslot = blocks[bi].n1b; n.y = c.A(p.blocks[bi].xn1); n.stats = c.F(p.blocks[bi].st1); }
    return n;
  };

  auto block_fwd = [&](int bi, const bf16* x_in, bool xn1_ready, const NextLn& next, bool* next_done) -> int {
    const BlockDef& b = blocks[bi];
    const BlockBuf& bb = p.blocks[bi];
    const int Hs = H0 >> b.stage, Ws = W0 >> b.stage, C = E << b.stage, T = B * Hs * Ws;
    const float* ds1 = drop_scales ? drop_scales + (long)(2 * b.index) * B : nullptr;
    const float* ds2 = drop_scales ? drop_scales + (long)(2 * b.index + 1) * B : nullptr;
    const bool mlp_fused = bb.hpre < 0 && mlp_block_supported(T, C);
    bool xn2_ready = false;
    *next_done = false;
    at(b.stage, 1);
    {
      // attention half as ONE kernel where the shape qualifies and no backward pass needs the intermediates (wmsa.cu)
      AttnArgs a0 = attn_args(c, b, nullptr);
      if (inference && wmsa_block_supported(B, Hs, Ws, C, a0.heads, a0.Mh, a0.Mw)) {
        join_pack();
        WmsaBlockArgs w;
        memset(&w, 0, sizeof w);
        w.x = x_in; w.y = c.A(bb.xmid); w.ln_w = c.P(b.n1w); w.ln_b = c.P(b.n1b);
        w.wqkv = c.W(linears[b.qkv]); w.bqkv = c.bias(linears[b.qkv]); w.wproj = c.W(linears[b.proj]); w.bproj = c.bias(linears[b.proj]);
        w.bias_table = a0.bias_table; w.row_scale = ds1;
        w.B = B; w.H = Hs; w.W = Ws; w.C = C; w.heads = a0.heads; w.Mh = a0.Mh; w.Mw = a0.Mw; w.sh = a0.sh; w.sw = a0.sw;
        w.masked = a0.masked; w.bMh = a0.bMh; w.bMw = a0.bMw; w.eps = cfg.ln_eps;
        tag(K_WMSA_FWD, 8.0 * T * C * C + 64.0 * T * C, 4.0 * T * C);
        RUN(wmsa_block_fwd(w, st));
        goto mlp_half;
      }
    }
    if (!xn1_ready) RUN(ln(x_in, b.n1w, b.n1b, c.A(bb.xn1), c.F(bb.st1), T, C, 0, 0, 0));
    join_pack();
    {
      const Linear& l = linears[b.qkv];
      GemmArgs g = nt_args(c.A(bb.xn1), C, c.W(l), C, T, 3 * C, C, c.bias(l), c.A(bb.qkv), 3 * C);
      RUN_NT(g, EPI_STORE);
    }
    {
      AttnArgs a = attn_args(c, b, c.A(bb.qkv));
      a.out = c.A(bb.ao);
      tag(K_ATTN_FWD, 64.0 * T * C, 8.0 * T * C);
      RUN(win_attn_fwd(a, st));
    }
    {
      const Linear& l = linears[b.proj];
      GemmArgs g = nt_args(c.A(bb.ao), C, c.W(l), C, T, C, C, c.bias(l), c.A(bb.xmid), C);
      g.aux = x_in; g.ldaux = C; g.row_scale = ds1; g.rows_per_sample = Hs * Ws;
      int epi = EPI_RESID;
      if (!mlp_fused) {
        NextLn n2; n2.wslot = b.n2w; n2.bslot = b.n2b; n2.y = c.A(bb.xn2); n2.stats = c.F(bb.st2);
        xn2_ready = fuse_ln(g, epi, n2);
      }
      RUN_NT(g, epi);
    }
  mlp_half:
    at(b.stage, 2);
    if (mlp_fused) {
      // MLP half as ONE kernel (mlp.cu); in training it also stores the LayerNorm output, its statistics and the activated
      // hidden tensor, which is everything the backward pass reads
      join_pack();
      MlpBlockArgs m;
      memset(&m, 0, sizeof m);
      m.x = c.A(bb.xmid); m.y = c.A(bb.xout); m.ln_w = c.P(b.n2w); m.ln_b = c.P(b.n2b);
      m.w1 = c.W(linears[b.fc1]); m.b1 = c.bias(linears[b.fc1]); m.w2 = c.W(linears[b.fc2]); m.b2 = c.bias(linears[b.fc2]);
      m.row_scale = ds2; m.rows_per_sample = Hs * Ws;
      if (!inference) { m.xn = c.A(bb.xn2); m.stats = c.F(bb.st2); m.hact = c.A(bb.hact); }
      m.T = T; m.C = C; m.eps = cfg.ln_eps;
      tag(K_MLP_FWD, 16.0 * T * C * C, (inference ? 4.0 : 14.0) * T * C);
      RUN(mlp_block_fwd(m, st));
      return TULIP_OK;
    }
    if (!xn2_ready) RUN(ln(c.A(bb.xmid), b.n2w, b.n2b, c.A(bb.xn2), c.F(bb.st2), T, C, 0, 0, 0));
    {
      const Linear& l = linears[b.fc1];
      GemmArgs g = nt_args(c.A(bb.xn2), C, c.W(l), C, T, 4 * C, C, c.bias(l), c.A(bb.hact), 4 * C);
      if (bb.hpre >= 0) { g.out2 = c.A(bb.hpre); g.ldo2 = 4 * C; }      // otherwise the backward recomputes it (EPI_DGELU2)
      RUN_NT(g, EPI_GELU);
    }
    {
      const Linear& l = linears[b.fc2];
      GemmArgs g = nt_args(c.A(bb.hact), 4 * C, c.W(l), 4 * C, T, C, 4 * C, c.bias(l), c.A(bb.xout), C);
      g.aux = c.A(bb.xmid); g.ldaux = C; g.row_scale = ds2; g.rows_per_sample = Hs * Ws;
      int epi = EPI_RESID;
      *next_done = fuse_ln(g, epi, next);
      RUN_NT(g, epi);
    }
    return TULIP_OK;
  };
  // the blocks of one stage in order; `first_ready`: LN1 of the first block was written by the producer of x; `after`: the
  // LayerNorm that follows the last block (or none)
  auto run_blocks = [&](const std::vector<int>& list, const bf16*& x, bool first_ready, const NextLn& after, bool* after_done) -> int {
    bool ready = first_ready;
    *after_done = false;
    for (size_t k = 0; k < list.size(); ++k) {
      const int bi = list[k];
      const NextLn next = (k + 1 < list.size()) ? ln1_of(list[k + 1]) : after;
      bool done = false;
      const int rc_ = block_fwd(bi, x, ready, next, &done);
      if (rc_) return rc_;
      x = c.A(p.blocks[bi].xout);
      ready = done;
    }
    if (!list.empty()) *after_done = ready;
    return TULIP_OK;
  };

  auto unmerge_fwd = [&](const Linear& l, const bf16* x, bf16* out, int Hs, int Ws, int C, int nw, int nb, long pre, long stats) -> int {
    // 1x1 conv C -> 2C + PixelShuffle(2) as a GEMM with a scatter epilogue (tulip.py:117-123); PatchExpanding (nw >= 0,
    // tulip.py:135-141) is the same scatter of an un-permuted, bias-free Linear into `pre`, followed by LayerNorm(C/2)
    GemmArgs g = nt_args(x, C, c.W(l), C, B * Hs * Ws, 2 * C, C, c.bias(l), nw >= 0 ? c.A(pre) : out, C / 2);
    g.g_H = Hs; g.g_W = Ws; g.g_Cc = C / 2;
    RUN_NT(g, EPI_PIXSHUF);
    if (nw >= 0) RUN(ln(c.A(pre), nw, nb, out, c.F(stats), 4 * B * Hs * Ws, C / 2, 0, 0, 0));
    return TULIP_OK;
  };

  std::vector<const bf16*> x_save(L);
  const bf16* x = c.A(p.pe_out);
  bool first_ready = false, after_done = false;
  const NextLn none;
  for (int s = 0; s < L; ++s) {
    x_save[s] = x;
    rc = run_blocks(enc_blocks[s], x, first_ready, none, &after_done);
    if (rc) return rc;
    first_ready = false;
    at(s, 0);
    if (s < L - 1) {
      const int Hs = H0 >> s, Ws = W0 >> s, C = E << s, T4 = B * Hs * Ws / 4;
      RUN(ln(x, merge_nw[s], merge_nb[s], c.A(p.xn_m[s]), c.F(p.st_m[s]), T4, 4 * C, 1, Hs / 2, Ws / 2));
      const Linear& l = linears[merge_lin[s]];
      GemmArgs g = nt_args(c.A(p.xn_m[s]), 4 * C, c.W(l), 4 * C, T4, 2 * C, 4 * C, nullptr, c.A(p.x_merged[s]), 2 * C);
      int epi = EPI_STORE;
      first_ready = fuse_ln(g, epi, ln1_of(enc_blocks[s + 1].empty() ? -1 : enc_blocks[s + 1][0]));
      RUN_NT(g, epi);
      x = c.A(p.x_merged[s]);
    }
  }
  at(L - 1, 0);
  rc = unmerge_fwd(linears[fpe_lin], x, c.A(p.x_fpe), H0 >> (L - 1), W0 >> (L - 1), E << (L - 1), fpe_nw, fpe_nb, p.x_fpe_pre, p.st_fpe);
  if (rc) return rc;
  x = c.A(p.x_fpe);
  for (int u = 0; u < L - 1; ++u) {
    const int s = L - u - 2;
    const int Hs = H0 >> s, Ws = W0 >> s, C = E << s, T = B * Hs * Ws;
    at(s, 0);
    {
      // Linear(2C -> C) on cat([x, x_save[s]], -1) without materialising the concat (tulip.py:715-716)
      const Linear& l = linears[skip_lin[u]];
      GemmArgs g = nt_args(x, C, c.W(l), 2 * C, T, C, 2 * C, c.bias(l), c.A(p.x_skip[u]), C);
      g.A2 = x_save[s]; g.lda2 = C; g.K1 = C;
      int epi = EPI_STORE;
      first_ready = fuse_ln(g, epi, ln1_of(dec_blocks[u].empty() ? -1 : dec_blocks[u][0]));
      RUN_NT(g, epi);
      x = c.A(p.x_skip[u]);
    }
    {
      NextLn after;                                       // norm_up follows the last decoder stage
      if (u == L - 2) { after.wslot = slot_normup_w; after.bslot = slot_normup_b; after.y = c.A(p.xn_up); after.stats = c.F(p.st_up); }
      rc = run_blocks(dec_blocks[u], x, first_ready, after, &after_done);
      if (rc) return rc;
      first_ready = false;
    }
    at(s, 0);
    if (u < L - 2) {
      rc = unmerge_fwd(linears[up_lin[u]], x, c.A(p.x_up[u]), Hs, Ws, C, up_nw[u], up_nb[u], p.x_up_pre[u], p.st_upx[u]);
      if (rc) return rc;
      x = c.A(p.x_up[u]);
    }
  }
  // norm_up -> ps_head -> decoder_pred, fused: the E*r^2-channel tensor never exists (tulip.py:720-731)
  const int T0 = B * H0 * W0;
  at(0, 3);                                               // part 3 = head + loss
  if (!(L >= 2 && after_done)) RUN(ln(x, slot_normup_w, slot_normup_b, c.A(p.xn_up), c.F(p.st_up), T0, E, 0, 0, 0));
  {
    const Linear& l = linears[head_lin];
    if (E != 96) TULIP_CUDA(cudaMemsetAsync(pred, 0, (size_t)T0 * r * r * sizeof(float), st));   // partial sums over channel groups
    GemmArgs g = nt_args(c.A(p.xn_up), E, c.W(l), E, T0, E * r * r, E, c.bias(l), nullptr, 0);
    g.wd = c.P(slot_dec_w); g.pred = pred; g.hd_H = H0; g.hd_W = W0; g.hd_r = r; g.hd_E = E;
    if (cfg.expanding_head) {                             // FinalPatchExpanding: LayerNorm per output pixel instead of bias + LeakyReLU
      g.hd_ln = 1; g.ln_w = c.P(slot_fh_nw); g.ln_b = c.P(slot_fh_nb); g.ln_eps = cfg.ln_eps; g.ln_ystats = c.F(p.st_head);
    }
    RUN_NT(g, EPI_HEAD);
  }
  tag(K_LOSS, 0, 8.0 * T0 * r * r);
  if (target) RUN(l1_loss(pred, target, (long)T0 * r * r, cfg.log_transform, c.F(p.loss_acc), losses, st));
  return TULIP_OK;
}

int tulip_net::backward(int B, const float* params_, const int64_t* offs, float* grads, const float* x_lo, const float* target,
                        const float* pred, const float* grad_loss, const float* drop_scales, const int* win_mode, void* ws,
                        cudaStream_t st, int phase_lo, int phase_hi) {
  TULIP_REQUIRE(B > 0 && target && pred && grad_loss, "tulip backward: needs target, pred and grad_loss");
  TULIP_REQUIRE(warena != nullptr, "tulip backward: no forward has run on this net");
  const Plan p = plan(B);
  Ctx c{this, B, params_, offs, grads, drop_scales, win_mode, reinterpret_cast<unsigned char*>(ws), st};
  const int E = cfg.embed_dim;
  const int T0 = B * H0 * W0;
  int rc;
  cur_dir = 1; at(0, 3);
  // The host walk below always covers the whole pass (buffer rotation and scratch offsets are the same in every call); `live`
  // says whether the launches of the part being walked are issued (backward phases, net.h).
  live = phase_lo <= 0;

  // Weight-gradient GEMMs (gemm_tn) are leaves of the backward graph: nothing on the dX chain reads them.  They run on a
  // side stream, forked after the kernel that produced their dY operand and joined before that buffer is overwritten, so
  // at the deep stages (few CTAs per kernel) they fill idle SMs and at every stage they leave the critical path.
  const bool use_side = !profiling && !side_stream_disabled();
  if (use_side && !side) TULIP_CUDA(cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking));
  sync_used = 0;
  bool side_pending = false;
  // zero the whole flat gradient span covered by the parameters.  Only the side-stream GEMMs and the final sum_copies write
  // into it (every other gradient goes to the scratch copies), so the 108 MB fill runs on the side stream under the head kernels.
  {
    long lo = offs[0], hi = offs[0] + params[0].numel;
    for (size_t i = 0; i < params.size(); ++i) {
      if (offs[i] < lo) lo = offs[i];
      if (offs[i] + params[i].numel > hi) hi = offs[i] + params[i].numel;
    }
    if (live) {
      if (use_side) {
        cudaEvent_t e = next_sync_event();
        cudaEventRecord(e, st);
        cudaStreamWaitEvent(side, e, 0);
        side_pending = true;
      }
      TULIP_CUDA(cudaMemsetAsync(grads + lo, 0, (size_t)(hi - lo) * sizeof(float), use_side ? side : st));
    }
  }
  auto fork = [&]() {                                     // side stream is ordered after everything issued on `st` so far
    if (!live) return;
    cudaEvent_t e = next_sync_event();
    cudaEventRecord(e, st);
    cudaStreamWaitEvent(side, e, 0);
  };
  // Weight-gradient GEMMs with plain operands are collected and issued together (gemm_tn_group: one persistent launch for
  // the two of a half-block); flush_tn() sends what has been collected, ordered after everything issued on `st` so far.
  // A deferred launch reads its operands later than the point it was registered at: callers flush before any kernel that
  // overwrites them is issued.
  std::vector<GemmTNArgs> tn_pending;
  int tn_rc = TULIP_OK;                                   // first error of a deferred launch (checked by TN_FLUSH / at the end)
  auto flush_tn_impl = [&]() -> int {
    if (tn_pending.empty()) return TULIP_OK;
    std::vector<GemmTNArgs> batch;
    batch.swap(tn_pending);
    if (!live) return TULIP_OK;
    const int nb = (int)batch.size();
    if (!use_side) {
      double fl = 0.0, by = 0.0;
      for (const GemmTNArgs& g_ : batch) {
        fl += 2.0 * g_.M * g_.N * g_.K;
        by += 2.0 * ((double)g_.M * g_.N + (double)g_.M * g_.K) + 4.0 * g_.N * g_.K;
      }
      tag(K_TN, fl, by);
      RUN(gemm_tn_group(batch.data(), nb, st));
      return TULIP_OK;
    }
    fork();
    const int rc_ = gemm_tn_group(batch.data(), nb, side);
    if (rc_ != TULIP_OK) return rc_;
    ++kernel_launches;
    side_pending = true;
    return TULIP_OK;
  };
  auto flush_tn = [&]() {
    const int rc_ = flush_tn_impl();
    if (rc_ != TULIP_OK && tn_rc == TULIP_OK) tn_rc = rc_;
  };
  auto join = [&]() {                                     // `st` waits for everything issued on the side stream so far
    flush_tn();
    if (!side_pending) return;
    cudaEvent_t e = next_sync_event();
    cudaEventRecord(e, side);
    cudaStreamWaitEvent(st, e, 0);
    side_pending = false;
  };
  // marker on the side stream after what has been issued there so far; wait_side(e) orders `st` after it (and nothing later)
  auto side_marker = [&]() -> cudaEvent_t {
    if (!live || !use_side || !side_pending) return nullptr;
    cudaEvent_t e = next_sync_event();
    cudaEventRecord(e, side);
    return e;
  };
  auto wait_side = [&](cudaEvent_t e) { if (e && live) cudaStreamWaitEvent(st, e, 0); };
  auto run_tn = [&](const GemmTNArgs& gw) -> int {        // one weight-gradient GEMM, on the side stream when enabled
    if (!live || debug_skip_tn()) return TULIP_OK;
    if (gemm_tn_groupable(gw)) {
      if ((int)tn_pending.size() == TN_GROUP_MAX) flush_tn();
      tn_pending.push_back(gw);
      if ((int)tn_pending.size() >= tn_group_limit()) flush_tn();
      return tn_rc;
    }
    flush_tn();
    if (tn_rc != TULIP_OK) return tn_rc;
    if (!use_side) { RUN_TN(gw); return TULIP_OK; }
    fork();
    const int rc_ = gemm_tn(gw, side);
    if (rc_ != TULIP_OK) return rc_;
    ++kernel_launches;
    side_pending = true;
    return TULIP_OK;
  };
#define TN_FLUSH()                                  \
  do {                                              \
    flush_tn();                                     \
    if (tn_rc != TULIP_OK) return tn_rc;            \
  } while (0)
#define TN_SIDE(g)                      \
  do {                                  \
    const int rc__ = run_tn(g);         \
    if (rc__ != TULIP_OK) return rc__;  \
  } while (0)

  // gradient copies (net.h GRAD_COPIES): grad_scratch(n) hands out GRAD_COPIES x n zeroed floats; sum_to() registers where
  // elements [off, off + n) of every copy are finally added
  if (live) TULIP_CUDA(cudaMemsetAsync(c.F(p.gscr), 0, (size_t)p.gscr_bytes, st));
  long gscr_used = 0;
  std::vector<SumCopiesItem> sum_items;
  bool gscr_overflow = false;
  auto grad_scratch = [&](int n) -> float* {
    float* ptr = c.F(p.gscr) + gscr_used;
    gscr_used += align_up((long)n * GRAD_COPIES, 64);
    if (gscr_used * 4 > p.gscr_bytes) { gscr_overflow = true; return c.F(p.gscr); }
    return ptr;
  };
  bool sum_clash = false;                                 // two folds into one destination would race inside sum_copies
  auto sum_to = [&](float* dst, const float* scr, int off, int n, int stride) {
    if (!live) return;
    for (const SumCopiesItem& it : sum_items)
      if (dst < it.dst + it.n && it.dst < dst + n) sum_clash = true;
    sum_items.push_back(SumCopiesItem{dst, scr + off, n, stride});
  };
  // fold the gradient copies registered so far into the flat gradient (end of a phase / end of the pass)
  auto flush_sums = [&]() -> int {
    for (size_t first = 0; first < sum_items.size(); first += 128) {   // 128 (dst, src) pairs per launch (kernel-parameter space)
      SumCopiesArgs sum_args;
      memset(&sum_args, 0, sizeof sum_args);
      sum_args.copies = GRAD_COPIES;
      sum_args.count = (int)std::min<size_t>(128, sum_items.size() - first);
      for (int i = 0; i < sum_args.count; ++i) sum_args.item[i] = sum_items[first + i];
      tag(K_ELEMWISE, 0, 8.0 * GRAD_COPIES * 64.0 * sum_args.count);
      RUN(sum_copies(sum_args, st));
    }
    sum_items.clear();
    return TULIP_OK;
  };
  // phase boundary: what the finished phase produced is complete on `st` (side stream joined, copies folded)
  auto enter_phase = [&](int ph) -> int {
    flush_tn();
    if (tn_rc != TULIP_OK) return tn_rc;
    if (live) {
      if (side_pending) { cudaEvent_t e = next_sync_event(); cudaEventRecord(e, side); cudaStreamWaitEvent(st, e, 0); side_pending = false; }
      const int rc_ = flush_sums();
      if (rc_) return rc_;
    }
    live = ph >= phase_lo && ph <= phase_hi;
    return TULIP_OK;
  };

  // optional fused DropPath scale for the consumer of dx: set before calling ln_bwd, consumed (reset) by it
  bf16* ln_dxs = nullptr; const float* ln_scale = nullptr; int ln_rps = 1;
  // Builds the arguments on every walk (live or not): the scratch offsets must not depend on the requested phases.
  auto ln_bwd_args = [&](const bf16* x, int wslot, int bslot, const float* stats, const bf16* dy, const bf16* dres, bf16* dx,
                         int rows, int C, int gather, int H2, int W2) -> LnArgs {
    LnArgs a;
    memset(&a, 0, sizeof a);
    a.dxs = ln_dxs; a.row_scale = ln_scale; a.rows_per_sample = ln_rps;
    ln_dxs = nullptr; ln_scale = nullptr;
    a.x = x; a.w = c.P(wslot); a.stats = const_cast<float*>(stats); a.dy = dy; a.dres = dres; a.dx = dx;
    a.rows = rows; a.C = C; a.eps = cfg.ln_eps; a.gather = gather; a.H2 = H2; a.W2 = W2;
    {
      float* scr = grad_scratch(2 * C);                    // per copy: [dgamma | dbeta]
      a.dw = scr; a.db = scr + C; a.dcopies = GRAD_COPIES; a.dstride = 2 * C;
      sum_to(c.G(wslot), scr, 0, C, 2 * C);
      sum_to(c.G(bslot), scr, C, C, 2 * C);
    }
    return a;
  };
  auto tag_ln_bwd = [&](const LnArgs& a) {
    tag(K_LN_BWD, 0, (a.dres ? 8.0 : 6.0) * a.rows * a.C + 8.0 * a.rows + (a.dxs ? 2.0 * a.rows * a.C : 0.0));
  };
#define LN_BWD(...)                               \
  do {                                            \
    const LnArgs la__ = ln_bwd_args(__VA_ARGS__); \
    tag_ln_bwd(la__);                             \
    RUN(layernorm_bwd(la__, st));                 \
  } while (0)
  // dX GEMM whose output rows are dL/dy of a LayerNorm: where a tile holds whole rows (C = 96 / 192) the LayerNorm backward
  // runs in the GEMM epilogue (EPI_LNBWD) and dy never reaches HBM; elsewhere the GEMM stores dy and layernorm_bwd follows.
  auto lnbwd_fused = [&](GemmArgs& g, const LnArgs& a) {
    g.aux = a.x; g.ldaux = a.C; g.aux2 = a.dres; g.ldaux2 = a.C;
    g.out = a.dx; g.ldo = a.C; g.out2 = a.dxs; g.ldo2 = a.C; g.row_scale = a.row_scale; g.rows_per_sample = a.rows_per_sample;
    g.ln_w = a.w; g.ln_stats = a.stats; g.ln_dw = a.dw; g.ln_db = a.db; g.ln_copies = a.dcopies; g.ln_stride = a.dstride;
  };
#define GEMM_LN_BWD(g, ...)                                                       \
  do {                                                                            \
    const LnArgs la__ = ln_bwd_args(__VA_ARGS__);                                 \
    if (la__.gather == 0 && gemm_nt_lnbwd_supported((g).M, (g).N, (g).K)) {       \
      lnbwd_fused((g), la__);                                                     \
      RUN_NT((g), EPI_LNBWD);                                                     \
    } else {                                                                      \
      RUN_NT((g), EPI_STORE);                                                     \
      tag_ln_bwd(la__);                                                           \
      RUN(layernorm_bwd(la__, st));                                               \
    }                                                                             \
  } while (0)
  auto dw = [&](const Linear& l, const bf16* dY, const bf16* X, int M) {
    GemmTNArgs g = tn_args(dY, l.N, X, l.K, M, l.N, l.K, c.G(l.slot_w), l.slot_b >= 0 ? c.G(l.slot_b) : nullptr);
    g.perm_R2 = l.perm_R2; g.perm_Cc = l.perm_R2 > 1 ? l.perm_Cc : 1;
    return g;
  };

  bf16* g_cur = c.A(p.gA);       // gradient w.r.t. the current activation
  bf16* g_alt = c.A(p.gB);
  // DropPath backward: the branch gradients are s * g.  Where g is produced by a LayerNorm backward the scaled copy is
  // written by that kernel (second output); gsM_ready says scr_gsm already holds s2(block) * g for the block about to run.
  bool gsM_ready = false;
  auto want_scaled_for = [&](int bi_next, int rows_per_sample) {      // call right before the ln_bwd that produces g for bi_next
    if (!drop_scales || bi_next < 0) return;
    ln_dxs = c.A(p.scr_gsm); ln_scale = drop_scales + (long)(2 * blocks[bi_next].index + 1) * B; ln_rps = rows_per_sample;
    gsM_ready = true;
  };

  // ---- head: dpred -> dh -> (dWe, dbe, dwd), dxn_up -> norm_up ----
  const bf16* x_last;
  {
    const int last_u = L - 2;
    const int bi = dec_blocks[last_u].back();
    x_last = c.A(p.blocks[bi].xout);
  }
  {
    const Linear& l = linears[head_lin];
    bf16* dh = c.A(p.scr_big);
    GemmArgs g = nt_args(c.A(p.xn_up), E, c.W(l), E, T0, E * r * r, E, c.bias(l), dh, (long)E * r * r);
    g.wd = c.P(slot_dec_w); g.pred = const_cast<float*>(pred); g.target = target; g.gscale = grad_loss;
    {
      float* scr = grad_scratch(E);                       // (handed out in both head forms: the scratch offsets that follow stay put)
      g.dwd = scr; g.dwd_copies = GRAD_COPIES;
      if (!cfg.expanding_head) sum_to(c.G(slot_dec_w), scr, 0, E, E);   // one fold per destination: sum_copies items must not share one
    }
    if (cfg.expanding_head) {                             // per copy: [d(gamma) | d(beta) | d(decoder_pred.weight)]
      float* scr = grad_scratch(3 * E);
      g.hd_ln = 1; g.ln_w = c.P(slot_fh_nw); g.ln_b = c.P(slot_fh_nb); g.ln_eps = cfg.ln_eps; g.ln_stats = c.F(p.st_head);
      g.ln_dw = scr; g.ln_db = scr + E; g.dwd = scr + 2 * E; g.ln_copies = GRAD_COPIES; g.ln_stride = 3 * E;
      sum_to(c.G(slot_fh_nw), scr, 0, E, 3 * E);
      sum_to(c.G(slot_fh_nb), scr, E, E, 3 * E);
      sum_to(c.G(slot_dec_w), scr, 2 * E, E, 3 * E);
    }
    g.hd_H = H0; g.hd_W = W0; g.hd_r = r; g.hd_E = E; g.hd_inv_npix = 1.0f / ((float)T0 * r * r);
    const bool fused_head = !cfg.expanding_head && head_bwd_fused_supported(E, r);
    if (fused_head) {
      // ONE launch (mlp.cu, MODE 1): recomputed pre-activation -> dh chunk in shared memory -> dxn accumulated over the 16
      // shuffle slots; dh leaves once for the weight-gradient GEMM and is never read back by a dX GEMM
      HeadBwdArgs hb;
      memset(&hb, 0, sizeof hb);
      hb.xn = c.A(p.xn_up); hb.dxn = c.A(p.scr_dxn); hb.we = c.W(l); hb.wet = c.Wt(l); hb.bias = c.bias(l); hb.wd = g.wd;
      hb.pred = pred; hb.target = target; hb.gscale = grad_loss; hb.dh = dh; hb.dwd = g.dwd; hb.dwd_copies = g.dwd_copies;
      hb.T = T0; hb.E = E; hb.H = H0; hb.W = W0; hb.r = r;
      tag(K_NT_HEAD_BWD, 4.0 * T0 * E * r * r * E, 2.0 * T0 * E * (2.0 + r * r));
      RUN(head_bwd_fused(hb, st));
    } else {
      RUN_NT(g, EPI_HEAD_BWD);
    }
    // dWe' += dh^T . xn_up (rows un-permuted on store); bias gradient = column sums of dh, same un-permutation
    GemmTNArgs gw = dw(l, dh, c.A(p.xn_up), T0);
    TN_SIDE(gw);
    TN_FLUSH();
    want_scaled_for(dec_blocks[L - 2].back(), H0 * W0);
    if (fused_head) {
      LN_BWD(x_last, slot_normup_w, slot_normup_b, c.F(p.st_up), c.A(p.scr_dxn), nullptr, g_cur, T0, E, 0, 0, 0);
    } else {
      // dxn_up = dh . We'          (A = dh [T0, E r^2], B = We'^T stored as Wt' [E, E r^2])
      GemmArgs gx = nt_args(dh, (long)E * r * r, c.Wt(l), (long)E * r * r, T0, E, E * r * r, nullptr, c.A(p.scr_dxn), E);
      GEMM_LN_BWD(gx, x_last, slot_normup_w, slot_normup_b, c.F(p.st_up), c.A(p.scr_dxn), nullptr, g_cur, T0, E, 0, 0, 0);
    }
  }

  auto block_bwd = [&](int bi, const bf16* x_in, bf16* g_io, bf16* g_tmp, int bi_next) -> int {
    // g_io holds dL/dx_out on entry and dL/dx_in on exit; g_tmp is a same-sized scratch
    const BlockDef& b = blocks[bi];
    const BlockBuf& bb = p.blocks[bi];
    const int Hs = H0 >> b.stage, Ws = W0 >> b.stage, C = E << b.stage, T = B * Hs * Ws;
    const float* ds1 = drop_scales ? drop_scales + (long)(2 * b.index) * B : nullptr;
    const float* ds2 = drop_scales ? drop_scales + (long)(2 * b.index + 1) * B : nullptr;
    // ---- MLP half: x_out = x_mid + s2 * fc2(gelu(fc1(LN2(x_mid)))) ----
    at(b.stage, 2);
    join();                                               // side work issued outside blocks still reads the g buffers
    const bf16* gy = g_io;
    if (ds2) {
      if (!gsM_ready) {                                   // g came from a GEMM epilogue: scale it here
        tag(K_ELEMWISE, 0, 4.0 * T * C);
        RUN(scale_rows_bf16(c.A(p.scr_gsm), g_io, ds2, T, C, Hs * Ws, st));
      }
      gy = c.A(p.scr_gsm);
    }
    gsM_ready = false;
    cudaEvent_t mlp_dw_done = nullptr;
    {
      const Linear& l2 = linears[b.fc2];
      const Linear& lf1 = linears[b.fc1];
      TN_SIDE(dw(l2, gy, c.A(bb.hact), T));
      GemmArgs g = nt_args(gy, C, c.Wt(l2), C, T, 4 * C, C, nullptr, c.A(p.scr_big), 4 * C);   // dh = (gy . W2) o gelu'(pre)
      if (bb.hpre >= 0) {
        g.aux = c.A(bb.hpre); g.ldaux = 4 * C;
        RUN_NT(g, EPI_DGELU);
      } else {
        // pre = xn2 . W1^T + b1 is recomputed by a second accumulator of the same kernel: 2x the (cheap, HBM-bound)
        // MMA work instead of writing and re-reading the [T, 4C] pre-activation
        g.A2 = c.A(bb.xn2); g.lda2 = C; g.B2 = c.W(lf1); g.ldb2 = C; g.K2 = C; g.bias = c.bias(lf1);
        RUN_NT(g, EPI_DGELU2);
      }
      const Linear& l1 = linears[b.fc1];
      TN_SIDE(dw(l1, c.A(p.scr_big), c.A(bb.xn2), T));
      TN_FLUSH();                                         // fc2 + fc1 weight gradients: one launch beside the LayerNorm-backward GEMM
      mlp_dw_done = side_marker();
      GemmArgs g1 = nt_args(c.A(p.scr_big), 4 * C, c.Wt(l1), 4 * C, T, C, 4 * C, nullptr, c.A(p.scr_dxn), C);
      if (ds1) { ln_dxs = c.A(p.scr_gs); ln_scale = ds1; ln_rps = Hs * Ws; }     // scaled copy for the attention branch
      GEMM_LN_BWD(g1, c.A(bb.xmid), b.n2w, b.n2b, c.F(bb.st2), c.A(p.scr_dxn), g_io, g_tmp, T, C, 0, 0, 0);   // g_tmp = dL/dx_mid
    }
    // ---- attention half: x_mid = x_in + s1 * proj(attn(qkv(LN1(x_in)))) ----
    at(b.stage, 1);
    gy = ds1 ? c.A(p.scr_gs) : g_tmp;
    {
      const Linear& lp = linears[b.proj];
      TN_SIDE(dw(lp, gy, c.A(bb.ao), T));
      GemmArgs g = nt_args(gy, C, c.Wt(lp), C, T, C, C, nullptr, c.A(p.scr_do), C);
      RUN_NT(g, EPI_STORE);
      AttnArgs a = attn_args(c, b, c.A(bb.qkv));
      a.dout = c.A(p.scr_do); a.dqkv = c.A(p.scr_dqkv);
      {
        const int n = a.nbias * a.heads;
        float* scr = grad_scratch(n);
        a.dbias_table = scr; a.dbias_copies = GRAD_COPIES;
        sum_to(c.G(b.table), scr, 0, n, n);
      }
      tag(K_ATTN_BWD, 160.0 * T * C, 16.0 * T * C);
      RUN(win_attn_bwd(a, st));
      const Linear& lq = linears[b.qkv];
      TN_SIDE(dw(lq, c.A(p.scr_dqkv), c.A(bb.xn1), T));
      TN_FLUSH();                                         // proj + qkv weight gradients: one launch, reads neither buffer written below
      wait_side(mlp_dw_done);                             // the LayerNorm backward below overwrites g_io / the scaled copies (fc2's dY)
      GemmArgs gq = nt_args(c.A(p.scr_dqkv), 3 * C, c.Wt(lq), 3 * C, T, C, 3 * C, nullptr, c.A(p.scr_dxn), C);
      want_scaled_for(bi_next, Hs * Ws);                  // next block in backward order lives on the same grid
      GEMM_LN_BWD(gq, x_in, b.n1w, b.n1b, c.F(bb.st1), c.A(p.scr_dxn), g_tmp, g_io, T, C, 0, 0, 0);            // g_io = dL/dx_in
    }
    return TULIP_OK;
  };

  auto unmerge_bwd = [&](const Linear& l, const bf16* x_in, const bf16* g_out, bf16* g_in, int Hs, int Ws, int C, int nw, int nb,
                         long pre, long stats) -> int {
    // g_out: [B, 2Hs, 2Ws, C/2]; gathered view A[m, ij*C/2 + c] (PixelShuffle backward), then dX and dW
    const int T = B * Hs * Ws;
    join();
    if (nw >= 0) {                                        // PatchExpanding: back through LayerNorm(C/2) first
      LN_BWD(c.A(pre), nw, nb, c.F(stats), g_out, nullptr, c.A(p.scr_do), 4 * T, C / 2, 0, 0, 0);
      g_out = c.A(p.scr_do);
    }
    GemmTNArgs gw = dw(l, g_out, x_in, T);
    gw.ldy = C / 2; gw.y_mode = A_UNSHUFFLE; gw.g_H = Hs; gw.g_W = Ws; gw.g_Cc = C / 2;
    TN_SIDE(gw);
    GemmArgs g = nt_args(g_out, C / 2, c.Wt(l), 2 * C, T, C, 2 * C, nullptr, g_in, C);
    g.a_mode = A_UNSHUFFLE; g.g_H = Hs; g.g_W = Ws; g.g_Cc = C / 2;
    RUN_NT(g, EPI_STORE);
    return TULIP_OK;
  };

  // ---- decoder, last stage first ----
  for (int u = L - 2; u >= 0; --u) {
    const int s = L - u - 2;
    const int Hs = H0 >> s, Ws = W0 >> s, C = E << s, T = B * Hs * Ws;
    at(s, 0);
    if (u < L - 2) {
      const bf16* x_before = c.A(p.blocks[dec_blocks[u].back()].xout);
      rc = unmerge_bwd(linears[up_lin[u]], x_before, g_cur, g_alt, Hs, Ws, C, up_nw[u], up_nb[u], p.x_up_pre[u], p.st_upx[u]);
      if (rc) return rc;
      std::swap(g_cur, g_alt);
    }
    for (int k = (int)dec_blocks[u].size() - 1; k >= 0; --k) {
      const int bi = dec_blocks[u][k];
      const bf16* x_in = k > 0 ? c.A(p.blocks[dec_blocks[u][k - 1]].xout) : c.A(p.x_skip[u]);
      rc = block_bwd(bi, x_in, g_cur, g_alt, k > 0 ? dec_blocks[u][k - 1] : -1);
      if (rc) return rc;
    }
    at(s, 0);
    {
      // skip Linear backward: d[x | skip] = g . Wskip ; the skip half is parked until the encoder stage is reached
      const Linear& l = linears[skip_lin[u]];
      const bf16* x_prev = (u == 0) ? c.A(p.x_fpe) : c.A(p.x_up[u - 1]);
      const bf16* x_enc = (s == 0) ? c.A(p.pe_out) : c.A(p.x_merged[s - 1]);
      join();
      GemmTNArgs gw = dw(l, g_cur, x_prev, T);
      gw.ldx = C; gw.X2 = x_enc; gw.ldx2 = C; gw.K1 = C;
      TN_SIDE(gw);
      GemmArgs g = nt_args(g_cur, C, c.Wt(l), C, T, 2 * C, C, nullptr, g_alt, C);
      g.out2 = c.A(p.g_save[s]); g.ldo2 = C; g.split_col = C;
      RUN_NT(g, EPI_SPLIT2);
      std::swap(g_cur, g_alt);
    }
  }
  // ---- first_patch_expanding ----
  {
    const int s = L - 1;
    at(s, 0);
    const bf16* x_top = c.A(p.blocks[enc_blocks[s].back()].xout);
    rc = unmerge_bwd(linears[fpe_lin], x_top, g_cur, g_alt, H0 >> s, W0 >> s, E << s, fpe_nw, fpe_nb, p.x_fpe_pre, p.st_fpe);
    if (rc) return rc;
    std::swap(g_cur, g_alt);
  }
  // ---- encoder, top stage first ----
  for (int s = L - 1; s >= 0; --s) {
    const int Hs = H0 >> s, Ws = W0 >> s, C = E << s, T = B * Hs * Ws;
    if (s == L - 1) { rc = enter_phase(1); if (rc) return rc; }
    if (s == L - 2) { rc = enter_phase(2); if (rc) return rc; }
    at(s, 0);
    if (s < L - 1) {
      // PatchMerging backward: g_cur is dL/d(x_merged[s]) [T/4, 2C]
      const Linear& l = linears[merge_lin[s]];
      const bf16* x_stage_out = c.A(p.blocks[enc_blocks[s].back()].xout);
      join();
      TN_SIDE(dw(l, g_cur, c.A(p.xn_m[s]), T / 4));
      TN_FLUSH();
      GemmArgs g = nt_args(g_cur, 2 * C, c.Wt(l), 2 * C, T / 4, 4 * C, 2 * C, nullptr, c.A(p.scr_big), 4 * C);
      RUN_NT(g, EPI_STORE);
      want_scaled_for(enc_blocks[s].back(), (Hs / 2) * (Ws / 2));       // rows here are merged (2x2) tokens
      LN_BWD(x_stage_out, merge_nw[s], merge_nb[s], c.F(p.st_m[s]), c.A(p.scr_big), nullptr, g_alt, T / 4, 4 * C, 1, Hs / 2,
                 Ws / 2);
      std::swap(g_cur, g_alt);
    }
    for (int k = (int)enc_blocks[s].size() - 1; k >= 0; --k) {
      const int bi = enc_blocks[s][k];
      const bf16* x_in = k > 0 ? c.A(p.blocks[enc_blocks[s][k - 1]].xout) : (s == 0 ? c.A(p.pe_out) : c.A(p.x_merged[s - 1]));
      rc = block_bwd(bi, x_in, g_cur, g_alt, k > 0 ? enc_blocks[s][k - 1] : -1);
      if (rc) return rc;
    }
    at(s, 0);
    tag(K_ELEMWISE, 0, 6.0 * T * C);
    if (s < L - 1) RUN(add_inplace_bf16(g_cur, c.A(p.g_save[s]), (long)T * C, st));     // skip-connection gradient (tulip.py:708,715)
  }
  {
    EmbedArgs e;
    memset(&e, 0, sizeof e);
    e.x = x_lo; e.w = c.P(slot_pe_w); e.b = c.P(slot_pe_b); e.ln_w = c.P(slot_pe_nw); e.ln_b = c.P(slot_pe_nb);
    e.B = B; e.Himg = cfg.img_h; e.Wimg = cfg.img_w; e.ph = cfg.patch_h; e.E = E; e.eps = cfg.ln_eps;
    e.dy = g_cur;
    {
      float* scr = grad_scratch(11 * E);                   // per copy: [dW (E x 8) | db | dgamma | dbeta]
      e.dw = scr; e.db = scr + 8 * E; e.dln_w = scr + 9 * E; e.dln_b = scr + 10 * E; e.dcopies = GRAD_COPIES; e.dstride = 11 * E;
      sum_to(c.G(slot_pe_w), scr, 0, 8 * E, 11 * E);
      sum_to(c.G(slot_pe_b), scr, 8 * E, E, 11 * E);
      sum_to(c.G(slot_pe_nw), scr, 9 * E, E, 11 * E);
      sum_to(c.G(slot_pe_nb), scr, 10 * E, E, 11 * E);
    }
    tag(K_EMBED_BWD, 0, 4.0 * B * cfg.img_h * cfg.img_w + 2.0 * B * H0 * W0 * E);
    RUN(patch_embed_bwd(e, st));
  }
  TULIP_REQUIRE(!gscr_overflow, "tulip_b200: gradient-copy scratch exhausted (grad_scratch_bytes() out of step with backward())");
  TULIP_REQUIRE(!sum_clash, "tulip_b200: two gradient-copy folds share a destination (sum_copies items must be disjoint)");
  TN_FLUSH();
  join();                                                 // the gradient fill and every weight-gradient GEMM precede the fold
  if (live) {
    rc = flush_sums();
    if (rc) return rc;
  }
  join();                                                 // every gradient is complete on `st` when backward returns
  live = true;
#undef TN_SIDE
#undef TN_FLUSH
#undef LN_BWD
#undef GEMM_LN_BWD
  return TULIP_OK;
}
