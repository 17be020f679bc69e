// Internal launcher declarations (host side) for the non-GEMM kernels.
#pragma once
#include "common.cuh"

struct AttnArgs {
  const bf16* qkv;            // [B*H*W, 3C] natural NHWC token order, feature f = t*C + head*32 + d (tulip.py:298)
  bf16* out;                  // fwd: [B*H*W, C] head-major concat (tulip.py:317), natural token order
  const bf16* dout;           // bwd: grad of `out`
  bf16* dqkv;                 // bwd: [B*H*W, 3C]
  const float* bias_table;    // [nbias, heads] fp32 (the nn.Parameter itself)
  float* dbias_table;         // bwd: [nbias, heads] fp32, accumulated
  int dbias_copies;           // bwd: CTAs spread their atomics over this many copies of the table (stride nbias*heads), >= 1
  int B, H, W, C, heads;
  int Mh, Mw;                 // window used for partitioning (backup window when H < win_h)
  int sh, sw;                 // cyclic shift (0,0 for W-MSA)
  int masked;                 // 1 for shifted blocks (mask is built even when a shift component is 0)
  int bMh, bMw, nbias;        // window the bias index buffer was built for (never rebuilt, tulip.py:228-240)
  float scale;
  int hint;                   // set by the launcher: q|k|v / dO rows are read once here, load them with the L2 evict_first priority
};
int win_attn_fwd(const AttnArgs& a, cudaStream_t st);
int win_attn_bwd(const AttnArgs& a, cudaStream_t st);

// fused W-MSA / SW-MSA half-block (wmsa.cu): y = x + row_scale[b] * proj(attn(qkv(LN1(x))))   (tulip.py:338-346, 282-324)
struct WmsaBlockArgs {
  const bf16* x; bf16* y;                 // [B*H*W, C] natural NHWC token order
  const float* ln_w; const float* ln_b;   // norm1
  const bf16* wqkv; const float* bqkv;    // [3C, C] bf16 (feature f = t*C + head*32 + d), [3C]
  const bf16* wproj; const float* bproj;  // [C, C] bf16, [C]
  const float* bias_table;                // [nbias, heads] fp32
  const float* row_scale;                 // [B] DropPath scales or null
  int B, H, W, C, heads;
  int Mh, Mw, sh, sw, masked, bMh, bMw;   // as AttnArgs
  float eps;
};
bool wmsa_block_supported(int B, int H, int W, int C, int heads, int Mh, int Mw);
int wmsa_block_fwd(const WmsaBlockArgs& a, cudaStream_t st);

// fused MLP half-block (mlp.cu): y = x + row_scale[b] * fc2(gelu(fc1(LN2(x))))   (tulip.py:347-352, 194-200)
struct MlpBlockArgs {
  const bf16* x; bf16* y;                 // [T, C]
  const float* ln_w; const float* ln_b;   // norm2
  const bf16* w1; const float* b1;        // [4C, C] bf16, [4C]
  const bf16* w2; const float* b2;        // [C, 4C] bf16, [C]
  const float* row_scale; int rows_per_sample;   // per-sample DropPath scale or null
  // training by-products (all three or none): LayerNorm output [T, C], its (mean, rstd) [T, 2], activated hidden tensor [T, 4C]
  bf16* xn; float* stats; bf16* hact;
  int T, C;
  float eps;
};
bool mlp_block_supported(int T, int C);
int mlp_block_fwd(const MlpBlockArgs& a, cudaStream_t st);

// fused head backward (mlp.cu, MODE 1) for PixelShuffleHead + decoder_pred + L1 at embed_dim 96, r = 4:
//   dh [T, E r^2] (bf16, n' = ij*E + c order) = dpred[m, ij] * wd[c] * LeakyReLU'(xn . We'^T + bias'),  dxn = dh . We',
//   dwd += colsum over (m, ij) of dpred * LeakyReLU(pre);  dpred = sign(pred - target) * gscale / (T r^2)
struct HeadBwdArgs {
  const bf16* xn; bf16* dxn;              // norm_up output [T, E], its gradient [T, E]
  const bf16* we; const bf16* wet;        // We' [E r^2, E] (rows in shuffle-slot order) and its transpose [E, E r^2]
  const float* bias;                      // bias' [E r^2], same order
  const float* wd;                        // decoder_pred.weight [E]
  const float* pred; const float* target; const float* gscale;
  bf16* dh;                               // out: [T, E r^2], read by the weight-gradient GEMM
  float* dwd; int dwd_copies;             // CTA b adds into copy b % dwd_copies (E floats apart)
  int T, E, H, W, r;
};
bool head_bwd_fused_supported(int E, int r);
int head_bwd_fused(const HeadBwdArgs& a, cudaStream_t st);

struct LnArgs {
  const bf16* x;              // LN input rows [rows, C]; with gather: source tensor [B, 2*H2, 2*W2, C/4]
  const float* w; const float* b;
  bf16* y;                    // [rows, C]
  float* stats;               // [rows, 2] (mean, rstd)
  int rows, C; float eps;
  int gather; int H2, W2;     // PatchMerging 2x2 gather (tulip.py:92-99): rows = B*H2*W2, C = 4*Csrc
  // backward
  const bf16* dy; const bf16* dres; bf16* dx; float* dw; float* db;
  // optional second output: dxs = row_scale[sample] * dx  (DropPath backward scale for the branch that consumes dx next)
  bf16* dxs; const float* row_scale; int rows_per_sample;
  // dw / db may point at `dcopies` scratch copies `dstride` floats apart: CTA b accumulates into copy b % dcopies
  // (same-address atomics from every CTA of the grid serialise in L2); 0 or 1 = accumulate in place
  int dcopies; int dstride;
};
int layernorm_fwd(const LnArgs& a, cudaStream_t st);
int layernorm_bwd(const LnArgs& a, cudaStream_t st);

struct EmbedArgs {
  const float* x;             // [B, 1, Himg, Wimg] fp32
  const float* w; const float* b;       // conv weight [E,1,ph,8], bias [E]
  const float* ln_w; const float* ln_b;
  bf16* y;                    // [B, Himg/ph, Wimg/4, E]
  int B, Himg, Wimg, ph, E; float eps;
  const bf16* dy; float* dw; float* db; float* dln_w; float* dln_b;
  int dcopies; int dstride;   // as in LnArgs, applied to all four gradient pointers
};
int patch_embed_fwd(const EmbedArgs& a, cudaStream_t st);
int patch_embed_bwd(const EmbedArgs& a, cudaStream_t st);

// table-driven fp32 -> bf16 weight repack (one launch for the whole model)
struct PackItem {
  long src_off;               // element offset into the fp32 flat parameter buffer
  long dst_off;               // element offset into the bf16 arena: W' [rows, cols] (rows permuted)
  long dstT_off;              // element offset of the transposed copy Wt' [cols, rows] or -1
  int rows, cols;
  int perm_R2, perm_Cc;       // destination row n' = ij*Cc + c  <-  source row c*R2 + ij   (R2 == 1: identity)
  int tile_begin;             // first 32x32 tile index of this item
};
int pack_weights(const float* flat, bf16* arena, const PackItem* items_dev, int n_items, int n_tiles, cudaStream_t st);
int permute_bias(const float* src, float* dst, int n, int R2, int Cc, cudaStream_t st);

int add_inplace_bf16(bf16* dst, const bf16* src, long n, cudaStream_t st);
// dst_j[i] += sum_c src_j[c * n_j + i] for up to 64 (dst, src, n) triples in one launch
struct SumCopiesItem { float* dst; const float* src; int n; int stride; };   // copy c of element i is src[c * stride + i]
struct SumCopiesArgs { SumCopiesItem item[128]; int count; int copies; };
int sum_copies(const SumCopiesArgs& a, cudaStream_t st);
int scale_rows_bf16(bf16* dst, const bf16* src, const float* row_scale, int rows, int C, int rows_per_sample, cudaStream_t st);
// evaluation post-processing (engine_upsampling.py:174-244): losses [B][2] = {pixel loss, low-res-row loss}, scratch [2B] floats
int eval_postprocess(const float* pred, const float* lo, const float* hi, float* out, float* losses, float* scratch, int B, int H, int W,
                     int h_lo, int log_transform, float clip_lo, int keep_low_res, cudaStream_t st);
// Monte-Carlo-dropout aggregation (engine_upsampling.py:423-427): preds [n, npix] -> out [npix] (mean, zeroed where std > threshold * mean)
int mc_aggregate(const float* preds, float* out, float* std_out, int n, long npix, float threshold, cudaStream_t st);
int l1_loss(const float* pred, const float* target, long n, int log_transform, float* acc2, float* out2, cudaStream_t st);
int stage_inputs(const float* const* src, float* const* dst, const long* n, cudaStream_t st);   // up to three fp32 copies, one launch

// stand-alone index ops (bit-exact tests of the index arithmetic used inside the fused kernels)
int window_gather(const bf16* x, bf16* out, int B, int H, int W, int C, int Mh, int Mw, int sh, int sw, cudaStream_t st);
int window_scatter(const bf16* xw, bf16* out, int B, int H, int W, int C, int Mh, int Mw, int sh, int sw, cudaStream_t st);
int shift_mask(float* out, int H, int W, int Mh, int Mw, int sh, int sw, cudaStream_t st);
int rel_bias_gather(const float* table, float* out, int heads, int Mh, int Mw, cudaStream_t st);
int merge_gather(const bf16* x, bf16* out, int B, int H, int W, int C, cudaStream_t st);
int pixel_shuffle_nhwc(const bf16* x, bf16* out, int B, int H, int W, int Cout, int r, cudaStream_t st);

// evaluation metrics (metrics.cu; reference tulip/util/evaluation.py)
int range_to_points(const float* img, const float* sin_h, const float* cos_h, const float* sin_v, const float* cos_v, float max_range,
                    float* points, int B, int H, int W, cudaStream_t st);
long voxel_metrics_workspace_bytes(int n);
int voxel_metrics(const float* pts_pred, const float* pts_gt, int n, float grid_size, void* workspace, double* out4, cudaStream_t st);
int chamfer_distance(const float* a, const float* b, int na, int nb, float* dist_a, float* dist_b, float* out3, cudaStream_t st);

// input pipeline (reference tulip/util/datasets.py transform chains): raw [B,H,W,channels] -> hi [B,1,H,W], lo [B,1,H/rf,W/cf]
int preprocess_range(const float* raw, int channels, float scale, int filter, float min_range, float max_range, int row_factor,
                     int col_factor, int log_transform, float* hi, float* lo, int B, int H, int W, cudaStream_t st);

// CARLA .rimg payload (size1 rows of size0 fp16 values per frame) -> frames [B, size0, size1] fp32 (datasets.py:181-193)
int rimg_decode(const void* rows_f16, float* frame, int B, int size0, int size1, cudaStream_t st);

// fused AdamW / gradient norm over the flat buffers (layout mirrored by tulip_adamw_segment / tulip_adamw_hyper in the C header)
struct AdamwSegment { long offset; long numel; int group; int pad; };
struct AdamwHyper {
  float lr[64]; float weight_decay[64];          // per parameter group (param_groups_layer_decay makes ~2 * (layers + 2) of them)
  float beta1, beta2, eps, bias_correction1, bias_correction2_sqrt, grad_scale;
};
int adamw_step(float* p, const float* g, float* m, float* v, const AdamwSegment* segs_dev, int n_segs, long span, const AdamwHyper& hp,
               cudaStream_t st);
int grad_norm(const float* g, const AdamwSegment* segs_dev, int n_segs, long span, double* scratch, float* out, cudaStream_t st);
int voxel_metrics_f64(const double* pts_pred, const double* pts_gt, int n, double grid_size, void* workspace, double* out4, cudaStream_t st);
int range_to_points_durlar(const float* img, const double* ca, const double* sa, const double* ce, const double* se, const double* cel,
                           const double* sel, const int* offset_lut, float max_range, double origin_offset, double z_offset,
                           double* points, int B, int H, int W, cudaStream_t st);
