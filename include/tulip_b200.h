/* tulip_b200 -- C ABI of the B200-native TULIP Swin forward/backward path.
 *
 * The reference (ethz-asl/TULIP) has no FFI: its hot path is the Python nn.Module in
 * tulip/model/tulip.py, driven by engine_upsampling.py:77-80 (train), :168-171 (eval), :417-419 (MC).
 * This header is the boundary the replacement is built behind: plain C, raw device pointers, explicit
 * sizes and a cudaStream_t -- no torch types.  `tulip_b200/model/tulip.py` binds it with ctypes and
 * re-exposes the reference's module API (same constructor kwargs, forward signature, state_dict).
 *
 * Conventions
 *   - every function returns 0 (TULIP_OK) or an error code; tulip_last_error() gives the message;
 *     nothing here calls exit()/abort();
 *   - all pointers are DEVICE pointers unless the name ends in _host;
 *   - activations are bf16 NHWC token-major [B*H*W, C]; parameters and gradients are fp32;
 *   - kernels are enqueued on `stream` (passed as void* = cudaStream_t) and never synchronise;
 *   - no function allocates device memory except tulip_net_create (persistent bf16 weight arena).
 * Each entry cites the reference code it replaces (paths relative to the reference root).
 */
#ifndef TULIP_B200_H
#define TULIP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TULIP_MAX_STAGES 8

typedef struct tulip_config {
  int img_h, img_w;           /* low-res input, tulip.py:531 img_size            */
  int tgt_h, tgt_w;           /* high-res target, target_img_size                */
  int patch_h, patch_w;       /* patch_size (patch_w must be 4: circular conv k=(ph,8), s=(ph,4), tulip.py:41) */
  int in_chans;               /* must be 1                                       */
  int embed_dim;              /* 96 in both factories, tulip.py:741,750          */
  int win_h, win_w;           /* window_size; win_h*win_w must be 16             */
  int num_layers;             /* len(depths)                                     */
  int depths[TULIP_MAX_STAGES];
  int num_heads[TULIP_MAX_STAGES];   /* head_dim = C/heads must be 32            */
  int mlp_ratio;              /* 4                                               */
  float ln_eps;               /* 1e-6, tulip.py:744                              */
  int log_transform;          /* second loss term in expm1 space, tulip.py:695-696 */
  int patch_expanding;        /* 0: PatchUnmerging (--patch_unmerging, every shipped script); 1: PatchExpanding = Linear(C, 2C,
                                 no bias) + '(P1 P2 C)' rearrange + LayerNorm(C/2) in the decoder and as first_patch_expanding
                                 (patch_unmerging=False, tulip.py:126-141, 472, 565) */
  int expanding_head;         /* 0: PixelShuffleHead (--pixel_shuffle); 1: FinalPatchExpanding = Linear(E, r^2 E, no bias) + '(P1 P2 C)'
                                 rearrange + LayerNorm(E) per output pixel (pixel_shuffle=False, tulip.py:144-159, 582, 727-729);
                                 embed_dim 96 only */
} tulip_config;

typedef struct tulip_net tulip_net;

const char* tulip_last_error(void);
int tulip_abi_version(void);

/* ---- whole network: TULIP.__init__ / TULIP.forward / autograd backward (tulip.py:531-584, 702-737) ---- */
int tulip_net_create(const tulip_config* cfg, tulip_net** out);
void tulip_net_destroy(tulip_net* net);
/* parameter schema in state_dict order without the int64 index buffers */
int tulip_net_num_params(const tulip_net* net);
int tulip_net_param_info(const tulip_net* net, int i, char* name, int name_cap, int64_t shape[4], int* ndim);
int tulip_net_num_blocks(const tulip_net* net);       /* Swin blocks, encoder then decoder, execution order */
int tulip_net_block_info(const tulip_net* net, int i, int* stage, int* shifted, int* H, int* W);
int64_t tulip_net_workspace_bytes(const tulip_net* net, int batch);
int64_t tulip_net_kernel_launches(const tulip_net* net);   /* kernels launched by this net so far */

/* built-in per-launch profiler: CUDA events around every kernel launch of forward/backward, aggregated per
 * kernel function.  enable=1 clears and starts recording, 0 stops.  read() synchronises on the recorded events and
 * returns, for one tag, the summed device time (ms), algorithmic FLOPs and bytes, and the launch count. */
int tulip_net_profile(tulip_net* net, int enable);
int tulip_net_profile_num_tags(void);
int tulip_net_profile_read(tulip_net* net, int tag, char* name, int name_cap, double* ms, double* flops, double* bytes,
                           int64_t* launches);
/* the same records one launch at a time, in launch order: returns the number of records; if i is in range also fills
 * tag, device time (ms), algorithmic FLOPs and bytes of launch i */
int tulip_net_profile_record(tulip_net* net, int i, int* tag, double* ms, double* flops, double* bytes);
/* where launch i of the recording belongs: stage (0 = widest), part (0 glue: merge / unmerge / skip / embed, 1 attention
 * half-block, 2 MLP half-block, 3 head + loss), backward (0 / 1).  Returns the number of records. */
int tulip_net_profile_where(tulip_net* net, int i, int* stage, int* part, int* backward);

/* forward: x_lo [B,1,h,w] fp32, target [B,1,H,W] fp32 or NULL (mc_drop=True, tulip.py:733-734)
 * params: flat fp32 buffer; param_offsets_host[i] = element offset of parameter i (schema order)
 * drop_scales: NULL (eval) or [2*num_blocks, B] fp32 per-sample DropPath scales, 0 or 1/keep (tulip.py:25-29)
 * win_mode_host: per block 0 = configured window, 1 = backup window (1,16)/shift (0,8) (tulip.py:284-287)
 * outputs: pred [B,1,H,W] fp32; losses[2] = {total L1, pixel L1} (tulip.py:690-700), untouched if target == NULL */
int tulip_net_forward(tulip_net* net, int batch, const float* params, const int64_t* param_offsets_host,
                      const float* x_lo, const float* target, const float* drop_scales, const int* win_mode_host,
                      void* workspace, float* pred, float* losses, void* stream);
/* Parameter version: any number that changes whenever the fp32 parameters change (e.g. the sum of the tensors' version
 * counters).  While it stays the same (and `params` is the same buffer) tulip_net_forward skips the fp32 -> bf16 weight re-pack
 * (216 MB of traffic per call for tulip_base).  A negative version (the default) means "unknown": re-pack on every forward. */
int tulip_net_set_params_version(tulip_net* net, long long version);
/* forward_only = 1: the following tulip_net_forward calls will not be followed by tulip_net_backward (evaluate(), MCdrop(),
 * torch.no_grad()): the fused half-block kernels run and no intermediate is saved.  0 (default) restores training mode. */
int tulip_net_set_inference(tulip_net* net, int forward_only);
/* backward of total_loss * grad_loss[0] (device scalar; GradScaler's 65536 arrives here, misc.py:294-295).
 * grads: flat fp32 buffer laid out like params; it is OVERWRITTEN (zeroed, then accumulated).
 * workspace must be the buffer the matching forward ran on, untouched since. */
int tulip_net_backward(tulip_net* net, int batch, const float* params, const int64_t* param_offsets_host,
                       float* grads, const float* x_lo, const float* target, const float* pred, const float* grad_loss,
                       const float* drop_scales, const int* win_mode_host, void* workspace, void* stream);
/* The same pass in up to three calls, so that a data-parallel host can start the all-reduce of finished gradient slices under
 * the rest of the pass (reference: DDP's bucketed reduce during backward, main_lidar_upsampling.py:277).  Phases, in order:
 *   0  head, decoder (layers_up.*), first_patch_expanding, skip_connection_layers, norm_up   (also zero-fills `grads`)
 *   1  top encoder stage (layers.{L-1}.*)
 *   2  remaining encoder stages (layers.{0..L-2}.*) and patch_embed
 * Call with [phase_lo, phase_hi] = [0,0], [1,1], [2,2] (or [0,2] = tulip_net_backward); when a call returns, the gradients of its
 * phases are complete on `stream`. */
int tulip_net_backward_phases(tulip_net* net, int batch, const float* params, const int64_t* param_offsets_host,
                              float* grads, const float* x_lo, const float* target, const float* pred, const float* grad_loss,
                              const float* drop_scales, const int* win_mode_host, void* workspace, void* stream,
                              int phase_lo, int phase_hi);

/* ---- dense contractions ----
 * NT: out[M,N] = A[M,K] . W[N,K]^T (+bias) -- every nn.Linear / 1x1 Conv2d on the path
 *     (tulip.py:298 qkv, :318 proj, :195/:198 fc1/fc2, :105 reduction, :119 expand, :716 skip, :175 conv_expand)
 * epilogue: 0 store, 1 bias+GELU (out2 = pre-activation), 2 residual: out = aux + row_scale*(acc+bias)
 * TN: dW[N,K] += dY[M,N]^T . X[M,K]; db[N] += colsum(dY)   (autograd of the above)
 * impl: 0 (2 is accepted as an alias).  There is ONE implementation, the tcgen05 / TMEM / TMA GEMM; N % 96 == 0, K % 8 == 0 and
 * 16-byte aligned operands are required and anything else is an error (round 1's warp-MMA backend, impl 1, was removed). */
int tulip_gemm_nt(const void* A, const void* W, const float* bias, void* out, void* out2, const void* aux,
                  const float* row_scale, int rows_per_sample, int M, int N, int K, int epilogue, int impl, void* stream);
/* host-side tiling decision of tulip_gemm_nt for (M, N, K, epilogue) on the tcgen05 path -- no device work:
 * out10 = {tile width BN, schedule (0 tile-major, 1 resident-weight panels, 2 CTA pairs = cta_group::2 on 256-row tiles), n_chunks, tiles per chunk, workers, A ring stages, K blocks,
 *          MMA steps in the last K block, grid size, operand ring stages} */
int tulip_gemm_nt_plan(int M, int N, int K, int epilogue, int save_pre, int* out10);
/* CTA-pair schedule of the NT GEMM (cta_group::2: two CTAs of a cluster run one M = 256 MMA, each loads its own A rows and
 * half of the B rows).  mode 0 = off (default: measured slower at this model's sizes), 1 = every eligible launch, 2 = launches
 * with K >= 384; anything else only queries.  Returns the previous mode (-1: not decided yet, env TULIP_B200_CG2 or 0). */
int tulip_gemm_nt_pairs_mode(int mode);
/* SM budget of the persistent kernels (GEMMs, fused half-block kernels, grouped weight gradients): n > 0 sizes their grids for
 * n SMs instead of the whole device, n <= 0 lifts the limit.  The data-parallel exchange uses it: while NCCL reduces the finished
 * gradient slices under the rest of the backward pass (the reference's DDP buckets, main_lidar_upsampling.py:277), the pass runs
 * on the SMs NCCL's CTAs leave free instead of queueing a second wave behind them.  Returns the previous budget (0 = none).
 * Launch sequences captured as CUDA graphs keep the budget they were captured under. */
int tulip_set_sm_budget(int n);
int tulip_gemm_tn(const void* dY, const void* X, float* dW, float* db, int M, int N, int K, int impl, void* stream);

/* The same two contractions with every operand mode and fused epilogue of the path spelled out (per-kernel parity tests of
 * PatchUnmerging, the skip Linear and the head bind these).  bf16 operands, fp32 parameters; 0 / NULL = unused.
 * epilogue: 0 store | 1 bias+GELU | 2 residual | 3 NHWC PixelShuffle(2) scatter (PatchUnmerging, tulip.py:117-123; W rows in
 *   the order n' = (2i+j)*Cc + c) | 4 split columns at split_col into out / out2 (skip Linear backward, tulip.py:715-716) |
 *   5 acc * gelu'(aux) | 6 head: pred[pixel] = sum_c wd[c] * leaky(acc + bias) (tulip.py:174-178, 731) |
 *   7 head backward: recomputes the pre-activation, out = dh, dwd += ... with dpred = sign(pred - target) * gscale / npix |
 *   8 row scale | 9 (A.B^T) * gelu'(A2.B2^T + bias)
 * a_mode 1: A[m=(b,h,w), k=(2i+j)*Cc+c] = src[b, 2h+i, 2w+j, c] (PixelShuffle(2) backward gather), geometry g_H, g_W, g_Cc */
typedef struct tulip_gemm_desc {
  const void* A; int64_t lda;
  const void* A2; int64_t lda2; int K1;        /* columns k >= K1 of the virtual A come from A2 (concat); K1 = K when unused */
  const void* B; int64_t ldb;
  const void* B2; int64_t ldb2; int K2;        /* epilogue 9 */
  int M, N, K;
  int a_mode, g_H, g_W, g_Cc;
  const float* bias;
  void* out; int64_t ldo;
  void* out2; int64_t ldo2;
  const void* aux; int64_t ldaux;
  const float* row_scale; int rows_per_sample;
  int split_col;
  const float* wd; const float* target; float* pred; const float* gscale; float* dwd;
  int hd_H, hd_W, hd_r, hd_E;
  /* epilogue 10 (LayerNorm backward on the output rows, N = 96 or 192): aux = LN input rows, aux2 = gradient added to dx or
   * NULL, ln_w = gamma [N], ln_stats = (mean, rstd) per row, ln_dw / ln_db += d(gamma) / d(beta); out2 (optional) =
   * row_scale[sample] * out */
  const void* aux2; int64_t ldaux2;
  const float* ln_w; const float* ln_stats; float* ln_dw; float* ln_db;
  /* epilogues 11 / 12 (= 0 / 2 on whole rows, N = 96 or 192, plus the LayerNorm that reads the output next): ln_w = gamma,
   * ln_b = beta, ln_y [M, N] bf16 = LayerNorm(out), ln_ystats [M, 2] = (mean, rstd) */
  const float* ln_b; void* ln_y; float* ln_ystats; float ln_eps;
  /* epilogues 6 / 7 in their FinalPatchExpanding form (tulip.py:144-159; hd_E = 96, no bias): pred[pixel] = sum_c wd[c] *
   * LayerNorm(acc row; ln_w, ln_b, ln_eps)[c]; 6 writes (mean, rstd) per pixel to ln_ystats [pixels, 2]; 7 reads them from
   * ln_stats, out = d(acc) (bf16), ln_dw / ln_db / dwd += d(gamma) / d(beta) / d(decoder_pred.weight) */
  int hd_ln;
} tulip_gemm_desc;
int tulip_gemm_nt_ex(const tulip_gemm_desc* d, int epilogue, void* stream);
/* dW[N,K] += dY^T . [X | X2]; rows written un-permuted when perm_R2 > 1 (row n' = ij*Cc + c -> c*R2 + ij); y_mode 1 gathers
 * dY through the PixelShuffle(2) backward view */
typedef struct tulip_gemm_tn_desc {
  const void* dY; int64_t ldy;
  const void* X; int64_t ldx;
  const void* X2; int64_t ldx2; int K1;
  int M, N, K;
  int y_mode, g_H, g_W, g_Cc;
  float* dW; int64_t lddw;
  float* db;
  int perm_R2, perm_Cc;
} tulip_gemm_tn_desc;
int tulip_gemm_tn_ex(const tulip_gemm_tn_desc* d, void* stream);
/* Up to 4 independent weight gradients (plain operands: no X2, y_mode 0) in ONE persistent launch: the way the executor
 * issues the weight-gradient GEMMs of a Swin half-block (fc2 + fc1, proj + qkv; autograd of tulip.py:194-200, 282-324).
 * CTAs walk (problem, dW tile, token range) work items with double-buffered TMEM accumulators.
 * tulip_gemm_tn_group_plan: host-side cut of the token ranges for `sms` SMs: per4[p] = 64-token blocks per work item of
 * problem p, *items = work items of the launch (tests, tooling; no GPU needed). */
int tulip_gemm_tn_group(const tulip_gemm_tn_desc* d, int n, void* stream);
int tulip_gemm_tn_group_plan(const int* M, const int* N, const int* K, int n, int sms, int* per4, int* items);

/* ---- fused head backward (PixelShuffleHead + decoder_pred + L1, tulip.py:161-178, 727-731, 692-693; embed_dim 96, r = 4):
 * ONE launch for   dh[m, ij*E + c] = dpred[m, ij] * wd[c] * LeakyReLU'(xn[m,:] . we[ij*E + c, :] + bias[ij*E + c])   (bf16 out),
 *                  dxn = dh . we  (bf16 [T, E]),   dwd[c] += sum_{m, ij} dpred[m, ij] * LeakyReLU(pre),
 * dpred = sign(pred - target) * gscale[0] / (T r^2).  we [E r^2, E] bf16 with rows in shuffle-slot order n' = ij*E + c, wet = its
 * transpose [E, E r^2]; pred / target fp32 [B, 1, H r, W r]; T = B H W. */
int tulip_head_bwd_fused_supported(int E, int r);
int tulip_head_bwd_fused(const void* xn, void* dxn, const void* we, const void* wet, const float* bias, const float* wd,
                         const float* pred, const float* target, const float* gscale, void* dh, float* dwd, int T, int E, int H,
                         int W, int r, void* stream);

/* ---- fused W-MSA / SW-MSA half-block (SURVEY 8b tulip_wmsa_block_fwd): ONE launch for
 *   y = x + row_scale[b] * proj(attn(qkv(LayerNorm(x))))          tulip.py:338-346 with WindowAttention.forward :282-324
 * LayerNorm, cyclic shift, window partition / reverse, rel-pos bias, shift mask, softmax, both Linears and the residual; the
 * two weight matrices stay resident in shared memory, q/k/v/S/P/O never reach HBM.  x, y: bf16 [B*H*W, C]; wqkv [3C, C] and
 * wproj [C, C] bf16 row-major (nn.Linear layout); everything else fp32.  row_scale: per-sample DropPath scale or NULL.
 * Built for C = 96 (3 heads of 32, stage 0 of both factories); other shapes return
 * an error (tulip_wmsa_block_supported says which) and the caller runs the unfused chain. */
int tulip_wmsa_block_supported(int B, int H, int W, int C, int heads, int Mh, int Mw);
int tulip_wmsa_block_fwd(const void* x, void* y, const float* ln_w, const float* ln_b, const void* wqkv, const float* bqkv,
                         const void* wproj, const float* bproj, const float* bias_table, const float* row_scale,
                         int B, int H, int W, int C, int heads, int Mh, int Mw, int sh, int sw, int masked,
                         int bias_Mh, int bias_Mw, float eps, void* stream);

/* ---- fused MLP half-block (SURVEY 8b tulip_mlp_block_fwd): ONE launch for
 *   y = x + row_scale[sample] * fc2(gelu(fc1(LayerNorm(x))))      tulip.py:347-352 with Mlp.forward :194-200 (exact-erf GELU)
 * x, y: bf16 [T, C]; w1 [4C, C], w2 [C, 4C] bf16 row-major (nn.Linear layout); everything else fp32.  The hidden tensor lives
 * in TMEM / shared memory only.  Training: pass xn [T, C] bf16, stats [T, 2] fp32 and hact [T, 4C] bf16 (all three or none) and
 * the same launch also stores the LayerNorm output, its (mean, rstd) and the activated hidden tensor for the backward pass.
 * Built for C = 96 (stage 0 of both factories); tulip_mlp_block_supported says whether a shape qualifies. */
int tulip_mlp_block_supported(int T, int C);
int tulip_mlp_block_fwd(const void* x, void* y, const float* ln_w, const float* ln_b, const void* w1, const float* b1,
                        const void* w2, const float* b2, const float* row_scale, int rows_per_sample,
                        void* xn, float* stats, void* hact, int T, int C, float eps, void* stream);

/* ---- window attention core: tulip.py:289-317 without the two Linears; shift/partition/mask/bias in-kernel ---- */
int tulip_window_attention_fwd(const void* qkv, const float* bias_table, void* out, int B, int H, int W, int C, int heads,
                               int Mh, int Mw, int sh, int sw, int masked, int bias_Mh, int bias_Mw, void* stream);
int tulip_window_attention_bwd(const void* qkv, const float* bias_table, const void* dout, void* dqkv, float* dbias_table,
                               int B, int H, int W, int C, int heads, int Mh, int Mw, int sh, int sw, int masked,
                               int bias_Mh, int bias_Mw, void* stream);

/* ---- LayerNorm (tulip.py:330,334,80,569), optional PatchMerging 2x2 gather on the input (tulip.py:92-99) ---- */
int tulip_layernorm_fwd(const void* x, const float* w, const float* b, void* y, float* stats, int rows, int C, float eps,
                        int merge_gather, int H2, int W2, void* stream);
int tulip_layernorm_bwd(const void* x, const float* w, const float* stats, const void* dy, const void* dres, void* dx,
                        float* dw, float* db, int rows, int C, int merge_gather, int H2, int W2, void* stream);

/* ---- PatchEmbedding: circular pad + Conv2d(1->E,(ph,8),(ph,4)) + LayerNorm (tulip.py:59-73) ---- */
int tulip_patch_embed_fwd(const float* x, const float* w, const float* b, const float* ln_w, const float* ln_b, void* y,
                          int B, int Himg, int Wimg, int ph, int E, float eps, void* stream);
int tulip_patch_embed_bwd(const float* x, const float* w, const float* b, const float* ln_w, const void* dy, float* dw,
                          float* db, float* dln_w, float* dln_b, int B, int Himg, int Wimg, int ph, int E, float eps,
                          void* stream);

/* ---- loss (tulip.py:690-700); scratch2 and out2 are 2 floats each ---- */
/* ---- evaluation post-processing: engine_upsampling.py:174-244 of evaluate(), fused into one pass ----
 * pred, target [B,1,H,W], x_lo [B,1,h_lo,W] fp32 as the model saw / produced them (log1p space when log_transform).
 * out [B,1,H,W]: expm1 (:177-180), clipped to [clip_lo, 1] else 0 (:183-188; 2/80 kitti & carla, 0.3/120 durlar), and with every
 * (H/h_lo)-th row overwritten by the sensor's own row when keep_low_res (:214-221, :236-244).
 * losses [B][2]: per frame {mean |out_before_overwrite - target| (:192-193), mean over the sensor rows |pred_row - x_lo| (:216-219)}.
 * scratch: 2*B floats. */
int tulip_eval_postprocess(const float* pred, const float* x_lo, const float* target, float* out, float* losses, float* scratch,
                           int B, int H, int W, int h_lo, int log_transform, float clip_lo, int keep_low_res, void* stream);
/* Monte-Carlo-dropout aggregation of MCdrop() (engine_upsampling.py:423-427): preds [n_passes, npix] fp32 (the stacked mc_drop
 * forward passes of ONE frame) -> out [npix] = mean over the passes, zeroed where std (unbiased) > threshold * mean; std_out
 * [npix] or NULL. */
int tulip_mc_dropout_aggregate(const float* preds, float* out, float* std_out, int n_passes, int64_t npix, float threshold,
                               void* stream);
/* ---- evaluation metrics on the device (reference tulip/util/evaluation.py, engine_upsampling.py:223-277) ----
 * range image -> points (evaluation.py:52-116, img_to_pcd_kitti / img_to_pcd_carla): img [B,H,W] normalised range,
 * points [B, H*W, 3]; x = (sin_h[w] cos_v[h]) r, y = (cos_h[w] cos_v[h]) r, z = sin_v[h] r, r = img * max_range.  The four
 * float32 tables hold the sines / cosines of the sensor's column and row angles (host side: tulip_b200.metrics.angle_tables). */
int tulip_range_to_points(const float* img, const float* sin_h, const float* cos_h, const float* sin_v, const float* cos_v,
                          float max_range, float* points, int B, int H, int W, void* stream);
/* voxel IoU / precision / recall / F1 of two clouds of n points (evaluation.py:148-175 + engine_upsampling.py:259-277): the voxel
 * index of a point is int((p - min) / grid_size) with min over both clouds; occupied voxels are kept in hash sets (no dense grid).
 * out4: doubles {iou, precision, recall, f1}. */
int64_t tulip_voxel_metrics_workspace_bytes(int n_points);
int tulip_voxel_metrics(const float* pts_pred, const float* pts_gt, int n_points, float grid_size, void* workspace, double* out4,
                        void* stream);
/* DurLAR (Ouster OS1-128) projection, evaluation.py:19-50 (img_to_pcd_durlar): float64 points [B, H*W, 3]; per-column tables
 * cos / sin(encoder + azimuth), cos / sin(encoder), per-row cos / sin(elevation) and the per-row column offset LUT come from the host
 * (tulip_b200.metrics); the point of pixel (row, col) is stored at row * W + (col + W - offset_lut[row]) % W. */
int tulip_range_to_points_durlar(const float* img, const double* cos_ea, const double* sin_ea, const double* cos_e, const double* sin_e,
                                 const double* cos_el, const double* sin_el, const int* offset_lut, float max_range, double origin_offset,
                                 double z_offset, double* points, int B, int H, int W, void* stream);
/* tulip_voxel_metrics for float64 clouds (the DurLAR projection is float64 in the reference) */
int tulip_voxel_metrics_f64(const double* pts_pred, const double* pts_gt, int n_points, double grid_size, void* workspace, double* out4,
                            void* stream);
/* Chamfer distance as evaluation.py:125-134 uses it: dist_a[i] = min_j |a_i - b_j|^2, dist_b likewise (the un-vendored
 * github.com/otaheri/chamfer_distance extension), out3 = {mean(dist_a) + mean(dist_b), mean(dist_a), mean(dist_b)}. */
int tulip_chamfer_distance(const float* a, const float* b, int na, int nb, float* dist_a, float* dist_b, float* out3, void* stream);
/* ---- input pipeline: the transform chains of tulip/util/datasets.py:244-369 in one pass ----
 * raw [B,H,W,channels] fp32 range frames in metres as the loaders return them (channel 0 = range, npy_loader :187-191);
 * v = raw * scale (ScaleTensor :140-144; 1/80 kitti & carla, 1/120 durlar); if filter: v outside [min_range, max_range] -> 0
 * (FilterInvalidPixels :146-154; durlar 0.3/120, carla 2/80; kitti has none); if log_transform: log1p (:73-75).
 * hi [B,1,H,W] gets every pixel, lo [B,1,H/row_factor,W/col_factor] rows 0, rf, 2rf, ... and columns 0, cf, ... of it
 * (DownsampleTensor :120-128, DownsampleTensorWidth :130-138). */
int tulip_preprocess_range(const float* raw, int channels, float scale, int filter, float min_range, float max_range, int row_factor,
                           int col_factor, int log_transform, float* hi, float* lo, int B, int H, int W, void* stream);
/* The CARLA `.rimg` container (rimg_loader, tulip/util/datasets.py:181-193): a file is two native unsigned longs (size[0], size[1])
 * followed by size[1] rows of size[0] float16 values; the loader returns flip(transpose(rows)) widened to float32.
 * rows_f16: the payloads of B files back to back (device, header stripped); frames [B, size0, size1] fp32:
 * frames[b][i][j] = rows[b][size1 - 1 - j][size0 - 1 - i] -- what tulip_preprocess_range takes as `raw` with channels = 1. */
int tulip_rimg_decode(const void* rows_f16, float* frames, int B, int size0, int size1, void* stream);
/* ---- optimizer step over the flat buffers: torch.optim.AdamW(param_groups_layer_decay(...), betas=(0.9, 0.95))
 * (main_lidar_upsampling.py:281-283) and get_grad_norm_ (util/misc.py:317-329) as one launch each ----
 * segments (device array, sorted by offset): parameter i occupies [offset, offset + numel) of the flat fp32 buffers and belongs to
 * optimizer group `group`; hyper (host struct, passed by value to the kernel): per-group lr and weight decay for this step (the
 * reference's scheduler rewrites lr every iteration, util/lr_sched.py), betas, eps, the two bias corrections of the step and a factor
 * applied to the gradients on the fly (1 / loss scale; 1 if the gradients were unscaled already).  Update = torch's
 * _single_tensor_adamw (decoupled weight decay, no amsgrad).  span = extent of the flat buffers in floats (multiple of 4). */
typedef struct { int64_t offset; int64_t numel; int group; int pad; } tulip_adamw_segment;
typedef struct {
  float lr[64]; float weight_decay[64];
  float beta1, beta2, eps, bias_correction1, bias_correction2_sqrt, grad_scale;
} tulip_adamw_hyper;
int tulip_adamw_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, const tulip_adamw_segment* segments_dev,
                     int n_segments, int64_t span, const tulip_adamw_hyper* hyper_host, void* stream);
/* global L2 norm of the gradients of all segments -> out[0] (device float); scratch: one double */
int tulip_grad_norm(const float* grads, const tulip_adamw_segment* segments_dev, int n_segments, int64_t span, double* scratch, float* out,
                    void* stream);
int tulip_l1_loss(const float* pred, const float* target, int64_t n, int log_transform, float* scratch2, float* out2,
                  void* stream);
/* inputs of a step (low-res frames, targets, DropPath scales; any pair may be NULL) copied into persistent buffers by ONE launch:
 * a host that keeps its buffers gets tulip_net_forward / _backward replayed as CUDA graphs (stable pointers) */
int tulip_stage_inputs(const float* lo, float* lo_dst, int64_t n_lo, const float* hi, float* hi_dst, int64_t n_hi, const float* drop,
                       float* drop_dst, int64_t n_drop, void* stream);

/* ---- stand-alone index ops (bit-exact; the same device functions the fused kernels use) ----
 * window_partition (tulip.py:248-252) composed with torch.roll(-sh,-sw) (tulip.py:290); reverse = tulip.py:320,323 */
int tulip_window_partition(const void* x, void* out, int B, int H, int W, int C, int Mh, int Mw, int sh, int sw, void* stream);
int tulip_window_reverse(const void* xw, void* out, int B, int H, int W, int C, int Mh, int Mw, int sh, int sw, void* stream);
int tulip_shift_mask(float* out, int H, int W, int Mh, int Mw, int sh, int sw, void* stream);          /* tulip.py:254-280 */
int tulip_rel_bias_gather(const float* table, float* out, int heads, int Mh, int Mw, void* stream);  /* tulip.py:304-307 */
int tulip_merge_gather(const void* x, void* out, int B, int H, int W, int C, void* stream);          /* tulip.py:92-99 */
int tulip_pixel_shuffle(const void* x, void* out, int B, int H, int W, int Cout, int r, void* stream); /* nn.PixelShuffle, NHWC */

#ifdef __cplusplus
}
#endif
#endif /* TULIP_B200_H */
