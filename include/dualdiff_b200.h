/* dualdiff_b200 — C ABI of libdualdiff_sm100.so
 *
 * The drop-in boundary for the DualDiff denoising-step hot path.  The reference is
 * pure Python and reaches its kernels through torch/diffusers/xformers; it has no FFI of its own, so
 * each entry point below names the reference call site (file:line under
 * /root/reference/MD_txt_con_fusion/magicdrive) whose arithmetic it replaces.  The Python modules in
 * dualdiff_b200/ (same class names / forward signatures / state-dict keys as the reference) bind these
 * with ctypes; INTEGRATION.md shows the stub.
 *
 * Conventions
 *   - every pointer is a raw CUDA device pointer (tensor.data_ptr()); PyTorch owns all memory;
 *   - activations are bf16, channels-last: "compact" = [rows = img*H*W, C]; "padded" = the zero-haloed
 *     pixel layout [img][H+1][W+1][C] consumed by the 3x3 implicit-GEMM convolution;
 *   - every call is stream-ordered on `stream` (a cudaStream_t passed as void*), performs no
 *     allocation and no host synchronisation, and is safe under CUDA-graph capture;
 *   - return 0 on success, negative on error; dd_last_error() returns a thread-local message.
 *     There is no CPU fallback: without an sm_100 device the calls fail.
 */
#ifndef DUALDIFF_B200_H_
#define DUALDIFF_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DD_VERSION 100
#if defined(__GNUC__)
#define DD_API __attribute__((visibility("default")))
#else
#define DD_API
#endif

DD_API int dd_version(void);
DD_API const char* dd_last_error(void);
/* number of kernels launched by this library since load (bench.py `gpu_launches`) */
DD_API long long dd_launch_count(void);

/* ---- GEMM / implicit-GEMM convolution on tcgen05 tensor cores ------------------------------------
 * out = epilogue( A[M, K(*taps)] x W[N, taps*K]^T )
 * replaces: torch.nn.Linear / Conv2d(1x1, 3x3 s1 p1) inside diffusers ResnetBlock2D, Transformer2DModel,
 * Attention.to_q/k/v/out, FeedForward (GEGLU), Upsample2D.conv, zero convs
 * (networks/unet_addon_rawbox.py:965,1031-1039; networks/unet_2d_condition_multiview.py:443,522;
 *  networks/blocks.py:72-83).
 */
typedef struct dd_gemm_args {
  const void* a;      /* bf16 [M, k1 or K] row-major, leading dim a_ld (elements)               */
  const void* a2;     /* optional second K-source (skip-connection concat), cols [k1, K)        */
  const void* w;      /* bf16 [N, taps*K] row-major (tap-major, channel-minor), leading dim w_ld */
  void* out;          /* bf16 or fp32 [rows, n_store], leading dim out_ld                        */
  const float* bias;  /* fp32 [N] or NULL                                                        */
  const float* rowvec;/* fp32 [n_img, rowvec_ld]: per-image vector added to every row, or NULL   */
  const void* res1;   /* bf16 residual [rows, >=n_store] or NULL                                 */
  const void* res2;   /* second bf16 residual or NULL                                            */
  int M, N, K;        /* K = channels per tap                                                    */
  int k1;             /* split point when a2 != NULL (multiple of 64)                            */
  int taps;           /* 1 = linear / 1x1 conv; 9 = 3x3 stride-1 pad-1 conv over the padded layout */
  int conv_h, conv_w; /* taps == 9: image height/width (padded layout has (H+1) x (W+1) rows)    */
  long long a_ld, a2_ld, w_ld, out_ld, res1_ld, res2_ld;
  int rowvec_ld, rows_per_img;
  int out_f32;        /* 1: fp32 output                                                          */
  int geglu;          /* 1: out[:, j] = v[:, j] * gelu(g[:, j]); W packed as 128-value / 128-gate
                         column groups per 256-wide tile (dd pack_geglu in dualdiff_b200/packing.py) */
  int force_bn;       /* 0 = auto tile width; testing hook                                        */
  int act;            /* 0 none, 1 SiLU applied after bias/residuals (ControlNetConditioningEmbedding) */
  int no_tma_epilogue;/* testing hook: 1 forces the register/smem-transpose epilogue                 */
  int one_cta;        /* testing hook: 1 forces the single-CTA kernel (default: CTA pairs, cta_group::2) */
  int stream_k;       /* 0 auto (used when the tiles fill the last wave badly and K is long), 1 force, -1 never */
  void* workspace;    /* optional fp32 scratch for stream-K partial tiles (caller-owned, one per stream) or NULL */
  long long workspace_bytes;
} dd_gemm_args;
DD_API int dd_gemm(const dd_gemm_args* args, void* stream);

/* ---- GroupNorm(32) [+SiLU] ------------------------------------------------------------------------
 * replaces diffusers ResnetBlock2D.norm1/norm2 + nonlinearity, Transformer2DModel.norm,
 * conv_norm_out + conv_act (networks/unet_2d_condition_multiview.py:519-521).
 * Input: compact channels-last rows [n_img*H*W, c1] (+ optional second source [.., c2], concatenated along
 * channels = the up-block skip concat).  Output: compact rows, or (padded_out=1) the zero-haloed layout
 * [n_img][(H+1)][(W+1)][C] that dd_gemm(taps=9) consumes.  stats: fp32 scratch of dd_groupnorm_scratch_floats()
 * elements (per-CTA partial sums, then per-channel scale/shift; no atomics -> bit-reproducible). */
typedef struct dd_groupnorm_args {
  const void* x1; const void* x2; void* out; float* stats;
  const float* gamma; const float* beta;
  long long x1_ld, x2_ld, out_ld;
  int n_img, h, w, c1, c2, groups;
  float eps;
  int silu, padded_out;
  int two_pass;   /* testing hook: 1 forces the two-kernel (statistics, apply) form; 0 = single-pass cluster kernel when it fits */
} dd_groupnorm_args;
DD_API int dd_groupnorm(const dd_groupnorm_args* args, void* stream);
DD_API long long dd_groupnorm_scratch_floats(int n_img, int c, int hw);

/* ---- LayerNorm over token rows (networks/blocks.py:163,177,192,225) -------------------------------- */
typedef struct dd_layernorm_args {
  const void* x; void* out; const float* gamma; const float* beta;
  long long x_ld, out_ld;
  int rows, c;
  float eps;
} dd_layernorm_args;
DD_API int dd_layernorm(const dd_layernorm_args* args, void* stream);

/* ---- fused flash attention on tcgen05 (self / text-cross / cross-view / SFA) -------------------------
 * replaces xformers.ops.memory_efficient_attention behind diffusers Attention (networks/blocks.py:166-222,
 * networks/txt_con_fusion.py:156-162).  out[img, row, h*head_dim + c] = sum_src softmax(q k_src^T * scale) v_src.
 * q/k/v are bf16 row-major matrices [n_img*l, ld]; head h of q lives in columns q_col0 + h*q_head_stride ...
 * (a fused QKV projection output is addressed in place).  head_dim 40 requires the Q and K heads to be
 * zero-padded to a 48-column stride (packing.py:pack_qkv).  n_src = 2 with kv_map[img*2 + s] = the two
 * neighbour images implements the reference's `neighboring_attn_type == "add"` cross-view attention. */
typedef struct dd_attention_args {
  const void* q; const void* k; const void* v; void* out;
  const int* kv_map;          /* device int32 [n_img * n_src] or NULL (kv image == query image) */
  long long q_ld, k_ld, v_ld, out_ld;
  int q_cols, k_cols, v_cols; /* logical widths of the q/k/v matrices (TMA bounds)             */
  int q_col0, k_col0, v_col0;
  int q_head_stride, k_head_stride, v_head_stride;
  int n_img, n_kv_img, heads, head_dim, lq, lk, n_src;   /* n_src: K/V sources per query image, 1..8 (kv_map columns) */
  float scale;
  int variant;                /* testing hook: 0 = auto.  head_dim 40: 0 = two query tiles per CTA, 25 % of the exponentials on
                                 the FMA pipe, 1 = one-tile kernel, 2 = two-tile kernel with every exponential on MUFU;
                                 head_dim 80 / 160: 1 = one CTA per work item instead of the persistent item loop */
  int v_ones;                 /* head_dim 40 only: the V heads are padded to a 48-column stride and column 40 of every head
                                 holds 1.0 (the projection's bias writes it).  O[:, 40] = sum_k P[:, k] is then the softmax
                                 denominator, accumulated by the tensor core: the kernel keeps no row sum of its own. */
  int concat;                 /* n_src > 1: 0 = one softmax per K/V source, outputs summed (neighboring_attn_type "add",
                                 networks/blocks.py:112-121); 1 = the sources are ONE key sequence under a single softmax
                                 ("concat", blocks.py:122-133; with all views of the scene as sources: "self", :134-137) */
} dd_attention_args;
DD_API int dd_attention(const dd_attention_args* args, void* stream);

/* ---- temporal attention over the frames of a clip (BASELINE.json configs[4], "DualDiff+ video") ------------
 * The reference repository has NO temporal block; dualdiff_b200 defines it as
 *   x_f += W_o MHA(q = LN(x_f), k = v = LN(x_f'), f' over all frames) + b_o     per (scene, view, token), bidirectional,
 * inserted after the cross-view attention of BasicMultiviewTransformerBlock (networks/blocks.py:190-222 is the sibling it
 * is modelled on).  Images are ordered [outer][frame][view]; q holds the frames_q local frames, k/v hold frames_kv frames
 * laid out as frames_kv / frames_per_rank blocks of [outer][frames_per_rank][view] images, kv_rank_stride images apart
 * (= the NCCL all-gather layout when frames are sharded over ranks; one block when they are not).
 * out[img, token, h*head_dim + c] = sum_f' softmax_f'(q k_f'^T * scale) v_f'. */
typedef struct dd_temporal_attention_args {
  const void* q; const void* k; const void* v; void* out;
  long long q_ld, k_ld, v_ld, out_ld;
  int q_col0, k_col0, v_col0;
  int q_head_stride, k_head_stride, v_head_stride;
  int n_outer, n_view, tokens, heads, head_dim;
  int frames_q, frames_kv, frames_per_rank;
  long long kv_rank_stride;
  float scale;
  /* Query / output frames in rank blocks as well (the token-sharded layout after FrameShard's all-to-all, where a rank holds
   * ALL frames of its share of the tokens): frames_q frames as blocks of [n_outer, frames_q_per_rank, n_view] images,
   * q_rank_stride images apart.  0 / 0 = one block (queries are this rank's own frames). */
  int frames_q_per_rank;
  long long q_rank_stride;
} dd_temporal_attention_args;
DD_API int dd_temporal_attention(const dd_temporal_attention_args* args, void* stream);

/* ---- Occupancy Ray-shape Sampling projector (networks/occ3d_proj.py:50-113, OccupancyRay.project) ---------------
 * origins / dirs: fp32 [n_pix, 3] ray origin and unit direction per (camera, y, x) pixel of the compressed image grid
 * (occ3d_proj.py:26-42, computed on the host exactly as the reference does); sem: uint8 Occ3D labels [D, H, W]
 * (200 x 200 x 16).  For sample s of pixel p the label of the voxel nearest to origin + s*sample_step*dir is looked up
 * (grid_sample 'nearest', align_corners=False semantics; class 17 outside the volume).
 * ids (optional): uint8 [n_pix, sample_point] = the reference's output.  rows (optional): bf16 [n_pix, sample_point] =
 * filter(ids) / 17, the channels-last form of the (B*6, 320, h, w) tensor the foreground branch consumes
 * (dataset/utils.py:412-420; keep_fg = 0 maps classes <= 10 to 17, keep_bg = 0 maps classes >= 11 to 17). */
DD_API int dd_ors_project(const float* origins, const float* dirs, const unsigned char* sem, unsigned char* ids, void* rows,
                          long long n_pix, int sample_point, float sample_step, int D, int H, int W, int keep_fg,
                          int keep_bg, void* stream);

/* ---- prompt encoder pieces -----------------------------------------------------------------
 * The pipeline encodes the prompt once per sample with transformers' CLIPTextModel through diffusers' `_encode_prompt`
 * (pipeline/pipeline_bev_controlnet.py:273-281).  Projections / MLP run on dd_gemm and the LayerNorms on dd_layernorm;
 * these are the remaining pieces (dualdiff_b200/networks/clip_text.py strings them together). */
/* CLIPTextEmbeddings: out[r, :] = tok_emb[ids[r], :] + pos_emb[r % seq_len, :]   (fp32 tables [vocab, c] / [seq_len, c],
 * int64 ids [n_tok], bf16 rows out).  An id outside [0, vocab) yields a NaN row (the host wrapper validates first). */
DD_API int dd_clip_embed(const long long* ids, const float* tok_emb, const float* pos_emb, void* out, long long out_ld,
                         long long n_tok, int seq_len, int c, int vocab, void* stream);
/* attention over short sequences (seq_len <= 128, head_dim 64), optionally causal -- CLIPAttention with the causal mask of
 * the text model: out[s*L + i, h*64 + c] = sum_{j <= i} softmax_j(q_i k_j * scale) v_j.  q/k/v are bf16 row-major
 * [n_seq*seq_len, ld] matrices addressed like dd_attention (column offset + head stride: a fused QKV output in place). */
typedef struct dd_seq_attention_args {
  const void* q; const void* k; const void* v; void* out;
  long long q_ld, k_ld, v_ld, out_ld;
  int q_col0, k_col0, v_col0;
  int q_head_stride, k_head_stride, v_head_stride;
  int n_seq, seq_len, heads, head_dim;
  int causal;
  float scale;
} dd_seq_attention_args;
DD_API int dd_seq_attention(const dd_seq_attention_args* args, void* stream);
/* out = x * sigmoid(1.702 x) over bf16 (CLIP `quick_gelu`), n elements (multiple of 8); in place allowed */
DD_API int dd_quick_gelu(const void* x, void* out, long long n, void* stream);

/* ---- layout / gather kernels feeding the implicit-GEMM convolution ---------------------------------- */
/* NCHW (fp32 or bf16, arbitrary outer strides) -> padded channels-last bf16 with channel zero-padding to cp.
 * Image index = outer * n_view + view; source offset = outer*stride_outer + view*stride_view + c*stride_c +
 * y*stride_h + x.  Used for conv_in on the latents (CFG duplication = stride_outer 0) and for the
 * 6-way panorama split of the bg condition image (networks/map_embedder.py:116-125). */
typedef struct dd_to_padded_args {
  const void* src; void* out;
  long long stride_outer, stride_view, stride_c, stride_h;
  int n_outer, n_view, c, h, w, cp;
  int src_f32;
} dd_to_padded_args;
DD_API int dd_nchw_to_padded(const dd_to_padded_args* args, void* stream);
/* EXPERIMENTAL (parity-checked, not on the default path yet): 3x3 stride-1 pad-1 patch rows of a few-channel NCHW
 * image, same addressing arguments as dd_nchw_to_padded: out[(img, y, x), tap*c + ch] = src[img, ch, y+kh-1, x+kw-1] (zero
 * outside), row length cp >= 9*c (multiple of 8, tail zero).  Turns conv_in on the 4-channel latents into one K = 40 GEMM. */
DD_API int dd_nchw_patches(const dd_to_padded_args* args, void* stream);
/* 3x3 stride-2 pad-1 patches of a compact activation -> [n_img*Ho*Wo, 9*C] (diffusers Downsample2D.conv and the
 * stride-2 convs of ControlNetConditioningEmbedding) */
DD_API int dd_im2col_s2(const void* x, long long x_ld, void* out, int n_img, int h, int w, int c, void* stream);
/* nearest-neighbour resize to (h2, w2) written in the padded layout (diffusers Upsample2D with explicit size,
 * networks/unet_2d_condition_multiview.py:363-374,500-501): src = floor(dst * in / out) */
DD_API int dd_upsample_pad(const void* x, long long x_ld, void* out, int n_img, int h, int w, int c, int h2, int w2,
                           void* stream);
/* compact -> padded copy (no normalisation), for 3x3 convs whose input is not a GroupNorm output */
DD_API int dd_pad_rows(const void* x, long long x_ld, void* out, int n_img, int h, int w, int c, void* stream);

/* ---- small fp32 pieces (time / camera / box embedders; hard part 4: fp32 sin/cos and MLPs) ----------- */
/* y[M,N] = act(x[M,K] W[N,K]^T + b); act: 0 none, 1 SiLU.  y fp32 (ld y_ld) and/or bf16 copy y16 (ld y16_ld). */
typedef struct dd_linear_f32_args {
  const float* x; const float* w; const float* b; float* y; void* y16;
  long long x_ld, y_ld, y16_ld;
  int M, N, K, act;
} dd_linear_f32_args;
DD_API int dd_linear_f32(const dd_linear_f32_args* args, void* stream);
/* diffusers Timesteps(dim, flip_sin_to_cos=True, freq_shift=0): out[i] = cat[cos(t_i f), sin(t_i f)] fp32 */
DD_API int dd_timestep_embedding(const float* t, float* out, int n, int dim, void* stream);
/* networks/embedder.py:5-40: [x, sin(x 2^k), cos(x 2^k)]_{k<nfreq} over the last dim (3) of x[rows,3] */
DD_API int dd_fourier_embed(const float* x, float* out, long long rows, int nfreq, void* stream);
/* networks/bbox_embedder.py:180-194: pos = fourier(boxes)*m + null_pos*(1-m); cls = class_tokens[c]*m + null_cls*(1-m) */
DD_API int dd_box_features(const float* boxes, const long long* classes, const unsigned char* masks,
                           const float* class_tokens, const float* null_pos, const float* null_cls,
                           float* pos_out, long long pos_ld, float* cls_out, long long cls_ld,
                           long long n_box, int n_pts, int cls_dim, int n_classes, void* stream);
/* out = silu(x) (fp32 -> bf16), n elements */
DD_API int dd_silu_to_bf16(const float* x, void* out, long long n, void* stream);
/* out = a + b (+ c) elementwise over bf16, n elements (multiple of 8) */
DD_API int dd_add_bf16(const void* a, const void* b, const void* c, void* out, long long n, void* stream);
/* out[r, c] = softmax_c(x[r, :] * scale) (fp32 scores -> bf16 probabilities), one warp per row: the single-head 512-wide
 * attention of the VAE decoder's mid block (diffusers AutoencoderKL, called at pipeline_bev_controlnet.py:101-113) */
DD_API int dd_softmax_rows(const float* x, long long x_ld, void* out, long long out_ld, int rows, int cols, float scale,
                           void* stream);
/* fp32 -> bf16 / bf16 -> fp32 NCHW<->channels-last conversions at the module boundary */
DD_API int dd_nchw_to_rows(const void* src, int src_f32, void* out, int n_img, int c, int hw, void* stream);
DD_API int dd_rows_to_nchw(const void* rows, int rows_f32, long long ld, void* out, int out_f32, int n_img, int c,
                           int hw, void* stream);

/* ---- CFG combine + UniPC/DDIM update, one pass (pipeline/pipeline_bev_controlnet.py:487-499) -------- */
/* eps: fp32 channels-last [2*n_img*hw, c] (uncond half first) or [n_img*hw, c] when cfg == 0
 * (eps_nchw = 1: same but NCHW, the layout UNet2DConditionModelMultiview.forward returns).
 * State (fp32, NCHW, n_img*c*hw each): x (latents, in/out), last (corrected sample), m0, m1 (x0 history).
 * coef (device, fp32[16]): {guidance, sigma_t, 1/alpha_t, a_last, a_m0, a_m1, a_x0, a_x, b_xc, b_x0, b_m0, ...}
 *   x0  = (x - sigma_t*eps) / alpha_t
 *   xc  = a_x*x + a_last*last + a_m0*m0 + a_m1*m1 + a_x0*x0          (UniC corrector; identity on step 0)
 *   x'  = b_xc*xc + b_x0*x0 + b_m0*m0                                 (UniP predictor / DDIM)
 *   then last <- xc, m1 <- m0, m0 <- x0, x <- x'.   Coefficients are computed on the host in fp64. */
DD_API int dd_cfg_sched_step(const float* eps, float* x, float* last, float* m0, float* m1, const float* coef,
                             int n_img, int c, int hw, int cfg, int eps_nchw, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DUALDIFF_B200_H_ */
