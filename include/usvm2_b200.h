/* usvm2_b200.h -- C ABI of libusvm2_b200.so: hand-written sm_100a kernels for MedSAM2's per-frame video
 * propagation path (reference: Joungjimin/US-Video-MedSAM2, config sam2.1_hiera_t512).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the parameter name ends in `_host`; sizes are element counts;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); calls only enqueue work;
 *   - return value: 0 on success, <0 on error (USVM_ERR_*); nothing is thrown, nothing is allocated;
 *   - activations are channels-last ("token major"): [rows, C] with C contiguous; bf16 is raw uint16 storage;
 *   - no torch types cross this boundary.
 *
 * Each entry point names the reference code it replaces (paths relative to the reference repo root).
 */
#ifndef USVM2_B200_H
#define USVM2_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define USVM_ABI_VERSION 1
int usvm_abi_version(void);
/* compute capability major*10+minor of the current device, or <0 */
int usvm_device_sm(void);
/* Number of SMs persistent kernels may occupy (0 = all of the device).  Host-side launch heuristic only: set before
 * launching / capturing work that runs on an SM partition (CUDA green context); grids are baked into captured graphs. */
int usvm_set_sm_budget(int sms);

/* ------------------------------------------------------------------------------------------------
 * (a14) connected components + hole filling
 * replaces sam2._C.get_connected_componnets  (sam2/csrc/connected_components.cu:213-282; Python binding
 * sam2/utils/misc.py:47-63) and fill_holes_in_mask_scores (sam2/utils/misc.py:312-338)
 * ---------------------------------------------------------------------------------------------- */
/* img uint8 [N,H,W] (nonzero = foreground, 8-connectivity); labels, counts int32 [N,H,W]; H, W even.
 * labels = 1 + smallest 2x2-block anchor index of the component, counts = component area; 0 on background. */
int usvm_cc2d_label_u8(const uint8_t* img, int32_t* labels, int32_t* counts, int N, int H, int W, void* stream);
/* scores fp32 [N,H,W]: background components (score <= 0) of area <= max_area are set to fill_value; may run in
 * place.  scratch_* (int32 [N,H,W]) are only needed when the 2x2-block grid exceeds shared memory (H*W > ~128K). */
int usvm_fill_holes_f32(const float* scores_in, float* scores_out, int32_t* scratch_labels, int32_t* scratch_counts,
                        int N, int H, int W, int max_area, float fill_value, void* stream);

/* Largest 26-connected component of a binary volume -- the CT driver's post-step
 * `labels = skimage.measure.label(seg); seg = labels == argmax(bincount(labels.flat)[1:]) + 1`
 * (medsam2_infer_3D_CT.py:76-79, 285).  vol, out uint8 [D,H,W] (nonzero = foreground; out is 0 / 1, all 0 for an empty
 * volume); scratch_parent, scratch_count int32 [D,H,W]; scratch_best one 64-bit word.  Ties in area go to the component
 * whose first voxel in raster order comes first (the lowest skimage label, np.argmax's first maximum). */
int usvm_cc3d_largest_u8(const uint8_t* vol, uint8_t* out, int32_t* scratch_parent, int32_t* scratch_count,
                         unsigned long long* scratch_best, int D, int H, int W, void* stream);

/* ------------------------------------------------------------------------------------------------
 * GEMM with fused epilogue:  C[M,N] = epi(A[M,K] . W[N,K]^T)
 * replaces nn.Linear / 1x1 Conv2d / ConvTranspose2d(k2,s2) / im2col convs dispatched to cuBLAS / cuDNN
 * (hieradet.py:60,77,139,164; image_encoder.py:114; memory_attention.py:36-38; sam/transformer.py:237-240;
 *  memory_encoder.py:55,94-96,152,157; mask_decoder.py:66-83)
 * ---------------------------------------------------------------------------------------------- */
#define USVM_ACT_NONE 0
#define USVM_ACT_RELU 1
#define USVM_ACT_GELU 2 /* exact erf form (nn.GELU default) */

typedef struct usvm_gemm_epilogue {
  const float* bias;      /* [N] or NULL */
  const float* col_scale; /* [N] or NULL: applied after the activation (CXBlock gamma, memory_encoder.py:113) */
  const float* residual;  /* fp32 [*, ldr] or NULL: added last; row index = row % res_mod when res_mod > 0 */
  int ldr;
  int res_mod;
  int act;
  float* out_f32; /* either or both outputs */
  int ldo_f32;
  void* out_bf16;
  int ldo_bf16;
  /* optional fused axial RoPE (position_encoding.py:194-221) on output columns [0, rope_cols): applied after the bias
   * to adjacent column pairs of rows whose index inside their batch of rope_rows_per_batch is < rope_n_rope, with
   * table row pos = (index % rope_table_rows) and table column tc = ((col % 256) / 2).  The tables hold
   * rope_table_rows x 128 fp32 values in a TILED order, element (pos, tc) at
   * ((pos / 32) * 32 + tc / 4) * 128 + (pos % 32) * 4 + tc % 4  (rope_table_rows % 32 == 0): the 32 accumulator rows an
   * epilogue warp owns then read contiguous memory.  tensor-core kernel only; rope_cos == NULL disables it */
  const float* rope_cos;
  const float* rope_sin;
  int rope_cols, rope_rows_per_batch, rope_n_rope, rope_table_rows;
  /* optional fused LayerNorm over the output row (N == 256 only, tensor-core bf16 kernel): out_f32 receives the
   * epilogue result x as usual and out_bf16 receives LayerNorm(x) * ln_w + ln_b (then exact GELU when ln_gelu != 0)
   * instead of a bf16 copy of x -- the "residual GEMM followed by the next block's norm" pair of every transformer
   * layer as one launch.  ln_w == NULL disables it. */
  const float* ln_w;
  const float* ln_b;
  float ln_eps;
  int ln_gelu;
  /* grouped residual rows (several videos batched into one launch: the objects of video g share that video's residual
   * block): when res_div > 0 the residual row is (row / res_div) * res_mod + row % res_mod */
  int res_div;
} usvm_gemm_epilogue;

/* bf16 operands, fp32 accumulation in TMEM: TMA (128B swizzle) -> tcgen05.mma -> tcgen05.ld epilogue.
 * A bf16 [M,K] row pitch lda, W bf16 [N,K] row pitch ldw (pitches % 8 == 0, bases 16-byte aligned).
 * block_n: 0 = auto; 32 / 64 / 128 / 256 = one 128 x block_n tile per CTA (latency-bound shapes); negative = the
 * persistent kernel (one CTA per SM walks the tiles, accumulator double-buffered in TMEM so the epilogue of tile i
 * overlaps the mainloop of tile i+1) with tile width -block_n (a multiple of 32, <= 256), -1 = its own choice.
 * Auto picks the persistent kernel when the problem has at least one full wave (148) of 128 x 128 tiles. */
int usvm_gemm_bf16_tc5(const void* A, int lda, const void* W, int ldw, const usvm_gemm_epilogue* ep_host, int M, int N,
                       int K, int block_n, void* stream);
/* Same kernel on fp32 operands as tf32 (10-bit mantissa products, fp32 accumulate): image-side projections of the SAM
 * mask decoder's two-way transformer (sam/transformer.py:257-286) and its upscaling convolutions, where bf16 operands
 * are too coarse.  A fp32 [M,K] pitch lda, W fp32 [N,K] pitch ldw (pitches % 4 == 0, 16-byte aligned); no fused RoPE.
 * block_n: 0 = auto, 32 / 64 / 128. */
int usvm_gemm_tf32_tc5(const float* A, int lda, const float* W, int ldw, const usvm_gemm_epilogue* ep_host, int M, int N,
                       int K, int block_n, void* stream);
/* Feed-forward block of a memory-attention layer as one kernel (memory_attention.py:92-98):
 *   out[M,256] = x + relu(h . W1^T + b1) . W2^T + b2,   h = LayerNorm(x) bf16 [M,256], x / out fp32 [M,256] (out may alias
 * neither), W1 bf16 [2048,256], W2 bf16 [256,2048].  M % 128 == 0, d_model == 256, hidden == 2048.  A cluster of 8 CTAs per
 * 128-row tile: the hidden activations stay in shared memory, the 8 partial outputs are reduce-scattered through
 * distributed shared memory in rank order (deterministic). */
int usvm_ffn_fused_tc5(const void* h_bf16, const float* x, const void* w1, const float* b1, const void* w2,
                       const float* b2, float* out, int M, int d_model, int hidden, void* stream);
/* developer aid: clock64() stamps of the persistent GEMM's first 64 tiles on CTA 0 (USVM2_PGEMM_DEBUG & 4), 64 x 16 words */
int usvm_debug_pgemm_profile(unsigned long long* host_out_1024);
/* fp32-accumulate SIMT GEMM, operands fp32 or bf16 (flags), any shape: fp32 decoder tail + checker. */
int usvm_gemm_simt(const void* A, int a_is_bf16, int lda, const void* W, int w_is_bf16, int ldw,
                   const usvm_gemm_epilogue* ep_host, int M, int N, int K, void* stream);

/* ------------------------------------------------------------------------------------------------
 * attention -- replaces F.scaled_dot_product_attention (hieradet.py:70; sam/transformer.py:270,281,344,355)
 * ---------------------------------------------------------------------------------------------- */
typedef struct usvm_fmha_params {
  const void* q; /* bf16; element (b, row, h, c) at q + b*q_bs + row*q_rs + h*q_hs + c */
  const void* k;
  const void* v;
  void* o; /* bf16, same addressing with o_* */
  long long q_bs, k_bs, v_bs, o_bs;
  int q_rs, k_rs, v_rs, o_rs;
  int q_hs, k_hs, v_hs, o_hs;
  int B, H, Nq, Nk, head_dim; /* head_dim 96 or 256 */
  int num_splits;             /* >1: keys are split across CTAs (flash-decoding) and merged by a second kernel */
  float* o_part;              /* fp32 [num_splits, B*H, Nq, head_dim] when num_splits > 1 */
  float* ml_part;             /* fp32 [num_splits, B*H, Nq, 2] */
  float scale;                /* 1/sqrt(head_dim) */
  int part_bf16;              /* 1: o_part holds bf16 instead of fp32 (usvm_fmha_tc5 TS variant + usvm_fmha_combine only):
                                 halves the L2 traffic of the split partials; each partial is rounded once, like the
                                 bf16 output itself */
} usvm_fmha_params;
int usvm_fmha_bf16(const usvm_fmha_params* p_host, void* stream);
/* Fused Hiera window attention for windows of <= 256 keys (hieradet.py:46-95 + backbones/utils.py:17-58): window
 * partition, QKV split, optional 2x2 query max-pool, attention per (window, head of 96) and window un-partition in one
 * kernel.  qkv bf16 [F, Hg, Wg, 3*C] raster order (C = heads*96), qkv_bias fp32 [3*C] (value of the tokens in the zero
 * padding of partial windows), out bf16 [F, Ho, Wo, C] raster order (Ho = Hg/2 when pool). */
int usvm_window_attn_bf16(const void* qkv, const float* qkv_bias, void* out, int F, int Hg, int Wg, int ws, int pool,
                          int C, int heads, void* stream);
/* tcgen05 / TMEM / TMA flash attention for the memory-attention shapes: head_dim 256, H == 1, Nq % 128 == 0, batches
 * contiguous (x_bs == N * x_rs).  Q, S = QK^T (double buffered), P and O live in TMEM, V is consumed in place as an
 * MN-major operand.  With num_splits > 1 it writes partials only: follow with usvm_fmha_combine. */
int usvm_fmha_tc5(const usvm_fmha_params* p_host, void* stream);
/* Image-encoder attention on tcgen05 / TMEM / TMA straight from the raster-order qkv tensor (replaces window_partition +
 * F.scaled_dot_product_attention + window_unpartition of MultiScaleAttention / the ViT blocks, hieradet.py:56-81,
 * backbones/utils.py:17-61, efficient_track_anything/modeling/backbones/vitdet.py): qkv bf16 [F, H, W, 3*dim] with
 * (q | k | v) x (head, channel) columns, out bf16 [F, H, W, dim]; head_dim = dim / heads in {64, 96}.
 * window == 0: global attention, H*W % 128 == 0.  window == 14 or 7: non-overlapping windows over the grid padded to a
 * multiple of the window; the padding tokens' q / k / v are the projection bias (qkv_bias fp32 [3*dim]), as in the reference
 * where the zero padding follows the LayerNorm.  pool != 0 (window 14, even H and W): the queries are the 2 x 2 max-pool
 * of each window's q (do_pool, hieradet.py:25-36, 60-67) and out is bf16 [F, H/2, W/2, dim]. */
typedef struct usvm_hiera_attn_params {
  const void* qkv;
  void* out;
  const float* qkv_bias;
  int F, H, W, dim, heads, window;
  float scale; /* 1/sqrt(head_dim) */
  int pool;
} usvm_hiera_attn_params;
int usvm_hiera_attn_tc5(const usvm_hiera_attn_params* p_host, void* stream);
int usvm_fmha_combine(const usvm_fmha_params* p_host, void* stream);
/* fp32, one warp per query; head_dim 16 or 32, Nk <= 1024 (SAM decoder two-way transformer) */
int usvm_attn_small_f32(const float* q, const float* k, const float* v, float* out, int B, int H, int Nq, int Nk,
                        int head_dim, int q_rs, int k_rs, int v_rs, int o_rs, float scale, void* stream);

/* ------------------------------------------------------------------------------------------------
 * normalisation / position encoding / layout kernels
 * ---------------------------------------------------------------------------------------------- */
/* nn.LayerNorm over C (hieradet.py:101,124; memory_attention.py:42-44; transformer.py:162-177) */
int usvm_layernorm(const float* x, int ldx, const float* w, const float* b, float eps, int gelu, float* out_f32,
                   int ldo_f32, void* out_bf16, int ldo_bf16, int rows, int C, void* stream);
/* out = alpha*x[row % x_mod] + beta*y[row % y_mod] (mod 0: same rows); pos-enc adds (memory_attention.py:134-135).
 * x_div > 0: x row = (row / x_div) * x_mod + row % x_mod -- the objects of one video share that video's feature rows */
int usvm_axpby_rows(const float* x, const float* y, float alpha, float beta, int x_mod, int y_mod, int x_div,
                    float* out_f32, void* out_bf16, long long rows, int C, void* stream);
int usvm_cast_f32_bf16(const float* x, void* y, long long n, void* stream);
/* apply_rotary_enc (position_encoding.py:194-221): rows inside each batch of rows_per_batch with index < n_rope are
 * rotated with table row (index % table_rows); cos/sin tables [table_rows, dim/2]; output bf16 */
int usvm_rope_bf16(const float* x, int ldx, const float* cos_t, const float* sin_t, void* out, int ldo, long long rows,
                   int rows_per_batch, int n_rope, int table_rows, int dim, void* stream);
/* window_partition + qkv split + q max-pool (backbones/utils.py:17-37, hieradet.py:56-67); qkv bf16 [F,Hg,Wg,3C] */
int usvm_window_gather(const void* qkv, const float* qkv_bias, void* Qw, void* Kw, void* Vw, int F, int Hg, int Wg,
                       int ws, int pool, int C, void* stream);
/* window_unpartition (backbones/utils.py:40-61): [F*nW, wq*wq, C] -> [F, Ho, Wo, C] */
int usvm_window_scatter(const void* Ow, void* out, int F, int Ho, int Wo, int wq, int C, void* stream);
/* MaxPool2d(2,2) on NHWC fp32 (hieradet.py:25-36,139-140) */
int usvm_maxpool2_nhwc(const float* x, float* y, int F, int H, int W, int C, void* stream);
/* FPN nearest-x2 top-down add (image_encoder.py:116-126): fine += up2(coarse); optional bf16 copy */
int usvm_upsample2_add(float* fine, const float* coarse, void* fine_bf16, int F, int H, int W, int C, void* stream);
/* PatchEmbed 7x7/s4/p3 as im2col (backbones/utils.py:64-94): img fp32 [F,3,S,S] -> bf16 [F*(S/4)^2, KP>=147] */
int usvm_im2col_patch(const float* img, void* A, int F, int S, int KP, void* stream);
/* ViT PatchEmbed with kernel = stride = P, no padding (efficient_track_anything/modeling/backbones/vitdet.py:214-220,
 * utils.py:64-94): img fp32 [F,3,S,S] -> bf16 [F*(S/P)^2, 3*P*P], column (c, ky, kx); P % 8 == 0 */
int usvm_im2col_patch_grid(const float* img, void* A, int F, int S, int P, void* stream);
/* frame ingest (sam2/utils/misc.py:253-276): uint8 gray [F,H,W] -> fp32 [F,3,H,W] (x/255 - mean)/std */
int usvm_normalize_gray_u8(const uint8_t* gray, float* out, int F, int H, int W, const float* mean3_host,
                           const float* std3_host, void* stream);

/* JPEG-folder ingest (sam2/utils/misc.py:92-101, 213-277): uint8 RGB [F,H,W,3] as decoded and resized on the host ->
 * fp32 [F,3,H,W], (x/255 - mean_c)/std_c -- the reference's arithmetic, on a quarter of the host-to-device bytes */
int usvm_normalize_rgb_u8(const uint8_t* rgb_hwc, float* out, int F, int H, int W, const float* mean3_host,
                          const float* std3_host, void* stream);

/* memory-bank assembly (sam2_base.py:1344-1437) */
#define USVM_MAX_MEMORY_FRAMES 32
typedef struct usvm_memory_frames {
  const void* mem[USVM_MAX_MEMORY_FRAMES]; /* bf16 [B, T, Cm] token-major spatial memories */
  int tpos_index[USVM_MAX_MEMORY_FRAMES];  /* row of maskmem_tpos_enc to add (num_maskmem - t_pos - 1) */
  int count;
} usvm_memory_frames;
/* k_in = mem + pos + tpos, v_in = mem for rows [row_offset, row_offset + count*T + n_ptr_tokens) of [B, Nk_total, Cm];
 * pointer tokens fp32 [B, n_ptr_tokens, Cm] with their own pos [n_ptr_tokens, Cm] */
int usvm_build_memory(const usvm_memory_frames* frames_host, const float* pos, const float* tpos, const float* ptrs,
                      const float* ptr_pos, void* k_in, void* v_in, int B, int T, int Cm, int n_ptr_tokens,
                      int Nk_total, int row_offset, void* stream);
/* Per-frame control block kept in DEVICE memory: everything that varies from frame to frame (and from session to
 * session) on the steady-state tracking path -- where the frame store lives, which stored frames feed the memory bank
 * (sam2_base.py:1296-1394) and which slot this frame writes.  Kernels read it through a pointer, so ONE captured CUDA
 * graph serves every frame of every session with the same (objects, #memories, #pointers) signature. */
#define USVM_MAX_PTRS 48
typedef struct usvm_frame_ctrl {
  void* mem_store;   /* bf16 [slots][B][T][Cm]   spatial memories            */
  float* ptr_store;  /* fp32 [slots][B][4*Cm]    object pointers             */
  float* score_store; /* fp32 [slots][B]         object score logits         */
  float* mask_store; /* fp32 [slots][B][h*w]     low-res mask logits (hole-filled) */
  long long mem_slot_stride, ptr_slot_stride, score_slot_stride, mask_slot_stride; /* elements per slot */
  int cur_frame; /* slot written by this frame */
  int n_mem, n_ptr, reserved;
  int mem_frame[USVM_MAX_MEMORY_FRAMES]; /* slot of each spatial memory, in concatenation order */
  int mem_tpos[USVM_MAX_MEMORY_FRAMES];  /* row of maskmem_tpos_enc for each */
  int ptr_frame[USVM_MAX_PTRS];          /* slot of each object pointer */
  float ptr_rel[USVM_MAX_PTRS];          /* signed temporal distance / (max_obj_ptrs - 1) */
} usvm_frame_ctrl;
/* writes *ctrl_host into device memory through a by-value kernel parameter (asynchronous, no staging copy) */
int usvm_set_frame_ctrl(usvm_frame_ctrl* ctrl_dev, const usvm_frame_ctrl* ctrl_host, void* stream);
/* the same plus up to 4 device-to-device copies (src[i] -> dst[i], bytes[i] a multiple of 16, 16-byte aligned) in the
 * same launch: the frame's backbone features (sam2_video_predictor.py:879-910 `_get_image_feature`) into the static
 * inputs of the captured tracked-frame graph */
int usvm_frame_prologue(usvm_frame_ctrl* ctrl_dev, const usvm_frame_ctrl* ctrl_host, const void* const* src,
                        void* const* dst, const long long* bytes, int n_seg, void* stream);
/* memory feature epilogue (sam2_base.py:1488-1496; sam2_video_predictor.py:956): + no_obj_embed_spatial where
 * score <= 0 (score stride in elements), rounded to bf16 into mem_bf16 [B, T, Cm], or -- when mem_bf16 is NULL --
 * into slot ctrl_dev->cur_frame of ctrl_dev->mem_store */
int usvm_finalize_memory(const float* x, const float* score, int score_stride, const float* no_obj_embed,
                         void* mem_bf16, int B, int T, int Cm, const usvm_frame_ctrl* ctrl_dev, void* stream);
/* get_1d_sine_pe + obj_ptr_tpos_proj (sam2_utils.py:64-74, sam2_base.py:1402-1408) from ctrl->ptr_rel:
 * out fp32 [n_ptr*4, 64] */
int usvm_ptr_tpos(const usvm_frame_ctrl* ctrl_dev, const float* W, const float* bias, float* out, int n_ptr,
                  void* stream);
/* usvm_build_memory reading the frame store through the control block;
 * k_in / v_in bf16 [B, n_mem*T + n_ptr*4, Cm] */
int usvm_build_memory_store(const usvm_frame_ctrl* ctrl_dev, const float* pos, const float* tpos, const float* ptr_pos,
                            void* k_in, void* v_in, int B, int T, int Cm, int n_mem, int n_ptr, void* stream);
/* slot ctrl->cur_frame of the pointer / score / mask stores <- obj_ptr [B,256], score (stride score_stride), masks [B,hw] */
int usvm_store_outputs(const usvm_frame_ctrl* ctrl_dev, const float* obj_ptr, const float* score, int score_stride,
                       const float* masks, int B, int ptr_dim, int hw, void* stream);

/* ------------------------------------------------------------------------------------------------
 * small convolutions / resize (memory encoder, prompt encoder)
 * ---------------------------------------------------------------------------------------------- */
/* NHWC fp32 direct conv, weights [k][k][Cin][Cout] (<= 48 KB), Cout <= 64, optional LayerNorm2d + GELU
 * (memory_encoder.py:40-56; prompt_encoder.py:57-66; sam2_base.py:863) */
int usvm_conv2d_small(const float* x, const float* w_kkio, const float* bias, const float* ln_w, const float* ln_b,
                      float eps, int gelu, float* out_f32, void* out_bf16, int B, int H, int W, int Cin, int Cout,
                      int k, int stride, int pad, void* stream);
/* First stage of the MaskDownSampler (Conv2d 1 -> 4, k x k / stride, LayerNorm2d, GELU) reading a VIRTUAL H x W input: the
 * bilinear (align_corners=False) upsampling of low [B, hi, wi] followed by post_mode (sigmoid * scale + bias or binarise,
 * sam2_base.py:1472-1484) is evaluated while the footprint is staged, so the upsampled mask is never written.
 * out_f32 [B, Ho, Wo, 4]; Ho, Wo multiples of 16. */
int usvm_conv2d_mask_first(const float* low, int hi, int wi, int post_mode, float post_scale, float post_bias,
                           const float* w_kkio, const float* bias, const float* ln_w, const float* ln_b, float eps, int gelu,
                           float* out_f32, int B, int H, int W, int k, int stride, int pad, void* stream);
int usvm_im2col_nhwc(const float* x, void* A, int B, int H, int W, int C, int k, int stride, int pad, void* stream);
/* CXBlock depthwise 7x7 + LayerNorm2d (memory_encoder.py:104-108); weights [49][C]; C == 256; bf16 out */
int usvm_dwconv7_ln(const float* x, const float* w_49c, const float* bias, const float* ln_w, const float* ln_b,
                    float eps, void* out_bf16, int B, int H, int W, int C, void* stream);
#define USVM_POST_NONE 0
#define USVM_POST_SIGMOID_AFFINE 1  /* sigmoid(v)*scale + bias  (sam2_base.py:1477-1484) */
#define USVM_POST_BINARIZE_AFFINE 2 /* (v > 0)*scale + bias     (sam2_base.py:1472-1474) */
int usvm_resize_bilinear(const float* x, float* y, long long planes, int Hi, int Wi, int Ho, int Wo, int post_mode,
                         float post_scale, float post_bias, void* stream);
/* SAM2Base._apply_non_overlapping_constraints (sam2/modeling/sam2_base.py:1663-1681) on x fp32 [B, HW]: inside each
 * group of `group` consecutive objects (the objects of one video; group <= 0: all B) only the object with the highest
 * logit at a pixel keeps it, the others are clamped to <= -10; then the post transform of usvm_resize_bilinear.  Ties
 * go to the lowest object index (torch.argmax).  May run in place. */
int usvm_non_overlap_f32(const float* x, float* y, int B, long long HW, int group, int post_mode, float post_scale,
                         float post_bias, void* stream);
int usvm_resize_bilinear_aa(const float* x, float* y, long long planes, int Hi, int Wi, int Ho, int Wo,
                            int binarize_half, void* stream);

/* ------------------------------------------------------------------------------------------------
 * SAM mask decoder tail (fp32)
 * ---------------------------------------------------------------------------------------------- */
/* feat_group > 0: feat_s1 / feat_s0 hold one frame per group of feat_group consecutive objects (the objects of a video
 * share its frame; feat_group >= B: one frame for the whole batch); 0: one frame per object */
int usvm_upscale1_ln_gelu(const float* g1, const float* feat_s1, const float* ln_w, const float* ln_b, float eps,
                          float* out, int B, int Hc, int Wc, int C, int feat_group, void* stream);
int usvm_upscale2_masks(const float* g2, const float* feat_s0, const float* hyper, int hyper_bs, float* masks, int B,
                        int Hc, int Wc, int feat_group, void* stream);
int usvm_small_mlp3(const float* x, long long x_row_stride, long long x_inst_stride, const int* row_select,
                    const float* w1, const float* b1, const float* w2, const float* b2, const float* w3,
                    const float* b3, int out_dim, int sigmoid_out, float* y, long long y_row_stride,
                    long long y_inst_stride, int rows, int instances, void* stream);
/* Latency-oriented fp32 GEMM for a handful of rows (the 8 decoder tokens per object, the stacked 3-layer heads):
 * out[i][m][n] = act((x + x2)[i][m] . w[i][n] + bias[i][n]) + residual[i][m][n];  strides in elements;
 * row_select (int32 [M], optional) adds row_select[m] * x_sel_stride to the x row address (token picked on device). */
typedef struct usvm_skinny_params {
  const float* x;
  long long x_is, x_rs;
  const float* x2; /* optional second addend, NULL = none */
  long long x2_is, x2_rs;
  const int* row_select;
  long long x_sel_stride;
  const float* w; /* [instances][N][K] */
  long long w_is;
  const float* bias; /* [instances][N] or NULL */
  long long b_is;
  const float* residual;
  long long r_is, r_rs;
  float* out;
  long long o_is, o_rs;
  int M, N, K, instances, act;
  int x2_cols; /* x2 is added for output columns < x2_cols only (a multiple of 4); 0 or >= N: for all columns */
  /* optional LayerNorm of the input rows on load (K == 256 only; transformer.py:192-211 norm1..3, :131 norm_final_attn):
   * x <- LN(x) * ln_w + ln_b before x2 is added; the normalised rows are also written to ln_out (row (m) of instance i at
   * ln_out + i * ln_is + m * ln_rs) by the first CTA of each (instance, row block), bit-identical to usvm_layernorm */
  const float* ln_w;
  const float* ln_b;
  float* ln_out;
  long long ln_is, ln_rs;
  float ln_eps;
} usvm_skinny_params;
int usvm_gemm_skinny_f32(const usvm_skinny_params* p_host, void* stream);
/* token -> image attention (transformer.py:194-198): q [B*Nt, H*16], k/v rows of stride kv_rs, Nt <= 16 */
int usvm_attn_t2i_f32(const float* q, int q_rs, const float* k, const float* v, int kv_rs, float* out, int o_rs, int B,
                      int H, int Nt, int Nk, float scale, void* stream);
/* Same attention split over the keys (128 per CTA, <= 64 splits): B*H*ceil(Nk/128) CTAs write partials
 * [B*H][splits][Nt][18] floats and the last CTA of each head (ticket in counters[B*H], zero before the first launch and
 * left zero afterwards) merges them in split order. */
int usvm_attn_t2i_split_f32(const float* q, int q_rs, const float* k, const float* v, int kv_rs, float* out, int o_rs,
                            int B, int H, int Nt, int Nk, float scale, float* partials, int* counters, void* stream);
/* image -> token attention (transformer.py:205-210): q [B*Nq, .] row stride q_rs, k/v [B*Nt, .] row stride kv_rs */
int usvm_attn_i2t_f32(const float* q, int q_rs, const float* k, const float* v, int kv_rs, float* out, int o_rs, int B,
                      int H, int Nq, int Nt, float scale, void* stream);
/* multimask argmax-IoU / stability fallback + NO_OBJ_SCORE gating (mask_decoder.py:146-153,247-295;
 * sam2_base.py:1112-1141); iou: 4 values per object (logits when iou_is_logit), score: 1 value per object */
int usvm_sam_select(const float* masks, const float* iou, int iou_stride, int iou_is_logit, const float* score,
                    int score_stride, int multimask, float stab_delta, float stab_thresh, float no_obj_score,
                    float* low_res, int* token_index, float* iou_out, int B, int HW, void* stream);
int usvm_objptr_mix(float* ptr, const float* score, int score_stride, const float* no_obj_ptr, int B, int C,
                    void* stream);
int usvm_point_embed(const float* coords, const int* labels, const float* gauss, const float* table, float image_size,
                     float* out, int n_points, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* USVM2_B200_H */
