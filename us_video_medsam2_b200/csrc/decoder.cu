// fp32 tail of the SAM mask decoder (mask_decoder.py:168-295, sam2_base.py:1010-1166).
//
//   usvm_upscale1_ln_gelu   pixel-shuffle of the first ConvTranspose2d(k2,s2) GEMM + feat_s1 add +
//                           LayerNorm2d + GELU                        (mask_decoder.py:217-221)
//   usvm_upscale2_masks     pixel-shuffle of the second ConvTranspose GEMM + feat_s0 add + GELU fused with the
//                           hyper-network mask product [4 x 32] . [32]  (:222-231) -- the 32-channel
//                           128x128 embedding never reaches HBM
//   usvm_small_mlp3         3-layer MLPs on a handful of rows: 4 hyper-networks, IoU head, object-score
//                           head, obj_ptr_proj                          (:224-238, sam2_base.py:1146)
//   usvm_sam_select         multimask argmax-IoU / stability fallback, object-score gating to -1024,
//                           obj-pointer no-object mixing                (mask_decoder.py:146-153,247-295;
//                                                                        sam2_base.py:1112-1156)
//   usvm_point_embed        random-Fourier point prompt embedding       (prompt_encoder.py:79-103)
#include "common.cuh"
#include "usvm2_b200.h"

namespace {

// g1: [B*Hc*Wc, 4*C] fp32, column (dy*2+dx)*C + c (bias already added); feat: [B, 2Hc, 2Wc, C]
template <int C>
__global__ void __launch_bounds__(256)
upscale1_kernel(const float* __restrict__ g1, const float* __restrict__ feat, const float* __restrict__ ln_w,
                const float* __restrict__ ln_b, float eps, float* __restrict__ out, int B, int Hc, int Wc,
                int feat_group) {
  constexpr int CPL = C / 32;
  const int lane = threadIdx.x & 31;
  float gw[CPL], gb[CPL];  // LayerNorm parameters: constants, fetched before the programmatic-dependency wait
#pragma unroll
  for (int j = 0; j < CPL; ++j) {
    gw[j] = __ldg(ln_w + j * 32 + lane);
    gb[j] = __ldg(ln_b + j * 32 + lane);
  }
  PDL_ENTRY();
  const int Ho = 2 * Hc, Wo = 2 * Wc;
  const long long pix = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (pix >= (long long)B * Ho * Wo) return;
  const int X = (int)(pix % Wo), Y = (int)((pix / Wo) % Ho), b = (int)(pix / ((long long)Wo * Ho));
  const long long src = ((long long)b * Hc + (Y >> 1)) * Wc + (X >> 1);
  const int q = (Y & 1) * 2 + (X & 1);
  // a video's objects share its frame of features (feat_group consecutive objects per frame; 0: one frame each)
  const long long fpix = feat_group > 0 ? ((long long)(b / feat_group) * Ho + Y) * Wo + X : pix;
  float v[CPL];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < CPL; ++j) {
    const int c = j * 32 + lane;
    v[j] = g1[src * 4 * C + q * C + c] + feat[fpix * C + c];
    s += v[j];
  }
  const float mean = warp_sum(s) / C;
  float qq = 0.f;
#pragma unroll
  for (int j = 0; j < CPL; ++j) {
    v[j] -= mean;
    qq = fmaf(v[j], v[j], qq);
  }
  const float rstd = 1.0f / sqrtf(warp_sum(qq) / C + eps);
#pragma unroll
  for (int j = 0; j < CPL; ++j) {
    const int c = j * 32 + lane;
    out[pix * C + c] = gelu_erf(v[j] * rstd * gw[j] + gb[j]);
  }
}

// g2: [B*Hc*Wc, 4*32]; feat_s0: [B, 2Hc, 2Wc, 32]; hyper: [B, 4, 32]; masks: [B, 4, 2Hc, 2Wc]
// One thread per output pixel: its 32 up-scaled channels are two 128-byte reads (GEMM row segment + feature row), the four
// hyper-network vectors of the block's object sit in shared memory (broadcast reads), the four mask planes are written
// coalesced along x.  (A first version gave a warp to each pixel and reduced over the lanes: 20 shuffles per pixel, 171 us
// for 32 objects against ~75 MB of traffic.)
constexpr int UP2_THREADS = 128;
__global__ void __launch_bounds__(UP2_THREADS)
upscale2_masks_kernel(const float* __restrict__ g2, const float* __restrict__ feat, const float* __restrict__ hyper,
                      int hyper_bs, float* __restrict__ masks, int B, int Hc, int Wc, int feat_group) {
  __shared__ __align__(16) float hs[128];
  PDL_ENTRY();
  const int Ho = 2 * Hc, Wo = 2 * Wc, per_obj = Ho * Wo;
  const int blocks_per_obj = per_obj / UP2_THREADS;
  const int b = blockIdx.x / blocks_per_obj;
  const int p = (blockIdx.x - b * blocks_per_obj) * UP2_THREADS + threadIdx.x;
  hs[threadIdx.x] = hyper[(long long)b * hyper_bs + threadIdx.x];
  __syncthreads();
  const int Y = p / Wo, X = p - Y * Wo;
  const long long src = ((long long)b * Hc + (Y >> 1)) * Wc + (X >> 1);
  const int q = (Y & 1) * 2 + (X & 1);
  const long long fpix = feat_group > 0 ? ((long long)(b / feat_group) * Ho + Y) * Wo + X : (long long)b * per_obj + p;
  const float4* gp = reinterpret_cast<const float4*>(g2 + src * 128 + q * 32);
  const float4* fp = reinterpret_cast<const float4*>(feat + fpix * 32);
  float4 ga[8], fa[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    ga[j] = gp[j];
    fa[j] = __ldg(fp + j);
  }
  float m[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float v0 = gelu_erf(ga[j].x + fa[j].x), v1 = gelu_erf(ga[j].y + fa[j].y);
    const float v2 = gelu_erf(ga[j].z + fa[j].z), v3 = gelu_erf(ga[j].w + fa[j].w);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float4 h = *reinterpret_cast<const float4*>(hs + k * 32 + 4 * j);
      m[k] = fmaf(v0, h.x, m[k]);
      m[k] = fmaf(v1, h.y, m[k]);
      m[k] = fmaf(v2, h.z, m[k]);
      m[k] = fmaf(v3, h.w, m[k]);
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) masks[(((long long)b * 4 + k) * Ho + Y) * Wo + X] = m[k];
}

// y = L3(relu(L2(relu(L1 x))))  (+ sigmoid); grid (rows, instances); 256 threads; hidden = 256
__global__ void __launch_bounds__(256)
small_mlp3_kernel(const float* __restrict__ x, long long x_row_stride, long long x_inst_stride,
                  const int* __restrict__ row_select, const float* __restrict__ w1, const float* __restrict__ b1,
                  const float* __restrict__ w2, const float* __restrict__ b2, const float* __restrict__ w3,
                  const float* __restrict__ b3, int out_dim, int sigmoid_out, float* __restrict__ y,
                  long long y_row_stride, long long y_inst_stride) {
  PDL_ENTRY();
  __shared__ float h0[256], h1[256];
  const int row = blockIdx.x, inst = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sel = row_select ? row_select[row] : 0;
  const float* xp = x + row * x_row_stride + (inst + sel) * x_inst_stride;
  h0[threadIdx.x] = xp[threadIdx.x];
  __syncthreads();
  const float* W1 = w1 + (long long)inst * 256 * 256;
  const float* W2 = w2 + (long long)inst * 256 * 256;
  const float* W3 = w3 + (long long)inst * out_dim * 256;
  auto layer = [&](const float* Wm, const float* bv, const float* in, float* outp, int n_out, bool relu, bool to_global,
                   bool sig) {
    for (int o = warp; o < n_out; o += 8) {
      const float* wr = Wm + (long long)o * 256;
      float acc = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) acc = fmaf(wr[j * 32 + lane], in[j * 32 + lane], acc);
      acc = warp_sum(acc) + bv[o];
      if (relu) acc = fmaxf(acc, 0.f);
      if (sig) acc = 1.0f / (1.0f + expf(-acc));
      if (lane == 0) {
        if (to_global) outp[o] = acc;
        else outp[o] = acc;
      }
    }
  };
  layer(W1, b1 + inst * 256, h0, h1, 256, true, false, false);
  __syncthreads();
  layer(W2, b2 + inst * 256, h1, h0, 256, true, false, false);
  __syncthreads();
  layer(W3, b3 + inst * out_dim, h0, y + row * y_row_stride + inst * y_inst_stride, out_dim, false, true,
        sigmoid_out != 0);
}

// One block per object.  masks [B,4,HW]; iou [B,4]; score [B].  Writes the chosen low-res mask (gated),
// the chosen token index (for the object pointer) and the chosen IoU.
__global__ void __launch_bounds__(1024)
sam_select_kernel(const float* __restrict__ masks, const float* __restrict__ iou, int iou_stride, int iou_is_logit,
                  const float* __restrict__ score, int score_stride, int multimask, float stab_delta, float stab_thresh,
                  float no_obj_score, float* __restrict__ low_res, int* __restrict__ token_index,
                  float* __restrict__ iou_out, int HW) {
  PDL_ENTRY();
  __shared__ int s_cnt[2];
  __shared__ int s_pick;
  const int b = blockIdx.x;
  const float* mb = masks + (long long)b * 4 * HW;
  const float* ib = iou + (long long)b * iou_stride;
  if (threadIdx.x == 0) {
    s_cnt[0] = 0;
    s_cnt[1] = 0;
  }
  __syncthreads();
  int best = 1;  // torch.argmax: first maximum
  if (ib[2] > ib[best]) best = 2;
  if (ib[3] > ib[best]) best = 3;
  if (!multimask) {
    int ci = 0, cu = 0;
    for (int i = threadIdx.x; i < HW; i += blockDim.x) {
      const float v = mb[i];
      ci += v > stab_delta;
      cu += v > -stab_delta;
    }
    ci = __reduce_add_sync(0xffffffffu, ci);
    cu = __reduce_add_sync(0xffffffffu, cu);
    if ((threadIdx.x & 31) == 0) {
      atomicAdd(&s_cnt[0], ci);
      atomicAdd(&s_cnt[1], cu);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      const float stab = s_cnt[1] > 0 ? (float)s_cnt[0] / (float)s_cnt[1] : 1.0f;
      s_pick = stab >= stab_thresh ? 0 : best;
    }
  } else if (threadIdx.x == 0) {
    s_pick = best;
  }
  __syncthreads();
  const int pick = s_pick;
  const bool present = score[(long long)b * score_stride] > 0.f;
  if (threadIdx.x == 0) {
    // object pointer token: the multimask token when multimask output is on, else always token 0
    token_index[b] = multimask ? pick : 0;
    iou_out[b] = iou_is_logit ? 1.0f / (1.0f + expf(-ib[pick])) : ib[pick];
  }
  const float* src = mb + (long long)pick * HW;
  float* dst = low_res + (long long)b * HW;
  for (int i = threadIdx.x; i < HW; i += blockDim.x) dst[i] = present ? src[i] : no_obj_score;
}

// obj_ptr = lam * ptr + (1 - lam) * no_obj_ptr, lam = score > 0  (sam2_base.py:1146-1156)
__global__ void objptr_mix_kernel(float* __restrict__ ptr, const float* __restrict__ score, int score_stride,
                                  const float* __restrict__ no_obj_ptr, int B, int C) {
  PDL_ENTRY();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * C) return;
  const int b = i / C, c = i - b * C;
  if (!(score[(long long)b * score_stride] > 0.f)) ptr[i] = no_obj_ptr[c];
}

// coords [B,P,2] in model pixels (+0.5 applied by the caller; padding point = 0), labels [B,P] int32
// table: rows 0..3 point_embeddings, row 4 not_a_point_embed; gauss [2,128]; out [B,P,256]
__global__ void point_embed_kernel(const float* __restrict__ coords, const int* __restrict__ labels,
                                   const float* __restrict__ gauss, const float* __restrict__ table, float inv_size,
                                   float* __restrict__ out, int n_points) {
  PDL_ENTRY();
  const int p = blockIdx.x;
  if (p >= n_points) return;
  const int c = threadIdx.x;  // 256
  const int lab = labels[p];
  float v;
  if (lab == -1) {
    v = table[4 * 256 + c];
  } else {
    const float x = 2.f * (coords[p * 2] * inv_size) - 1.f, y = 2.f * (coords[p * 2 + 1] * inv_size) - 1.f;
    const int j = c & 127;
    const float ang = 6.283185307179586f * (x * gauss[j] + y * gauss[128 + j]);
    v = c < 128 ? sinf(ang) : cosf(ang);
    if (lab >= 0 && lab < 4) v += table[lab * 256 + c];
  }
  out[(long long)p * 256 + c] = v;
}

}  // namespace

#define STREAM reinterpret_cast<cudaStream_t>(stream)

extern "C" int usvm_upscale1_ln_gelu(const float* g1, const float* feat_s1, const float* ln_w, const float* ln_b,
                                     float eps, float* out, int B, int Hc, int Wc, int C, int feat_group,
                                     void* stream) {
  if (!g1 || !feat_s1 || !ln_w || !ln_b || !out || C != 64) return USVM_ERR_ARG;
  usvm_launch(upscale1_kernel<64>, dim3(cdiv((long long)B * 4 * Hc * Wc, 8)), dim3(256), 0, STREAM, g1, feat_s1, ln_w, ln_b, eps, out, B, Hc, Wc, feat_group);
  return usvm_check_launch();
}

extern "C" int usvm_upscale2_masks(const float* g2, const float* feat_s0, const float* hyper, int hyper_bs,
                                   float* masks, int B, int Hc, int Wc, int feat_group, void* stream) {
  if (!g2 || !feat_s0 || !hyper || !masks || B <= 0 || ((4 * Hc * Wc) % UP2_THREADS) ||
      (reinterpret_cast<uintptr_t>(g2) & 15) || (reinterpret_cast<uintptr_t>(feat_s0) & 15))
    return USVM_ERR_ARG;
  usvm_launch(upscale2_masks_kernel, dim3(B * (4 * Hc * Wc / UP2_THREADS)), dim3(UP2_THREADS), 0, STREAM, g2, feat_s0, hyper,
              hyper_bs, masks, B, Hc, Wc, feat_group);
  return usvm_check_launch();
}

extern "C" int usvm_small_mlp3(const float* x, long long x_row_stride, long long x_inst_stride, const int* row_select,
                               const float* w1, const float* b1, const float* w2, const float* b2, const float* w3,
                               const float* b3, int out_dim, int sigmoid_out, float* y, long long y_row_stride,
                               long long y_inst_stride, int rows, int instances, void* stream) {
  if (!x || !w1 || !b1 || !w2 || !b2 || !w3 || !b3 || !y || rows <= 0 || instances <= 0 || out_dim <= 0)
    return USVM_ERR_ARG;
  usvm_launch(small_mlp3_kernel, dim3(dim3(rows, instances)), dim3(256), 0, STREAM, x, x_row_stride, x_inst_stride, row_select, w1, b1, w2,
                                                               b2, w3, b3, out_dim, sigmoid_out, y, y_row_stride,
                                                               y_inst_stride);
  return usvm_check_launch();
}

extern "C" int usvm_sam_select(const float* masks, const float* iou, int iou_stride, int iou_is_logit,
                               const float* score, int score_stride, int multimask, float stab_delta,
                               float stab_thresh, float no_obj_score, float* low_res, int* token_index, float* iou_out,
                               int B, int HW, void* stream) {
  if (!masks || !iou || !score || !low_res || !token_index || !iou_out || B <= 0) return USVM_ERR_ARG;
  usvm_launch(sam_select_kernel, dim3(B), dim3(1024), 0, STREAM, masks, iou, iou_stride, iou_is_logit, score, score_stride, multimask,
                                            stab_delta, stab_thresh, no_obj_score, low_res, token_index, iou_out, HW);
  return usvm_check_launch();
}

extern "C" int usvm_objptr_mix(float* ptr, const float* score, int score_stride, const float* no_obj_ptr, int B, int C,
                               void* stream) {
  if (!ptr || !score || !no_obj_ptr || B <= 0) return USVM_ERR_ARG;
  usvm_launch(objptr_mix_kernel, dim3(cdiv((long long)B * C, 256)), dim3(256), 0, STREAM, ptr, score, score_stride, no_obj_ptr, B, C);
  return usvm_check_launch();
}

extern "C" int usvm_point_embed(const float* coords, const int* labels, const float* gauss, const float* table,
                                float image_size, float* out, int n_points, void* stream) {
  if (!coords || !labels || !gauss || !table || !out || n_points <= 0) return USVM_ERR_ARG;
  usvm_launch(point_embed_kernel, dim3(n_points), dim3(256), 0, STREAM, coords, labels, gauss, table, 1.0f / image_size, out, n_points);
  return usvm_check_launch();
}
