// Small-convolution, resize and memory-encoder bandwidth kernels (NHWC activations).
//
//   usvm_conv2d_small       direct k x k / stride s conv (+ LayerNorm2d + GELU), weights staged in shared
//                           memory; covers MaskDownSampler's 3x3/s2 stages (memory_encoder.py:40-56),
//                           PromptEncoder.mask_downscaling 2x2/s2 (prompt_encoder.py:57-66) and
//                           SAM2Base.mask_downsample 4x4/s4 (sam2_base.py:863)
//   usvm_im2col_nhwc        im2col for the one mask-downsampler stage big enough for the tensor cores
//   usvm_dwconv7_ln         CXBlock depthwise 7x7 + LayerNorm2d (memory_encoder.py:104-108), bf16 out
//   usvm_resize_bilinear    F.interpolate(bilinear, align_corners=False) with optional fused
//                           sigmoid*scale+bias / binarise (sam2_base.py:1125-1131,1472-1484)
//   usvm_resize_bilinear_aa antialiased bilinear (sam2_base.py:1079,1177; predictor :342)
#include "common.cuh"
#include "usvm2_b200.h"

namespace {

__device__ __forceinline__ float post_op(float v, int mode, float scale, float bias) {
  if (mode == USVM_POST_SIGMOID_AFFINE) return (1.0f / (1.0f + expf(-v))) * scale + bias;
  if (mode == USVM_POST_BINARIZE_AFFINE) return (v > 0.f ? 1.0f : 0.0f) * scale + bias;
  return v;
}

// Virtual input of the mask down-sampler's first stage: instead of reading a materialised [B, Ho, Wo] image, sample it
// from the low-resolution logits -- F.interpolate(bilinear, align_corners=False) followed by post_op, the arithmetic of
// resize_bilinear_kernel -- so the 1 MiB-per-object upsampled mask never exists (_encode_new_memory, sam2_base.py:1472-1484).
struct UpSrc {
  const float* src;  // [B, hi, wi] or nullptr
  int hi, wi, mode;
  float scale, bias;
};
__device__ __forceinline__ float up_sample(const UpSrc& u, int b, int oy, int ox, int Ho, int Wo) {
  const float sh = (float)u.hi / Ho, sw = (float)u.wi / Wo;
  float fy = sh * (oy + 0.5f) - 0.5f, fx = sw * (ox + 0.5f) - 0.5f;
  fy = fy < 0.f ? 0.f : fy;
  fx = fx < 0.f ? 0.f : fx;
  const int y0 = (int)fy, x0 = (int)fx;
  const int y1 = y0 + (y0 < u.hi - 1 ? 1 : 0), x1 = x0 + (x0 < u.wi - 1 ? 1 : 0);
  const float ly = fy - y0, lx = fx - x0, hy = 1.f - ly, hx = 1.f - lx;
  const float* p = u.src + (long long)b * u.hi * u.wi;
  const float v = hy * (hx * __ldg(p + y0 * u.wi + x0) + lx * __ldg(p + y0 * u.wi + x1)) +
                  ly * (hx * __ldg(p + y1 * u.wi + x0) + lx * __ldg(p + y1 * u.wi + x1));
  return post_op(v, u.mode, u.scale, u.bias);
}

// lanes cover (pixel, output channel) pairs: a warp handles 32 / Cout pixels when Cout < 32 (Cout a power of two),
// one pixel with 2 channels per lane when Cout == 64; weights [k*k*Cin][Cout] staged in shared memory
__global__ void __launch_bounds__(256)
conv2d_small_kernel(const float* __restrict__ x, const float* __restrict__ wt, const float* __restrict__ bias,
                    const float* __restrict__ ln_w, const float* __restrict__ ln_b, float eps, int gelu,
                    float* out_f32, bf16* out_bf16, int B, int H, int W, int Cin, int Cout, int k, int s, int pad,
                    int Ho, int Wo) {
  PDL_ENTRY();
  extern __shared__ float s_w[];
  const int K = k * k * Cin;
  for (int i = threadIdx.x; i < K * Cout; i += blockDim.x) s_w[i] = wt[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int grp = Cout < 32 ? Cout : 32;       // lanes per pixel
  const int ppw = 32 / grp;                    // pixels per warp
  const int cpl = Cout > 32 ? Cout / 32 : 1;   // channels per lane (1 or 2)
  const int c0 = lane % grp;
  const long long total = (long long)B * Ho * Wo;
  const long long warp_id = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  for (long long base = warp_id * ppw; base < total; base += (long long)gridDim.x * 8 * ppw) {
    const long long pix = base + lane / grp;
    const bool live = pix < total;
    float acc0 = 0.f, acc1 = 0.f;
    if (live) {
      const int ox = (int)(pix % Wo), oy = (int)((pix / Wo) % Ho), b = (int)(pix / ((long long)Wo * Ho));
      for (int ky = 0; ky < k; ++ky) {
        const int y = oy * s - pad + ky;
        if (y < 0 || y >= H) continue;
        for (int kx = 0; kx < k; ++kx) {
          const int xx = ox * s - pad + kx;
          if (xx < 0 || xx >= W) continue;
          const float* xp = x + (((long long)b * H + y) * W + xx) * Cin;
          const float* wp = s_w + (ky * k + kx) * Cin * Cout;
          for (int ci = 0; ci < Cin; ++ci) {
            const float v = xp[ci];
            acc0 = fmaf(v, wp[ci * Cout + c0], acc0);
            if (cpl == 2) acc1 = fmaf(v, wp[ci * Cout + c0 + 32], acc1);
          }
        }
      }
      acc0 += bias[c0];
      if (cpl == 2) acc1 += bias[c0 + 32];
    }
    if (ln_w) {
      float sum = acc0 + (cpl == 2 ? acc1 : 0.f);
      for (int o = grp >> 1; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      const float mean = sum / Cout;
      const float d0 = acc0 - mean, d1 = cpl == 2 ? acc1 - mean : 0.f;
      float var = d0 * d0 + d1 * d1;
      for (int o = grp >> 1; o > 0; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
      const float rstd = 1.0f / sqrtf(var / Cout + eps);
      acc0 = d0 * rstd * ln_w[c0] + ln_b[c0];
      if (cpl == 2) acc1 = d1 * rstd * ln_w[c0 + 32] + ln_b[c0 + 32];
    }
    if (gelu) {
      acc0 = gelu_erf(acc0);
      acc1 = gelu_erf(acc1);
    }
    if (live) {
      if (out_f32) out_f32[pix * Cout + c0] = acc0;
      if (out_bf16) out_bf16[pix * Cout + c0] = __float2bfloat16(acc0);
      if (cpl == 2) {
        if (out_f32) out_f32[pix * Cout + c0 + 32] = acc1;
        if (out_bf16) out_bf16[pix * Cout + c0 + 32] = __float2bfloat16(acc1);
      }
    }
  }
}


// Tiled variant for the mask down-sampler stages (Cout in {4, 16, 64}, square output tile of 1024 / Cout pixels per
// CTA): the input footprint and the weights are staged in shared memory with all global loads of a thread in flight at
// once (the kernel above walks taps x channels with one dependent global load per multiply); a thread then produces
// four consecutive output channels of one pixel (float4 weights, broadcast input), LayerNorm statistics are shuffles
// inside the pixel's Cout / 4 lanes, and the store is one float4.
template <int CIN, int COUT>
__global__ void __launch_bounds__(256)
conv2d_tile_kernel(const float* __restrict__ x, const float* __restrict__ wt, const float* __restrict__ bias,
                   const float* __restrict__ ln_w, const float* __restrict__ ln_b, float eps, int gelu,
                   float* __restrict__ out_f32, bf16* __restrict__ out_bf16, int H, int W, int k, int s, int pad, int Ho,
                   int Wo, const UpSrc up) {
  constexpr int LPP = COUT / 4;           // lanes per pixel
  constexpr int PIX = 256 / LPP;          // pixels per CTA
  constexpr int TW = PIX == 256 ? 16 : PIX == 64 ? 8 : 4;  // square tile
  extern __shared__ __align__(16) float cv_smem[];
  const int K = k * k * CIN;
  float* s_w = cv_smem;                   // [K][COUT]
  float* s_x = cv_smem + K * COUT;        // [IH][IW][CIN]
  const int IW = (TW - 1) * s + k;
  const int tiles_x = Wo / TW, tiles_y = Ho / TW;
  const int tile = blockIdx.x;
  const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, b = tile / (tiles_x * tiles_y);
  const int iy0 = ty * TW * s - pad, ix0 = tx * TW * s - pad;
  const int tid = threadIdx.x;
  {  // weights: K * COUT / 4 float4, batches of 8 per thread
    const int nv = K * COUT / 4;
    for (int i0 = tid; i0 < nv; i0 += 256 * 8) {
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (i0 + u * 256 < nv) v[u] = __ldg(reinterpret_cast<const float4*>(wt) + i0 + u * 256);
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (i0 + u * 256 < nv) reinterpret_cast<float4*>(s_w)[i0 + u * 256] = v[u];
    }
  }
  const int p = tid / LPP, c4 = (tid % LPP) * 4;
  const int py = p / TW, px = p - py * TW;
  const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + c4));
  float4 lw = make_float4(1.f, 1.f, 1.f, 1.f), lb = make_float4(0.f, 0.f, 0.f, 0.f);
  if (ln_w) {
    lw = __ldg(reinterpret_cast<const float4*>(ln_w + c4));
    lb = __ldg(reinterpret_cast<const float4*>(ln_b + c4));
  }
  // everything above is a constant of the model (weights, bias, LayerNorm parameters): staged before the
  // programmatic-dependency wait, under the predecessor's tail
  PDL_ENTRY();
  {  // input footprint, zero outside the image; a row of the footprint is contiguous in memory (IW * CIN floats)
    const int row_f = IW * CIN, nf = IW * row_f;
    for (int i0 = tid; i0 < nf; i0 += 256 * 6) {
      float v[6];
#pragma unroll
      for (int u = 0; u < 6; ++u) {
        const int i = i0 + u * 256;
        v[u] = 0.f;
        if (i < nf) {
          const int r = i / row_f, c = i - r * row_f;
          const int y = iy0 + r, xx = ix0 + c / CIN;
          if (y >= 0 && y < H && xx >= 0 && xx < W)
            v[u] = (CIN == 1 && up.src) ? up_sample(up, b, y, xx, H, W)
                                        : __ldg(x + (((long long)b * H + y) * W + ix0) * CIN + c);
        }
      }
#pragma unroll
      for (int u = 0; u < 6; ++u)
        if (i0 + u * 256 < nf) s_x[i0 + u * 256] = v[u];
    }
  }
  __syncthreads();
  float4 acc = b4;
  for (int ky = 0; ky < k; ++ky) {
    for (int kx = 0; kx < k; ++kx) {
      const float* xp = s_x + ((py * s + ky) * IW + px * s + kx) * CIN;
      const float* wp = s_w + (ky * k + kx) * CIN * COUT + c4;
#pragma unroll
      for (int ci = 0; ci < CIN; ++ci) {
        const float v = xp[ci];
        const float4 w4 = *reinterpret_cast<const float4*>(wp + ci * COUT);
        acc.x = fmaf(v, w4.x, acc.x); acc.y = fmaf(v, w4.y, acc.y);
        acc.z = fmaf(v, w4.z, acc.z); acc.w = fmaf(v, w4.w, acc.w);
      }
    }
  }
  if (ln_w) {
    float sum = (acc.x + acc.y) + (acc.z + acc.w);
#pragma unroll
    for (int o = LPP >> 1; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum / COUT;
    acc.x -= mean; acc.y -= mean; acc.z -= mean; acc.w -= mean;
    float var = acc.x * acc.x + acc.y * acc.y + acc.z * acc.z + acc.w * acc.w;
#pragma unroll
    for (int o = LPP >> 1; o > 0; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
    const float rstd = 1.0f / sqrtf(var / COUT + eps);
    acc.x = acc.x * rstd * lw.x + lb.x; acc.y = acc.y * rstd * lw.y + lb.y;
    acc.z = acc.z * rstd * lw.z + lb.z; acc.w = acc.w * rstd * lw.w + lb.w;
  }
  if (gelu) {
    acc.x = gelu_erf(acc.x); acc.y = gelu_erf(acc.y); acc.z = gelu_erf(acc.z); acc.w = gelu_erf(acc.w);
  }
  const long long pix = ((long long)b * Ho + ty * TW + py) * Wo + tx * TW + px;
  if (out_f32) *reinterpret_cast<float4*>(out_f32 + pix * COUT + c4) = acc;
  if (out_bf16) {
    uint2 pk;
    pk.x = pack_bf16x2(acc.x, acc.y);
    pk.y = pack_bf16x2(acc.z, acc.w);
    *reinterpret_cast<uint2*>(out_bf16 + pix * COUT + c4) = pk;
  }
}

// Batched variant: a thread produces four consecutive output channels of a 2 x 2 pixel block (one float4 weight load
// feeds 16 FMAs instead of 4) and the CTA an output tile four times as large, so the weights (36 KB for 16 -> 64) are
// staged once per 4x the pixels.  Same tap order per output: results identical to conv2d_tile_kernel.
template <int CIN, int COUT>
__global__ void __launch_bounds__(256)
conv2d_tile4_kernel(const float* __restrict__ x, const float* __restrict__ wt, const float* __restrict__ bias,
                    const float* __restrict__ ln_w, const float* __restrict__ ln_b, float eps, int gelu,
                    float* __restrict__ out_f32, bf16* __restrict__ out_bf16, int H, int W, int k, int s, int pad, int Ho,
                    int Wo, const UpSrc up) {
  constexpr int LPP = COUT / 4;           // lanes per pixel block
  constexpr int GRP = 256 / LPP;          // 2 x 2 pixel blocks per CTA
  constexpr int GW = GRP == 256 ? 16 : GRP == 64 ? 8 : 4;  // square arrangement of the blocks
  constexpr int TW = 2 * GW;              // output tile edge
  extern __shared__ __align__(16) float cv_smem[];
  const int K = k * k * CIN;
  float* s_w = cv_smem;                   // [K][COUT]
  float* s_x = cv_smem + K * COUT;        // [IW][IW][CIN]
  const int IW = (TW - 1) * s + k;
  const int tiles_x = Wo / TW, tiles_y = Ho / TW;
  const int tile = blockIdx.x;
  const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, b = tile / (tiles_x * tiles_y);
  const int iy0 = ty * TW * s - pad, ix0 = tx * TW * s - pad;
  const int tid = threadIdx.x;
  {
    const int nv = K * COUT / 4;
    for (int i0 = tid; i0 < nv; i0 += 256 * 8) {
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (i0 + u * 256 < nv) v[u] = __ldg(reinterpret_cast<const float4*>(wt) + i0 + u * 256);
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (i0 + u * 256 < nv) reinterpret_cast<float4*>(s_w)[i0 + u * 256] = v[u];
    }
  }
  const int gidx = tid / LPP, c4 = (tid % LPP) * 4;
  const int gy = gidx / GW, gx = gidx - gy * GW;
  const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + c4));
  float4 lw = make_float4(1.f, 1.f, 1.f, 1.f), lb = make_float4(0.f, 0.f, 0.f, 0.f);
  if (ln_w) {
    lw = __ldg(reinterpret_cast<const float4*>(ln_w + c4));
    lb = __ldg(reinterpret_cast<const float4*>(ln_b + c4));
  }
  PDL_ENTRY();
  {
    const int row_f = IW * CIN, nf = IW * row_f;
    for (int i0 = tid; i0 < nf; i0 += 256 * 6) {
      float v[6];
#pragma unroll
      for (int u = 0; u < 6; ++u) {
        const int i = i0 + u * 256;
        v[u] = 0.f;
        if (i < nf) {
          const int r = i / row_f, c = i - r * row_f;
          const int y = iy0 + r, xx = ix0 + c / CIN;
          if (y >= 0 && y < H && xx >= 0 && xx < W)
            v[u] = (CIN == 1 && up.src) ? up_sample(up, b, y, xx, H, W)
                                        : __ldg(x + (((long long)b * H + y) * W + ix0) * CIN + c);
        }
      }
#pragma unroll
      for (int u = 0; u < 6; ++u)
        if (i0 + u * 256 < nf) s_x[i0 + u * 256] = v[u];
    }
  }
  __syncthreads();
  float4 acc[4] = {b4, b4, b4, b4};
  for (int ky = 0; ky < k; ++ky) {
    for (int kx = 0; kx < k; ++kx) {
      const float* wp = s_w + (ky * k + kx) * CIN * COUT + c4;
      const float* xp[4];
#pragma unroll
      for (int q = 0; q < 4; ++q)
        xp[q] = s_x + (((2 * gy + (q >> 1)) * s + ky) * IW + (2 * gx + (q & 1)) * s + kx) * CIN;
#pragma unroll
      for (int ci = 0; ci < CIN; ++ci) {
        const float4 w4 = *reinterpret_cast<const float4*>(wp + ci * COUT);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float v = xp[q][ci];
          acc[q].x = fmaf(v, w4.x, acc[q].x); acc[q].y = fmaf(v, w4.y, acc[q].y);
          acc[q].z = fmaf(v, w4.z, acc[q].z); acc[q].w = fmaf(v, w4.w, acc[q].w);
        }
      }
    }
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float4 a = acc[q];
    if (ln_w) {
      float sum = (a.x + a.y) + (a.z + a.w);
#pragma unroll
      for (int o = LPP >> 1; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      const float mean = sum / COUT;
      a.x -= mean; a.y -= mean; a.z -= mean; a.w -= mean;
      float var = a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
#pragma unroll
      for (int o = LPP >> 1; o > 0; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
      const float rstd = 1.0f / sqrtf(var / COUT + eps);
      a.x = a.x * rstd * lw.x + lb.x; a.y = a.y * rstd * lw.y + lb.y;
      a.z = a.z * rstd * lw.z + lb.z; a.w = a.w * rstd * lw.w + lb.w;
    }
    if (gelu) {
      a.x = gelu_erf(a.x); a.y = gelu_erf(a.y); a.z = gelu_erf(a.z); a.w = gelu_erf(a.w);
    }
    const long long pix = ((long long)b * Ho + ty * TW + 2 * gy + (q >> 1)) * Wo + tx * TW + 2 * gx + (q & 1);
    if (out_f32) *reinterpret_cast<float4*>(out_f32 + pix * COUT + c4) = a;
    if (out_bf16) {
      uint2 pk;
      pk.x = pack_bf16x2(a.x, a.y);
      pk.y = pack_bf16x2(a.z, a.w);
      *reinterpret_cast<uint2*>(out_bf16 + pix * COUT + c4) = pk;
    }
  }
}

template <int CIN, int COUT>
int launch_conv_tile(const float* x, const float* w, const float* bias, const float* ln_w, const float* ln_b, float eps,
                     int gelu, float* out_f32, bf16* out_bf16, int B, int H, int W, int k, int s, int pad, int Ho, int Wo,
                     cudaStream_t stream, const UpSrc up = UpSrc{nullptr, 0, 0, 0, 0.f, 0.f}) {
  constexpr int PIX = 256 / (COUT / 4);
  constexpr int TW = PIX == 256 ? 16 : PIX == 64 ? 8 : 4;
  // batched path: 2 x 2 pixels per thread once the larger tiles still fill the device twice over
  if (Ho % (2 * TW) == 0 && Wo % (2 * TW) == 0 && (long long)B * (Ho / (2 * TW)) * (Wo / (2 * TW)) >= 2 * 148) {
    const int IW4 = (2 * TW - 1) * s + k;
    const size_t smem4 = ((size_t)k * k * CIN * COUT + (size_t)IW4 * IW4 * CIN) * sizeof(float);
    if (smem4 <= 96 * 1024) {
      static UsvmPerDeviceOnce configured4 = {};
      if (usvm_need_setup(configured4)) {
        if (cudaFuncSetAttribute(conv2d_tile4_kernel<CIN, COUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024) !=
            cudaSuccess)
          return USVM_ERR_CUDA;
        usvm_setup_done(configured4);
      }
      usvm_launch(conv2d_tile4_kernel<CIN, COUT>, dim3(B * (Ho / (2 * TW)) * (Wo / (2 * TW))), dim3(256), smem4, stream, x, w,
                  bias, ln_w, ln_b, eps, gelu, out_f32, out_bf16, H, W, k, s, pad, Ho, Wo, up);
      return usvm_check_launch();
    }
  }
  const int IW = (TW - 1) * s + k;
  const size_t smem = ((size_t)k * k * CIN * COUT + (size_t)IW * IW * CIN) * sizeof(float);
  static UsvmPerDeviceOnce configured = {};
  if (usvm_need_setup(configured)) {
    if (cudaFuncSetAttribute(conv2d_tile_kernel<CIN, COUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024) !=
        cudaSuccess)
      return USVM_ERR_CUDA;
    usvm_setup_done(configured);
  }
  if (smem > 96 * 1024) return USVM_ERR_ARG;
  usvm_launch(conv2d_tile_kernel<CIN, COUT>, dim3(B * (Ho / TW) * (Wo / TW)), dim3(256), smem, stream, x, w, bias, ln_w, ln_b,
              eps, gelu, out_f32, out_bf16, H, W, k, s, pad, Ho, Wo, up);
  return usvm_check_launch();
}

// A[pix, (ky*k + kx)*C + c] = x[b, oy*s - pad + ky, ox*s - pad + kx, c] (zero outside), bf16
__global__ void im2col_nhwc_kernel(const float* __restrict__ x, bf16* __restrict__ A, int B, int H, int W, int C, int k,
                                   int s, int pad, int Ho, int Wo) {
  PDL_ENTRY();
  const int K = k * k * C;
  const long long total = (long long)B * Ho * Wo * K;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int kk = (int)(i % K);
    const long long pix = i / K;
    const int c = kk % C, tap = kk / C, ky = tap / k, kx = tap - ky * k;
    const int ox = (int)(pix % Wo), oy = (int)((pix / Wo) % Ho), b = (int)(pix / ((long long)Wo * Ho));
    const int y = oy * s - pad + ky, xx = ox * s - pad + kx;
    float v = 0.f;
    if (y >= 0 && y < H && xx >= 0 && xx < W) v = x[(((long long)b * H + y) * W + xx) * C + c];
    A[i] = __float2bfloat16(v);
  }
}

// same, 8 channels per thread (C % 8 == 0): two 16-byte loads, one 16-byte store, index arithmetic once per 8 elements
// (the scalar kernel took 151 us for the 32-object mask down-sampler stage: 18.9 M two-byte stores)
__global__ void im2col_nhwc_vec8_kernel(const float* __restrict__ x, bf16* __restrict__ A, int B, int H, int W, int C, int k,
                                        int s, int pad, int Ho, int Wo) {
  PDL_ENTRY();
  const int C8 = C >> 3, K8 = k * k * C8;
  const long long total = (long long)B * Ho * Wo * K8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int kk = (int)(i % K8);
    const long long pix = i / K8;
    const int c8 = kk % C8, tap = kk / C8, ky = tap / k, kx = tap - ky * k;
    const int ox = (int)(pix % Wo), oy = (int)((pix / Wo) % Ho), b = (int)(pix / ((long long)Wo * Ho));
    const int y = oy * s - pad + ky, xx = ox * s - pad + kx;
    uint4 o = make_uint4(0u, 0u, 0u, 0u);
    if (y >= 0 && y < H && xx >= 0 && xx < W) {
      const float4* p = reinterpret_cast<const float4*>(x + (((long long)b * H + y) * W + xx) * C + c8 * 8);
      const float4 a = __ldg(p), c = __ldg(p + 1);
      o.x = pack_bf16x2(a.x, a.y);
      o.y = pack_bf16x2(a.z, a.w);
      o.z = pack_bf16x2(c.x, c.y);
      o.w = pack_bf16x2(c.z, c.w);
    }
    reinterpret_cast<uint4*>(A)[i] = o;
  }
}

// depthwise 7x7 (pad 3) + LayerNorm over channels; one warp per pixel, C = 256 -> 8 channels per lane
template <int CPL>
__global__ void __launch_bounds__(256)
dwconv7_ln_kernel(const float* __restrict__ x, const float* __restrict__ wt /* [49][C] */, const float* __restrict__ bias,
                  const float* __restrict__ ln_w, const float* __restrict__ ln_b, float eps, bf16* __restrict__ out,
                  int B, int H, int W) {
  PDL_ENTRY();
  constexpr int C = CPL * 32;
  const int lane = threadIdx.x & 31;
  const long long pix = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (pix >= (long long)B * H * W) return;
  const int ox = (int)(pix % W), oy = (int)((pix / W) % H), b = (int)(pix / ((long long)W * H));
  float acc[CPL];
#pragma unroll
  for (int j = 0; j < CPL; ++j) acc[j] = bias[j * 32 + lane];
  for (int ky = 0; ky < 7; ++ky) {
    const int y = oy - 3 + ky;
    if (y < 0 || y >= H) continue;
    for (int kx = 0; kx < 7; ++kx) {
      const int xx = ox - 3 + kx;
      if (xx < 0 || xx >= W) continue;
      const float* xp = x + (((long long)b * H + y) * W + xx) * C;
      const float* wp = wt + (ky * 7 + kx) * C;
#pragma unroll
      for (int j = 0; j < CPL; ++j) acc[j] = fmaf(xp[j * 32 + lane], wp[j * 32 + lane], acc[j]);
    }
  }
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < CPL; ++j) s += acc[j];
  const float mean = warp_sum(s) / C;
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < CPL; ++j) {
    acc[j] -= mean;
    q = fmaf(acc[j], acc[j], q);
  }
  const float rstd = 1.0f / sqrtf(warp_sum(q) / C + eps);
#pragma unroll
  for (int j = 0; j < CPL; ++j) {
    const int c = j * 32 + lane;
    out[pix * C + c] = __float2bfloat16(acc[j] * rstd * ln_w[c] + ln_b[c]);
  }
}


// Tiled variant (C = 256, W % 8 == 0): one CTA per 8 consecutive pixels of an image row, thread <-> channel.  The
// 7 x 14 input footprint of the tile is staged in shared memory with all loads in flight at once (the one-warp-per-pixel
// kernel above re-reads every input 49 times and waits for L2 tap by tap); the thread keeps its channel's 49 weights
// in registers, so the inner loop is 98 conflict-free LDS and 392 FMA.  LayerNorm statistics are block reductions.
// TR = rows of the output tile: 1 for the latency path (128 CTAs at one object), 8 for the batched path -- an 8 x 8 tile
// reads a 14 x 14 footprint (3.1 x its output) instead of 7 x 14 per 8 pixels (12 x): the 32-object launch was bound by
// those shared-memory / L2 re-reads (110 us).  Taps are accumulated in the same order in both, results are identical.
constexpr int DW_TILE = 8;
constexpr int DW_FOOT = DW_TILE + 6;
template <int TR>
__global__ void __launch_bounds__(256)
dwconv7_ln_tile_kernel(const float* __restrict__ x, const float* __restrict__ wt /* [49][256] */,
                       const float* __restrict__ bias, const float* __restrict__ ln_w, const float* __restrict__ ln_b,
                       float eps, bf16* __restrict__ out, int B, int H, int W) {
  constexpr int C = 256;
  constexpr int FR = TR + 6;        // footprint rows
  constexpr int NP = TR * DW_TILE;  // output pixels per CTA
  extern __shared__ __align__(16) float dw_smem[];
  float* xs = dw_smem;                        // [FR][14][256]
  float* red = dw_smem + FR * DW_FOOT * C;    // [8 warps][NP pixels]
  const int c = threadIdx.x, warp = c >> 5, lane = c & 31;
  // constants of the model first (49 taps, bias, LayerNorm parameters of this thread's channel): before the
  // programmatic-dependency wait
  float w[49];
#pragma unroll
  for (int k = 0; k < 49; ++k) w[k] = __ldg(wt + k * C + c);
  const float bc = __ldg(bias + c);
  const float lw = __ldg(ln_w + c), lb = __ldg(ln_b + c);
  PDL_ENTRY();
  const int tiles_x = W / DW_TILE, tiles_y = H / TR;
  const int tile = blockIdx.x;
  const int ox0 = (tile % tiles_x) * DW_TILE, oy0 = ((tile / tiles_x) % tiles_y) * TR, b = tile / (tiles_x * tiles_y);
  // stage the footprint: FR * 14 * 64 float4, zero outside the image
  constexpr int NV = FR * DW_FOOT * (C / 4);
  for (int i0 = c; i0 < NV; i0 += 256 * 7) {
    float4 v[7];
#pragma unroll
    for (int u = 0; u < 7; ++u) {
      const int i = i0 + u * 256;
      v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i < NV) {
        const int r = i / (DW_FOOT * 64), rem = i - r * (DW_FOOT * 64), px = rem >> 6, c4 = rem & 63;
        const int y = oy0 - 3 + r, xx = ox0 - 3 + px;
        if (y >= 0 && y < H && xx >= 0 && xx < W)
          v[u] = __ldg(reinterpret_cast<const float4*>(x + (((long long)b * H + y) * W + xx) * C) + c4);
      }
    }
#pragma unroll
    for (int u = 0; u < 7; ++u) {
      const int i = i0 + u * 256;
      if (i < NV) *reinterpret_cast<float4*>(xs + 4 * i) = v[u];
    }
  }
  float acc[TR][DW_TILE];
#pragma unroll
  for (int q = 0; q < TR; ++q)
#pragma unroll
    for (int o = 0; o < DW_TILE; ++o) acc[q][o] = bc;
  __syncthreads();
#pragma unroll
  for (int r = 0; r < FR; ++r) {
#pragma unroll
    for (int px = 0; px < DW_FOOT; ++px) {
      const float xv = xs[(r * DW_FOOT + px) * C + c];
#pragma unroll
      for (int q = 0; q < TR; ++q) {
        const int ky = r - q;
        if (ky >= 0 && ky < 7) {
#pragma unroll
          for (int o = 0; o < DW_TILE; ++o) {
            const int kx = px - o;
            if (kx >= 0 && kx < 7) acc[q][o] = fmaf(xv, w[ky * 7 + kx], acc[q][o]);
          }
        }
      }
    }
  }
  if constexpr (TR > 1) {
    // LayerNorm of the batched tile: transpose through shared memory (the footprint is dead) so that warp w owns pixels
    // 8w .. 8w + 7 with 8 channels per lane -- 8 loads + one warp reduction per statistic per pixel instead of every
    // thread reducing every pixel (10 shuffles + 16 shared loads per pixel and thread: that tail cost more than the conv)
    __syncthreads();  // every thread is done reading the footprint
#pragma unroll
    for (int q = 0; q < TR; ++q)
#pragma unroll
      for (int o = 0; o < DW_TILE; ++o) xs[(q * DW_TILE + o) * C + c] = acc[q][o];
    __syncthreads();
    float lwv[8], lbv[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      lwv[j] = __ldg(ln_w + j * 32 + lane);
      lbv[j] = __ldg(ln_b + j * 32 + lane);
    }
#pragma unroll 1
    for (int p = warp * (NP / 8); p < (warp + 1) * (NP / 8); ++p) {
      float v[8], s = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        v[j] = xs[p * C + j * 32 + lane];
        s += v[j];
      }
      const float mean = warp_sum(s) / C;
      float qq = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        v[j] -= mean;
        qq = fmaf(v[j], v[j], qq);
      }
      const float rstd = 1.0f / sqrtf(warp_sum(qq) / C + eps);
      const long long pix = ((long long)b * H + oy0 + p / DW_TILE) * W + ox0 + p % DW_TILE;
#pragma unroll
      for (int j = 0; j < 8; ++j) out[pix * C + j * 32 + lane] = __float2bfloat16(v[j] * rstd * lwv[j] + lbv[j]);
    }
    return;
  }
  // LayerNorm over the 256 channels of each of the NP pixels (two-pass)
#pragma unroll
  for (int q = 0; q < TR; ++q)
#pragma unroll
    for (int o = 0; o < DW_TILE; ++o) {
      const float s = warp_sum(acc[q][o]);
      if (lane == 0) red[warp * NP + q * DW_TILE + o] = s;
    }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < TR; ++q)
#pragma unroll
    for (int o = 0; o < DW_TILE; ++o) {
      float s = 0.f;
#pragma unroll
      for (int wv = 0; wv < 8; ++wv) s += red[wv * NP + q * DW_TILE + o];
      acc[q][o] -= s / C;
    }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < TR; ++q)
#pragma unroll
    for (int o = 0; o < DW_TILE; ++o) {
      const float s = warp_sum(acc[q][o] * acc[q][o]);
      if (lane == 0) red[warp * NP + q * DW_TILE + o] = s;
    }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < TR; ++q) {
    const long long pix0 = ((long long)b * H + oy0 + q) * W + ox0;
#pragma unroll
    for (int o = 0; o < DW_TILE; ++o) {
      float s = 0.f;
#pragma unroll
      for (int wv = 0; wv < 8; ++wv) s += red[wv * NP + q * DW_TILE + o];
      const float rstd = 1.0f / sqrtf(s / C + eps);
      out[(pix0 + o) * C + c] = __float2bfloat16(acc[q][o] * rstd * lw + lb);
    }
  }
}


// F.interpolate(mode="bilinear", align_corners=False); planes = N*C
__global__ void resize_bilinear_kernel(const float* __restrict__ x, float* __restrict__ y, long long planes, int Hi,
                                       int Wi, int Ho, int Wo, int mode, float pscale, float pbias) {
  PDL_ENTRY();
  const float sh = (float)Hi / Ho, sw = (float)Wi / Wo;
  const long long total = planes * Ho * Wo;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(i % Wo), oy = (int)((i / Wo) % Ho);
    const long long pl = i / ((long long)Wo * Ho);
    float fy = sh * (oy + 0.5f) - 0.5f, fx = sw * (ox + 0.5f) - 0.5f;
    fy = fy < 0.f ? 0.f : fy;
    fx = fx < 0.f ? 0.f : fx;
    const int y0 = (int)fy, x0 = (int)fx;
    const int y1 = y0 + (y0 < Hi - 1 ? 1 : 0), x1 = x0 + (x0 < Wi - 1 ? 1 : 0);
    const float ly = fy - y0, lx = fx - x0, hy = 1.f - ly, hx = 1.f - lx;
    const float* p = x + pl * Hi * Wi;
    const float v = hy * (hx * p[(long long)y0 * Wi + x0] + lx * p[(long long)y0 * Wi + x1]) +
                    ly * (hx * p[(long long)y1 * Wi + x0] + lx * p[(long long)y1 * Wi + x1]);
    y[i] = post_op(v, mode, pscale, pbias);
  }
}

// SAM2Base._apply_non_overlapping_constraints (sam2_base.py:1663-1681): per pixel, only the object with the highest logit
// of its group keeps it (ties: lowest index, as torch.argmax), the others are clamped to <= -10; then post_op.
// One thread per (group, pixel): the group's objects are read with stride HW (coalesced across the warp).
__global__ void non_overlap_kernel(const float* __restrict__ x, float* __restrict__ y, int B, long long HW, int group,
                                   int mode, float pscale, float pbias) {
  PDL_ENTRY();
  const int groups = B / group;
  const long long total = (long long)groups * HW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long g = i / HW, p = i - g * HW;
    const float* xp = x + g * group * HW + p;
    float best = xp[0];
    int win = 0;
    for (int b = 1; b < group; ++b) {
      const float v = xp[(long long)b * HW];
      if (v > best) { best = v; win = b; }
    }
    float* yp = y + g * group * HW + p;
    for (int b = 0; b < group; ++b) {
      const float v = xp[(long long)b * HW];
      yp[(long long)b * HW] = post_op(b == win ? v : fminf(v, -10.0f), mode, pscale, pbias);
    }
  }
}

// antialiased bilinear (triangle filter widened by the scale), ATen _upsample_bilinear2d_aa semantics
__device__ __forceinline__ void aa_span(int o, float scale, int in_size, int& lo, int& size, float& center,
                                        float& invscale) {
  const float support = scale >= 1.f ? scale : 1.f;  // interp_size(2) * 0.5 * scale
  invscale = scale >= 1.f ? 1.f / scale : 1.f;
  center = scale * (o + 0.5f);
  lo = max((int)(center - support + 0.5f), 0);
  size = min((int)(center + support + 0.5f), in_size) - lo;
}
__global__ void resize_bilinear_aa_kernel(const float* __restrict__ x, float* __restrict__ y, long long planes, int Hi,
                                          int Wi, int Ho, int Wo, int binarize_half) {
  PDL_ENTRY();
  const float sh = (float)Hi / Ho, sw = (float)Wi / Wo;
  const long long total = planes * Ho * Wo;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(i % Wo), oy = (int)((i / Wo) % Ho);
    const long long pl = i / ((long long)Wo * Ho);
    int ylo, ysz, xlo, xsz;
    float yc, yinv, xc, xinv;
    aa_span(oy, sh, Hi, ylo, ysz, yc, yinv);
    aa_span(ox, sw, Wi, xlo, xsz, xc, xinv);
    float wysum = 0.f, wxsum = 0.f;
    for (int j = 0; j < ysz; ++j) wysum += fmaxf(0.f, 1.f - fabsf((j + ylo - yc + 0.5f) * yinv));
    for (int j = 0; j < xsz; ++j) wxsum += fmaxf(0.f, 1.f - fabsf((j + xlo - xc + 0.5f) * xinv));
    const float* p = x + pl * Hi * Wi;
    // ATen resamples horizontally first, then vertically; accumulate in the same order
    float acc = 0.f;
    for (int jy = 0; jy < ysz; ++jy) {
      const float wy = fmaxf(0.f, 1.f - fabsf((jy + ylo - yc + 0.5f) * yinv)) / wysum;
      float row = 0.f;
      for (int jx = 0; jx < xsz; ++jx) {
        const float wx = fmaxf(0.f, 1.f - fabsf((jx + xlo - xc + 0.5f) * xinv)) / wxsum;
        row = fmaf(wx, p[(long long)(ylo + jy) * Wi + xlo + jx], row);
      }
      acc = fmaf(wy, row, acc);
    }
    y[i] = binarize_half ? (acc >= 0.5f ? 1.f : 0.f) : acc;
  }
}

inline int grid_for(long long total, int threads = 256) {
  long long b = (total + threads - 1) / threads;
  const long long cap = 148LL * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

#define STREAM reinterpret_cast<cudaStream_t>(stream)

extern "C" int usvm_conv2d_mask_first(const float* low, int hi, int wi, int post_mode, float post_scale, float post_bias,
                                      const float* w_kkio, const float* bias, const float* ln_w, const float* ln_b,
                                      float eps, int gelu, float* out_f32, int B, int H, int W, int k, int stride, int pad,
                                      void* stream) {
  if (!low || !w_kkio || !bias || !out_f32 || hi <= 0 || wi <= 0 || H <= 0 || W <= 0) return USVM_ERR_ARG;
  const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  if (k > 4 || stride > 4 || (Ho % 16) || (Wo % 16)) return USVM_ERR_ARG;
  if ((reinterpret_cast<uintptr_t>(w_kkio) & 15) || (reinterpret_cast<uintptr_t>(bias) & 15) ||
      (reinterpret_cast<uintptr_t>(out_f32) & 15) ||
      (ln_w && ((reinterpret_cast<uintptr_t>(ln_w) & 15) || (reinterpret_cast<uintptr_t>(ln_b) & 15))))
    return USVM_ERR_ARG;
  return launch_conv_tile<1, 4>(low, w_kkio, bias, ln_w, ln_b, eps, gelu, out_f32, nullptr, B, H, W, k, stride, pad, Ho, Wo,
                                STREAM, UpSrc{low, hi, wi, post_mode, post_scale, post_bias});
}

extern "C" int usvm_conv2d_small(const float* x, const float* w_kkio, const float* bias, const float* ln_w,
                                 const float* ln_b, float eps, int gelu, float* out_f32, void* out_bf16, int B, int H,
                                 int W, int Cin, int Cout, int k, int stride, int pad, void* stream) {
  if (!x || !w_kkio || !bias || (!out_f32 && !out_bf16) || Cin <= 0) return USVM_ERR_ARG;
  if (!(Cout == 1 || Cout == 2 || Cout == 4 || Cout == 8 || Cout == 16 || Cout == 32 || Cout == 64)) return USVM_ERR_ARG;
  const size_t smem = (size_t)k * k * Cin * Cout * sizeof(float);
  if (smem > 48 * 1024) return USVM_ERR_ARG;
  const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  {  // mask down-sampler shapes: shared-memory tiled kernel
    bf16* ob = reinterpret_cast<bf16*>(out_bf16);
    const bool aligned = !(reinterpret_cast<uintptr_t>(w_kkio) & 15) && !(reinterpret_cast<uintptr_t>(bias) & 15) &&
                         !(reinterpret_cast<uintptr_t>(out_f32) & 15) && !(reinterpret_cast<uintptr_t>(out_bf16) & 7) &&
                         (!ln_w || (!(reinterpret_cast<uintptr_t>(ln_w) & 15) && !(reinterpret_cast<uintptr_t>(ln_b) & 15)));
    cudaStream_t st = STREAM;
    if (aligned && k <= 4 && stride <= 4) {
      if (Cin == 1 && Cout == 4 && Ho % 16 == 0 && Wo % 16 == 0)
        return launch_conv_tile<1, 4>(x, w_kkio, bias, ln_w, ln_b, eps, gelu, out_f32, ob, B, H, W, k, stride, pad, Ho, Wo, st);
      if (Cin == 4 && Cout == 16 && Ho % 8 == 0 && Wo % 8 == 0)
        return launch_conv_tile<4, 16>(x, w_kkio, bias, ln_w, ln_b, eps, gelu, out_f32, ob, B, H, W, k, stride, pad, Ho, Wo, st);
      if (Cin == 16 && Cout == 64 && Ho % 4 == 0 && Wo % 4 == 0)
        return launch_conv_tile<16, 64>(x, w_kkio, bias, ln_w, ln_b, eps, gelu, out_f32, ob, B, H, W, k, stride, pad, Ho, Wo, st);
    }
  }
  const long long pixels = (long long)B * Ho * Wo;
  const int ppw = Cout < 32 ? 32 / Cout : 1;
  const int grid = (int)max(1LL, min((long long)148 * 8, (pixels + 8LL * ppw - 1) / (8LL * ppw)));
  usvm_launch(conv2d_small_kernel, dim3(grid), dim3(256), smem, STREAM, x, w_kkio, bias, ln_w, ln_b, eps, gelu, out_f32,
                                                   reinterpret_cast<bf16*>(out_bf16), B, H, W, Cin, Cout, k, stride,
                                                   pad, Ho, Wo);
  return usvm_check_launch();
}

extern "C" int usvm_im2col_nhwc(const float* x, void* A, int B, int H, int W, int C, int k, int stride, int pad,
                                void* stream) {
  if (!x || !A) return USVM_ERR_ARG;
  const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  if ((C % 8) == 0 && !(reinterpret_cast<uintptr_t>(x) & 15) && !(reinterpret_cast<uintptr_t>(A) & 15)) {
    usvm_launch(im2col_nhwc_vec8_kernel, dim3(grid_for((long long)B * Ho * Wo * k * k * (C / 8))), dim3(256), 0, STREAM, x,
                reinterpret_cast<bf16*>(A), B, H, W, C, k, stride, pad, Ho, Wo);
    return usvm_check_launch();
  }
  usvm_launch(im2col_nhwc_kernel, dim3(grid_for((long long)B * Ho * Wo * k * k * C)), dim3(256), 0, STREAM, 
      x, reinterpret_cast<bf16*>(A), B, H, W, C, k, stride, pad, Ho, Wo);
  return usvm_check_launch();
}

extern "C" int usvm_dwconv7_ln(const float* x, const float* w_49c, const float* bias, const float* ln_w,
                               const float* ln_b, float eps, void* out_bf16, int B, int H, int W, int C,
                               void* stream) {
  if (!x || !w_49c || !bias || !ln_w || !ln_b || !out_bf16 || C != 256) return USVM_ERR_ARG;
  if (W % DW_TILE == 0 && !(reinterpret_cast<uintptr_t>(x) & 15)) {
    // batched path: 8 x 8 output tiles once they fill the device
    if ((H % 8) == 0 && (long long)B * (H / 8) * (W / DW_TILE) >= 148) {
      const int smem = (14 * DW_FOOT * 256 + 8 * 8 * DW_TILE) * (int)sizeof(float);
      static UsvmPerDeviceOnce configured8 = {};
      if (usvm_need_setup(configured8)) {
        if (cudaFuncSetAttribute(dwconv7_ln_tile_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
          return USVM_ERR_CUDA;
        usvm_setup_done(configured8);
      }
      usvm_launch(dwconv7_ln_tile_kernel<8>, dim3(B * (H / 8) * (W / DW_TILE)), dim3(256), smem, STREAM, x, w_49c, bias, ln_w,
                  ln_b, eps, reinterpret_cast<bf16*>(out_bf16), B, H, W);
      return usvm_check_launch();
    }
    const int smem = (7 * DW_FOOT * 256 + 8 * DW_TILE) * (int)sizeof(float);
    static UsvmPerDeviceOnce configured = {};
    if (usvm_need_setup(configured)) {
      if (cudaFuncSetAttribute(dwconv7_ln_tile_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
        return USVM_ERR_CUDA;
      usvm_setup_done(configured);
    }
    usvm_launch(dwconv7_ln_tile_kernel<1>, dim3(B * H * (W / DW_TILE)), dim3(256), smem, STREAM, x, w_49c, bias, ln_w, ln_b,
                eps, reinterpret_cast<bf16*>(out_bf16), B, H, W);
    return usvm_check_launch();
  }
  usvm_launch(dwconv7_ln_kernel<8>, dim3(cdiv((long long)B * H * W, 8)), dim3(256), 0, STREAM, x, w_49c, bias, ln_w, ln_b, eps,
                                                                           reinterpret_cast<bf16*>(out_bf16), B, H, W);
  return usvm_check_launch();
}

extern "C" int usvm_resize_bilinear(const float* x, float* y, long long planes, int Hi, int Wi, int Ho, int Wo,
                                    int post_mode, float post_scale, float post_bias, void* stream) {
  if (!x || !y || planes <= 0 || Hi <= 0 || Wi <= 0 || Ho <= 0 || Wo <= 0) return USVM_ERR_ARG;
  usvm_launch(resize_bilinear_kernel, dim3(grid_for(planes * Ho * Wo)), dim3(256), 0, STREAM, x, y, planes, Hi, Wi, Ho, Wo, post_mode,
                                                                         post_scale, post_bias);
  return usvm_check_launch();
}

extern "C" int usvm_non_overlap_f32(const float* x, float* y, int B, long long HW, int group, int post_mode,
                                    float post_scale, float post_bias, void* stream) {
  if (group <= 0) group = B;
  if (!x || !y || B <= 0 || HW <= 0 || B % group) return USVM_ERR_ARG;
  usvm_launch(non_overlap_kernel, dim3(grid_for((long long)(B / group) * HW)), dim3(256), 0, STREAM, x, y, B, HW, group,
              post_mode, post_scale, post_bias);
  return usvm_check_launch();
}

extern "C" int usvm_resize_bilinear_aa(const float* x, float* y, long long planes, int Hi, int Wi, int Ho, int Wo,
                                       int binarize_half, void* stream) {
  if (!x || !y || planes <= 0 || Hi <= 0 || Wi <= 0 || Ho <= 0 || Wo <= 0) return USVM_ERR_ARG;
  usvm_launch(resize_bilinear_aa_kernel, dim3(grid_for(planes * Ho * Wo)), dim3(256), 0, STREAM, x, y, planes, Hi, Wi, Ho, Wo,
                                                                            binarize_half);
  return usvm_check_launch();
}
