// Bandwidth kernels around the GEMM / attention tiles: normalisation, rotary encoding, window
// (un)partition with Hiera's zero-pad and q-pool semantics, pooling, FPN top-down add, im2col for the
// patch embedding, and memory-bank assembly.  All are coalesced over the channel (innermost) axis,
// vectorised where the channel count allows, and sized by rows x channels so one launch covers every
// frame / object in the batch.
#include "common.cuh"
#include "usvm2_b200.h"

namespace {

// ---------------------------------------------------------------------------------------------
// LayerNorm over the last axis (nn.LayerNorm, hieradet.py:101/124, memory_attention.py:42-44,
// transformer.py norms): one warp per row, two-pass mean / variance in fp32.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ w, const float* __restrict__ b,
                 float eps, int gelu, float* out_f32, int ldo_f32, bf16* out_bf16, int ldo_bf16, int rows, int C) {
  PDL_ENTRY();
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* xr = x + row * ldx;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += xr[c];
  const float mean = warp_sum(s) / C;
  float q = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float d = xr[c] - mean;
    q = fmaf(d, d, q);
  }
  const float rstd = 1.0f / sqrtf(warp_sum(q) / C + eps);
  for (int c = lane; c < C; c += 32) {
    float y = (xr[c] - mean) * rstd * w[c] + b[c];
    if (gelu) y = gelu_erf(y);
    if (out_f32) out_f32[row * ldo_f32 + c] = y;
    if (out_bf16) out_bf16[row * ldo_bf16 + c] = __float2bfloat16(y);
  }
}

// single-read variant: the row lives in registers (NV = C / 32 values per lane)
template <int NV>
__global__ void __launch_bounds__(256)
layernorm_reg_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ w, const float* __restrict__ b,
                     float eps, int gelu, float* out_f32, int ldo_f32, bf16* out_bf16, int ldo_bf16, int rows) {
  constexpr int C = NV * 32;
  const int lane = threadIdx.x & 31;
  // weight and bias are constants of the model: fetched before the programmatic-dependency wait, so their latency
  // overlaps the predecessor's tail instead of following the row statistics
  float gw[NV], gb[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    gw[j] = __ldg(w + j * 32 + lane);
    gb[j] = __ldg(b + j * 32 + lane);
  }
  PDL_ENTRY();
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* xr = x + row * ldx;
  float v[NV];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    v[j] = xr[j * 32 + lane];
    s += v[j];
  }
  const float mean = warp_sum(s) / C;
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    v[j] -= mean;
    q = fmaf(v[j], v[j], q);
  }
  const float rstd = 1.0f / sqrtf(warp_sum(q) / C + eps);
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int c = j * 32 + lane;
    float y = v[j] * rstd * gw[j] + gb[j];
    if (gelu) y = gelu_erf(y);
    if (out_f32) out_f32[row * ldo_f32 + c] = y;
    if (out_bf16) out_bf16[row * ldo_bf16 + c] = __float2bfloat16(y);
  }
}

// narrow rows (C = 96 / 192: Hiera stages 1-2, 16 K / 4 K rows per frame): LPR lanes per row, 12 channels per lane as three
// float4 -- 4 (2) rows per warp, reductions over 8 (16) lanes; a warp per 384-byte row was mostly shuffles and loop overhead
template <int LPR, int NV4 = 3>
__global__ void __launch_bounds__(256)
layernorm_narrow_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ w, const float* __restrict__ b,
                        float eps, int gelu, float* out_f32, int ldo_f32, bf16* out_bf16, int ldo_bf16, int rows) {
  constexpr int C = LPR * 4 * NV4, RPW = 32 / LPR;  // (LPR = 32: one row per warp, NV4 float4 per lane -- C = 384 / 768)
  const int lane = threadIdx.x & 31, sub = lane % LPR, rw = lane / LPR;
  float4 gw[NV4], gb[NV4];
#pragma unroll
  for (int j = 0; j < NV4; ++j) {
    gw[j] = __ldg(reinterpret_cast<const float4*>(w) + j * LPR + sub);
    gb[j] = __ldg(reinterpret_cast<const float4*>(b) + j * LPR + sub);
  }
  PDL_ENTRY();
  const long long row = ((long long)blockIdx.x * 8 + (threadIdx.x >> 5)) * RPW + rw;
  const bool ok = row < rows;  // (every lane stays for the shuffles)
  float4 v[NV4];
  const float4* xr = reinterpret_cast<const float4*>(x + (ok ? row : 0) * ldx);
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < NV4; ++j) {
    v[j] = xr[j * LPR + sub];
    s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
  }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / C;
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < NV4; ++j) {
    v[j].x -= mean; v[j].y -= mean; v[j].z -= mean; v[j].w -= mean;
    q = fmaf(v[j].x, v[j].x, q); q = fmaf(v[j].y, v[j].y, q); q = fmaf(v[j].z, v[j].z, q); q = fmaf(v[j].w, v[j].w, q);
  }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = 1.0f / sqrtf(q / C + eps);
  if (!ok) return;
#pragma unroll
  for (int j = 0; j < NV4; ++j) {
    float4 y;
    y.x = v[j].x * rstd * gw[j].x + gb[j].x; y.y = v[j].y * rstd * gw[j].y + gb[j].y;
    y.z = v[j].z * rstd * gw[j].z + gb[j].z; y.w = v[j].w * rstd * gw[j].w + gb[j].w;
    if (gelu) { y.x = gelu_erf(y.x); y.y = gelu_erf(y.y); y.z = gelu_erf(y.z); y.w = gelu_erf(y.w); }
    const int c = (j * LPR + sub) * 4;
    if (out_f32) *reinterpret_cast<float4*>(out_f32 + row * ldo_f32 + c) = y;
    if (out_bf16) {
      uint2 pk;
      pk.x = pack_bf16x2(y.x, y.y);
      pk.y = pack_bf16x2(y.z, y.w);
      *reinterpret_cast<uint2*>(out_bf16 + row * ldo_bf16 + c) = pk;
    }
  }
}

// the same for a channel count that is not a multiple of 32 (Hiera-B+ stage 1: 112 of 128): lanes past C hold nothing
template <int NV>
__global__ void __launch_bounds__(256)
layernorm_regp_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ w, const float* __restrict__ b,
                      float eps, int gelu, float* out_f32, int ldo_f32, bf16* out_bf16, int ldo_bf16, int rows, int C) {
  const int lane = threadIdx.x & 31;
  float gw[NV], gb[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const bool ok = j * 32 + lane < C;
    gw[j] = ok ? __ldg(w + j * 32 + lane) : 0.f;
    gb[j] = ok ? __ldg(b + j * 32 + lane) : 0.f;
  }
  PDL_ENTRY();
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* xr = x + row * ldx;
  float v[NV];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    v[j] = j * 32 + lane < C ? xr[j * 32 + lane] : 0.f;
    s += v[j];
  }
  const float mean = warp_sum(s) / C;
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    if (j * 32 + lane < C) {
      v[j] -= mean;
      q = fmaf(v[j], v[j], q);
    }
  }
  const float rstd = 1.0f / sqrtf(warp_sum(q) / C + eps);
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int c = j * 32 + lane;
    if (c < C) {
      float y = v[j] * rstd * gw[j] + gb[j];
      if (gelu) y = gelu_erf(y);
      if (out_f32) out_f32[row * ldo_f32 + c] = y;
      if (out_bf16) out_bf16[row * ldo_bf16 + c] = __float2bfloat16(y);
    }
  }
}

// out[r, c] = alpha * x[r % x_mod, c] + beta * y[r % y_mod, c]   (fp32 in; fp32 and/or bf16 out)
__global__ void axpby_rows_kernel(const float* __restrict__ x, const float* __restrict__ y, float alpha, float beta,
                                  int x_mod, int y_mod, int x_div, float* out_f32, bf16* out_bf16, long long rows, int C) {
  PDL_ENTRY();
  const long long total = rows * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / C;
    const int c = (int)(i - r * C);
    const long long xr = x_div > 0 ? (r / x_div) * x_mod + r % x_mod : (x_mod > 0 ? r % x_mod : r);
    float v = alpha * x[xr * C + c];
    if (y) v += beta * y[(y_mod > 0 ? r % y_mod : r) * C + c];
    if (out_f32) out_f32[i] = v;
    if (out_bf16) out_bf16[i] = __float2bfloat16(v);
  }
}

__global__ void cast_f32_bf16_kernel(const float* __restrict__ x, bf16* __restrict__ y, long long n) {
  PDL_ENTRY();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = __float2bfloat16(x[i]);
}

// ---------------------------------------------------------------------------------------------
// Axial RoPE (position_encoding.py:174-221): rotate adjacent channel pairs of fp32 rows by the
// (cos, sin) of token position (row_in_batch % table_rows); rows >= n_rope inside each batch of
// rows_per_batch are copied unrotated (object-pointer tokens).  Output bf16 (attention operand).
// ---------------------------------------------------------------------------------------------
__global__ void rope_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ cs,
                            const float* __restrict__ sn, bf16* __restrict__ out, int ldo, long long rows,
                            int rows_per_batch, int n_rope, int table_rows, int half) {
  PDL_ENTRY();
  const long long total = rows * half;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / half;
    const int j = (int)(i - r * half);
    const int rb = (int)(r % rows_per_batch);
    const float2 v = *reinterpret_cast<const float2*>(x + r * ldx + 2 * j);
    float a = v.x, b = v.y;
    if (rb < n_rope) {
      const int pos = rb % table_rows;
      const float c = cs[pos * half + j], s = sn[pos * half + j];
      a = v.x * c - v.y * s;
      b = v.x * s + v.y * c;
    }
    *reinterpret_cast<uint32_t*>(out + r * ldo + 2 * j) = pack_bf16x2(a, b);
  }
}

// ---------------------------------------------------------------------------------------------
// Hiera window partition (backbones/utils.py:17-37) fused with the QKV split and the optional 2x2
// q max-pool (hieradet.py:63-67).  Source: qkv bf16 [F, Hg, Wg, 3*C] in raster order.  Tokens that
// fall in the zero padding carry qkv = bias (the reference pads *after* norm1, so the padded rows
// enter the qkv Linear as zeros).  Destination: window-major Q [F*nW, nq, C], K/V [F*nW, nk, C].
// ---------------------------------------------------------------------------------------------
__global__ void window_gather_kernel(const bf16* __restrict__ qkv, const float* __restrict__ bias,
                                     bf16* __restrict__ Qw, bf16* __restrict__ Kw, bf16* __restrict__ Vw, int F,
                                     int Hg, int Wg, int ws, int pool, int C) {
  PDL_ENTRY();
  const int nwx = (Wg + ws - 1) / ws, nwy = (Hg + ws - 1) / ws;
  const int nk = ws * ws;
  const int wq = pool ? ws / 2 : ws;
  const int nq = wq * wq;
  const int C8 = C / 8;
  const long long per_win = (long long)(nq + 2 * nk) * C8;
  const long long total = (long long)F * nwy * nwx * per_win;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long win = i / per_win;
    long long rem = i - win * per_win;
    const int f = (int)(win / (nwy * nwx));
    const int wi = (int)(win - (long long)f * nwy * nwx);
    const int wy = wi / nwx, wx = wi - wy * nwx;
    const int c8 = (int)(rem % C8);
    int tok = (int)(rem / C8);
    int which;  // 0 q, 1 k, 2 v
    if (tok < nq) which = 0;
    else if (tok < nq + nk) { which = 1; tok -= nq; }
    else { which = 2; tok -= nq + nk; }
    auto fetch = [&](int ly, int lx, float (&o)[8]) {
      const int y = wy * ws + ly, x = wx * ws + lx;
      if (y < Hg && x < Wg) {
        const uint4 u = *reinterpret_cast<const uint4*>(qkv + (((long long)f * Hg + y) * Wg + x) * 3 * C + which * C + c8 * 8);
        const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
        o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y; o[4] = c.x; o[5] = c.y; o[6] = d.x; o[7] = d.y;
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = __bfloat162float(__float2bfloat16(bias[which * C + c8 * 8 + e]));
      }
    };
    float o[8];
    if (which == 0 && pool) {
      const int qy = tok / wq, qx = tok - qy * wq;
      float t[8];
      fetch(2 * qy, 2 * qx, o);
      fetch(2 * qy, 2 * qx + 1, t);
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = fmaxf(o[e], t[e]);
      fetch(2 * qy + 1, 2 * qx, t);
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = fmaxf(o[e], t[e]);
      fetch(2 * qy + 1, 2 * qx + 1, t);
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = fmaxf(o[e], t[e]);
    } else {
      const int w_ = which == 0 ? wq : ws;
      const int ly = tok / w_, lx = tok - ly * w_;
      fetch(ly, lx, o);
    }
    uint4 pk;
    pk.x = pack_bf16x2(o[0], o[1]); pk.y = pack_bf16x2(o[2], o[3]);
    pk.z = pack_bf16x2(o[4], o[5]); pk.w = pack_bf16x2(o[6], o[7]);
    bf16* dst = which == 0 ? Qw + (win * nq + tok) * C : (which == 1 ? Kw : Vw) + (win * nk + tok) * C;
    *reinterpret_cast<uint4*>(dst + c8 * 8) = pk;
  }
}

// window_unpartition (backbones/utils.py:40-61): window-major [F*nW, wq*wq, C] -> raster [F, Ho, Wo, C]
__global__ void window_scatter_kernel(const bf16* __restrict__ Ow, bf16* __restrict__ out, int F, int Ho, int Wo,
                                      int wq, int C) {
  PDL_ENTRY();
  const int nwx = (Wo + wq - 1) / wq, nwy = (Ho + wq - 1) / wq;
  const int C8 = C / 8;
  const long long total = (long long)F * Ho * Wo * C8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % C8);
    const long long tokid = i / C8;
    const int x = (int)(tokid % Wo);
    const int y = (int)((tokid / Wo) % Ho);
    const int f = (int)(tokid / ((long long)Wo * Ho));
    const int wy = y / wq, wx = x / wq;
    const long long win = ((long long)f * nwy + wy) * nwx + wx;
    const int tok = (y - wy * wq) * wq + (x - wx * wq);
    *reinterpret_cast<uint4*>(out + tokid * C + c8 * 8) =
        *reinterpret_cast<const uint4*>(Ow + (win * wq * wq + tok) * C + c8 * 8);
  }
}

// 2x2 / stride-2 max pool on NHWC fp32 (do_pool on the projected shortcut, hieradet.py:139-140)
__global__ void maxpool2_nhwc_kernel(const float* __restrict__ x, float* __restrict__ y, int F, int H, int W, int C) {
  PDL_ENTRY();
  const int Ho = H / 2, Wo = W / 2, C4 = C / 4;
  const long long total = (long long)F * Ho * Wo * C4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % C4);
    const long long t = i / C4;
    const int ox = (int)(t % Wo), oy = (int)((t / Wo) % Ho), f = (int)(t / ((long long)Wo * Ho));
    const float* p = x + (((long long)f * H + 2 * oy) * W + 2 * ox) * C + c4 * 4;
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + C);
    const float4 c = *reinterpret_cast<const float4*>(p + (long long)W * C), d = *reinterpret_cast<const float4*>(p + (long long)W * C + C);
    float4 m;
    m.x = fmaxf(fmaxf(a.x, b.x), fmaxf(c.x, d.x));
    m.y = fmaxf(fmaxf(a.y, b.y), fmaxf(c.y, d.y));
    m.z = fmaxf(fmaxf(a.z, b.z), fmaxf(c.z, d.z));
    m.w = fmaxf(fmaxf(a.w, b.w), fmaxf(c.w, d.w));
    *reinterpret_cast<float4*>(y + t * C + c4 * 4) = m;
  }
}

// FPN top-down: fine[f,y,x,:] += coarse[f,y/2,x/2,:]  (nearest x2, image_encoder.py:116-126)
__global__ void upsample2_add_kernel(float* __restrict__ fine, const float* __restrict__ coarse, bf16* fine_bf16,
                                     int F, int H, int W, int C) {
  PDL_ENTRY();
  const long long total = (long long)F * H * W * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long t = i / C;
    const int x = (int)(t % W), y = (int)((t / W) % H), f = (int)(t / ((long long)W * H));
    const float v = fine[i] + coarse[(((long long)f * (H / 2) + y / 2) * (W / 2) + x / 2) * C + c];
    fine[i] = v;
    if (fine_bf16) fine_bf16[i] = __float2bfloat16(v);
  }
}

// ---------------------------------------------------------------------------------------------
// im2col of the 7x7 / stride 4 / pad 3 patch embedding (backbones/utils.py:64-94):
// img fp32 NCHW [F,3,S,S] -> A bf16 [F*(S/4)^2, KP], column k = c*49 + ky*7 + kx, zero beyond 147
// ---------------------------------------------------------------------------------------------
// one thread per 8 consecutive columns of a token row: eight scattered (L1-resident) 4-byte reads, one 16-byte store
__global__ void im2col_patch_kernel(const float* __restrict__ img, bf16* __restrict__ A, int F, int S, int KP) {
  PDL_ENTRY();
  const int G = S / 4, KP8 = KP >> 3;
  const long long total = (long long)F * G * G * KP8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int k0 = (int)(i % KP8) << 3;
    const long long t = i / KP8;
    const int ox = (int)(t % G), oy = (int)((t / G) % G), f = (int)(t / ((long long)G * G));
    const float* base = img + (long long)f * 3 * S * S;
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int k = k0 + u;
      v[u] = 0.f;
      if (k < 147) {
        const int c = k / 49, r = k - c * 49, ky = r / 7, kx = r - ky * 7;
        const int y = oy * 4 - 3 + ky, x = ox * 4 - 3 + kx;
        if (y >= 0 && y < S && x >= 0 && x < S) v[u] = __ldg(base + ((long long)c * S + y) * S + x);
      }
    }
    uint4 pk;
    pk.x = pack_bf16x2(v[0], v[1]); pk.y = pack_bf16x2(v[2], v[3]);
    pk.z = pack_bf16x2(v[4], v[5]); pk.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(A + t * KP + k0) = pk;
  }
}

// Non-overlapping P x P patches (ViT PatchEmbed, kernel = stride = P, no padding; efficient_track_anything/modeling/
// backbones/utils.py:64-94 with vitdet.py:214-220): img fp32 [F,3,S,S] -> bf16 [F*(S/P)^2, 3*P*P], k = (c, ky, kx).
// One thread moves 8 consecutive kx: a 32-byte read from an image row, a 16-byte write.
__global__ void im2col_patch_grid_kernel(const float* __restrict__ img, bf16* __restrict__ A, int F, int S, int P) {
  PDL_ENTRY();
  const int G = S / P, K = 3 * P * P, K8 = K >> 3, P8 = P >> 3;
  const long long total = (long long)F * G * G * K8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int k8 = (int)(i % K8);
    const long long t = i / K8;
    const int ox = (int)(t % G), oy = (int)((t / G) % G), f = (int)(t / ((long long)G * G));
    const int kx0 = (k8 % P8) << 3, ky = (k8 / P8) % P, c = k8 / (P8 * P);
    const float4* src = reinterpret_cast<const float4*>(img + (((long long)f * 3 + c) * S + (oy * P + ky)) * S + ox * P + kx0);
    const float4 a = __ldg(src), b = __ldg(src + 1);
    uint4 pk;
    pk.x = pack_bf16x2(a.x, a.y); pk.y = pack_bf16x2(a.z, a.w);
    pk.z = pack_bf16x2(b.x, b.y); pk.w = pack_bf16x2(b.z, b.w);
    *reinterpret_cast<uint4*>(A + t * K + (k8 << 3)) = pk;
  }
}

// uint8 grayscale [F,S,S] -> normalised fp32 [F,3,S,S] ((g/255 - mean_c)/std_c, misc.py:253-276)
__global__ void normalize_gray_kernel(const uint8_t* __restrict__ g, float* __restrict__ out, long long frames_px,
                                      long long px, float m0, float m1, float m2, float s0, float s1, float s2) {
  PDL_ENTRY();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < frames_px; i += (long long)gridDim.x * blockDim.x) {
    const long long f = i / px, o = i - f * px;
    const float v = (float)g[i] / 255.0f;
    out[(f * 3 + 0) * px + o] = (v - m0) / s0;
    out[(f * 3 + 1) * px + o] = (v - m1) / s1;
    out[(f * 3 + 2) * px + o] = (v - m2) / s2;
  }
}

// uint8 RGB frames as PIL decodes them, [F,S,S,3] interleaved -> normalised fp32 planes [F,3,S,S] (misc.py:92-101,
// 268-276: /255, - mean, / std): one thread per pixel reads 3 bytes, writes one float to each plane (coalesced per plane)
__global__ void normalize_rgb_kernel(const uint8_t* __restrict__ rgb, float* __restrict__ out, long long frames_px,
                                     long long px, float m0, float m1, float m2, float s0, float s1, float s2) {
  PDL_ENTRY();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < frames_px; i += (long long)gridDim.x * blockDim.x) {
    const long long f = i / px, o = i - f * px;
    const uint8_t* p = rgb + i * 3;
    out[(f * 3 + 0) * px + o] = ((float)p[0] / 255.0f - m0) / s0;
    out[(f * 3 + 1) * px + o] = ((float)p[1] / 255.0f - m1) / s1;
    out[(f * 3 + 2) * px + o] = ((float)p[2] / 255.0f - m2) / s2;
  }
}

// ---------------------------------------------------------------------------------------------
// Memory-bank assembly (sam2_base.py:1344-1437): for every selected memory frame f (token-major bf16
// [B, T, 64]) write  k_in = mem + pos + tpos[f]  and  v_in = mem  into the concatenated key/value
// inputs [B, Nk, 64]; object-pointer tokens (fp32 [B, P, 64], pos [P, 64]) follow at row `ptr_row0`.
// ---------------------------------------------------------------------------------------------
__global__ void build_memory_kernel(const usvm_memory_frames fr, const float* __restrict__ pos,
                                    const float* __restrict__ tpos, const float* __restrict__ ptrs,
                                    const float* __restrict__ ptr_pos, bf16* __restrict__ k_in, bf16* __restrict__ v_in,
                                    int B, int T, int Cm, int n_ptr_tokens, int Nk, int Nk_total, int row_offset) {
  PDL_ENTRY();
  const int ptr_row0 = fr.count * T;
  const long long total = (long long)B * Nk * Cm;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % Cm);
    const long long t = i / Cm;
    const int row = (int)(t % Nk);
    const int b = (int)(t / Nk);
    float m, pe;
    if (row < ptr_row0) {
      const int f = row / T, tok = row - f * T;
      m = __bfloat162float(reinterpret_cast<const bf16*>(fr.mem[f])[((long long)b * T + tok) * Cm + c]);
      pe = pos[tok * Cm + c] + tpos[fr.tpos_index[f] * Cm + c];
    } else {
      const int pr = row - ptr_row0;
      m = ptrs[((long long)b * n_ptr_tokens + pr) * Cm + c];
      pe = ptr_pos[pr * Cm + c];
    }
    const long long o = ((long long)b * Nk_total + row_offset + row) * Cm + c;
    k_in[o] = __float2bfloat16(m + pe);
    v_in[o] = __float2bfloat16(m);
  }
}

// memory feature epilogue (sam2_base.py:1488-1496, predictor :956): add no_obj_embed_spatial where the object
// score is <= 0, round to bf16 into the memory-bank slot
__global__ void finalize_memory_kernel(const float* __restrict__ x, const float* __restrict__ score, int score_stride,
                                       const float* __restrict__ no_obj_embed, bf16* __restrict__ mem, int B, int T,
                                       int Cm, const usvm_frame_ctrl* __restrict__ ctrl) {
  PDL_ENTRY();
  const long long total = (long long)B * T * Cm;
  if (!mem) mem = reinterpret_cast<bf16*>(ctrl->mem_store) + (long long)ctrl->cur_frame * ctrl->mem_slot_stride;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % Cm);
    const int b = (int)(i / ((long long)T * Cm));
    float v = x[i];
    if (!(score[(long long)b * score_stride] > 0.f)) v += no_obj_embed[c];
    mem[i] = __float2bfloat16(v);
  }
}


// ---------------------------------------------------------------------------------------------
// Device-indexed variants for the CUDA-graph steady-state frame: which stored frames feed the memory
// bank is read from a small control block in device memory, so one captured graph serves every frame.
// ---------------------------------------------------------------------------------------------
// temporal position encoding of the object pointers (get_1d_sine_pe + obj_ptr_tpos_proj,
// sam2_utils.py:64-74, sam2_base.py:1402-1408): out[p*4 + q, c] = W[c, :] . [sin(rel/dim_t), cos(rel/dim_t)] + b[c]
// one CTA per pointer; warp w owns output channels 8w .. 8w+7: the lanes stride the 256-long reduction with float4 loads
// (coalesced, all 16 in flight at once) and the channel sums are warp reductions
__global__ void __launch_bounds__(256)
ptr_tpos_kernel(const usvm_frame_ctrl* __restrict__ ctrl, const float* __restrict__ W, const float* __restrict__ bias,
                float* __restrict__ out) {
  PDL_ENTRY();
  __shared__ __align__(16) float pe[256];
  const int p = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float4 wv[8][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float* wr = W + (warp * 8 + i) * 256 + lane * 4;
    wv[i][0] = __ldg(reinterpret_cast<const float4*>(wr));
    wv[i][1] = __ldg(reinterpret_cast<const float4*>(wr + 128));
  }
  const float rel = ctrl->ptr_rel[p];
  if (tid < 128) {
    const float dim_t = powf(10000.0f, (float)(2 * (tid / 2)) / 128.0f);
    const float e = rel / dim_t;
    pe[tid] = sinf(e);
    pe[128 + tid] = cosf(e);
  }
  __syncthreads();
  const float4 p0 = *reinterpret_cast<const float4*>(pe + lane * 4);
  const float4 p1 = *reinterpret_cast<const float4*>(pe + 128 + lane * 4);
  float mine = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    // same summation tree for every channel: fixed order inside the lane, then the warp butterfly
    float a = wv[i][0].x * p0.x;
    a = fmaf(wv[i][0].y, p0.y, a); a = fmaf(wv[i][0].z, p0.z, a); a = fmaf(wv[i][0].w, p0.w, a);
    a = fmaf(wv[i][1].x, p1.x, a); a = fmaf(wv[i][1].y, p1.y, a); a = fmaf(wv[i][1].z, p1.z, a); a = fmaf(wv[i][1].w, p1.w, a);
    a = warp_sum(a);
    if (lane == i) mine = a;
  }
  if (lane < 8) {
    const int c = warp * 8 + lane;
    const float v = mine + bias[c];
#pragma unroll
    for (int q = 0; q < 4; ++q) out[(p * 4 + q) * 64 + c] = v;
  }
}

// 8 channels per thread: one 16-byte bf16 load of the stored memory (or two float4 of a pointer token), four float4 of
// position tables, two 16-byte stores (the scalar version moved 2 bytes per access: 89 us for 32 objects, 90 MB)
__global__ void build_memory_store_kernel(const usvm_frame_ctrl* __restrict__ ctrl, const float* __restrict__ pos,
                                          const float* __restrict__ tpos, const float* __restrict__ ptr_pos,
                                          bf16* __restrict__ k_in, bf16* __restrict__ v_in, int B, int T, int Cm,
                                          int n_mem, int n_ptr) {
  PDL_ENTRY();
  const bf16* __restrict__ mem_store = reinterpret_cast<const bf16*>(ctrl->mem_store);
  const float* __restrict__ ptr_store = ctrl->ptr_store;
  const long long mem_frame_stride = ctrl->mem_slot_stride, ptr_frame_stride = ctrl->ptr_slot_stride;
  const int ptr_row0 = n_mem * T;
  const int Nk = ptr_row0 + n_ptr * 4;
  const int C8 = Cm >> 3;
  const long long total = (long long)B * Nk * C8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C8) << 3;
    const long long t = i / C8;
    const int row = (int)(t % Nk);
    const int b = (int)(t / Nk);
    float m[8], pe[8];
    if (row < ptr_row0) {
      const int f = row / T, tok = row - f * T;
      const uint4 u = *reinterpret_cast<const uint4*>(mem_store + (long long)ctrl->mem_frame[f] * mem_frame_stride +
                                                      ((long long)b * T + tok) * Cm + c);
      const float2 a0 = unpack_bf16x2(u.x), a1 = unpack_bf16x2(u.y), a2 = unpack_bf16x2(u.z), a3 = unpack_bf16x2(u.w);
      m[0] = a0.x; m[1] = a0.y; m[2] = a1.x; m[3] = a1.y; m[4] = a2.x; m[5] = a2.y; m[6] = a3.x; m[7] = a3.y;
      const float4* pp = reinterpret_cast<const float4*>(pos + tok * Cm + c);
      const float4* tp = reinterpret_cast<const float4*>(tpos + ctrl->mem_tpos[f] * Cm + c);
      const float4 p0 = __ldg(pp), p1 = __ldg(pp + 1), t0 = __ldg(tp), t1 = __ldg(tp + 1);
      pe[0] = p0.x + t0.x; pe[1] = p0.y + t0.y; pe[2] = p0.z + t0.z; pe[3] = p0.w + t0.w;
      pe[4] = p1.x + t1.x; pe[5] = p1.y + t1.y; pe[6] = p1.z + t1.z; pe[7] = p1.w + t1.w;
    } else {
      const int pr = row - ptr_row0;  // pointer p = pr / 4 contributes channels [(pr % 4) * Cm, +Cm)
      const float4* mp = reinterpret_cast<const float4*>(ptr_store + (long long)ctrl->ptr_frame[pr >> 2] * ptr_frame_stride +
                                                         (long long)b * (4 * Cm) + (pr & 3) * Cm + c);
      const float4* pp = reinterpret_cast<const float4*>(ptr_pos + pr * Cm + c);
      const float4 m0 = mp[0], m1 = mp[1], p0 = pp[0], p1 = pp[1];
      m[0] = m0.x; m[1] = m0.y; m[2] = m0.z; m[3] = m0.w; m[4] = m1.x; m[5] = m1.y; m[6] = m1.z; m[7] = m1.w;
      pe[0] = p0.x; pe[1] = p0.y; pe[2] = p0.z; pe[3] = p0.w; pe[4] = p1.x; pe[5] = p1.y; pe[6] = p1.z; pe[7] = p1.w;
    }
    uint4 ko, vo;
    ko.x = pack_bf16x2(m[0] + pe[0], m[1] + pe[1]); ko.y = pack_bf16x2(m[2] + pe[2], m[3] + pe[3]);
    ko.z = pack_bf16x2(m[4] + pe[4], m[5] + pe[5]); ko.w = pack_bf16x2(m[6] + pe[6], m[7] + pe[7]);
    vo.x = pack_bf16x2(m[0], m[1]); vo.y = pack_bf16x2(m[2], m[3]);
    vo.z = pack_bf16x2(m[4], m[5]); vo.w = pack_bf16x2(m[6], m[7]);
    *reinterpret_cast<uint4*>(k_in + i * 8) = ko;
    *reinterpret_cast<uint4*>(v_in + i * 8) = vo;
  }
}

__global__ void set_frame_ctrl_kernel(usvm_frame_ctrl* dst, const usvm_frame_ctrl v) {
  PDL_ENTRY();
  if (threadIdx.x == 0 && blockIdx.x == 0) *dst = v;
}

// Everything a replay of the tracked-frame graph needs refreshed, in ONE launch: the control block (by-value parameter)
// and the frame's backbone features copied into the graph's static input buffers (up to 4 segments, 16-byte units).
struct FrameSegments {
  const uint4* src[4];
  uint4* dst[4];
  long long n16[4];  // 16-byte units per segment
};
__global__ void frame_prologue_kernel(usvm_frame_ctrl* dst, const usvm_frame_ctrl v, const FrameSegments seg) {
  PDL_ENTRY();
  if (threadIdx.x == 0 && blockIdx.x == 0) *dst = v;
  const long long stride = (long long)gridDim.x * blockDim.x;
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    const uint4* __restrict__ a = seg.src[s];
    uint4* __restrict__ b = seg.dst[s];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < seg.n16[s]; i += stride) b[i] = __ldg(a + i);
  }
}

// slot ctrl->cur_frame of the pointer / score / mask stores <- this frame's outputs
__global__ void store_outputs_kernel(const usvm_frame_ctrl* __restrict__ ctrl, const float* __restrict__ obj_ptr,
                                     const float* __restrict__ score, int score_stride, const float* __restrict__ masks,
                                     int B, int ptr_dim, int hw) {
  PDL_ENTRY();
  const long long slot = ctrl->cur_frame;
  const long long n0 = (long long)B * ptr_dim, n1 = B, n2 = (long long)B * hw;
  float* d0 = ctrl->ptr_store + slot * ctrl->ptr_slot_stride;
  float* d1 = ctrl->score_store + slot * ctrl->score_slot_stride;
  float* d2 = ctrl->mask_store + slot * ctrl->mask_slot_stride;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n0 + n1 + n2; i += (long long)gridDim.x * blockDim.x) {
    if (i < n0) d0[i] = obj_ptr[i];
    else if (i < n0 + n1) d1[i - n0] = score[(i - n0) * score_stride];
    else d2[i - n0 - n1] = masks[i - n0 - n1];
  }
}

inline int grid_for(long long total, int threads = 256) {
  long long b = (total + threads - 1) / threads;
  const long long cap = 148LL * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

#define STREAM reinterpret_cast<cudaStream_t>(stream)

extern "C" int usvm_layernorm(const float* x, int ldx, const float* w, const float* b, float eps, int gelu,
                              float* out_f32, int ldo_f32, void* out_bf16, int ldo_bf16, int rows, int C,
                              void* stream) {
  if (!x || !w || !b || rows <= 0 || C <= 0 || (!out_f32 && !out_bf16)) return USVM_ERR_ARG;
  bf16* ob = reinterpret_cast<bf16*>(out_bf16);
#define LN_CASE(NV)                                                                                              \
  case NV * 32:                                                                                                  \
    usvm_launch(layernorm_reg_kernel<NV>, dim3(cdiv(rows, 8)), dim3(256), 0, STREAM, x, ldx, w, b, eps, gelu, out_f32, ldo_f32, ob,   \
                                                                ldo_bf16, rows);                                 \
    break;
  const bool vec_ok = !(ldx & 3) && !(reinterpret_cast<uintptr_t>(x) & 15) && !(reinterpret_cast<uintptr_t>(w) & 15) &&
                      !(reinterpret_cast<uintptr_t>(b) & 15) &&
                      (!out_f32 || (!(ldo_f32 & 3) && !(reinterpret_cast<uintptr_t>(out_f32) & 15))) &&
                      (!ob || (!(ldo_bf16 & 3) && !(reinterpret_cast<uintptr_t>(ob) & 7)));
  if (vec_ok && (C == 224 || C == 448 || C == 896)) {  // Hiera-B+ stages 2-4: 7 float4 per lane
    if (C == 224)
      usvm_launch(layernorm_narrow_kernel<8, 7>, dim3(cdiv(rows, 32)), dim3(256), 0, STREAM, x, ldx, w, b, eps, gelu,
                  out_f32, ldo_f32, ob, ldo_bf16, rows);
    else if (C == 448)
      usvm_launch(layernorm_narrow_kernel<16, 7>, dim3(cdiv(rows, 16)), dim3(256), 0, STREAM, x, ldx, w, b, eps, gelu,
                  out_f32, ldo_f32, ob, ldo_bf16, rows);
    else
      usvm_launch(layernorm_narrow_kernel<32, 7>, dim3(cdiv(rows, 8)), dim3(256), 0, STREAM, x, ldx, w, b, eps, gelu,
                  out_f32, ldo_f32, ob, ldo_bf16, rows);
    return usvm_check_launch();
  }
  if (vec_ok && (C == 96 || C == 192 || C == 384 || C == 768)) {  // the Hiera-tiny / small stage widths
    if (C == 96)
      usvm_launch(layernorm_narrow_kernel<8>, dim3(cdiv(rows, 32)), dim3(256), 0, STREAM, x, ldx, w, b, eps, gelu, out_f32,
                  ldo_f32, ob, ldo_bf16, rows);
    else if (C == 192)
      usvm_launch(layernorm_narrow_kernel<16>, dim3(cdiv(rows, 16)), dim3(256), 0, STREAM, x, ldx, w, b, eps, gelu, out_f32,
                  ldo_f32, ob, ldo_bf16, rows);
    else if (C == 384)
      usvm_launch(layernorm_narrow_kernel<32, 3>, dim3(cdiv(rows, 8)), dim3(256), 0, STREAM, x, ldx, w, b, eps, gelu,
                  out_f32, ldo_f32, ob, ldo_bf16, rows);
    else
      usvm_launch(layernorm_narrow_kernel<32, 6>, dim3(cdiv(rows, 8)), dim3(256), 0, STREAM, x, ldx, w, b, eps, gelu,
                  out_f32, ldo_f32, ob, ldo_bf16, rows);
    return usvm_check_launch();
  }
  switch (C) {
    LN_CASE(2) LN_CASE(3) LN_CASE(6) LN_CASE(7) LN_CASE(8) LN_CASE(12) LN_CASE(14) LN_CASE(24) LN_CASE(28)
    default:
      if (C < 128)
        usvm_launch(layernorm_regp_kernel<4>, dim3(cdiv(rows, 8)), dim3(256), 0, STREAM, x, ldx, w, b, eps, gelu, out_f32,
                    ldo_f32, ob, ldo_bf16, rows, C);
      else
      usvm_launch(layernorm_kernel, dim3(cdiv(rows, 8)), dim3(256), 0, STREAM, x, ldx, w, b, eps, gelu, out_f32, ldo_f32, ob, ldo_bf16, rows, C);
  }
#undef LN_CASE
  return usvm_check_launch();
}

extern "C" int usvm_axpby_rows(const float* x, const float* y, float alpha, float beta, int x_mod, int y_mod, int x_div,
                               float* out_f32, void* out_bf16, long long rows, int C, void* stream) {
  if (!x || rows <= 0 || C <= 0 || (x_div > 0 && x_mod <= 0)) return USVM_ERR_ARG;
  usvm_launch(axpby_rows_kernel, dim3(grid_for(rows * C)), dim3(256), 0, STREAM, x, y, alpha, beta, x_mod, y_mod, x_div, out_f32,
                                                            reinterpret_cast<bf16*>(out_bf16), rows, C);
  return usvm_check_launch();
}

extern "C" int usvm_cast_f32_bf16(const float* x, void* y, long long n, void* stream) {
  if (!x || !y || n <= 0) return USVM_ERR_ARG;
  usvm_launch(cast_f32_bf16_kernel, dim3(grid_for(n)), dim3(256), 0, STREAM, x, reinterpret_cast<bf16*>(y), n);
  return usvm_check_launch();
}

extern "C" int usvm_rope_bf16(const float* x, int ldx, const float* cos_t, const float* sin_t, void* out, int ldo,
                              long long rows, int rows_per_batch, int n_rope, int table_rows, int dim,
                              void* stream) {
  if (!x || !cos_t || !sin_t || !out || rows <= 0 || (dim & 1) || (ldx & 1) || (ldo & 1)) return USVM_ERR_ARG;
  usvm_launch(rope_kernel, dim3(grid_for(rows * (dim / 2))), dim3(256), 0, STREAM, x, ldx, cos_t, sin_t, reinterpret_cast<bf16*>(out), ldo,
                                                               rows, rows_per_batch, n_rope, table_rows, dim / 2);
  return usvm_check_launch();
}

extern "C" int usvm_window_gather(const void* qkv, const float* qkv_bias, void* Qw, void* Kw, void* Vw, int F, int Hg,
                                  int Wg, int ws, int pool, int C, void* stream) {
  if (!qkv || !qkv_bias || !Qw || !Kw || !Vw || ws <= 0 || (C % 8) || (pool && (ws & 1))) return USVM_ERR_ARG;
  const int nw = cdiv(Hg, ws) * cdiv(Wg, ws);
  const int nq = pool ? (ws / 2) * (ws / 2) : ws * ws;
  const long long total = (long long)F * nw * (nq + 2 * ws * ws) * (C / 8);
  usvm_launch(window_gather_kernel, dim3(grid_for(total)), dim3(256), 0, STREAM, 
      reinterpret_cast<const bf16*>(qkv), qkv_bias, reinterpret_cast<bf16*>(Qw), reinterpret_cast<bf16*>(Kw),
      reinterpret_cast<bf16*>(Vw), F, Hg, Wg, ws, pool, C);
  return usvm_check_launch();
}

extern "C" int usvm_window_scatter(const void* Ow, void* out, int F, int Ho, int Wo, int wq, int C, void* stream) {
  if (!Ow || !out || wq <= 0 || (C % 8)) return USVM_ERR_ARG;
  usvm_launch(window_scatter_kernel, dim3(grid_for((long long)F * Ho * Wo * (C / 8))), dim3(256), 0, STREAM, 
      reinterpret_cast<const bf16*>(Ow), reinterpret_cast<bf16*>(out), F, Ho, Wo, wq, C);
  return usvm_check_launch();
}

extern "C" int usvm_maxpool2_nhwc(const float* x, float* y, int F, int H, int W, int C, void* stream) {
  if (!x || !y || (H & 1) || (W & 1) || (C % 4)) return USVM_ERR_ARG;
  usvm_launch(maxpool2_nhwc_kernel, dim3(grid_for((long long)F * (H / 2) * (W / 2) * (C / 4))), dim3(256), 0, STREAM, x, y, F, H, W, C);
  return usvm_check_launch();
}

extern "C" int usvm_upsample2_add(float* fine, const float* coarse, void* fine_bf16, int F, int H, int W, int C,
                                  void* stream) {
  if (!fine || !coarse || (H & 1) || (W & 1)) return USVM_ERR_ARG;
  usvm_launch(upsample2_add_kernel, dim3(grid_for((long long)F * H * W * C)), dim3(256), 0, STREAM, 
      fine, coarse, reinterpret_cast<bf16*>(fine_bf16), F, H, W, C);
  return usvm_check_launch();
}

extern "C" int usvm_im2col_patch(const float* img, void* A, int F, int S, int KP, void* stream) {
  if (!img || !A || (S % 4) || KP < 147 || (KP % 8) || (reinterpret_cast<uintptr_t>(A) & 15)) return USVM_ERR_ARG;
  usvm_launch(im2col_patch_kernel, dim3(grid_for((long long)F * (S / 4) * (S / 4) * (KP / 8))), dim3(256), 0, STREAM, 
      img, reinterpret_cast<bf16*>(A), F, S, KP);
  return usvm_check_launch();
}

extern "C" int usvm_im2col_patch_grid(const float* img, void* A, int F, int S, int P, void* stream) {
  if (!img || !A || F <= 0 || S <= 0 || P <= 0 || (P % 8) || (S % P) || (reinterpret_cast<uintptr_t>(img) & 15) ||
      (reinterpret_cast<uintptr_t>(A) & 15))
    return USVM_ERR_ARG;
  const long long total = (long long)F * (S / P) * (S / P) * (3 * P * P / 8);
  usvm_launch(im2col_patch_grid_kernel, dim3(grid_for(total)), dim3(256), 0, STREAM, img, reinterpret_cast<bf16*>(A), F, S, P);
  return usvm_check_launch();
}

extern "C" int usvm_normalize_gray_u8(const uint8_t* gray, float* out, int F, int H, int W, const float* mean3,
                                      const float* std3, void* stream) {
  if (!gray || !out || !mean3 || !std3 || F <= 0) return USVM_ERR_ARG;
  const long long px = (long long)H * W;
  usvm_launch(normalize_gray_kernel, dim3(grid_for(F * px)), dim3(256), 0, STREAM, gray, out, F * px, px, mean3[0], mean3[1], mean3[2],
                                                              std3[0], std3[1], std3[2]);
  return usvm_check_launch();
}

extern "C" int usvm_normalize_rgb_u8(const uint8_t* rgb, float* out, int F, int H, int W, const float* mean3,
                                     const float* std3, void* stream) {
  if (!rgb || !out || !mean3 || !std3 || F <= 0 || H <= 0 || W <= 0) return USVM_ERR_ARG;
  const long long px = (long long)H * W;
  usvm_launch(normalize_rgb_kernel, dim3(grid_for(F * px)), dim3(256), 0, STREAM, rgb, out, F * px, px, mean3[0], mean3[1],
              mean3[2], std3[0], std3[1], std3[2]);
  return usvm_check_launch();
}

extern "C" int usvm_build_memory(const usvm_memory_frames* frames, const float* pos, const float* tpos,
                                 const float* ptrs, const float* ptr_pos, void* k_in, void* v_in, int B, int T,
                                 int Cm, int n_ptr_tokens, int Nk_total, int row_offset, void* stream) {
  if (!frames || frames->count < 0 || frames->count > USVM_MAX_MEMORY_FRAMES || !pos || !tpos || !k_in || !v_in)
    return USVM_ERR_ARG;
  if (n_ptr_tokens > 0 && (!ptrs || !ptr_pos)) return USVM_ERR_ARG;
  const int Nk = frames->count * T + n_ptr_tokens;
  if (Nk <= 0 || row_offset < 0 || row_offset + Nk > Nk_total) return USVM_ERR_ARG;
  usvm_launch(build_memory_kernel, dim3(grid_for((long long)B * Nk * Cm)), dim3(256), 0, STREAM, 
      *frames, pos, tpos, ptrs, ptr_pos, reinterpret_cast<bf16*>(k_in), reinterpret_cast<bf16*>(v_in), B, T, Cm,
      n_ptr_tokens, Nk, Nk_total, row_offset);
  return usvm_check_launch();
}

extern "C" int usvm_finalize_memory(const float* x, const float* score, int score_stride, const float* no_obj_embed,
                                    void* mem_bf16, int B, int T, int Cm, const usvm_frame_ctrl* ctrl_dev,
                                    void* stream) {
  if (!x || !score || !no_obj_embed || (!mem_bf16 && !ctrl_dev) || B <= 0) return USVM_ERR_ARG;
  usvm_launch(finalize_memory_kernel, dim3(grid_for((long long)B * T * Cm)), dim3(256), 0, STREAM, 
      x, score, score_stride, no_obj_embed, reinterpret_cast<bf16*>(mem_bf16), B, T, Cm, ctrl_dev);
  return usvm_check_launch();
}

extern "C" int usvm_ptr_tpos(const usvm_frame_ctrl* ctrl_dev, const float* W, const float* bias, float* out, int n_ptr,
                             void* stream) {
  if (!ctrl_dev || !W || !bias || !out || n_ptr <= 0 || n_ptr > USVM_MAX_PTRS) return USVM_ERR_ARG;
  usvm_launch(ptr_tpos_kernel, dim3(n_ptr), dim3(256), 0, STREAM, ctrl_dev, W, bias, out);
  return usvm_check_launch();
}

extern "C" int usvm_build_memory_store(const usvm_frame_ctrl* ctrl_dev, const float* pos, const float* tpos,
                                       const float* ptr_pos, void* k_in, void* v_in, int B, int T, int Cm, int n_mem,
                                       int n_ptr, void* stream) {
  if (!ctrl_dev || !pos || !tpos || !k_in || !v_in || n_mem < 0 || n_mem > USVM_MAX_MEMORY_FRAMES || n_ptr < 0 ||
      n_ptr > USVM_MAX_PTRS || Cm != 64)
    return USVM_ERR_ARG;
  if (n_ptr > 0 && !ptr_pos) return USVM_ERR_ARG;
  const long long total = (long long)B * (n_mem * T + n_ptr * 4) * (Cm / 8);
  if (total <= 0) return USVM_ERR_ARG;
  usvm_launch(build_memory_store_kernel, dim3(grid_for(total)), dim3(256), 0, STREAM, 
      ctrl_dev, pos, tpos, ptr_pos, reinterpret_cast<bf16*>(k_in), reinterpret_cast<bf16*>(v_in), B, T, Cm, n_mem, n_ptr);
  return usvm_check_launch();
}

extern "C" int usvm_store_outputs(const usvm_frame_ctrl* ctrl_dev, const float* obj_ptr, const float* score,
                                  int score_stride, const float* masks, int B, int ptr_dim, int hw, void* stream) {
  if (!ctrl_dev || !obj_ptr || !score || !masks || B <= 0) return USVM_ERR_ARG;
  usvm_launch(store_outputs_kernel, dim3(grid_for((long long)B * (ptr_dim + 1 + hw))), dim3(256), 0, STREAM, ctrl_dev, obj_ptr, score,
                                                                                       score_stride, masks, B, ptr_dim, hw);
  return usvm_check_launch();
}

extern "C" int usvm_frame_prologue(usvm_frame_ctrl* ctrl_dev, const usvm_frame_ctrl* ctrl_host, const void* const* src,
                                   void* const* dst, const long long* bytes, int n_seg, void* stream) {
  if (!ctrl_dev || !ctrl_host || n_seg < 0 || n_seg > 4 || (n_seg > 0 && (!src || !dst || !bytes))) return USVM_ERR_ARG;
  FrameSegments seg{};
  long long total = 0;
  for (int i = 0; i < n_seg; ++i) {
    if (!src[i] || !dst[i] || bytes[i] < 0 || (bytes[i] & 15) || (reinterpret_cast<uintptr_t>(src[i]) & 15) ||
        (reinterpret_cast<uintptr_t>(dst[i]) & 15))
      return USVM_ERR_ARG;
    seg.src[i] = reinterpret_cast<const uint4*>(src[i]);
    seg.dst[i] = reinterpret_cast<uint4*>(dst[i]);
    seg.n16[i] = bytes[i] / 16;
    total += seg.n16[i];
  }
  const int grid = (int)max(1LL, min(148LL * 4, (total + 255) / 256));
  usvm_launch(frame_prologue_kernel, dim3(grid), dim3(256), 0, STREAM, ctrl_dev, *ctrl_host, seg);
  return usvm_check_launch();
}

extern "C" int usvm_set_frame_ctrl(usvm_frame_ctrl* ctrl_dev, const usvm_frame_ctrl* ctrl_host, void* stream) {
  if (!ctrl_dev || !ctrl_host) return USVM_ERR_ARG;
  usvm_launch(set_frame_ctrl_kernel, dim3(1), dim3(32), 0, STREAM, ctrl_dev, *ctrl_host);  // by-value kernel parameter: no H2D copy, no sync
  return usvm_check_launch();
}
