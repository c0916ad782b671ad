// Latency-oriented fp32 kernels for the token side of the SAM mask decoder's two-way transformer
// (sam/transformer.py:137-212): a handful of tokens (8 per object) against 1024 image tokens.
//
//   usvm_gemm_skinny_f32  y[inst][m][n] = act((x + x2)[m] . W[inst][n] + b) (+ residual), M <= 8 rows per CTA pass;
//                         one warp owns 4 output columns and streams their weight rows once (coalesced float4),
//                         so a [8 x 256] x [2048 x 256] layer spreads over 64 CTAs instead of one.  Also runs the
//                         stacked 3-layer heads (hyper-networks, IoU, object score, obj_ptr_proj).
//   usvm_attn_t2i_f32     token -> image attention: one CTA per (object, head); scores [Nt x 1024] in shared memory
//   usvm_attn_i2t_f32     image -> token attention: one thread per (image token, head), keys/values in shared memory
#include "common.cuh"
#include "usvm2_b200.h"

namespace {

constexpr int SK_ROWS = 8;       // rows per pass
constexpr int SK_COLS_WARP = 4;  // output columns per warp
constexpr int SK_WARPS = 8;

__global__ void __launch_bounds__(256)
gemm_skinny_kernel(const usvm_skinny_params p) {
  extern __shared__ __align__(16) float xs[];  // [rows][K]
  const int inst = blockIdx.y;
  const int m0 = blockIdx.z * SK_ROWS;
  const int rows = min(SK_ROWS, p.M - m0);
  const int K = p.K;
  // stage (x + x2) rows
  for (int i = threadIdx.x; i < rows * K; i += blockDim.x) {
    const int m = i / K, k = i - m * K;
    const long long sel = p.row_select ? (long long)p.row_select[m0 + m] * p.x_sel_stride : 0;
    float v = p.x[(long long)inst * p.x_is + (long long)(m0 + m) * p.x_rs + sel + k];
    if (p.x2) v += p.x2[(long long)inst * p.x2_is + (long long)(m0 + m) * p.x2_rs + k];
    xs[i] = v;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = (blockIdx.x * SK_WARPS + warp) * SK_COLS_WARP;
  if (n0 >= p.N) return;
  float acc[SK_COLS_WARP][SK_ROWS];
#pragma unroll
  for (int c = 0; c < SK_COLS_WARP; ++c)
#pragma unroll
    for (int m = 0; m < SK_ROWS; ++m) acc[c][m] = 0.f;
  const float* Wb = p.w + (long long)inst * p.w_is;
  for (int k = lane * 4; k < K; k += 128) {
    float4 wv[SK_COLS_WARP];
#pragma unroll
    for (int c = 0; c < SK_COLS_WARP; ++c) {
      const int n = min(n0 + c, p.N - 1);
      wv[c] = *reinterpret_cast<const float4*>(Wb + (long long)n * K + k);
    }
#pragma unroll
    for (int m = 0; m < SK_ROWS; ++m) {
      if (m < rows) {
        const float4 xv = *reinterpret_cast<const float4*>(xs + m * K + k);
#pragma unroll
        for (int c = 0; c < SK_COLS_WARP; ++c)
          acc[c][m] += xv.x * wv[c].x + xv.y * wv[c].y + xv.z * wv[c].z + xv.w * wv[c].w;
      }
    }
  }
#pragma unroll
  for (int c = 0; c < SK_COLS_WARP; ++c)
#pragma unroll
    for (int m = 0; m < SK_ROWS; ++m) acc[c][m] = warp_sum(acc[c][m]);
  // lane (c * 8 + m) writes element (m, n0 + c)
  const int c = lane >> 3, m = lane & 7;
  float v = 0.f;
#pragma unroll
  for (int cc = 0; cc < SK_COLS_WARP; ++cc)
#pragma unroll
    for (int mm = 0; mm < SK_ROWS; ++mm)
      if (cc == c && mm == m) v = acc[cc][mm];
  const int n = n0 + c;
  if (m < rows && n < p.N) {
    if (p.bias) v += p.bias[(long long)inst * p.b_is + n];
    if (p.act == USVM_ACT_RELU) v = fmaxf(v, 0.f);
    else if (p.act == USVM_ACT_GELU) v = gelu_erf(v);
    if (p.residual) v += p.residual[(long long)inst * p.r_is + (long long)(m0 + m) * p.r_rs + n];
    p.out[(long long)inst * p.o_is + (long long)(m0 + m) * p.o_rs + n] = v;
  }
}

// ---- token -> image attention --------------------------------------------------------------------
constexpr int T2I_DH = 16;
constexpr int T2I_MAX_NT = 16;

__global__ void __launch_bounds__(256)
attn_t2i_kernel(const float* __restrict__ q, int q_rs, const float* __restrict__ k, const float* __restrict__ v,
                int kv_rs, float* __restrict__ out, int o_rs, int H, int Nt, int Nk, float scale) {
  extern __shared__ float sm[];
  float* s_p = sm;                       // [Nt][Nk]
  float* s_q = sm + Nt * Nk;             // [Nt][16]
  float* s_red = s_q + Nt * T2I_DH;      // [16 slices][Nt][16]
  const int b = blockIdx.x / H, h = blockIdx.x - b * H;
  const int tid = threadIdx.x;
  const float* qb = q + (long long)b * Nt * q_rs + h * T2I_DH;
  const float* kb = k + (long long)b * Nk * kv_rs + h * T2I_DH;
  const float* vb = v + (long long)b * Nk * kv_rs + h * T2I_DH;
  for (int i = tid; i < Nt * T2I_DH; i += blockDim.x) s_q[i] = qb[(i / T2I_DH) * q_rs + (i % T2I_DH)] * scale;
  __syncthreads();
  // scores: thread per key
  for (int j = tid; j < Nk; j += blockDim.x) {
    float kr[T2I_DH];
    const float4* kp = reinterpret_cast<const float4*>(kb + (long long)j * kv_rs);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 t4 = kp[i];
      kr[4 * i] = t4.x; kr[4 * i + 1] = t4.y; kr[4 * i + 2] = t4.z; kr[4 * i + 3] = t4.w;
    }
    for (int t = 0; t < Nt; ++t) {
      float d = 0.f;
#pragma unroll
      for (int c = 0; c < T2I_DH; ++c) d = fmaf(s_q[t * T2I_DH + c], kr[c], d);
      s_p[t * Nk + j] = d;
    }
  }
  __syncthreads();
  // softmax per token: one warp per token row (looping if Nt > 8 warps)
  const int warp = tid >> 5, lane = tid & 31;
  for (int t = warp; t < Nt; t += 8) {
    float mx = -INFINITY;
    for (int j = lane; j < Nk; j += 32) mx = fmaxf(mx, s_p[t * Nk + j]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < Nk; j += 32) {
      const float e = __expf(s_p[t * Nk + j] - mx);
      s_p[t * Nk + j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    for (int j = lane; j < Nk; j += 32) s_p[t * Nk + j] *= inv;
  }
  __syncthreads();
  // P.V: 16 key slices x 16 channels; every thread accumulates all Nt tokens for its (slice, channel)
  const int slice = tid >> 4, c = tid & 15;
  float acc[T2I_MAX_NT];
#pragma unroll
  for (int t = 0; t < T2I_MAX_NT; ++t) acc[t] = 0.f;
  for (int j = slice; j < Nk; j += 16) {
    const float vv = vb[(long long)j * kv_rs + c];
#pragma unroll
    for (int t = 0; t < T2I_MAX_NT; ++t)
      if (t < Nt) acc[t] = fmaf(s_p[t * Nk + j], vv, acc[t]);
  }
#pragma unroll
  for (int t = 0; t < T2I_MAX_NT; ++t)
    if (t < Nt) s_red[(slice * Nt + t) * T2I_DH + c] = acc[t];
  __syncthreads();
  for (int i = tid; i < Nt * T2I_DH; i += blockDim.x) {
    const int t = i / T2I_DH, cc = i - t * T2I_DH;
    float s = 0.f;
#pragma unroll
    for (int sl = 0; sl < 16; ++sl) s += s_red[(sl * Nt + t) * T2I_DH + cc];
    out[((long long)b * Nt + t) * o_rs + h * T2I_DH + cc] = s;
  }
}

// ---- image -> token attention --------------------------------------------------------------------
__global__ void __launch_bounds__(256)
attn_i2t_kernel(const float* __restrict__ q, int q_rs, const float* __restrict__ k, const float* __restrict__ v,
                int kv_rs, float* __restrict__ out, int o_rs, int H, int Nq, int Nt, float scale) {
  extern __shared__ float sm[];  // k [Nt][H*16], v [Nt][H*16]
  const int C = H * T2I_DH;
  float* s_k = sm;
  float* s_v = sm + Nt * C;
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < Nt * C; i += blockDim.x) {
    const int t = i / C, cc = i - t * C;
    s_k[i] = k[((long long)b * Nt + t) * kv_rs + cc];
    s_v[i] = v[((long long)b * Nt + t) * kv_rs + cc];
  }
  __syncthreads();
  const int item = blockIdx.x * blockDim.x + threadIdx.x;  // (query, head)
  if (item >= Nq * H) return;
  const int qi = item / H, h = item - qi * H;
  float qr[T2I_DH];
  const float4* qp = reinterpret_cast<const float4*>(q + ((long long)b * Nq + qi) * q_rs + h * T2I_DH);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 t4 = qp[i];
    qr[4 * i] = t4.x * scale; qr[4 * i + 1] = t4.y * scale; qr[4 * i + 2] = t4.z * scale; qr[4 * i + 3] = t4.w * scale;
  }
  float sc[T2I_MAX_NT];
  float mx = -INFINITY;
#pragma unroll
  for (int t = 0; t < T2I_MAX_NT; ++t) {
    if (t < Nt) {
      float d = 0.f;
#pragma unroll
      for (int c = 0; c < T2I_DH; ++c) d = fmaf(qr[c], s_k[t * C + h * T2I_DH + c], d);
      sc[t] = d;
      mx = fmaxf(mx, d);
    }
  }
  float sum = 0.f;
#pragma unroll
  for (int t = 0; t < T2I_MAX_NT; ++t)
    if (t < Nt) {
      sc[t] = __expf(sc[t] - mx);
      sum += sc[t];
    }
  const float inv = 1.0f / sum;
  float o[T2I_DH];
#pragma unroll
  for (int c = 0; c < T2I_DH; ++c) o[c] = 0.f;
#pragma unroll
  for (int t = 0; t < T2I_MAX_NT; ++t)
    if (t < Nt) {
      const float pw = sc[t] * inv;
#pragma unroll
      for (int c = 0; c < T2I_DH; ++c) o[c] = fmaf(pw, s_v[t * C + h * T2I_DH + c], o[c]);
    }
  float4* op = reinterpret_cast<float4*>(out + ((long long)b * Nq + qi) * o_rs + h * T2I_DH);
#pragma unroll
  for (int i = 0; i < 4; ++i) op[i] = make_float4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
}

}  // namespace

#define STREAM reinterpret_cast<cudaStream_t>(stream)

extern "C" int usvm_gemm_skinny_f32(const usvm_skinny_params* p, void* stream) {
  if (!p || !p->x || !p->w || !p->out || p->M <= 0 || p->N <= 0 || p->K <= 0 || p->instances <= 0) return USVM_ERR_ARG;
  if ((p->K % 4) || (reinterpret_cast<uintptr_t>(p->w) & 15) || (p->w_is % 4)) return USVM_ERR_ARG;
  const size_t smem = (size_t)SK_ROWS * p->K * sizeof(float);
  static size_t configured = 48 * 1024;
  if (smem > configured) {
    if (smem > 200 * 1024) return USVM_ERR_ARG;
    if (cudaFuncSetAttribute(gemm_skinny_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess)
      return USVM_ERR_CUDA;
    configured = 200 * 1024;
  }
  dim3 grid(cdiv(p->N, SK_WARPS * SK_COLS_WARP), p->instances, cdiv(p->M, SK_ROWS));
  gemm_skinny_kernel<<<grid, 256, smem, STREAM>>>(*p);
  return usvm_check_launch();
}

extern "C" int usvm_attn_t2i_f32(const float* q, int q_rs, const float* k, const float* v, int kv_rs, float* out,
                                 int o_rs, int B, int H, int Nt, int Nk, float scale, void* stream) {
  if (!q || !k || !v || !out || B <= 0 || H <= 0 || Nt <= 0 || Nt > T2I_MAX_NT || Nk <= 0 || (kv_rs % 4)) return USVM_ERR_ARG;
  const size_t smem = ((size_t)Nt * Nk + Nt * T2I_DH + 16 * Nt * T2I_DH) * sizeof(float);
  static size_t configured = 48 * 1024;
  if (smem > configured) {
    if (smem > 200 * 1024) return USVM_ERR_ARG;
    if (cudaFuncSetAttribute(attn_t2i_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess)
      return USVM_ERR_CUDA;
    configured = 200 * 1024;
  }
  attn_t2i_kernel<<<B * H, 256, smem, STREAM>>>(q, q_rs, k, v, kv_rs, out, o_rs, H, Nt, Nk, scale);
  return usvm_check_launch();
}

extern "C" int usvm_attn_i2t_f32(const float* q, int q_rs, const float* k, const float* v, int kv_rs, float* out,
                                 int o_rs, int B, int H, int Nq, int Nt, float scale, void* stream) {
  if (!q || !k || !v || !out || B <= 0 || H <= 0 || Nt <= 0 || Nt > T2I_MAX_NT || Nq <= 0 || (q_rs % 4) || (o_rs % 4))
    return USVM_ERR_ARG;
  const size_t smem = (size_t)2 * Nt * H * T2I_DH * sizeof(float);
  attn_i2t_kernel<<<dim3(cdiv((long long)Nq * H, 256), B), 256, smem, STREAM>>>(q, q_rs, k, v, kv_rs, out, o_rs, H, Nq,
                                                                                 Nt, scale);
  return usvm_check_launch();
}
