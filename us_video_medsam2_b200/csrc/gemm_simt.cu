// fp32-accumulate SIMT GEMM with the same fused epilogue as the tensor-core kernel.
//   C[M,N] = epilogue(A[M,K] * W[N,K]^T),  A and W fp32 or bf16, any M/N/K.
// Used (a) for the fp32 tail of the path -- the SAM mask decoder (mask_decoder.py:168-245,
// sam/transformer.py:90-286) whose logits sit within ~0.05 of the 0 threshold at random init, so its
// contractions stay in full fp32 (DESIGN.md "precision plan") -- and (b) as the independent checker
// for usvm_gemm_bf16_tc5 in the GPU tests.
#include "common.cuh"
#include "usvm2_b200.h"

namespace {

constexpr int TM = 64, TN = 64, TK = 16;

__device__ __forceinline__ float ld_as_float(const float* p) { return *p; }
__device__ __forceinline__ float ld_as_float(const bf16* p) { return __bfloat162float(*p); }

__device__ __forceinline__ float act_fn(float v, int act) {
  if (act == USVM_ACT_RELU) return fmaxf(v, 0.0f);
  if (act == USVM_ACT_GELU) return gelu_erf(v);
  return v;
}

template <typename TA, typename TW>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(const TA* __restrict__ A, int lda, const TW* __restrict__ W, int ldw, const usvm_gemm_epilogue ep,
                 int M, int N, int K) {
  PDL_ENTRY();
  __shared__ float As[TK][TM + 4];
  __shared__ float Ws[TK][TN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads, 4 x 4 outputs each
  const int m0 = blockIdx.x * TM, n0 = blockIdx.y * TN;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += TK) {
    // 64 rows x 16 k per operand = 1024 elements, 4 per thread; consecutive threads walk k (contiguous)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = tid + i * 256;
      const int r = e >> 4, k = e & 15;
      const int gm = m0 + r, gn = n0 + r, gk = k0 + k;
      As[k][r] = (gm < M && gk < K) ? ld_as_float(A + (long long)gm * lda + gk) : 0.0f;
      Ws[k][r] = (gn < N && gk < K) ? ld_as_float(W + (long long)gn * ldw + gk) : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 w = *reinterpret_cast<const float4*>(&Ws[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = m0 + ty * 4 + i;
    if (row >= M) continue;
    const long long rrow = (ep.res_div > 0 ? ((long long)(row / ep.res_div) * ep.res_mod + row % ep.res_mod)
                                           : (ep.res_mod > 0 ? (row % ep.res_mod) : row));
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = n0 + tx * 4 + j;
      if (col >= N) continue;
      float x = acc[i][j];
      if (ep.bias) x += ep.bias[col];
      x = act_fn(x, ep.act);
      if (ep.col_scale) x *= ep.col_scale[col];
      if (ep.residual) x += ep.residual[rrow * ep.ldr + col];
      if (ep.out_f32) ep.out_f32[(long long)row * ep.ldo_f32 + col] = x;
      if (ep.out_bf16) reinterpret_cast<bf16*>(ep.out_bf16)[(long long)row * ep.ldo_bf16 + col] = __float2bfloat16(x);
    }
  }
}

// fp32 x fp32 fast path: 16-byte global loads, register prefetch of the next k-tile while the current one is
// multiplied (double-buffered shared memory), 64 x 64 x 16 tiles, 4 x 4 outputs per thread.
// Requires K % 4 == 0, lda % 4 == 0, ldw % 4 == 0 and 16-byte aligned bases.
__global__ void __launch_bounds__(256)
gemm_simt_f32_pipelined_kernel(const float* __restrict__ A, int lda, const float* __restrict__ W, int ldw,
                               const usvm_gemm_epilogue ep, int M, int N, int K) {
  PDL_ENTRY();
  __shared__ __align__(16) float As[2][TK][TM + 4];
  __shared__ __align__(16) float Ws[2][TK][TN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * TM, n0 = blockIdx.y * TN;
  const int lr = tid >> 2, lk = (tid & 3) << 2;  // this thread's load slot: row lr, k offset lk..lk+3
  const bool a_ok = m0 + lr < M, w_ok = n0 + lr < N;
  const float* ap = A + (long long)(a_ok ? m0 + lr : 0) * lda + lk;
  const float* wp = W + (long long)(w_ok ? n0 + lr : 0) * ldw + lk;
  auto fetch = [&](int k0, float4& a, float4& w) {
    const bool k_ok = k0 + lk < K;  // K % 4 == 0: a float4 is entirely inside or outside
    a = (a_ok && k_ok) ? *reinterpret_cast<const float4*>(ap + k0) : make_float4(0.f, 0.f, 0.f, 0.f);
    w = (w_ok && k_ok) ? *reinterpret_cast<const float4*>(wp + k0) : make_float4(0.f, 0.f, 0.f, 0.f);
  };
  auto stash = [&](int buf, const float4& a, const float4& w) {
    As[buf][lk][lr] = a.x; As[buf][lk + 1][lr] = a.y; As[buf][lk + 2][lr] = a.z; As[buf][lk + 3][lr] = a.w;
    Ws[buf][lk][lr] = w.x; Ws[buf][lk + 1][lr] = w.y; Ws[buf][lk + 2][lr] = w.z; Ws[buf][lk + 3][lr] = w.w;
  };
  float acc[4][4] = {};
  float4 a_next, w_next;
  fetch(0, a_next, w_next);
  stash(0, a_next, w_next);
  __syncthreads();
  const int nk = (K + TK - 1) / TK;
  for (int t = 0; t < nk; ++t) {
    const int buf = t & 1;
    if (t + 1 < nk) fetch((t + 1) * TK, a_next, w_next);  // in flight while this tile is multiplied
#pragma unroll
    for (int k = 0; k < TK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 w = *reinterpret_cast<const float4*>(&Ws[buf][k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
    }
    if (t + 1 < nk) {
      stash(buf ^ 1, a_next, w_next);  // the other buffer was last read in iteration t - 1, before the barrier below
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = m0 + ty * 4 + i;
    if (row >= M) continue;
    const long long rrow = (ep.res_div > 0 ? ((long long)(row / ep.res_div) * ep.res_mod + row % ep.res_mod)
                                           : (ep.res_mod > 0 ? (row % ep.res_mod) : row));
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = n0 + tx * 4 + j;
      if (col >= N) continue;
      float x = acc[i][j];
      if (ep.bias) x += ep.bias[col];
      x = act_fn(x, ep.act);
      if (ep.col_scale) x *= ep.col_scale[col];
      if (ep.residual) x += ep.residual[rrow * ep.ldr + col];
      if (ep.out_f32) ep.out_f32[(long long)row * ep.ldo_f32 + col] = x;
      if (ep.out_bf16) reinterpret_cast<bf16*>(ep.out_bf16)[(long long)row * ep.ldo_bf16 + col] = __float2bfloat16(x);
    }
  }
}

template <typename TA, typename TW>
int launch(const void* A, int lda, const void* W, int ldw, const usvm_gemm_epilogue* ep, int M, int N, int K,
           cudaStream_t s) {
  dim3 grid(cdiv(M, TM), cdiv(N, TN));
  usvm_launch(gemm_simt_kernel<TA, TW>, dim3(grid), dim3(256), 0, s, static_cast<const TA*>(A), lda, static_cast<const TW*>(W), ldw, *ep,
                                                M, N, K);
  return usvm_check_launch();
}

}  // namespace

extern "C" int usvm_gemm_simt(const void* A, int a_is_bf16, int lda, const void* W, int w_is_bf16, int ldw,
                              const usvm_gemm_epilogue* ep, int M, int N, int K, void* stream) {
  if (!A || !W || !ep || M <= 0 || N <= 0 || K <= 0) return USVM_ERR_ARG;
  if (ep->rope_cos) return USVM_ERR_ARG;  // fused RoPE exists on the tensor-core kernel only
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (a_is_bf16 && w_is_bf16) return launch<bf16, bf16>(A, lda, W, ldw, ep, M, N, K, s);
  if (!a_is_bf16 && !w_is_bf16) {
    const bool vec_ok = (K % 4 == 0) && (lda % 4 == 0) && (ldw % 4 == 0) &&
                        !(reinterpret_cast<uintptr_t>(A) & 15) && !(reinterpret_cast<uintptr_t>(W) & 15);
    if (vec_ok) {
      dim3 grid(cdiv(M, TM), cdiv(N, TN));
      usvm_launch(gemm_simt_f32_pipelined_kernel, dim3(grid), dim3(256), 0, s, static_cast<const float*>(A), lda,
                                                           static_cast<const float*>(W), ldw, *ep, M, N, K);
      return usvm_check_launch();
    }
    return launch<float, float>(A, lda, W, ldw, ep, M, N, K, s);
  }
  if (a_is_bf16) return launch<bf16, float>(A, lda, W, ldw, ep, M, N, K, s);
  return launch<float, bf16>(A, lda, W, ldw, ep, M, N, K, s);
}
