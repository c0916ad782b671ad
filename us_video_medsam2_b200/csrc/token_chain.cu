// Token side of the SAM mask decoder's two-way transformer as ONE kernel per dependency chain
// (sam/transformer.py:137-212 TwoWayAttentionBlock, mask_decoder.py:215-253 hyper-network / IoU / score heads,
// sam2_base.py:1143-1156 obj_ptr_proj).
//
// A tracked frame carries 8 decoder tokens per object.  Every token-side layer is a [8 x K] x [N x K]^T product whose
// cost is streaming the fp32 weight matrix once; as separate launches each of the ~38 layers paid a full kernel
// boundary (~3-4 us) for ~1 us of work.  Here a thread-block CLUSTER (8 or 16 CTAs, one cluster per object) runs a
// whole chain of such layers: the cluster's warps split the output columns of each step (so the weight stream is
// spread over 8-16 SMs' L2 ports), step outputs go to small global buffers, and steps are separated by a
// release/acquire cluster barrier instead of a kernel boundary.  Input-side transforms are recomputed per CTA because
// they are tiny: LayerNorm of the 8 rows, adding the positional tokens, the 8 x 8 token self-attention, and the
// merge of the per-CTA token->image attention partials.
//
// Step kinds:
//   LINEAR       out[m, n] = act((T(x)[m] (+ x2[m] for n < x2_cols)) . W[n] + b[n]) (+ residual[m, n])
//                T = identity | LayerNorm (optionally also written out) | self-attention over the q|k|v columns of x
//                  | softmax-merge of the T2I partials; w_is != 0 selects "row m uses matrix m" (stacked heads).
//   T2I_PARTIAL  token->image attention (8 heads x 16): CTA r of the cluster handles keys [r*Nk/CL, (r+1)*Nk/CL) of
//                all heads and writes unnormalised (max, sum, P.V) partials to the scratch buffer.
#include <cooperative_groups.h>

#include "common.cuh"
#include "usvm2_b200.h"

namespace {

constexpr int TC_THREADS = 256;
constexpr int TC_WARPS = 8;
constexpr int TC_ROWS = USVM_CHAIN_ROWS;  // 8
constexpr int TC_KMAX = 2048;
constexpr int TC_X2_FLOATS = TC_ROWS * 768;  // positional-token copy of the input (K <= 768) / q|k|v staging
constexpr int TC_HEADS = 8;
constexpr int TC_T2I_DH = 16;
constexpr int TC_T2I_C = TC_HEADS * TC_T2I_DH;  // 128
constexpr int TC_KEYS = 128;                    // keys per CTA in a T2I_PARTIAL step (Nk <= 128 * cluster)
constexpr size_t TC_SMEM = (size_t)(TC_ROWS * TC_KMAX + TC_X2_FLOATS + TC_ROWS * TC_T2I_C + 64) * sizeof(float);

__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// release/acquire at cluster scope: global writes of every CTA before the barrier are visible to all CTAs after it
__device__ __forceinline__ void cluster_barrier() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }

// sum over the warp of 32 per-lane values a[0..31]; lane l ends up with the total of a[l] (31 shuffles)
__device__ __forceinline__ float transpose_reduce(float (&a)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = upper ? a[i] : a[i + off];
      const float keep = upper ? a[i + off] : a[i];
      a[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return a[0];
}

// ---- input transforms: fill xs[rows][K] (and xs2 = xs + x2) --------------------------------------------------
__device__ void load_rows(const usvm_chain_step& st, int obj, int rank, float* xs, float* xs2) {
  const int K = st.K, K4 = K >> 2, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* xb = st.x + (long long)obj * st.x_os + (st.row_select ? (long long)st.row_select[obj] * st.sel_stride : 0);
  for (int i = tid; i < TC_ROWS * K4; i += TC_THREADS) {
    const int m = i / K4, k = (i - m * K4) << 2;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (m < st.rows) v = ldcg4(xb + (long long)m * st.x_rs + k);
    *reinterpret_cast<float4*>(xs + m * K + k) = v;
  }
  __syncthreads();
  if (st.ln_w) {  // warp m normalises row m (same two-pass statistics as layernorm_reg_kernel)
    const int m = warp;
    if (m < st.rows) {
      float s = 0.f;
      for (int c = lane; c < K; c += 32) s += xs[m * K + c];
      const float mean = warp_sum(s) / K;
      float q = 0.f;
      for (int c = lane; c < K; c += 32) {
        const float d = xs[m * K + c] - mean;
        q = fmaf(d, d, q);
      }
      const float rstd = 1.0f / sqrtf(warp_sum(q) / K + st.ln_eps);
      float* lo = (st.ln_out && rank == 0) ? st.ln_out + (long long)obj * st.ln_os + (long long)m * st.ln_rs : nullptr;
      for (int c = lane; c < K; c += 32) {
        const float y = (xs[m * K + c] - mean) * rstd * st.ln_w[c] + st.ln_b[c];
        xs[m * K + c] = y;
        if (lo) lo[c] = y;
      }
    }
    __syncthreads();
  }
  if (st.x2) {
    const float* x2b = st.x2 + (long long)obj * st.x2_os;
    for (int i = tid; i < TC_ROWS * K4; i += TC_THREADS) {
      const int m = i / K4, k = (i - m * K4) << 2;
      float4 v = *reinterpret_cast<const float4*>(xs + m * K + k);
      if (m < st.rows) {
        const float4 a = ldcg4(x2b + (long long)m * st.x2_rs + k);
        v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
      }
      *reinterpret_cast<float4*>(xs2 + m * K + k) = v;
    }
    __syncthreads();
  }
}

// self-attention of the rows over themselves: x rows hold q | k | v at column offsets attn_q/k/v, 8 heads x 32
__device__ void load_self_attention(const usvm_chain_step& st, int obj, float* xs, float* qkv, float* sc) {
  constexpr int C = 256, DH = 32;
  const int tid = threadIdx.x, rows = st.rows;
  const float* xb = st.x + (long long)obj * st.x_os;
  for (int i = tid; i < TC_ROWS * 3 * (C / 4); i += TC_THREADS) {  // stage q | k | v of every row
    const int m = i / (3 * C / 4), r = i - m * (3 * C / 4), part = r / (C / 4), k = (r - part * (C / 4)) << 2;
    const int off = part == 0 ? st.attn_q : part == 1 ? st.attn_k : st.attn_v;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (m < rows) v = ldcg4(xb + (long long)m * st.x_rs + off + k);
    *reinterpret_cast<float4*>(qkv + (m * 3 + part) * C + k) = v;
  }
  __syncthreads();
  const float scale = 0.17677669529663687f;  // 1 / sqrt(32)
  for (int i = tid; i < TC_HEADS * TC_ROWS * TC_ROWS; i += TC_THREADS) {  // scores [head][query][key]
    const int h = i >> 6, qi = (i >> 3) & 7, kj = i & 7;
    float d = -INFINITY;
    if (qi < rows && kj < rows) {
      d = 0.f;
      const float* qp = qkv + (qi * 3 + 0) * C + h * DH;
      const float* kp = qkv + (kj * 3 + 1) * C + h * DH;
#pragma unroll
      for (int c = 0; c < DH; ++c) d = fmaf(qp[c] * scale, kp[c], d);
    }
    sc[i] = d;
  }
  __syncthreads();
  if (tid < TC_HEADS * TC_ROWS) {  // softmax over the 8 keys of one (head, query)
    float* s = sc + tid * 8;
    const int qi = tid & 7;
    if (qi < rows) {
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 8; ++j) mx = fmaxf(mx, s[j]);
      float e[8], sum = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        e[j] = __expf(s[j] - mx);
        sum += e[j];
      }
      const float inv = 1.0f / sum;
#pragma unroll
      for (int j = 0; j < 8; ++j) s[j] = e[j] * inv;
    }
  }
  __syncthreads();
  for (int i = tid; i < TC_ROWS * C; i += TC_THREADS) {  // out[query][h*32 + c]
    const int qi = i / C, col = i - qi * C, h = col / DH;
    float o = 0.f;
    if (qi < rows) {
      const float* pr = sc + (h * 8 + qi) * 8;
      for (int j = 0; j < rows; ++j) o = fmaf(pr[j], qkv[(j * 3 + 2) * C + col], o);
    }
    xs[qi * st.K + col] = o;
  }
  __syncthreads();
}

// merge of the cluster's token->image partials: xs[row][128] = sum_r e^{m_r - M} o_r / sum_r e^{m_r - M} l_r
__device__ void load_t2i_merge(const usvm_chain_step& st, const float* scratch, int cl, float* xs) {
  const float* po = scratch;                                     // [cl][8][128]
  const float* ml = scratch + (size_t)cl * TC_ROWS * TC_T2I_C;   // [cl][8][8][2]
  for (int i = threadIdx.x; i < TC_ROWS * TC_T2I_C; i += TC_THREADS) {
    const int row = i / TC_T2I_C, ch = i - row * TC_T2I_C, h = ch / TC_T2I_DH;
    float o = 0.f;
    if (row < st.rows) {
      float M = -INFINITY;
      for (int r = 0; r < cl; ++r) M = fmaxf(M, __ldcg(ml + ((r * TC_ROWS + row) * TC_HEADS + h) * 2));
      float L = 0.f;
      for (int r = 0; r < cl; ++r) {
        const float2 v = __ldcg(reinterpret_cast<const float2*>(ml + ((r * TC_ROWS + row) * TC_HEADS + h) * 2));
        const float wgt = __expf(v.x - M);
        L = fmaf(v.y, wgt, L);
        o = fmaf(__ldcg(po + (r * TC_ROWS + row) * TC_T2I_C + ch), wgt, o);
      }
      o /= L;
    }
    xs[row * st.K + ch] = o;
  }
  __syncthreads();
}

// ---- one LINEAR step: this CTA's share of the output columns ------------------------------------------------
__device__ void linear_step(const usvm_chain_step& st, int obj, int rank, int cl, const float* xs, const float* xs2) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int K = st.K, N = st.N, rows = st.rows;
  const int gpn = (N + 3) >> 2;                          // 4-column groups per matrix
  const int groups = st.w_is ? rows * gpn : gpn;
  for (int g = rank * TC_WARPS + warp; g < groups; g += cl * TC_WARPS) {
    const int inst = st.w_is ? g / gpn : 0;
    const int n0 = (st.w_is ? g - inst * gpn : g) << 2;
    const float* Wb = st.w + (long long)inst * st.w_is;
    const float* xin = (st.x2 && n0 < st.x2_cols) ? xs2 : xs;
    float acc[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = 0.f;
    if (st.w_is) {  // one input row (the instance's) against 4 weight rows
#pragma unroll 2
      for (int k = lane * 4; k < K; k += 128) {
        const float4 xv = *reinterpret_cast<const float4*>(xin + inst * K + k);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float4 wv = __ldg(reinterpret_cast<const float4*>(Wb + (long long)min(n0 + c, N - 1) * K + k));
          acc[c * 8] += xv.x * wv.x + xv.y * wv.y + xv.z * wv.z + xv.w * wv.w;
        }
      }
    } else {
#pragma unroll 2
      for (int k = lane * 4; k < K; k += 128) {
        float4 wv[4];
#pragma unroll
        for (int c = 0; c < 4; ++c)
          wv[c] = __ldg(reinterpret_cast<const float4*>(Wb + (long long)min(n0 + c, N - 1) * K + k));
#pragma unroll
        for (int m = 0; m < TC_ROWS; ++m) {
          const float4 xv = *reinterpret_cast<const float4*>(xin + m * K + k);
#pragma unroll
          for (int c = 0; c < 4; ++c)
            acc[c * 8 + m] += xv.x * wv[c].x + xv.y * wv[c].y + xv.z * wv[c].z + xv.w * wv[c].w;
        }
      }
    }
    float v = transpose_reduce(acc, lane);  // lane = c * 8 + m
    const int c = lane >> 3, mm = lane & 7;
    const int m = st.w_is ? inst : mm;
    const int n = n0 + c;
    const bool mine = st.w_is ? (mm == 0) : (mm < rows);
    if (mine && n < N) {
      if (st.bias) v += st.bias[(long long)inst * st.b_is + n];
      if (st.act == USVM_ACT_RELU) v = fmaxf(v, 0.f);
      else if (st.act == USVM_ACT_GELU) v = gelu_erf(v);
      if (st.residual) v += __ldcg(st.residual + (long long)obj * st.r_os + (long long)m * st.r_rs + n);
      st.out[(long long)obj * st.o_os + (long long)m * st.o_rs + n] = v;
    }
  }
}

// ---- one T2I_PARTIAL step -----------------------------------------------------------------------------------
__device__ void t2i_partial_step(const usvm_chain_step& st, int obj, int rank, int cl, float* scratch, float* sc,
                                 float* sq) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, rows = st.rows;
  const int per = (st.Nk + cl - 1) / cl;  // <= TC_KEYS
  const int key0 = rank * per;
  const int nkeys = max(0, min(per, st.Nk - key0));
  const float* xb = st.x + (long long)obj * st.x_os;
  for (int i = tid; i < TC_ROWS * TC_T2I_C; i += TC_THREADS) {
    const int m = i / TC_T2I_C, c = i - m * TC_T2I_C;
    sq[i] = m < rows ? __ldcg(xb + (long long)m * st.x_rs + c) * 0.25f : 0.f;  // 1 / sqrt(16)
  }
  __syncthreads();
  const float* kb = st.k + (long long)obj * st.kv_os + (long long)key0 * st.kv_rs;
  const float* vb = st.v + (long long)obj * st.kv_os + (long long)key0 * st.kv_rs;
  {  // scores: thread <-> (key, half of the heads); sc[(row * 8 + head) * 128 + key]
    const int j = tid & (TC_KEYS - 1), hh = tid >> 7;
    if (j < nkeys) {
#pragma unroll
      for (int h4 = 0; h4 < 4; ++h4) {
        const int h = hh * 4 + h4;
        float kr[TC_T2I_DH];
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          const float4 t = __ldg(reinterpret_cast<const float4*>(kb + (long long)j * st.kv_rs + h * TC_T2I_DH + c4 * 4));
          kr[c4 * 4] = t.x; kr[c4 * 4 + 1] = t.y; kr[c4 * 4 + 2] = t.z; kr[c4 * 4 + 3] = t.w;
        }
#pragma unroll
        for (int m = 0; m < TC_ROWS; ++m) {
          float d = 0.f;
#pragma unroll
          for (int c = 0; c < TC_T2I_DH; ++c) d = fmaf(sq[m * TC_T2I_C + h * TC_T2I_DH + c], kr[c], d);
          sc[(m * TC_HEADS + h) * TC_KEYS + j] = d;
        }
      }
    } else {
#pragma unroll
      for (int h4 = 0; h4 < 4; ++h4)
#pragma unroll
        for (int m = 0; m < TC_ROWS; ++m) sc[(m * TC_HEADS + hh * 4 + h4) * TC_KEYS + j] = -INFINITY;
    }
  }
  __syncthreads();
  float* po = scratch + ((size_t)rank * TC_ROWS) * TC_T2I_C;
  float* ml = scratch + (size_t)cl * TC_ROWS * TC_T2I_C + ((size_t)rank * TC_ROWS) * TC_HEADS * 2;
  for (int pr = warp; pr < TC_ROWS * TC_HEADS; pr += TC_WARPS) {  // (row, head) pairs: local max / exp / sum
    float* s = sc + pr * TC_KEYS;
    float e[4], mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      e[i] = s[lane + 32 * i];
      mx = fmaxf(mx, e[i]);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    if (mx == -INFINITY) {  // this CTA holds no keys
#pragma unroll
      for (int i = 0; i < 4; ++i) s[lane + 32 * i] = 0.f;
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        e[i] = __expf(e[i] - mx);
        s[lane + 32 * i] = e[i];
        sum += e[i];
      }
    }
    sum = warp_sum(sum);
    if (lane == 0) {
      ml[pr * 2] = mx;
      ml[pr * 2 + 1] = sum;
    }
  }
  __syncthreads();
  {  // P.V: thread <-> (channel, half of the rows)
    const int ch = tid & (TC_T2I_C - 1), rh = tid >> 7, h = ch / TC_T2I_DH;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
    for (int j = 0; j < nkeys; ++j) {
      const float vv = __ldg(vb + (long long)j * st.kv_rs + ch);
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[r] = fmaf(sc[((rh * 4 + r) * TC_HEADS + h) * TC_KEYS + j], vv, acc[r]);
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) po[(rh * 4 + r) * TC_T2I_C + ch] = acc[r];
  }
}

__global__ void __launch_bounds__(TC_THREADS, 1) token_chain_kernel(const __grid_constant__ usvm_chain_params p) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ __align__(16) float smem_f[];
  float* xs = smem_f;                         // [8][K]; doubles as the score buffer of T2I_PARTIAL / self-attention
  float* xs2 = xs + TC_ROWS * TC_KMAX;        // [8][K <= 768] input + positional tokens; q|k|v staging
  float* sq = xs2 + TC_X2_FLOATS;             // [8][128] scaled queries of T2I_PARTIAL
  const int cl = p.cluster;
  const int rank = (int)cluster_rank();
  const int obj = blockIdx.x / cl;
  float* scratch = p.scratch + (size_t)obj * cl * TC_ROWS * (TC_T2I_C + TC_HEADS * 2);
  for (int s = 0; s < p.n_steps; ++s) {
    const usvm_chain_step& st = p.steps[s];
    if (st.kind == USVM_CHAIN_T2I_PARTIAL) {
      t2i_partial_step(st, obj, rank, cl, scratch, xs, sq);
    } else {
      if (st.in_kind == USVM_CHAIN_IN_SELF_ATTN) load_self_attention(st, obj, xs, xs2, xs + TC_ROWS * 256);
      else if (st.in_kind == USVM_CHAIN_IN_T2I_MERGE) load_t2i_merge(st, scratch, cl, xs);
      else load_rows(st, obj, rank, xs, xs2);
      linear_step(st, obj, rank, cl, xs, xs2);
    }
    if (s + 1 < p.n_steps) cluster_barrier();  // also orders this CTA's shared-memory reuse (all threads take part)
  }
}

}  // namespace

extern "C" int usvm_token_chain(const usvm_chain_params* p, void* stream) {
  if (!p || p->n_steps <= 0 || p->n_steps > USVM_CHAIN_MAX_STEPS || p->n_obj <= 0) return USVM_ERR_ARG;
  if (p->cluster != 8 && p->cluster != 16) return USVM_ERR_ARG;
  for (int s = 0; s < p->n_steps; ++s) {
    const usvm_chain_step& st = p->steps[s];
    if (st.rows <= 0 || st.rows > USVM_CHAIN_ROWS || !st.x) return USVM_ERR_ARG;
    if (st.kind == USVM_CHAIN_T2I_PARTIAL) {
      if (!st.k || !st.v || !p->scratch || st.Nk <= 0 || st.Nk > TC_KEYS * p->cluster || (st.kv_rs % 4)) return USVM_ERR_ARG;
    } else if (st.kind == USVM_CHAIN_LINEAR) {
      if (!st.w || !st.out || st.N <= 0 || st.K <= 0 || (st.K % 4) || st.K > TC_KMAX) return USVM_ERR_ARG;
      if ((reinterpret_cast<uintptr_t>(st.w) & 15) || (st.w_is % 4)) return USVM_ERR_ARG;
      if (st.in_kind == USVM_CHAIN_IN_ROWS) {
        if ((st.x_rs % 4) || (st.x_os % 4) || (reinterpret_cast<uintptr_t>(st.x) & 15)) return USVM_ERR_ARG;
        if (st.x2 && (st.K > 768 || (st.x2_rs % 4) || (st.x2_os % 4) || (st.x2_cols % 4))) return USVM_ERR_ARG;
      } else if (st.in_kind == USVM_CHAIN_IN_SELF_ATTN) {
        if (st.K != 256 || (st.x_rs % 4) || ((st.attn_q | st.attn_k | st.attn_v) % 4) || st.x2) return USVM_ERR_ARG;
      } else if (st.in_kind == USVM_CHAIN_IN_T2I_MERGE) {
        if (st.K != 128 || !p->scratch || st.x2) return USVM_ERR_ARG;
      } else {
        return USVM_ERR_ARG;
      }
    } else {
      return USVM_ERR_ARG;
    }
  }
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(token_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM) != cudaSuccess ||
        cudaFuncSetAttribute(token_chain_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess)
      return USVM_ERR_CUDA;
    configured = true;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(p->n_obj * p->cluster);
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = TC_SMEM;
  cfg.stream = reinterpret_cast<cudaStream_t>(stream);
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = p->cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = usvm_pdl_enabled() ? 2 : 1;
  if (cudaLaunchKernelEx(&cfg, token_chain_kernel, *p) != cudaSuccess) return USVM_ERR_CUDA;
  return usvm_check_launch();
}
