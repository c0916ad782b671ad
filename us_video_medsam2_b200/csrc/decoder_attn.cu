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
constexpr int SK_COLS = 4;       // output columns per warp group
constexpr int SK_WARPS = 8;

// KSPLIT = 1: each warp owns 4 output columns (CTA = 32 columns), lanes stride K.
// KSPLIT = 8: the 8 warps of a CTA share 4 output columns and split K (long reductions, e.g. K = 2048), partial sums
//             are combined through shared memory.
template <int KSPLIT>
__global__ void __launch_bounds__(256)
gemm_skinny_kernel(const usvm_skinny_params p) {
  extern __shared__ __align__(16) float xs[];  // [8][K] x + x2, [8][32] partials, then [8][K] plain x when x2_cols < N
  const int inst = blockIdx.y;
  const int m0 = blockIdx.z * SK_ROWS;
  const int rows = min(SK_ROWS, p.M - m0);
  const int K = p.K, K4 = K >> 2;
  // The weights are constants of the model: nothing launched before this kernel writes them, so this warp's weight
  // rows are fetched BEFORE the programmatic-dependency wait -- their L2 / HBM latency overlaps the predecessor's tail and
  // the staging of x instead of following it.  (256-wide k ranges: K = 256, or K = 2048 split over the 8 warps.)
  const int pw_warp = threadIdx.x >> 5, pw_lane = threadIdx.x & 31;
  const int pw_n0 = (KSPLIT == 1 ? (blockIdx.x * SK_WARPS + pw_warp) : blockIdx.x) * SK_COLS;
  const int pw_chunk = KSPLIT == 1 ? K : (((K + KSPLIT - 1) / KSPLIT + 3) & ~3);
  const bool prefetched = pw_chunk == 256 && pw_n0 < p.N && (KSPLIT == 1 || (pw_warp + 1) * 256 <= K);
  float4 wpre[2][SK_COLS];
  if (prefetched) {
    const float* Wp = p.w + (long long)inst * p.w_is + (KSPLIT == 1 ? 0 : pw_warp * 256) + pw_lane * 4;
#pragma unroll
    for (int it = 0; it < 2; ++it)
#pragma unroll
      for (int c = 0; c < SK_COLS; ++c)
        wpre[it][c] = __ldg(reinterpret_cast<const float4*>(Wp + (long long)min(pw_n0 + c, p.N - 1) * K + it * 128));
  }
  // epilogue operands of this lane's output element (m, n0 + c): the bias is a constant (before the wait), the residual
  // is read right after the wait -- both long before the reduction needs them instead of after it
  const int e_c = pw_lane >> 3, e_m = pw_lane & 7, e_n = pw_n0 + e_c;
  const bool e_ok = e_m < rows && e_n < p.N;
  float e_bias = 0.f, e_res = 0.f;
  if (e_ok && p.bias) e_bias = __ldg(p.bias + (long long)inst * p.b_is + e_n);
  PDL_ENTRY();
  if (e_ok && p.residual) e_res = p.residual[(long long)inst * p.r_is + (long long)(m0 + e_m) * p.r_rs + e_n];
  // x2 (the positional tokens) applies to output columns < x2_cols only: projections that read `queries + pe` and
  // projections that read `queries` share one launch, with both variants of the input staged
  const bool two = p.x2 && p.x2_cols > 0 && p.x2_cols < p.N;
  float* xs_plain = xs + SK_ROWS * K + SK_WARPS * 32;
  if (p.ln_w) {
    // LayerNorm on load (K == 256): warp m owns row m; same lane <-> column mapping and operation order as
    // layernorm_reg_kernel<8>, so the result is bit-identical to the separate launch it replaces
    const int wrow = threadIdx.x >> 5, ln_lane = threadIdx.x & 31;
    if (wrow < rows) {
      const long long sel = p.row_select ? (long long)p.row_select[m0 + wrow] * p.x_sel_stride : 0;
      const float* xr = p.x + (long long)inst * p.x_is + (long long)(m0 + wrow) * p.x_rs + sel;
      const float* x2r = p.x2 ? p.x2 + (long long)inst * p.x2_is + (long long)(m0 + wrow) * p.x2_rs : nullptr;
      // every global load of the row is issued before the first reduction: one memory round trip, not three
      float v[8], gw[8], gb[8], a2[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int col = j * 32 + ln_lane;
        v[j] = xr[col];
        gw[j] = __ldg(p.ln_w + col);
        gb[j] = __ldg(p.ln_b + col);
        a2[j] = x2r ? x2r[col] : 0.f;
      }
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) sum += v[j];
      const float mean = warp_sum(sum) / 256;
      float q = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        v[j] -= mean;
        q = fmaf(v[j], v[j], q);
      }
      const float rstd = 1.0f / sqrtf(warp_sum(q) / 256 + p.ln_eps);
      float* lo = (p.ln_out && blockIdx.x == 0) ? p.ln_out + (long long)inst * p.ln_is + (long long)(m0 + wrow) * p.ln_rs
                                                  : nullptr;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int col = j * 32 + ln_lane;
        const float y = v[j] * rstd * gw[j] + gb[j];
        if (lo) lo[col] = y;
        if (two) xs_plain[wrow * K + col] = y;
        xs[wrow * K + col] = x2r ? y + a2[j] : y;
      }
    }
  } else
  for (int i = threadIdx.x; i < rows * K4; i += blockDim.x) {
    const int m = i / K4, k = (i - m * K4) << 2;
    const long long sel = p.row_select ? (long long)p.row_select[m0 + m] * p.x_sel_stride : 0;
    const float* xp = p.x + (long long)inst * p.x_is + (long long)(m0 + m) * p.x_rs + sel + k;
    float4 v = make_float4(xp[0], xp[1], xp[2], xp[3]);
    if (two) *reinterpret_cast<float4*>(xs_plain + m * K + k) = v;
    if (p.x2) {
      const float* x2 = p.x2 + (long long)inst * p.x2_is + (long long)(m0 + m) * p.x2_rs + k;
      v.x += x2[0]; v.y += x2[1]; v.z += x2[2]; v.w += x2[3];
    }
    *reinterpret_cast<float4*>(xs + m * K + k) = v;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = (KSPLIT == 1 ? (blockIdx.x * SK_WARPS + warp) : blockIdx.x) * SK_COLS;
  float acc[SK_COLS][SK_ROWS];
#pragma unroll
  for (int c = 0; c < SK_COLS; ++c)
#pragma unroll
    for (int m = 0; m < SK_ROWS; ++m) acc[c][m] = 0.f;
  if (n0 < p.N) {
    const float* Wb = p.w + (long long)inst * p.w_is;
    const float* xin = (two && n0 >= p.x2_cols) ? xs_plain : xs;
    const int kchunk = KSPLIT == 1 ? K : (((K + KSPLIT - 1) / KSPLIT + 3) & ~3);
    const int kbeg = KSPLIT == 1 ? 0 : warp * kchunk;
    const int kend = min(K, kbeg + kchunk);
    if (prefetched) {
#pragma unroll
      for (int it = 0; it < 2; ++it) {
        const int k = kbeg + lane * 4 + it * 128;
#pragma unroll
        for (int m = 0; m < SK_ROWS; ++m) {
          if (m < rows) {
            const float4 xv = *reinterpret_cast<const float4*>(xin + m * K + k);
#pragma unroll
            for (int c = 0; c < SK_COLS; ++c)
              acc[c][m] += xv.x * wpre[it][c].x + xv.y * wpre[it][c].y + xv.z * wpre[it][c].z + xv.w * wpre[it][c].w;
          }
        }
      }
    } else
    for (int k = kbeg + lane * 4; k < kend; k += 128) {
      float4 wv[SK_COLS];
#pragma unroll
      for (int c = 0; c < SK_COLS; ++c) {
        const int n = min(n0 + c, p.N - 1);
        wv[c] = *reinterpret_cast<const float4*>(Wb + (long long)n * K + k);
      }
#pragma unroll
      for (int m = 0; m < SK_ROWS; ++m) {
        if (m < rows) {
          const float4 xv = *reinterpret_cast<const float4*>(xin + m * K + k);
#pragma unroll
          for (int c = 0; c < SK_COLS; ++c)
            acc[c][m] += xv.x * wv[c].x + xv.y * wv[c].y + xv.z * wv[c].z + xv.w * wv[c].w;
        }
      }
    }
  }
#pragma unroll
  for (int c = 0; c < SK_COLS; ++c)
#pragma unroll
    for (int m = 0; m < SK_ROWS; ++m) acc[c][m] = warp_sum(acc[c][m]);
  // lane (c * 8 + m) holds element (m, n0 + c)
  const int c = lane >> 3, m = lane & 7;
  float v = 0.f;
#pragma unroll
  for (int cc = 0; cc < SK_COLS; ++cc)
#pragma unroll
    for (int mm = 0; mm < SK_ROWS; ++mm)
      if (cc == c && mm == m) v = acc[cc][mm];
  if (KSPLIT > 1) {
    float* part = xs + SK_ROWS * K;  // [8 warps][32]
    part[warp * 32 + lane] = v;
    __syncthreads();
    if (warp != 0) return;
    v = 0.f;
#pragma unroll
    for (int w = 0; w < SK_WARPS; ++w) v += part[w * 32 + lane];
  }
  const int n = n0 + c;
  if (m < rows && n < p.N) {  // (c, m, n) == (e_c, e_m, e_n): the operands were fetched at the top
    v += e_bias;
    if (p.act == USVM_ACT_RELU) v = fmaxf(v, 0.f);
    else if (p.act == USVM_ACT_GELU) v = gelu_erf(v);
    v += e_res;
    p.out[(long long)inst * p.o_is + (long long)(m0 + m) * p.o_rs + n] = v;
  }
}

// ---- token -> image attention --------------------------------------------------------------------
constexpr int T2I_DH = 16;
constexpr int T2I_MAX_NT = 16;
constexpr int T2I_LD = 17;  // padded row (16 channels + 1) -> conflict-free per-thread row reads

// one CTA per (object, head): the head's K and V slices ([Nk x 16] each) are first staged in shared memory with
// independent 16-byte loads (all in flight at once), then scores / softmax / P.V run entirely out of shared memory
__global__ void __launch_bounds__(256)
attn_t2i_kernel(const float* __restrict__ q, int q_rs, const float* __restrict__ k, const float* __restrict__ v,
                int kv_rs, float* __restrict__ out, int o_rs, int H, int Nt, int Nk, float scale) {
  PDL_ENTRY();
  extern __shared__ float sm[];
  float* s_k = sm;                         // [Nk][17]
  float* s_v = s_k + Nk * T2I_LD;          // [Nk][17]
  float* s_p = s_v + Nk * T2I_LD;          // [Nt][Nk]
  float* s_q = s_p + Nt * Nk;              // [Nt][16]
  float* s_red = s_q + Nt * T2I_DH;        // [16 slices][Nt][16]
  const int b = blockIdx.x / H, h = blockIdx.x - b * H;
  const int tid = threadIdx.x;
  const float* qb = q + (long long)b * Nt * q_rs + h * T2I_DH;
  const float* kb = k + (long long)b * Nk * kv_rs + h * T2I_DH;
  const float* vb = v + (long long)b * Nk * kv_rs + h * T2I_DH;
  for (int i = tid; i < Nk * 4; i += blockDim.x) {  // 4 float4 per key row, K and V
    const int j = i >> 2, c4 = (i & 3) << 2;
    const float4 kk = *reinterpret_cast<const float4*>(kb + (long long)j * kv_rs + c4);
    const float4 vv = *reinterpret_cast<const float4*>(vb + (long long)j * kv_rs + c4);
    float* dk = s_k + j * T2I_LD + c4;
    float* dv = s_v + j * T2I_LD + c4;
    dk[0] = kk.x; dk[1] = kk.y; dk[2] = kk.z; dk[3] = kk.w;
    dv[0] = vv.x; dv[1] = vv.y; dv[2] = vv.z; dv[3] = vv.w;
  }
  for (int i = tid; i < Nt * T2I_DH; i += blockDim.x) s_q[i] = qb[(i / T2I_DH) * q_rs + (i % T2I_DH)] * scale;
  __syncthreads();
  for (int j = tid; j < Nk; j += blockDim.x) {
    float kr[T2I_DH];
#pragma unroll
    for (int c = 0; c < T2I_DH; ++c) kr[c] = s_k[j * T2I_LD + c];
    for (int t = 0; t < Nt; ++t) {
      float d = 0.f;
#pragma unroll
      for (int c = 0; c < T2I_DH; ++c) d = fmaf(s_q[t * T2I_DH + c], kr[c], d);
      s_p[t * Nk + j] = d;
    }
  }
  __syncthreads();
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
    const float vv = s_v[j * T2I_LD + c];
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


// ---- token -> image attention, split over the keys ------------------------------------------------------------------
// grid (object x head, key split): each CTA scores its 128 keys for all tokens and writes an unnormalised partial
// (max, sum, P.V); the CTA that finishes last for a head (ticket counter) merges the partials in split order, so the
// result does not depend on arrival order.  8 heads x 8 splits = 64 CTAs instead of 8 for one object.
constexpr int T2S_KEYS = 128;
__global__ void __launch_bounds__(T2S_KEYS)
attn_t2i_split_kernel(const float* __restrict__ q, int q_rs, const float* __restrict__ k, const float* __restrict__ v,
                      int kv_rs, float* __restrict__ out, int o_rs, int H, int Nt, int Nk, float scale,
                      float* __restrict__ part, int* __restrict__ counters) {
  PDL_ENTRY();
  __shared__ float s_q[T2I_MAX_NT * T2I_DH];
  __shared__ float s_p[T2I_MAX_NT * T2S_KEYS];
  __shared__ float s_v[T2S_KEYS * T2I_LD];
  __shared__ float s_ml[T2I_MAX_NT * 2];
  __shared__ int s_last;
  const int bh = blockIdx.x, split = blockIdx.y, S = gridDim.y;
  const int b = bh / H, h = bh - b * H;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int j = split * T2S_KEYS + tid;
  const bool live = j < Nk;
  float4 kr[4], vr[4];
  if (live) {
    const float* kp = k + ((long long)b * Nk + j) * kv_rs + h * T2I_DH;
    const float* vp = v + ((long long)b * Nk + j) * kv_rs + h * T2I_DH;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      kr[i] = __ldg(reinterpret_cast<const float4*>(kp) + i);
      vr[i] = __ldg(reinterpret_cast<const float4*>(vp) + i);
    }
  }
  for (int i = tid; i < Nt * T2I_DH; i += T2S_KEYS)
    s_q[i] = q[((long long)b * Nt + i / T2I_DH) * q_rs + h * T2I_DH + (i % T2I_DH)] * scale;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float* dv = s_v + tid * T2I_LD + 4 * i;
    const float4 x = live ? vr[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    dv[0] = x.x; dv[1] = x.y; dv[2] = x.z; dv[3] = x.w;
  }
  for (int t = 0; t < Nt; ++t) {
    float d = -INFINITY;
    if (live) {
      d = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 qq = *reinterpret_cast<const float4*>(s_q + t * T2I_DH + 4 * i);
        d = fmaf(qq.x, kr[i].x, d); d = fmaf(qq.y, kr[i].y, d); d = fmaf(qq.z, kr[i].z, d); d = fmaf(qq.w, kr[i].w, d);
      }
    }
    s_p[t * T2S_KEYS + tid] = d;
  }
  __syncthreads();
  for (int t = warp; t < Nt; t += T2S_KEYS / 32) {  // local max / exp / sum of one token's 128 scores
    float e[4], mx = -INFINITY;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      e[u] = s_p[t * T2S_KEYS + lane + 32 * u];
      mx = fmaxf(mx, e[u]);
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      e[u] = mx == -INFINITY ? 0.f : __expf(e[u] - mx);
      s_p[t * T2S_KEYS + lane + 32 * u] = e[u];
      sum += e[u];
    }
    sum = warp_sum(sum);
    if (lane == 0) {
      s_ml[2 * t] = mx;
      s_ml[2 * t + 1] = sum;
    }
  }
  __syncthreads();
  // partial layout: [bh][split][Nt][16 + 2]
  float* mypart = part + ((long long)bh * S + split) * Nt * (T2I_DH + 2);
  for (int i = tid; i < Nt * T2I_DH; i += T2S_KEYS) {
    const int t = i / T2I_DH, c = i - t * T2I_DH;
    float acc = 0.f;
#pragma unroll 8
    for (int jj = 0; jj < T2S_KEYS; ++jj) acc = fmaf(s_p[t * T2S_KEYS + jj], s_v[jj * T2I_LD + c], acc);
    mypart[t * (T2I_DH + 2) + c] = acc;
  }
  for (int i = tid; i < Nt * 2; i += T2S_KEYS) mypart[(i >> 1) * (T2I_DH + 2) + T2I_DH + (i & 1)] = s_ml[i];
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(&counters[bh], 1) == S - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const float* hp = part + (long long)bh * S * Nt * (T2I_DH + 2);
  for (int i = tid; i < Nt * T2I_DH; i += T2S_KEYS) {
    const int t = i / T2I_DH, c = i - t * T2I_DH;
    if (S > 16) {  // 64 x 64 feature maps (32 splits): the same merge, in split order, as two loops
      float M = -INFINITY;
      for (int s2 = 0; s2 < S; ++s2) M = fmaxf(M, __ldcg(hp + ((long long)s2 * Nt + t) * (T2I_DH + 2) + T2I_DH));
      float L = 0.f, acc = 0.f;
      for (int s2 = 0; s2 < S; ++s2) {
        const float* pp = hp + ((long long)s2 * Nt + t) * (T2I_DH + 2);
        const float w = __expf(__ldcg(pp + T2I_DH) - M);
        L = fmaf(__ldcg(pp + T2I_DH + 1), w, L);
        acc = fmaf(__ldcg(pp + c), w, acc);
      }
      out[((long long)b * Nt + t) * o_rs + h * T2I_DH + c] = acc / L;
      continue;
    }
    float m[16], l[16], o[16];
#pragma unroll
    for (int s2 = 0; s2 < 16; ++s2) {
      if (s2 < S) {
        const float* pp = hp + ((long long)s2 * Nt + t) * (T2I_DH + 2);
        m[s2] = __ldcg(pp + T2I_DH);
        l[s2] = __ldcg(pp + T2I_DH + 1);
        o[s2] = __ldcg(pp + c);
      }
    }
    float M = -INFINITY;
#pragma unroll
    for (int s2 = 0; s2 < 16; ++s2)
      if (s2 < S) M = fmaxf(M, m[s2]);
    float L = 0.f, acc = 0.f;
#pragma unroll
    for (int s2 = 0; s2 < 16; ++s2) {
      if (s2 < S) {
        const float w = __expf(m[s2] - M);
        L = fmaf(l[s2], w, L);
        acc = fmaf(o[s2], w, acc);
      }
    }
    out[((long long)b * Nt + t) * o_rs + h * T2I_DH + c] = acc / L;
  }
  if (tid == 0) counters[bh] = 0;  // ready for the next launch (stream-ordered)
}

// ---- image -> token attention --------------------------------------------------------------------
__global__ void __launch_bounds__(256)
attn_i2t_kernel(const float* __restrict__ q, int q_rs, const float* __restrict__ k, const float* __restrict__ v,
                int kv_rs, float* __restrict__ out, int o_rs, int H, int Nq, int Nt, float scale) {
  PDL_ENTRY();
  extern __shared__ float sm[];  // k [Nt][H*16], v [Nt][H*16]
  const int C = H * T2I_DH;
  float* s_k = sm;
  float* s_v = sm + Nt * C;
  const int b = blockIdx.y;
  const int item = blockIdx.x * blockDim.x + threadIdx.x;  // (query, head)
  const bool active = item < Nq * H;
  const int qi = active ? item / H : 0, h = active ? item - qi * H : 0;
  // every global load of this thread -- its query row and its share of the token keys / values -- is issued before
  // anything waits: one memory round trip for the whole kernel
  float4 q4[4];
  const float4* qp = reinterpret_cast<const float4*>(q + ((long long)b * Nq + qi) * q_rs + h * T2I_DH);
#pragma unroll
  for (int i = 0; i < 4; ++i) q4[i] = active ? qp[i] : make_float4(0.f, 0.f, 0.f, 0.f);
  constexpr int STG = 8;  // staging slots per thread: Nt * C <= 16 * 128 elements, 256 threads
  float kk[STG], vv[STG];
#pragma unroll
  for (int u = 0; u < STG; ++u) {
    const int i = threadIdx.x + u * 256;
    kk[u] = vv[u] = 0.f;
    if (i < Nt * C) {
      const int t = i / C, cc = i - t * C;
      kk[u] = k[((long long)b * Nt + t) * kv_rs + cc];
      vv[u] = v[((long long)b * Nt + t) * kv_rs + cc];
    }
  }
#pragma unroll
  for (int u = 0; u < STG; ++u) {
    const int i = threadIdx.x + u * 256;
    if (i < Nt * C) {
      s_k[i] = kk[u];
      s_v[i] = vv[u];
    }
  }
  __syncthreads();
  if (!active) return;
  float qr[T2I_DH];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 t4 = q4[i];
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
  if (p->x2_cols % SK_COLS) return USVM_ERR_ARG;
  if (p->ln_w && (p->K != 256 || !p->ln_b)) return USVM_ERR_ARG;
  const size_t smem = ((size_t)SK_ROWS * p->K * 2 + SK_WARPS * 32) * sizeof(float);
  if (smem > 200 * 1024) return USVM_ERR_ARG;
  static UsvmPerDeviceOnce configured = {};
  if (usvm_need_setup(configured)) {
    if (cudaFuncSetAttribute(gemm_skinny_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess ||
        cudaFuncSetAttribute(gemm_skinny_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess)
      return USVM_ERR_CUDA;
    usvm_setup_done(configured);
  }
  if (p->K >= 1024) {  // long reduction: the CTA's 8 warps split K, one CTA per 4 columns
    dim3 grid(cdiv(p->N, SK_COLS), p->instances, cdiv(p->M, SK_ROWS));
    usvm_launch(gemm_skinny_kernel<8>, dim3(grid), dim3(256), smem, STREAM, *p);
  } else {
    dim3 grid(cdiv(p->N, SK_WARPS * SK_COLS), p->instances, cdiv(p->M, SK_ROWS));
    usvm_launch(gemm_skinny_kernel<1>, dim3(grid), dim3(256), smem, STREAM, *p);
  }
  return usvm_check_launch();
}

extern "C" int usvm_attn_t2i_f32(const float* q, int q_rs, const float* k, const float* v, int kv_rs, float* out,
                                 int o_rs, int B, int H, int Nt, int Nk, float scale, void* stream) {
  if (!q || !k || !v || !out || B <= 0 || H <= 0 || Nt <= 0 || Nt > T2I_MAX_NT || Nk <= 0 || (kv_rs % 4)) return USVM_ERR_ARG;
  const size_t smem = ((size_t)2 * Nk * T2I_LD + (size_t)Nt * Nk + Nt * T2I_DH + 16 * Nt * T2I_DH) * sizeof(float);
  if (smem > 200 * 1024) return USVM_ERR_ARG;
  static UsvmPerDeviceOnce configured = {};
  if (smem > 48 * 1024 && usvm_need_setup(configured)) {
    if (cudaFuncSetAttribute(attn_t2i_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess)
      return USVM_ERR_CUDA;
    usvm_setup_done(configured);
  }
  usvm_launch(attn_t2i_kernel, dim3(B * H), dim3(256), smem, STREAM, q, q_rs, k, v, kv_rs, out, o_rs, H, Nt, Nk, scale);
  return usvm_check_launch();
}

extern "C" int usvm_attn_t2i_split_f32(const float* q, int q_rs, const float* k, const float* v, int kv_rs, float* out,
                                       int o_rs, int B, int H, int Nt, int Nk, float scale, float* partials,
                                       int* counters, void* stream) {
  if (!q || !k || !v || !out || !partials || !counters || B <= 0 || H <= 0 || Nt <= 0 || Nt > T2I_MAX_NT || Nk <= 0 ||
      (kv_rs % 4) || (reinterpret_cast<uintptr_t>(k) & 15) || (reinterpret_cast<uintptr_t>(v) & 15))
    return USVM_ERR_ARG;
  const int S = cdiv(Nk, T2S_KEYS);
  if (S > 64) return USVM_ERR_ARG;
  usvm_launch(attn_t2i_split_kernel, dim3(B * H, S), dim3(T2S_KEYS), 0, STREAM, q, q_rs, k, v, kv_rs, out, o_rs, H, Nt, Nk,
              scale, partials, counters);
  return usvm_check_launch();
}

extern "C" int usvm_attn_i2t_f32(const float* q, int q_rs, const float* k, const float* v, int kv_rs, float* out,
                                 int o_rs, int B, int H, int Nq, int Nt, float scale, void* stream) {
  if (!q || !k || !v || !out || B <= 0 || H <= 0 || Nt <= 0 || Nt > T2I_MAX_NT || Nq <= 0 || (q_rs % 4) || (o_rs % 4))
    return USVM_ERR_ARG;
  if (Nt * H * T2I_DH > 8 * 256) return USVM_ERR_ARG;  // staging slots of the kernel
  const size_t smem = (size_t)2 * Nt * H * T2I_DH * sizeof(float);
  usvm_launch(attn_i2t_kernel, dim3(dim3(cdiv((long long)Nq * H, 256), B)), dim3(256), smem, STREAM, q, q_rs, k, v, kv_rs, out, o_rs, H, Nq,
                                                                                 Nt, scale);
  return usvm_check_launch();
}
