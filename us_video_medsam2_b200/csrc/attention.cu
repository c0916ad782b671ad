// Fused attention kernels.
//
//  usvm_fmha_bf16      flash attention, bf16 operands, fp32 softmax/accumulate, head_dim 96 or 256,
//                      arbitrary Nq / Nk, strided Q/K/V/O, optional split over the key range
//                      (flash-decoding) + usvm_fmha_combine.  Replaces F.scaled_dot_product_attention
//                      at hieradet.py:70 (windowed + global Hiera attention, head_dim 96) and
//                      sam/transformer.py:344/355 (RoPEAttention self- and cross-attention over the
//                      memory bank, one head of 256, up to 7232 keys).
//  usvm_attn_small_f32 fp32 attention for the SAM mask decoder's 8-token two-way transformer
//                      (sam/transformer.py:257-286; 8 heads of 32 or 16).
//
// The bf16 kernel is a register-resident (FlashAttention-2 style) mma.sync.m16n8k16 kernel: 64 queries x
// 64 keys per step, 4 warps x 16 query rows, K/V tiles double-buffered through cp.async.  It is the
// round-1 implementation; the tcgen05/TMEM variant (S and O accumulators in TMEM) is the next step
// recorded in DESIGN.md.
#include "common.cuh"
#include "usvm2_b200.h"

namespace {

constexpr int FM = 64;   // queries per CTA
constexpr int FN = 64;   // keys per step
constexpr int FTHREADS = 128;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
  const uint32_t d = smem_u32(smem_dst);
  const int sz = valid ? 16 : 0;  // src-size 0 => 16 bytes of zeros
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(p)));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int D>
struct FmhaSmem {
  static constexpr int LD = D + 8;  // padded row: ldmatrix rows land in distinct banks
  static constexpr int TILE = FM * LD;  // elements per 64-row tile
  static constexpr int BYTES = 5 * TILE * 2;  // Q + 2 x (K, V)
};

// copy a [64 x D] tile (rows row0.. of a strided matrix) into padded smem; rows >= nrows are zero-filled
template <int D>
__device__ __forceinline__ void load_tile(bf16* s, const bf16* g, int row_stride, int row0, int nrows) {
  constexpr int CPR = D / 8;  // 16-byte chunks per row
  for (int i = threadIdx.x; i < FM * CPR; i += FTHREADS) {
    const int r = i / CPR, c = i - r * CPR;
    const bool ok = row0 + r < nrows;
    const bf16* src = g + (long long)(ok ? row0 + r : 0) * row_stride + c * 8;
    cp_async16(s + r * FmhaSmem<D>::LD + c * 8, src, ok);
  }
}

template <int D>
__global__ void __launch_bounds__(FTHREADS)
fmha_bf16_kernel(const usvm_fmha_params p) {
  PDL_ENTRY();
  extern __shared__ __align__(16) uint8_t fsm[];
  using S = FmhaSmem<D>;
  bf16* sQ = reinterpret_cast<bf16*>(fsm);
  bf16* sK = sQ + S::TILE;       // two stages
  bf16* sV = sK + 2 * S::TILE;   // two stages
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int q0 = blockIdx.x * FM;
  const int bh = blockIdx.y;
  const int b = bh / p.H, h = bh - b * p.H;
  const int split = blockIdx.z;

  const bf16* Q = reinterpret_cast<const bf16*>(p.q) + (long long)b * p.q_bs + (long long)h * p.q_hs;
  const bf16* K = reinterpret_cast<const bf16*>(p.k) + (long long)b * p.k_bs + (long long)h * p.k_hs;
  const bf16* V = reinterpret_cast<const bf16*>(p.v) + (long long)b * p.v_bs + (long long)h * p.v_hs;

  // key range of this split (tile aligned)
  const int ntiles = (p.Nk + FN - 1) / FN;
  const int per = (ntiles + p.num_splits - 1) / p.num_splits;
  const int t_begin = split * per;
  const int t_end = min(ntiles, t_begin + per);

  load_tile<D>(sQ, Q, p.q_rs, q0, p.Nq);
  if (t_begin < t_end) {
    load_tile<D>(sK, K, p.k_rs, t_begin * FN, p.Nk);
    load_tile<D>(sV, V, p.v_rs, t_begin * FN, p.Nk);
  }
  cp_async_commit();

  constexpr int ND = D / 8;    // output n-tiles (8 channels each)
  constexpr int KS = D / 16;   // k-steps of Q.K^T
  float o[ND][4];
#pragma unroll
  for (int i = 0; i < ND; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY};
  float l_run[2] = {0.f, 0.f};
  const float sl2 = p.scale * 1.4426950408889634f;  // softmax in base 2

  for (int tile = t_begin; tile < t_end; ++tile) {
    const int st = (tile - t_begin) & 1;
    if (tile + 1 < t_end) {  // prefetch next K/V tile into the other stage
      load_tile<D>(sK + (st ^ 1) * S::TILE, K, p.k_rs, (tile + 1) * FN, p.Nk);
      load_tile<D>(sV + (st ^ 1) * S::TILE, V, p.v_rs, (tile + 1) * FN, p.Nk);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const bf16* cK = sK + st * S::TILE;
    const bf16* cV = sV + st * S::TILE;

    // ---- S = Q K^T (16 x 64 per warp) ----
    float s[FN / 8][4];
#pragma unroll
    for (int i = 0; i < FN / 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      uint32_t a[4];
      ldsm_x4(a, sQ + (warp * 16 + (lane & 15)) * S::LD + ks * 16 + (lane >> 4) * 8);
#pragma unroll
      for (int nt = 0; nt < FN / 8; nt += 2) {
        uint32_t kb[4];
        ldsm_x4(kb, cK + (nt * 8 + (lane & 7) + (lane >> 4) * 8) * S::LD + ks * 16 + ((lane >> 3) & 1) * 8);
        mma_bf16(s[nt], a, kb[0], kb[1]);
        mma_bf16(s[nt + 1], a, kb[2], kb[3]);
      }
    }
    // ---- mask keys beyond Nk, online softmax ----
    const int key0 = tile * FN;
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < FN / 8; ++nt) {
      const int kc = key0 + nt * 8 + 2 * t;
      if (kc >= p.Nk) s[nt][0] = s[nt][2] = -INFINITY;
      if (kc + 1 >= p.Nk) s[nt][1] = s[nt][3] = -INFINITY;
      mx[0] = fmaxf(mx[0], fmaxf(s[nt][0], s[nt][1]));
      mx[1] = fmaxf(mx[1], fmaxf(s[nt][2], s[nt][3]));
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
    }
    float corr[2], mnew[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mnew[r] = fmaxf(m_run[r], mx[r]);  // finite: every tile has at least one valid key
      corr[r] = exp2f((m_run[r] - mnew[r]) * sl2);
      m_run[r] = mnew[r];
      l_run[r] *= corr[r];
    }
    uint32_t pf[FN / 16][4];
    float rs[2] = {0.f, 0.f};
#pragma unroll
    for (int nt = 0; nt < FN / 8; ++nt) {
      const float p0 = exp2f((s[nt][0] - mnew[0]) * sl2);
      const float p1 = exp2f((s[nt][1] - mnew[0]) * sl2);
      const float p2 = exp2f((s[nt][2] - mnew[1]) * sl2);
      const float p3 = exp2f((s[nt][3] - mnew[1]) * sl2);
      rs[0] += p0 + p1;
      rs[1] += p2 + p3;
      pf[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16x2(p0, p1);
      pf[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16x2(p2, p3);
    }
    l_run[0] += rs[0];
    l_run[1] += rs[1];
#pragma unroll
    for (int i = 0; i < ND; ++i) {
      o[i][0] *= corr[0];
      o[i][1] *= corr[0];
      o[i][2] *= corr[1];
      o[i][3] *= corr[1];
    }
    // ---- O += P V ----
#pragma unroll
    for (int kk = 0; kk < FN / 16; ++kk) {
#pragma unroll
      for (int dt = 0; dt < ND; dt += 2) {
        uint32_t vb[4];
        ldsm_x4_t(vb, cV + (kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * S::LD + dt * 8 + (lane >> 4) * 8);
        mma_bf16(o[dt], pf[kk], vb[0], vb[1]);
        mma_bf16(o[dt + 1], pf[kk], vb[2], vb[3]);
      }
    }
    __syncthreads();  // everyone done with stage `st` before it is refilled
  }
  if (t_begin >= t_end) cp_async_wait<0>();

  // row sums live per quad: reduce across the 4 lanes that share a row
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
  }
  const int row_a = q0 + warp * 16 + g, row_b = row_a + 8;
  if (p.num_splits == 1) {
    bf16* O = reinterpret_cast<bf16*>(p.o) + (long long)b * p.o_bs + (long long)h * p.o_hs;
    const float inv0 = l_run[0] > 0.f ? 1.f / l_run[0] : 0.f;
    const float inv1 = l_run[1] > 0.f ? 1.f / l_run[1] : 0.f;
#pragma unroll
    for (int dt = 0; dt < ND; ++dt) {
      const int c = dt * 8 + 2 * t;
      if (row_a < p.Nq)
        *reinterpret_cast<uint32_t*>(O + (long long)row_a * p.o_rs + c) = pack_bf16x2(o[dt][0] * inv0, o[dt][1] * inv0);
      if (row_b < p.Nq)
        *reinterpret_cast<uint32_t*>(O + (long long)row_b * p.o_rs + c) = pack_bf16x2(o[dt][2] * inv1, o[dt][3] * inv1);
    }
  } else {
    // partial results: unnormalised O (fp32) + (m, l) per row;  layout [split][b*H + h][Nq][...]
    const long long prow = ((long long)split * gridDim.y + bh) * p.Nq;
    float* OP = p.o_part + prow * D;
    float* ML = p.ml_part + prow * 2;
#pragma unroll
    for (int dt = 0; dt < ND; ++dt) {
      const int c = dt * 8 + 2 * t;
      if (row_a < p.Nq) *reinterpret_cast<float2*>(OP + (long long)row_a * D + c) = make_float2(o[dt][0], o[dt][1]);
      if (row_b < p.Nq) *reinterpret_cast<float2*>(OP + (long long)row_b * D + c) = make_float2(o[dt][2], o[dt][3]);
    }
    if (t == 0) {
      if (row_a < p.Nq) *reinterpret_cast<float2*>(ML + (long long)row_a * 2) = make_float2(m_run[0], l_run[0]);
      if (row_b < p.Nq) *reinterpret_cast<float2*>(ML + (long long)row_b * 2) = make_float2(m_run[1], l_run[1]);
    }
  }
}

// merge split partials: one thread per (row, 4 channels).  The kernel is pure latency / bandwidth (up to 18 partials of
// 128 KB per query tile), so all (max, sum) pairs are fetched up front and the partial rows in register batches of 8:
// every load of a batch is in flight before the first one is consumed.
constexpr int COMBINE_MAX_SPLITS = 32;
template <int D, int NSMAX>
__global__ void __launch_bounds__(256, NSMAX <= 24 ? 2 : 1)
fmha_combine_kernel(const usvm_fmha_params p) {
  PDL_ENTRY();
  constexpr int C4 = D / 4;
  const long long total_rows = (long long)p.B * p.H * p.Nq;
  const long long item = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (item >= total_rows * C4) return;
  const long long row = item / C4;
  const int c4 = (int)(item - row * C4) * 4;
  const int bh = (int)(row / p.Nq);
  const int qi = (int)(row - (long long)bh * p.Nq);
  const int b = bh / p.H, h = bh - b * p.H;
  const float sl2 = p.scale * 1.4426950408889634f;
  const int ns = p.num_splits;
  const float2* mlp = reinterpret_cast<const float2*>(p.ml_part) + row;
  const float* op = p.o_part + row * D + c4;
  const long long ostride = total_rows * D;
  float2 ml[NSMAX];
#pragma unroll
  for (int s = 0; s < NSMAX; ++s) ml[s] = s < ns ? __ldg(mlp + (long long)s * total_rows) : make_float2(0.f, 0.f);
  float M = -INFINITY;
#pragma unroll
  for (int s = 0; s < NSMAX; ++s)
    if (ml[s].y > 0.f) M = fmaxf(M, ml[s].x);
  float L = 0.f;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (p.part_bf16) {
    // bf16 partials (8 bytes per thread and split): the first 24 splits are fetched in ONE batch -- the 18-way split of
    // the one-object cross-attention then costs a single L2 round trip instead of three -- the rest in a second one
    const bf16* opb = reinterpret_cast<const bf16*>(p.o_part) + row * D + c4;
    constexpr int B0 = NSMAX < 24 ? NSMAX : 24;
    uint2 raw[B0];
#pragma unroll
    for (int u = 0; u < B0; ++u)
      raw[u] = u < ns ? __ldg(reinterpret_cast<const uint2*>(opb + (long long)u * ostride)) : make_uint2(0u, 0u);
#pragma unroll
    for (int u = 0; u < B0; ++u) {
      const float2 m = ml[u];
      const float w = m.y > 0.f ? exp2f((m.x - M) * sl2) : 0.f;
      const float2 lo = unpack_bf16x2(raw[u].x), hi = unpack_bf16x2(raw[u].y);
      L += m.y * w;
      acc.x += lo.x * w; acc.y += lo.y * w; acc.z += hi.x * w; acc.w += hi.y * w;
    }
    if (NSMAX > 24 && ns > 24) {
#pragma unroll
      for (int u = 0; u < 8; ++u)
        raw[u] = 24 + u < ns ? __ldg(reinterpret_cast<const uint2*>(opb + (long long)(24 + u) * ostride)) : make_uint2(0u, 0u);
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const float2 m = ml[NSMAX > 24 ? 24 + u : 0];
        const float w = m.y > 0.f ? exp2f((m.x - M) * sl2) : 0.f;
        const float2 lo = unpack_bf16x2(raw[u].x), hi = unpack_bf16x2(raw[u].y);
        L += m.y * w;
        acc.x += lo.x * w; acc.y += lo.y * w; acc.z += hi.x * w; acc.w += hi.y * w;
      }
    }
  } else {
#pragma unroll
    for (int s0 = 0; s0 < NSMAX; s0 += 8) {
      if (s0 < ns) {
        float4 o[8];
#pragma unroll
        for (int u = 0; u < 8; ++u)
          o[u] = s0 + u < ns ? __ldg(reinterpret_cast<const float4*>(op + (long long)(s0 + u) * ostride))
                             : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const float2 m = ml[s0 + u];
          const float w = m.y > 0.f ? exp2f((m.x - M) * sl2) : 0.f;
          L += m.y * w;
          acc.x += o[u].x * w; acc.y += o[u].y * w; acc.z += o[u].z * w; acc.w += o[u].w * w;
        }
      }
    }
  }
  const float inv = L > 0.f ? 1.f / L : 0.f;
  bf16* O = reinterpret_cast<bf16*>(p.o) + (long long)b * p.o_bs + (long long)h * p.o_hs + (long long)qi * p.o_rs + c4;
  uint2 pk;
  pk.x = pack_bf16x2(acc.x * inv, acc.y * inv);
  pk.y = pack_bf16x2(acc.z * inv, acc.w * inv);
  *reinterpret_cast<uint2*>(O) = pk;
}

template <int D>
void launch_combine(const usvm_fmha_params* p, long long rows, cudaStream_t s) {
  const dim3 grid(cdiv(rows * (D / 4), 256));
  if (p->num_splits <= 8) usvm_launch(fmha_combine_kernel<D, 8>, grid, dim3(256), 0, s, *p);
  else if (p->num_splits <= 24) usvm_launch(fmha_combine_kernel<D, 24>, grid, dim3(256), 0, s, *p);
  else usvm_launch(fmha_combine_kernel<D, COMBINE_MAX_SPLITS>, grid, dim3(256), 0, s, *p);
}

template <int D>
int launch_fmha(const usvm_fmha_params* p, cudaStream_t s) {
  static UsvmPerDeviceOnce attr = {};
  if (usvm_need_setup(attr)) {
    if (cudaFuncSetAttribute(fmha_bf16_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, FmhaSmem<D>::BYTES) !=
        cudaSuccess)
      return USVM_ERR_CUDA;
    usvm_setup_done(attr);
  }
  dim3 grid(cdiv(p->Nq, FM), p->B * p->H, p->num_splits);
  usvm_launch(fmha_bf16_kernel<D>, dim3(grid), dim3(FTHREADS), FmhaSmem<D>::BYTES, s, *p);
  if (p->num_splits > 1) {
    const long long rows = (long long)p->B * p->H * p->Nq;
    launch_combine<D>(p, rows, s);
  }
  return usvm_check_launch();
}

// ---------------------------------------------------------------------------------------------
// fp32 small attention: one warp per (batch, head, query)
// ---------------------------------------------------------------------------------------------
constexpr int SMALL_MAX_NK = 1024;

__global__ void __launch_bounds__(128)
attn_small_f32_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                      float* __restrict__ out, int B, int H, int Nq, int Nk, int dh, int q_rs, int k_rs, int v_rs,
                      int o_rs, float scale) {
  PDL_ENTRY();
  __shared__ float s_scores[4][SMALL_MAX_NK];
  __shared__ float s_q[4][32];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long item = (long long)blockIdx.x * 4 + w;
  const long long total = (long long)B * H * Nq;
  if (item >= total) return;  // whole warp exits together
  const int qi = (int)(item % Nq);
  const int h = (int)((item / Nq) % H);
  const int b = (int)(item / ((long long)Nq * H));
  const float* qp = q + ((long long)b * Nq + qi) * q_rs + h * dh;
  const float* kp = k + (long long)b * Nk * k_rs + h * dh;
  const float* vp = v + (long long)b * Nk * v_rs + h * dh;
  if (lane < dh) s_q[w][lane] = qp[lane] * scale;
  __syncwarp();
  float mx = -INFINITY;
  for (int j = lane; j < Nk; j += 32) {
    const float* kr = kp + (long long)j * k_rs;
    float d = 0.f;
    for (int c = 0; c < dh; ++c) d = fmaf(s_q[w][c], kr[c], d);
    s_scores[w][j] = d;
    mx = fmaxf(mx, d);
  }
  mx = warp_max(mx);
  float sum = 0.f;
  for (int j = lane; j < Nk; j += 32) {
    const float e = __expf(s_scores[w][j] - mx);
    s_scores[w][j] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  __syncwarp();
  // lanes own channels; 32 / dh key groups run in parallel and are reduced by shuffles (dh = 16 or 32)
  const int groups = 32 / dh;
  const int c = lane % dh, grp = lane / dh;
  float acc = 0.f;
  if (grp < groups)
    for (int j = grp; j < Nk; j += groups) acc = fmaf(s_scores[w][j], vp[(long long)j * v_rs + c], acc);
  for (int o = dh; o < 32; o <<= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane < dh) out[((long long)b * Nq + qi) * o_rs + h * dh + lane] = acc / sum;
}

}  // namespace

extern "C" int usvm_fmha_bf16(const usvm_fmha_params* p, void* stream) {
  if (!p || !p->q || !p->k || !p->v || !p->o || p->B <= 0 || p->H <= 0 || p->Nq <= 0 || p->Nk <= 0)
    return USVM_ERR_ARG;
  if (p->num_splits < 1 || p->num_splits > COMBINE_MAX_SPLITS || p->part_bf16) return USVM_ERR_ARG;
  if (p->num_splits > 1 && (!p->o_part || !p->ml_part)) return USVM_ERR_ARG;
  if (p->num_splits > cdiv(p->Nk, FN)) return USVM_ERR_ARG;
  if ((p->q_rs % 8) || (p->k_rs % 8) || (p->v_rs % 8) || (p->o_rs % 2) || (p->q_hs % 8) || (p->k_hs % 8) ||
      (p->v_hs % 8) || (p->q_bs % 8) || (p->k_bs % 8) || (p->v_bs % 8))
    return USVM_ERR_ARG;  // 16-byte cp.async chunks
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (p->head_dim == 64) return launch_fmha<64>(p, s);
  if (p->head_dim == 96) return launch_fmha<96>(p, s);
  if (p->head_dim == 256) return launch_fmha<256>(p, s);
  return USVM_ERR_ARG;
}

// merge the split-KV partials written by usvm_fmha_bf16 / usvm_fmha_tc5 (head_dim 96 or 256)
extern "C" int usvm_fmha_combine(const usvm_fmha_params* p, void* stream) {
  if (!p || !p->o || !p->o_part || !p->ml_part || p->num_splits < 1 || p->num_splits > COMBINE_MAX_SPLITS || p->B <= 0 || p->H <= 0 || p->Nq <= 0)
    return USVM_ERR_ARG;
  const long long rows = (long long)p->B * p->H * p->Nq;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if ((p->o_rs % 4) || (p->o_hs % 4) || (p->o_bs % 4)) return USVM_ERR_ARG;
  if (p->head_dim == 64) launch_combine<64>(p, rows, s);
  else if (p->head_dim == 96) launch_combine<96>(p, rows, s);
  else if (p->head_dim == 256) launch_combine<256>(p, rows, s);
  else return USVM_ERR_ARG;
  return usvm_check_launch();
}

extern "C" int usvm_attn_small_f32(const float* q, const float* k, const float* v, float* out, int B, int H, int Nq,
                                   int Nk, int head_dim, int q_rs, int k_rs, int v_rs, int o_rs, float scale,
                                   void* stream) {
  if (!q || !k || !v || !out || B <= 0 || H <= 0 || Nq <= 0 || Nk <= 0) return USVM_ERR_ARG;
  if (Nk > SMALL_MAX_NK || (head_dim != 16 && head_dim != 32)) return USVM_ERR_ARG;
  const long long total = (long long)B * H * Nq;
  usvm_launch(attn_small_f32_kernel, dim3(cdiv(total, 4)), dim3(128), 0, reinterpret_cast<cudaStream_t>(stream), 
      q, k, v, out, B, H, Nq, Nk, head_dim, q_rs, k_rs, v_rs, o_rs, scale);
  return usvm_check_launch();
}
