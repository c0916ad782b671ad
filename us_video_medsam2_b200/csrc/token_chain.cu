// Token side of the SAM mask decoder's two-way transformer as ONE kernel per dependency chain
// (sam/transformer.py:137-212 TwoWayAttentionBlock, mask_decoder.py:215-253 hyper-network / IoU / score heads,
// sam2_base.py:1143-1156 obj_ptr_proj).
//
// A tracked frame carries 8 decoder tokens per object.  Every token-side layer is a [8 x K] x [N x K]^T product whose
// cost is streaming the fp32 weight matrix once; as separate launches each of the ~38 layers paid a full kernel
// boundary (~3-4 us) for well under 1 us of work.  Here a thread-block CLUSTER (8 or 16 CTAs, one cluster per object)
// runs a whole chain of such layers:
//   * the cluster's CTAs split the weight rows (= output columns) of each step, so the weight stream is spread over
//     8-16 SMs; a producer warp per CTA streams the rows through a 4 x 32 KB shared-memory ring with bulk async copies
//     and runs ahead of the consumers (weights never depend on earlier steps or kernels);
//   * products run on the tensor cores as 3 x tf32 (mma.sync m16n8k8 on hi / lo splits): fp32-level accuracy with
//     ~20x fewer issue slots than CUDA-core FMAs -- the chain is bound by single-warp instruction latency, not by
//     bandwidth, so instruction count is what matters; for the same reason all index arithmetic is division-free and
//     the per-step schedule (rows per CTA, chunk shape, k-slicing) is computed on the host;
//   * step outputs go to small global buffers and steps are separated by a release/acquire cluster barrier instead of
//     a kernel boundary;
//   * input-side transforms are recomputed per CTA because they are tiny: LayerNorm of the 8 rows (in registers, one
//     warp per row), adding the positional tokens, the 8 x 8 token self-attention, the merge of the per-CTA
//     token->image attention partials.
//
// Step kinds:
//   LINEAR       out[m, n] = act((T(x)[m] (+ x2[m] for n < x2_cols)) . W[n] + b[n]) (+ residual[m, n])
//                T = identity | LayerNorm (optionally also written out) | self-attention over the q|k|v columns of x
//                  | softmax-merge of the T2I partials; w_is != 0 selects "row m uses matrix m" (stacked heads).
//   T2I_PARTIAL  token->image attention (8 heads x 16): CTA r of the cluster handles keys [r*Nk/CL, (r+1)*Nk/CL) of
//                all heads and writes unnormalised (max, sum, P.V) partials to the scratch buffer.
#include "common.cuh"
#include "usvm2_b200.h"

namespace {

constexpr int TC_THREADS = 256;  // consumer threads (8 warps); warp 8 is the weight-stream producer
constexpr int TC_WARPS = 8;
constexpr int TC_BLOCK = TC_THREADS + 32;
constexpr int TC_RING = 4;                 // weight ring: 4 x 32 KB bulk-copy chunks in flight per SM
constexpr int TC_CHUNK_BYTES = 32 * 1024;
constexpr int TC_CHUNK_FLOATS = TC_CHUNK_BYTES / 4;
constexpr int TC_ROWS = USVM_CHAIN_ROWS;  // 8
constexpr int TC_KMAX = 2048;
constexpr int TC_X2_FLOATS = TC_ROWS * 768;  // positional-token copy of the input (K <= 768) / q|k|v staging
constexpr int TC_HEADS = 8;
constexpr int TC_T2I_DH = 16;
constexpr int TC_T2I_C = TC_HEADS * TC_T2I_DH;  // 128
constexpr int TC_KEYS = 128;                    // keys per CTA in a T2I_PARTIAL step (Nk <= 128 * cluster)
constexpr int TC_RED_FLOATS = 2 * TC_WARPS * 128;  // double-buffered k-slice partial accumulators (4 per lane)
constexpr size_t TC_SMEM = (size_t)TC_RING * TC_CHUNK_BYTES +
                           (size_t)(TC_ROWS * TC_KMAX + TC_X2_FLOATS + TC_RED_FLOATS) * sizeof(float) +
                           2 * TC_RING * sizeof(uint64_t) + 64;

// Device-side step: the public descriptor narrowed to 32-bit strides plus the host-computed schedule.
struct KStep {
  const float *x, *ln_w, *ln_b, *x2, *w, *bias, *residual, *k, *v;
  float *ln_out, *out;
  const int* row_select;
  int x_os, x_rs, sel_stride, ln_os, ln_rs, x2_os, x2_rs, b_is, r_os, r_rs, o_os, o_rs, kv_os, kv_rs;
  int kind, in_kind, rows, N, K, act, x2_cols, Nk, attn_q, attn_k, attn_v, stacked;
  float ln_eps;
  // schedule (LINEAR): weight rows [rank*per, +per) per CTA, chunks of rb rows x kw columns, nkc k-chunks per row block
  int total_rows, per, rb_shift, kw, nkc, ng_shift;  // ng = (rb/16) = 1 << ng_shift row groups, 8 >> ng_shift k-slices
  int n_shift;                                       // stacked: N = 1 << n_shift
  int keys_per;                                      // T2I_PARTIAL: keys per CTA
};
struct KParams {
  float* scratch;
  unsigned long long* timing;
  int n_steps, n_obj, cluster, precise;
  KStep steps[USVM_CHAIN_MAX_STEPS];
};

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
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
// barrier among the 256 consumer threads only (the producer warp runs ahead on its own)
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// fp32 -> (hi, lo) tf32 pair: hi + lo carries ~21 mantissa bits, so three tensor-core products (lo*hi, hi*lo, hi*hi)
// reproduce the fp32 product to ~1e-6 relative -- the token side of the decoder keeps fp32-level accuracy
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
  const float r = x - __uint_as_float(hi);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(r));
}
__device__ __forceinline__ void mma_tf32_16x8x8(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                                uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// ---- input transforms: fill xs[8][K] (and xs2 = xs + x2) -----------------------------------------------------
// NOTE on all loaders: a global load cannot be hoisted above an earlier store the compiler cannot prove disjoint, so
// every loop issues its whole batch of loads into registers first and stores afterwards.
// Warp m owns token row m: lane l holds columns 128 j + 4 l .. + 3.  LayerNorm (K <= 512) runs on those registers.
__device__ __forceinline__ void load_rows(const KStep& st, int obj, int rank, float* xs, float* xs2) {
  const int K = st.K, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m = warp;
  const bool live = m < st.rows;
  const float* xr = st.x + obj * st.x_os + (st.row_select ? st.row_select[obj] * st.sel_stride : 0) + m * st.x_rs;
  const float* x2r = st.x2 ? st.x2 + obj * st.x2_os + m * st.x2_rs : nullptr;
  float* xd = xs + m * K;
  float* x2d = xs2 + m * K;
  const bool ln = st.ln_w != nullptr;
  for (int kb = lane * 4; kb < K; kb += 512) {  // batches of four float4 per lane (one batch when K <= 512)
    float4 v[4], a[4], lw[4], lb[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int k = kb + 128 * u;
      v[u] = a[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k < K) {
        if (live) {
          v[u] = ldcg4(xr + k);
          if (x2r) a[u] = ldcg4(x2r + k);
        }
        if (ln) {
          lw[u] = ldg4(st.ln_w + k);
          lb[u] = ldg4(st.ln_b + k);
        }
      }
    }
    if (ln) {  // the whole row is in this batch (K <= 512, checked on the host); two-pass statistics
      float s = 0.f;
#pragma unroll
      for (int u = 0; u < 4; ++u) s += (v[u].x + v[u].y) + (v[u].z + v[u].w);  // absent columns hold zeros
      const float mean = warp_sum(s) / K;
      float q = 0.f;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (kb + 128 * u < K) {
          v[u].x -= mean; v[u].y -= mean; v[u].z -= mean; v[u].w -= mean;
          q = fmaf(v[u].x, v[u].x, q); q = fmaf(v[u].y, v[u].y, q);
          q = fmaf(v[u].z, v[u].z, q); q = fmaf(v[u].w, v[u].w, q);
        }
      }
      const float rstd = 1.0f / sqrtf(warp_sum(q) / K + st.ln_eps);
      float* lo = (st.ln_out && rank == 0 && live) ? st.ln_out + obj * st.ln_os + m * st.ln_rs : nullptr;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int k = kb + 128 * u;
        if (k < K) {
          v[u].x = v[u].x * rstd * lw[u].x + lb[u].x; v[u].y = v[u].y * rstd * lw[u].y + lb[u].y;
          v[u].z = v[u].z * rstd * lw[u].z + lb[u].z; v[u].w = v[u].w * rstd * lw[u].w + lb[u].w;
          if (!live) v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (lo) *reinterpret_cast<float4*>(lo + k) = v[u];
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int k = kb + 128 * u;
      if (k < K) {
        *reinterpret_cast<float4*>(xd + k) = v[u];
        if (x2r) *reinterpret_cast<float4*>(x2d + k) = make_float4(v[u].x + a[u].x, v[u].y + a[u].y, v[u].z + a[u].z, v[u].w + a[u].w);
      }
    }
  }
  consumer_sync();
}

// self-attention of the rows over themselves: x rows hold q | k | v at column offsets attn_q/k/v, 8 heads x 32
__device__ __forceinline__ void load_self_attention(const KStep& st, int obj, float* xs, float* qkv, float* sc) {
  constexpr int C = 256, DH = 32;
  const int tid = threadIdx.x, rows = st.rows;
  const float* xb = st.x + obj * st.x_os;
  {  // stage q | k | v of every row: 8 * 3 * 64 float4 = 6 per thread, all loads first
    float4 v[6];
#pragma unroll
    for (int u = 0; u < 6; ++u) {
      const int i = tid + u * TC_THREADS;
      const int m = i / (3 * C / 4), r = i - m * (3 * C / 4), part = r / (C / 4), k = (r - part * (C / 4)) << 2;
      const int off = part == 0 ? st.attn_q : part == 1 ? st.attn_k : st.attn_v;
      v[u] = m < rows ? ldcg4(xb + m * st.x_rs + off + k) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < 6; ++u) *reinterpret_cast<float4*>(qkv + 4 * (tid + u * TC_THREADS)) = v[u];  // [(m*3+part)*256 + k]
  }
  consumer_sync();
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
  consumer_sync();
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
  consumer_sync();
#pragma unroll
  for (int u = 0; u < TC_ROWS; ++u) {  // out[query u][col = tid], col = h * 32 + c
    const int col = tid, h = col >> 5;
    float o = 0.f;
    if (u < rows) {
      const float* pr = sc + (h * 8 + u) * 8;
#pragma unroll
      for (int j = 0; j < TC_ROWS; ++j)
        if (j < rows) o = fmaf(pr[j], qkv[(j * 3 + 2) * C + col], o);
    }
    xs[u * C + col] = o;
  }
  consumer_sync();
}

// merge of the cluster's token->image partials: xs[row][128] = sum_r e^{m_r - M} o_r / sum_r e^{m_r - M} l_r
// thread <-> (row, head, 4 of the head's 16 channels): every load of the thread is issued before the first use
__device__ __forceinline__ void load_t2i_merge(const KStep& st, const float* scratch, int cl, float* xs) {
  const float* po = scratch;                                     // [cl][8][128]
  const float* ml = scratch + cl * (TC_ROWS * TC_T2I_C);         // [cl][8][8][2]
  const int tid = threadIdx.x;
  const int row = tid >> 5, h = (tid >> 2) & 7, q4 = (tid & 3) << 2;
  float2 mlv[16];
  float4 ov[16];
#pragma unroll
  for (int r = 0; r < 16; ++r) {
    if (r < cl) {
      mlv[r] = __ldcg(reinterpret_cast<const float2*>(ml + ((r * TC_ROWS + row) * TC_HEADS + h) * 2));
      ov[r] = ldcg4(po + (r * TC_ROWS + row) * TC_T2I_C + h * TC_T2I_DH + q4);
    }
  }
  float M = -INFINITY;
#pragma unroll
  for (int r = 0; r < 16; ++r)
    if (r < cl) M = fmaxf(M, mlv[r].x);
  float L = 0.f;
  float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int r = 0; r < 16; ++r) {
    if (r < cl) {
      const float wgt = __expf(mlv[r].x - M);
      L = fmaf(mlv[r].y, wgt, L);
      o.x = fmaf(ov[r].x, wgt, o.x); o.y = fmaf(ov[r].y, wgt, o.y);
      o.z = fmaf(ov[r].z, wgt, o.z); o.w = fmaf(ov[r].w, wgt, o.w);
    }
  }
  const float inv = row < st.rows ? 1.0f / L : 0.f;
  *reinterpret_cast<float4*>(xs + row * TC_T2I_C + h * TC_T2I_DH + q4) = make_float4(o.x * inv, o.y * inv, o.z * inv, o.w * inv);
  consumer_sync();
}

// ---- one LINEAR step: this CTA's share of the output columns -------------------------------------------------------
// out^T[n][m] = sum_k W[n][k] x[m][k] on mma.sync m16n8k8 tf32: A = 16 weight rows from the ring chunk, B = the 8 token
// rows from xs.  Each lane fetches float4s, i.e. four consecutive k: MMA j of a 32-wide k-block uses element j of every
// lane's float4 -- a fixed permutation of k inside the block that A and B share, so no shuffles are needed.
// PRECISE: three products on (hi, lo) tf32 splits (fp32-level accuracy); otherwise one product on round-to-nearest tf32.
template <bool PRECISE>
__device__ __forceinline__ void mma_kblock(float (&acc)[4], const float* wrow, const float* wrow8, const float* xk) {
  const float4 a_lo0 = *reinterpret_cast<const float4*>(wrow);
  const float4 a_hi0 = *reinterpret_cast<const float4*>(wrow + 16);
  const float4 a_lo8 = *reinterpret_cast<const float4*>(wrow8);
  const float4 a_hi8 = *reinterpret_cast<const float4*>(wrow8 + 16);
  const float4 x_lo = *reinterpret_cast<const float4*>(xk);
  const float4 x_hi = *reinterpret_cast<const float4*>(xk + 16);
  const float av[4][4] = {{a_lo0.x, a_lo8.x, a_hi0.x, a_hi8.x}, {a_lo0.y, a_lo8.y, a_hi0.y, a_hi8.y},
                          {a_lo0.z, a_lo8.z, a_hi0.z, a_hi8.z}, {a_lo0.w, a_lo8.w, a_hi0.w, a_hi8.w}};
  const float bv[4][2] = {{x_lo.x, x_hi.x}, {x_lo.y, x_hi.y}, {x_lo.z, x_hi.z}, {x_lo.w, x_hi.w}};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint32_t ah[4], al[4], bh[2], bl[2];
    if (PRECISE) {
#pragma unroll
      for (int i = 0; i < 4; ++i) split_tf32(av[j][i], ah[i], al[i]);
#pragma unroll
      for (int i = 0; i < 2; ++i) split_tf32(bv[j][i], bh[i], bl[i]);
      mma_tf32_16x8x8(acc, al[0], al[1], al[2], al[3], bh[0], bh[1]);
      mma_tf32_16x8x8(acc, ah[0], ah[1], ah[2], ah[3], bl[0], bl[1]);
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(ah[i]) : "f"(av[j][i]));
#pragma unroll
      for (int i = 0; i < 2; ++i) asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(bh[i]) : "f"(bv[j][i]));
    }
    mma_tf32_16x8x8(acc, ah[0], ah[1], ah[2], ah[3], bh[0], bh[1]);
  }
}

// Work split.  K <= 512 (one k-chunk per row block): every 16-row group of a chunk belongs to ONE warp, round-robin over
// the warps, which runs the whole reduction and the epilogue on its own -- no cross-warp traffic, and the warps work on
// different ring slots at the same time.  K > 512: the 16-row group of a chunk is cut into 8 k-slices, one per warp;
// the partial accumulators meet in shared memory after the last k-chunk.
template <bool PRECISE>
__device__ __forceinline__ void linear_step(const KStep& st, int obj, int rank, const float* xs, const float* xs2,
                                            const float* ring, uint64_t* full_bar, uint64_t* empty_bar, float* red,
                                            uint32_t& it) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int K = st.K, rows = st.rows, kw = st.kw, nkc = st.nkc;
  const int ng = 1 << st.ng_shift;
  const int r_lo = min(st.total_rows, rank * st.per), r_hi = min(st.total_rows, r_lo + st.per);
  const int rb = 1 << st.rb_shift;
  const bool stacked = st.stacked != 0;
  const float* bias = st.bias;
  const float* resid = st.residual ? st.residual + obj * st.r_os : nullptr;
  float* outp = st.out + obj * st.o_os;
  const int r_rs = st.r_rs, o_rs = st.o_rs, act = st.act, b_is = st.b_is;
  const int x2_cols = st.x2 ? st.x2_cols : 0;

  // bias / residual of this lane's four outputs of a group: (row g, token 2t), (g, 2t+1), (g+8, 2t), (g+8, 2t+1)
  auto epilogue = [&](const float (&acc)[4], const float (&bpre)[4], const float (&rpre)[4], const bool (&mine)[4], int n0) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (mine[e]) {
        float o = acc[e] + bpre[e];
        if (act == USVM_ACT_RELU) o = fmaxf(o, 0.f);
        else if (act == USVM_ACT_GELU) o = gelu_erf(o);
        outp[(2 * t + (e & 1)) * o_rs + n0 + g + (e >> 1) * 8] = o + rpre[e];
      }
    }
  };
  auto prefetch = [&](float (&bpre)[4], float (&rpre)[4], bool (&mine)[4], int grp, int nrows, int inst, int n0, bool owner) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int r = grp * 16 + g + (e >> 1) * 8;
      const int m = 2 * t + (e & 1);
      mine[e] = owner && r < nrows && (stacked ? (m == inst) : (m < rows));
      bpre[e] = rpre[e] = 0.f;
      if (mine[e]) {
        const int n = n0 + g + (e >> 1) * 8;
        if (bias) bpre[e] = __ldg(bias + inst * b_is + n);
        if (resid) rpre[e] = __ldcg(resid + m * r_rs + n);
      }
    }
  };

  if (nkc == 1) {
    int q_base = 0;  // (row block, group) pairs are dealt to the warps round-robin
    for (int row0 = r_lo; row0 < r_hi; row0 += rb, q_base += ng, ++it) {
      const uint32_t slot = it & (TC_RING - 1), ph = (it / TC_RING) & 1u;
      const int nrows = min(rb, r_hi - row0);
      const int grp = (warp - q_base) & (TC_WARPS - 1);  // the group of this chunk that falls to this warp, if < ng
      const bool owner = grp < ng && grp * 16 < nrows;
      const int wrow = row0 + grp * 16;
      const int inst = stacked ? wrow >> st.n_shift : 0;
      const int n0 = stacked ? wrow - (inst << st.n_shift) : wrow;
      float bpre[4], rpre[4], acc[4] = {0.f, 0.f, 0.f, 0.f};
      bool mine[4];
      prefetch(bpre, rpre, mine, grp, nrows, inst, n0, owner);
      mbar_wait(&full_bar[slot], ph);
      if (owner) {
        const float* wr = ring + slot * TC_CHUNK_FLOATS + (grp * 16 + g) * K + 4 * t;
        const float* xk = (n0 < x2_cols ? xs2 : xs) + g * K + 4 * t;
#pragma unroll 2
        for (int kb = 0; kb < K; kb += 32) mma_kblock<PRECISE>(acc, wr + kb, wr + 8 * K + kb, xk + kb);
        epilogue(acc, bpre, rpre, mine, n0);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[slot]);  // this warp is done with the ring slot
    }
    return;
  }

  const int nsl = TC_WARPS >> st.ng_shift;
  const int group = warp & (ng - 1), slice = warp >> st.ng_shift;
  int rbk = 0;
  for (int row0 = r_lo; row0 < r_hi; row0 += rb, ++rbk) {
    const int nrows = min(rb, r_hi - row0);
    const bool active = group * 16 < nrows;
    const int wrow = row0 + group * 16;                  // first weight row of this warp's group
    const int inst = stacked ? wrow >> st.n_shift : 0;
    const int n0 = stacked ? wrow - (inst << st.n_shift) : wrow;  // output column of that row
    const float* xin = (n0 < x2_cols ? xs2 : xs) + g * K + 4 * t;
    float bpre[4], rpre[4], acc[4] = {0.f, 0.f, 0.f, 0.f};
    bool mine[4];
    prefetch(bpre, rpre, mine, group, nrows, inst, n0, slice == 0);
    for (int kc = 0; kc < nkc; ++kc, ++it) {
      const uint32_t slot = it & (TC_RING - 1), ph = (it / TC_RING) & 1u;
      const int kbase = kc * kw;
      const int kwid = min(kw, K - kbase);
      const float* wch = ring + slot * TC_CHUNK_FLOATS + (group * 16 + g) * kwid + 4 * t;
      mbar_wait(&full_bar[slot], ph);
      if (active)
        for (int kb = slice * 32; kb < kwid; kb += nsl * 32)
          mma_kblock<PRECISE>(acc, wch + kb, wch + 8 * kwid + kb, xin + kbase + kb);
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[slot]);
    }
    // ---- meet the k-slices, then bias / activation / residual / store by the slice-0 warp of each group
    float* rbuf = red + (rbk & 1) * (TC_WARPS * 128);
    *reinterpret_cast<float4*>(rbuf + (warp * 32 + lane) * 4) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    consumer_sync();
    if (slice == 0 && active) {
      for (int sl = 1; sl < nsl; ++sl) {
        const float4 o = *reinterpret_cast<const float4*>(rbuf + (((sl << st.ng_shift) + group) * 32 + lane) * 4);
        acc[0] += o.x; acc[1] += o.y; acc[2] += o.z; acc[3] += o.w;
      }
      epilogue(acc, bpre, rpre, mine, n0);
    }
  }
}

// ---- one T2I_PARTIAL step -----------------------------------------------------------------------------------------
__device__ __forceinline__ void t2i_partial_step(const KStep& st, int obj, int rank, int cl, float* scratch, float* sc,
                                                 float* sq) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, rows = st.rows;
  const int key0 = rank * st.keys_per;
  const int nkeys = max(0, min(st.keys_per, st.Nk - key0));
  const int kv_rs = st.kv_rs;
  const float* xb = st.x + obj * st.x_os;
  const float* kb = st.k + obj * st.kv_os + key0 * kv_rs;
  const float* vb = st.v + obj * st.kv_os + key0 * kv_rs;
  const int j = tid & (TC_KEYS - 1), hh = tid >> 7;
  // this thread's key row (4 heads x 16) and its share of the queries travel together
  float4 kr4[16];
  if (j < nkeys) {
#pragma unroll
    for (int i = 0; i < 16; ++i) kr4[i] = ldg4(kb + j * kv_rs + hh * 64 + i * 4);
  }
  {
    float qv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = tid + u * TC_THREADS;
      const int m = i >> 7, c = i & 127;
      qv[u] = m < rows ? __ldcg(xb + m * st.x_rs + c) * 0.25f : 0.f;  // 1 / sqrt(16)
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) sq[tid + u * TC_THREADS] = qv[u];
  }
  consumer_sync();
  // scores: thread <-> (key, half of the heads); sc[(row * 8 + head) * 128 + key]
  if (j < nkeys) {
#pragma unroll
    for (int h4 = 0; h4 < 4; ++h4) {
      const int h = hh * 4 + h4;
#pragma unroll
      for (int m = 0; m < TC_ROWS; ++m) {
        const float* qm = sq + m * TC_T2I_C + h * TC_T2I_DH;
        float d = 0.f;
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          const float4 kk = kr4[h4 * 4 + c4];
          const float4 qq = *reinterpret_cast<const float4*>(qm + c4 * 4);
          d = fmaf(qq.x, kk.x, d); d = fmaf(qq.y, kk.y, d);
          d = fmaf(qq.z, kk.z, d); d = fmaf(qq.w, kk.w, d);
        }
        sc[(m * TC_HEADS + h) * TC_KEYS + j] = d;
      }
    }
  } else {
#pragma unroll
    for (int h4 = 0; h4 < 4; ++h4)
#pragma unroll
      for (int m = 0; m < TC_ROWS; ++m) sc[(m * TC_HEADS + hh * 4 + h4) * TC_KEYS + j] = -INFINITY;
  }
  consumer_sync();
  float* po = scratch + rank * (TC_ROWS * TC_T2I_C);
  float* ml = scratch + cl * (TC_ROWS * TC_T2I_C) + rank * (TC_ROWS * TC_HEADS * 2);
#pragma unroll
  for (int i = 0; i < TC_ROWS; ++i) {  // (row, head) pairs: local max / exp / sum; 8 pairs per warp
    const int pr = warp * TC_ROWS + i;
    float* s = sc + pr * TC_KEYS;
    float e[4], mx = -INFINITY;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      e[u] = s[lane + 32 * u];
      mx = fmaxf(mx, e[u]);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    if (mx == -INFINITY) {  // this CTA holds no keys
#pragma unroll
      for (int u = 0; u < 4; ++u) s[lane + 32 * u] = 0.f;
    } else {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        e[u] = __expf(e[u] - mx);
        s[lane + 32 * u] = e[u];
        sum += e[u];
      }
    }
    sum = warp_sum(sum);
    if (lane == 0) *reinterpret_cast<float2*>(ml + pr * 2) = make_float2(mx, sum);
  }
  consumer_sync();
  {  // P.V: thread <-> (channel, half of the rows); value loads in register batches of 32 (one L2 trip per batch)
    const int ch = tid & (TC_T2I_C - 1), rh = tid >> 7, h = ch / TC_T2I_DH;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const float* scr = sc + (rh * 4 * TC_HEADS + h) * TC_KEYS;
    for (int j0 = 0; j0 < nkeys; j0 += 32) {
      float vv[32];
#pragma unroll
      for (int u = 0; u < 32; ++u) vv[u] = (j0 + u < nkeys) ? __ldg(vb + (j0 + u) * kv_rs + ch) : 0.f;
#pragma unroll
      for (int u = 0; u < 32; ++u) {
        const int jj = min(j0 + u, TC_KEYS - 1);
#pragma unroll
        for (int r = 0; r < 4; ++r) acc[r] = fmaf(scr[r * TC_HEADS * TC_KEYS + jj], vv[u], acc[r]);
      }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) po[(rh * 4 + r) * TC_T2I_C + ch] = acc[r];
  }
}

__global__ void __launch_bounds__(TC_BLOCK, 1) token_chain_kernel(const __grid_constant__ KParams p) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  float* ring = reinterpret_cast<float*>(smem_raw);                 // [TC_RING][32 KB] weight chunks (bulk copies)
  float* xs = ring + TC_RING * TC_CHUNK_FLOATS;                     // [8][K]; also the score buffer of the attention steps
  float* xs2 = xs + TC_ROWS * TC_KMAX;                              // [8][K <= 768] input + positional tokens; q|k|v staging
  float* sq = xs2;                                                  // [8][128] scaled queries of T2I_PARTIAL (xs2 is idle then)
  float* red = xs2 + TC_X2_FLOATS;                                  // k-slice partial accumulators
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(red + TC_RED_FLOATS);
  uint64_t* empty_bar = full_bar + TC_RING;
  const int cl = p.cluster;
  const int rank = (int)cluster_rank();
  const int obj = blockIdx.x / cl;
  const int warp = threadIdx.x >> 5;
  // The step descriptors move from the kernel-parameter constant bank to shared memory once: they are read with a
  // runtime index by every warp in every phase.
  __shared__ KStep s_steps[USVM_CHAIN_MAX_STEPS];
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(p.steps);
    uint32_t* dst = reinterpret_cast<uint32_t*>(s_steps);
    const int words = p.n_steps * (int)(sizeof(KStep) / 4);
    for (int i = threadIdx.x; i < words; i += TC_BLOCK) dst[i] = src[i];
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < TC_RING; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], TC_WARPS);
    }
    mbar_fence_init();
  }
  __syncthreads();
  float* scratch = p.scratch + (size_t)obj * cl * (TC_ROWS * (TC_T2I_C + TC_HEADS * 2));

  if (warp == TC_WARPS) {
    // ---- producer warp: streams every step's weight rows through the ring.  Weights never depend on earlier steps
    // (or on earlier kernels), so it runs ahead: before the grid dependency resolves, and across cluster barriers by
    // up to the ring depth.
    const int plane = threadIdx.x & 31;
    const bool lead = plane == 0;
    uint32_t it = 0;
    // whole warp: chunk (row block rbk, k-chunk kc) of a step; a full-K chunk is one contiguous copy, a k-cut chunk is
    // one copy per weight row (issued by the lanes in parallel, all counted on the slot's barrier)
    auto issue = [&](const KStep& st, int r_lo, int r_hi, int rbk, int kc) {
      const uint32_t slot = it & (TC_RING - 1), ph = (it / TC_RING) & 1u;
      const int row = r_lo + (rbk << st.rb_shift);
      const int nrows = min(1 << st.rb_shift, r_hi - row);
      const int kbase = kc * st.kw, kwid = min(st.kw, st.K - kbase);
      float* dst = ring + slot * TC_CHUNK_FLOATS;
      if (lead) {
        mbar_wait(&empty_bar[slot], ph ^ 1u);
        mbar_arrive_expect_tx(&full_bar[slot], (uint32_t)(nrows * kwid) * 4u);
      }
      __syncwarp();
      if (st.nkc == 1) {
        if (lead) bulk_load(dst, st.w + (size_t)row * st.K, (uint32_t)(nrows * kwid) * 4u, &full_bar[slot]);
      } else {
        for (int r = plane; r < nrows; r += 32)
          bulk_load(dst + r * kwid, st.w + (size_t)(row + r) * st.K + kbase, (uint32_t)kwid * 4u, &full_bar[slot]);
      }
      ++it;
    };
    // (rbk, kc) cursor of the step being issued; chunks are issued strictly in consumption order
    int cur = 0, rbk = 0, kc = 0, r_lo = 0, r_hi = 0;
    bool open = false;  // cursor initialised for step `cur`
    auto advance = [&](int limit_step, int budget) {  // issue up to `budget` chunks of steps <= limit_step
      while (cur <= limit_step && cur < p.n_steps && budget > 0) {
        const KStep& st = s_steps[cur];
        if (st.kind != USVM_CHAIN_LINEAR) {
          ++cur;
          open = false;
          continue;
        }
        if (!open) {
          r_lo = min(st.total_rows, rank * st.per);
          r_hi = min(st.total_rows, r_lo + st.per);
          rbk = 0;
          kc = 0;
          open = true;
        }
        if (r_lo + (rbk << st.rb_shift) >= r_hi) {
          ++cur;
          open = false;
          continue;
        }
        issue(st, r_lo, r_hi, rbk, kc);
        --budget;
        if (++kc == st.nkc) {
          kc = 0;
          ++rbk;
        }
      }
    };
    for (int s = 0; s < p.n_steps; ++s) {
      advance(s, 1 << 30);        // everything step s still needs (blocks on ring slots as the consumers free them)
      if (s == 0) pdl_wait();
      if (s + 1 < p.n_steps) {
        advance(s + 1, TC_RING);  // run ahead into the next step by up to the ring depth, then meet the barrier
        __syncwarp();
        cluster_barrier();
      }
    }
    return;
  }

  pdl_wait();
  pdl_trigger();
  unsigned long long* stamp = (p.timing && blockIdx.x == 0 && threadIdx.x == 0) ? p.timing : nullptr;
  uint32_t it = 0;
  for (int s = 0; s < p.n_steps; ++s) {
    const KStep& st = s_steps[s];
    if (stamp) stamp[s * 8 + 0] = clock64();
    if (st.kind == USVM_CHAIN_T2I_PARTIAL) {
      t2i_partial_step(st, obj, rank, cl, scratch, xs, sq);
      if (stamp) stamp[s * 8 + 1] = stamp[s * 8 + 2] = clock64();
    } else {
      if (st.in_kind == USVM_CHAIN_IN_SELF_ATTN) load_self_attention(st, obj, xs, xs2, xs + TC_ROWS * 256);
      else if (st.in_kind == USVM_CHAIN_IN_T2I_MERGE) load_t2i_merge(st, scratch, cl, xs);
      else load_rows(st, obj, rank, xs, xs2);
      if (stamp) stamp[s * 8 + 1] = clock64();
      if (p.precise) linear_step<true>(st, obj, rank, xs, xs2, ring, full_bar, empty_bar, red, it);
      else linear_step<false>(st, obj, rank, xs, xs2, ring, full_bar, empty_bar, red, it);
      if (stamp) stamp[s * 8 + 2] = clock64();
    }
    if (s + 1 < p.n_steps) cluster_barrier();  // also orders this CTA's shared-memory reuse
    if (stamp) stamp[s * 8 + 3] = clock64();
  }
}

bool fits_i32(long long v) { return v >= 0 && v < (1LL << 30); }
int ilog2(int v) {
  int s = 0;
  while ((1 << s) < v) ++s;
  return s;
}

}  // namespace

extern "C" int usvm_token_chain(const usvm_chain_params* p, void* stream) {
  if (!p || p->n_steps <= 0 || p->n_steps > USVM_CHAIN_MAX_STEPS || p->n_obj <= 0) return USVM_ERR_ARG;
  if (p->cluster != 8 && p->cluster != 16) return USVM_ERR_ARG;
  KParams kp;
  memset(&kp, 0, sizeof(kp));
  kp.scratch = p->scratch;
  kp.timing = p->timing;
  kp.n_steps = p->n_steps;
  kp.n_obj = p->n_obj;
  kp.cluster = p->cluster;
  kp.precise = p->precise;
  for (int s = 0; s < p->n_steps; ++s) {
    const usvm_chain_step& st = p->steps[s];
    KStep& k = kp.steps[s];
    if (st.rows <= 0 || st.rows > USVM_CHAIN_ROWS || !st.x) return USVM_ERR_ARG;
    const long long strides[] = {st.x_os, st.x_rs, st.sel_stride, st.ln_os, st.ln_rs, st.x2_os, st.x2_rs, st.b_is,
                                 st.r_os,  st.r_rs, st.o_os,      st.o_rs,  st.kv_os, st.kv_rs, st.w_is};
    for (long long v : strides)
      if (!fits_i32(v * p->n_obj)) return USVM_ERR_ARG;
    k.x = st.x; k.ln_w = st.ln_w; k.ln_b = st.ln_b; k.x2 = st.x2; k.w = st.w; k.bias = st.bias; k.residual = st.residual;
    k.k = st.k; k.v = st.v; k.ln_out = st.ln_out; k.out = st.out; k.row_select = st.row_select;
    k.x_os = (int)st.x_os; k.x_rs = (int)st.x_rs; k.sel_stride = (int)st.sel_stride; k.ln_os = (int)st.ln_os;
    k.ln_rs = (int)st.ln_rs; k.x2_os = (int)st.x2_os; k.x2_rs = (int)st.x2_rs; k.b_is = (int)st.b_is;
    k.r_os = (int)st.r_os; k.r_rs = (int)st.r_rs; k.o_os = (int)st.o_os; k.o_rs = (int)st.o_rs;
    k.kv_os = (int)st.kv_os; k.kv_rs = (int)st.kv_rs;
    k.kind = st.kind; k.in_kind = st.in_kind; k.rows = st.rows; k.N = st.N; k.K = st.K; k.act = st.act;
    k.x2_cols = st.x2_cols; k.Nk = st.Nk; k.attn_q = st.attn_q; k.attn_k = st.attn_k; k.attn_v = st.attn_v;
    k.stacked = st.w_is != 0; k.ln_eps = st.ln_eps;
    if (st.kind == USVM_CHAIN_T2I_PARTIAL) {
      if (!st.k || !st.v || !p->scratch || st.Nk <= 0 || st.Nk > TC_KEYS * p->cluster || (st.kv_rs % 4) ||
          (st.x_rs % 4) || (reinterpret_cast<uintptr_t>(st.k) & 15) || (reinterpret_cast<uintptr_t>(st.v) & 15))
        return USVM_ERR_ARG;
      k.keys_per = (st.Nk + p->cluster - 1) / p->cluster;
    } else if (st.kind == USVM_CHAIN_LINEAR) {
      if (!st.w || !st.out || st.N <= 0 || st.K <= 0 || (st.K % 32) || st.K > TC_KMAX) return USVM_ERR_ARG;
      if ((reinterpret_cast<uintptr_t>(st.w) & 15)) return USVM_ERR_ARG;
      if (st.w_is) {  // stacked matrices must be dense ([rows][N][K]) and a 16-row MMA tile must stay inside one of them
        if (st.w_is != (long long)st.N * st.K || (st.N % 16) || (st.N & (st.N - 1))) return USVM_ERR_ARG;
        k.n_shift = ilog2(st.N);
      }
      if (st.in_kind == USVM_CHAIN_IN_ROWS) {
        if ((st.x_rs % 4) || (st.x_os % 4) || (st.sel_stride % 4) || (reinterpret_cast<uintptr_t>(st.x) & 15)) return USVM_ERR_ARG;
        if (st.x2 && (st.K > 768 || (st.x2_rs % 4) || (st.x2_os % 4) || (st.x2_cols % 16) ||
                      (reinterpret_cast<uintptr_t>(st.x2) & 15)))
          return USVM_ERR_ARG;
        if (st.ln_w && (st.K > 512 || !st.ln_b || (reinterpret_cast<uintptr_t>(st.ln_w) & 15) ||
                        (reinterpret_cast<uintptr_t>(st.ln_b) & 15)))
          return USVM_ERR_ARG;
        if (st.ln_out && ((st.ln_rs % 4) || (st.ln_os % 4) || (reinterpret_cast<uintptr_t>(st.ln_out) & 15))) return USVM_ERR_ARG;
      } else if (st.in_kind == USVM_CHAIN_IN_SELF_ATTN) {
        if (st.K != 256 || (st.x_rs % 4) || (st.x_os % 4) || ((st.attn_q | st.attn_k | st.attn_v) % 4) || st.x2 ||
            (reinterpret_cast<uintptr_t>(st.x) & 15))
          return USVM_ERR_ARG;
      } else if (st.in_kind == USVM_CHAIN_IN_T2I_MERGE) {
        if (st.K != 128 || !p->scratch || st.x2) return USVM_ERR_ARG;
      } else {
        return USVM_ERR_ARG;
      }
      // schedule: rows per CTA (multiple of the 16-row MMA tile), chunk = rb rows x kw columns <= 32 KB
      k.total_rows = st.w_is ? st.rows * st.N : st.N;
      k.per = (((k.total_rows + p->cluster - 1) / p->cluster) + 15) & ~15;
      k.kw = st.K < 512 ? st.K : 512;
      int rb = 64;
      while (rb > 16 && rb * k.kw * 4 > TC_CHUNK_BYTES) rb >>= 1;
      k.rb_shift = ilog2(rb);
      k.ng_shift = k.rb_shift - 4;
      k.nkc = (st.K + k.kw - 1) / k.kw;
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
  cfg.blockDim = dim3(TC_BLOCK);
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
  if (cudaLaunchKernelEx(&cfg, token_chain_kernel, kp) != cudaSuccess) return USVM_ERR_CUDA;
  return usvm_check_launch();
}
