// Image-encoder attention on the 5th-gen tensor cores: the global blocks and the 14 x 14 windowed blocks of the Hiera
// trunk (MultiScaleAttention.forward, hieradet.py:56-81, heads of 96) and of the EfficientTAM ViT trunk (heads of 64),
// straight from the raster-order qkv tensor the projection GEMM wrote -- window_partition / window_unpartition
// (backbones/utils.py:17-61) are TMA box coordinates and an output row index here, not passes over memory.
//
// One CTA owns up to 128 queries of one (frame, head) and walks that head's key tiles:
//     global   128 consecutive tokens x (T / 128) tiles of 128 keys             2-D tensor map [F*T, 3C], box {64, 128}
//     window   7 rows of a 14 x 14 window (98 queries) x 1-2 tiles of 7 window   4-D tensor map [F, H, W, 3C],
//              rows (98 keys, MMA N / K = 112); a whole 7 x 7 window (49 / 64);  box {64, ws, 7, 1}: OOB -> zero fill
//              or the 49 max-pooled queries of a 14 x 14 window (q-pool blocks)
// TMEM (512 columns): Q bf16 [0,64) -- S fp32 [64,192) / [192,320), double buffered; P (bf16 pairs) overwrites the first
// 64 columns of its S buffer -- O fp32 [320, 320 + HD).  Both MMAs take their A operand (Q, P) from TMEM; K / V tiles
// arrive by TMA (128-byte swizzle) through a 3-stage ring, V is consumed in place as an MN-major operand.  A head of 96 is
// loaded as two 64-column boxes; the second one's upper half belongs to the neighbouring head and is never multiplied.
//
// Tokens in the zero padding of a partial window (the reference pads AFTER norm1, so their k / v are the projection
// bias, hieradet.py:150-158) never reach shared memory: all npad of them share one score q . b_k, so their softmax mass
// npad * exp(q . b_k - m) and their output contribution (that mass times b_v) are added in closed form by the softmax
// threads -- the zero-filled OOB rows of the TMA box are simply masked.
//
// Warp roles (320 threads): warp 0 TMA producer, warp 1 tcgen05.mma issuer (S_{j+1} is issued before the softmax of
// tile j is awaited), warps 2-9 softmax: thread <-> (query row, 64 of the tile's 128 key columns), row max exchanged
// through shared memory + a 64-thread named barrier, base-2 online softmax with lazy rescaling of O.
#include "common.cuh"
#include "usvm2_b200.h"

namespace {

constexpr int QM = 128;
constexpr int KN = 128;          // key rows per smem tile (window mode fills 98 of them, the MMAs read 112)
constexpr int STAGES = 3;
constexpr int THREADS = 320;
constexpr int CHUNK_BYTES = KN * 128;  // one 64-column (128-byte) chunk of a K or V tile
constexpr int XCH_BYTES = 4 * QM * 4;
constexpr int ALIGN_SLACK = 1024;
constexpr int WIN_ROWS = 7;  // window rows per key tile / query tile: 14 x 14 windows are two tiles of 98, 7 x 7 windows one of 49
__host__ __device__ constexpr int win_n(int ws) { return ws == 14 ? 112 : 64; }  // MMA N / K covering 7 * ws keys

__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void st_tmem_x16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                       uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// MN-major operand (rows = K index, 128 bytes = 64 consecutive N elements per row, 128B swizzle):
// 8-row groups 1024 B apart (SBO), 64-element N chunks `lbo_bytes` apart (LBO)
__device__ __forceinline__ uint64_t desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t lo = ((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16);
  uint64_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
  return lo | (hi << 32);
}
// D fp32, A/B bf16, A K-major, B K-major (b_mn = 0) or MN-major (b_mn = 1)
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

template <int HD>
struct Cfg {
  static constexpr int NCH = (HD + 63) / 64;            // 64-column chunks per K / V row
  static constexpr int KV_BYTES = NCH * CHUNK_BYTES;    // one operand, one stage
  static constexpr int QW = HD / 2;                     // 32-bit words of a Q row
  static constexpr int QW0 = HD == 96 ? 32 : 16;        // words written by column half 0 (the rest by half 1)
  static constexpr int SMEM = STAGES * 2 * KV_BYTES + XCH_BYTES + ALIGN_SLACK + 256;
};

template <int HD>
__global__ void __launch_bounds__(THREADS, 1)
hiera_attn_tc5_kernel(const __grid_constant__ CUtensorMap tmKV, const usvm_hiera_attn_params p) {
  using C = Cfg<HD>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem;
  uint8_t* sV = sK + STAGES * C::KV_BYTES;
  float* s_xch = reinterpret_cast<float*>(sV + STAGES * C::KV_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(s_xch) + XCH_BYTES);
  uint64_t* q_ready = bars;       // 256 arrivals
  uint64_t* kv_full = bars + 1;   // STAGES
  uint64_t* kv_empty = bars + 4;  // STAGES
  uint64_t* s_full = bars + 7;    // 2
  uint64_t* p_full = bars + 9;    // 2 (256 arrivals each)
  uint64_t* pv_done = bars + 11;  // 2
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.y, f = blockIdx.z;
  const int dim = p.dim, ld = 3 * dim;
  const bool windowed = p.window > 0;
  // ---- what this CTA owns ----
  int n = 0;                 // key tiles
  int wx0 = 0, wy0 = 0;      // window origin (tokens)
  int qy0 = 0;               // first window row of the query tile
  int npad = 0;
  const int ws = p.window;                                        // 0, 7 or 14
  const bool pooled = p.pool != 0;                                // queries = 2 x 2 max-pool of the window's q (ws == 14)
  const int qtiles = (ws == 14 && !pooled) ? 2 : 1;               // query tiles per window
  const int nq = pooled ? 49 : WIN_ROWS * ws;                     // query rows of a tile
  if (windowed) {
    const int nwx = (p.W + ws - 1) / ws;
    const int win = blockIdx.x / qtiles, qt = blockIdx.x % qtiles;
    wy0 = (win / nwx) * ws;
    wx0 = (win % nwx) * ws;
    qy0 = qt * WIN_ROWS;
    if (wy0 + qy0 >= p.H) return;  // this half of the window lies entirely in the padding: nothing to write
    n = (ws == 14 && wy0 + WIN_ROWS < p.H) ? 2 : 1;
    npad = ws * ws - min(ws, p.W - wx0) * min(ws, p.H - wy0);
  } else {
    n = (p.H * p.W) / KN;
  }
  const int WIN_KEYS = WIN_ROWS * ws, WIN_N = win_n(ws);          // (window mode) keys per tile, MMA N / K
  const int rows_per_tile = windowed ? WIN_KEYS : KN;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmKV);
    mbar_init(q_ready, 256);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&p_full[s], 256);
      mbar_init(&pv_done[s], 1);
    }
    mbar_fence_init();
  }
  if (windowed) {
    // rows [98, 112) / [49, 64) of every V chunk are read by the second MMA with P = 0: they must hold finite values.
    // TMA never writes them (the box has 98 / 49 rows), so clearing them once is enough.
    for (int i = threadIdx.x; i < STAGES * C::NCH * (WIN_N - WIN_KEYS) * 8; i += THREADS) {
      const int chunk = i / ((WIN_N - WIN_KEYS) * 8), rem = i % ((WIN_N - WIN_KEYS) * 8);
      *reinterpret_cast<uint4*>(sV + chunk * CHUNK_BYTES + (WIN_KEYS + rem / 8) * 128 + (rem % 8) * 16) =
          make_uint4(0u, 0u, 0u, 0u);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) tc5_alloc(tmem_slot, 512);
  tc5_fence_before();
  __syncthreads();
  tc5_fence_after();
  pdl_wait();
  pdl_trigger();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tmem_Q = tmem, tmem_S = tmem + 64, tmem_O = tmem + 320;

  if (warp == 0) {
    // (whole warp walks the loop, one elected lane issues: see elect_one() in common.cuh)
    const uint32_t tx_bytes = 2u * C::NCH * (uint32_t)rows_per_tile * 128u;
    for (int j = 0; j < n; ++j) {
      const int st = j % STAGES;
      mbar_wait(&kv_empty[st], ((j / STAGES) & 1) ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&kv_full[st], tx_bytes);
#pragma unroll
        for (int c = 0; c < C::NCH; ++c) {
          uint8_t* kd = sK + st * C::KV_BYTES + c * CHUNK_BYTES;
          uint8_t* vd = sV + st * C::KV_BYTES + c * CHUNK_BYTES;
          const int kc = dim + h * HD + c * 64, vc = 2 * dim + h * HD + c * 64;
          if (windowed) {
            tma_load_4d(kd, &tmKV, &kv_full[st], kc, wx0, wy0 + j * WIN_ROWS, f);
            tma_load_4d(vd, &tmKV, &kv_full[st], vc, wx0, wy0 + j * WIN_ROWS, f);
          } else {
            const int row = f * p.H * p.W + j * KN;
            tma_load_2d(kd, &tmKV, &kv_full[st], kc, row);
            tma_load_2d(vd, &tmKV, &kv_full[st], vc, row);
          }
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    const uint32_t idesc_s = idesc_bf16(QM, windowed ? WIN_N : KN, 0);
    constexpr uint32_t idesc_o = idesc_bf16(QM, HD, 1);
    const int ksteps_pv = (windowed ? WIN_N : KN) / 16;
    auto issue_s = [&](int j) {
      const int st = j % STAGES;
      mbar_wait(&kv_full[st], (j / STAGES) & 1);
      tc5_fence_after();
      const uint32_t k_addr = smem_u32(sK + st * C::KV_BYTES);
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < HD / 16; ++kk) {
          const uint32_t koff = (kk >> 2) * CHUNK_BYTES + (kk & 3) * 32;
          mma_ts(tmem_S + (j & 1) * KN, tmem_Q + kk * 8, umma_desc_k_sw128(k_addr + koff), idesc_s, kk > 0 ? 1u : 0u);
        }
        tc5_commit(&s_full[j & 1]);
      }
      __syncwarp();
    };
    mbar_wait(q_ready, 0);
    tc5_fence_after();
    issue_s(0);
    for (int j = 0; j < n; ++j) {
      if (j + 1 < n) issue_s(j + 1);
      mbar_wait(&p_full[j & 1], (j >> 1) & 1);
      tc5_fence_after();
      const uint32_t v_addr = smem_u32(sV + (j % STAGES) * C::KV_BYTES);
      if (elect_one()) {
        if (windowed) {
#pragma unroll
          for (int kk = 0; kk < 112 / 16; ++kk)
            if (kk < WIN_N / 16)
              mma_ts(tmem_O, tmem_S + (j & 1) * KN + kk * 8, desc_mn_sw128(v_addr + kk * 2048, CHUNK_BYTES), idesc_o,
                     (j > 0 || kk > 0) ? 1u : 0u);
        } else {
#pragma unroll
          for (int kk = 0; kk < KN / 16; ++kk)
            mma_ts(tmem_O, tmem_S + (j & 1) * KN + kk * 8, desc_mn_sw128(v_addr + kk * 2048, CHUNK_BYTES), idesc_o,
                   (j > 0 || kk > 0) ? 1u : 0u);
        }
        tc5_commit(&kv_empty[j % STAGES]);
        tc5_commit(&pv_done[j & 1]);
      }
      __syncwarp();
    }
    (void)ksteps_pv;
  } else {
    const int lane_grp = warp & 3;
    const int half = (warp - 2) >> 2;
    const int r = lane_grp * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(lane_grp * 32) << 16;
    const float sl2 = p.scale * 1.4426950408889634f;
    const int bar_id = 1 + lane_grp;
    // ---- which token this query row is ----
    bool row_ok;
    long long tok;  // row of qkv / out
    long long tok_out;   // row of out (the pooled grid for q-pool blocks)
    int pty = 0, ptx = 0;  // (pooled) top-left token of this query's 2 x 2 pooling cell
    if (windowed && pooled) {
      const int i = r / 7, j2 = r % 7;
      pty = wy0 + 2 * i;
      ptx = wx0 + 2 * j2;
      const int Hp = p.H >> 1, Wp = p.W >> 1;
      row_ok = r < nq && (pty >> 1) < Hp && (ptx >> 1) < Wp;
      tok = ((long long)f * p.H + pty) * p.W + ptx;
      tok_out = ((long long)f * Hp + (pty >> 1)) * Wp + (ptx >> 1);
    } else if (windowed) {
      const int ty = wy0 + qy0 + r / ws, tx = wx0 + r % ws;
      row_ok = r < nq && ty < p.H && tx < p.W;
      tok = ((long long)f * p.H + ty) * p.W + tx;
      tok_out = tok;
    } else {
      row_ok = true;
      tok = (long long)f * p.H * p.W + blockIdx.x * QM + r;
      tok_out = tok;
    }
    // ---- Q row -> TMEM; in window mode also this row's score against a padding token, q . b_k ----
    float spad = 0.f;
    {
      constexpr int W0 = C::QW0, W1 = C::QW - C::QW0;
      constexpr int NW = W0 > W1 ? W0 : W1;
      const int w_lo = half ? W0 : 0, nw = half ? W1 : W0;
      uint32_t w[NW];
      const bf16* qrow = reinterpret_cast<const bf16*>(p.qkv) + tok * ld + h * HD + w_lo * 2;
#pragma unroll
      for (int i = 0; i < NW / 4; ++i) {
        uint4 u = make_uint4(0u, 0u, 0u, 0u);
        if (row_ok && i * 4 < nw) {
          if (!pooled) {
            u = *reinterpret_cast<const uint4*>(qrow + i * 8);
          } else {
            // 2 x 2 max-pool of q over the cell; cell tokens in the zero padding carry q = bias (hieradet.py:60-67)
            __nv_bfloat162 m[4];
            bool first = true;
#pragma unroll
            for (int a = 0; a < 4; ++a) {
              const int yy = pty + (a >> 1), xx = ptx + (a & 1);
              uint4 t;
              if (yy < p.H && xx < p.W) {
                t = *reinterpret_cast<const uint4*>(qrow + ((long long)(a >> 1) * p.W + (a & 1)) * ld + i * 8);
              } else {
                const float* bq = p.qkv_bias + h * HD + w_lo * 2 + i * 8;
                t.x = pack_bf16x2(bq[0], bq[1]); t.y = pack_bf16x2(bq[2], bq[3]);
                t.z = pack_bf16x2(bq[4], bq[5]); t.w = pack_bf16x2(bq[6], bq[7]);
              }
              const __nv_bfloat162* tp = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
              for (int e = 0; e < 4; ++e) m[e] = first ? tp[e] : __hmax2(m[e], tp[e]);
              first = false;
            }
            u = *reinterpret_cast<const uint4*>(m);
          }
        }
        w[4 * i] = u.x; w[4 * i + 1] = u.y; w[4 * i + 2] = u.z; w[4 * i + 3] = u.w;
      }
      if (npad > 0) {
        const float* bk = p.qkv_bias + dim + h * HD + w_lo * 2;
#pragma unroll
        for (int i = 0; i < NW; ++i) {
          if (i < nw) {
            const float2 q2 = unpack_bf16x2(w[i]);
            spad = fmaf(q2.x, __bfloat162float(__float2bfloat16(bk[2 * i])), spad);
            spad = fmaf(q2.y, __bfloat162float(__float2bfloat16(bk[2 * i + 1])), spad);
          }
        }
      }
      if (half == 0) {
        if constexpr (W0 == 32) tc5_st_32x32(tmem_Q + lane_addr, reinterpret_cast<uint32_t(&)[32]>(w));
        else st_tmem_x16(tmem_Q + lane_addr, w);
      } else {
        st_tmem_x16(tmem_Q + lane_addr + W0, w);
      }
      tc5_wait_st_all();
      tc5_fence_before();
      mbar_arrive(q_ready);
    }
    // ---- key masks of this thread's 64 columns, per tile (window mode) ----
    unsigned long long kmask[2] = {~0ull, ~0ull};
    if (windowed) {  // key c of a tile is window row c / ws, column c % ws: valid columns [0, vx) of valid rows [0, vy)
      const int vx = min(ws, p.W - wx0);
      const unsigned long long rowbits = (1ull << vx) - 1ull;
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const int vy = max(0, min(WIN_ROWS, p.H - wy0 - t * WIN_ROWS));
        unsigned long long lo = 0ull, hi = 0ull;
#pragma unroll
        for (int cy = 0; cy < WIN_ROWS; ++cy) {
          if (cy < vy) {
            const int sh = cy * ws;
            if (sh < 64) {
              lo |= rowbits << sh;
              if (sh + ws > 64) hi |= rowbits >> (64 - sh);
            } else {
              hi |= rowbits << (sh - 64);
            }
          }
        }
        kmask[t] = half ? hi : lo;
      }
    }
    float m_ref = -INFINITY, l = 0.f;
    if (npad > 0) {  // the padding tokens are part of the softmax from the start
      s_xch[half * QM + r] = spad;
      asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
      spad = (spad + s_xch[(half ^ 1) * QM + r]) * sl2;
      asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
      m_ref = spad;
      l = half == 0 ? (float)npad : 0.f;
    }
    for (int j = 0; j < n; ++j) {
      const int st = j & 1;
      mbar_wait(&s_full[st], (j >> 1) & 1);
      tc5_fence_after();
      uint32_t sa[64];
      tc5_ld_32x32(tmem_S + lane_addr + st * KN + half * 64, reinterpret_cast<uint32_t(&)[32]>(sa[0]));
      tc5_ld_32x32(tmem_S + lane_addr + st * KN + half * 64 + 32, reinterpret_cast<uint32_t(&)[32]>(sa[32]));
      tc5_wait_ld();
      if (windowed) {
        const unsigned long long km = j == 0 ? kmask[0] : kmask[1];
#pragma unroll
        for (int i = 0; i < 64; ++i)
          if (!((km >> i) & 1ull)) sa[i] = 0xff800000u;
      }
      float mx = -INFINITY;
#pragma unroll
      for (int i = 0; i < 64; ++i) mx = fmaxf(mx, __uint_as_float(sa[i]));
      s_xch[(st * 2 + half) * QM + r] = mx;
      // also orders "both halves have read S_j" before either half overwrites S columns with P_j
      asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
      mx = fmaxf(mx, s_xch[(st * 2 + (half ^ 1)) * QM + r]) * sl2;
      float corr = 1.f;
      bool rescale = false;
      if (mx > m_ref + 8.0f) {  // (first tile without padding tokens: m_ref = -inf, corr = 0, l = 0)
        corr = exp2f(m_ref - mx);
        m_ref = mx;
        l *= corr;
        rescale = true;
      }
      float sum = 0.f;
      uint32_t pk[32];
#pragma unroll
      for (int i = 0; i < 64; i += 2) {
        const float p0 = ex2_approx(fmaf(__uint_as_float(sa[i]), sl2, -m_ref));
        const float p1 = ex2_approx(fmaf(__uint_as_float(sa[i + 1]), sl2, -m_ref));
        sum += p0 + p1;
        pk[i >> 1] = pack_bf16x2(p0, p1);
      }
      l += sum;
      if (j > 0 && __any_sync(0xffffffffu, rescale)) {
        mbar_wait(&pv_done[(j - 1) & 1], ((j - 1) >> 1) & 1);
        tc5_fence_after();
#pragma unroll 1
        for (int c = half * 32; c < HD; c += 64) {
          uint32_t o[32];
          tc5_ld_32x32(tmem_O + lane_addr + c, o);
          tc5_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * corr);
          tc5_st_32x32(tmem_O + lane_addr + c, o);
        }
      }
      // P_j (bf16 pairs) -> first 64 columns of S buffer j&1: this half's 32 words at [half*32, half*32 + 32).
      // S buffer j&1 is rewritten by S_{j+2}, which the MMA warp issues after O += P_j V_j (in-order tensor pipe).
      tc5_st_32x32(tmem_S + lane_addr + st * KN + half * 32, pk);
      tc5_wait_st_all();
      tc5_fence_before();
      mbar_arrive(&p_full[st]);
    }
    asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
    s_xch[half * QM + r] = l;
    asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
    l += s_xch[(half ^ 1) * QM + r];
    mbar_wait(&pv_done[(n - 1) & 1], ((n - 1) >> 1) & 1);
    tc5_fence_after();
    const float inv = 1.f / l;
    const float wpad = npad > 0 ? (float)npad * exp2f(spad - m_ref) : 0.f;  // softmax mass of the padding tokens
#pragma unroll 1
    for (int c = half * 32; c < HD; c += 64) {
      uint32_t o[32];
      tc5_ld_32x32(tmem_O + lane_addr + c, o);
      tc5_wait_ld();
      if (!row_ok) continue;
      float v[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(o[i]);
      if (npad > 0) {
        const float* bv = p.qkv_bias + 2 * dim + h * HD + c;
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = fmaf(wpad, __bfloat162float(__float2bfloat16(bv[i])), v[i]);
      }
      bf16* O = reinterpret_cast<bf16*>(p.out) + tok_out * dim + h * HD + c;
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        uint4 v4;
        v4.x = pack_bf16x2(v[i] * inv, v[i + 1] * inv);
        v4.y = pack_bf16x2(v[i + 2] * inv, v[i + 3] * inv);
        v4.z = pack_bf16x2(v[i + 4] * inv, v[i + 5] * inv);
        v4.w = pack_bf16x2(v[i + 6] * inv, v[i + 7] * inv);
        *reinterpret_cast<uint4*>(O + i) = v4;
      }
    }
  }
  tc5_fence_before();
  __syncthreads();
  if (warp == 1) tc5_dealloc(tmem, 512);
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<PFN_encodeTiled>(ptr);
  }
  return fn;
}

template <int HD>
int launch(const usvm_hiera_attn_params* p, cudaStream_t stream) {
  PFN_encodeTiled enc = encode_fn();
  if (!enc) return USVM_ERR_DRIVER;
  CUtensorMap tm;
  const cuuint64_t ld = 3ull * p->dim;
  CUresult r;
  if (p->window > 0) {
    cuuint64_t gdim[4] = {ld, (cuuint64_t)p->W, (cuuint64_t)p->H, (cuuint64_t)p->F};
    cuuint64_t gstr[3] = {ld * 2, ld * 2 * p->W, ld * 2 * p->W * p->H};
    cuuint32_t box[4] = {64u, (cuuint32_t)p->window, (cuuint32_t)WIN_ROWS, 1u};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(p->qkv), gdim, gstr, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else {
    cuuint64_t gdim[2] = {ld, (cuuint64_t)p->F * p->H * p->W};
    cuuint64_t gstr[1] = {ld * 2};
    cuuint32_t box[2] = {64u, (cuuint32_t)KN};
    cuuint32_t estr[2] = {1, 1};
    r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(p->qkv), gdim, gstr, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  if (r != CUDA_SUCCESS) return USVM_ERR_DRIVER;
  if (cudaFuncSetAttribute(hiera_attn_tc5_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<HD>::SMEM) !=
      cudaSuccess)
    return USVM_ERR_CUDA;
  dim3 grid;
  if (p->window > 0)
    grid = dim3(((p->window == 14 && !p->pool) ? 2 : 1) * cdiv(p->H, p->window) * cdiv(p->W, p->window), p->heads, p->F);
  else
    grid = dim3(p->H * p->W / QM, p->heads, p->F);
  usvm_launch(hiera_attn_tc5_kernel<HD>, grid, dim3(THREADS), Cfg<HD>::SMEM, stream, tm, *p);
  return usvm_check_launch();
}

}  // namespace

extern "C" int usvm_hiera_attn_tc5(const usvm_hiera_attn_params* p, void* stream) {
  if (!p || !p->qkv || !p->out || p->F <= 0 || p->H <= 0 || p->W <= 0 || p->heads <= 0 || p->dim <= 0)
    return USVM_ERR_ARG;
  if (p->dim % p->heads) return USVM_ERR_ARG;
  const int hd = p->dim / p->heads;
  if (hd != 96 && hd != 64) return USVM_ERR_ARG;
  if (p->window != 0 && p->window != 14 && p->window != 7) return USVM_ERR_ARG;
  if (p->pool && (p->window != 14 || (p->H & 1) || (p->W & 1))) return USVM_ERR_ARG;
  if (p->window == 0 && ((p->H * p->W) % KN)) return USVM_ERR_ARG;
  if (p->window > 0 && !p->qkv_bias) return USVM_ERR_ARG;
  if ((reinterpret_cast<uintptr_t>(p->qkv) & 15) || (reinterpret_cast<uintptr_t>(p->out) & 15) || (p->dim % 8))
    return USVM_ERR_ARG;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  return hd == 96 ? launch<96>(p, s) : launch<64>(p, s);
}
