// Fused Hiera window attention (windows of up to 256 keys): window partition, QKV split, optional 2x2 max-pool
// of the queries, per-(window, head) attention and window un-partition in ONE kernel, reading the raster-order fused qkv
// tensor in place and writing the raster-order output (hieradet.py:46-95 MultiScaleAttention / MultiScaleBlock,
// backbones/utils.py:17-58 window_partition / window_unpartition).
//
// The generic path (usvm_window_gather -> usvm_fmha_bf16 -> usvm_window_scatter) moves q, k and v through HBM twice
// more than necessary and runs a 64 x 64-tile flash kernel on 16 x 16 problems; for the four stage-1/2 blocks that is
// 487 us per 8-frame pass against ~70 us of unavoidable traffic.  Here one CTA owns one window:
//   * rows of q / k / v are fetched with cp.async straight from qkv[f, y, x, which*C + h*96 ...]; tokens that fall into
//     the zero padding of a partial window carry qkv = bias (the reference pads after norm1, before the qkv Linear);
//     pooled queries are the element-wise max of their 2 x 2 source tokens;
//   * a work unit is (head, 16-query slab): S = Q K^T with mma.sync m16n8k16 bf16 over <= 64 keys in registers, row
//     softmax in base 2, O = P V with P re-used as the A fragment (FlashAttention-2 register pipeline, one key tile);
//   * the slab's rows go back to out[f, oy, ox, h*96 ...] -- raster order, padding rows dropped.
// Windows of <= 64 keys: one CTA per window, 4 heads per shared-memory pass (3 x 64 x (4*96 + 8) bf16 = 147 KB).
// Larger windows (14 x 14 = 196 keys): one CTA per (window, head); the keys are walked in tiles of 64 with the online
// softmax of the flash kernel (720 x 104 bf16 = 146 KB).
#include "common.cuh"
#include "usvm2_b200.h"

namespace {

constexpr int WA_THREADS = 128;
constexpr int WA_ROWS = 64;  // max keys (and max queries) per window
// (head dim D is a template parameter: 96 = Hiera-tiny / small, 64 = Hiera-B+ heads of 56 zero-padded by the packer)
constexpr int WA_HC = 4;     // heads per pass

__device__ __forceinline__ void wa_cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void wa_ldsm_x4(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(p)));
}
__device__ __forceinline__ void wa_ldsm_x4_t(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(p)));
}
__device__ __forceinline__ void wa_mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint4 bias_chunk(const float* b) {  // 8 fp32 -> 8 bf16
  uint4 u;
  u.x = pack_bf16x2(b[0], b[1]); u.y = pack_bf16x2(b[2], b[3]);
  u.z = pack_bf16x2(b[4], b[5]); u.w = pack_bf16x2(b[6], b[7]);
  return u;
}
__device__ __forceinline__ uint32_t max_bf16x2(uint32_t a, uint32_t b) {
  const float2 x = unpack_bf16x2(a), y = unpack_bf16x2(b);
  return pack_bf16x2(fmaxf(x.x, y.x), fmaxf(x.y, y.y));
}

// one (head, 16-query slab) unit; NKT = key n-tiles of 8 (2 -> 16 keys, 8 -> up to 64 keys)
template <int NKT, int WA_D>
__device__ __forceinline__ void attend_unit(const bf16* sQ, const bf16* sK, const bf16* sV, int LD, int hoff, int slab,
                                            int nk, float sl2, float (&o)[WA_D / 8][4], float (&inv)[2]) {
  const int lane = threadIdx.x & 31, t = lane & 3;
  float s[NKT][4];
#pragma unroll
  for (int i = 0; i < NKT; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
  for (int ks = 0; ks < WA_D / 16; ++ks) {
    uint32_t a[4];
    wa_ldsm_x4(a, sQ + (slab * 16 + (lane & 15)) * LD + hoff + ks * 16 + (lane >> 4) * 8);
#pragma unroll
    for (int nt = 0; nt < NKT; nt += 2) {
      uint32_t kb[4];
      wa_ldsm_x4(kb, sK + (nt * 8 + (lane & 7) + (lane >> 4) * 8) * LD + hoff + ks * 16 + ((lane >> 3) & 1) * 8);
      wa_mma(s[nt], a, kb[0], kb[1]);
      wa_mma(s[nt + 1], a, kb[2], kb[3]);
    }
  }
  float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
  for (int nt = 0; nt < NKT; ++nt) {
    const int kc = nt * 8 + 2 * t;
    if (kc >= nk) s[nt][0] = s[nt][2] = -INFINITY;
    if (kc + 1 >= nk) s[nt][1] = s[nt][3] = -INFINITY;
    mx[0] = fmaxf(mx[0], fmaxf(s[nt][0], s[nt][1]));
    mx[1] = fmaxf(mx[1], fmaxf(s[nt][2], s[nt][3]));
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
    mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
  }
  uint32_t pf[NKT / 2][4];
  float rs[2] = {0.f, 0.f};
#pragma unroll
  for (int nt = 0; nt < NKT; ++nt) {
    const float p0 = exp2f((s[nt][0] - mx[0]) * sl2);
    const float p1 = exp2f((s[nt][1] - mx[0]) * sl2);
    const float p2 = exp2f((s[nt][2] - mx[1]) * sl2);
    const float p3 = exp2f((s[nt][3] - mx[1]) * sl2);
    rs[0] += p0 + p1;
    rs[1] += p2 + p3;
    pf[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16x2(p0, p1);
    pf[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16x2(p2, p3);
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    rs[r] += __shfl_xor_sync(0xffffffffu, rs[r], 1);
    rs[r] += __shfl_xor_sync(0xffffffffu, rs[r], 2);
    inv[r] = 1.0f / rs[r];  // at least one key is valid
  }
#pragma unroll
  for (int i = 0; i < WA_D / 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
#pragma unroll
  for (int kk = 0; kk < NKT / 2; ++kk) {
#pragma unroll
    for (int dt = 0; dt < WA_D / 8; dt += 2) {
      uint32_t vb[4];
      wa_ldsm_x4_t(vb, sV + (kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LD + hoff + dt * 8 + (lane >> 4) * 8);
      wa_mma(o[dt], pf[kk], vb[0], vb[1]);
      wa_mma(o[dt + 1], pf[kk], vb[2], vb[3]);
    }
  }
}


// the same unit over several 64-key tiles (online softmax, FlashAttention-2 recurrence)
template <int WA_D>
__device__ __forceinline__ void attend_unit_tiles(const bf16* sQ, const bf16* sK, const bf16* sV, int LD, int hoff,
                                                  int slab, int nk, float sl2, float (&o)[WA_D / 8][4], float (&inv)[2]) {
  const int lane = threadIdx.x & 31, t = lane & 3;
#pragma unroll
  for (int i = 0; i < WA_D / 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
  uint32_t aq[WA_D / 16][4];
#pragma unroll
  for (int ks = 0; ks < WA_D / 16; ++ks)
    wa_ldsm_x4(aq[ks], sQ + (slab * 16 + (lane & 15)) * LD + hoff + ks * 16 + (lane >> 4) * 8);
  for (int key0 = 0; key0 < nk; key0 += 64) {
    const bf16* cK = sK + key0 * LD;
    const bf16* cV = sV + key0 * LD;
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < WA_D / 16; ++ks) {
#pragma unroll
      for (int nt = 0; nt < 8; nt += 2) {
        uint32_t kb[4];
        wa_ldsm_x4(kb, cK + (nt * 8 + (lane & 7) + (lane >> 4) * 8) * LD + hoff + ks * 16 + ((lane >> 3) & 1) * 8);
        wa_mma(s[nt], aq[ks], kb[0], kb[1]);
        wa_mma(s[nt + 1], aq[ks], kb[2], kb[3]);
      }
    }
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int kc = key0 + nt * 8 + 2 * t;
      if (kc >= nk) s[nt][0] = s[nt][2] = -INFINITY;
      if (kc + 1 >= nk) s[nt][1] = s[nt][3] = -INFINITY;
      mx[0] = fmaxf(mx[0], fmaxf(s[nt][0], s[nt][1]));
      mx[1] = fmaxf(mx[1], fmaxf(s[nt][2], s[nt][3]));
    }
    float corr[2], mnew[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      mnew[r] = fmaxf(m_run[r], mx[r]);  // finite: every tile holds at least one valid key
      corr[r] = exp2f((m_run[r] - mnew[r]) * sl2);
      m_run[r] = mnew[r];
      l_run[r] *= corr[r];
    }
    uint32_t pf[4][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float p0 = exp2f((s[nt][0] - mnew[0]) * sl2);
      const float p1 = exp2f((s[nt][1] - mnew[0]) * sl2);
      const float p2 = exp2f((s[nt][2] - mnew[1]) * sl2);
      const float p3 = exp2f((s[nt][3] - mnew[1]) * sl2);
      l_run[0] += p0 + p1;
      l_run[1] += p2 + p3;
      pf[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16x2(p0, p1);
      pf[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16x2(p2, p3);
    }
#pragma unroll
    for (int i = 0; i < WA_D / 8; ++i) {
      o[i][0] *= corr[0]; o[i][1] *= corr[0];
      o[i][2] *= corr[1]; o[i][3] *= corr[1];
    }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
      for (int dt = 0; dt < WA_D / 8; dt += 2) {
        uint32_t vb[4];
        wa_ldsm_x4_t(vb, cV + (kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LD + hoff + dt * 8 + (lane >> 4) * 8);
        wa_mma(o[dt], pf[kk], vb[0], vb[1]);
        wa_mma(o[dt + 1], pf[kk], vb[2], vb[3]);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
    inv[r] = 1.0f / l_run[r];
  }
}

// TILED: the variant for windows of more than 64 keys (online softmax over key tiles) needs 168 registers; the <= 64-key
// windows this kernel actually serves in the encoder get their own instantiation at 128 (4 CTAs per SM instead of 3: the
// kernel is bound by load latency at 12 warps per SM, ncu: 25 % of the DRAM peak, 49 % of the cycles without an eligible warp)
template <int WA_D, bool TILED>
__global__ void __launch_bounds__(WA_THREADS, TILED ? 1 : 4)
window_attn_kernel(const bf16* __restrict__ qkv, const float* __restrict__ bias, bf16* __restrict__ out, int F, int Hg,
                   int Wg, int ws, int pool, int C, int H, float scale, int rows_q, int rows_k, int heads_per_cta) {
  PDL_ENTRY();
  extern __shared__ __align__(16) uint8_t wa_smem[];
  const int nwx = (Wg + ws - 1) / ws, nwy = (Hg + ws - 1) / ws;
  const int nk = ws * ws;
  const int wq = pool ? ws / 2 : ws, nq = wq * wq;
  const int Ho = pool ? Hg / 2 : Hg, Wo = pool ? Wg / 2 : Wg;
  const int win = blockIdx.x;
  const int f = win / (nwy * nwx), wi = win - f * (nwy * nwx);
  const int wy = wi / nwx, wx = wi - wy * nwx;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const float sl2 = scale * 1.4426950408889634f;
  const int nk_pad = nk <= 16 ? 16 : (nk + 63) & ~63;
  const int nslab = (nq + 15) >> 4;
  const bf16* fbase = qkv + (long long)f * Hg * Wg * 3 * C;

  const int h_begin = blockIdx.y * heads_per_cta, h_end = min(H, h_begin + heads_per_cta);
  const int pass = nk > WA_ROWS ? 1 : WA_HC;  // heads per shared-memory pass
  for (int h0 = h_begin; h0 < h_end; h0 += pass) {
    const int hc = min(pass, h_end - h0);
    const int Cc = hc * WA_D, LD = Cc + 8, CH = Cc / 8;  // channels / padded row / 16-byte chunks of this pass
    bf16* sQ = reinterpret_cast<bf16*>(wa_smem);
    bf16* sK = sQ + rows_q * LD;  // rows_q / rows_k: padded query / key rows of this window size
    bf16* sV = sK + rows_k * LD;
    // (row, 16-byte chunk) of this thread: one division per pass instead of two per chunk -- the kernel is bound by
    // instruction issue (ncu: IPC 2 with half of the cycles without an eligible warp), and the loops below were 40 % of it
    const int rows_per_iter = WA_THREADS / CH;
    const int r_first = tid / CH, c8 = tid - r_first * CH;
    const bool loader = r_first < rows_per_iter;
    const bool ws_p2 = (ws & (ws - 1)) == 0;
    const int ws_sh = __ffs(ws) - 1;
    const int colh = h0 * WA_D + c8 * 8;
    // ---- keys and values: rows < nk from the window (bias where the window hangs over the image), zeros up to nk_pad
    if (loader) {
      for (int r = r_first; r < nk_pad; r += rows_per_iter) {
        bf16* dk = sK + r * LD + c8 * 8;
        bf16* dv = sV + r * LD + c8 * 8;
        if (r < nk) {
          const int ly = ws_p2 ? r >> ws_sh : r / ws, lx = r - ly * ws;
          const int y = wy * ws + ly, x = wx * ws + lx;
          if (y < Hg && x < Wg) {
            const bf16* src = fbase + ((long long)y * Wg + x) * 3 * C + colh;
            wa_cp_async16(dk, src + C);
            wa_cp_async16(dv, src + 2 * C);
          } else {
            *reinterpret_cast<uint4*>(dk) = bias_chunk(bias + C + colh);
            *reinterpret_cast<uint4*>(dv) = bias_chunk(bias + 2 * C + colh);
          }
        } else {
          *reinterpret_cast<uint4*>(dk) = make_uint4(0u, 0u, 0u, 0u);
          *reinterpret_cast<uint4*>(dv) = make_uint4(0u, 0u, 0u, 0u);
        }
      }
      // ---- queries: rows < nq (max over the 2 x 2 source tokens when pooling), zeros up to the end of the last slab
      const bool wq_p2 = (wq & (wq - 1)) == 0;
      const int wq_sh = __ffs(wq) - 1;
      for (int r = r_first; r < nslab * 16; r += rows_per_iter) {
        bf16* dst = sQ + r * LD + c8 * 8;
        if (r >= nq) {
          *reinterpret_cast<uint4*>(dst) = make_uint4(0u, 0u, 0u, 0u);
        } else if (!pool) {
          const int ly = ws_p2 ? r >> ws_sh : r / ws, lx = r - ly * ws;
          const int y = wy * ws + ly, x = wx * ws + lx;
          if (y < Hg && x < Wg) wa_cp_async16(dst, fbase + ((long long)y * Wg + x) * 3 * C + colh);
          else *reinterpret_cast<uint4*>(dst) = bias_chunk(bias + colh);
        } else {
          const int qy = wq_p2 ? r >> wq_sh : r / wq, qx = r - qy * wq;
          uint4 m = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
          for (int d = 0; d < 4; ++d) {
            const int y = wy * ws + 2 * qy + (d >> 1), x = wx * ws + 2 * qx + (d & 1);
            const uint4 u = (y < Hg && x < Wg) ? __ldg(reinterpret_cast<const uint4*>(fbase + ((long long)y * Wg + x) * 3 * C + colh))
                                               : bias_chunk(bias + colh);
            if (d == 0) m = u;
            else {
              m.x = max_bf16x2(m.x, u.x); m.y = max_bf16x2(m.y, u.y);
              m.z = max_bf16x2(m.z, u.z); m.w = max_bf16x2(m.w, u.w);
            }
          }
          *reinterpret_cast<uint4*>(dst) = m;
        }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    // ---- units: (head of this pass, 16-query slab), dealt to the warps
    for (int u = warp; u < hc * nslab; u += WA_THREADS / 32) {
      const int hh = u / nslab, slab = u - hh * nslab;
      float o[WA_D / 8][4], inv[2];
      if constexpr (TILED) {
        attend_unit_tiles<WA_D>(sQ, sK, sV, LD, hh * WA_D, slab, nk, sl2, o, inv);
      } else {
        if (nk_pad == 16) attend_unit<2, WA_D>(sQ, sK, sV, LD, hh * WA_D, slab, nk, sl2, o, inv);
        else attend_unit<8, WA_D>(sQ, sK, sV, LD, hh * WA_D, slab, nk, sl2, o, inv);
      }
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int r = slab * 16 + g + half * 8;
        if (r < nq) {
          const int oy = wy * wq + r / wq, ox = wx * wq + r % wq;
          if (oy < Ho && ox < Wo) {
            bf16* dst = out + (((long long)f * Ho + oy) * Wo + ox) * C + (h0 + hh) * WA_D + 2 * t;
#pragma unroll
            for (int dt = 0; dt < WA_D / 8; ++dt)
              *reinterpret_cast<uint32_t*>(dst + dt * 8) =
                  pack_bf16x2(o[dt][half * 2] * inv[half], o[dt][half * 2 + 1] * inv[half]);
          }
        }
      }
    }
    __syncthreads();  // the next pass overwrites the tiles
  }
}

}  // namespace

template <int WA_D>
static int launch_window_attn(const void* qkv, const float* qkv_bias, void* out, int F, int Hg, int Wg, int ws, int pool, int C,
                              int heads, void* stream) {
  const int nk = ws * ws, nq = pool ? nk / 4 : nk;
  const bool big = nk > WA_ROWS;                 // several key tiles: one CTA per (window, head)
  const int heads_per_cta = big ? 1 : heads;
  const int hc = big ? 1 : (heads < WA_HC ? heads : WA_HC);
  const int rows_k = nk <= 16 ? 16 : (nk + 63) & ~63, rows_q = (nq + 15) & ~15;
  const size_t smem = (size_t)(rows_q + 2 * rows_k) * (hc * WA_D + 8) * sizeof(bf16);
  if (smem > 200 * 1024) return USVM_ERR_ARG;
  const int nw = cdiv(Hg, ws) * cdiv(Wg, ws);
  const dim3 grid(F * nw, cdiv(heads, heads_per_cta));
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const float scale = 1.0f / sqrtf((float)WA_D);
  if (big) {
    static UsvmPerDeviceOnce configured = {};
    if (usvm_need_setup(configured)) {
      if (cudaFuncSetAttribute(window_attn_kernel<WA_D, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) !=
          cudaSuccess)
        return USVM_ERR_CUDA;
      usvm_setup_done(configured);
    }
    usvm_launch(window_attn_kernel<WA_D, true>, grid, dim3(WA_THREADS), smem, s, reinterpret_cast<const bf16*>(qkv), qkv_bias,
                reinterpret_cast<bf16*>(out), F, Hg, Wg, ws, pool, C, heads, scale, rows_q, rows_k, heads_per_cta);
  } else {
    static UsvmPerDeviceOnce configured = {};
    if (usvm_need_setup(configured)) {
      if (cudaFuncSetAttribute(window_attn_kernel<WA_D, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) !=
              cudaSuccess ||
          cudaFuncSetAttribute(window_attn_kernel<WA_D, false>, cudaFuncAttributePreferredSharedMemoryCarveout,
                               cudaSharedmemCarveoutMaxShared) != cudaSuccess)
        return USVM_ERR_CUDA;
      usvm_setup_done(configured);
    }
    usvm_launch(window_attn_kernel<WA_D, false>, grid, dim3(WA_THREADS), smem, s, reinterpret_cast<const bf16*>(qkv), qkv_bias,
                reinterpret_cast<bf16*>(out), F, Hg, Wg, ws, pool, C, heads, scale, rows_q, rows_k, heads_per_cta);
  }
  return usvm_check_launch();
}

extern "C" int usvm_window_attn_bf16(const void* qkv, const float* qkv_bias, void* out, int F, int Hg, int Wg, int ws,
                                     int pool, int C, int heads, void* stream) {
  if (!qkv || !qkv_bias || !out || F <= 0 || Hg <= 0 || Wg <= 0 || ws <= 0 || heads <= 0) return USVM_ERR_ARG;
  if ((C != heads * 96 && C != heads * 64) || ws * ws > 256 || (pool && ((ws & 1) || (Hg & 1) || (Wg & 1))))
    return USVM_ERR_ARG;
  if ((reinterpret_cast<uintptr_t>(qkv) & 15) || (reinterpret_cast<uintptr_t>(out) & 3) ||
      (reinterpret_cast<uintptr_t>(qkv_bias) & 3))
    return USVM_ERR_ARG;
  if (C == heads * 96) return launch_window_attn<96>(qkv, qkv_bias, out, F, Hg, Wg, ws, pool, C, heads, stream);
  return launch_window_attn<64>(qkv, qkv_bias, out, F, Hg, Wg, ws, pool, C, heads, stream);
}
