// Flash attention on the 5th-gen tensor cores for the memory-attention path: one head of 256, 1024 queries per
// object, up to 7 x 1024 + 64 keys (RoPEAttention self- and cross-attention, sam/transformer.py:311-360).
//
// One CTA owns 128 queries and a contiguous range of 64-key tiles (split-KV; partials are merged by
// fmha_combine_kernel in attention.cu).  Accumulators live in TMEM:
//     columns [0,64) / [64,128)   S = Q K^T, double buffered (fp32, 128 lanes x 64)
//     columns [128,384)           O           (fp32, 128 lanes x 256)
// Warp roles (320 threads):
//   warp 0   TMA producer: Q once (4 boxes of 128 x 64), then K / V tiles (4 + 4 boxes of 64 x 64) through a
//            2-stage mbarrier ring, 128-byte swizzle
//   warp 1   tcgen05.mma issuer: S_{j+1} = Q K_{j+1}^T is issued before waiting for the softmax of tile j, and P is
//            double buffered, so both MMAs of a tile overlap the softmax of the next one; O += P_j V_j uses V in
//            place as an MN-major operand
//   warps 2-9 softmax: thread <-> (query row, half of the tile's 64 key columns); the row max is exchanged between the
//            two halves through shared memory + a 64-thread named barrier; tcgen05.ld S, online softmax in base 2
//            (ex2.approx) with lazy rescaling (O in TMEM is only rescaled when the running max grows by more than
//            2^8), P written as bf16 into the swizzled K-major smem tile that feeds the second MMA
#include "common.cuh"
#include "usvm2_b200.h"

namespace {

constexpr int QM = 128;   // queries per CTA
constexpr int KN = 64;    // keys per tile
constexpr int HD = 256;   // head dim
constexpr int NCH = HD / 64;  // 64-column (128-byte) chunks per row
constexpr int THREADS = 320;  // TMA warp, MMA warp, 8 softmax warps

constexpr int Q_BYTES = NCH * QM * 128;     // 65536
constexpr int KV_BYTES = NCH * KN * 128;    // 32768 per operand per stage
constexpr int P_BYTES = QM * 128;           // 16384
constexpr int XCH_BYTES = 4 * QM * 4;  // row max exchange between the two column halves, double buffered
constexpr int ALIGN_SLACK = 512;  // the dynamic smem base is 1024-aligned in practice; trap if it is not

__device__ __forceinline__ void tc5_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]),
        "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]),
        "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tc5_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// MN-major operand (rows = K index, 128 bytes = 64 consecutive N elements per row, 128B swizzle):
// 8-row groups 1024 B apart (SBO), 64-element N chunks `lbo_bytes` apart (LBO)
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t lo = ((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16);
  uint64_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
  return lo | (hi << 32);
}
// D fp32, A/B bf16, A K-major, B K-major (b_mn = 0) or MN-major (b_mn = 1)
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

// =================================================================================================================
// Q and P live in TENSOR MEMORY and feed the MMAs as the A operand (tcgen05.mma with A in TMEM), so the only
// shared-memory traffic of a tile is the K / V tile itself (TMA write + one MMA read each): 128 KB per tile.  (A first
// version kept Q and P in shared memory; re-reading the 64 KB Q tile and round-tripping P for every 64 keys made it
// shared-memory bandwidth bound at ~224 KB per tile.)
//   TMEM columns: [0,128) Q (bf16 pairs), [128,192) / [192,256) S_j fp32 -- P_j (bf16 pairs) overwrites the first 32
//   columns of its S buffer once both column halves have read S -- and [256,512) O.
//   Shared memory: 3-stage K / V ring (192 KB).
// =================================================================================================================
constexpr int TS_STAGES = 3;
constexpr int TS_SMEM_BYTES = TS_STAGES * 2 * KV_BYTES + XCH_BYTES + ALIGN_SLACK + 256;

__device__ __forceinline__ void tc5_mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc5_st_32x32_x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

__global__ void __launch_bounds__(THREADS, 1)
fmha_tc5_ts_kernel(const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                   const usvm_fmha_params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  if (smem - smem_raw > ALIGN_SLACK) __trap();
  uint8_t* sK = smem;                              // TS_STAGES stages
  uint8_t* sV = sK + TS_STAGES * KV_BYTES;         // TS_STAGES stages
  float* s_xch = reinterpret_cast<float*>(sV + TS_STAGES * KV_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(s_xch) + XCH_BYTES);
  uint64_t* q_ready = bars;           // 1 (256 arrivals)
  uint64_t* kv_full = bars + 1;       // 3
  uint64_t* kv_empty = bars + 4;      // 3
  uint64_t* s_full = bars + 7;        // 2
  uint64_t* p_full = bars + 9;        // 2 (256 arrivals each)
  uint64_t* pv_done = bars + 11;      // 2
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * QM;
  const int b = blockIdx.y;
  const int split = blockIdx.z;
  const int ntiles = (p.Nk + KN - 1) / KN;
  const int per = (ntiles + p.num_splits - 1) / p.num_splits;
  const int t_begin = split * per;
  const int t_end = min(ntiles, t_begin + per);
  const int n = max(0, t_end - t_begin);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_ready, 256);
    for (int s = 0; s < TS_STAGES; ++s) {
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
  if (warp == 1) tc5_alloc(tmem_slot, 512);
  tc5_fence_before();
  __syncthreads();
  tc5_fence_after();
  pdl_wait();  // barriers, descriptors and TMEM were set up under the previous kernel's tail
  pdl_trigger();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tmem_Q = tmem, tmem_S = tmem + 128, tmem_O = tmem + 256;

  if (warp == 0) {
    // (whole warp walks the loop, one elected lane issues: see elect_one() in common.cuh)
    for (int j = 0; j < n; ++j) {
      const int st = j % TS_STAGES;
      mbar_wait(&kv_empty[st], ((j / TS_STAGES) & 1) ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&kv_full[st], 2 * KV_BYTES);
        const int row = b * p.Nk + (t_begin + j) * KN;
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          tma_load_2d(sK + st * KV_BYTES + c * (KN * 128), &tmK, &kv_full[st], c * 64, row);
          tma_load_2d(sV + st * KV_BYTES + c * (KN * 128), &tmV, &kv_full[st], c * 64, row);
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    if (n > 0) {
      constexpr uint32_t idesc_s = idesc_bf16(QM, KN, 0);
      constexpr uint32_t idesc_o = idesc_bf16(QM, HD, 1);
      auto issue_s = [&](int j) {
        const int st = j % TS_STAGES;
        mbar_wait(&kv_full[st], (j / TS_STAGES) & 1);
        tc5_fence_after();
        const uint32_t k_addr = smem_u32(sK + st * KV_BYTES);
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < HD / 16; ++kk) {
            const uint32_t koff = (kk >> 2) * (KN * 128) + (kk & 3) * 32;
            tc5_mma_f16_ts(tmem_S + (j & 1) * KN, tmem_Q + kk * 8, umma_desc_k_sw128(k_addr + koff), idesc_s,
                           kk > 0 ? 1u : 0u);
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
        const uint32_t v_addr = smem_u32(sV + (j % TS_STAGES) * KV_BYTES);
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < KN / 16; ++kk) {
            tc5_mma_f16_ts(tmem_O, tmem_S + (j & 1) * KN + kk * 8, umma_desc_mn_sw128(v_addr + kk * 2048, KN * 128),
                           idesc_o, (j > 0 || kk > 0) ? 1u : 0u);
          }
          tc5_commit(&kv_empty[j % TS_STAGES]);
          tc5_commit(&pv_done[j & 1]);
        }
        __syncwarp();
      }
    }
  } else {
    const int lane_grp = warp & 3;
    const int half = (warp - 2) >> 2;
    const int r = lane_grp * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(lane_grp * 32) << 16;
    const float sl2 = p.scale * 1.4426950408889634f;
    // ---- Q row r, columns [half*128, half*128 + 128) -> TMEM (bf16 pairs: 64 words per half row) ----
    {
      const bf16* qrow = reinterpret_cast<const bf16*>(p.q) + (long long)b * p.q_bs + (long long)(q0 + r) * p.q_rs +
                         half * 128;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t w[32];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint4 u = *reinterpret_cast<const uint4*>(qrow + c * 64 + i * 8);
          w[4 * i] = u.x; w[4 * i + 1] = u.y; w[4 * i + 2] = u.z; w[4 * i + 3] = u.w;
        }
        tc5_st_32x32(tmem_Q + lane_addr + half * 64 + c * 32, w);
      }
      tc5_wait_st();
      tc5_fence_before();
      mbar_arrive(q_ready);
    }
    float m_ref = -INFINITY, l = 0.f;
    for (int j = 0; j < n; ++j) {
      const int st = j & 1;
      mbar_wait(&s_full[st], (j >> 1) & 1);
      tc5_fence_after();
      uint32_t sa[32];
      tc5_ld_32x32(tmem_S + lane_addr + st * KN + half * 32, sa);
      tc5_wait_ld();
      const int key0 = (t_begin + j) * KN + half * 32;
      if (key0 + 32 > p.Nk) {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (key0 + i >= p.Nk) sa[i] = 0xff800000u;
      }
      float mx = -INFINITY;
#pragma unroll
      for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(sa[i]));
      s_xch[(st * 2 + half) * QM + r] = mx;
      // also orders "both halves have read S_j" before either half overwrites S columns with P_j
      asm volatile("bar.sync %0, 64;" ::"r"(1 + lane_grp) : "memory");
      mx = fmaxf(mx, s_xch[(st * 2 + (half ^ 1)) * QM + r]) * sl2;
      float corr = 1.f;
      bool rescale = false;
      if (j == 0) {
        m_ref = mx;
      } else if (mx > m_ref + 8.0f) {
        corr = exp2f(m_ref - mx);
        m_ref = mx;
        l *= corr;
        rescale = true;
      }
      float sum = 0.f;
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
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
        for (int c = half * 128; c < half * 128 + 128; c += 32) {
          uint32_t o[32];
          tc5_ld_32x32(tmem_O + lane_addr + c, o);
          tc5_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * corr);
          tc5_st_32x32(tmem_O + lane_addr + c, o);
        }
      }
      // P_j (bf16 pairs) -> first 32 columns of S buffer j&1: this half's 16 words at [half*16, half*16 + 16).
      // S buffer j&1 is rewritten by S_{j+2}, which the MMA warp issues after O += P_j V_j (in-order tensor pipe).
      tc5_st_32x32_x16(tmem_S + lane_addr + st * KN + half * 16, pk);
      tc5_wait_st();
      tc5_fence_before();
      mbar_arrive(&p_full[st]);
    }
    asm volatile("bar.sync %0, 64;" ::"r"(1 + lane_grp) : "memory");
    s_xch[half * QM + r] = l;
    asm volatile("bar.sync %0, 64;" ::"r"(1 + lane_grp) : "memory");
    l += s_xch[(half ^ 1) * QM + r];
    const int row = q0 + r;
    if (n > 0) {
      mbar_wait(&pv_done[(n - 1) & 1], ((n - 1) >> 1) & 1);
      tc5_fence_after();
    }
    const float inv = (p.num_splits == 1 && l > 0.f) ? 1.f / l : 1.f;
#pragma unroll 1
    for (int c = half * 128; c < half * 128 + 128; c += 32) {
      uint32_t o[32];
      if (n > 0) {
        tc5_ld_32x32(tmem_O + lane_addr + c, o);
        tc5_wait_ld();
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] = 0u;
      }
      if (p.num_splits == 1) {
        bf16* O = reinterpret_cast<bf16*>(p.o) + (long long)b * p.o_bs + (long long)row * p.o_rs + c;
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          uint4 v4;
          v4.x = pack_bf16x2(__uint_as_float(o[i]) * inv, __uint_as_float(o[i + 1]) * inv);
          v4.y = pack_bf16x2(__uint_as_float(o[i + 2]) * inv, __uint_as_float(o[i + 3]) * inv);
          v4.z = pack_bf16x2(__uint_as_float(o[i + 4]) * inv, __uint_as_float(o[i + 5]) * inv);
          v4.w = pack_bf16x2(__uint_as_float(o[i + 6]) * inv, __uint_as_float(o[i + 7]) * inv);
          *reinterpret_cast<uint4*>(O + i) = v4;
        }
      } else if (p.part_bf16) {
        bf16* OP = reinterpret_cast<bf16*>(p.o_part) + (((long long)split * gridDim.y + b) * p.Nq + row) * HD + c;
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          uint4 v4;
          v4.x = pack_bf16x2(__uint_as_float(o[i]), __uint_as_float(o[i + 1]));
          v4.y = pack_bf16x2(__uint_as_float(o[i + 2]), __uint_as_float(o[i + 3]));
          v4.z = pack_bf16x2(__uint_as_float(o[i + 4]), __uint_as_float(o[i + 5]));
          v4.w = pack_bf16x2(__uint_as_float(o[i + 6]), __uint_as_float(o[i + 7]));
          *reinterpret_cast<uint4*>(OP + i) = v4;
        }
      } else {
        float* OP = p.o_part + (((long long)split * gridDim.y + b) * p.Nq + row) * HD + c;
#pragma unroll
        for (int i = 0; i < 32; i += 4)
          *reinterpret_cast<float4*>(OP + i) = make_float4(__uint_as_float(o[i]), __uint_as_float(o[i + 1]),
                                                           __uint_as_float(o[i + 2]), __uint_as_float(o[i + 3]));
      }
    }
    if (p.num_splits > 1 && half == 0) {
      float* ML = p.ml_part + (((long long)split * gridDim.y + b) * p.Nq + row) * 2;
      ML[0] = n > 0 ? m_ref / sl2 : -INFINITY;
      ML[1] = l;
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

int make_map(CUtensorMap* map, const void* base, long long rows, int cols, long long pitch_elems, int box_rows) {
  PFN_encodeTiled enc = encode_fn();
  if (!enc) return USVM_ERR_DRIVER;
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)pitch_elems * 2};
  cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? USVM_OK : USVM_ERR_DRIVER;
}

}  // namespace

// Same parameter block as usvm_fmha_bf16.  Requirements: head_dim 256, H == 1, Nq % 128 == 0, contiguous batches
// (q_bs == Nq * q_rs, k_bs == Nk * k_rs, v_bs == Nk * v_rs), 16-byte aligned bases, row strides % 8 == 0.
// Partials (num_splits > 1) use the o_part / ml_part layout of usvm_fmha_bf16; call usvm_fmha_combine afterwards.
extern "C" int usvm_fmha_tc5(const usvm_fmha_params* p, void* stream) {
  if (!p || !p->q || !p->k || !p->v || !p->o || p->B <= 0 || p->Nq <= 0 || p->Nk <= 0) return USVM_ERR_ARG;
  if (p->head_dim != HD || p->H != 1 || (p->Nq % QM) || p->num_splits < 1 || p->num_splits > 32) return USVM_ERR_ARG;
  if (p->q_bs != (long long)p->Nq * p->q_rs || p->k_bs != (long long)p->Nk * p->k_rs ||
      p->v_bs != (long long)p->Nk * p->v_rs)
    return USVM_ERR_ARG;
  if ((p->q_rs % 8) || (p->k_rs % 8) || (p->v_rs % 8) || (p->o_rs % 8)) return USVM_ERR_ARG;
  if ((reinterpret_cast<uintptr_t>(p->q) & 15) || (reinterpret_cast<uintptr_t>(p->k) & 15) ||
      (reinterpret_cast<uintptr_t>(p->v) & 15) || (reinterpret_cast<uintptr_t>(p->o) & 15))
    return USVM_ERR_ARG;
  if (p->num_splits > 1 && (!p->o_part || !p->ml_part)) return USVM_ERR_ARG;
  if (p->num_splits > (p->Nk + KN - 1) / KN) return USVM_ERR_ARG;
  CUtensorMap tk, tv;
  int rc = make_map(&tk, p->k, (long long)p->B * p->Nk, HD, p->k_rs, KN);
  if (rc) return rc;
  rc = make_map(&tv, p->v, (long long)p->B * p->Nk, HD, p->v_rs, KN);
  if (rc) return rc;
  static UsvmPerDeviceOnce attr = {};
  if (usvm_need_setup(attr)) {
    if (cudaFuncSetAttribute(fmha_tc5_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TS_SMEM_BYTES) !=
        cudaSuccess)
      return USVM_ERR_CUDA;
    usvm_setup_done(attr);
  }
  dim3 grid(p->Nq / QM, p->B, p->num_splits);
  usvm_launch(fmha_tc5_ts_kernel, dim3(grid), dim3(THREADS), TS_SMEM_BYTES, reinterpret_cast<cudaStream_t>(stream), tk, tv, *p);
  return usvm_check_launch();
}
