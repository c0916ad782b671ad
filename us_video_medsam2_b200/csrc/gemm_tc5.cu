// bf16 GEMM on the 5th-gen tensor cores:  C[M,N] = epilogue(A[M,K] * W[N,K]^T)
//
// Every Linear / 1x1 conv / im2col conv on the propagation path that is run in bf16 goes through
// this kernel (reference call sites: hieradet.py:60,77,164 qkv/proj/mlp; image_encoder.py:114 FPN
// laterals; memory_attention.py + sam/transformer.py:257-286 projections and FFN;
// memory_encoder.py:104-117,170-176 pix_feat_proj / pwconv / out_proj).
//
// Structure (one 128 x BN output tile per CTA, 192 threads):
//   warp 0   : TMA producer  -- cp.async.bulk.tensor 2D boxes {64 x 128} of A and {64 x BN} of W,
//              128-byte swizzle, mbarrier ring of min(4, K/64) stages (zero fill beyond M / N / K edges)
//   warp 1   : TMEM allocator + single-thread tcgen05.mma issuer (M=128, N=BN, K=16, fp32 accum in
//              TMEM), tcgen05.commit releases smem stages and finally signals the epilogue
//   warps 2-5: epilogue -- tcgen05.ld 32x32b.x32 (one accumulator row per thread), fused
//              bias / RoPE / activation / per-column scale / residual, then each warp stages its 32 x 32 block in
//              shared memory and writes it with ONE TMA store (coalesced, edge clipping for free)
// The ring depth follows K, so short-K GEMMs (the Hiera blocks: K = 96 ... 768) use little shared memory and 2-3 CTAs
// share an SM: one CTA's epilogue overlaps another's loads and MMAs without a persistent scheduler.
#include "common.cuh"
#include "usvm2_b200.h"

namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 B = one swizzle row
constexpr int STAGES = 8;  // maximum ring depth (barrier storage); the launch picks the depth per problem
constexpr int GEMM_THREADS = 192;

constexpr int STG_F32 = 32 * 128;  // per-warp staging of a 32 x 32 fp32 block (128-byte rows, 128B swizzle)
constexpr int STG_BF16 = 32 * 64;  // per-warp staging of a 32 x 32 bf16 block (64-byte rows)
constexpr int STG_BYTES = 4 * (STG_F32 + STG_BF16);

template <int BN>
struct SmemLayout {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int total(int stages) {
    return stages * STAGE_BYTES + STG_BYTES + 1024 /* alignment slack */ + 128 /* barriers */;
  }
};

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// at most one bulk group (the previous column block's stores) may still be reading shared memory
__device__ __forceinline__ void tma_store_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_shared(uint32_t local_addr, uint32_t cta_rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(cta_rank));
  return r;
}
__device__ __forceinline__ void st_shared_cluster_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

__host__ __device__ constexpr int tmem_cols(int bn) { return bn <= 32 ? 32 : bn <= 64 ? 64 : bn <= 128 ? 128 : 256; }

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == USVM_ACT_RELU) return fmaxf(v, 0.0f);
  if (act == USVM_ACT_GELU) return gelu_erf(v);
  return v;
}

// Rotary tables are stored TILED, [pos / 32][col / 4][pos % 32][col % 4] (fp32, 128 columns per position): in the epilogue
// a thread owns one position (accumulator row) and needs 16 consecutive table columns, so with a row-major table every
// 16-byte load of a warp touched 32 different cache lines (256 L1 wavefronts per 32 x 32 block -- the bank key projection
// ran 380 us against 163 us without the rotation).  Tiled, the 32 lanes of a warp -- 32 consecutive positions -- read 512
// contiguous bytes per load.  Chunk j of the 16 columns starting at table column tcol0 (a multiple of 16):
__device__ __forceinline__ const float4* rope_chunk(const float* table, int pos, int tcol0, int j) {
  return reinterpret_cast<const float4*>(table + ((long long)(pos >> 5) * 32 + (tcol0 >> 2) + j) * 128 + (pos & 31) * 4);
}

// Fused epilogue of one 32-column block of one accumulator row: bias -> RoPE -> activation -> per-column scale ->
// residual (skipped when the caller adds a prefetched residual itself).
__device__ __forceinline__ void epilogue_block(float (&v)[32], const usvm_gemm_epilogue& ep, const int row,
                                               const bool row_ok, const long long rrow, const int col0, const int N,
                                               const bool add_residual, const float4* bias_pre = nullptr,
                                               const float4* rope_pre = nullptr) {
  const int ncol = min(32, N - col0);
  if (ncol == 32) {
    if (bias_pre) {  // fetched while the mainloop was still running
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        v[4 * j] += bias_pre[j].x; v[4 * j + 1] += bias_pre[j].y; v[4 * j + 2] += bias_pre[j].z; v[4 * j + 3] += bias_pre[j].w;
      }
    } else if (ep.bias) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 b = *reinterpret_cast<const float4*>(ep.bias + col0 + j);
        v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
      }
    }
    if (ep.rope_cos && col0 < ep.rope_cols) {
      const int rb = row % ep.rope_rows_per_batch;
      if (rb < ep.rope_n_rope) {
        const int pos = rb % ep.rope_table_rows, tcol0 = (col0 & 255) >> 1;
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          const float4 c4 = rope_pre ? rope_pre[j >> 3] : __ldg(rope_chunk(ep.rope_cos, pos, tcol0, j >> 3));
          const float4 s4 = rope_pre ? rope_pre[4 + (j >> 3)] : __ldg(rope_chunk(ep.rope_sin, pos, tcol0, j >> 3));
          const float cs[4] = {c4.x, c4.y, c4.z, c4.w}, sn[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float a = v[j + 2 * q], b = v[j + 2 * q + 1];
            v[j + 2 * q] = a * cs[q] - b * sn[q];
            v[j + 2 * q + 1] = a * sn[q] + b * cs[q];
          }
        }
      }
    }
    // (one branch per block, not per element: the 32 independent GELU chains then interleave -- with the switch inside
    // the loop every element was a serial chain of ~17 dependent instructions, 3400 cycles per block)
    if (ep.act == USVM_ACT_GELU) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
    } else if (ep.act == USVM_ACT_RELU) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.0f);
    }
    if (ep.col_scale) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 b = *reinterpret_cast<const float4*>(ep.col_scale + col0 + j);
        v[j] *= b.x; v[j + 1] *= b.y; v[j + 2] *= b.z; v[j + 3] *= b.w;
      }
    }
    if (add_residual && ep.residual && row_ok) {
      const float* r = ep.residual + rrow * ep.ldr + col0;
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 b = *reinterpret_cast<const float4*>(r + j);
        v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
      }
    }
  } else {  // ragged last column block: columns >= N are clipped by the TMA store, never read
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      if (j < ncol) {
        float x = v[j];
        if (ep.bias) x += ep.bias[col0 + j];
        x = apply_act(x, ep.act);
        if (ep.col_scale) x *= ep.col_scale[col0 + j];
        if (add_residual && ep.residual && row_ok) x += ep.residual[rrow * ep.ldr + col0 + j];
        v[j] = x;
      }
    }
  }
}

// TF32 = true: the same kernel on fp32 operands (tf32 tensor-core math, fp32 accumulate) -- a k-block is then 32
// floats (still one 128-byte swizzle row) and each tcgen05.mma covers K = 8.
template <int BN, bool TF32>
__global__ void __launch_bounds__(GEMM_THREADS, 2)
gemm_bf16_tc5_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const __grid_constant__ CUtensorMap tmO32, const __grid_constant__ CUtensorMap tmO16,
                     const usvm_gemm_epilogue ep, const int M, const int N, const int K, const int stages,
                     const int ksplit) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  using L = SmemLayout<BN>;
  uint8_t* staging = smem + stages * L::STAGE_BYTES;  // 1024-aligned: STAGE_BYTES is a multiple of 1024
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + STG_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
  // split-K (ksplit CTAs of one cluster share an output tile, blockIdx.z = k-slice): the leader (slice 0) receives the
  // other slices' accumulators here through distributed shared memory; [slice - 1][column / 4][row] float4
  float4* partials = reinterpret_cast<float4*>(staging + STG_BYTES + 256);  // past the barrier block
  const int kslice = blockIdx.z;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tile_m = blockIdx.x;
  const int tile_n = blockIdx.y;
  constexpr int BKE = TF32 ? BK / 2 : BK;  // elements per k-block (128 bytes)
  const int num_kb_all = (K + BKE - 1) / BKE;
  const int kb_per = (num_kb_all + ksplit - 1) / ksplit;
  const int kb_lo = kslice * kb_per;
  const int num_kb = max(0, min(num_kb_all, kb_lo + kb_per) - kb_lo);  // k-blocks of this CTA

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (ep.out_f32) tma_prefetch_desc(&tmO32);
    if (ep.out_bf16) tma_prefetch_desc(&tmO16);
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    mbar_fence_init();
  }
  if (warp == 1) tc5_alloc(tmem_slot, tmem_cols(BN));
  tc5_fence_before();
  __syncthreads();
  tc5_fence_after();
  // everything above (barriers, descriptor prefetch, TMEM) overlapped the previous kernel's tail
  pdl_wait();
  pdl_trigger();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // (whole warp walks the loop, one elected lane issues: see elect_one() in common.cuh)
    for (int kb = 0; kb < num_kb; ++kb) {
      const int s = kb % stages;
      const uint32_t ph = (kb / stages) & 1;
      mbar_wait(&empty_bar[s], ph ^ 1);
      if (elect_one()) {
        uint8_t* a_dst = smem + s * L::STAGE_BYTES;
        uint8_t* b_dst = a_dst + L::A_BYTES;
        mbar_arrive_expect_tx(&full_bar[s], L::STAGE_BYTES);
        tma_load_2d(a_dst, &tmA, &full_bar[s], (kb_lo + kb) * BKE, tile_m * BM);
        tma_load_2d(b_dst, &tmB, &full_bar[s], (kb_lo + kb) * BKE, tile_n * BN);
      }
      __syncwarp();
    }
    if (ksplit > 1) cluster_sync_all();
  } else if (warp == 1) {
    constexpr uint32_t idesc = TF32 ? umma_idesc_tf32(BM, BN) : umma_idesc_bf16(BM, BN);
    for (int kb = 0; kb < num_kb; ++kb) {
      const int s = kb % stages;
      const uint32_t ph = (kb / stages) & 1;
      mbar_wait(&full_bar[s], ph);
      tc5_fence_after();
      const uint32_t a_addr = smem_u32(smem + s * L::STAGE_BYTES);
      const uint32_t b_addr = a_addr + L::A_BYTES;
      const uint64_t a_desc = umma_desc_k_sw128(a_addr);
      const uint64_t b_desc = umma_desc_k_sw128(b_addr);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          // advance 16 bf16 (8 tf32) = 32 B along K inside the 128 B swizzle row: +2 in the (addr >> 4) field
          if (TF32)
            tc5_mma_tf32(tmem_base, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc,
                         (kb > 0 || k > 0) ? 1u : 0u);
          else
            tc5_mma_f16(tmem_base, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc,
                        (kb > 0 || k > 0) ? 1u : 0u);
        }
        tc5_commit(&empty_bar[s]);  // frees this smem stage once the MMAs above have read it
        if (kb == num_kb - 1) tc5_commit(tmem_full_bar);  // accumulator complete
      }
      __syncwarp();
    }
    if (num_kb == 0 && elect_one()) tc5_commit(tmem_full_bar);
    __syncwarp();
    if (ksplit > 1) cluster_sync_all();
  } else {
    // ---- epilogue: thread <-> accumulator row; warp <-> 32-row slab written by TMA ----
    const int lane_grp = warp & 3;  // TMEM lanes [32*lane_grp, 32*lane_grp + 32) are visible to this warp
    const int row0 = tile_m * BM + lane_grp * 32;
    const int row = row0 + lane;
    const bool row_ok = row < M;
    const long long rrow = (ep.res_div > 0 ? ((long long)(row / ep.res_div) * ep.res_mod + row % ep.res_mod)
                                           : (ep.res_mod > 0 ? (row % ep.res_mod) : row));
    uint8_t* stg32 = staging + lane_grp * STG_F32;
    uint8_t* stg16 = staging + 4 * STG_F32 + lane_grp * STG_BF16;
    bool pending = false;  // this warp has a TMA store in flight that still reads its staging buffers
    // The residual segment of a column block does not depend on the accumulator: it is fetched while the mainloop
    // (first block) or the TMEM load (later blocks) is still in flight, which takes an L2 round trip off the tail.
    float4 rs[8];
    auto fetch_residual = [&](int c0, float4 (&dst)[8]) -> bool {
      const int col0 = tile_n * BN + c0;
      if (!(ep.residual != nullptr && row_ok && c0 < BN && col0 + 32 <= N)) return false;
      const float4* r = reinterpret_cast<const float4*>(ep.residual + rrow * ep.ldr + col0);
#pragma unroll
      for (int j = 0; j < 8; ++j) dst[j] = __ldg(r + j);
      return true;
    };
    bool have = fetch_residual(0, rs);
    // ... and so does the bias segment (a constant of the model: one L2 round trip less after the accumulator is ready)
    // (only the single-block tile, BN = 32 -- the latency-bound shapes; wider tiles have no registers to spare)
    constexpr bool PRE_BIAS = BN <= 64;   // bias of the first column block
    constexpr bool PRE_ROPE = BN == 32;   // rotary tables: single-block tile only
    float4 bs[PRE_BIAS ? 8 : 1];
    bool have_b = false;
    if constexpr (PRE_BIAS) {
      const int col0 = tile_n * BN;
      if (ep.bias != nullptr && col0 + 32 <= N) {
        const float4* bp = reinterpret_cast<const float4*>(ep.bias + col0);
#pragma unroll
        for (int j = 0; j < 8; ++j) bs[j] = __ldg(bp + j);
        have_b = true;
      }
    }
    // ... and the rotary tables of this row (constants too)
    float4 rp[PRE_ROPE ? 8 : 1];
    bool have_r = false;
    if constexpr (PRE_ROPE) {
      const int col0 = tile_n * BN;
      if (ep.rope_cos && col0 < ep.rope_cols && col0 + 32 <= N) {
        const int rb = row % ep.rope_rows_per_batch;
        if (rb < ep.rope_n_rope) {
          const int pos = rb % ep.rope_table_rows, tcol0 = (col0 & 255) >> 1;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            rp[j] = __ldg(rope_chunk(ep.rope_cos, pos, tcol0, j));
            rp[4 + j] = __ldg(rope_chunk(ep.rope_sin, pos, tcol0, j));
          }
          have_r = true;
        }
      }
    }
    // fused LayerNorm (full 256-wide rows in this tile): pass 1 below also sums the row and parks the epilogue result
    // back in TMEM; the statistics and the normalised bf16 output follow after the loop
    const bool ln_mode = BN == 256 && ep.ln_w != nullptr;
    float ln_sum = 0.f;
    mbar_wait(tmem_full_bar, 0);
    tc5_fence_after();
    if (ksplit > 1) {
      if (kslice > 0) {  // hand this slice's accumulator to the leader CTA of the cluster
        const uint32_t remote = mapa_shared(smem_u32(partials), 0);
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
          uint32_t acc[32];
          tc5_ld_32x32(tmem_base + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)c0, acc);
          tc5_wait_ld();
#pragma unroll
          for (int c = 0; c < 8; ++c)
            st_shared_cluster_v4(remote + (uint32_t)((((kslice - 1) * (BN / 4) + (c0 >> 2) + c) * BM + lane_grp * 32 + lane) * 16),
                                 acc[4 * c], acc[4 * c + 1], acc[4 * c + 2], acc[4 * c + 3]);
        }
      }
      tc5_fence_before();
      cluster_sync_all();  // release / acquire at cluster scope: the partials are visible to the leader
      tc5_fence_after();
    }
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      const int col0 = tile_n * BN + c0;
      if (col0 >= N || row0 >= M || kslice > 0) break;  // warp-uniform
      uint32_t acc[32];
      tc5_ld_32x32(tmem_base + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)c0, acc);
      float4 rn[8];
      const bool have_next = fetch_residual(c0 + 32, rn);
      tc5_wait_ld();
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]);
      for (int sl = 0; sl + 1 < ksplit; ++sl) {  // split-K: add the other slices (fixed order: deterministic)
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 pv = partials[((sl * (BN / 4) + (c0 >> 2) + c) * BM) + lane_grp * 32 + lane];
          v[4 * c] += pv.x; v[4 * c + 1] += pv.y; v[4 * c + 2] += pv.z; v[4 * c + 3] += pv.w;
        }
      }
      epilogue_block(v, ep, row, row_ok, rrow, col0, N, !have, (PRE_BIAS && have_b && c0 == 0) ? bs : nullptr,
                     (PRE_ROPE && have_r && c0 == 0) ? rp : nullptr);
      if (have) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          v[4 * j] += rs[j].x; v[4 * j + 1] += rs[j].y; v[4 * j + 2] += rs[j].z; v[4 * j + 3] += rs[j].w;
        }
      }
      have = have_next;
#pragma unroll
      for (int j = 0; j < 8; ++j) rs[j] = rn[j];
      if (ln_mode) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          ln_sum += v[j];
          acc[j] = __float_as_uint(v[j]);
        }
        tc5_st_32x32(tmem_base + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)c0, acc);
      }
      if (pending) {  // the previous block's TMA store must have finished reading the staging buffers
        if (lane == 0) tma_store_wait_read();
        __syncwarp();
      }
      if (ep.out_f32) {  // 128-byte rows, 16-byte chunk c of row r at r*128 + ((c ^ (r & 7)) * 16)
        uint8_t* prow = stg32 + lane * 128;
#pragma unroll
        for (int c = 0; c < 8; ++c)
          *reinterpret_cast<float4*>(prow + ((c ^ (lane & 7)) << 4)) =
              make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
      }
      if (ep.out_bf16 && !ln_mode) {  // 64-byte rows, no swizzle
        uint8_t* prow = stg16 + lane * 64;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint4 pk;
          pk.x = pack_bf16x2(v[8 * c], v[8 * c + 1]);
          pk.y = pack_bf16x2(v[8 * c + 2], v[8 * c + 3]);
          pk.z = pack_bf16x2(v[8 * c + 4], v[8 * c + 5]);
          pk.w = pack_bf16x2(v[8 * c + 6], v[8 * c + 7]);
          *reinterpret_cast<uint4*>(prow + (c << 4)) = pk;
        }
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) {
        if (ep.out_f32) tma_store_2d(&tmO32, stg32, col0, row0);
        if (ep.out_bf16 && !ln_mode) tma_store_2d(&tmO16, stg16, col0, row0);
        tma_store_commit();
      }
      pending = true;
    }
    if (ln_mode && row0 < M) {
      tc5_wait_st_all();
      const float mean = ln_sum * (1.0f / 256.0f);
      float sq = 0.f;
#pragma unroll 1
      for (int c0 = 0; c0 < 256; c0 += 32) {
        uint32_t acc[32];
        tc5_ld_32x32(tmem_base + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)c0, acc);
        tc5_wait_ld();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float d = __uint_as_float(acc[j]) - mean;
          sq = fmaf(d, d, sq);
        }
      }
      const float rstd = 1.0f / sqrtf(sq * (1.0f / 256.0f) + ep.ln_eps);
#pragma unroll 1
      for (int c0 = 0; c0 < 256; c0 += 32) {
        uint32_t acc[32];
        tc5_ld_32x32(tmem_base + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)c0, acc);
        float4 lw[8], lb[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          lw[j] = __ldg(reinterpret_cast<const float4*>(ep.ln_w + c0) + j);
          lb[j] = __ldg(reinterpret_cast<const float4*>(ep.ln_b + c0) + j);
        }
        tc5_wait_ld();
        float y[32];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          y[4 * j] = (__uint_as_float(acc[4 * j]) - mean) * rstd * lw[j].x + lb[j].x;
          y[4 * j + 1] = (__uint_as_float(acc[4 * j + 1]) - mean) * rstd * lw[j].y + lb[j].y;
          y[4 * j + 2] = (__uint_as_float(acc[4 * j + 2]) - mean) * rstd * lw[j].z + lb[j].z;
          y[4 * j + 3] = (__uint_as_float(acc[4 * j + 3]) - mean) * rstd * lw[j].w + lb[j].w;
        }
        if (ep.ln_gelu) {
#pragma unroll
          for (int j = 0; j < 32; ++j) y[j] = gelu_erf(y[j]);
        }
        if (pending) {
          if (lane == 0) tma_store_wait_read();
          __syncwarp();
        }
        uint8_t* prow = stg16 + lane * 64;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint4 pk;
          pk.x = pack_bf16x2(y[8 * c], y[8 * c + 1]);
          pk.y = pack_bf16x2(y[8 * c + 2], y[8 * c + 3]);
          pk.z = pack_bf16x2(y[8 * c + 4], y[8 * c + 5]);
          pk.w = pack_bf16x2(y[8 * c + 6], y[8 * c + 7]);
          *reinterpret_cast<uint4*>(prow + (c << 4)) = pk;
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&tmO16, stg16, tile_n * BN + c0, row0);
          tma_store_commit();
        }
        pending = true;
      }
    }
    if (pending && lane == 0) tma_store_wait_read();  // smem must outlive the bulk reads
    __syncwarp();
  }
  tc5_fence_before();
  __syncthreads();
  if (warp == 1) tc5_dealloc(tmem_base, tmem_cols(BN));
}

// ---------------------------------------------------------------------------------------------------------------
// Persistent variant for the throughput-bound GEMMs (the batched Hiera encoder, the memory-bank K/V projections, the
// batched memory attention): one CTA per SM walks output tiles of 128 x BN (BN a runtime multiple of 32, <= 256, chosen to
// divide N), the accumulator is double-buffered in TMEM (columns [0,256) and [256,512)), and eight epilogue warps drain
// tile i while the TMA / MMA warps are already running tile i+1.
//   warp 0    : TMA producer, one continuous smem ring across tiles
//   warp 1    : MMA issuer; waits acc_empty[buf], issues K/16 tcgen05.mma, commits acc_full[buf]
//   warps 2-9 : epilogue; warp w owns TMEM lanes 32*(w%4).. and every second 32-column block of the tile
// Two schedules:
//   * WEIGHT-STATIONARY (ws_groups > 0; chosen whenever the [BN, K] slab of W fits next to a 3-stage A ring -- every
//     Hiera / memory-attention GEMM with K <= 384): CTA c keeps column slab c % tiles_n of W resident in shared memory and
//     walks row tiles c / tiles_n, + ws_groups, ...; only A streams.  Measured per-tile timeline of the streaming
//     schedule (tools/pgemm_timeline.py, 128 x 192 x 384): the 24 MMAs of a tile took 5970 cycles against 2304 at the
//     tensor pipe's rate, because A + W k-blocks (40 KB per 384 MMA cycles = 107 B/clk) exceed what one SM draws from L2
//     (~55 B/clk sustained with all SMs loading); with W resident the feed is 16 KB per k-block = 43 B/clk.
//   * streaming (ws_groups == 0): tiles n-fastest so the CTAs that run concurrently share the same A rows in L2.
// Epilogue: a 32 x 32 block per warp per step.  TMEM reads are the floor (64 B/clk per SM: 128 x BN x 4 B / 64 cycles per
// tile), so everything else overlaps them: the tcgen05.ld of block i+1 is issued before the math of block i, the
// accumulator is handed back to the MMA warp right after the tile's last TMEM read (not after its last store), bias and
// column scale of the warp's blocks sit in a per-warp shared-memory strip (filled before the accumulator is awaited, read
// back as broadcast 16-byte loads), residual / rotary operands are fetched one block ahead, and the block goes out
// through double-buffered staging + one TMA store.
// ---------------------------------------------------------------------------------------------------------------
// USVM2_PGEMM_DEBUG & 4: CTA 0 records clock64() stamps of its first 64 tiles here (usvm_debug_pgemm_profile reads them):
// [tile][0] MMA: accumulator free, [1] MMA: first stage full, [2] MMA: tile committed; [3] epilogue warp 0: accumulator
// full, [4] first TMEM load done, [5] first block's math done, [6] last block handed to TMA, [7] accumulator released
__device__ unsigned long long g_pgemm_prof[64 * 16];

// ---- CTA-pair (cta_group::2) forms: the two CTAs of a cluster run ONE 256 x BN tile; each holds its 128 rows of A and
// HALF of the W rows in its own shared memory (the tensor pipes read the peer's half), its 128 accumulator rows in its own
// TMEM.  The leader (rank 0) issues the MMAs; TMA loads of both CTAs complete on the leader's barrier; commits are
// multicast to both CTAs' barriers.
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc5_mma_f16_2sm(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                                uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the barrier at the same shared-memory offset in BOTH CTAs of the pair
__device__ __forceinline__ void tc5_commit_2sm(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void tc5_alloc_2sm(uint32_t* smem_slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc5_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar) : "memory");
}
constexpr int P_EPI_WARPS = 12;  // three per scheduler (14 warps -> 128 registers each): the epilogue is a chain of
                                 // long-latency steps, warps are what hides them
constexpr int P_COL_GROUPS = P_EPI_WARPS / 4;  // a warp owns every P_COL_GROUPS-th 32-column block of its 32 rows
constexpr int P_MAX_BLOCKS = (256 / 32 + P_COL_GROUPS - 1) / P_COL_GROUPS;
constexpr int P_THREADS = 64 + 32 * P_EPI_WARPS;
constexpr int P_STRIP_BYTES = P_EPI_WARPS * 2 * P_MAX_BLOCKS * 32 * 4;  // per warp: bias + column scale of its blocks
// staging: per epilogue warp TWO buffers of one 32 x 32 block per output present (fp32 4 KB, bf16 2 KB): the TMA store of
// block i still reads its buffer while block i + 1 is computed and written (single-buffered, a warp's second block waited
// ~1.5 k cycles for the first one's store to finish reading shared memory)
__host__ __device__ constexpr int p_stg_per_buf(bool f32, bool b16) { return (f32 ? STG_F32 : 0) + (b16 ? STG_BF16 : 0); }
// (the host picks one buffer instead when two would cost the W slab its residency, or for fp32 outputs: 96 KB)
__host__ __device__ constexpr int p_stg_bytes(bool f32, bool b16, int bufs) { return P_EPI_WARPS * bufs * p_stg_per_buf(f32, b16); }
constexpr int P_A_BYTES = BM * BK * 2;

// (bn = W rows held by ONE CTA: the tile width, or half of it for a CTA pair)
__host__ __device__ constexpr int p_stage_bytes(int bn, bool ws) { return P_A_BYTES + (ws ? 0 : bn * BK * 2); }
__host__ __device__ constexpr int p_slab_bytes(int bn, int num_kb, bool ws) { return ws ? num_kb * bn * BK * 2 : 0; }
__host__ __device__ constexpr int p_smem_total(int bn, int num_kb, int stages, bool f32, bool b16, bool ws, int bufs) {
  return p_slab_bytes(bn, num_kb, ws) + stages * p_stage_bytes(bn, ws) + p_stg_bytes(f32, b16, bufs) + P_STRIP_BYTES +
         1024 /* alignment slack */ + 256 /* barriers */;
}

// TF32: fp32 operands read as tf32 (a k-block is then 32 floats, each tcgen05.mma covers K = 8) -- the batched mask decoder
// PF = false: the variant for epilogues without residual / rotary operands -- the registers their one-block-ahead prefetch
// would hold carry a second accumulator block instead (two column blocks per iteration, see the epilogue loop)
template <bool PAIR, bool TF32 = false, bool PF = true>
__global__ void __launch_bounds__(P_THREADS, 1)
gemm_bf16_tc5_persistent_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                const __grid_constant__ CUtensorMap tmO32, const __grid_constant__ CUtensorMap tmO16,
                                const usvm_gemm_epilogue ep, const int M, const int N, const int K, const int BN,
                                const int stages_flags, const int tiles_n, const int tiles_m, const int ws_groups) {
  const int stages = stages_flags & 0xf;
  const int stg_bufs = (stages_flags >> 4) & 3;  // staging buffers per epilogue warp: 1 or 2
  const int dbg = stages_flags >> 8;  // experiment switches (USVM2_PGEMM_DEBUG): 1 = no TMA store, 2 = no staging either
  const bool ws = ws_groups > 0;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int BKE = TF32 ? BK / 2 : BK;  // elements per k-block (128 bytes)
  const int num_kb = (K + BKE - 1) / BKE;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;      // CTA pair: 0 = leader (issues the MMAs)
  const int bnl = PAIR ? BN / 2 : BN;                       // W rows this CTA holds
  const int tile_rows = PAIR ? 2 * BM : BM;                 // output rows of one tile
  const int cta_id = PAIR ? (int)blockIdx.x >> 1 : (int)blockIdx.x;      // scheduling unit: CTA or CTA pair
  const int num_units = PAIR ? (int)gridDim.x >> 1 : (int)gridDim.x;
  const int b_bytes = bnl * BK * 2;  // one k-block of this CTA's part of the W slab / tile (a multiple of 1024)
  const int stage_bytes = p_stage_bytes(bnl, ws);
  uint8_t* slab = smem;                                   // weight-stationary: num_kb k-blocks of W, resident
  uint8_t* ring = smem + p_slab_bytes(bnl, num_kb, ws);   // A (and, streaming, W) stages
  uint8_t* staging = ring + stages * stage_bytes;
  const int stg_per_buf = p_stg_per_buf(ep.out_f32 != nullptr, ep.out_bf16 != nullptr);
  float* strips = reinterpret_cast<float*>(staging + P_EPI_WARPS * stg_bufs * stg_per_buf);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(strips) + P_STRIP_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* acc_full = empty_bar + STAGES;
  uint64_t* acc_empty = acc_full + 2;
  uint64_t* slab_full = acc_empty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(slab_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // this CTA's tiles: i-th tile -> (tile_m, tile_n)
  const int my_n = ws ? cta_id % tiles_n : 0;
  const int my_group = ws ? cta_id / tiles_n : 0;
  const int num_tiles = tiles_m * tiles_n;
  const int my_count = ws ? (tiles_m - my_group + ws_groups - 1) / ws_groups
                          : (num_tiles - cta_id + num_units - 1) / num_units;
  auto tile_at = [&](int i, int& tile_m, int& tile_n) {
    if (ws) {
      tile_m = my_group + i * ws_groups;
      tile_n = my_n;
    } else {
      const int tile = cta_id + i * num_units;
      tile_m = tile / tiles_n;
      tile_n = tile - tile_m * tiles_n;
    }
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (ep.out_f32) tma_prefetch_desc(&tmO32);
    if (ep.out_bf16) tma_prefetch_desc(&tmO16);
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], PAIR ? 2 * P_EPI_WARPS : P_EPI_WARPS);  // the leader's is released by both CTAs' epilogues
    }
    mbar_init(slab_full, 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    if (PAIR) tc5_alloc_2sm(tmem_slot, 512);
    else tc5_alloc(tmem_slot, 512);
  }
  tc5_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();  // both CTAs' barriers are initialised before either signals the other's
  tc5_fence_after();
  pdl_wait();
  pdl_trigger();
  const uint32_t tmem_base = *tmem_slot;
  // (pair) barriers of the LEADER, as shared::cluster addresses: TMA loads of both CTAs complete there, both epilogues release there
  const uint32_t leader_slab_full = PAIR ? mapa_shared(smem_u32(slab_full), 0) : 0u;

  if (warp == 0) {
    if (ws && my_count > 0 && elect_one()) {  // the W slab of this CTA (pair: this CTA's half of it), once
      if (PAIR) {
        if (rank == 0) mbar_arrive_expect_tx(slab_full, (uint32_t)(2 * num_kb * b_bytes));
        for (int kb = 0; kb < num_kb; ++kb)
          tma_load_2d_2sm(slab + kb * b_bytes, &tmB, leader_slab_full, kb * BKE, my_n * BN + (int)rank * bnl);
      } else {
        mbar_arrive_expect_tx(slab_full, (uint32_t)(num_kb * b_bytes));
        for (int kb = 0; kb < num_kb; ++kb) tma_load_2d(slab + kb * b_bytes, &tmB, slab_full, kb * BKE, my_n * BN);
      }
    }
    __syncwarp();
    uint32_t it = 0;
    for (int i = 0; i < my_count; ++i) {
      int tile_m, tile_n;
      tile_at(i, tile_m, tile_n);
      for (int kb = 0; kb < num_kb; ++kb, ++it) {
        const uint32_t s = it % (uint32_t)stages;
        const uint32_t ph = (it / (uint32_t)stages) & 1u;
        mbar_wait(&empty_bar[s], ph ^ 1u);
        if (elect_one()) {
          uint8_t* a_dst = ring + s * stage_bytes;
          if (PAIR) {
            const uint32_t leader_full = mapa_shared(smem_u32(&full_bar[s]), 0);
            if (rank == 0) mbar_arrive_expect_tx(&full_bar[s], (uint32_t)(2 * stage_bytes));
            tma_load_2d_2sm(a_dst, &tmA, leader_full, kb * BKE, tile_m * tile_rows + (int)rank * BM);
            if (!ws) tma_load_2d_2sm(a_dst + P_A_BYTES, &tmB, leader_full, kb * BKE, tile_n * BN + (int)rank * bnl);
          } else if (dbg & 16) {  // experiment: no operand traffic at all (the MMAs run on whatever the ring holds)
            mbar_arrive(&full_bar[s]);
          } else {
            mbar_arrive_expect_tx(&full_bar[s], (uint32_t)stage_bytes);
            tma_load_2d(a_dst, &tmA, &full_bar[s], kb * BKE, tile_m * BM);
            if (!ws) tma_load_2d(a_dst + P_A_BYTES, &tmB, &full_bar[s], kb * BKE, tile_n * BN);
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 1 && rank == 0) {
    // the whole warp walks the loop (uniform control flow, waits included); one elected lane issues
    const uint32_t idesc = TF32 ? umma_idesc_tf32(tile_rows, BN) : umma_idesc_bf16(tile_rows, BN);
    uint32_t it = 0;
    if (ws && my_count > 0) {
      mbar_wait(slab_full, 0);
      tc5_fence_after();
    }
    for (int lt = 0; lt < my_count; ++lt) {
      const uint32_t buf = lt & 1u;
      mbar_wait(&acc_empty[buf], ((lt >> 1) & 1u) ^ 1u);  // the epilogue has drained this accumulator
      tc5_fence_after();
      const bool prof = (dbg & 4) && blockIdx.x == 0 && lt < 64 && lane == 0;
      if (prof) g_pgemm_prof[lt * 16 + 0] = clock64();
      const uint32_t d_tmem = tmem_base + buf * 256u;
      for (int kb = 0; kb < num_kb; ++kb, ++it) {
        const uint32_t s = it % (uint32_t)stages;
        const uint32_t ph = (it / (uint32_t)stages) & 1u;
        mbar_wait(&full_bar[s], ph);
        tc5_fence_after();
        if (prof && kb == 0) g_pgemm_prof[lt * 16 + 1] = clock64();
        const uint32_t a_addr = smem_u32(ring + s * stage_bytes);
        const uint64_t a_desc = umma_desc_k_sw128(a_addr);
        const uint64_t b_desc = umma_desc_k_sw128(ws ? smem_u32(slab + kb * b_bytes) : a_addr + P_A_BYTES);
        if (elect_one()) {
          if (PAIR) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              tc5_mma_f16_2sm(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc,
                              (kb > 0 || k > 0) ? 1u : 0u);
            tc5_commit_2sm(&empty_bar[s]);
            if (kb == num_kb - 1) tc5_commit_2sm(&acc_full[buf]);
          } else {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              if (TF32)
                tc5_mma_tf32(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc,
                             (kb > 0 || k > 0) ? 1u : 0u);
              else
                tc5_mma_f16(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc,
                            (kb > 0 || k > 0) ? 1u : 0u);
            }
            tc5_commit(&empty_bar[s]);
            if (kb == num_kb - 1) tc5_commit(&acc_full[buf]);
          }
        }
        __syncwarp();
      }
      if (prof) g_pgemm_prof[lt * 16 + 2] = clock64();
    }
  } else if (warp == 1) {
    // (pair) the peer's MMA warp has nothing to issue
  } else {
    const int ew = warp - 2;        // 0..P_EPI_WARPS-1
    const int lane_grp = warp & 3;  // TMEM lanes [32*lane_grp, +32) are the ones this warp may read
    const int quarter = ew >> 2;    // which 32-column blocks of the tile: quarter, quarter + P_COL_GROUPS, ...
    uint8_t* stg_warp = staging + ew * stg_bufs * stg_per_buf;  // one or two buffers: [fp32 block | bf16 block] each
    const int stg16_off = ep.out_f32 ? STG_F32 : 0;
    uint32_t nstored = 0;  // column blocks this warp has handed to TMA so far (buffer = nstored & 1)
    bool full_wait = false;  // a paired store group (both buffers) is the most recent one
    float* strip_bias = strips + ew * (2 * P_MAX_BLOCKS * 32);   // [blocks][32 columns]
    float* strip_scale = strip_bias + P_MAX_BLOCKS * 32;
    int strip_n = -1;      // tile_n the strips currently hold
    for (int lt = 0; lt < my_count; ++lt) {
      int tile_m, tile_n;
      tile_at(lt, tile_m, tile_n);
      const uint32_t buf = lt & 1u;
      // bias / column scale of this warp's blocks -> its strip, before the accumulator is awaited
      if (tile_n != strip_n && (ep.bias != nullptr || ep.col_scale != nullptr)) {
        __syncwarp();
#pragma unroll
        for (int b = 0; b < P_MAX_BLOCKS; ++b) {
          const int col = tile_n * BN + (quarter + b * P_COL_GROUPS) * 32 + lane;
          const bool ok = (quarter + b * P_COL_GROUPS) * 32 < BN && col < N;
          if (ep.bias) strip_bias[b * 32 + lane] = ok ? __ldg(ep.bias + col) : 0.f;
          if (ep.col_scale) strip_scale[b * 32 + lane] = ok ? __ldg(ep.col_scale + col) : 0.f;
        }
        __syncwarp();
        strip_n = tile_n;
      }
      const int row0 = tile_m * tile_rows + (int)rank * BM + lane_grp * 32;
      const int row = row0 + lane;
      const bool row_ok = row < M;
      const long long rrow = (ep.res_div > 0 ? ((long long)(row / ep.res_div) * ep.res_mod + row % ep.res_mod)
                                           : (ep.res_mod > 0 ? (row % ep.res_mod) : row));
      // Operands of a column block that do not depend on the accumulator -- the residual row segment, or the rotary
      // cos / sin of this row (the two never occur together) -- are requested BEFORE the accumulator is awaited / read.
      const bool rope_row = ep.rope_cos != nullptr && (row % ep.rope_rows_per_batch) < ep.rope_n_rope;
      const int rope_pos = rope_row ? (row % ep.rope_rows_per_batch) % ep.rope_table_rows : 0;
      float4 cur[8];
      // kind: 0 nothing, 1 residual, 2 rotary tables
      auto prefetch = [&](int c0) -> int {
        if constexpr (!PF) return 0;
        const int col0 = tile_n * BN + c0;
        if (c0 >= BN || col0 + 32 > N || row0 >= M) return 0;
        if (ep.residual != nullptr && row_ok) {
          const float4* r = reinterpret_cast<const float4*>(ep.residual + rrow * ep.ldr + col0);
#pragma unroll
          for (int j = 0; j < 8; ++j) cur[j] = __ldg(r + j);
          return 1;
        }
        if (rope_row && col0 < ep.rope_cols) {
          const int tcol0 = (col0 & 255) >> 1;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            cur[j] = __ldg(rope_chunk(ep.rope_cos, rope_pos, tcol0, j));
            cur[4 + j] = __ldg(rope_chunk(ep.rope_sin, rope_pos, tcol0, j));
          }
          return 2;
        }
        return 0;
      };
      int kind = prefetch(quarter * 32);
      mbar_wait(&acc_full[buf], (lt >> 1) & 1u);
      tc5_fence_after();
      const bool prof = (dbg & 4) && blockIdx.x == 0 && lt < 64 && ew == 0 && lane == 0;
      if (prof) g_pgemm_prof[lt * 16 + 3] = clock64();
      bool released = false;
      if (row0 < M && !(dbg & 8)) {  // (dbg & 8: experiment, the accumulator is released unread)
        const uint32_t t_lane = tmem_base + buf * 256u + ((uint32_t)(lane_grp * 32) << 16);
        int blk = 0;
#pragma unroll 1
        for (int c0 = quarter * 32; c0 < BN; c0 += 32 * P_COL_GROUPS, ++blk) {
          const int col0 = tile_n * BN + c0;
          if (col0 >= N) break;  // warp-uniform
          if constexpr (!PF) {
            // Two column blocks at a time (the warp has two staging buffers): both TMEM loads are issued before the wait and
            // the pair shares one staging wait, one proxy fence and one bulk-store group -- the per-block chain (load
            // latency, fence, commit), not bandwidth, bounds the output-heavy GEMMs (tools/pgemm_timeline.py).
            const int c1 = c0 + 32 * P_COL_GROUPS, col1 = tile_n * BN + c1;
            if (stg_bufs == 2 && !(dbg & 6) && c1 < BN && col1 + 32 <= N) {
              uint32_t acc[32], acc2[32];
              tc5_ld_32x32(t_lane + (uint32_t)c0, acc);
              tc5_ld_32x32(t_lane + (uint32_t)c1, acc2);
              tc5_wait_ld();
              if (!(c1 + 32 * P_COL_GROUPS < BN && col1 + 32 * P_COL_GROUPS < N)) {  // the warp's last reads of this tile
                tc5_fence_before();
                __syncwarp();
                if (lane == 0) {
                  if (PAIR) mbar_arrive_cluster(mapa_shared(smem_u32(&acc_empty[buf]), 0));
                  else mbar_arrive(&acc_empty[buf]);
                }
                released = true;
              }
              float v[32], u[32];
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                v[j] = __uint_as_float(acc[j]);
                u[j] = __uint_as_float(acc2[j]);
              }
              if (ep.bias) {
                const float4* bs = reinterpret_cast<const float4*>(strip_bias + blk * 32);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const float4 b = bs[j], b2 = bs[8 + j];
                  v[4 * j] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
                  u[4 * j] += b2.x; u[4 * j + 1] += b2.y; u[4 * j + 2] += b2.z; u[4 * j + 3] += b2.w;
                }
              }
              if (ep.act == USVM_ACT_GELU) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                  v[j] = gelu_erf(v[j]);
                  u[j] = gelu_erf(u[j]);
                }
              } else if (ep.act == USVM_ACT_RELU) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                  v[j] = fmaxf(v[j], 0.0f);
                  u[j] = fmaxf(u[j], 0.0f);
                }
              }
              if (ep.col_scale) {
                const float4* ss = reinterpret_cast<const float4*>(strip_scale + blk * 32);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const float4 b = ss[j], b2 = ss[8 + j];
                  v[4 * j] *= b.x; v[4 * j + 1] *= b.y; v[4 * j + 2] *= b.z; v[4 * j + 3] *= b.w;
                  u[4 * j] *= b2.x; u[4 * j + 1] *= b2.y; u[4 * j + 2] *= b2.z; u[4 * j + 3] *= b2.w;
                }
              }
              if (nstored) {  // every earlier store of this warp has finished reading both staging buffers
                if (elect_one()) tma_store_wait_read();
                __syncwarp();
              }
              uint8_t* sa32 = stg_warp;
              uint8_t* sb32 = stg_warp + stg_per_buf;
              if (ep.out_f32) {
                uint8_t* pa = sa32 + lane * 128;
                uint8_t* pb = sb32 + lane * 128;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                  *reinterpret_cast<float4*>(pa + ((c ^ (lane & 7)) << 4)) =
                      make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
                  *reinterpret_cast<float4*>(pb + ((c ^ (lane & 7)) << 4)) =
                      make_float4(u[4 * c], u[4 * c + 1], u[4 * c + 2], u[4 * c + 3]);
                }
              }
              if (ep.out_bf16) {
                uint8_t* pa = sa32 + stg16_off + lane * 64;
                uint8_t* pb = sb32 + stg16_off + lane * 64;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                  uint4 pk, pk2;
                  pk.x = pack_bf16x2(v[8 * c], v[8 * c + 1]);
                  pk.y = pack_bf16x2(v[8 * c + 2], v[8 * c + 3]);
                  pk.z = pack_bf16x2(v[8 * c + 4], v[8 * c + 5]);
                  pk.w = pack_bf16x2(v[8 * c + 6], v[8 * c + 7]);
                  pk2.x = pack_bf16x2(u[8 * c], u[8 * c + 1]);
                  pk2.y = pack_bf16x2(u[8 * c + 2], u[8 * c + 3]);
                  pk2.z = pack_bf16x2(u[8 * c + 4], u[8 * c + 5]);
                  pk2.w = pack_bf16x2(u[8 * c + 6], u[8 * c + 7]);
                  *reinterpret_cast<uint4*>(pa + (c << 4)) = pk;
                  *reinterpret_cast<uint4*>(pb + (c << 4)) = pk2;
                }
              }
              fence_async_smem();
              __syncwarp();
              if (!(dbg & 1) && elect_one()) {
                if (ep.out_f32) {
                  tma_store_2d(&tmO32, sa32, col0, row0);
                  tma_store_2d(&tmO32, sb32, col1, row0);
                }
                if (ep.out_bf16) {
                  tma_store_2d(&tmO16, sa32 + stg16_off, col0, row0);
                  tma_store_2d(&tmO16, sb32 + stg16_off, col1, row0);
                }
                tma_store_commit();
              }
              __syncwarp();
              nstored += 2;
              full_wait = true;  // the next single block must not reuse a buffer this group may still be reading
              c0 = c1;
              ++blk;
              continue;
            }
          }
          uint32_t acc[32];
          tc5_ld_32x32(t_lane + (uint32_t)c0, acc);
          if (blk > 0) kind = prefetch(c0);
          tc5_wait_ld();
          if (!(c0 + 32 * P_COL_GROUPS < BN && col0 + 32 * P_COL_GROUPS < N)) {
            // the tile's last TMEM read of this warp is complete: hand the accumulator back before the math / stores
            tc5_fence_before();
            __syncwarp();
            if (lane == 0) {
              if (PAIR) mbar_arrive_cluster(mapa_shared(smem_u32(&acc_empty[buf]), 0));
              else mbar_arrive(&acc_empty[buf]);
            }
            released = true;
          }
          if (prof && blk == 0) g_pgemm_prof[lt * 16 + 4] = clock64();
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]);
          if (N - col0 >= 32) {
            if (ep.bias) {
              const float4* bs = reinterpret_cast<const float4*>(strip_bias + blk * 32);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 b = bs[j];
                v[4 * j] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
              }
            }
            if (kind == 2) {
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                const float4 c4 = cur[j >> 3], s4 = cur[4 + (j >> 3)];
                const float cs[4] = {c4.x, c4.y, c4.z, c4.w}, sn[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  const float a = v[j + 2 * q], b = v[j + 2 * q + 1];
                  v[j + 2 * q] = a * cs[q] - b * sn[q];
                  v[j + 2 * q + 1] = a * sn[q] + b * cs[q];
                }
              }
            }
            if (ep.act == USVM_ACT_GELU) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
            } else if (ep.act == USVM_ACT_RELU) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.0f);
            }
            if (ep.col_scale) {
              const float4* ss = reinterpret_cast<const float4*>(strip_scale + blk * 32);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 b = ss[j];
                v[4 * j] *= b.x; v[4 * j + 1] *= b.y; v[4 * j + 2] *= b.z; v[4 * j + 3] *= b.w;
              }
            }
            if (kind == 1) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                v[4 * j] += cur[j].x; v[4 * j + 1] += cur[j].y; v[4 * j + 2] += cur[j].z; v[4 * j + 3] += cur[j].w;
              }
            }
          } else {  // ragged last column block (no rotary encoding there: N % 32 == 0 is required with rope)
            epilogue_block(v, ep, row, row_ok, rrow, col0, N, true);
          }
          if (prof && blk == 0) g_pgemm_prof[lt * 16 + 5] = clock64() + (v[0] == 1.2345e30f);
          // (Stores straight from registers -- 16-byte pieces at the row pitch -- were measured 15-80 % slower than
          // staging + one TMA store per block: partial-sector writes.)  Two staging buffers per warp: the TMA store of
          // block i still reads its buffer while block i + 1 is written.
          uint8_t* stg32 = stg_warp + (stg_bufs == 2 ? (nstored & 1u) : 0u) * stg_per_buf;
          uint8_t* stg16 = stg32 + stg16_off;
          if (nstored >= (uint32_t)stg_bufs || full_wait) {  // the store that last used this buffer has finished reading it
            if (elect_one()) {
              if (stg_bufs == 2 && !full_wait) tma_store_wait_read1();
              else tma_store_wait_read();
            }
            __syncwarp();
            full_wait = false;
          }
          if (prof && blk == 0) g_pgemm_prof[lt * 16 + 8] = clock64();
          if (ep.out_f32 && !(dbg & 2)) {
            uint8_t* prow = stg32 + lane * 128;
#pragma unroll
            for (int c = 0; c < 8; ++c)
              *reinterpret_cast<float4*>(prow + ((c ^ (lane & 7)) << 4)) =
                  make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
          }
          if (ep.out_bf16 && !(dbg & 2)) {
            uint8_t* prow = stg16 + lane * 64;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              uint4 pk;
              pk.x = pack_bf16x2(v[8 * c], v[8 * c + 1]);
              pk.y = pack_bf16x2(v[8 * c + 2], v[8 * c + 3]);
              pk.z = pack_bf16x2(v[8 * c + 4], v[8 * c + 5]);
              pk.w = pack_bf16x2(v[8 * c + 6], v[8 * c + 7]);
              *reinterpret_cast<uint4*>(prow + (c << 4)) = pk;
            }
          }
          if (dbg & 2) {  // keep the values alive
            float keep = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) keep += v[j];
            if (keep == 1.2345e30f) stg_warp[0] = 1;
          }
          if (prof && blk == 0) g_pgemm_prof[lt * 16 + 9] = clock64();
          fence_async_smem();
          __syncwarp();
          if (prof && blk == 0) g_pgemm_prof[lt * 16 + 10] = clock64();
          if (!(dbg & 1) && elect_one()) {  // (one elected lane of the converged warp: plain UTMASTG, no ELECT loop)
            if (ep.out_f32) tma_store_2d(&tmO32, stg32, col0, row0);
            if (ep.out_bf16) tma_store_2d(&tmO16, stg16, col0, row0);
            tma_store_commit();
          }
          __syncwarp();
          ++nstored;
          if (prof && blk == 0) g_pgemm_prof[lt * 16 + 11] = clock64();
        }
      }
      if (prof) g_pgemm_prof[lt * 16 + 6] = clock64();
      if (!released) {  // (row slab beyond M, or no column block of this warp inside N)
        tc5_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (PAIR) mbar_arrive_cluster(mapa_shared(smem_u32(&acc_empty[buf]), 0));
          else mbar_arrive(&acc_empty[buf]);
        }
      }
      if (prof) g_pgemm_prof[lt * 16 + 7] = clock64();
    }
    if (nstored && elect_one()) tma_store_wait_read();
    __syncwarp();
  }
  tc5_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();  // the peer's shared memory / TMEM stay alive until the pair is done
  if (warp == 1) {
    if (PAIR) tc5_dealloc_2sm(tmem_base, 512);
    else tc5_dealloc(tmem_base, 512);
  }
}

// ---- host side: tensor maps through the driver entry point (no link-time libcuda dependency) ----
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

// Eager launches (prompted frames, encoder batches outside a captured graph) describe the same few tensors over and over
// -- weights always, activations whenever the caching allocator hands the same block back: a small direct-mapped cache of
// encoded maps per host thread replaces the driver call (a tensor map is a pure function of these arguments).
struct MapKey {
  const void* base;
  long long rows, cols, ld;
  int dt, box_cols, box_rows, swz;
  bool operator==(const MapKey& o) const {
    return base == o.base && rows == o.rows && cols == o.cols && ld == o.ld && dt == o.dt && box_cols == o.box_cols &&
           box_rows == o.box_rows && swz == o.swz;
  }
};
struct MapSlot {
  MapKey key;
  bool valid;
  CUtensorMap map;
};
constexpr int MAP_CACHE_SLOTS = 1024;

// 2D row-major [rows, cols] with row pitch ld (elements); box = {box_cols, box_rows}
int make_map(CUtensorMap* map, CUtensorMapDataType dt, int esize, const void* base, long long rows, long long cols,
             long long ld, int box_cols, int box_rows, CUtensorMapSwizzle swz) {
  static thread_local MapSlot cache[MAP_CACHE_SLOTS] = {};
  const MapKey key{base, rows, cols, ld, (int)dt, box_cols, box_rows, (int)swz};
  unsigned long long h = reinterpret_cast<unsigned long long>(base) >> 4;
  h = (h ^ (unsigned long long)rows * 0x9E3779B97F4A7C15ull ^ (unsigned long long)cols * 0xC2B2AE3D27D4EB4Full ^
       (unsigned long long)(box_rows * 131 + box_cols * 7 + (int)dt)) * 0xFF51AFD7ED558CCDull;
  MapSlot& slot = cache[(h >> 40) % MAP_CACHE_SLOTS];
  if (slot.valid && slot.key == key) {
    *map = slot.map;
    return USVM_OK;
  }
  const int rc = [&]() -> int {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return USVM_ERR_DRIVER;
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * esize};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, dt, 2, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? USVM_OK : USVM_ERR_DRIVER;
  }();
  if (rc == USVM_OK) {
    slot.key = key;
    slot.map = *map;
    slot.valid = true;
  }
  return rc;
}
int make_map_bf16(CUtensorMap* map, const void* base, long long rows, long long cols, long long ld, int box_rows) {
  return make_map(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, rows, cols, ld, BK, box_rows,
                  CU_TENSOR_MAP_SWIZZLE_128B);
}

template <int BN, bool TF32>
int launch(const void* A, int lda, const void* W, int ldw, const usvm_gemm_epilogue* ep, int M, int N, int K,
           cudaStream_t stream, int ksplit = 1) {
  CUtensorMap tmA, tmB, tmO32, tmO16;
  int rc = TF32 ? make_map(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, A, M, K, lda, BK / 2, BM, CU_TENSOR_MAP_SWIZZLE_128B)
                : make_map_bf16(&tmA, A, M, K, lda, BM);
  if (rc) return rc;
  rc = TF32 ? make_map(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, W, N, K, ldw, BK / 2, BN, CU_TENSOR_MAP_SWIZZLE_128B)
            : make_map_bf16(&tmB, W, N, K, ldw, BN);
  if (rc) return rc;
  tmO32 = tmA;  // placeholders keep the kernel parameters valid when an output is absent
  tmO16 = tmA;
  if (ep->out_f32) {
    rc = make_map(&tmO32, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, ep->out_f32, M, N, ep->ldo_f32, 32, 32,
                  CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  if (ep->out_bf16) {
    rc = make_map(&tmO16, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ep->out_bf16, M, N, ep->ldo_bf16, 32, 32,
                  CU_TENSOR_MAP_SWIZZLE_NONE);
    if (rc) return rc;
  }
  static UsvmPerDeviceOnce attr_set = {};
  if (usvm_need_setup(attr_set)) {
    const int want = 227 * 1024;  // ring + staging (+ split-K partials) are sized per launch, always below this
    if (cudaFuncSetAttribute(gemm_bf16_tc5_kernel<BN, TF32>, cudaFuncAttributeMaxDynamicSharedMemorySize, want) !=
        cudaSuccess)
      return USVM_ERR_CUDA;
    usvm_setup_done(attr_set);
  }
  const int num_kb = cdiv(K, TF32 ? BK / 2 : BK);
  dim3 grid(cdiv(M, BM), cdiv(N, BN));
  // Ring depth: a problem with fewer CTAs than ~2 per SM is latency bound -> as deep as shared memory allows (the
  // k-loop then streams at TMA issue rate instead of TMA latency); big problems keep <= 4 stages so that 2-3 CTAs
  // share an SM and overlap each other's epilogues.
  const long long ctas = (long long)grid.x * grid.y;
  int max_stages = ctas * ksplit <= 2 * 148 ? STAGES : 4;
  const int part_bytes = ksplit > 1 ? 256 + (ksplit - 1) * BM * BN * 4 : 0;
  while (max_stages > 1 && SmemLayout<BN>::total(max_stages) + part_bytes > 200 * 1024) --max_stages;
  const int kb_per = cdiv(num_kb, ksplit);
  const int stages = kb_per < max_stages ? kb_per : max_stages;
  const size_t smem = SmemLayout<BN>::total(stages) + part_bytes;
  if (ksplit == 1) {
    usvm_launch(gemm_bf16_tc5_kernel<BN, TF32>, dim3(grid), dim3(GEMM_THREADS), smem, stream, tmA, tmB, tmO32, tmO16, *ep,
                M, N, K, stages, 1);
    return usvm_check_launch();
  }
  // split-K: the ksplit CTAs of an output tile form a thread-block cluster along z
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid.x, grid.y, ksplit);
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = ksplit;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = usvm_pdl_enabled() ? 2 : 1;
  if (cudaLaunchKernelEx(&cfg, gemm_bf16_tc5_kernel<BN, TF32>, tmA, tmB, tmO32, tmO16, *ep, M, N, K, stages, ksplit) !=
      cudaSuccess)
    return USVM_ERR_CUDA;
  return usvm_check_launch();
}

// SMs a persistent grid may occupy (0 = all): set while the image encoder is captured on an SM partition (green
// context) so that one CTA per *partition* SM walks the tiles instead of 148 CTAs queueing in waves.
int g_sm_budget = 0;

int pgemm_debug() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("USVM2_PGEMM_DEBUG");
    v = e ? atoi(e) : 0;
  }
  return v;
}

int persistent_block_n(int N) {
  // the widest multiple of 32 (<= 256) that wastes the fewest padded columns
  int best = 32, best_waste = 1 << 30;
  for (int bn = 32; bn <= 256; bn += 32) {
    const int waste = cdiv(N, bn) * bn - N;
    if (waste < best_waste || (waste == best_waste && bn > best)) {
      best = bn;
      best_waste = waste;
    }
  }
  return best;
}

int device_sm_count() {
  static int sm_count = 0;
  if (!sm_count) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sm_count <= 0)
      sm_count = 0;
  }
  return sm_count;
}

// weight-stationary groups for tile width bn (0: the slab does not fit / too few row tiles per scheduling unit).
// pair: the scheduling unit is a CTA pair (256-row tiles, each CTA holds bn / 2 rows of W), units = sm_limit / 2.
int ws_groups_for(int bn, int M, int N, int K, bool o32, bool o16, int sm_limit, bool pair, int* bufs_out = nullptr,
                  bool tf32 = false) {
  static const int ws_mode = [] { const char* e = getenv("USVM2_PGEMM_WS"); return e ? atoi(e) : 1; }();
  const int tiles_m = cdiv(M, pair ? 2 * BM : BM), tiles_n = cdiv(N, bn), num_kb = cdiv(K, tf32 ? BK / 2 : BK);
  const int units = pair ? sm_limit / 2 : sm_limit;
  if (!ws_mode || tiles_n > units) return 0;
  int bufs = o32 ? 1 : 2;
  if (bufs == 2 && p_smem_total(pair ? bn / 2 : bn, num_kb, 3, o32, o16, true, 2) > 227 * 1024) bufs = 1;
  if (p_smem_total(pair ? bn / 2 : bn, num_kb, 3, o32, o16, true, bufs) > 227 * 1024) return 0;
  if (bufs_out) *bufs_out = bufs;
  const int groups = units / tiles_n;
  return (groups >= 1 && tiles_m >= 2 * groups) ? groups : 0;
}

int pair_mode() {
  // opt-in (USVM2_PGEMM_PAIR=1).  Measured (tools/pgemm_timeline.py): with pairs a 256 x 192 x 16 MMA retires in ~120
  // cycles against 183 for the one-CTA 128 x 192 x 16 form -- the predicted 1.5x per-CTA MMA rate -- but at the batch
  // sizes of this path (16 frames, 32 objects) the tile period is then set by the epilogue (TMEM drains at 64 B/clk:
  // 1536 cycles per 128 x 192 fp32 tile, plus staging and stores) and by ramp / tail, and kernel times come out equal
  // (35.3 vs 32.3 us stage-3 qkv, 52.5 vs 52.2 us GELU MLP, 66.0 vs 66.4 us FFN linear2).
  static const int v = [] { const char* e = getenv("USVM2_PGEMM_PAIR"); return e ? atoi(e) : 0; }();
  return v;
}

// bn <= 0: choose -- the widest tile among those that waste the fewest padded columns, preferring one whose W slab can
// stay resident (weight-stationary schedule).  CTA pairs (cta_group::2, 256-row tiles) whenever the problem has at least
// two 256-row tiles per pair: each CTA then reads 4 KB of A + bn * 16 B of W per MMA instead of 4 KB + bn * 32 B.
int launch_persistent(const void* A, int lda, const void* W, int ldw, const usvm_gemm_epilogue* ep, int M, int N, int K,
                      int bn, cudaStream_t stream, bool tf32 = false) {
  const int sm_count = device_sm_count();
  if (!sm_count) return USVM_ERR_CUDA;
  const int sm_limit = (g_sm_budget > 0 && g_sm_budget < sm_count) ? g_sm_budget : sm_count;
  const bool o32 = ep->out_f32 != nullptr, o16 = ep->out_bf16 != nullptr;
  // enough work for every pair to see several 256-row tiles
  const bool pair_ok = !tf32 && pair_mode() && sm_limit >= 2 && (long long)cdiv(M, 2 * BM) * cdiv(N, 128) >= sm_limit;
  bool pair = false;
  if (bn <= 0) {
    // widest tile first: 256 (pair only: the slab halves fit), then the no-waste candidates
    bn = persistent_block_n(N);
    const int min_waste = cdiv(N, bn) * bn - N;
    if (pair_ok) {
      for (int b = 256; b >= 64; b -= 32)
        if ((b % 64) == 0 && cdiv(N, b) * b - N == min_waste && ws_groups_for(b, M, N, K, o32, o16, sm_limit, true)) {
          bn = b;
          pair = true;
          break;
        }
    }
    if (!pair && !ws_groups_for(bn, M, N, K, o32, o16, sm_limit, false, nullptr, tf32)) {
      for (int b = bn - 32; b >= 96; b -= 32)
        if (cdiv(N, b) * b - N == min_waste && ws_groups_for(b, M, N, K, o32, o16, sm_limit, false, nullptr, tf32)) {
          bn = b;
          break;
        }
    }
    if (!pair && pair_ok && (bn % 64) == 0) pair = true;  // streaming schedule on pairs
  } else {
    pair = pair_ok && (bn % 64) == 0;
  }
  if (bn > 256 || (bn % 32)) return USVM_ERR_ARG;
  const int bnl = pair ? bn / 2 : bn;
  CUtensorMap tmA, tmB, tmO32, tmO16;
  int rc = tf32 ? make_map(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, A, M, K, lda, BK / 2, BM, CU_TENSOR_MAP_SWIZZLE_128B)
                : make_map_bf16(&tmA, A, M, K, lda, BM);
  if (rc) return rc;
  rc = tf32 ? make_map(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, W, N, K, ldw, BK / 2, bnl, CU_TENSOR_MAP_SWIZZLE_128B)
            : make_map_bf16(&tmB, W, N, K, ldw, bnl);
  if (rc) return rc;
  tmO32 = tmA;
  tmO16 = tmA;
  if (ep->out_f32) {
    rc = make_map(&tmO32, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, ep->out_f32, M, N, ep->ldo_f32, 32, 32,
                  CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  if (ep->out_bf16) {
    rc = make_map(&tmO16, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ep->out_bf16, M, N, ep->ldo_bf16, 32, 32,
                  CU_TENSOR_MAP_SWIZZLE_NONE);
    if (rc) return rc;
  }
  const int tiles_m = cdiv(M, pair ? 2 * BM : BM), tiles_n = cdiv(N, bn), num_kb = cdiv(K, tf32 ? BK / 2 : BK);
  const int num_tiles = tiles_m * tiles_n;
  const int units = pair ? sm_limit / 2 : sm_limit;
  int bufs = o32 ? 1 : 2;
  const int ws_groups = ws_groups_for(bn, M, N, K, o32, o16, sm_limit, pair, &bufs, tf32);
  const bool ws = ws_groups > 0;
  int stages = STAGES;
  while (stages > 1 && p_smem_total(bnl, num_kb, stages, o32, o16, ws, bufs) > 227 * 1024) --stages;
  const int grid_units = ws ? ws_groups * tiles_n : (num_tiles < units ? num_tiles : units);
  const size_t smem = p_smem_total(bnl, num_kb, stages, o32, o16, ws, bufs);
  const int flags = stages | (bufs << 4) | (pgemm_debug() << 8);
  if (tf32) {
    if (cudaFuncSetAttribute(gemm_bf16_tc5_persistent_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             227 * 1024) != cudaSuccess)
      return USVM_ERR_CUDA;
    usvm_launch(gemm_bf16_tc5_persistent_kernel<false, true>, dim3(grid_units), dim3(P_THREADS), smem, stream, tmA, tmB,
                tmO32, tmO16, *ep, M, N, K, bn, flags, tiles_n, tiles_m, ws_groups);
    return usvm_check_launch();
  }
  if (!pair && ep->residual == nullptr && ep->rope_cos == nullptr) {  // no prefetched operands: two blocks per iteration
    if (cudaFuncSetAttribute(gemm_bf16_tc5_persistent_kernel<false, false, false>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
      return USVM_ERR_CUDA;
    usvm_launch(gemm_bf16_tc5_persistent_kernel<false, false, false>, dim3(grid_units), dim3(P_THREADS), smem, stream, tmA,
                tmB, tmO32, tmO16, *ep, M, N, K, bn, flags, tiles_n, tiles_m, ws_groups);
    return usvm_check_launch();
  }
  if (!pair) {
    if (cudaFuncSetAttribute(gemm_bf16_tc5_persistent_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             227 * 1024) != cudaSuccess)
      return USVM_ERR_CUDA;
    usvm_launch(gemm_bf16_tc5_persistent_kernel<false>, dim3(grid_units), dim3(P_THREADS), smem, stream, tmA, tmB, tmO32,
                tmO16, *ep, M, N, K, bn, flags, tiles_n, tiles_m, ws_groups);
    return usvm_check_launch();
  }
  if (cudaFuncSetAttribute(gemm_bf16_tc5_persistent_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           227 * 1024) != cudaSuccess)
    return USVM_ERR_CUDA;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(2 * grid_units);
  cfg.blockDim = dim3(P_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = usvm_pdl_enabled() ? 2 : 1;
  if (cudaLaunchKernelEx(&cfg, gemm_bf16_tc5_persistent_kernel<true>, tmA, tmB, tmO32, tmO16, *ep, M, N, K, bn, flags,
                         tiles_n, tiles_m, ws_groups) != cudaSuccess)
    return USVM_ERR_CUDA;
  return usvm_check_launch();
}

}  // namespace

extern "C" int usvm_debug_pgemm_profile(unsigned long long* host_out_1024) {
  if (!host_out_1024) return USVM_ERR_ARG;
  return cudaMemcpyFromSymbol(host_out_1024, g_pgemm_prof, sizeof(g_pgemm_prof)) == cudaSuccess ? USVM_OK : USVM_ERR_CUDA;
}

extern "C" int usvm_set_sm_budget(int sms) {
  g_sm_budget = sms > 0 ? sms : 0;
  return 0;
}

extern "C" int usvm_gemm_bf16_tc5(const void* A, int lda, const void* W, int ldw, const usvm_gemm_epilogue* ep,
                                  int M, int N, int K, int block_n, void* stream) {
  if (!A || !W || !ep || M <= 0 || N <= 0 || K <= 0) return USVM_ERR_ARG;
  if ((lda % 8) || (ldw % 8) || (reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(W) & 15))
    return USVM_ERR_ARG;  // TMA: 16-byte aligned base and row pitch
  if (ep->out_f32 && ((ep->ldo_f32 % 4) || (reinterpret_cast<uintptr_t>(ep->out_f32) & 15))) return USVM_ERR_ARG;
  if (ep->out_bf16 && ((ep->ldo_bf16 % 8) || (reinterpret_cast<uintptr_t>(ep->out_bf16) & 15))) return USVM_ERR_ARG;
  if (ep->residual && ((ep->ldr % 4) || (reinterpret_cast<uintptr_t>(ep->residual) & 15))) return USVM_ERR_ARG;
  if (ep->rope_cos && (!ep->rope_sin || (N % 32) || (ep->rope_cols % 32) || ep->rope_rows_per_batch <= 0 ||
                       ep->rope_table_rows <= 0 || (ep->rope_table_rows % 32) || (reinterpret_cast<uintptr_t>(ep->rope_cos) & 15) ||
                       (reinterpret_cast<uintptr_t>(ep->rope_sin) & 15)))
    return USVM_ERR_ARG;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (ep->ln_w) {  // fused LayerNorm: one 128 x 256 tile per CTA holds whole rows
    if (N != 256 || !ep->ln_b || !ep->out_bf16 || ep->rope_cos || (reinterpret_cast<uintptr_t>(ep->ln_w) & 15) ||
        (reinterpret_cast<uintptr_t>(ep->ln_b) & 15))
      return USVM_ERR_ARG;
    return launch<256, false>(A, lda, W, ldw, ep, M, N, K, s);
  }
  int bn = block_n;
  if (bn < 0) return launch_persistent(A, lda, W, ldw, ep, M, N, K, bn == -1 ? 0 : -bn, s);
  if (bn == 0 && (long long)cdiv(M, BM) * cdiv(N, 128) >= 148)  // throughput-bound: a full wave of 128 x 128 tiles or more
    return launch_persistent(A, lda, W, ldw, ep, M, N, K, 0, s);
  if (bn <= 0) {
    // latency-bound shapes: the narrowest tile (most CTAs) that still runs as a single wave -- a second wave costs a
    // whole extra tile time.  Slots per SM follow from the shared memory of the k-deep ring (<= 2 by launch bounds).
    const int mt = cdiv(M, BM), num_kb = cdiv(K, BK);
    const int ring = num_kb < STAGES ? num_kb : STAGES;
    auto one_wave = [&](int b, int total_smem) {
      const int occ = total_smem * 2 <= 227 * 1024 ? 2 : 1;
      return (long long)mt * cdiv(N, b) <= 148LL * occ;
    };
    bn = 128;
    if (one_wave(32, SmemLayout<32>::total(ring))) bn = 32;
    else if (one_wave(64, SmemLayout<64>::total(ring))) bn = 64;
    if (num_kb >= 32 && N >= 64 && bn == 32) bn = 64;  // long K: halve the re-reads of A (split-K restores the CTA count)
    if (N <= 32) bn = 32;
    else if (N <= 64 && bn > 64) bn = 64;
    else if (N <= 128 && bn > 128) bn = 128;
  }
  // long reductions on few tiles (the FFN's second linear: K = 2048 on 64 tiles): split K over a cluster of 2 or 4 CTAs
  int ksplit = 1;
  if (block_n <= 0 || block_n == bn) {
    const long long ctas = (long long)cdiv(M, BM) * cdiv(N, bn);
    const int num_kb = cdiv(K, BK);
    // (N <= 256 keeps every image-encoder GEMM out: its results must not depend on how many frames are batched)
    if (!ep->ln_w && bn <= 64 && N <= 256 && ctas <= 74 && num_kb >= 16)
      ksplit = (num_kb >= 32 && ctas * 4 <= 2 * 148) ? 4 : 2;
    if (const char* e = getenv("USVM2_GEMM_KSPLIT")) ksplit = (e[0] == '0' || e[0] == '1') ? 1 : ksplit;
  }
  switch (bn) {
    case 32: return launch<32, false>(A, lda, W, ldw, ep, M, N, K, s, ksplit);
    case 64: return launch<64, false>(A, lda, W, ldw, ep, M, N, K, s, ksplit);
    case 128: return launch<128, false>(A, lda, W, ldw, ep, M, N, K, s);
    case 256: return launch<256, false>(A, lda, W, ldw, ep, M, N, K, s);
    default: return USVM_ERR_ARG;
  }
}

extern "C" int usvm_gemm_tf32_tc5(const float* A, int lda, const float* W, int ldw, const usvm_gemm_epilogue* ep,
                                  int M, int N, int K, int block_n, void* stream) {
  if (!A || !W || !ep || M <= 0 || N <= 0 || K <= 0) return USVM_ERR_ARG;
  if ((lda % 4) || (ldw % 4) || (reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(W) & 15))
    return USVM_ERR_ARG;
  if (ep->out_f32 && ((ep->ldo_f32 % 4) || (reinterpret_cast<uintptr_t>(ep->out_f32) & 15))) return USVM_ERR_ARG;
  if (ep->out_bf16 && ((ep->ldo_bf16 % 8) || (reinterpret_cast<uintptr_t>(ep->out_bf16) & 15))) return USVM_ERR_ARG;
  if (ep->residual && ((ep->ldr % 4) || (reinterpret_cast<uintptr_t>(ep->residual) & 15))) return USVM_ERR_ARG;
  if (ep->rope_cos) return USVM_ERR_ARG;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  int bn = block_n;
  // throughput-bound (the batched decoder): two waves of 128 x 128 tiles or more -> persistent kernel
  if (bn == 0 && (long long)cdiv(M, BM) * cdiv(N, 128) >= 2 * 148 && (N % 32) == 0)
    return launch_persistent(A, lda, W, ldw, ep, M, N, K, 0, s, true);
  if (bn <= 0) {
    const int mt = cdiv(M, BM);
    bn = 128;
    while (bn > 32 && (long long)mt * cdiv(N, bn) < 148) bn >>= 1;
    if (N <= 32) bn = 32;
    else if (N <= 64 && bn > 64) bn = 64;
  }
  switch (bn) {
    case 32: return launch<32, true>(A, lda, W, ldw, ep, M, N, K, s);
    case 64: return launch<64, true>(A, lda, W, ldw, ep, M, N, K, s);
    case 128: return launch<128, true>(A, lda, W, ldw, ep, M, N, K, s);
    default: return USVM_ERR_ARG;
  }
}
