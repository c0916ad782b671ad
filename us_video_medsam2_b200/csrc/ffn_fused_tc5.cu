// Feed-forward block of a memory-attention layer as ONE kernel (MemoryAttentionLayer.forward, memory_attention.py:92-98:
// tgt = tgt + linear2(relu(linear1(norm3(tgt)))), d_model 256, hidden 2048):
//     out[M,256] = x + relu(h W1^T + b1) W2^T + b2            h = LayerNorm(x) in bf16 (written by the norm kernel)
// On the one-object frame the two GEMMs were a 7 us + 11 us pair of launches (128 CTAs each, the second one split-K over a
// cluster) with the 1024 x 2048 hidden activations making a round trip through L2 in between.  Here a cluster of 8 CTAs
// owns a 128-row tile; CTA c of the cluster owns hidden units [256c, 256c + 256):
//   GEMM 1   D1[128 x 256] = h_tile . W1_c^T          (tcgen05.mma M128 N256, accumulator in TMEM columns [0,256))
//   epilogue D1 + b1 -> ReLU -> bf16 -> shared memory, written in the swizzled K-major layout a UMMA A operand expects
//            (it overwrites the h tile, which is dead by then): the hidden activations never leave the SM
//   GEMM 2   D2[128 x 256] = H_c . W2[:, 256c : 256c + 256]^T   (TMEM columns [256,512)) -- a partial sum over this CTA's hidden units
//   reduce-scatter over the cluster through distributed shared memory: CTA q receives the seven other CTAs' columns
//            [32q, 32q + 32) of D2 (st.async into its free weight ring, mbarrier transaction counts instead of cluster
//            barriers), adds them in rank order (deterministic), adds b2 and the residual x, writes out.
// Measured at M = 1024 (graph replay): 14.8 us against 17.9 us for the two launches; phase timeline of one CTA in cycles:
// GEMM 1 5.0 k (16 MMAs + first loads), hidden epilogue 2.8 k, GEMM 2 2.5 k, reduce-scatter 8.4 k (DSMEM moves ~21 B/clk
// per SM), final sum + store 1.6 k.
// W1_c / W2_c k-blocks stream through a 5-stage TMA ring; they are constants of the model, so the first five are requested
// BEFORE griddepcontrol.wait and arrive under the previous kernel's tail.  64 CTAs at M = 1024: one wave even on the SMs
// the encoder partition leaves to the tracked frame.
#include "common.cuh"
#include "usvm2_b200.h"

namespace {

constexpr int BM = 128;
constexpr int DM = 256;        // d_model
constexpr int HC = 256;        // hidden units per CTA
constexpr int CL = 8;          // cluster size = hidden / HC
constexpr int BK = 64;
constexpr int NKB = DM / BK;   // 4 k-blocks per GEMM
constexpr int A_KB = BM * 128;   // 16 KB: one k-block of the h / H tile
constexpr int B_KB = 256 * 128;  // 32 KB: one k-block of a weight chunk (256 rows)
constexpr int RING = 5;
constexpr int THREADS = 320;
constexpr int SLOT = BM * 32 * 4;  // 16 KB: one peer's 128 x 32 fp32 block
constexpr int SMEM = NKB * A_KB + RING * B_KB + 1024 /* alignment slack */ + 256 /* barriers */;
static_assert(CL * SLOT <= RING * B_KB, "the receive slots reuse the weight ring");

__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa_shared(uint32_t local_addr, uint32_t cta_rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(cta_rank));
  return r;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// remote (peer CTA) mbarrier arrive with cluster-scope release, and a 16-byte asynchronous store into a peer's shared
// memory that completes 16 bytes of that peer's mbarrier transaction count when it lands
__device__ __forceinline__ void mbar_arrive_remote(uint32_t remote_bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote_bar) : "memory");
}
__device__ __forceinline__ void st_async_v4(uint32_t remote_addr, uint32_t remote_bar, uint32_t a, uint32_t b, uint32_t c,
                                            uint32_t d) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%2, %3, %4, %5}, [%1];"
               ::"r"(remote_addr), "r"(remote_bar), "r"(a), "r"(b), "r"(c), "r"(d)
               : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  const long long t0 = clock64();
  while (true) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    if (ok) return;
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}
__device__ __forceinline__ void ld_tmem_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// phase timeline of CTA (0, 0) for tools/ffn_timeline.py: build with EXTRA=-DUSVM2_FFN_PROFILE
__device__ unsigned long long g_ffn_prof[16];
#ifdef USVM2_FFN_PROFILE
#define FFN_STAMP(i) do { if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 64) g_ffn_prof[i] = clock64(); } while (0)
#else
#define FFN_STAMP(i) do { } while (0)
#endif
__global__ void __launch_bounds__(THREADS, 1)
ffn_fused_tc5_kernel(const __grid_constant__ CUtensorMap tmH, const __grid_constant__ CUtensorMap tmW1,
                     const __grid_constant__ CUtensorMap tmW2, const float* __restrict__ x,
                     const float* __restrict__ b1, const float* __restrict__ b2, float* __restrict__ out, const int M) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                  // h tile, then H_c (4 k-blocks of 128 rows x 128 B, 128B swizzle)
  uint8_t* ring = smem + NKB * A_KB;   // weight k-blocks; after GEMM 2 the peers' partial blocks
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + RING * B_KB);
  uint64_t* a_full = bars;             // 1
  uint64_t* b_full = bars + 1;         // RING
  uint64_t* b_empty = bars + 1 + RING; // RING
  uint64_t* d1_full = bars + 1 + 2 * RING;
  uint64_t* h_full = d1_full + 1;      // 256 arrivals
  uint64_t* d2_full = d1_full + 2;
  uint64_t* ready_bar = d1_full + 3;   // CL - 1 remote arrivals: "my weight ring is free, send your partials"
  uint64_t* data_bar = d1_full + 4;    // 1 arrival + (CL - 1) * SLOT bytes of st.async traffic
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d1_full + 5);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  FFN_STAMP(0);
  const int row0 = blockIdx.x * BM;
  const uint32_t rank = cluster_rank();  // = blockIdx.y: which 256 hidden units

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmH);
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmW2);
    mbar_init(a_full, 1);
    for (int s = 0; s < RING; ++s) {
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], 1);
    }
    mbar_init(d1_full, 1);
    mbar_init(h_full, 256);
    mbar_init(d2_full, 1);
    mbar_init(ready_bar, CL - 1);
    mbar_init(data_bar, 1);
    mbar_fence_init();
    mbar_arrive_expect_tx(data_bar, (CL - 1) * SLOT);
  }
  if (warp == 1) tc5_alloc(tmem_slot, 512);
  tc5_fence_before();
  __syncthreads();
  tc5_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tmem_D1 = tmem, tmem_D2 = tmem + 256;
  cluster_arrive();  // "my barriers are initialised": awaited (cluster_wait) only before the first remote operation

  // weight k-block i (0..3: W1_c, 4..7: W2_c) into ring stage i % RING
  auto load_weight = [&](int i) {
    const int s = i % RING;
    mbar_arrive_expect_tx(&b_full[s], B_KB);
    if (i < NKB) tma_load_2d(ring + s * B_KB, &tmW1, &b_full[s], i * BK, (int)rank * HC);
    else tma_load_2d(ring + s * B_KB, &tmW2, &b_full[s], (int)rank * HC + (i - NKB) * BK, 0);
  };

  float4 xr[4], bo[4];  // epilogue warps: residual and bias segment of the thread's 16 output columns
  if (warp == 0) {
    if (elect_one()) {  // constants of the model: requested before the predecessor has finished
      for (int i = 0; i < RING; ++i) load_weight(i);
    }
    __syncwarp();
  }
  FFN_STAMP(1);
  pdl_wait();
  pdl_trigger();
  FFN_STAMP(2);

  if (warp == 0) {
    if (elect_one()) {
      mbar_arrive_expect_tx(a_full, NKB * A_KB);
      for (int kb = 0; kb < NKB; ++kb) tma_load_2d(sA + kb * A_KB, &tmH, a_full, kb * BK, row0);
    }
    __syncwarp();
    for (int i = RING; i < 2 * NKB; ++i) {
      mbar_wait(&b_empty[i % RING], 0);  // first (and only) reuse of the stage
      if (elect_one()) load_weight(i);
      __syncwarp();
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = umma_idesc_bf16(BM, 256);
    mbar_wait(a_full, 0);
    for (int i = 0; i < 2 * NKB; ++i) {
      const int s = i % RING, kb = i % NKB;
      if (i == NKB) {  // GEMM 2 reads H_c, which the epilogue warps are writing over the h tile
        mbar_wait(h_full, 0);
      }
      mbar_wait(&b_full[s], (i / RING) & 1);
      tc5_fence_after();
      const uint64_t a_desc = umma_desc_k_sw128(smem_u32(sA + kb * A_KB));
      const uint64_t b_desc = umma_desc_k_sw128(smem_u32(ring + s * B_KB));
      const uint32_t d = i < NKB ? tmem_D1 : tmem_D2;
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < BK / 16; ++k)
          tc5_mma_f16(d, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc, (kb > 0 || k > 0) ? 1u : 0u);
        tc5_commit(&b_empty[s]);
        if (i == NKB - 1) tc5_commit(d1_full);
        if (i == 2 * NKB - 1) tc5_commit(d2_full);
      }
      __syncwarp();
    }
  } else {
    const int lane_grp = warp & 3;
    const int half = (warp - 2) >> 2;
    const int r = lane_grp * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(lane_grp * 32) << 16;
    // the residual and bias segment of this thread's 16 output columns: requested now, used at the very end
    const int ocol = (int)rank * 32 + half * 16;
    {
      const float4* xp = reinterpret_cast<const float4*>(x + (long long)(row0 + r) * DM + ocol);
      const float4* bp = reinterpret_cast<const float4*>(b2 + ocol);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        xr[j] = __ldg(xp + j);
        bo[j] = __ldg(bp + j);
      }
    }
    // ---- epilogue of GEMM 1: blocks of 32 hidden units; this warp owns blocks 4*half .. 4*half + 3 ----
    mbar_wait(d1_full, 0);
    tc5_fence_after();
    FFN_STAMP(3);
#pragma unroll 1
    for (int j = half * 4; j < half * 4 + 4; ++j) {
      uint32_t acc[32];
      tc5_ld_32x32(tmem_D1 + lane_addr + (uint32_t)(j * 32), acc);
      float4 bb[8];
      const float4* bp = reinterpret_cast<const float4*>(b1 + (int)rank * HC + j * 32);
#pragma unroll
      for (int q = 0; q < 8; ++q) bb[q] = __ldg(bp + q);
      tc5_wait_ld();
      uint8_t* rowp = sA + (j >> 1) * A_KB + r * 128;
      const int cb = (j & 1) * 4;
#pragma unroll
      for (int q = 0; q < 4; ++q) {  // chunk = 8 hidden units = 16 bytes
        const float4 b0 = bb[2 * q], b1v = bb[2 * q + 1];
        uint4 pk;
        pk.x = pack_bf16x2(fmaxf(__uint_as_float(acc[8 * q]) + b0.x, 0.f), fmaxf(__uint_as_float(acc[8 * q + 1]) + b0.y, 0.f));
        pk.y = pack_bf16x2(fmaxf(__uint_as_float(acc[8 * q + 2]) + b0.z, 0.f), fmaxf(__uint_as_float(acc[8 * q + 3]) + b0.w, 0.f));
        pk.z = pack_bf16x2(fmaxf(__uint_as_float(acc[8 * q + 4]) + b1v.x, 0.f), fmaxf(__uint_as_float(acc[8 * q + 5]) + b1v.y, 0.f));
        pk.w = pack_bf16x2(fmaxf(__uint_as_float(acc[8 * q + 6]) + b1v.z, 0.f), fmaxf(__uint_as_float(acc[8 * q + 7]) + b1v.w, 0.f));
        *reinterpret_cast<uint4*>(rowp + (((cb + q) ^ (r & 7)) << 4)) = pk;
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_arrive(h_full);
    FFN_STAMP(4);
    mbar_wait(d2_full, 0);
    tc5_fence_after();
    FFN_STAMP(5);
  }
  // ---- reduce-scatter of the 8 partial D2 through distributed shared memory, synchronised by mbarriers only (a
  // barrier.cluster with release / acquire costs ~1.2 k cycles each way and serialises the whole cluster twice):
  //   ready_bar of CTA q: the other 7 CTAs arrive once their GEMM 2 is complete, i.e. their ring can take q's blocks
  //   data_bar  of CTA q: completes when the 7 x 16 KB of st.async traffic addressed to q have landed
  cluster_wait();  // every CTA of the cluster has initialised its barriers (arrived on long ago)
  if (warp == 1) {
    mbar_wait(d2_full, 0);  // this CTA's ring is no longer read by the tensor pipe
    tc5_fence_after();
    if (lane < CL && (uint32_t)lane != rank) mbar_arrive_remote(mapa_shared(smem_u32(ready_bar), (uint32_t)lane));
    __syncwarp();
  }
  FFN_STAMP(6);
  if (warp >= 2) {
    const int lane_grp = warp & 3;
    const int half = (warp - 2) >> 2;
    const int r = lane_grp * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(lane_grp * 32) << 16;
    // receive slot (index = sender rank) layout [8 chunks of 4 columns][128 rows] x 16 B: the 32 lanes of a store write
    // 512 contiguous bytes, the 32 lanes of the final read hit 32 different banks (row-major slots -- 128-byte lane
    // stride -- made both the remote stores and the final loads 32-way serialised: 13 k + 8 k cycles)
    const uint32_t recv_local = smem_u32(ring) + rank * SLOT + (uint32_t)r * 16u;
    const uint32_t data_local = smem_u32(data_bar);
    mbar_wait_cluster(ready_bar, 0);
#pragma unroll 1
    for (int pp = 0; pp < CL / 2; ++pp) {
      const uint32_t peer = (uint32_t)(half * (CL / 2) + pp);
      if (peer == rank) continue;  // warp-uniform
      uint32_t acc[32];
      tc5_ld_32x32(tmem_D2 + lane_addr + peer * 32u, acc);
      const uint32_t dst = mapa_shared(recv_local, peer), dbar = mapa_shared(data_local, peer);
      tc5_wait_ld();
#pragma unroll
      for (int q = 0; q < 8; ++q)
        st_async_v4(dst + q * 2048, dbar, acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
    }
    FFN_STAMP(7);
    mbar_wait_cluster(data_bar, 0);
  }
  FFN_STAMP(8);
  if (warp >= 2) {
    const int lane_grp = warp & 3;
    const int half = (warp - 2) >> 2;
    const int r = lane_grp * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(lane_grp * 32) << 16;
    const int ocol = (int)rank * 32 + half * 16;
    uint32_t own[16];
    ld_tmem_x16(tmem_D2 + lane_addr + (uint32_t)ocol, own);
    tc5_wait_ld();
    float v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = 0.f;
    // partial sums in rank order (the own one from TMEM, the others from the receive slots)
#pragma unroll 1
    for (uint32_t p = 0; p < (uint32_t)CL; ++p) {
      if (p == rank) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] += __uint_as_float(own[i]);
      } else {
        const float4* sp = reinterpret_cast<const float4*>(ring + p * SLOT + (half * 4 * 128 + r) * 16);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 t = sp[j * 128];
          v[4 * j] += t.x; v[4 * j + 1] += t.y; v[4 * j + 2] += t.z; v[4 * j + 3] += t.w;
        }
      }
    }
    if (row0 + r < M) {
      float4* op = reinterpret_cast<float4*>(out + (long long)(row0 + r) * DM + ocol);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        op[j] = make_float4(v[4 * j] + bo[j].x + xr[j].x, v[4 * j + 1] + bo[j].y + xr[j].y, v[4 * j + 2] + bo[j].z + xr[j].z,
                            v[4 * j + 3] + bo[j].w + xr[j].w);
    }
  }
  FFN_STAMP(9);
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
int make_map(CUtensorMap* map, const void* base, long long rows, long long cols, long long pitch, int box_rows) {
  PFN_encodeTiled enc = encode_fn();
  if (!enc) return USVM_ERR_DRIVER;
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)pitch * 2};
  cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? USVM_OK : USVM_ERR_DRIVER;
}

}  // namespace

extern "C" int usvm_debug_ffn_profile(unsigned long long* host_out_16) {
  return cudaMemcpyFromSymbol(host_out_16, g_ffn_prof, sizeof(g_ffn_prof)) == cudaSuccess ? USVM_OK : USVM_ERR_CUDA;
}

extern "C" int usvm_ffn_fused_tc5(const void* h_bf16, const float* x, const void* w1, const float* b1, const void* w2,
                                  const float* b2, float* out, int M, int d_model, int hidden, void* stream) {
  if (!h_bf16 || !x || !w1 || !b1 || !w2 || !b2 || !out || M <= 0) return USVM_ERR_ARG;
  if (d_model != DM || hidden != HC * CL || (M % BM)) return USVM_ERR_ARG;
  for (const void* p : {h_bf16, (const void*)x, w1, (const void*)b1, w2, (const void*)b2, (const void*)out})
    if (reinterpret_cast<uintptr_t>(p) & 15) return USVM_ERR_ARG;
  CUtensorMap tmH, tmW1, tmW2;
  int rc = make_map(&tmH, h_bf16, M, DM, DM, BM);
  if (rc) return rc;
  rc = make_map(&tmW1, w1, hidden, DM, DM, HC);
  if (rc) return rc;
  rc = make_map(&tmW2, w2, DM, hidden, hidden, DM);
  if (rc) return rc;
  if (cudaFuncSetAttribute(ffn_fused_tc5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM) != cudaSuccess)
    return USVM_ERR_CUDA;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(M / BM, CL, 1);
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = SMEM;
  cfg.stream = reinterpret_cast<cudaStream_t>(stream);
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1;
  attr[0].val.clusterDim.y = CL;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = usvm_pdl_enabled() ? 2 : 1;
  if (cudaLaunchKernelEx(&cfg, ffn_fused_tc5_kernel, tmH, tmW1, tmW2, x, b1, b2, out, M) != cudaSuccess)
    return USVM_ERR_CUDA;
  return usvm_check_launch();
}
