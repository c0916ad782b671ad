// Connected components (8-connectivity) + fused small-hole filling.
//
// Replaces sam2._C.get_connected_componnets (reference: sam2/csrc/connected_components.cu:213-282,
// 6 launches per image, global atomics, three zero-filled int32 scratch maps) and the elementwise
// wrapper fill_holes_in_mask_scores (sam2/utils/misc.py:312-338).
//
// Result contract (bit-exact with the reference): for foreground pixel p
//   labels[p] = 1 + min{ (r & ~1) * W + (c & ~1) : (r, c) foreground pixel of p's component }
//   counts[p] = area of p's component;   both 0 on background.
//
// Hot-path shape is [B,1,128,128].  One CTA owns one image: the mask, the union-find parents of the
// 2x2 blocks and the per-root areas all live in shared memory (48 KB at 128^2), so the whole op is one
// launch for all images and touches HBM exactly once per input byte and once per output word.
// Images whose block grid does not fit in shared memory take a batched global-memory path (5 launches
// for all N).  HBM-bound: 1 B in + 8 B out per pixel (labelling) or 4 B in + 4 B out (hole filling).
#include "common.cuh"
#include "usvm2_b200.h"

namespace {

__device__ __forceinline__ int uf_find(const volatile int* par, int a) {
  int p = par[a];
  while (p != a) {
    a = p;
    p = par[a];
  }
  return a;
}

// lock-free union, links always point to the smaller index so the root is the component minimum
__device__ __forceinline__ void uf_union(int* par, int a, int b) {
  while (true) {
    a = uf_find(par, a);
    b = uf_find(par, b);
    if (a == b) return;
    if (a > b) {
      const int t = a;
      a = b;
      b = t;
    }
    const int old = atomicMin(&par[b], a);  // b was a root when read; if it still is, it now points to a
    if (old == b) return;
    b = old;  // somebody re-linked b first: retry from its new parent
  }
}

// Foreground test of the two input flavours.
struct FgU8 {
  const uint8_t* img;
  __device__ __forceinline__ bool operator()(long long i) const { return img[i] != 0; }
};
struct FgScore {  // hole filling: "foreground" of the CC pass is the mask's background, score <= 0
  const float* s;
  __device__ __forceinline__ bool operator()(long long i) const { return s[i] <= 0.0f; }
};

// Merge rule of one 2x2 block with its TL / T / TR / L neighbour blocks (connected_components.cu:72-117):
// `px(r,c)` must return false outside the image.
template <typename PX, typename UN>
__device__ __forceinline__ void merge_block(int r, int c, PX px, UN unite) {
  const bool tl = px(r, c), tr = px(r, c + 1), bl = px(r + 1, c);
  if (!(tl || tr || bl)) return;
  if (tl && px(r - 1, c - 1)) unite(-1, -1);
  if ((tl || tr) && (px(r - 1, c) || px(r - 1, c + 1))) unite(-1, 0);
  if (tr && px(r - 1, c + 2)) unite(-1, 1);
  if ((tl || bl) && (px(r, c - 1) || px(r + 1, c - 1))) unite(0, -1);
}

// ---------------------------------------------------------------------------------------------
// shared-memory path: one CTA per image
// ---------------------------------------------------------------------------------------------
template <bool FILL>
__global__ void __launch_bounds__(1024)
cc_smem_kernel(const uint8_t* __restrict__ img, const float* scores_in, float* scores_out, int32_t* __restrict__ labels,
               int32_t* __restrict__ counts, int H, int W, int max_area, float fill_value) {
  PDL_ENTRY();
  extern __shared__ int s_mem[];
  const int HB = H >> 1, WB = W >> 1, NB = HB * WB, HW = H * W;
  int* s_par = s_mem;
  int* s_area = s_mem + NB;
  uint8_t* s_fg = reinterpret_cast<uint8_t*>(s_mem + 2 * NB);
  const long long base = (long long)blockIdx.x * HW;
  const int tid = threadIdx.x, nt = blockDim.x;

  for (int i = tid; i < HW; i += nt) s_fg[i] = FILL ? (scores_in[base + i] <= 0.0f) : (img[base + i] != 0);
  for (int b = tid; b < NB; b += nt) {
    s_par[b] = b;
    s_area[b] = 0;
  }
  __syncthreads();
  auto px = [&](int r, int c) { return r >= 0 && r < H && c >= 0 && c < W && s_fg[r * W + c]; };
  for (int b = tid; b < NB; b += nt) {
    const int by = b / WB, bx = b - by * WB;
    merge_block(2 * by, 2 * bx, px, [&](int dy, int dx) { uf_union(s_par, b, (by + dy) * WB + (bx + dx)); });
  }
  __syncthreads();
  for (int b = tid; b < NB; b += nt) {
    const int root = uf_find(s_par, b);
    const int by = b / WB, bx = b - by * WB;
    const int p = 2 * by * W + 2 * bx;
    const int k = s_fg[p] + s_fg[p + 1] + s_fg[p + W] + s_fg[p + W + 1];
    {  // warp-aggregated: lanes that share a root (a giant component!) issue one atomic instead of 32
      const unsigned peers = __match_any_sync(__activemask(), root);
      const int total = __reduce_add_sync(peers, k);
      if (total && (threadIdx.x & 31) == (__ffs(peers) - 1)) atomicAdd(&s_area[root], total);
    }
    // path compression is deferred to a private write: other threads may still be walking through b
    // towards the root, and pointing b at the root keeps every chain valid
    s_par[b] = root;
  }
  __syncthreads();
  for (int i = tid; i < HW; i += nt) {
    const int r = i / W, c = i - r * W;
    const bool fg = s_fg[i];
    const int root = s_par[(r >> 1) * WB + (c >> 1)];
    if (FILL) {
      const float v = scores_in[base + i];
      scores_out[base + i] = (fg && s_area[root] <= max_area) ? fill_value : v;
    } else {
      const int ry = root / WB, rx = root - ry * WB;
      labels[base + i] = fg ? (2 * ry * W + 2 * rx + 1) : 0;
      counts[base + i] = fg ? s_area[root] : 0;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// global-memory path for large images (parents live in `labels` at the block anchors, areas in `counts`)
// ---------------------------------------------------------------------------------------------
template <typename FG>
__global__ void cc_g_init(FG fg, int32_t* labels, int32_t* counts, int H, int W) {
  PDL_ENTRY();
  const long long base = (long long)blockIdx.z * H * W;
  const int bx = blockIdx.x * blockDim.x + threadIdx.x, by = blockIdx.y * blockDim.y + threadIdx.y;
  if (2 * bx >= W || 2 * by >= H) return;
  const int a = 2 * by * W + 2 * bx;
  labels[base + a] = a;
  counts[base + a] = 0;
}
template <typename FG>
__global__ void cc_g_merge(FG fg, int32_t* labels, int H, int W) {
  PDL_ENTRY();
  const long long base = (long long)blockIdx.z * H * W;
  const int bx = blockIdx.x * blockDim.x + threadIdx.x, by = blockIdx.y * blockDim.y + threadIdx.y;
  if (2 * bx >= W || 2 * by >= H) return;
  const int r = 2 * by, c = 2 * bx, a = r * W + c;
  int* par = labels + base;
  auto px = [&](int rr, int cc) { return rr >= 0 && rr < H && cc >= 0 && cc < W && fg(base + (long long)rr * W + cc); };
  merge_block(r, c, px, [&](int dy, int dx) { uf_union(par, a, a + 2 * dy * W + 2 * dx); });
}
template <typename FG>
__global__ void cc_g_compress(FG fg, int32_t* labels, int32_t* counts, int H, int W) {
  PDL_ENTRY();
  const long long base = (long long)blockIdx.z * H * W;
  const int bx = blockIdx.x * blockDim.x + threadIdx.x, by = blockIdx.y * blockDim.y + threadIdx.y;
  if (2 * bx >= W || 2 * by >= H) return;
  const int a = 2 * by * W + 2 * bx;
  const int root = uf_find(labels + base, a);
  labels[base + a] = root;
  const int k = (int)fg(base + a) + (int)fg(base + a + 1) + (int)fg(base + a + W) + (int)fg(base + a + W + 1);
  if (k) atomicAdd(&counts[base + root], k);
}
// writes everything except the anchor slots of root blocks (they still hold parent / area for readers)
template <typename FG, bool FILL>
__global__ void cc_g_final(FG fg, int32_t* labels, int32_t* counts, const float* scores_in, float* scores_out, int H,
                           int W, int max_area, float fill_value) {
  PDL_ENTRY();
  const long long base = (long long)blockIdx.z * H * W;
  const int bx = blockIdx.x * blockDim.x + threadIdx.x, by = blockIdx.y * blockDim.y + threadIdx.y;
  if (2 * bx >= W || 2 * by >= H) return;
  const int a = 2 * by * W + 2 * bx;
  const int root = labels[base + a];
  const int area = counts[base + root];
#pragma unroll
  for (int d = 3; d >= 0; --d) {
    const int p = a + (d >> 1) * W + (d & 1);
    if (d == 0 && root == a) break;  // root anchor: finalised by cc_g_roots
    const bool f = fg(base + p);
    if (FILL) {
      scores_out[base + p] = (f && area <= max_area) ? fill_value : scores_in[base + p];
    } else {
      labels[base + p] = f ? root + 1 : 0;
      counts[base + p] = f ? area : 0;
    }
  }
}
template <typename FG, bool FILL>
__global__ void cc_g_roots(FG fg, int32_t* labels, int32_t* counts, const float* scores_in, float* scores_out, int H,
                           int W, int max_area, float fill_value) {
  PDL_ENTRY();
  const long long base = (long long)blockIdx.z * H * W;
  const int bx = blockIdx.x * blockDim.x + threadIdx.x, by = blockIdx.y * blockDim.y + threadIdx.y;
  if (2 * bx >= W || 2 * by >= H) return;
  const int a = 2 * by * W + 2 * bx;
  if (labels[base + a] != a) return;  // non-root anchors were rewritten to root + 1 (!= a) or 0 (a == 0 is a root)
  const bool f = fg(base + a);
  const int area = counts[base + a];
  if (FILL) {
    scores_out[base + a] = (f && area <= max_area) ? fill_value : scores_in[base + a];
  } else {
    labels[base + a] = f ? a + 1 : 0;
    counts[base + a] = f ? area : 0;
  }
}

// ---------------------------------------------------------------------------------------------
// Hole filling without global labelling.  Only components of area <= max_area matter, and such a component has
// graph diameter < max_area, so `max_area` SYNCHRONOUS rounds of min-label propagation over the 3x3 neighbourhood
// make every small component converge to its minimum pixel index.  A label group that is (a) closed -- no member
// touches a background pixel carrying a different label -- and (b) of size <= max_area is exactly one small
// component: a closed group is a whole component, and an unconverged larger component always contains an adjacent
// pair of different labels, which marks both groups open.  Synchronous (ping-pong) rounds also bound every group to
// the (2 max_area + 1)^2 window around its label, so the per-label atomic counters never see more than a few hundred
// increments (a giant component would otherwise serialise thousands of atomics on one address).
// ---------------------------------------------------------------------------------------------
constexpr int kNoLabel = 0x7fffffff;

// TILED over the image: because a component of area <= max_area containing a given pixel lies within Chebyshev
// distance max_area - 1 of it, a CTA that owns a 32 x 32 core only needs the window core + R (R = max_area rounded up to
// a multiple of 4) to decide its own pixels.  Window edges that are NOT image edges carry a one-pixel border of
// kForeign: a group touching it is marked open (a fragment of a component that continues outside the window always has a
// member on the window's outermost ring, a genuine small hole of a core pixel never reaches that ring), image edges carry
// kNoLabel as before.  16 CTAs of 256 threads per 128 x 128 image instead of one CTA of 1024 (47 us on a single SM).
// Each thread owns 4 x 4 pixel tiles (labels in registers, 6 x 6 halo read from shared memory once per round).
constexpr int kForeign = 0x7ffffffe;
constexpr int kFillCore = 32;

__global__ void __launch_bounds__(256)
fill_holes_local_kernel(const float* scores_in, float* scores_out, int H, int W, int max_area, float fill_value, int R) {
  PDL_ENTRY();
  extern __shared__ int s_mem[];
  // window of this CTA (clamped to the image); all coordinates multiples of 4
  const int cx0 = blockIdx.x * kFillCore, cy0 = blockIdx.y * kFillCore;
  const int wx0 = max(0, cx0 - R), wy0 = max(0, cy0 - R);
  const int wx1 = min(W, cx0 + kFillCore + R), wy1 = min(H, cy0 + kFillCore + R);
  const int ww = wx1 - wx0, wh = wy1 - wy0;
  const int Wp = ww + 2, plane = (wh + 2) * Wp, npix = ww * wh;
  int* s_a = s_mem;                // labels, ping   [(wh+2) x (ww+2)]
  int* s_b = s_mem + plane;        // labels, pong
  int* s_cnt = s_mem + 2 * plane;  // per-label closed-member count (open members add 2^20)
  const long long base = (long long)blockIdx.z * H * W;
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int i = tid; i < plane; i += nt) {
    const int r = i / Wp - 1, c = i - (r + 1) * Wp - 1;  // window coordinates of this plane entry (-1 / wh / ww: border)
    int v = kNoLabel;
    if (r < 0 || r >= wh || c < 0 || c >= ww) {
      const int gy = wy0 + r, gx = wx0 + c;
      v = (gy >= 0 && gy < H && gx >= 0 && gx < W) ? kForeign : kNoLabel;
    } else if (scores_in[base + (long long)(wy0 + r) * W + wx0 + c] <= 0.0f) {
      v = r * ww + c;
    }
    s_a[i] = v;
    s_b[i] = v;  // borders stay put in both planes; interior entries are rewritten every round
  }
  for (int i = tid; i < npix; i += nt) s_cnt[i] = 0;
  __syncthreads();
  const int tiles_x = ww >> 2, tiles = (wh >> 2) * tiles_x;
  int* cur = s_a;
  int* nxt = s_b;
  for (int round = 0; round < max_area; ++round) {
    for (int t = tid; t < tiles; t += nt) {
      const int ty = t / tiles_x, tx = t - ty * tiles_x;
      const int* src = cur + (4 * ty) * Wp + 4 * tx;  // top-left of the 6 x 6 halo window
      int v[6][6];
#pragma unroll
      for (int y = 0; y < 6; ++y)
#pragma unroll
        for (int x = 0; x < 6; ++x) v[y][x] = src[y * Wp + x];
      int* dst = nxt + (4 * ty + 1) * Wp + 4 * tx + 1;
#pragma unroll
      for (int y = 1; y <= 4; ++y)
#pragma unroll
        for (int x = 1; x <= 4; ++x) {
          int m = min(min(min(v[y - 1][x - 1], v[y - 1][x]), min(v[y - 1][x + 1], v[y][x - 1])),
                      min(min(v[y][x + 1], v[y + 1][x - 1]), min(v[y + 1][x], v[y + 1][x + 1])));
          m = min(m, v[y][x]);
          // (kForeign is larger than every pixel label: it never wins the minimum of a mask pixel)
          dst[(y - 1) * Wp + (x - 1)] = v[y][x] == kNoLabel ? kNoLabel : m;
        }
    }
    __syncthreads();
    int* tswap = cur;
    cur = nxt;
    nxt = tswap;
  }
  for (int t = tid; t < tiles; t += nt) {
    const int ty = t / tiles_x, tx = t - ty * tiles_x;
    const int* src = cur + (4 * ty) * Wp + 4 * tx;
    int v[6][6];
#pragma unroll
    for (int y = 0; y < 6; ++y)
#pragma unroll
      for (int x = 0; x < 6; ++x) v[y][x] = src[y * Wp + x];
#pragma unroll
    for (int y = 1; y <= 4; ++y)
#pragma unroll
      for (int x = 1; x <= 4; ++x) {
        const int me = v[y][x];
        if (me == kNoLabel) continue;
        bool open = false;
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
          for (int dx = -1; dx <= 1; ++dx) {
            const int q = v[y + dy][x + dx];
            open |= (q != kNoLabel && q != me);
          }
        atomicAdd(&s_cnt[me], open ? (1 << 20) : 1);
      }
  }
  __syncthreads();
  // write back the core only
  const int cw = min(kFillCore, W - cx0), ch = min(kFillCore, H - cy0);
  for (int i = tid; i < cw * ch; i += nt) {
    const int r = i / cw, c = i - r * cw;
    const int wr = cy0 + r - wy0, wc = cx0 + c - wx0;
    const int me = cur[(wr + 1) * Wp + wc + 1];
    const long long g = base + (long long)(cy0 + r) * W + cx0 + c;
    const bool hole = me != kNoLabel && s_cnt[me] <= max_area;
    scores_out[g] = hole ? fill_value : scores_in[g];
  }
}

// ---------------------------------------------------------------------------------------------
// Largest 3-D component of a binary volume (26-connectivity): the post-step of the reference's CT driver,
//   labels = skimage.measure.label(seg);  largest = labels == argmax(bincount(labels.flat)[1:]) + 1
// (medsam2_infer_3D_CT.py:76-79, 285).  Voxel-level union-find in global memory: every foreground voxel unites with its
// foreground neighbours among the 13 that precede it in raster order (links point to the smaller index, so a root is the
// component's first voxel in raster order -- the order in which skimage numbers its labels); areas are counted at the
// roots; the winner is the largest area, ties to the smallest root (np.argmax returns the first maximum).
// ---------------------------------------------------------------------------------------------
// init: a voxel's first link is the start of its run of consecutive foreground voxels along x (inside the 32 voxels its
// warp covers), found with one ballot -- solid regions start out as a few roots per row instead of one per voxel
__global__ void cc3_init(const uint8_t* __restrict__ vol, int32_t* __restrict__ par, int32_t* __restrict__ cnt,
                         long long n, unsigned long long* best, int W) {
  PDL_ENTRY();
  if (blockIdx.x == 0 && threadIdx.x == 0) *best = 0ull;
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long base = warp0 * 32; base < n; base += nwarps * 32) {
    const long long i = base + lane;
    const bool fg = i < n && vol[i];
    const unsigned bits = __ballot_sync(0xffffffffu, fg);
    const bool first = fg && (lane == 0 || (int)(i % W) == 0 || !((bits >> (lane - 1)) & 1u));
    const unsigned starts = __ballot_sync(0xffffffffu, first);
    if (i < n) {
      const int s0 = 31 - __clz((int)(starts & (0xffffffffu >> (31 - lane))));  // nearest run start at or below this lane
      par[i] = fg ? (int32_t)(base + s0) : -1;
      cnt[i] = 0;
    }
  }
}
// merge: unions with the 13 neighbours that precede the voxel in raster order, minus the redundant ones -- if the voxel's
// left neighbour p' and the neighbour's left neighbour q' are both foreground, p ~ p' and q ~ q' hold through their runs and
// p' ~ q' is the same direction's union one voxel to the left, so (p, q) adds nothing.  Inside a solid region every
// union is skipped; work is left at region boundaries and at the 32-voxel seams of the runs.
__global__ void cc3_merge(const uint8_t* __restrict__ vol, int32_t* par, int D, int H, int W) {
  PDL_ENTRY();
  const long long n = (long long)D * H * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    if (!vol[i]) continue;
    const int x = (int)(i % W), y = (int)((i / W) % H), z = (int)(i / ((long long)W * H));
    const bool left = x > 0 && vol[i - 1];
    if (left && (i & 31) == 0) uf_union(par, (int)i, (int)(i - 1));  // run seam between two warps' segments
#pragma unroll
    for (int dz = -1; dz <= 0; ++dz)
#pragma unroll
      for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) {
          if (dz == 0 && dy >= 0) continue;  // (0, 0, -1) is the run itself
          const int zz = z + dz, yy = y + dy, xx = x + dx;
          if (zz < 0 || yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
          const long long j = ((long long)zz * H + yy) * W + xx;
          if (!vol[j]) continue;
          if (left && xx > 0 && vol[j - 1]) continue;
          uf_union(par, (int)i, (int)j);
        }
  }
}
__global__ void cc3_count(int32_t* par, int32_t* __restrict__ cnt, long long n) {
  PDL_ENTRY();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    if (par[i] < 0) continue;
    const int root = uf_find(par, (int)i);
    // warp-aggregated: neighbouring voxels mostly share a root
    const unsigned peers = __match_any_sync(__activemask(), root);
    if ((threadIdx.x & 31) == (__ffs(peers) - 1)) atomicAdd(&cnt[root], __popc(peers));
  }
}
// key = area << 32 | (0xffffffff - root): the maximum key is the largest area, ties to the smallest root
__global__ void cc3_best(const int32_t* __restrict__ par, const int32_t* __restrict__ cnt, long long n,
                         unsigned long long* best) {
  PDL_ENTRY();
  unsigned long long mine = 0ull;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    if (cnt[i] > 0) {
      const unsigned long long key = ((unsigned long long)(unsigned)cnt[i] << 32) | (0xffffffffull - (unsigned)i);
      mine = key > mine ? key : mine;
    }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long other = __shfl_xor_sync(0xffffffffu, mine, o);
    mine = other > mine ? other : mine;
  }
  if ((threadIdx.x & 31) == 0 && mine) atomicMax(best, mine);
}
__global__ void cc3_select(int32_t* par, uint8_t* __restrict__ out, long long n, const unsigned long long* best) {
  PDL_ENTRY();
  const unsigned long long key = *best;
  const int root = key ? (int)(0xffffffffull - (key & 0xffffffffull)) : -2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = (par[i] >= 0 && uf_find(par, (int)i) == root) ? 1 : 0;
}

size_t smem_bytes(int H, int W) { return (size_t)(H / 2) * (W / 2) * 8 + (size_t)H * W; }
constexpr size_t kSmemLimit = 224 * 1024;

template <bool FILL>
int launch_smem(const uint8_t* img, const float* sin, float* sout, int32_t* labels, int32_t* counts, int N, int H,
                int W, int max_area, float fill, cudaStream_t s) {
  const size_t bytes = smem_bytes(H, W);
  static UsvmPerDeviceOnce configured = {};
  if (bytes > 48 * 1024 && usvm_need_setup(configured)) {
    if (cudaFuncSetAttribute(cc_smem_kernel<FILL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemLimit) !=
        cudaSuccess)
      return USVM_ERR_CUDA;
    usvm_setup_done(configured);
  }
  const int threads = (H * W >= 1024 * 4) ? 1024 : 256;
  usvm_launch(cc_smem_kernel<FILL>, dim3(N), dim3(threads), bytes, s, img, sin, sout, labels, counts, H, W, max_area, fill);
  return usvm_check_launch();
}

template <typename FG, bool FILL>
int launch_global(FG fg, int32_t* labels, int32_t* counts, const float* sin, float* sout, int N, int H, int W,
                  int max_area, float fill, cudaStream_t s) {
  dim3 block(32, 8);
  dim3 grid(cdiv(W / 2, 32), cdiv(H / 2, 8), N);
  usvm_launch(cc_g_init<FG>, dim3(grid), dim3(block), 0, s, fg, labels, counts, H, W);
  usvm_launch(cc_g_merge<FG>, dim3(grid), dim3(block), 0, s, fg, labels, H, W);
  usvm_launch(cc_g_compress<FG>, dim3(grid), dim3(block), 0, s, fg, labels, counts, H, W);
  usvm_launch(cc_g_final<FG, FILL>, dim3(grid), dim3(block), 0, s, fg, labels, counts, sin, sout, H, W, max_area, fill);
  usvm_launch(cc_g_roots<FG, FILL>, dim3(grid), dim3(block), 0, s, fg, labels, counts, sin, sout, H, W, max_area, fill);
  return usvm_check_launch();
}

}  // namespace

extern "C" int usvm_cc2d_label_u8(const uint8_t* img, int32_t* labels, int32_t* counts, int N, int H, int W,
                                  void* stream) {
  if (N < 0 || H <= 0 || W <= 0 || (H & 1) || (W & 1)) return USVM_ERR_ARG;
  if (N == 0) return USVM_OK;
  if (!img || !labels || !counts) return USVM_ERR_ARG;
  if ((long long)H * W >= (1LL << 31) - 1 || N > 65535) return USVM_ERR_ARG;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (smem_bytes(H, W) <= kSmemLimit)
    return launch_smem<false>(img, nullptr, nullptr, labels, counts, N, H, W, 0, 0.f, s);
  return launch_global<FgU8, false>(FgU8{img}, labels, counts, nullptr, nullptr, N, H, W, 0, 0.f, s);
}

extern "C" int usvm_fill_holes_f32(const float* scores_in, float* scores_out, int32_t* scratch_labels,
                                   int32_t* scratch_counts, int N, int H, int W, int max_area, float fill_value,
                                   void* stream) {
  if (N < 0 || H <= 0 || W <= 0 || (H & 1) || (W & 1) || max_area <= 0) return USVM_ERR_ARG;
  if (N == 0) return USVM_OK;
  if (!scores_in || !scores_out) return USVM_ERR_ARG;
  if ((long long)H * W >= (1LL << 31) - 1 || N > 65535) return USVM_ERR_ARG;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (max_area <= 32 && (H % 4) == 0 && (W % 4) == 0) {  // the propagation path's case: 128 x 128, max_area 8
    const int R = (max_area + 3) & ~3;
    const int win = kFillCore + 2 * R;
    const size_t local_bytes = ((size_t)2 * (win + 2) * (win + 2) + (size_t)win * win) * 4;
    static UsvmPerDeviceOnce configured = {};
    if (usvm_need_setup(configured)) {
      if (cudaFuncSetAttribute(fill_holes_local_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemLimit) !=
          cudaSuccess)
        return USVM_ERR_CUDA;
      usvm_setup_done(configured);
    }
    usvm_launch(fill_holes_local_kernel, dim3(cdiv(W, kFillCore), cdiv(H, kFillCore), N), dim3(256), local_bytes, s, scores_in,
                scores_out, H, W, max_area, fill_value, R);
    return usvm_check_launch();
  }
  if (smem_bytes(H, W) <= kSmemLimit)
    return launch_smem<true>(nullptr, scores_in, scores_out, nullptr, nullptr, N, H, W, max_area, fill_value, s);
  if (!scratch_labels || !scratch_counts) return USVM_ERR_ARG;  // the global path needs two int32 [N,H,W] maps
  return launch_global<FgScore, true>(FgScore{scores_in}, scratch_labels, scratch_counts, scores_in, scores_out, N,
                                      H, W, max_area, fill_value, s);
}

extern "C" int usvm_cc3d_largest_u8(const uint8_t* vol, uint8_t* out, int32_t* scratch_parent, int32_t* scratch_count,
                                    unsigned long long* scratch_best, int D, int H, int W, void* stream) {
  if (D <= 0 || H <= 0 || W <= 0 || !vol || !out || !scratch_parent || !scratch_count || !scratch_best)
    return USVM_ERR_ARG;
  const long long n = (long long)D * H * W;
  if (n >= (1LL << 31) - 1) return USVM_ERR_ARG;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int grid = (int)(n / 256 + 1 < 148 * 16 ? n / 256 + 1 : 148 * 16);
  usvm_launch(cc3_init, dim3(grid), dim3(256), 0, s, vol, scratch_parent, scratch_count, n, scratch_best, W);
  usvm_launch(cc3_merge, dim3(grid), dim3(256), 0, s, vol, scratch_parent, D, H, W);
  usvm_launch(cc3_count, dim3(grid), dim3(256), 0, s, scratch_parent, scratch_count, n);
  usvm_launch(cc3_best, dim3(grid), dim3(256), 0, s, scratch_parent, scratch_count, n, scratch_best);
  usvm_launch(cc3_select, dim3(grid), dim3(256), 0, s, scratch_parent, out, n, scratch_best);
  return usvm_check_launch();
}
