/* ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle/medsam2_ref.py header).
 *
 * Plain-C restatement of sam2._C.get_connected_componnets
 * (reference: sam2/csrc/connected_components.cu:62-282): 8-connectivity labelling by union-find
 * over 2x2 pixel blocks with the kernel's own merge rules (:72-117), min-anchor roots
 * (union_/atomicMin :42-60), label = root + 1 on foreground (:129-168), per-pixel component
 * area (:170-210).  Sequential; used as the CPU checker and as bench.py's cpu_baseline for the
 * hole-filling kernel.  Build: `make -C oracle` -> oracle/_build/libcc_ref.so.
 */
#include <stdint.h>
#include <stdlib.h>

static int32_t find_root(const int32_t* parent, int32_t a) {
  while (parent[a] != a) a = parent[a];
  return a;
}

static void unite(int32_t* parent, int32_t a, int32_t b) {
  a = find_root(parent, a);
  b = find_root(parent, b);
  if (a < b) parent[b] = a;
  else if (b < a) parent[a] = b;
}

/* img: uint8 [N,H,W] (nonzero = foreground); labels, counts: int32 [N,H,W].  Returns 0, or -1 on
 * odd H/W (the reference asserts even sizes, connected_components.cu:226-227). */
int cc_ref_label_u8(const uint8_t* img, int32_t* labels, int32_t* counts, int N, int H, int W) {
  if ((H & 1) || (W & 1) || N < 0) return -1;
  const size_t hw = (size_t)H * (size_t)W;
  int32_t* parent = (int32_t*)malloc(hw * sizeof(int32_t));
  int32_t* area = (int32_t*)malloc(hw * sizeof(int32_t));
  if (!parent || !area) { free(parent); free(area); return -2; }
  for (int n = 0; n < N; ++n) {
    const uint8_t* im = img + n * hw;
    int32_t* lab = labels + n * hw;
    int32_t* cnt = counts + n * hw;
#define PX(r, c) ((r) >= 0 && (r) < H && (c) >= 0 && (c) < W && im[(size_t)(r) * W + (c)])
    for (int r = 0; r < H; r += 2)
      for (int c = 0; c < W; c += 2) { parent[(size_t)r * W + c] = r * W + c; area[(size_t)r * W + c] = 0; }
    for (int r = 0; r < H; r += 2)
      for (int c = 0; c < W; c += 2) {
        const int32_t idx = r * W + c;
        const int tl = PX(r, c), tr = PX(r, c + 1), bl = PX(r + 1, c);
        if (tl && PX(r - 1, c - 1)) unite(parent, idx, idx - 2 * W - 2);
        if ((tl || tr) && (PX(r - 1, c) || PX(r - 1, c + 1))) unite(parent, idx, idx - 2 * W);
        if (tr && PX(r - 1, c + 2)) unite(parent, idx, idx - 2 * W + 2);
        if ((tl || bl) && (PX(r, c - 1) || PX(r + 1, c - 1))) unite(parent, idx, idx - 2);
      }
    for (int r = 0; r < H; ++r)
      for (int c = 0; c < W; ++c)
        if (im[(size_t)r * W + c]) area[find_root(parent, (r & ~1) * W + (c & ~1))] += 1;
    for (int r = 0; r < H; ++r)
      for (int c = 0; c < W; ++c) {
        const size_t p = (size_t)r * W + c;
        if (im[p]) {
          const int32_t root = find_root(parent, (r & ~1) * W + (c & ~1));
          lab[p] = root + 1;
          cnt[p] = area[root];
        } else {
          lab[p] = 0;
          cnt[p] = 0;
        }
      }
#undef PX
  }
  free(parent);
  free(area);
  return 0;
}

/* fill_holes_in_mask_scores (sam2/utils/misc.py:312-338): background = (score <= 0); components of
 * background with area <= max_area are overwritten with `fill` (0.1).  In place on fp32 [N,H,W]. */
int cc_ref_fill_holes_f32(float* scores, int N, int H, int W, int max_area, float fill) {
  const size_t total = (size_t)N * H * W;
  uint8_t* bg = (uint8_t*)malloc(total);
  int32_t* lab = (int32_t*)malloc(total * sizeof(int32_t));
  int32_t* cnt = (int32_t*)malloc(total * sizeof(int32_t));
  if (!bg || !lab || !cnt) { free(bg); free(lab); free(cnt); return -2; }
  for (size_t i = 0; i < total; ++i) bg[i] = scores[i] <= 0.0f;
  int rc = cc_ref_label_u8(bg, lab, cnt, N, H, W);
  if (rc == 0)
    for (size_t i = 0; i < total; ++i)
      if (lab[i] > 0 && cnt[i] <= max_area) scores[i] = fill;
  free(bg); free(lab); free(cnt);
  return rc;
}
