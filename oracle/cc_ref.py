"""ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle/medsam2_ref.py header).

CPU restatement of `sam2._C.get_connected_componnets` (sam2/csrc/connected_components.cu:213-282):
8-connectivity component labelling of uint8 [N,1,H,W] masks by block-based union-find on 2x2
pixel blocks.  The CUDA code's observable result (derived from union_/atomicMin :42-60 and
final_labeling :129-168) is

    labels[p] = 1 + min{ (r & ~1) * W + (c & ~1) : (r, c) foreground pixel of p's component }
    counts[p] = area of p's component                                   (both 0 on background)

Two independent restatements live here: `connected_components_ref` (scipy labelling + the
closed form above) and `connected_components_emulated` (a sequential emulation of the kernel's
own merge rules :72-117 with a plain union-find), cross-checked in tests/test_cc_oracle.py.  The
plain-C twin is oracle/cc_ref.c (built by oracle/Makefile into oracle/_build/).
"""
import numpy as np
from scipy import ndimage


def connected_components_ref(img):
    """img: uint8 [N,1,H,W] (nonzero = foreground) -> (labels int32, counts int32)."""
    img = np.asarray(img)
    assert img.ndim == 4 and img.shape[1] == 1
    N, _, H, W = img.shape
    assert H % 2 == 0 and W % 2 == 0, "height and width must be even (connected_components.cu:226-227)"
    labels = np.zeros((N, 1, H, W), np.int32)
    counts = np.zeros((N, 1, H, W), np.int32)
    rr, cc = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    anchor = ((rr & ~1) * W + (cc & ~1)).astype(np.int64)
    for n in range(N):
        fg = img[n, 0] != 0
        comp, k = ndimage.label(fg, structure=np.ones((3, 3), np.int32))
        if k == 0:
            continue
        idx = np.arange(1, k + 1)
        root = ndimage.minimum(anchor, comp, idx).astype(np.int64)
        area = ndimage.sum(fg, comp, idx).astype(np.int64)
        lut_root = np.zeros(k + 1, np.int64)
        lut_area = np.zeros(k + 1, np.int64)
        lut_root[1:] = root + 1
        lut_area[1:] = area
        labels[n, 0] = lut_root[comp]
        counts[n, 0] = lut_area[comp]
    return labels, counts


def connected_components_emulated(img):
    """Sequential emulation of the kernels' own merge rules (connected_components.cu:72-117):
    one union-find node per 2x2 block anchor; a block is united with its top-left / top /
    top-right / left neighbour block when one of its TL / TR / BL pixels 8-touches a
    foreground pixel of that neighbour; links always point to the smaller anchor index."""
    img = np.asarray(img)
    N, _, H, W = img.shape
    labels = np.zeros((N, 1, H, W), np.int32)
    counts = np.zeros((N, 1, H, W), np.int32)
    for n in range(N):
        im = img[n, 0] != 0
        parent = {r * W + c: r * W + c for r in range(0, H, 2) for c in range(0, W, 2)}

        def find(a):
            while parent[a] != a:
                a = parent[a]
            return a

        def union(a, b):
            a, b = find(a), find(b)
            if a < b:
                parent[b] = a
            elif b < a:
                parent[a] = b

        def px(r, c):
            return 0 <= r < H and 0 <= c < W and bool(im[r, c])

        for r in range(0, H, 2):
            for c in range(0, W, 2):
                idx = r * W + c
                tl, tr, bl = px(r, c), px(r, c + 1), px(r + 1, c)
                if tl and px(r - 1, c - 1):
                    union(idx, idx - 2 * W - 2)
                if (tl or tr) and (px(r - 1, c) or px(r - 1, c + 1)):
                    union(idx, idx - 2 * W)
                if tr and px(r - 1, c + 2):
                    union(idx, idx - 2 * W + 2)
                if (tl or bl) and (px(r, c - 1) or px(r + 1, c - 1)):
                    union(idx, idx - 2)
        area = {}
        for r in range(H):
            for c in range(W):
                if im[r, c]:
                    root = find((r & ~1) * W + (c & ~1))
                    area[root] = area.get(root, 0) + 1
        for r in range(H):
            for c in range(W):
                if im[r, c]:
                    root = find((r & ~1) * W + (c & ~1))
                    labels[n, 0, r, c] = root + 1
                    counts[n, 0, r, c] = area[root]
    return labels, counts


def largest_component_3d_ref(seg):
    """getLargestCC of medsam2_infer_3D_CT.py:76-79 restated with scipy: skimage.measure.label's default connectivity for
    a 3-D input is full (26 neighbours) and it numbers components in raster order of their first voxel, exactly like
    scipy.ndimage.label with a 3x3x3 structure; np.argmax takes the first maximum."""
    seg = np.asarray(seg) != 0
    labels, k = ndimage.label(seg, structure=np.ones((3, 3, 3), np.int32))
    if k == 0:
        return np.zeros(seg.shape, bool)
    return labels == (np.argmax(np.bincount(labels.ravel())[1:]) + 1)
