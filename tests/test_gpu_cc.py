"""(a14) connected components / hole filling: CUDA kernel vs the CPU oracle, bit-exact, through the C ABI."""
import numpy as np
import pytest
import torch

from oracle.cc_ref import connected_components_ref

pytestmark = pytest.mark.gpu


def _run(img_np):
    from sam2 import _C  # the drop-in op, same spelling as the reference

    x = torch.from_numpy(img_np).cuda()
    labels, counts = _C.get_connected_componnets(x)
    torch.cuda.synchronize()
    return labels.cpu().numpy(), counts.cpu().numpy()


@pytest.mark.parametrize("shape", [(1, 1, 2, 2), (3, 1, 128, 128), (2, 1, 64, 96), (5, 1, 30, 18), (1, 1, 256, 256),
                                   (2, 1, 512, 640)])
@pytest.mark.parametrize("density", [0.0, 0.08, 0.45, 0.6, 0.93, 1.0])
def test_labels_and_counts_bit_exact(shape, density):
    rng = np.random.default_rng(hash((shape, density)) % (2 ** 32))
    img = (rng.random(shape) < density).astype(np.uint8)
    want_l, want_c = connected_components_ref(img)
    got_l, got_c = _run(img)
    assert got_l.dtype == np.int32 and got_c.dtype == np.int32
    assert np.array_equal(got_l, want_l)
    assert np.array_equal(got_c, want_c)


def test_structured_masks_bit_exact():
    """Spirals / checkerboards / diagonals: long union-find chains and pure 8-connectivity joins."""
    H = W = 128
    imgs = np.zeros((4, 1, H, W), np.uint8)
    yy, xx = np.mgrid[0:H, 0:W]
    imgs[0, 0] = ((yy + xx) % 2 == 0)                      # checkerboard: one 8-connected component
    imgs[1, 0] = (yy == xx) | (yy + xx == W - 1)           # two crossing diagonals
    imgs[2, 0] = ((yy // 2) % 2 == 0) & ~((xx == W - 1) & ((yy // 4) % 2 == 0)) & ~((xx == 0) & ((yy // 4) % 2 == 1))
    imgs[3, 0] = (yy % 4 == 0) | ((xx % 8 == 0) & (yy % 8 < 4))
    want_l, want_c = connected_components_ref(imgs)
    got_l, got_c = _run(imgs)
    assert np.array_equal(got_l, want_l) and np.array_equal(got_c, want_c)


def test_empty_batch_and_argument_errors():
    from sam2 import _C

    l, c = _C.get_connected_componnets(torch.zeros((0, 1, 8, 8), dtype=torch.uint8, device="cuda"))
    assert l.shape == (0, 1, 8, 8) and c.shape == (0, 1, 8, 8)
    with pytest.raises(RuntimeError):
        _C.get_connected_componnets(torch.zeros((1, 1, 8, 8), dtype=torch.uint8))  # CPU tensor
    with pytest.raises(RuntimeError):
        _C.get_connected_componnets(torch.zeros((1, 1, 7, 8), dtype=torch.uint8, device="cuda"))  # odd height
    with pytest.raises(RuntimeError):
        _C.get_connected_componnets(torch.zeros((1, 1, 8, 8), dtype=torch.float32, device="cuda"))
    with pytest.raises(RuntimeError):
        _C.get_connected_componnets(torch.zeros((1, 2, 8, 8), dtype=torch.uint8, device="cuda"))


@pytest.mark.parametrize("max_area", [8, 3, 40])
@pytest.mark.parametrize("shape", [(4, 1, 128, 128), (1, 1, 512, 640), (3, 1, 64, 96)])
def test_fill_holes_matches_reference_composition(shape, max_area):
    """fused kernel == where((labels > 0) & (areas <= 8), 0.1, mask) with labels/areas of (mask <= 0)."""
    from sam2.utils.misc import fill_holes_in_mask_scores

    g = torch.Generator().manual_seed(7)
    scores = torch.randn(shape, generator=g) * 0.07 + 0.02
    scores[0, 0, :4, :4] = 0.0  # exact zeros count as background
    labels, areas = connected_components_ref((scores <= 0).numpy().astype(np.uint8))
    want = torch.where(torch.from_numpy((labels > 0) & (areas <= max_area)), torch.full_like(scores, 0.1), scores)
    got = fill_holes_in_mask_scores(scores.cuda(), max_area).cpu()
    assert torch.equal(got, want)
    assert int((got != scores).sum()) > 0  # the case really exercises the fill


def test_fill_holes_structured_masks():
    """Thin / diagonal / ring-shaped background structures around the max_area threshold (local-propagation kernel)."""
    from sam2.utils.misc import fill_holes_in_mask_scores

    H = W = 128
    s = torch.ones((5, 1, H, W))
    s[0, 0, 10, 10:18] = -1          # horizontal bar, area 8  -> filled
    s[0, 0, 20, 10:19] = -1          # area 9                  -> kept
    for i in range(8):
        s[1, 0, 30 + i, 40 + i] = -1  # diagonal, area 8 (8-connected) -> filled
    for i in range(9):
        s[1, 0, 60 + i, 40 - i] = -1  # anti-diagonal, area 9 -> kept
    s[2, 0, 50:53, 50:53] = -1
    s[2, 0, 51, 51] = 1              # ring of 8 around a foreground pixel -> filled
    s[3, 0] = -1                     # everything background: one huge component -> kept
    s[3, 0, 64, 64] = 1
    yy, xx = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    s[4, 0][((yy + xx) % 2 == 0)] = -1  # checkerboard: one giant 8-connected background component
    s[4, 0, 0:2, :] = 1
    s[4, 0, 0, 5] = -1               # isolated single pixel in the cleared band -> filled
    labels, areas = connected_components_ref((s <= 0).numpy().astype(np.uint8))
    want = torch.where(torch.from_numpy((labels > 0) & (areas <= 8)), torch.full_like(s, 0.1), s)
    got = fill_holes_in_mask_scores(s.cuda(), 8).cpu()
    assert torch.equal(got, want)
    assert float(got[0, 0, 10, 12]) == pytest.approx(0.1) and float(got[0, 0, 20, 12]) == -1.0


def test_idempotent_and_area_checksum():
    """Size-independent properties at the full batch of BASELINE config 3 (4 objects x many frames)."""
    from sam2 import _C

    g = torch.Generator().manual_seed(3)
    img = (torch.rand((256, 1, 128, 128), generator=g) < 0.5).to(torch.uint8).cuda()
    labels, counts = _C.get_connected_componnets(img)
    assert torch.equal((labels > 0), img.bool())
    # every pixel of a component reports the same area, and the areas of the distinct components sum to #fg
    flat_l, flat_c = labels.flatten(1), counts.flatten(1)
    for n in (0, 17, 255):
        roots, inv = torch.unique(flat_l[n][flat_l[n] > 0], return_inverse=True)
        area = torch.zeros(len(roots), dtype=torch.int64, device="cuda").scatter_add_(
            0, inv, torch.ones_like(inv, dtype=torch.int64))
        assert torch.equal(area[inv], flat_c[n][flat_l[n] > 0].long())
        assert int(area.sum()) == int(img[n].sum())
    # labelling the "label > 0" mask again gives the same labels
    l2, c2 = _C.get_connected_componnets((labels > 0).to(torch.uint8))
    assert torch.equal(l2, labels) and torch.equal(c2, counts)


# ------------------------------------------------------------------------------------------------
# against the REFERENCE's own CUDA kernel (oracle/_ref/ref_sam2_C.so, compiled for sm_100a from
# /root/reference/sam2/csrc/connected_components.cu by oracle/build_ref_cc.py in the build container)
# ------------------------------------------------------------------------------------------------
def _reference_op():
    from oracle.build_ref_cc import load

    op = load()
    if op is None:
        pytest.skip("oracle/_ref/ref_sam2_C.so not built (python oracle/build_ref_cc.py in the build container)")
    return op


@pytest.mark.parametrize("shape", [(1, 1, 2, 2), (4, 1, 128, 128), (2, 1, 64, 96), (3, 1, 30, 18), (1, 1, 256, 256),
                                   (2, 1, 512, 640), (64, 1, 128, 128)])
@pytest.mark.parametrize("density", [0.0, 0.1, 0.45, 0.6, 0.9, 1.0])
def test_bit_exact_against_reference_kernel(shape, density):
    ref = _reference_op()
    rng = np.random.default_rng(hash(("ref", shape, density)) % (2 ** 32))
    img = (rng.random(shape) < density).astype(np.uint8)
    x = torch.from_numpy(img).cuda()
    want_l, want_c = ref(x)
    torch.cuda.synchronize()
    got_l, got_c = _run(img)
    assert np.array_equal(got_l, want_l.cpu().numpy())
    assert np.array_equal(got_c, want_c.cpu().numpy())
    # ... and the CPU oracle used by every GPU-less test is pinned to the same kernel
    o_l, o_c = connected_components_ref(img)
    assert np.array_equal(o_l, want_l.cpu().numpy()) and np.array_equal(o_c, want_c.cpu().numpy())


def test_fill_holes_against_reference_kernel_composition():
    """fill_holes_in_mask_scores (sam2/utils/misc.py:312-338) composed from the reference kernel vs the fused kernel."""
    from us_video_medsam2_b200.cc import fill_holes_in_mask_scores

    ref = _reference_op()
    g = torch.Generator().manual_seed(5)
    for n, scale in ((1, 0.05), (4, 0.5), (16, 1.0)):
        mask = (torch.randn((n, 1, 128, 128), generator=g) * scale + 0.3 * scale).cuda()
        mask[:, :, 40:60, 40:60] = mask[:, :, 40:60, 40:60].abs() + 0.01   # a solid blob with a few punched holes
        mask[:, :, 45, 45] = -1.0
        mask[:, :, 50:52, 50:53] = -0.5
        labels, areas = ref((mask <= 0).to(torch.uint8))
        is_hole = (labels > 0) & (areas <= 8)
        want = torch.where(is_hole, torch.full_like(mask, 0.1), mask)
        got = fill_holes_in_mask_scores(mask, 8)
        assert torch.equal(got, want)
        assert int(is_hole.sum()) > 0


def test_postprocess_masks_against_reference_kernel_composition():
    """SAM2Transforms.postprocess_masks (sam2/utils/transforms.py:76-115), the second caller of the CC op: hole filling and
    sprinkle removal at a non-zero threshold, then the resize -- against the same formula evaluated with the reference's
    own kernel."""
    from sam2.utils.transforms import SAM2Transforms

    ref = _reference_op()
    tr = SAM2Transforms(resolution=512, mask_threshold=0.25, max_hole_area=12, max_sprinkle_area=7)
    g = torch.Generator().manual_seed(11)
    masks = (torch.randn((2, 3, 128, 128), generator=g) * 0.6 + 0.3).cuda()
    masks[:, :, 30:70, 30:70] = masks[:, :, 30:70, 30:70].abs() + 0.3    # a blob with punched holes ...
    masks[:, :, 40:42, 40:43] = -1.0
    masks[:, :, 90:120, 10:60] = -masks[:, :, 90:120, 10:60].abs()          # ... and an empty region with sprinkles
    masks[:, :, 100, 30] = 2.0
    masks[:, :, 105:107, 40:42] = 1.5
    got = tr.postprocess_masks(masks, (300, 420))
    flat = masks.flatten(0, 1).unsqueeze(1)
    lab, area = ref((flat <= 0.25).to(torch.uint8))
    hole = ((lab > 0) & (area <= 12)).reshape_as(masks)
    want = torch.where(hole, 0.25 + 10.0, masks)
    lab, area = ref((flat > 0.25).to(torch.uint8))
    spr = ((lab > 0) & (area <= 7)).reshape_as(masks)
    want = torch.where(spr, 0.25 - 10.0, want)
    assert int(hole.sum()) > 0 and int(spr.sum()) > 0
    want = torch.nn.functional.interpolate(want, (300, 420), mode="bilinear", align_corners=False)
    assert got.shape == (2, 3, 300, 420) and torch.equal(got, want)
    pts = torch.tensor([[[210.0, 150.0]]])
    assert torch.allclose(tr.transform_coords(pts, normalize=True, orig_hw=(300, 420)), torch.tensor([[[256.0, 256.0]]]))


@pytest.mark.parametrize("shape,density,seed", [((9, 32, 40), 0.3, 0), ((16, 64, 64), 0.12, 1), ((5, 20, 28), 0.6, 2),
                                                ((64, 128, 128), 0.05, 3), ((1, 16, 16), 0.4, 4),
                                                ((24, 96, 100), 0.93, 5), ((12, 70, 33), 0.5, 6), ((6, 40, 129), 0.8, 7)])
def test_largest_component_3d_matches_scipy(shape, density, seed):
    """usvm_cc3d_largest_u8 (the CT driver's getLargestCC on the device) bit-exact against the scipy restatement, on
    random volumes with blobs -- including ties in area (small random components) and 26-connectivity-only links."""
    from oracle.cc_ref import largest_component_3d_ref
    from us_video_medsam2_b200.cc import get_largest_cc

    rng = np.random.default_rng(seed)
    vol = rng.random(shape) < density
    D, H, W = shape
    if D > 4:  # two solid blobs touching only through a diagonal voxel pair
        vol[1:4, 2:8, 2:8] = True
        vol[4:6, 8:12, 8:12] = True
    got = get_largest_cc(torch.from_numpy(vol).cuda())
    assert got.dtype == torch.bool and got.shape == shape
    assert np.array_equal(got.cpu().numpy(), largest_component_3d_ref(vol))


def test_largest_component_3d_edge_cases():
    from oracle.cc_ref import largest_component_3d_ref
    from us_video_medsam2_b200.cc import get_largest_cc

    empty = torch.zeros((4, 8, 8), dtype=torch.uint8, device="cuda")
    assert not bool(get_largest_cc(empty).any())
    full = torch.ones((3, 6, 10), dtype=torch.uint8, device="cuda")
    assert bool(get_largest_cc(full).all())
    tie = np.zeros((3, 8, 8), bool)   # two components of equal size: the one that comes first in raster order wins
    tie[0, 0:2, 0:2] = True
    tie[2, 5:7, 5:7] = True
    got = get_largest_cc(torch.from_numpy(tie).cuda()).cpu().numpy()
    assert np.array_equal(got, largest_component_3d_ref(tie)) and got[0, 0, 0] and not got[2, 5, 5]
    with pytest.raises(RuntimeError):
        get_largest_cc(torch.zeros((4, 8, 8), dtype=torch.uint8))


def test_fill_holes_across_tile_boundaries():
    """The fill kernel is tiled (32 x 32 cores + a halo of max_area): small and just-too-large background structures that
    straddle core boundaries, touch the halo's outermost ring or the image edge, and a long thin component whose fragment
    inside a window is small (it must stay open through the foreign border)."""
    from sam2.utils.misc import fill_holes_in_mask_scores

    H, W = 128, 160
    s = torch.ones((3, 1, H, W))
    s[0, 0, 31, 28:36] = -1           # area 8 across x = 32            -> filled
    s[0, 0, 40, 28:37] = -1           # area 9 across x = 32            -> kept
    s[0, 0, 60:68, 63] = -1           # vertical bar of 8 across y = 64 -> filled
    for i in range(8):
        s[0, 0, 92 + i, 92 + i] = -1  # diagonal across (96, 96)        -> filled
    s[0, 0, 0, 60:68] = -1            # on the image edge, across x = 64 -> filled
    s[0, 0, H - 1, W - 8:W] = -1      # in the image corner              -> filled
    s[1, 0, 33, :] = -1               # one image-wide line: every window sees a short-looking fragment -> kept
    s[1, 0, 20:60, 70] = -1           # long vertical line                -> kept
    s[1, 0, 100, 30:38] = -1          # area 8 next to nothing            -> filled
    s[2, 0, 24:40, 24:40] = -1        # 16 x 16 block over a core corner  -> kept
    s[2, 0, 31:33, 95:97] = -1        # 2 x 2 over the corner (32, 96)    -> filled
    s[2, 0, 70, 56:64] = -1           # area 8 ending at x = 63           -> filled
    s[2, 0, 71, 64] = -1              # ... touching it diagonally: now area 9 -> both kept
    labels, areas = connected_components_ref((s <= 0).numpy().astype(np.uint8))
    want = torch.where(torch.from_numpy((labels > 0) & (areas <= 8)), torch.full_like(s, 0.1), s)
    got = fill_holes_in_mask_scores(s.cuda(), 8).cpu()
    assert torch.equal(got, want)
    assert float(got[0, 0, 31, 30]) == pytest.approx(0.1) and float(got[0, 0, 40, 30]) == -1.0
    assert float(got[1, 0, 33, 5]) == -1.0 and float(got[2, 0, 70, 60]) == -1.0
