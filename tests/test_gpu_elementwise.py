"""Bandwidth / layout / small-conv / decoder-tail kernels vs plain PyTorch fp32 references of the same op."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


def _g(seed=0):
    return torch.Generator(device="cuda").manual_seed(seed)


@pytest.mark.parametrize("rows,C,eps", [(1000, 96, 1e-6), (4096, 768, 1e-6), (2048, 256, 1e-5), (7, 64, 1e-6),
                                        (777, 224, 1e-6), (1030, 448, 1e-6), (515, 896, 1e-6), (333, 100, 1e-6),
                                        (64, 1000, 1e-6), (16390, 96, 1e-6), (4099, 192, 1e-6)])
def test_layernorm(rows, C, eps):
    from us_video_medsam2_b200 import ops

    g = _g(rows)
    x = torch.randn((rows, C), generator=g, device="cuda") * 3 + 1
    w, b = torch.randn(C, generator=g, device="cuda"), torch.randn(C, generator=g, device="cuda")
    o32, o16 = ops.layernorm(x, w, b, eps, f32=True, bf16=True)
    want = F.layer_norm(x, (C,), w, b, eps)
    assert (o32 - want).abs().max().item() < 2e-5
    assert (o16.float() - want).abs().max().item() < 4e-2
    o32g, _ = ops.layernorm(x, w, b, eps, f32=True, gelu=True)
    assert (o32g - F.gelu(want)).abs().max().item() < 2e-5


def test_layernorm_over_the_real_channels_of_a_padded_row():
    """Hiera-B+ stage 1: 112 channels in rows of 128 (engine._cpad); the norm sees 112, the padding of the output is zero."""
    from us_video_medsam2_b200 import ops

    g = _g(5)
    x = torch.zeros((999, 128), device="cuda")
    x[:, :112] = torch.randn((999, 112), generator=g, device="cuda") * 2 - 0.5
    w, b = torch.randn(112, generator=g, device="cuda"), torch.randn(112, generator=g, device="cuda")
    o32, o16 = ops.layernorm(x, w, b, 1e-6, f32=True, bf16=True, valid=112)
    want = F.layer_norm(x[:, :112], (112,), w, b, 1e-6)
    assert o32.shape == (999, 128) and (o32[:, :112] - want).abs().max().item() < 2e-5
    assert (o16[:, :112].float() - want).abs().max().item() < 4e-2
    assert not o32[:, 112:].any() and not o16[:, 112:].any()


def test_axpby_and_cast():
    from us_video_medsam2_b200 import ops

    g = _g(1)
    x = torch.randn((1024, 256), generator=g, device="cuda")
    y = torch.randn((1024, 256), generator=g, device="cuda")
    o, _ = ops.axpby(x, y, 1.0, 0.1, rows=3 * 1024, x_mod=1024, y_mod=1024)
    assert torch.allclose(o, (x + 0.1 * y).repeat(3, 1), atol=1e-6)
    row = torch.randn((1, 256), generator=g, device="cuda")
    o, _ = ops.axpby(x, row, rows=1024, y_mod=1)
    assert torch.allclose(o, x + row, atol=1e-6)
    assert torch.equal(ops.cast_bf16(x), x.to(BF))


def test_rope_matches_complex_multiply():
    from oracle.medsam2_ref import apply_rope, axial_rope_table
    from us_video_medsam2_b200 import ops
    from us_video_medsam2_b200.engine import _rope_tables

    c, s = _rope_tables(256, 32, 32)
    c2, s2 = axial_rope_table(256, 32, 32)
    assert torch.equal(c, c2) and torch.equal(s, s2)
    g = _g(2)
    B, Nk, n_ptr = 2, 2 * 1024 + 12, 12
    x = torch.randn((B * Nk, 768), generator=g, device="cuda")
    got = ops.rope(x, 256, c.cuda(), s.cuda(), Nk, Nk - n_ptr)
    xv = x[:, 256:512].cpu().view(B, Nk, 256)
    want = torch.cat([apply_rope(xv[:, : Nk - n_ptr], c.repeat(2, 1), s.repeat(2, 1)), xv[:, Nk - n_ptr:]], dim=1)
    assert (got.float().cpu().view(B, Nk, 256) - want).abs().max().item() < 2e-2
    assert torch.equal(got.view(B, Nk, 256)[:, Nk - n_ptr:].cpu(), xv[:, Nk - n_ptr:].to(BF))


@pytest.mark.parametrize("Hg,ws,pool,heads", [(128, 8, False, 1), (128, 8, True, 2), (64, 4, False, 2), (64, 4, True, 4),
                                              (32, 14, False, 4), (32, 14, True, 8), (16, 7, False, 8)])
def test_window_gather_scatter(Hg, ws, pool, heads):
    """vs window_partition / q max-pool / window_unpartition restated in oracle/medsam2_ref.py."""
    from oracle.medsam2_ref import RefModel
    from us_video_medsam2_b200 import ops

    Fr, C = 2, heads * 96
    g = _g(Hg + ws)
    qkv = torch.randn((Fr, Hg, Hg, 3 * C), generator=g, device="cuda").to(BF)
    bias = torch.randn(3 * C, generator=g, device="cuda")
    Qw, Kw, Vw, nw, nq, nk = ops.window_gather(qkv, bias, Fr, Hg, Hg, ws, pool, C)
    # reference: pad positions hold the bias (qkv of a zero row), then partition
    x = qkv.float().cpu()
    pad = (ws - Hg % ws) % ws
    if pad:
        xp = bias.to(BF).float().cpu().expand(Fr, Hg + pad, Hg + pad, 3 * C).clone()
        xp[:, :Hg, :Hg] = x
        x = xp
    win, _ = RefModel._to_windows(x, ws)
    q, k, v = win[..., :C], win[..., C:2 * C], win[..., 2 * C:]
    if pool:
        q = RefModel._maxpool2(q)
    assert torch.equal(Kw.float().cpu(), k.reshape(Fr * nw, nk, C))
    assert torch.equal(Vw.float().cpu(), v.reshape(Fr * nw, nk, C))
    assert torch.equal(Qw.float().cpu(), q.reshape(Fr * nw, nq, C))
    wq = ws // 2 if pool else ws
    Ho = Hg // 2 if pool else Hg
    out = ops.window_scatter(Qw, Fr, Ho, Ho, wq, C)
    hp = Ho + (wq - Ho % wq) % wq
    want = RefModel._from_windows(q, wq, (hp, hp), (Ho, Ho))
    assert torch.equal(out.float().cpu().view(Fr, Ho, Ho, C), want)


def test_maxpool_upsample_im2col():
    from us_video_medsam2_b200 import ops

    g = _g(4)
    x = torch.randn((2, 64, 64, 192), generator=g, device="cuda")
    y = ops.maxpool2(x.view(-1, 192), 2, 64, 64, 192)
    want = F.max_pool2d(x.permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1)
    assert torch.equal(y.view(2, 32, 32, 192), want)
    fine = torch.randn((2, 32, 32, 256), generator=g, device="cuda")
    coarse = torch.randn((2, 16, 16, 256), generator=g, device="cuda")
    want = fine + F.interpolate(coarse.permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest").permute(0, 2, 3, 1)
    f2 = fine.clone().view(-1, 256)
    b16 = ops.upsample2_add_(f2, coarse.view(-1, 256), 2, 32, 32, 256, bf16=True)
    assert torch.allclose(f2.view_as(want), want, atol=1e-6) and torch.equal(b16, f2.to(BF))
    img = torch.randn((2, 3, 512, 512), generator=g, device="cuda")
    A = ops.im2col_patch(img)
    cols = F.unfold(img, kernel_size=7, padding=3, stride=4).transpose(1, 2).reshape(-1, 147)
    assert torch.equal(A[:, :147], cols.to(BF)) and float(A[:, 147:].abs().max()) == 0.0


def test_normalize_gray_u8():
    from us_video_medsam2_b200 import ops, synth

    g = torch.randint(0, 256, (3, 64, 48), dtype=torch.uint8, device="cuda")
    out = ops.normalize_gray_u8(g, synth.IMG_MEAN, synth.IMG_STD)
    mean = torch.tensor(synth.IMG_MEAN, device="cuda")[None, :, None, None]
    std = torch.tensor(synth.IMG_STD, device="cuda")[None, :, None, None]
    want = (g.float()[:, None] / 255.0 - mean) / std
    assert torch.allclose(out, want, atol=1e-6)


def test_build_memory_and_finalize():
    from us_video_medsam2_b200 import ops

    g = _g(5)
    B, T, Cm, P = 2, 1024, 64, 5
    frames = [torch.randn((B, T, Cm), generator=g, device="cuda").to(BF) for _ in range(3)]
    pos = torch.randn((T, Cm), generator=g, device="cuda")
    tpos = torch.randn((7, Cm), generator=g, device="cuda")
    ptrs = torch.randn((B, P * 4, Cm), generator=g, device="cuda")
    ppos = torch.randn((P * 4, Cm), generator=g, device="cuda")
    rows = [6, 0, 3]
    k_in, v_in, Nk = ops.build_memory(frames, rows, pos, tpos, ptrs, ppos, B)
    assert Nk == 3 * T + 4 * P
    wk = torch.cat([f.float() + (pos + tpos[r]) for f, r in zip(frames, rows)] + [ptrs + ppos], dim=1)
    wv = torch.cat([f.float() for f in frames] + [ptrs], dim=1)
    assert torch.equal(k_in, wk.to(BF)) and torch.equal(v_in, wv.to(BF))
    x = torch.randn((B * T, Cm), generator=g, device="cuda")
    score = torch.tensor([0.3, -0.2], device="cuda")
    emb = torch.randn(Cm, generator=g, device="cuda")
    mem = ops.finalize_memory(x, score, emb, B)
    want = x.view(B, T, Cm).clone()
    want[1] += emb
    assert torch.equal(mem, want.to(BF))


@pytest.mark.parametrize("Cin,Cout,k,s,p,H", [(1, 4, 3, 2, 1, 512), (4, 16, 3, 2, 1, 256), (16, 64, 3, 2, 1, 128),
                                              (1, 4, 2, 2, 0, 128), (4, 16, 2, 2, 0, 64), (1, 1, 4, 4, 0, 512),
                                              (4, 16, 3, 2, 1, 36), (16, 64, 3, 1, 1, 20)])
def test_conv2d_small(Cin, Cout, k, s, p, H):
    from us_video_medsam2_b200 import ops

    g = _g(Cin * 10 + k)
    B = 2
    x = torch.randn((B, Cin, H, H), generator=g, device="cuda")
    w = torch.randn((Cout, Cin, k, k), generator=g, device="cuda") / (Cin * k * k) ** 0.5
    b = torch.randn(Cout, generator=g, device="cuda")
    lw, lb = torch.randn(Cout, generator=g, device="cuda"), torch.randn(Cout, generator=g, device="cuda")
    use_ln = Cout > 1
    out, Ho, Wo = ops.conv2d_small(x.permute(0, 2, 3, 1).contiguous(), w.permute(2, 3, 1, 0).contiguous(), b, B, H, H,
                                   Cin, Cout, k, s, p, ln=(lw, lb) if use_ln else None, gelu=use_ln)
    y = F.conv2d(x, w, b, stride=s, padding=p)
    if use_ln:
        u = y.mean(1, keepdim=True)
        v = (y - u).pow(2).mean(1, keepdim=True)
        y = F.gelu((y - u) / torch.sqrt(v + 1e-6) * lw[None, :, None, None] + lb[None, :, None, None])
    assert (out.view(B, Ho, Wo, Cout) - y.permute(0, 2, 3, 1)).abs().max().item() < 1e-4
    if (Cin, Cout, k) in ((1, 4, 3), (4, 16, 3), (16, 64, 3)) and H >= 128:
        # batched path (2 x 2 pixels per thread, four times the tile): same taps in the same order -> bit-identical
        Bb = 40
        xb = torch.randn((Bb, H, H, Cin), generator=g, device="cuda")
        xb[:B] = x.permute(0, 2, 3, 1)
        big, _, _ = ops.conv2d_small(xb, w.permute(2, 3, 1, 0).contiguous(), b, Bb, H, H, Cin, Cout, k, s, p,
                                     ln=(lw, lb) if use_ln else None, gelu=use_ln)
        assert torch.equal(big.view(Bb, -1)[:B], out.view(B, -1))


def test_im2col_nhwc_and_dwconv():
    from us_video_medsam2_b200 import ops

    g = _g(6)
    B = 2
    x = torch.randn((B, 64, 64, 64), generator=g, device="cuda")
    A = ops.im2col_nhwc(x.view(-1, 64), B, 64, 64, 64, 3, 2, 1)
    w = torch.randn((256, 64, 3, 3), generator=g, device="cuda") / 24
    want = F.conv2d(x.permute(0, 3, 1, 2).to(BF).float(), w, stride=2, padding=1).permute(0, 2, 3, 1).reshape(-1, 256)
    got = A.float() @ w.permute(0, 2, 3, 1).reshape(256, 576).t()
    assert (got - want).abs().max().item() < 2e-3
    x = torch.randn((B, 256, 32, 32), generator=g, device="cuda")
    dw = torch.randn((256, 1, 7, 7), generator=g, device="cuda") / 7
    db = torch.randn(256, generator=g, device="cuda")
    lw, lb = torch.randn(256, generator=g, device="cuda"), torch.randn(256, generator=g, device="cuda")
    out = ops.dwconv7_ln(x.permute(0, 2, 3, 1).contiguous().view(-1, 256), dw.reshape(256, 49).t().contiguous(), db, lw,
                         lb, B, 32, 32)
    y = F.conv2d(x, dw, db, padding=3, groups=256).permute(0, 2, 3, 1)
    y = F.layer_norm(y, (256,), lw, lb, 1e-6)
    assert (out.float().view_as(y) - y).abs().max().item() < 5e-2
    assert (out.float().view_as(y) - y.to(BF).float()).abs().max().item() < 4e-2
    # batched path (8 x 8 output tiles from 10 images on): same taps in the same order, LayerNorm statistics reduced in a
    # different order -> equal to the row-tile kernel's per-image results up to one bf16 rounding
    B2 = 12
    xb = torch.randn((B2, 32, 32, 256), generator=g, device="cuda")
    args = (dw.reshape(256, 49).t().contiguous(), db, lw, lb)
    big = ops.dwconv7_ln(xb.view(-1, 256), *args, B2, 32, 32).view(B2, -1)
    yb = F.layer_norm(F.conv2d(xb.permute(0, 3, 1, 2), dw, db, padding=3, groups=256).permute(0, 2, 3, 1), (256,), lw, lb, 1e-6)
    assert (big.float().view_as(yb) - yb).abs().max().item() < 5e-2
    for i in (0, 5, 11):
        one = ops.dwconv7_ln(xb[i].reshape(-1, 256).contiguous(), *args, 1, 32, 32).view(-1)
        d = (big[i].float() - one.float()).abs()
        assert d.max().item() <= 2 ** -5 and (d > 0).float().mean().item() < 0.02
    # width not a multiple of the 8-pixel tile: the one-warp-per-pixel kernel
    x = torch.randn((1, 256, 12, 20), generator=g, device="cuda")
    out = ops.dwconv7_ln(x.permute(0, 2, 3, 1).contiguous().view(-1, 256), dw.reshape(256, 49).t().contiguous(), db, lw,
                         lb, 1, 12, 20)
    y = F.layer_norm(F.conv2d(x, dw, db, padding=3, groups=256).permute(0, 2, 3, 1), (256,), lw, lb, 1e-6)
    assert (out.float().view_as(y) - y).abs().max().item() < 5e-2


@pytest.mark.parametrize("Hg,ws,pool,heads", [(64, 8, False, 1), (64, 8, True, 2), (32, 4, False, 2), (32, 4, True, 4),
                                              (16, 7, False, 8), (24, 8, True, 5), (20, 6, False, 1), (32, 14, False, 4),
                                              (32, 14, True, 8), (30, 10, False, 2)])
def test_fused_window_attention(Hg, ws, pool, heads):
    """One-kernel window attention == gather -> flash attention -> scatter (the generic path), incl. partial windows
    (bias-valued padding tokens), pooled queries and more heads than one shared-memory pass holds."""
    from us_video_medsam2_b200 import ops

    Fr, C = 2, heads * 96
    g = _g(Hg * 3 + ws)
    qkv = torch.randn((Fr * Hg * Hg, 3 * C), generator=g, device="cuda").to(BF)
    bias = torch.randn(3 * C, generator=g, device="cuda")
    got = ops.window_attn(qkv, bias, Fr, Hg, Hg, ws, pool, C, heads)
    Qw, Kw, Vw, nw, nq, nk = ops.window_gather(qkv, bias, Fr, Hg, Hg, ws, pool, C)
    Ow = ops.fmha(Qw, Kw, Vw, Fr * nw, heads, nq, nk, 96, (0, nq * C, C, 96), (0, nk * C, C, 96), (0, nk * C, C, 96))
    Ho = Hg // 2 if pool else Hg
    want = ops.window_scatter(Ow, Fr, Ho, Ho, ws // 2 if pool else ws, C)
    # same bf16 inputs, fp32 accumulation, bf16 output: differences are summation order only
    assert (got.float() - want.float()).abs().max().item() < 2e-2
    # and against plain fp32 attention on the gathered windows
    q, k, v = (t.float().view(Fr * nw, -1, heads, 96).transpose(1, 2) for t in (Qw, Kw, Vw))
    ref = (torch.softmax(q @ k.transpose(-1, -2) / 96 ** 0.5, dim=-1) @ v).transpose(1, 2).reshape(Fr * nw, nq, C)
    ref = ops.window_scatter(ref.to(BF).contiguous(), Fr, Ho, Ho, ws // 2 if pool else ws, C)
    assert (got.float() - ref.float()).abs().max().item() < 3e-2


@pytest.mark.parametrize("Hi,Ho", [(128, 512), (128, 360), (512, 128), (100, 37)])
def test_resize_bilinear(Hi, Ho):
    from us_video_medsam2_b200 import ops

    g = _g(Hi)
    x = torch.randn((3, 1, Hi, Hi + 8), generator=g, device="cuda")
    got = ops.resize_bilinear(x, Ho, Ho + 4)
    want = F.interpolate(x, size=(Ho, Ho + 4), mode="bilinear", align_corners=False)
    assert (got - want).abs().max().item() < 1e-5
    got = ops.resize_bilinear(x, Ho, Ho + 4, ops.POST_SIGMOID_AFFINE, 20.0, -10.0)
    assert (got - (torch.sigmoid(want) * 20 - 10)).abs().max().item() < 1e-4


@pytest.mark.parametrize("Hi,Ho", [(512, 128), (720, 512), (300, 512), (128, 128)])
def test_resize_bilinear_antialias(Hi, Ho):
    from us_video_medsam2_b200 import ops

    g = torch.Generator().manual_seed(Hi)
    x = torch.rand((2, 1, Hi, Hi + 16), generator=g)
    want = F.interpolate(x, size=(Ho, Ho), mode="bilinear", align_corners=False, antialias=True)  # CPU = oracle side
    got = ops.resize_bilinear_aa(x.cuda(), Ho, Ho).cpu()
    assert (got - want).abs().max().item() < 1e-5


def test_decoder_tail_kernels():
    from us_video_medsam2_b200 import ops

    g = _g(8)
    B = 2
    # two transposed convs as pixel-shuffle GEMMs + fused epilogues
    src = torch.randn((B, 256, 32, 32), generator=g, device="cuda")
    w1 = torch.randn((256, 64, 2, 2), generator=g, device="cuda") / 16
    b1 = torch.randn(64, generator=g, device="cuda")
    w2 = torch.randn((64, 32, 2, 2), generator=g, device="cuda") / 8
    b2 = torch.randn(32, generator=g, device="cuda")
    lw, lb = torch.randn(64, generator=g, device="cuda"), torch.randn(64, generator=g, device="cuda")
    s1 = torch.randn((1, 64, 64, 64), generator=g, device="cuda")
    s0 = torch.randn((1, 32, 128, 128), generator=g, device="cuda")
    hyper = torch.randn((B, 4, 32), generator=g, device="cuda")
    y = F.conv_transpose2d(src, w1, b1, stride=2) + s1
    u = y.mean(1, keepdim=True)
    v = (y - u).pow(2).mean(1, keepdim=True)
    y = F.gelu((y - u) / torch.sqrt(v + 1e-6) * lw[None, :, None, None] + lb[None, :, None, None])
    y2 = F.gelu(F.conv_transpose2d(y, w2, b2, stride=2) + s0)
    want_masks = (hyper @ y2.view(B, 32, -1)).view(B, 4, 128, 128)
    keys = src.permute(0, 2, 3, 1).reshape(-1, 256).contiguous()
    g1 = ops.gemm_f32(keys, w1.permute(2, 3, 1, 0).reshape(256, 256).contiguous(), b1.repeat(4))
    u1 = ops.upscale1_ln_gelu(g1, s1.permute(0, 2, 3, 1).contiguous(), lw, lb, B, 32, 32, B)  # one frame for all B
    assert (u1.view(B, 64, 64, 64) - y.permute(0, 2, 3, 1)).abs().max().item() < 2e-4
    g2 = ops.gemm_f32(u1, w2.permute(2, 3, 1, 0).reshape(128, 64).contiguous(), b2.repeat(4))
    masks = ops.upscale2_masks(g2, s0.permute(0, 2, 3, 1).contiguous(), hyper.contiguous(), B, 64, 64, B)
    assert (masks - want_masks).abs().max().item() < 5e-4
    # feature groups (several videos in one launch): 4 objects, 2 per video -> object b reads frame b // 2
    B4 = 4
    src4 = torch.randn((B4, 256, 32, 32), generator=g, device="cuda")
    s1v = torch.randn((2, 64, 64, 64), generator=g, device="cuda")
    s0v = torch.randn((2, 32, 128, 128), generator=g, device="cuda")
    hyper4 = torch.randn((B4, 4, 32), generator=g, device="cuda")
    y = F.conv_transpose2d(src4, w1, b1, stride=2) + s1v.repeat_interleave(2, dim=0)
    u = y.mean(1, keepdim=True)
    v = (y - u).pow(2).mean(1, keepdim=True)
    y = F.gelu((y - u) / torch.sqrt(v + 1e-6) * lw[None, :, None, None] + lb[None, :, None, None])
    y2 = F.gelu(F.conv_transpose2d(y, w2, b2, stride=2) + s0v.repeat_interleave(2, dim=0))
    want4 = (hyper4 @ y2.view(B4, 32, -1)).view(B4, 4, 128, 128)
    g1 = ops.gemm_f32(src4.permute(0, 2, 3, 1).reshape(-1, 256).contiguous(),
                      w1.permute(2, 3, 1, 0).reshape(256, 256).contiguous(), b1.repeat(4))
    u1 = ops.upscale1_ln_gelu(g1, s1v.permute(0, 2, 3, 1).contiguous(), lw, lb, B4, 32, 32, 2)
    g2 = ops.gemm_f32(u1, w2.permute(2, 3, 1, 0).reshape(128, 64).contiguous(), b2.repeat(4))
    masks4 = ops.upscale2_masks(g2, s0v.permute(0, 2, 3, 1).contiguous(), hyper4.contiguous(), B4, 64, 64, 2)
    assert (masks4 - want4).abs().max().item() < 5e-4
    # grouped rows of the row-broadcast kernels: axpby x_div, GEMM res_div
    x = torch.randn((2 * 8, 64), generator=g, device="cuda")
    pos = torch.randn((8, 64), generator=g, device="cuda")
    got, _ = ops.axpby(x, pos, 1.0, 0.1, rows=4 * 8, x_mod=8, y_mod=8, x_div=2 * 8)
    want = x.view(2, 1, 8, 64).expand(2, 2, 8, 64).reshape(32, 64) + 0.1 * pos.repeat(4, 1)
    assert (got - want).abs().max().item() < 1e-6  # (the kernel fuses the multiply-add)
    a = torch.randn((4 * 128, 64), generator=g, device="cuda").to(torch.bfloat16)
    w = (torch.randn((256, 64), generator=g, device="cuda") / 8).to(torch.bfloat16)
    res = torch.randn((2 * 128, 256), generator=g, device="cuda")
    o32, _ = ops.gemm_bf16(a, w, residual=res, res_mod=128, res_div=2 * 128, f32=True)
    want = a.float() @ w.float().t() + res.view(2, 1, 128, 256).expand(2, 2, 128, 256).reshape(512, 256)
    assert (o32 - want).abs().max().item() < 2e-3
    # small MLP, instance-stacked + row select
    Ws = [torch.randn((4, 256, 256), generator=g, device="cuda") / 16, torch.randn((4, 256), generator=g, device="cuda"),
          torch.randn((4, 256, 256), generator=g, device="cuda") / 16, torch.randn((4, 256), generator=g, device="cuda"),
          torch.randn((4, 32, 256), generator=g, device="cuda") / 16, torch.randn((4, 32), generator=g, device="cuda")]
    hs = torch.randn((B, 8, 256), generator=g, device="cuda")
    got = ops.small_mlp3(hs.data_ptr() + 4 * 2 * 256, 8 * 256, 256, None, tuple(Ws), 32, B, 4, hs)
    for i in range(4):
        x = hs[:, 2 + i]
        want = F.relu(F.relu(x @ Ws[0][i].t() + Ws[1][i]) @ Ws[2][i].t() + Ws[3][i]) @ Ws[4][i].t() + Ws[5][i]
        assert (got[:, i] - want).abs().max().item() < 1e-4
    idx = torch.tensor([3, 1], dtype=torch.int32, device="cuda")
    one = tuple(t[:1].contiguous() for t in Ws)
    got = ops.small_mlp3(hs.data_ptr() + 4 * 2 * 256, 8 * 256, 256, idx, one, 32, B, 1, hs, sigmoid=True)
    for b in range(B):
        x = hs[b, 2 + int(idx[b])]
        want = torch.sigmoid(F.relu(F.relu(x @ Ws[0][0].t() + Ws[1][0]) @ Ws[2][0].t() + Ws[3][0]) @ Ws[4][0].t() + Ws[5][0])
        assert (got[b, 0] - want).abs().max().item() < 1e-5


def test_sam_select_and_point_embed():
    from oracle.medsam2_ref import random_fourier_pe
    from us_video_medsam2_b200 import ops

    g = _g(9)
    B = 3
    masks = torch.randn((B, 4, 128, 128), generator=g, device="cuda") * 0.07
    masks[1, 0] = masks[1, 0].abs() + 0.2  # stable single mask for object 1
    iou = torch.tensor([[0.9, 0.2, 0.5, 0.4], [0.1, 0.3, 0.3, 0.2], [0.5, 0.1, 0.2, 0.6]], device="cuda")
    score = torch.tensor([[0.4], [0.2], [-0.1]], device="cuda")
    low, idx, sel = ops.sam_select(masks, iou, score, True, 0.05, 0.98, -1024.0)
    assert idx.tolist() == [2, 1, 3] and torch.allclose(sel.flatten(), torch.tensor([0.5, 0.3, 0.6], device="cuda"))
    assert torch.equal(low[0, 0], masks[0, 2]) and torch.equal(low[1, 0], masks[1, 1])
    assert float(low[2].max()) == -1024.0 and float(low[2].min()) == -1024.0
    low, idx, sel = ops.sam_select(masks, iou, score, False, 0.05, 0.98, -1024.0)
    assert idx.tolist() == [0, 0, 0]
    assert torch.equal(low[0, 0], masks[0, 2])  # unstable -> best multimask
    assert torch.equal(low[1, 0], masks[1, 0])  # stable -> single mask
    gauss = torch.randn((2, 128), generator=g, device="cuda")
    table = torch.randn((5, 256), generator=g, device="cuda")
    coords = torch.tensor([[[100.5, 200.5], [30.5, 40.5], [0.0, 0.0]]], device="cuda")
    labels = torch.tensor([[1, 3, -1]], dtype=torch.int32, device="cuda")
    got = ops.point_embed(coords, labels, gauss, table, 512)
    pe = random_fourier_pe(coords.cpu() / 512, gauss.cpu())
    want = torch.stack([pe[0, 0] + table[1].cpu(), pe[0, 1] + table[3].cpu(), table[4].cpu()])
    assert (got[0].cpu() - want).abs().max().item() < 1e-4


def test_skinny_gemm_and_decoder_attention():
    from us_video_medsam2_b200 import ops

    g = _g(21)
    B, Nt = 3, 8
    x = torch.randn((B * Nt, 256), generator=g, device="cuda")
    x2 = torch.randn((B * Nt, 256), generator=g, device="cuda")
    w = torch.randn((2048, 256), generator=g, device="cuda") / 16
    b = torch.randn(2048, generator=g, device="cuda")
    got = ops.gemm_skinny(x, w, b, x2=x2, act=ops.ACT_RELU)
    part = ops.gemm_skinny(x, w, b, x2=x2, x2_cols=8)  # positional add on the first 8 output columns only
    want_part = torch.cat([(x + x2) @ w[:8].t(), x @ w[8:].t()], dim=1) + b
    assert (part - want_part).abs().max().item() < 1e-4
    want = F.relu((x + x2) @ w.t() + b)
    assert (got - want).abs().max().item() < 1e-4
    w2 = torch.randn((256, 2048), generator=g, device="cuda") / 45
    res = torch.randn((B * Nt, 256), generator=g, device="cuda")
    got2 = ops.gemm_skinny(got, w2, None, residual=res)
    assert (got2 - (want @ w2.t() + res)).abs().max().item() < 2e-4
    # stacked instances reading consecutive token rows, odd N
    hs = torch.randn((B, Nt, 256), generator=g, device="cuda")
    W = torch.randn((6, 30, 256), generator=g, device="cuda") / 16
    bb = torch.randn((6, 30), generator=g, device="cuda")
    y = ops.gemm_skinny(None, W, bb, M=B, x_ptr=hs.data_ptr(), x_rs=Nt * 256, x_is=256, instances=6)
    want = torch.einsum("bik,ink->bin", hs[:, :6], W) + bb
    assert (y.view(B, 6, 30) - want).abs().max().item() < 1e-4
    idx = torch.tensor([3, 0, 2], dtype=torch.int32, device="cuda")
    y = ops.gemm_skinny(None, W[0], bb[0], M=B, x_ptr=hs.data_ptr() + 4 * 2 * 256, x_rs=Nt * 256, row_select=idx,
                        x_sel_stride=256)
    want = torch.stack([hs[i, 2 + int(idx[i])] @ W[0].t() + bb[0] for i in range(B)])
    assert (y - want).abs().max().item() < 1e-4
    # LayerNorm on load: same product as LayerNorm followed by the GEMM, and the side output equals the LayerNorm kernel
    lw, lb = torch.randn(256, generator=g, device="cuda"), torch.randn(256, generator=g, device="cuda")
    xn, _ = ops.layernorm(x, lw, lb, 1e-5, f32=True)
    side = torch.empty_like(x)
    got_ln = ops.gemm_skinny(x, w, b, x2=x2, x2_cols=128, act=ops.ACT_RELU, ln=(lw, lb, 1e-5), ln_out=side)
    assert torch.equal(side, xn)
    assert torch.equal(got_ln, ops.gemm_skinny(xn, w, b, x2=x2, x2_cols=128, act=ops.ACT_RELU))
    want_ln = F.layer_norm(x, (256,), lw, lb, 1e-5)
    assert (side - want_ln).abs().max().item() < 1e-5
    hs_n = torch.full_like(hs, float("nan"))
    y = ops.gemm_skinny(None, W, bb, M=B, x_ptr=hs.data_ptr(), x_rs=Nt * 256, x_is=256, instances=6,
                        ln=(lw, lb, 1e-5), ln_out=hs_n, ln_rs=Nt * 256, ln_is=256)
    want_hs = F.layer_norm(hs[:, :6], (256,), lw, lb, 1e-5)
    assert (hs_n[:, :6] - want_hs).abs().max().item() < 1e-5 and bool(torch.isnan(hs_n[:, 6:]).all())
    assert (y.view(B, 6, 30) - (torch.einsum("bik,ink->bin", want_hs, W) + bb)).abs().max().item() < 1e-4
    # token->image and image->token attention (8 heads x 16) on column views of a fused buffer
    for Nt_ in (8, 11):
        q = torch.randn((B * Nt_, 128), generator=g, device="cuda")
        img = torch.randn((B * 1024, 384), generator=g, device="cuda")
        o = ops.attn_t2i(q, img[:, 0:128], img[:, 128:256], B, Nt_, 1024)
        sp = lambda t, n: t.reshape(B, n, 8, 16).permute(0, 2, 1, 3)
        att = torch.softmax(sp(q, Nt_) @ sp(img[:, 0:128], 1024).transpose(-1, -2) / 4.0, dim=-1) @ sp(img[:, 128:256], 1024)
        assert (o - att.permute(0, 2, 1, 3).reshape(B * Nt_, 128)).abs().max().item() < 2e-5
        k2 = torch.randn((B * Nt_, 128), generator=g, device="cuda")
        v2 = torch.randn((B * Nt_, 128), generator=g, device="cuda")
        o = ops.attn_i2t(img[:, 256:384], k2, v2, B, 1024, Nt_)
        att = torch.softmax(sp(img[:, 256:384], 1024) @ sp(k2, Nt_).transpose(-1, -2) / 4.0, dim=-1) @ sp(v2, Nt_)
        assert (o - att.permute(0, 2, 1, 3).reshape(B * 1024, 128)).abs().max().item() < 2e-5


def test_frame_store_kernels():
    from oracle.medsam2_ref import sine_pos_1d
    from us_video_medsam2_b200 import ops

    g = _g(22)
    B, T, Cm, S = 2, 1024, 64, 12
    store = ops.FrameStore(S, B, torch.device("cuda"))
    store.mem.copy_(torch.randn((S, B, T, Cm), generator=g, device="cuda"))
    store.ptr.copy_(torch.randn((S, B, 256), generator=g, device="cuda"))
    pos = torch.randn((T, Cm), generator=g, device="cuda")
    tpos = torch.randn((7, Cm), generator=g, device="cuda")
    W = torch.randn((64, 256), generator=g, device="cuda") / 16
    bias = torch.randn(64, generator=g, device="cuda")
    ctrl = ops.new_frame_ctrl(torch.device("cuda"))
    mem_frames, mem_tpos = [0, 9, 10, 11], [6, 2, 1, 0]
    ptr_frames, ptr_rel = [0, 11, 10, 9, 8], [11 / 15, 1 / 15, 2 / 15, 3 / 15, 4 / 15]
    want_pp = (sine_pos_1d(torch.tensor(ptr_rel), 256).cuda() @ W.t() + bias).repeat_interleave(4, dim=0)
    for obj0, Bq in ((0, 2), (1, 1)):
        ops.set_frame_ctrl(ctrl, store, obj0, 5, mem_frames, mem_tpos, ptr_frames, ptr_rel)
        pp = ops.ptr_tpos(ctrl, W, bias, len(ptr_frames))
        assert (pp - want_pp).abs().max().item() < 1e-4
        k_in, v_in, Nk = ops.build_memory_store(ctrl, pos, tpos, pp, Bq, 4, 5)
        sl = slice(obj0, obj0 + Bq)
        pt = torch.stack([store.ptr[f, sl] for f in ptr_frames], 1).reshape(Bq, 20, 64)
        wk = torch.cat([store.mem[f, sl].float() + (pos + tpos[r]) for f, r in zip(mem_frames, mem_tpos)] + [pt + pp], dim=1)
        wv = torch.cat([store.mem[f, sl].float() for f in mem_frames] + [pt], dim=1)
        assert Nk == 4 * T + 20 and torch.equal(k_in, wk.to(BF)) and torch.equal(v_in, wv.to(BF))
    ops.set_frame_ctrl(ctrl, store, 0, 5, mem_frames, mem_tpos, ptr_frames, ptr_rel)
    x = torch.randn((B * T, Cm), generator=g, device="cuda")
    score = torch.tensor([[0.3], [-0.2]], device="cuda")
    emb = torch.randn(Cm, generator=g, device="cuda")
    before = store.mem.clone()
    assert ops.finalize_memory(x, score, emb, B, ctrl=ctrl) is None
    want = x.view(B, T, Cm).clone()
    want[1] += emb
    assert torch.equal(store.mem[5], want.to(BF)) and torch.equal(store.mem[4], before[4]) and torch.equal(store.mem[6], before[6])
    a, c = torch.randn((B, 256), generator=g, device="cuda"), torch.randn((B, 1, 128, 128), generator=g, device="cuda")
    ops.store_outputs(ctrl, a, score, c)
    assert torch.equal(store.ptr[5], a) and torch.equal(store.score[5], score) and torch.equal(store.masks[5], c)
    assert float(store.masks[4].abs().max()) == 0


def test_im2col_patch_grid_matches_relayout():
    """ViT patch embedding im2col (kernel = stride = 16): equals the pure re-layout of the image."""
    from us_video_medsam2_b200 import ops

    img = torch.randn((3, 3, 64, 64), generator=_g(5), device="cuda")
    got = ops.im2col_patch_grid(img, 16)
    want = img.view(3, 3, 4, 16, 4, 16).permute(0, 2, 4, 1, 3, 5).reshape(3 * 16, 768).to(torch.bfloat16)
    assert torch.equal(got, want)


def test_gelu_epilogue_is_fp32_accurate():
    """The GELU of every fused epilogue (common.cuh: erfc by Abramowitz-Stegun 7.1.26 on the FMA pipe + rcp / ex2) against
    the exact erf form in float64 over [-12, 12]: within fp32 rounding of nn.GELU (torch's own fp32 kernel errs by 1.2e-6)."""
    from us_video_medsam2_b200 import ops

    n = 64 * 32768
    x = torch.linspace(-12.0, 12.0, n, device="cuda", dtype=torch.float64).float().view(-1, 64).contiguous()
    eye = torch.eye(64, device="cuda")
    got = ops.gemm_f32(x, eye, act=ops.ACT_GELU).double()
    want = F.gelu(x.double())
    err = (got - want).abs()
    assert err.max().item() < 1e-6, err.max().item()
    # relative accuracy: fp32-level where the value is O(1e-2) or more; the polynomial's 1.5e-7 ABSOLUTE error shows as
    # a relative one only in the far negative tail (|gelu| ~ 1e-4 at x ~ -4), where it stays below a quarter of a bf16 ulp
    m = want.abs() > 1e-2
    assert (err[m] / want[m].abs()).max().item() < 2e-5
    m = want.abs() > 1e-4
    assert (err[m] / want[m].abs()).max().item() < 1e-3
    assert (got - F.gelu(x).double()).abs().max().item() < 2.5e-6  # vs torch's fp32 kernel


def test_tcgen05_gemm_gelu_epilogue_is_fp32_accurate():
    """The GELU of the tcgen05 GEMM epilogues against the exact erf form in float64: an identity GEMM over every bf16 value
    of [-12, 12] (bf16 operands are exact), both through the one-tile kernel and the persistent one."""
    from us_video_medsam2_b200 import ops

    bits = torch.arange(0, 65536, dtype=torch.int32)
    vals = bits.to(torch.int16).view(torch.bfloat16).float()
    vals = vals[torch.isfinite(vals) & (vals.abs() <= 12.0)]
    n = (vals.numel() // 64) * 64
    x = vals[:n].view(-1, 64).to(torch.bfloat16).cuda().contiguous()
    eye = torch.eye(64, device="cuda", dtype=torch.bfloat16)
    want = F.gelu(x.double())
    for reps, block_n in ((1, 0), (64, -1)):  # the second case is large enough for the persistent kernel
        xs = x.repeat(reps, 1)
        got, _ = ops.gemm_bf16(xs, eye, act=ops.ACT_GELU, f32=True, block_n=block_n)
        err = (got[: x.shape[0]].double() - want).abs()
        assert err.max().item() < 1e-6, (block_n, err.max().item())
        m = want.abs() > 1e-2
        assert (err[m] / want[m].abs()).max().item() < 2e-5
        assert torch.equal(got[: x.shape[0]], got[-x.shape[0]:])


@pytest.mark.parametrize("B,binarize", [(2, False), (3, True), (40, False)])
def test_mask_first_conv_equals_resize_then_conv(B, binarize):
    """usvm_conv2d_mask_first (bilinear x4 + sigmoid / binarise evaluated inside the footprint load) against
    resize_bilinear followed by the ordinary first down-sampler stage (both the latency and the batched tile kernels):
    identical with the binarised input, within fp32 rounding of the interpolation (FMA contraction differs between the two
    kernels) with the sigmoid."""
    from us_video_medsam2_b200 import ops

    g = _g(B + 77)
    low = torch.randn((B, 1, 128, 128), generator=g, device="cuda") * 3
    w = torch.randn((3, 3, 1, 4), generator=g, device="cuda") / 3
    b, lw, lb = (torch.randn(4, generator=g, device="cuda") for _ in range(3))
    post = ops.POST_BINARIZE_AFFINE if binarize else ops.POST_SIGMOID_AFFINE
    full = ops.resize_bilinear(low, 512, 512, post, 20.0, -10.0)
    want, Ho, Wo = ops.conv2d_small(full.reshape(B, 512, 512, 1), w, b, B, 512, 512, 1, 4, 3, 2, 1, ln=(lw, lb), gelu=True)
    got, Ho2, Wo2 = ops.conv2d_mask_first(low, post, 20.0, -10.0, w, b, B, 512, 512, 3, 2, 1, ln=(lw, lb), gelu=True)
    assert (Ho, Wo) == (Ho2, Wo2) == (256, 256)
    if binarize:
        assert torch.equal(got, want)
    else:
        assert (got - want).abs().max().item() < 2e-5
