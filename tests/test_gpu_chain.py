"""Token-chain cluster kernel (csrc/token_chain.cu) vs plain PyTorch fp32 references of the same steps."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rand(g, *shape, scale=1.0):
    return torch.randn(shape, generator=g, device="cuda") * scale


# precise: 3 x tf32 products on (hi, lo) splits -> fp32-level agreement; otherwise one round-to-nearest tf32 product per
# pair (10-bit mantissa), i.e. ~5e-4 relative per product before accumulation
TOL = {True: 1.0, False: 40.0}


@pytest.mark.parametrize("precise", [True, False])
@pytest.mark.parametrize("cluster", [8, 16])
@pytest.mark.parametrize("B", [1, 3])
def test_chain_linear_ln_pe_residual(cluster, B, precise):
    """Three dependent LINEAR steps (LayerNorm on load + stored, positional tokens on the first 128 of 384 output
    columns, ReLU, K = 2048 reduction, residual) in one launch."""
    from us_video_medsam2_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(10 * cluster + B)
    x, pe = _rand(g, B * 8, 256), _rand(g, B * 8, 256)
    lw, lb = _rand(g, 256), _rand(g, 256)
    W1, b1 = _rand(g, 384, 256, scale=1 / 16), _rand(g, 384)
    W2, b2 = _rand(g, 2048, 384, scale=1 / 20), _rand(g, 2048)
    W3, b3 = _rand(g, 256, 2048, scale=1 / 45), _rand(g, 256)
    y1, y2, y3, xn = (torch.empty((B * 8, n), device="cuda") for n in (384, 2048, 256, 256))
    ops.token_chain([ops.chain_linear(x, W1, b1, y1, ln=(lw, lb), ln_out=xn, x2=pe, x2_cols=128),
                     ops.chain_linear(y1, W2, b2, y2, act=ops.ACT_RELU),
                     ops.chain_linear(y2, W3, b3, y3, residual=xn)], B, x, cluster=cluster, precise=precise)
    torch.cuda.synchronize()
    n = F.layer_norm(x, (256,), lw, lb, 1e-5)
    r1 = torch.cat([(n + pe) @ W1[:128].t(), n @ W1[128:].t()], dim=1) + b1
    r2 = F.relu(r1 @ W2.t() + b2)
    r3 = r2 @ W3.t() + b3 + n
    f = TOL[precise]
    assert (xn - n).abs().max().item() < 1e-5
    assert (y1 - r1).abs().max().item() < 1e-4 * f
    assert (y2 - r2).abs().max().item() < 2e-4 * f
    assert (y3 - r3).abs().max().item() < 5e-4 * f


@pytest.mark.parametrize("precise", [True, False])
@pytest.mark.parametrize("cluster", [8, 16])
def test_chain_self_attention_and_t2i(cluster, precise):
    """Self-attention input transform and the split-key token->image attention (partials + merge) against torch."""
    from us_video_medsam2_b200 import ops

    B, Nk = 2, 1024
    g = torch.Generator(device="cuda").manual_seed(cluster)
    qkv = _rand(g, B * 8, 1024)  # q at 128, k at 384, v at 768 (the packed layer-boundary layout)
    Wo, bo = _rand(g, 256, 256, scale=1 / 16), _rand(g, 256)
    res = _rand(g, B * 8, 256)
    out = torch.empty((B * 8, 256), device="cuda")
    ops.token_chain([ops.chain_linear(qkv, Wo, bo, out, in_kind=1, attn_cols=(128, 384, 768), residual=res)], B, qkv,
                    cluster=cluster, precise=precise)
    q, k, v = (qkv[:, o:o + 256].view(B, 8, 8, 32).transpose(1, 2) for o in (128, 384, 768))
    att = torch.softmax(q @ k.transpose(-1, -2) / 32 ** 0.5, dim=-1) @ v
    want = att.transpose(1, 2).reshape(B * 8, 256) @ Wo.t() + bo + res
    assert (out - want).abs().max().item() < 1e-4 * TOL[precise]

    img = _rand(g, B * Nk, 384)
    qt = _rand(g, B * 8, 384)
    W2, b2 = _rand(g, 256, 128, scale=1 / 11), _rand(g, 256)
    out2 = torch.empty((B * 8, 256), device="cuda")
    ops.token_chain([ops.chain_t2i(qt, img[:, 0:128], img[:, 128:256], Nk=Nk, q_off=128),
                     ops.chain_linear(qt, W2, b2, out2, in_kind=2)], B, qt, cluster=cluster, precise=precise)
    q = qt[:, 128:256].view(B, 8, 8, 16).transpose(1, 2)
    k = img[:, 0:128].view(B, Nk, 8, 16).transpose(1, 2)
    v = img[:, 128:256].view(B, Nk, 8, 16).transpose(1, 2)
    att = torch.softmax(q @ k.transpose(-1, -2) / 4.0, dim=-1) @ v
    want = att.transpose(1, 2).reshape(B * 8, 128) @ W2.t() + b2
    assert (out2 - want).abs().max().item() < 1e-4 * TOL[precise]


@pytest.mark.parametrize("precise", [True, False])
def test_chain_stacked_heads_and_row_select(precise):
    from us_video_medsam2_b200 import ops

    B = 3
    g = torch.Generator(device="cuda").manual_seed(3)
    hs = _rand(g, B * 8, 256)
    W1, b1 = _rand(g, 6, 256, 256, scale=1 / 16), _rand(g, 6, 256)
    W3, b3 = _rand(g, 6, 32, 256, scale=1 / 16), _rand(g, 6, 32)
    h1, y = torch.empty((B * 6, 256), device="cuda"), torch.empty((B, 192), device="cuda")
    ops.token_chain([ops.chain_linear(hs, W1, b1, h1, rows=6, x_os=8 * 256, stacked=True, act=ops.ACT_RELU),
                     ops.chain_linear(h1, W3, b3, y, rows=6, stacked=True, o_os=192, o_rs=32)], B, hs, precise=precise)
    x = hs.view(B, 8, 256)[:, :6]
    r1 = F.relu(torch.einsum("bik,ink->bin", x, W1) + b1)
    r2 = torch.einsum("bik,ink->bin", r1, W3) + b3
    assert (h1.view(B, 6, 256) - r1).abs().max().item() < 1e-4 * TOL[precise]
    assert (y.view(B, 6, 32) - r2).abs().max().item() < 1e-4 * TOL[precise]

    idx = torch.tensor([3, 0, 2], dtype=torch.int32, device="cuda")
    P, pb = _rand(g, 256, 256, scale=1 / 16), _rand(g, 256)
    t = torch.empty((B, 256), device="cuda")
    ops.token_chain([ops.chain_linear(hs, P, pb, t, rows=1, x_os=8 * 256, x_off=2 * 256, row_select=idx,
                                      sel_stride=256)], B, hs, precise=precise)
    want = torch.stack([hs.view(B, 8, 256)[b, 2 + int(idx[b])] for b in range(B)]) @ P.t() + pb
    assert (t - want).abs().max().item() < 1e-4 * TOL[precise]
