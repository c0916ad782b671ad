"""GEMM kernels vs a plain PyTorch fp32 reference of the same op (bf16-rounded operands, fp32 accumulate)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _ref(a, w, bias, act, scale, res, res_mod):
    y = a.float() @ w.float().t()
    if bias is not None:
        y = y + bias
    if act == 1:
        y = F.relu(y)
    elif act == 2:
        y = F.gelu(y)
    if scale is not None:
        y = y * scale
    if res is not None:
        idx = torch.arange(y.shape[0], device=y.device) % res_mod if res_mod else torch.arange(y.shape[0], device=y.device)
        y = y + res[idx]
    return y


SHAPES = [  # (M, N, K) -- hot-path shapes + ragged tails
    (128, 64, 64), (256, 96, 96), (16384, 288, 96), (4096, 768, 192), (1024, 2048, 256), (1024, 256, 2048),
    (7232, 256, 64), (1000, 100 * 1 + 4, 72), (130, 40, 200), (1024, 1152, 384), (16384, 96, 160), (256, 256, 768),
]


@pytest.mark.parametrize("M,N,K", SHAPES)
@pytest.mark.parametrize("block_n", [0, 32, 128])
def test_tc5_gemm_matches_fp32_reference(M, N, K, block_n):
    from us_video_medsam2_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    a = (torch.randn((M, K), generator=g, device="cuda")).to(torch.bfloat16)
    w = (torch.randn((N, K), generator=g, device="cuda") / K ** 0.5).to(torch.bfloat16)
    o32, _ = ops.gemm_bf16(a, w, f32=True, block_n=block_n, simt=False)
    torch.cuda.synchronize()
    want = _ref(a, w, None, 0, None, None, 0)
    err = (o32 - want).abs().max().item()
    assert err < 2e-3 * max(1.0, want.abs().max().item()), err


@pytest.mark.parametrize("act", [0, 1, 2])
def test_tc5_gemm_fused_epilogue(act):
    from us_video_medsam2_b200 import ops

    M, N, K = 2048, 256, 1024
    g = torch.Generator(device="cuda").manual_seed(act)
    a = torch.randn((M, K), generator=g, device="cuda").to(torch.bfloat16)
    w = (torch.randn((N, K), generator=g, device="cuda") / K ** 0.5).to(torch.bfloat16)
    bias = torch.randn(N, generator=g, device="cuda")
    scale = torch.rand(N, generator=g, device="cuda")
    res = torch.randn((1024, N), generator=g, device="cuda")
    o32, o16 = ops.gemm_bf16(a, w, bias=bias, act=act, col_scale=scale, residual=res, res_mod=1024, f32=True,
                             bf16=True, simt=False)
    want = _ref(a, w, bias, act, scale, res, 1024)
    assert (o32 - want).abs().max().item() < 3e-3
    assert (o16.float() - want).abs().max().item() < 3e-2
    # in-place residual (x = x + f(x)) as used by every transformer block
    x = torch.randn((M, N), generator=g, device="cuda")
    want = _ref(a, w, bias, 0, None, x.clone(), 0)
    ops.gemm_bf16(a, w, bias=bias, residual=x, out_f32=x, simt=False)
    assert (x - want).abs().max().item() < 3e-3


PERSISTENT_SHAPES = [  # (M, N, K, block_n): block_n < 0 selects the persistent kernel, -1 = its own tile choice
    (16384, 288, 96, -1), (8192, 1152, 384, -1), (8192, 1536, 384, -256), (8192, 384, 1536, -1), (7232, 1024, 64, -1),
    (2048, 3072, 768, -1), (131072, 96, 384, -1), (5000, 200, 72, -1), (300, 104, 200, -64), (40000, 288, 96, -96),
    (8192, 1152, 384, -192), (8192, 1152, 384, -32),
]


@pytest.mark.parametrize("M,N,K,block_n", PERSISTENT_SHAPES)
def test_tc5_persistent_gemm_matches_fp32_reference(M, N, K, block_n):
    """Persistent double-buffered-accumulator kernel: many tiles per CTA, ragged M / N / K edges, every tile width."""
    from us_video_medsam2_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(M + N * 3 + K * 5)
    a = (torch.randn((M, K), generator=g, device="cuda")).to(torch.bfloat16)
    w = (torch.randn((N, K), generator=g, device="cuda") / K ** 0.5).to(torch.bfloat16)
    o32, o16 = ops.gemm_bf16(a, w, f32=True, bf16=True, block_n=block_n, simt=False)
    torch.cuda.synchronize()
    want = _ref(a, w, None, 0, None, None, 0)
    tol = 2e-3 * max(1.0, want.abs().max().item())
    assert (o32 - want).abs().max().item() < tol
    assert (o16.float() - want).abs().max().item() < 16 * tol
    # the one-tile-per-CTA kernel accumulates the same bf16 products in the same order: bit-identical results
    p32, _ = ops.gemm_bf16(a, w, f32=True, block_n=128, simt=False)
    assert torch.equal(o32, p32)


@pytest.mark.parametrize("act", [0, 1, 2])
def test_tc5_persistent_gemm_fused_epilogue(act):
    from us_video_medsam2_b200 import ops

    M, N, K = 8192 + 64, 384, 256
    g = torch.Generator(device="cuda").manual_seed(100 + act)
    a = torch.randn((M, K), generator=g, device="cuda").to(torch.bfloat16)
    w = (torch.randn((N, K), generator=g, device="cuda") / K ** 0.5).to(torch.bfloat16)
    bias = torch.randn(N, generator=g, device="cuda")
    scale = torch.rand(N, generator=g, device="cuda")
    res = torch.randn((1024, N), generator=g, device="cuda")
    o32, o16 = ops.gemm_bf16(a, w, bias=bias, act=act, col_scale=scale, residual=res, res_mod=1024, f32=True,
                             bf16=True, simt=False, block_n=-1)
    want = _ref(a, w, bias, act, scale, res, 1024)
    assert (o32 - want).abs().max().item() < 3e-3
    assert (o16.float() - want).abs().max().item() < 3e-2
    x = torch.randn((M, N), generator=g, device="cuda")
    want = _ref(a, w, bias, 0, None, x.clone(), 0)
    ops.gemm_bf16(a, w, bias=bias, residual=x, out_f32=x, simt=False, block_n=-1)
    assert (x - want).abs().max().item() < 3e-3


@pytest.mark.parametrize("M,N,K", [(8, 256, 256), (1024, 128, 256), (4096, 128, 64), (37, 19, 53)])
def test_simt_gemm_fp32(M, N, K):
    from us_video_medsam2_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(5)
    a = torch.randn((M, K), generator=g, device="cuda")
    w = torch.randn((N, K), generator=g, device="cuda") / K ** 0.5
    bias = torch.randn(N, generator=g, device="cuda")
    res = torch.randn((M, N), generator=g, device="cuda")
    got = ops.gemm_f32(a, w, bias, act=1, residual=res)
    want = F.relu(a @ w.t() + bias) + res
    assert (got - want).abs().max().item() < 1e-4


def test_simt_and_tc5_agree_on_bf16_operands():
    from us_video_medsam2_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(11)
    a = torch.randn((1024, 384), generator=g, device="cuda").to(torch.bfloat16)
    w = (torch.randn((1536, 384), generator=g, device="cuda") / 20).to(torch.bfloat16)
    x, _ = ops.gemm_bf16(a, w, f32=True, simt=False)
    y, _ = ops.gemm_bf16(a, w, f32=True, simt=True)
    assert (x - y).abs().max().item() < 1e-3


@pytest.mark.parametrize("M,N,K", [(1024, 384, 256), (1024, 256, 128), (4096, 128, 64), (1000, 100, 72), (130, 40, 36),
                                   (32768, 256, 256), (32768 + 40, 128, 72)])  # the last two: persistent tf32 kernel
@pytest.mark.parametrize("block_n", [0, 32, 128])
def test_tc5_tf32_gemm(M, N, K, block_n):
    """fp32 operands on the tf32 tensor-core path: within tf32 rounding (10-bit mantissa products) of the fp32 result,
    and exact when the operands are representable in tf32."""
    from us_video_medsam2_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = torch.randn((M, K), generator=g, device="cuda")
    w = torch.randn((N, K), generator=g, device="cuda") / K ** 0.5
    bias = torch.randn(N, generator=g, device="cuda")
    res = torch.randn((M, N), generator=g, device="cuda")
    got = ops.gemm_f32(a, w, bias, act=1, residual=res, tf32=True, block_n=block_n)
    want = F.relu(a @ w.t() + bias) + res
    assert (got - want).abs().max().item() < 8e-3
    # operands that are exact in tf32 (bf16-representable): the only error left is fp32 accumulation order
    a16, w16 = a.to(torch.bfloat16).float(), w.to(torch.bfloat16).float()
    got = ops.gemm_f32(a16, w16, bias, residual=res, tf32=True, block_n=block_n)
    want = a16 @ w16.t() + bias + res
    assert (got - want).abs().max().item() < 2e-5 * K ** 0.5


@pytest.mark.parametrize("M,K,gelu", [(1024, 256, False), (2048 + 40, 576, True), (128, 2048, False)])
def test_tc5_gemm_fused_layernorm(M, K, gelu):
    """Residual GEMM + LayerNorm of the result row (N = 256) in one launch: fp32 output = epilogue result,
    bf16 output = LayerNorm (then GELU) of it."""
    from us_video_medsam2_b200 import ops

    N = 256
    g = torch.Generator(device="cuda").manual_seed(M + K)
    a = torch.randn((M, K), generator=g, device="cuda").to(torch.bfloat16)
    w = (torch.randn((N, K), generator=g, device="cuda") / K ** 0.5).to(torch.bfloat16)
    bias = torch.randn(N, generator=g, device="cuda")
    res = torch.randn((M, N), generator=g, device="cuda") * 2 + 0.5
    lw, lb = torch.randn(N, generator=g, device="cuda"), torch.randn(N, generator=g, device="cuda")
    x, h = ops.gemm_bf16(a, w, bias=bias, residual=res, f32=True, ln=(lw, lb, 1e-5, gelu), simt=False, ln_fused=True)
    want_x = _ref(a, w, bias, 0, None, res, 0)
    want_h = F.layer_norm(want_x, (N,), lw, lb, 1e-5)
    if gelu:
        want_h = F.gelu(want_h)
    assert (x - want_x).abs().max().item() < 3e-3
    assert (h.float() - want_h).abs().max().item() < 4e-2
    # agrees with the unfused pair of kernels to bf16 rounding
    x2, _ = ops.gemm_bf16(a, w, bias=bias, residual=res, f32=True, simt=False)
    _, h2 = ops.layernorm(x2, lw, lb, 1e-5, bf16=True, gelu=gelu)
    assert (x - x2).abs().max().item() < 1e-4  # (the unfused GEMM may split K: different fp32 summation order)
    assert (h.float() - h2.float()).abs().max().item() < 4e-2


@pytest.mark.parametrize("M,N,K", [(1024, 256, 2048), (1024, 256, 1024), (1000, 64, 4096), (256, 96, 1536), (128, 256, 2048)])
def test_tc5_gemm_cluster_split_k(M, N, K):
    """Long reductions on few tiles run as 2 or 4 k-slices per output tile (one thread-block cluster per tile, partial
    accumulators through distributed shared memory): same result as the fp32 reference, with the full epilogue."""
    from us_video_medsam2_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(K + N)
    a = torch.randn((M, K), generator=g, device="cuda").to(torch.bfloat16)
    w = (torch.randn((N, K), generator=g, device="cuda") / K ** 0.5).to(torch.bfloat16)
    bias = torch.randn(N, generator=g, device="cuda")
    res = torch.randn((M, N), generator=g, device="cuda")
    o32, o16 = ops.gemm_bf16(a, w, bias=bias, act=1, residual=res, f32=True, bf16=True, simt=False)
    want = _ref(a, w, bias, 1, None, res, 0)
    assert (o32 - want).abs().max().item() < 3e-3
    assert (o16.float() - want).abs().max().item() < 3e-2
    again, _ = ops.gemm_bf16(a, w, bias=bias, act=1, residual=res, f32=True, simt=False)
    assert torch.equal(o32, again)  # slices are added in a fixed order


@pytest.mark.parametrize("M,rows_per_batch,n_rope,N,cols,K", [
    (1024, 1024, 1024, 768, 512, 256),          # one-object qkv projection (one-tile kernel, prefetched tables)
    (2 * 7232, 7232, 7168, 1024, 1024, 64),     # bank key projection, 2 objects: pointer rows unrotated
    (3 * 7180, 7180, 7168, 1024, 1024, 64),     # 3 pointers: rows per object not a multiple of 32 (lanes straddle objects)
    (32768, 1024, 1024, 768, 512, 256),         # 32 objects: persistent kernel, tables fetched one block ahead
])
def test_fused_rope_epilogue_matches_rope_kernel(M, rows_per_batch, n_rope, N, cols, K):
    """RoPE fused into the GEMM epilogue (tiled tables, ops.tile_rope_table) against the stand-alone rotation kernel
    applied to the un-rotated fp32 GEMM output (row-major tables), both kernels of usvm_gemm_bf16_tc5."""
    from us_video_medsam2_b200 import ops
    from us_video_medsam2_b200.engine import _rope_tables

    g = torch.Generator(device="cuda").manual_seed(M + N)
    a = torch.randn((M, K), generator=g, device="cuda").to(torch.bfloat16)
    w = (torch.randn((N, K), generator=g, device="cuda") * K ** -0.5).to(torch.bfloat16)
    bias = torch.randn((N,), generator=g, device="cuda")
    c, s = _rope_tables(256, 32, 32)
    c, s = c.cuda(), s.cuda()
    _, got = ops.gemm_bf16(a, w, bias=bias, bf16=True,
                           rope=(ops.tile_rope_table(c), ops.tile_rope_table(s), cols, rows_per_batch, n_rope))
    plain, _ = ops.gemm_bf16(a, w, bias=bias, f32=True)
    want = plain.to(torch.bfloat16).clone()
    for c0 in range(0, cols, 256):
        want[:, c0:c0 + 256] = ops.rope(plain, c0, c, s, rows_per_batch, n_rope)
    # both round (nearly) the same fp32 value to bf16: they may land one bf16 ulp (2^-7 relative) apart, not more
    d = (got.float() - want.float()).abs()
    assert bool((d <= 2.0 ** -7 * want.float().abs().clamp_min(1.0)).all()), d.max().item()
    assert (d > 0).float().mean().item() < 0.02  # ... and only where the fp32 value sits on a rounding boundary
    assert torch.equal(got[:, cols:], plain[:, cols:].to(torch.bfloat16))  # columns beyond rope_cols untouched


@pytest.mark.parametrize("M", [128, 1024, 4096])
def test_ffn_fused_matches_two_gemms_and_fp32(M):
    """usvm_ffn_fused_tc5 (cluster of 8 CTAs per row tile, hidden activations in shared memory, DSMEM reduce-scatter) vs
    the fp32 composition on the same bf16 operands (hidden rounded to bf16 like the kernel does) and vs the two-launch
    path it replaces; two runs are bit-identical (fixed summation order)."""
    from us_video_medsam2_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(M)
    x = torch.randn((M, 256), generator=g, device="cuda")
    h = torch.randn((M, 256), generator=g, device="cuda").to(torch.bfloat16)
    w1 = (torch.randn((2048, 256), generator=g, device="cuda") * 256 ** -0.5).to(torch.bfloat16)
    w2 = (torch.randn((256, 2048), generator=g, device="cuda") * 2048 ** -0.5).to(torch.bfloat16)
    b1 = torch.randn((2048,), generator=g, device="cuda") * 0.1
    b2 = torch.randn((256,), generator=g, device="cuda") * 0.1
    out = ops.ffn_fused(h, x, w1, b1, w2, b2)
    torch.cuda.synchronize()
    hid = torch.relu(h.float() @ w1.float().t() + b1).to(torch.bfloat16).float()
    want = x + hid @ w2.float().t() + b2
    assert (out - want).abs().max().item() < 2e-3
    _, m = ops.gemm_bf16(h, w1, bias=b1, act=ops.ACT_RELU, bf16=True)
    two, _ = ops.gemm_bf16(m, w2, bias=b2, residual=x, f32=True)
    assert (out - two).abs().max().item() < 1e-3
    assert torch.equal(out, ops.ffn_fused(h, x, w1, b1, w2, b2))
