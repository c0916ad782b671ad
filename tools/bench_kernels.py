#!/usr/bin/env python
"""Micro-benchmarks of individual kernels at their propagation-path shapes (CUDA events on the launching stream,
L2-warm like in the real frame where producers and consumers are back to back).  Prints one line per case."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from us_video_medsam2_b200 import ops  # noqa: E402

dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(0)


def timeit(fn, iters=20, warm=3, replays=5):
    """GPU time per call: `iters` calls are captured in one CUDA graph (no host launch overhead between kernels, as on
    the real steady-state frame) and the graph is replayed `replays` times between two events."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for _ in range(iters):
            fn()
    graph.replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(replays):
        graph.replay()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / (iters * replays) * 1e3  # us


def rnd(*shape, dtype=torch.float32, scale=1.0):
    return (torch.randn(shape, generator=g, device=dev) * scale).to(dtype)


def attention():
    T, D = 1024, 256
    for B in (1, 4):
        for Nk in (1024, 7232):
            q = rnd(B * T, D, dtype=torch.bfloat16)
            kv = rnd(B * Nk, 4 * D, dtype=torch.bfloat16)
            args = (B, 1, T, Nk, D, (0, T * D, D, D), (D, Nk * 4 * D, 4 * D, D), (2 * D, Nk * 4 * D, 4 * D, D))
            flops = 4.0 * B * T * Nk * D
            for impl in ("tc5", "mma"):
                for splits in (1, 2, 4, 8, 9, 16, 18):
                    if splits > (Nk + 63) // 64:
                        continue
                    us = timeit(lambda: ops.fmha(q, kv, kv, *args, num_splits=splits, impl=impl))
                    print(f"fmha {impl} B={B} Nk={Nk} splits={splits:2d}: {us:8.1f} us  {flops / us / 1e6:7.1f} TFLOP/s", flush=True)


def gemms():
    for (M, N, K) in [(1024, 256, 256), (1024, 768, 256), (1024, 2048, 256), (1024, 256, 2048), (7232, 1024, 64),
                      (1024, 1024, 256), (1024, 256, 1024), (1024, 64, 256), (131072, 288, 96), (131072, 384, 96),
                      (131072, 96, 384), (32768, 768, 192), (8192, 1536, 384), (8192, 384, 1536), (8192, 1152, 384)]:
        a = rnd(M, K, dtype=torch.bfloat16)
        w = rnd(N, K, dtype=torch.bfloat16, scale=K ** -0.5)
        for bn in (0, 32, 64, 128, 256, -1, -128, -256):
            if abs(bn) > max(32, N) or (bn < 0 and M < 2048):
                continue
            us = timeit(lambda: ops.gemm_bf16(a, w, f32=True, block_n=bn), iters=10)
            print(f"gemm_tc5 M={M} N={N} K={K} bn={bn:3d}: {us:8.1f} us  {2.0 * M * N * K / us / 1e6:7.1f} TFLOP/s", flush=True)
    # the encoder's real epilogues: bias + exact GELU -> bf16 (mlp1), bias + fp32 residual -> fp32 (mlp2 / proj)
    for (M, N, K) in [(131072, 384, 96), (32768, 768, 192), (8192, 1536, 384), (2048, 3072, 768)]:
        a = rnd(M, K, dtype=torch.bfloat16)
        w = rnd(N, K, dtype=torch.bfloat16, scale=K ** -0.5)
        w2 = rnd(K, N, dtype=torch.bfloat16, scale=N ** -0.5)
        b1, b2, x = rnd(N), rnd(K), rnd(M, K)
        for bn in (128, 256, -1):
            us = timeit(lambda: ops.gemm_bf16(a, w, bias=b1, act=2, bf16=True, block_n=bn), iters=10)
            print(f"gemm_tc5 mlp1 gelu->bf16 M={M} N={N} K={K} bn={bn:3d}: {us:8.1f} us  {2.0 * M * N * K / us / 1e6:7.1f} TFLOP/s", flush=True)
        h = rnd(M, N, dtype=torch.bfloat16)
        for bn in (128, 256, -1):
            if abs(bn) > K and bn > 0:
                continue
            us = timeit(lambda: ops.gemm_bf16(h, w2, bias=b2, residual=x, out_f32=x, block_n=bn), iters=10)
            print(f"gemm_tc5 mlp2 +res->f32 M={M} N={K} K={N} bn={bn:3d}: {us:8.1f} us  {2.0 * M * N * K / us / 1e6:7.1f} TFLOP/s", flush=True)
    for (M, N, K) in [(1024, 384, 256), (1024, 256, 128), (1024, 256, 256), (4096, 128, 64)]:
        a, w = rnd(M, K), rnd(N, K, scale=K ** -0.5)
        us = timeit(lambda: ops.gemm_f32(a, w))
        print(f"gemm_f32 M={M} N={N} K={K}: {us:8.1f} us  {2.0 * M * N * K / us / 1e6:7.2f} TFLOP/s", flush=True)
    for (M, N, K) in [(8, 768, 256), (8, 256, 256), (8, 2048, 256), (8, 256, 2048), (8, 128, 256)]:
        a, w = rnd(M, K), rnd(N, K, scale=K ** -0.5)
        us = timeit(lambda: ops.gemm_skinny(a, w))
        print(f"gemm_skinny M={M} N={N} K={K}: {us:8.1f} us", flush=True)


def misc():
    B, Nt = 1, 8
    q = rnd(B * Nt, 128)
    img = rnd(B * 1024, 384)
    print(f"attn_t2i: {timeit(lambda: ops.attn_t2i(q, img[:, :128], img[:, 128:256], B, Nt, 1024)):8.1f} us")
    k2, v2 = rnd(B * Nt, 128), rnd(B * Nt, 128)
    print(f"attn_i2t: {timeit(lambda: ops.attn_i2t(img[:, 256:], k2, v2, B, 1024, Nt)):8.1f} us")
    s = rnd(B, 1, 128, 128, scale=0.07)
    print(f"fill_holes (local, area 8): {timeit(lambda: ops.fill_holes(s, 8)):8.1f} us")
    print(f"fill_holes (union-find, area 40): {timeit(lambda: ops.fill_holes(s, 40)):8.1f} us")
    x = rnd(1024, 256)
    w, b = rnd(256), rnd(256)
    print(f"layernorm 1024x256: {timeit(lambda: ops.layernorm(x, w, b, 1e-5, bf16=True)):8.1f} us")
    xm = rnd(B * 1024, 256)
    dw, db = rnd(49, 256), rnd(256)
    print(f"dwconv7_ln: {timeit(lambda: ops.dwconv7_ln(xm, dw, db, w, b, B, 32, 32)):8.1f} us")
    m = rnd(B, 512, 512, 1)
    cw, cb, lw, lb = rnd(3, 3, 1, 4), rnd(4), rnd(4), rnd(4)
    print(f"conv2d_small 1->4 @512: {timeit(lambda: ops.conv2d_small(m, cw, cb, B, 512, 512, 1, 4, 3, 2, 1, ln=(lw, lb), gelu=True)):8.1f} us")
    m3 = rnd(B, 128, 128, 16)
    cw3, cb3, lw3, lb3 = rnd(3, 3, 16, 64), rnd(64), rnd(64), rnd(64)
    print(f"conv2d_small 16->64 @128: {timeit(lambda: ops.conv2d_small(m3, cw3, cb3, B, 128, 128, 16, 64, 3, 2, 1, ln=(lw3, lb3), gelu=True)):8.1f} us")
    e = torch.empty(0, device=dev)
    print(f"empty-ish launch (axpby 1 row): {timeit(lambda: ops.axpby(x[:1], None)):8.1f} us")


def phases():
    """GPU time of the three phases of a tracked frame (graph-captured, L2-warm), one object, full memory bank."""
    from us_video_medsam2_b200 import synth
    from us_video_medsam2_b200.build_sam import build_sam2_video_predictor_npz
    pred = build_sam2_video_predictor_npz("configs/sam2.1_hiera_t512.yaml", device=dev, encoder_batch=8)
    pred.load_state_dict(synth.make_state_dict(19), strict=True)
    eng = pred.engine()
    B, T, Nk = 1, 1024, 7 * 1024 + 64
    feat = rnd(T, 256, scale=0.5)
    k_in, v_in = rnd(B, Nk, 64, dtype=torch.bfloat16), rnd(B, Nk, 64, dtype=torch.bfloat16)
    print(f"memory_attention: {timeit(lambda: eng.memory_attention(feat, k_in, v_in, Nk, 64, B), iters=4):7.1f} us", flush=True)
    pix = rnd(B * T, 256, scale=0.5)
    s0, s1 = rnd(16384, 32), rnd(4096, 64)
    print(f"sam_heads: {timeit(lambda: eng.sam_heads(pix, s0, s1, B, None, multimask=True), iters=4):7.1f} us", flush=True)
    low = rnd(B, 1, 128, 128, scale=0.07)
    score = rnd(B, 1)
    fb = rnd(T, 256, dtype=torch.bfloat16)
    print(f"mem_mask_input + encode_memory: "
          f"{timeit(lambda: eng.encode_memory(fb, eng.mem_mask_input(low, False), score, B), iters=4):7.1f} us", flush=True)
    x = rnd(T, 256)
    L = eng.w.ma_layers[0]
    cs, sn = eng.w.rope_cos, eng.w.rope_sin

    def sa_block():
        _, h = ops.layernorm(x, *L["n1"], 1e-5, bf16=True)
        _, qkv = ops.gemm_bf16(h, L["sa_qkv_w"], bias=L["sa_qkv_b"], bf16=True, rope=(cs, sn, 512, T, T))
        o = ops.fmha(qkv, qkv, qkv, B, 1, T, T, 256, (0, T * 768, 768, 256), (256, T * 768, 768, 256),
                     (512, T * 768, 768, 256), num_splits=8, impl="mma")
        return ops.gemm_bf16(o.view(B * T, 256), L["sa_o"][0], bias=L["sa_o"][1], residual=x, f32=True)

    def ffn_block():
        _, h = ops.layernorm(x, *L["n3"], 1e-5, bf16=True)
        _, m = ops.gemm_bf16(h, L["l1"][0], bias=L["l1"][1], act=1, bf16=True)
        return ops.gemm_bf16(m, L["l2"][0], bias=L["l2"][1], residual=x, f32=True)

    print(f"  self-attention sub-block (LN, qkv, fmha+combine, out-proj): {timeit(sa_block, iters=8):7.1f} us", flush=True)
    print(f"  FFN sub-block (LN, linear1, linear2): {timeit(ffn_block, iters=8):7.1f} us", flush=True)
    print(f"  layernorm alone: {timeit(lambda: ops.layernorm(x, *L['n1'], 1e-5, bf16=True), iters=20):7.1f} us", flush=True)
    hb = rnd(T, 256, dtype=torch.bfloat16)
    print(f"  qkv gemm alone: {timeit(lambda: ops.gemm_bf16(hb, L['sa_qkv_w'], bias=L['sa_qkv_b'], bf16=True, rope=(cs, sn, 512, T, T)), iters=20):7.1f} us", flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["attention", "gemms", "misc"]
    for w_ in which:
        globals()[w_]()
