#!/usr/bin/env python
"""Times the image-encoder attention kernels at the Hiera stage-3 shapes (16 frames x 32 x 32 tokens x 384, 4 heads of 96):
the tcgen05 kernel against the mma.sync path it replaces (global: fmha_bf16<96>; windowed: gather + fmha + scatter)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from us_video_medsam2_b200 import ops  # noqa: E402


def timeit(fn, reps=20):
    """Device time per call, replayed from a CUDA graph (the encoder runs from one): no host launch cost in it."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def main():
    Fr, H, W, dim, heads = int(os.environ.get("FRAMES", 16)), 32, 32, 384, 4
    T = H * W
    qkv = torch.randn((Fr * T, 3 * dim), device="cuda").to(torch.bfloat16)
    bias = torch.randn((3 * dim,), device="cuda")
    gf_global = 4.0 * Fr * heads * T * T * 96 / 1e9
    gf_win = 4.0 * Fr * heads * 9 * 196 * 196 * 96 / 1e9

    def old_global():
        return ops.fmha(qkv, qkv, qkv, Fr, heads, T, T, 96, (0, T * 3 * dim, 3 * dim, 96),
                        (dim, T * 3 * dim, 3 * dim, 96), (2 * dim, T * 3 * dim, 3 * dim, 96))

    def old_win():
        Qw, Kw, Vw, nw, nq, nk = ops.window_gather(qkv, bias, Fr, H, W, 14, False, dim)
        Ow = ops.fmha(Qw, Kw, Vw, Fr * nw, heads, nq, nk, 96, (0, nq * dim, dim, 96), (0, nk * dim, dim, 96),
                      (0, nk * dim, dim, 96))
        return ops.window_scatter(Ow, Fr, H, W, 14, dim)

    t = timeit(old_global)
    print(f"global  mma.sync : {t:8.1f} us  {gf_global / t * 1e3:7.1f} TFLOP/s")
    t = timeit(lambda: ops.hiera_attn(qkv, bias, Fr, H, W, dim, heads, 0))
    print(f"global  tcgen05  : {t:8.1f} us  {gf_global / t * 1e3:7.1f} TFLOP/s")
    t = timeit(old_win)
    print(f"window  mma.sync : {t:8.1f} us  (gather + flash + scatter)")
    t = timeit(lambda: ops.hiera_attn(qkv, bias, Fr, H, W, dim, heads, 14))
    print(f"window  tcgen05  : {t:8.1f} us  {gf_win / t * 1e3:7.1f} TFLOP/s (padded windows counted)")


if __name__ == "__main__":
    main()
