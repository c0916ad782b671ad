#!/usr/bin/env python
"""Memory-encoder pieces at 32 objects (graph replay): depthwise 7x7 + LayerNorm, im2col of the last down-sampler stage."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from us_video_medsam2_b200 import ops  # noqa: E402


def timeit(fn, reps=10):
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


for B in (1, 32):
    x = torch.randn((B * 1024, 256), device="cuda")
    dw, db, lw, lb = (torch.randn(s, device="cuda") for s in ((49, 256), (256,), (256,), (256,)))
    t = timeit(lambda: ops.dwconv7_ln(x, dw, db, lw, lb, B, 32, 32))
    print(f"dwconv7 + LN, {B:2d} objects: {t:7.1f} us  {B * 1024 * 256 * 6 / t / 1e3:7.1f} GB/s (fp32 in + bf16 out)")
    y = torch.randn((B * 64 * 64, 64), device="cuda")
    t = timeit(lambda: ops.im2col_nhwc(y, B, 64, 64, 64, 3, 2, 1))
    print(f"im2col 3x3/s2 of 64x64x64, {B:2d} objects: {t:7.1f} us  {B * (64 * 64 * 64 * 4 + 1024 * 576 * 2) / t / 1e3:7.1f} GB/s")
# batched mask decoder: image-side tf32 GEMMs at 32 objects
for M, N, K in ((32768, 256, 256), (32768, 128, 256), (32768, 256, 128)):
    a = torch.randn((M, K), device="cuda")
    w = torch.randn((N, K), device="cuda") / K ** 0.5
    b = torch.randn((N,), device="cuda")
    t = timeit(lambda: ops.gemm_f32(a, w, b, tf32=True))
    print(f"tf32 GEMM M {M} N {N} K {K}: {t:7.1f} us  {2.0 * M * N * K / t / 1e6:7.1f} TFLOP/s")
# mask down-sampler stages at 32 objects
for Cin, Cout, H in ((1, 4, 512), (4, 16, 256), (16, 64, 128)):
    for B in (1, 32):
        x = torch.randn((B, H, H, Cin), device="cuda")
        w = torch.randn((3, 3, Cin, Cout), device="cuda") / (9 * Cin) ** 0.5
        b, lw, lb = (torch.randn(Cout, device="cuda") for _ in range(3))
        t = timeit(lambda: ops.conv2d_small(x, w, b, B, H, H, Cin, Cout, 3, 2, 1, ln=(lw, lb), gelu=True))
        print(f"conv3x3/s2 {Cin:2d}->{Cout:2d} on {H}^2, {B:2d} objects: {t:7.1f} us")
