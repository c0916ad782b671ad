#!/usr/bin/env python
"""Memory-attention feed-forward block at one object (M = 1024): fused cluster kernel vs the two GEMM launches, replayed
from a CUDA graph of 20 back-to-back calls (PDL between them, as in the frame graph)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from us_video_medsam2_b200 import ops  # noqa: E402


def timeit(fn, reps=20):
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


for M in (1024, 4096):
    x = torch.randn((M, 256), device="cuda")
    h = torch.randn((M, 256), device="cuda").to(torch.bfloat16)
    w1 = (torch.randn((2048, 256), device="cuda") / 16).to(torch.bfloat16)
    w2 = (torch.randn((256, 2048), device="cuda") / 45).to(torch.bfloat16)
    b1, b2 = torch.randn((2048,), device="cuda"), torch.randn((256,), device="cuda")

    def two():
        _, m = ops.gemm_bf16(h, w1, bias=b1, act=ops.ACT_RELU, bf16=True)
        return ops.gemm_bf16(m, w2, bias=b2, residual=x, f32=True)[0]

    print(f"M {M}: two launches {timeit(two):6.1f} us   fused {timeit(lambda: ops.ffn_fused(h, x, w1, b1, w2, b2)):6.1f} us")
