"""Out-projection + residual + LayerNorm (N = K = 256): fused 128 x 256 epilogue-LN tiles against GEMM + LayerNorm launches,
over the row counts of this path (M = 1024 x objects, 4096 for Hiera-B+ at one object).  python tools/bench_lnfuse.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from us_video_medsam2_b200 import ops

dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(0)


def timeit(fn, n=20):
    fn()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(n):
            fn()
    gr.replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    gr.replay()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n * 1e3


with torch.inference_mode():
    for M in (2048, 4096, 8192, 12288, 16384, 24576, 32768):
        a = torch.randn((M, 256), generator=g, device=dev).to(torch.bfloat16)
        w = (torch.randn((256, 256), generator=g, device=dev) * 0.05).to(torch.bfloat16)
        b = torch.randn(256, generator=g, device=dev)
        res = torch.randn((M, 256), generator=g, device=dev)
        lw, lb = torch.randn(256, generator=g, device=dev), torch.randn(256, generator=g, device=dev)
        t = [timeit(lambda: ops.gemm_bf16(a, w, bias=b, residual=res, f32=True, ln=(lw, lb, 1e-5), ln_fused=f))
             for f in (True, False)]
        print(f"M = {M:6d}: fused {t[0]:6.1f} us   GEMM + LayerNorm {t[1]:6.1f} us")
