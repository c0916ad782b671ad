#!/usr/bin/env python
"""CUDA-event timing of the throughput GEMM shapes (persistent tcgen05 kernel) with an L2 flush before every launch:
us, TFLOP/s and the GB/s of algorithmic traffic (operands once + outputs once)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from us_video_medsam2_b200 import ops  # noqa: E402

dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(0)
rnd = lambda *s, dt=torch.bfloat16, sc=1.0: (torch.randn(s, generator=g, device=dev) * sc).to(dt)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
f32 = torch.float32


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        tot += s.elapsed_time(e)
    return tot / iters * 1e3


def case(name, M, N, K, act=0, rope=False, f32out=False, res=False, bf16out=True):
    a, w, b = rnd(M, K), rnd(N, K, sc=K ** -0.5), rnd(N, dt=f32)
    r = rnd(M, N, dt=f32) if res else None
    rp = None
    if rope:
        cs, sn = ops.tile_rope_table(rnd(1024, 128, dt=f32)), ops.tile_rope_table(rnd(1024, 128, dt=f32))
        rp = (cs, sn, min(N, 1024) if N >= 512 else N, M, M)
    o32 = torch.empty((M, N), dtype=f32, device=dev) if f32out else None
    o16 = torch.empty((M, N), dtype=torch.bfloat16, device=dev) if bf16out else None
    fn = lambda: ops.gemm_bf16(a, w, bias=b, act=act, residual=r, rope=rp, out_f32=o32, out_bf16=o16)
    us = timeit(fn)
    byts = 2 * M * K + 2 * N * K + (4 * M * N if f32out else 0) + (2 * M * N if bf16out else 0) + (4 * M * N if res else 0)
    print(f"{name:58s} M {M:7d} N {N:5d} K {K:5d}: {us:8.1f} us  {2 * M * N * K / us / 1e6:7.1f} TFLOP/s  {byts / us / 1e3:7.1f} GB/s",
          flush=True)


for fr in (8, 16):
    case(f"enc stage-1 qkv ({fr} fr)", fr * 16384, 288, 96)
    case(f"enc stage-1 MLP up GELU ({fr} fr)", fr * 16384, 384, 96, act=ops.ACT_GELU)
    case(f"enc stage-1 MLP down + res -> f32 ({fr} fr)", fr * 16384, 96, 384, f32out=True, res=True, bf16out=False)
    case(f"enc stage-3 qkv ({fr} fr)", fr * 1024, 1152, 384)
    case(f"enc stage-3 MLP up GELU ({fr} fr)", fr * 1024, 1536, 384, act=ops.ACT_GELU)
    case(f"enc stage-3 MLP down + res -> f32 ({fr} fr)", fr * 1024, 384, 1536, f32out=True, res=True, bf16out=False)
case("mem-attn qkv + RoPE (32 obj)", 32768, 768, 256, rope=True)
case("mem-attn out-proj + res -> f32 (32 obj)", 32768, 256, 256, f32out=True, res=True, bf16out=False)
case("mem-attn FFN linear1 ReLU (32 obj)", 32768, 2048, 256, act=ops.ACT_RELU)
case("mem-attn FFN linear2 + res -> f32 (32 obj)", 32768, 256, 2048, f32out=True, res=True, bf16out=False)
case("bank V projection (32 obj)", 32 * 7232, 1024, 64)
case("bank K projection + RoPE (32 obj)", 32 * 7232, 1024, 64, rope=True)
