#!/usr/bin/env python
"""Per-tile clock64() timeline of the persistent GEMM on CTA 0 (USVM2_PGEMM_DEBUG=4): where a tile's time goes."""
import ctypes as C
import os
import sys

os.environ["USVM2_PGEMM_DEBUG"] = os.environ.get("USVM2_PGEMM_DEBUG", "4")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from us_video_medsam2_b200 import _lib, ops  # noqa: E402

dev = torch.device("cuda")


def run(name, M, N, K, act=0, f32out=False, res=False):
    a = torch.randn((M, K), device=dev).to(torch.bfloat16)
    w = (torch.randn((N, K), device=dev) * K ** -0.5).to(torch.bfloat16)
    b = torch.randn((N,), device=dev)
    r = torch.randn((M, N), device=dev) if res else None
    for _ in range(2):
        ops.gemm_bf16(a, w, bias=b, act=act, residual=r, f32=f32out, bf16=not f32out)
    torch.cuda.synchronize()
    buf = (C.c_ulonglong * 1024)()
    _lib.call("usvm_debug_pgemm_profile", C.cast(buf, C.c_void_p))
    t = [[buf[i * 16 + j] for j in range(16)] for i in range(64)]
    print(f"== {name}: M {M} N {N} K {K}  (cycles; epilogue warp 0 columns are relative to its acc_full)")
    print("tile | MMA: acc_free->stage_full ->committed | EPI: wait(prev end->acc_full) ld  math  wait_rd1  st.shared  "
          "fence  tma_issue | rest-of-tile  release | per-tile")
    prev_end = None
    for i in range(2, 9):
        r = t[i]
        if prev_end is None:
            prev_end = t[i - 1][7]
        per = r[7] - prev_end
        print(f"{i:4d} | {r[1] - r[0]:8d} {r[2] - r[1]:10d} | {r[3] - prev_end:8d} {r[4] - r[3]:6d} {r[5] - r[4]:6d} "
              f"{r[8] - r[5]:7d} {r[9] - r[8]:7d} {r[10] - r[9]:7d} {r[11] - r[10]:7d} | {r[6] - r[11]:8d} {r[7] - r[6]:6d} | {per}")
        prev_end = r[7]


if os.environ.get("SHAPES") == "kv":  # the output-heavy bank projections (K = 64 -> 1024 columns)
    run("bank V projection (32 objects)", 32 * 7232, 1024, 64)
    run("out-proj + residual -> f32 (32 objects)", 32768, 256, 256, f32out=True, res=True)
    run("FFN linear1 ReLU (32 objects)", 32768, 2048, 256, act=ops.ACT_RELU)
    sys.exit(0)
run("stage-1 qkv", 262144, 288, 96)
run("stage-1 MLP up GELU", 262144, 384, 96, act=ops.ACT_GELU)
run("stage-3 qkv", 16384, 1152, 384)
run("stage-3 MLP up GELU", 16384 * 4, 1536, 384, act=ops.ACT_GELU)
run("stage-3 MLP down", 16384 * 4, 384, 1536, f32out=True, res=True)
