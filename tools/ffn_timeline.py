#!/usr/bin/env python
"""Phase timeline (clock64) of one CTA of the fused feed-forward kernel; needs `make -C us_video_medsam2_b200/csrc
EXTRA=-DUSVM2_FFN_PROFILE` (the stamps are compiled out by default)."""
import ctypes as C, sys, torch
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from us_video_medsam2_b200 import ops, _lib
M=1024
x = torch.randn((M, 256), device="cuda"); h = torch.randn((M, 256), device="cuda").to(torch.bfloat16)
w1 = (torch.randn((2048, 256), device="cuda") / 16).to(torch.bfloat16); w2 = (torch.randn((256, 2048), device="cuda") / 45).to(torch.bfloat16)
b1, b2 = torch.randn((2048,), device="cuda"), torch.randn((256,), device="cuda")
for _ in range(3): ops.ffn_fused(h, x, w1, b1, w2, b2)
torch.cuda.synchronize()
lib=_lib.lib(); lib.usvm_debug_ffn_profile.argtypes=[C.c_void_p]
buf=(C.c_ulonglong*16)(); lib.usvm_debug_ffn_profile(C.cast(buf,C.c_void_p))
t=[buf[i] for i in range(10)]
names=["start","pre-pdl","post-pdl","d1_full","h written","d2_full","cluster bar 1","pushed","cluster bar 2","end"]
for n,v in zip(names,t): print(f"{n:14s} {v-t[0]:8d}")
