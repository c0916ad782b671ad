"""The fused window-attention kernel (Hiera stages 1-2) at the encoder's shapes: us and GB/s of algorithmic traffic
(qkv read once + output written once), CUDA events, L2 flushed before every launch.  FRAMES=16 by default."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from us_video_medsam2_b200 import ops  # noqa: E402

dev = torch.device("cuda")
Fr = int(os.environ.get("FRAMES", 16))
g = torch.Generator(device=dev).manual_seed(0)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
#        name                         H    ws pool  C  heads
cases = [("block 0: 128^2, window 8", 128, 8, False, 96, 1), ("block 1: 128^2, window 8, q-pool", 128, 8, True, 192, 2),
         ("block 2: 64^2, window 4", 64, 4, False, 192, 2), ("block 3: 64^2, window 4, q-pool", 64, 4, True, 384, 4)]
only = os.environ.get("CASE")
for ci, (name, H, ws, pool, C, heads) in enumerate(cases):
    if only is not None and int(only) != ci:
        continue
    qkv = (torch.randn((Fr * H * H, 3 * C), generator=g, device=dev) * 0.5).to(torch.bfloat16)
    bias = torch.randn(3 * C, generator=g, device=dev) * 0.1
    fn = lambda: ops.window_attn(qkv, bias, Fr, H, H, ws, pool, C, heads)
    for _ in range(3):
        out = fn()
    tot = 0.0
    for _ in range(10):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        tot += s.elapsed_time(e)
    us = tot / 10 * 1e3
    byts = qkv.numel() * 2 + out.numel() * 2
    print(f"{name:36s} {us:7.1f} us  {byts / us / 1e3:7.1f} GB/s  ({byts / 1e6:.0f} MB)", flush=True)
