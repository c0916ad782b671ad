"""LayerNorm at the encoder's shapes (fp32 in -> bf16 out): us and GB/s, CUDA events, L2 flushed before every launch."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from us_video_medsam2_b200 import ops  # noqa: E402

dev = torch.device("cuda")
Fr = int(os.environ.get("FRAMES", 16))
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
for rows, C in ((Fr * 16384, 96), (Fr * 4096, 192), (Fr * 1024, 384), (Fr * 256, 768), (32768, 256)):
    x = torch.randn((rows, C), device=dev)
    w, b = torch.randn(C, device=dev), torch.randn(C, device=dev)
    fn = lambda: ops.layernorm(x, w, b, 1e-6, bf16=True)
    for _ in range(3):
        fn()
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
    print(f"rows {rows:7d} C {C:4d}: {us:7.1f} us  {rows * C * 6 / us / 1e3:7.1f} GB/s", flush=True)
