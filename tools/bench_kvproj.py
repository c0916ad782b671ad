"""Bank key / value projections at the batched shape (32 objects x 7232 rows, K = 64 -> 1024 columns, bf16 out): the most
output-heavy GEMMs of the path (474 MB written each).  USVM2_PGEMM_DEBUG=1 / 2 / 8 / 16 switch parts of the kernel off."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from us_video_medsam2_b200 import ops  # noqa: E402

dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(0)
rnd = lambda *s, dt=torch.bfloat16, sc=1.0: (torch.randn(s, generator=g, device=dev) * sc).to(dt)
B = int(os.environ.get("OBJECTS", 32))
M, N, K = B * 7232, 1024, 64
a, w, b = rnd(M, K), rnd(N, K, sc=K ** -0.5), rnd(N, dt=torch.float32)
cs, sn = ops.tile_rope_table(rnd(1024, 128, dt=torch.float32)), ops.tile_rope_table(rnd(1024, 128, dt=torch.float32))
o16 = torch.empty((M, N), dtype=torch.bfloat16, device=dev)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
for name, rope in (("V projection", None), ("K projection + RoPE", (cs, sn, 1024, 7232, 7168))):
    fn = lambda: ops.gemm_bf16(a, w, bias=b, rope=rope, out_bf16=o16)
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
    print(f"dbg={os.environ.get('USVM2_PGEMM_DEBUG', '0'):>2s} {name:22s} {us:7.1f} us  {2 * M * N / us / 1e3:7.1f} GB/s written", flush=True)
