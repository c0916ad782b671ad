import torch, sys
sys.path.insert(0,'.')
from us_video_medsam2_b200 import ops
for N in (1, 32):
    x = torch.randn((N,1,128,128), device='cuda')
    for _ in range(3): ops.fill_holes(x, 8, 0.1)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(20): ops.fill_holes(x, 8, 0.1)
    g.replay(); torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    us=e0.elapsed_time(e1)/20*1e3
    print(f"fill_holes N={N}: {us:.1f} us  {N*2*128*128*4/us/1e3:.1f} GB/s")
