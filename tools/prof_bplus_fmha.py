import sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tools')
from bench_bplus import time_cross_attention
from bench import measured_peaks
with torch.inference_mode():
    print(time_cross_attention(torch.device('cuda', 0), measured_peaks(), 4, iters=2))
