import sys, torch
sys.path.insert(0, '/root/repo')
from us_video_medsam2_b200 import ops
from us_video_medsam2_b200.ops import ACT_RELU
dev = torch.device('cuda')
g = torch.Generator(device=dev).manual_seed(0)
M = 4096
a = torch.randn((M, 256), generator=g, device=dev).to(torch.bfloat16)
w = (torch.randn((256, 256), generator=g, device=dev) * 0.05).to(torch.bfloat16)
b = torch.randn(256, generator=g, device=dev)
res = torch.randn((M, 256), generator=g, device=dev)
lw, lb = torch.randn(256, generator=g, device=dev), torch.randn(256, generator=g, device=dev)
w1 = (torch.randn((2048, 256), generator=g, device=dev) * 0.05).to(torch.bfloat16); b1 = torch.randn(2048, generator=g, device=dev)
w2 = (torch.randn((256, 2048), generator=g, device=dev) * 0.02).to(torch.bfloat16)
wq = (torch.randn((768, 256), generator=g, device=dev) * 0.05).to(torch.bfloat16); bq = torch.randn(768, generator=g, device=dev)
def timeit(fn, name, n=20):
    fn(); torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(n): fn()
    gr.replay(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); gr.replay(); e.record(); torch.cuda.synchronize()
    print(f"{name:50s} {s.elapsed_time(e) / n * 1e3:8.1f} us")
with torch.inference_mode():
    for fused in (True, False):
        timeit(lambda: ops.gemm_bf16(a, w, bias=b, residual=res, f32=True, ln=(lw, lb, 1e-5), ln_fused=fused), f"out-proj + residual + LN, fused={fused}")
    timeit(lambda: ops.gemm_bf16(a, wq, bias=bq, bf16=True), "qkv 4096x768x256")
    def ffn():
        _, m = ops.gemm_bf16(a, w1, bias=b1, act=ACT_RELU, bf16=True)
        ops.gemm_bf16(m, w2, bias=b, residual=res, f32=True)
    timeit(ffn, "FFN two GEMMs")
    timeit(lambda: ops.ffn_fused(a, res, w1, b1, w2, b), "FFN fused cluster kernel")
    x = torch.randn((M, 256), generator=g, device=dev)
    timeit(lambda: ops.layernorm(x, lw, lb, 1e-5, bf16=True), "layernorm 4096x256")
