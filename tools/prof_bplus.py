"""Per-kernel time table (torch.profiler / CUPTI) of the Hiera-B+ 1024^2 path: one encoder batch and steady-state tracked
frames (eager launches, no graphs, no SM partition).  python tools/prof_bplus.py [encoder|track]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("TQDM_DISABLE", "1")
import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile

from sam2.build_sam import build_sam2_video_predictor_npz
from us_video_medsam2_b200 import ops, synth

what = sys.argv[1] if len(sys.argv) > 1 else "encoder"
dev = torch.device("cuda", 0)
pred = build_sam2_video_predictor_npz("configs/sam2.1_hiera_b+.yaml", device=dev, encoder_batch=8, use_cuda_graphs=False)
pred.load_state_dict(synth.make_bplus_state_dict(11), strict=True)
T = 20
clip = synth.make_clip(T, size=1024, kind="speckle").to(dev)
BOX = np.array([190.0, 170.0, 340.0, 330.0], np.float32)
with torch.inference_mode():
    eng = pred.engine()
    if what == "encoder":
        x = clip[:8].contiguous()
        eng.encode_frames(x)
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            eng.encode_frames(x)
            torch.cuda.synchronize()
        div = 8
    else:
        st = pred.init_state(clip, 512, 512)
        pred.add_new_points_or_box(st, 0, 1, box=BOX * 2)
        it = pred.propagate_in_video(st)
        for _ in range(12):
            next(it)
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(4):
                next(it)
            torch.cuda.synchronize()
        div = 4
rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
tot = sum(e.device_time_total for e in rows)
print(f"{what}: {tot / div / 1e3:.3f} ms of kernel time per {'slice' if what == 'encoder' else 'tracked frame'}")
for e in rows[:28]:
    print(f"{e.device_time_total / div:9.1f} us  {e.count / div:6.1f} x  {e.key[:110]}")
