#!/usr/bin/env python
"""Where a lock-step frame of S videos x Bo objects goes: replays the predictor's own captured graphs (16-frame encoder
pass, B = S * Bo tracked frame) between CUDA events, times the full propagate_in_videos loop on the device and on the
host, and (USVM2_FORK=0 in the environment) the unforked frame."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("TQDM_DISABLE", "1")
import torch  # noqa: E402

from us_video_medsam2_b200 import ops, synth  # noqa: E402
from us_video_medsam2_b200.build_sam import build_sam2_video_predictor_npz  # noqa: E402


def replay_ms(graph, iters=10):
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        graph.replay()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def main():
    dev = torch.device("cuda", 0)
    S, Bo, T = int(os.environ.get("VIDEOS", 8)), int(os.environ.get("OBJECTS", 4)), int(os.environ.get("FRAMES", 40))
    pred = build_sam2_video_predictor_npz("configs/sam2.1_hiera_t512.yaml", device=dev, encoder_batch=16)
    pred.load_state_dict(synth.make_state_dict(19), strict=True)
    clips = [ops.normalize_gray_u8(synth.make_clip_u8(T, seed=1 + i).to(dev), synth.IMG_MEAN, synth.IMG_STD) for i in range(S)]
    masks = synth.multi_object_masks(Bo) if Bo > 1 else [synth.box_mask()]

    def one_pass():
        states = []
        for c in clips:
            st = pred.init_state(c, 512, 512)
            for j, m in enumerate(masks):
                pred.add_new_mask(st, 0, j + 1, m)
            states.append(st)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        n = sum(1 for _ in pred.propagate_in_videos(states))
        host = time.perf_counter() - t0
        e.record()
        torch.cuda.synchronize()
        return n, s.elapsed_time(e), host * 1e3

    with torch.inference_mode():
        for _ in range(2):
            one_pass()
        n, dev_ms, host_ms = one_pass()
        print(f"lock-step loop: {n} frames, device {dev_ms / n:.3f} ms/frame, host loop {host_ms / n:.3f} ms/frame")
        enc = pred._graphs.get(("encoder", 16, 0, False))
        keys = [k for k in pred._graphs if k[0] != "encoder"]
        steady = max(keys, key=lambda k: (k[0], k[1], k[2]))
        trk = pred._graphs[steady]
        print(f"tracked-frame graph {steady}: {trk[3]} kernels; encoder graph {enc[3]} kernels per 16 frames")
        print(f"tracked frame (B = {steady[0]}) graph replay: {replay_ms(trk[0]):.3f} ms")
        print(f"encoder graph replay (16 frames)            : {replay_ms(enc[0]):.3f} ms")


if __name__ == "__main__":
    main()
