#!/usr/bin/env python
"""Where a propagated frame's GPU time goes: replays the predictor's own captured CUDA graphs (batched image encoder,
steady-state tracked frame) between CUDA events, alone and overlapped on two streams.  Prints one line per case."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from us_video_medsam2_b200 import ops, synth  # noqa: E402
from us_video_medsam2_b200.build_sam import build_sam2_video_predictor_npz  # noqa: E402


def replay_ms(graphs, iters=20, streams=None):
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    main = torch.cuda.current_stream()
    s.record()
    if streams is None:
        for _ in range(iters):
            for g, reps in graphs:
                for _ in range(reps):
                    g.replay()
    else:
        done = []
        for (g, reps), st in zip(graphs, streams):
            st.wait_stream(main)
            with torch.cuda.stream(st):
                for _ in range(iters):
                    for _ in range(reps):
                        g.replay()
            done.append(st)
        for st in done:
            main.wait_stream(st)
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def main():
    dev = torch.device("cuda", 0)
    T = int(os.environ.get("FRAMES", 48))
    nb = int(os.environ.get("ENCODER_BATCH", 8))
    pred = build_sam2_video_predictor_npz("configs/sam2.1_hiera_t512.yaml", device=dev, encoder_batch=nb)
    pred.load_state_dict(synth.make_state_dict(19), strict=True)
    clip = ops.normalize_gray_u8(synth.make_clip_u8(T, seed=1).to(dev), synth.IMG_MEAN, synth.IMG_STD)
    for _ in range(2):
        st = pred.init_state(clip, 512, 512)
        pred.add_new_mask(st, 0, 1, synth.box_mask())
        for _ in pred.propagate_in_video(st):
            pass
    torch.cuda.synchronize()
    enc = pred._graphs[("encoder", nb)]
    track_keys = [k for k in pred._graphs if k[0] != "encoder"]
    steady = max(track_keys, key=lambda k: (k[1], k[2]))
    trk = pred._graphs[steady]
    print(f"encoder graph: {enc[3]} kernels per {nb} frames; tracked-frame graph {steady}: {trk[3]} kernels")
    t_enc = replay_ms([(enc[0], 1)])
    t_trk = replay_ms([(trk[0], nb)])
    print(f"encoder alone   : {t_enc:7.3f} ms per {nb} frames = {t_enc / nb:6.3f} ms/frame")
    print(f"tracking alone  : {t_trk:7.3f} ms per {nb} frames = {t_trk / nb:6.3f} ms/frame")
    t_seq = replay_ms([(enc[0], 1), (trk[0], nb)])
    print(f"back to back    : {t_seq:7.3f} ms per {nb} frames = {t_seq / nb:6.3f} ms/frame")
    s_enc, s_trk = torch.cuda.Stream(priority=0), torch.cuda.Stream(priority=-1)
    t_ovl = replay_ms([(enc[0], 1), (trk[0], nb)], streams=[s_enc, s_trk])
    print(f"two streams     : {t_ovl:7.3f} ms per {nb} frames = {t_ovl / nb:6.3f} ms/frame (encoder low, tracking high priority)")
    s_enc, s_trk = torch.cuda.Stream(), torch.cuda.Stream()
    t_ovl = replay_ms([(enc[0], 1), (trk[0], nb)], streams=[s_enc, s_trk])
    print(f"two streams     : {t_ovl:7.3f} ms per {nb} frames = {t_ovl / nb:6.3f} ms/frame (equal priority)")


if __name__ == "__main__":
    main()
