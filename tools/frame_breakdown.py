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
    enc = pred._graphs[("encoder", nb, 0, False)]
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
    # encoder on a green-context SM partition, tracked frames (captured for the remaining SMs) on an ordinary stream
    for sms in [int(x) for x in os.environ.get("PARTITION_SMS", "48,64").split(",") if x]:
        p2 = build_sam2_video_predictor_npz("configs/sam2.1_hiera_t512.yaml", device=dev, encoder_batch=nb, encoder_sms=sms)
        p2.load_state_dict(synth.make_state_dict(19), strict=True)
        for _ in range(3):
            st = p2.init_state(clip, 512, 512)
            p2.add_new_mask(st, 0, 1, synth.box_mask())
            for _ in p2.propagate_in_video(st):
                pass
        torch.cuda.synchronize()
        part = p2._partition_obj
        if part is None:
            print(f"partition {sms}: unavailable ({p2._partition_error})")
            continue
        enc2 = p2._graphs[("encoder", nb, 0, True)]
        steady2 = max([k for k in p2._graphs if k[0] != "encoder"], key=lambda k: (k[1], k[2]))
        trk2 = p2._graphs[steady2]
        t_e = replay_ms([(enc2[0], 1)], streams=[part.stream])
        t_t = replay_ms([(trk2[0], nb)])
        t_o = replay_ms([(enc2[0], 1), (trk2[0], nb)], streams=[part.stream, torch.cuda.Stream()])
        print(f"partition {part.sms:3d} SMs: encoder alone {t_e / nb:6.3f}, tracking alone (budget {part.total_sms - part.sms}) "
              f"{t_t / nb:6.3f}, concurrent {t_o / nb:6.3f} ms/frame", flush=True)
        # the same tracked frame captured on the COMPLEMENT partition (disjoint SM sets)
        try:
            p2.engine()._tail_stream = part.rest_stream2  # the forked user-facing tail stays in the same green context
            trk3 = p2._capture_graph(steady2, trk2[1], stream=part.rest_stream)
            t_t3 = replay_ms([(trk3[0], nb)], streams=[part.rest_stream])
            t_o3 = replay_ms([(enc2[0], 1), (trk3[0], nb)], streams=[part.stream, part.rest_stream])
            print(f"              tracking on the complement ({part.rest_sms} SMs) alone {t_t3 / nb:6.3f}, concurrent "
                  f"{t_o3 / nb:6.3f} ms/frame", flush=True)
        except Exception as e:
            print("              complement capture failed:", repr(e))
    # the other way round: the TRACKED FRAME on the first-class SM group (cluster-friendly), encoder on the remainder
    from us_video_medsam2_b200.pipeline import SmPartition
    for sms in [int(x) for x in os.environ.get("TRACK_SMS", "96,104,112").split(",") if x]:
        try:
            part = SmPartition(dev, sms)
            p3 = build_sam2_video_predictor_npz("configs/sam2.1_hiera_t512.yaml", device=dev, encoder_batch=nb)
            p3.load_state_dict(synth.make_state_dict(19), strict=True)
            eng = p3._sync_engine()
            eng.sm_budget = part.sms
            for _ in range(2):
                st = p3.init_state(clip, 512, 512)
                p3.add_new_mask(st, 0, 1, synth.box_mask())
                for _ in p3.propagate_in_video(st):
                    pass
            torch.cuda.synchronize()
            steady3 = max([k for k in p3._graphs if k[0] != "encoder"], key=lambda k: (k[1], k[2]))
            base = p3._graphs[steady3]
            ops.set_sm_budget(part.sms)
            eng._tail_stream = None
            # second stream of the same green context for the forked tail: reuse rest_stream2's constructor path
            from cuda.bindings import driver as drv
            err, raw2 = drv.cuGreenCtxStreamCreate(part._ctx, drv.CUstream_flags.CU_STREAM_NON_BLOCKING, 0)
            eng._tail_stream = torch.cuda.ExternalStream(int(raw2), device=dev)
            trk = p3._capture_graph(steady3, base[1], stream=part.stream)
            ops.set_sm_budget(part.rest_sms)
            # encoder captured on the remainder green context
            class _P:  # duck-typed partition view for _encoder_graph
                pass
            rp = _P()
            rp.stream, rp.sms, rp.total_sms = part.rest_stream, part.rest_sms, part.total_sms
            enc3 = p3._encoder_graph(nb, 0, rp)
            t_t = replay_ms([(trk[0], nb)], streams=[part.stream])
            t_e = replay_ms([(enc3[0], 1)], streams=[part.rest_stream])
            t_o = replay_ms([(enc3[0], 1), (trk[0], nb)], streams=[part.rest_stream, part.stream])
            print(f"tracking on group {part.sms:3d} SMs alone {t_t / nb:6.3f}, encoder on remainder {part.rest_sms} alone "
                  f"{t_e / nb:6.3f}, concurrent {t_o / nb:6.3f} ms/frame", flush=True)
        except Exception as e:
            import traceback
            traceback.print_exc()
            print(f"swap {sms}: failed {e!r}", flush=True)
        finally:
            ops.set_sm_budget(0)


if __name__ == "__main__":
    main()
