"""Step-by-step probe of the encoder-on-an-SM-partition path (green context): prints what works and how fast.
  python tools/probe_partition.py [sms ...]"""
import os
import sys
import time
import traceback

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("TQDM_DISABLE", "1")
import torch  # noqa: E402

from us_video_medsam2_b200 import ops, synth  # noqa: E402
from us_video_medsam2_b200.build_sam import build_sam2_video_predictor_npz  # noqa: E402
from us_video_medsam2_b200.pipeline import SmPartition  # noqa: E402


def timeit(fn, stream=None, iters=5):
    stream = stream or torch.cuda.current_stream()
    with torch.cuda.stream(stream):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    sms_list = [int(a) for a in sys.argv[1:]] or [56]
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    n = 8
    pred = build_sam2_video_predictor_npz("configs/sam2.1_hiera_t512.yaml", device=dev, encoder_batch=n)
    pred.load_state_dict(synth.make_state_dict(19), strict=True)
    eng = pred._sync_engine()
    imgs = synth.make_clip(n, kind="speckle").to(dev).float().contiguous()
    with torch.inference_mode():
        ref = eng.encode_frames(imgs)
        print(f"[probe] eager encoder, {n} frames, whole device: {timeit(lambda: eng.encode_frames(imgs)):.3f} ms", flush=True)
        for sms in sms_list:
            try:
                part = SmPartition(dev, sms)
                print(f"[probe] partition requested {sms} -> granted {part.sms} of {part.total_sms}", flush=True)
                ops.set_sm_budget(part.sms)
                part.stream.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(part.stream):
                    out = eng.encode_frames(imgs)
                torch.cuda.synchronize()
                same = all(torch.equal(out[k], ref[k]) for k in ref)
                print(f"[probe]   eager on partition: bit-identical={same}, "
                      f"{timeit(lambda: eng.encode_frames(imgs), part.stream):.3f} ms", flush=True)
                ops.set_sm_budget(0)
                graph, static_in, gout, nk = pred._encoder_graph(n, 0, part)
                static_in.copy_(imgs)
                with torch.cuda.stream(part.stream):
                    graph.replay()
                torch.cuda.synchronize()
                same = all(torch.equal(gout[k], ref[k]) for k in ref)
                print(f"[probe]   graph on partition ({nk} kernels): bit-identical={same}, "
                      f"{timeit(graph.replay, part.stream):.3f} ms", flush=True)
                pred._graphs = {}
            except Exception:
                print(f"[probe]   FAILED for sms={sms}", flush=True)
                traceback.print_exc()
                ops.set_sm_budget(0)
    print("[probe] done", flush=True)


if __name__ == "__main__":
    main()
