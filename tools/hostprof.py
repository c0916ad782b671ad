import os, sys, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); os.environ["TQDM_DISABLE"]="1"
import torch
from us_video_medsam2_b200 import synth, ops
from us_video_medsam2_b200.build_sam import build_sam2_video_predictor_npz
T=int(sys.argv[1]) if len(sys.argv)>1 else 128
pred = build_sam2_video_predictor_npz("configs/sam2.1_hiera_t512.yaml", encoder_batch=int(os.environ.get("ENCODER_BATCH", 8)),
                                      encoder_sms=int(os.environ.get("ENCODER_SMS", 0)))
pred.load_state_dict(synth.make_state_dict(19))
clip = synth.make_clip(T).cuda()
def one():
    st = pred.init_state(clip, 512, 512)
    pred.add_new_mask(st, 0, 1, synth.box_mask())
    n=0
    for t, ids, lg in pred.propagate_in_video(st): n+=1
    return n
with torch.inference_mode():
    for i in range(4):
        torch.cuda.synchronize(); t=time.perf_counter(); one(); th=time.perf_counter()-t; torch.cuda.synchronize(); tt=time.perf_counter()-t
        print(f"pass {i}: host {th*1e3:.1f} ms, total {tt*1e3:.1f} ms, per frame host {th/T*1e3:.3f} total {tt/T*1e3:.3f}", flush=True)
    pr = cProfile.Profile(); pr.enable(); one(); torch.cuda.synchronize(); pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(35)
