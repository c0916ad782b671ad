"""Replays the prompt scripts of oracle/make_golden.py on any predictor with the reference API."""
import numpy as np
import torch

from oracle.make_golden import CASES
from us_video_medsam2_b200 import synth


def replay(predictor, name, images=None, to_cpu=True):
    """Run case `name`; returns dict(frames, low [T,B,128,128], obj_ptr, score, state)."""
    cfg = CASES[name]
    clip = synth.make_clip(cfg["T"], kind="speckle") if images is None else images
    st = predictor.init_state(clip, 512, 512)
    for kind, t, oid, pl in cfg["prompts"]:
        if kind == "mask":
            predictor.add_new_mask(st, t, oid, synth.box_mask(512, **pl) if pl else synth.box_mask())
        elif kind == "box":
            predictor.add_new_points_or_box(st, t, oid, box=np.array(pl["box"], np.float32))
        else:
            predictor.add_new_points_or_box(st, t, oid, points=np.array(pl["points"], np.float32),
                                            labels=np.array(pl["labels"], np.int32))
    frames, low, video = [], [], []
    od = st["output_dict"]
    for t, ids, lg in predictor.propagate_in_video(st, **cfg["prop"]):
        frames.append(t)
        out = od["cond_frame_outputs"].get(t) or od["non_cond_frame_outputs"][t]
        low.append(out["pred_masks"][:, 0].float().cpu().clone())
        video.append(lg.float().cpu().clone() if to_cpu else lg)
    get = lambda t: od["cond_frame_outputs"].get(t) or od["non_cond_frame_outputs"][t]
    return dict(frames=frames, low=torch.stack(low), video=video,
                obj_ptr=torch.stack([get(t)["obj_ptr"].float().cpu() for t in frames]),
                score=torch.stack([get(t)["object_score_logits"].float().cpu() for t in frames]),
                maskmem_last=get(frames[-1])["maskmem_features"].float().cpu(), state=st)


def dice(a, b):
    A, B = a > 0, b > 0
    den = int(A.sum()) + int(B.sum())
    return 1.0 if den == 0 else 2.0 * int((A & B).sum()) / den
