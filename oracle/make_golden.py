"""TEST INFRASTRUCTURE -- generates tests/golden/*.npz by running the REAL reference
(/root/reference, build container only) on seeded synthetic inputs and seeded weights.

    python oracle/make_golden.py            # rewrites every fixture

Fixtures hold the reference's own outputs: low-res mask logits per frame/object (what
`propagate_in_video` stores as `pred_masks`; the yielded video-res logits are their bilinear
upsampling), object pointers, object scores, one bf16 memory feature map and strided samples
of the frame-0 backbone features.

Two variants of the mask logits are stored:
  * `low_res`        -- stock reference on CPU (its CUDA-only CC op raises, the fill-holes step is
                        skipped with a warning: sam2/utils/misc.py:321-336);
  * `low_res_filled` -- same run with `sam2.utils.misc.get_connected_components` replaced by the
                        scipy restatement in oracle/cc_ref.py, i.e. what the reference computes
                        on a GPU where `sam2._C` works.  Point/box prompts feed the hole-filled
                        prompt-frame mask into the memory encoder, so pointers, scores and memory
                        are stored per variant too (`*_filled`).
"""
import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("TQDM_DISABLE", "1")
warnings.filterwarnings("ignore")

from oracle.ref_loader import load_reference_predictor  # noqa: E402
from oracle.cc_ref import connected_components_ref  # noqa: E402
from us_video_medsam2_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

CASES = {
    # name: seed, frames, prompts [(kind, frame, obj_id, payload)], propagate kwargs
    "t512_mask_fwd": dict(seed=19, T=8, prompts=[("mask", 0, 1, dict())], prop=dict()),
    "t512_absent_fwd": dict(seed=28, T=4, prompts=[("mask", 0, 1, dict())], prop=dict()),
    "t512_two_obj_mask_box": dict(
        seed=19, T=5,
        prompts=[("mask", 0, 1, dict(y0=200, y1=300, x0=220, x1=330)),
                 ("box", 0, 2, dict(box=[60.0, 80.0, 200.0, 190.0]))],
        prop=dict()),
    "t512_points_reverse": dict(
        seed=26, T=5,
        prompts=[("points", 4, 7, dict(points=[[256.0, 250.0], [100.0, 400.0]], labels=[1, 0]))],
        prop=dict(reverse=True)),
    # apply_postprocessing=False (build_sam.py:108-122 overrides absent): a box prompt is 2 points, so the prompt frame
    # takes the single-mask branch WITHOUT the stability fallback (mask_decoder.py:160-166), memory-encoder input is
    # sigmoid-scaled instead of binarised, no hole filling
    "t512_box_nopost": dict(seed=19, T=4, prompts=[("box", 0, 3, dict(box=[220.0, 200.0, 330.0, 300.0]))],
                            prop=dict(), post=False),
}


def _patched_cc(mask):
    lab, cnt = connected_components_ref(mask.to(torch.uint8).cpu().numpy())
    return torch.from_numpy(lab), torch.from_numpy(cnt)


def run_case(model, cfg, fill):
    import sam2.utils.misc as misc  # the reference's module (ref_loader put it first on sys.path)

    stock = misc.get_connected_components
    if fill:
        misc.get_connected_components = _patched_cc
    try:
        sd = synth.make_state_dict(cfg["seed"])
        model.load_state_dict(sd, strict=True)
        clip = synth.make_clip(cfg["T"], kind="speckle")
        with torch.inference_mode():
            st = model.init_state(clip, 512, 512)
            prompt_out = []
            for kind, t, oid, pl in cfg["prompts"]:
                if kind == "mask":
                    _, _, lg = model.add_new_mask(st, t, oid, synth.box_mask(512, **pl) if pl else synth.box_mask())
                elif kind == "box":
                    _, _, lg = model.add_new_points_or_box(st, t, oid, box=np.array(pl["box"], np.float32))
                else:
                    _, _, lg = model.add_new_points_or_box(
                        st, t, oid, points=np.array(pl["points"], np.float32),
                        labels=np.array(pl["labels"], np.int32))
                prompt_out.append(lg[:, 0, ::4, ::4].numpy().copy())
            frames, low = [], []
            for t, ids, lg in model.propagate_in_video(st, **cfg["prop"]):
                frames.append(t)
                key = "cond_frame_outputs" if t in st["output_dict"]["cond_frame_outputs"] else "non_cond_frame_outputs"
                low.append(st["output_dict"][key][t]["pred_masks"][:, 0].numpy().copy())
        od = st["output_dict"]
        get = lambda t: od["cond_frame_outputs"].get(t) or od["non_cond_frame_outputs"][t]
        res = dict(frames=np.array(frames, np.int32), low=np.stack(low),
                   obj_ptr=np.stack([get(t)["obj_ptr"].numpy() for t in frames]),
                   score=np.stack([get(t)["object_score_logits"].numpy() for t in frames]),
                   maskmem_last=get(frames[-1])["maskmem_features"].float().numpy(),
                   prompt_video_res_s4=np.stack(prompt_out[-1:]))
        return res, clip, st
    finally:
        misc.get_connected_components = stock


def main():
    os.makedirs(OUT, exist_ok=True)
    only = sys.argv[1:]
    models = {}
    for name, cfg in CASES.items():
        if only and name not in only:
            continue
        post = cfg.get("post", True)
        if post not in models:
            models[post] = load_reference_predictor(apply_postprocessing=post)
        model = models[post]
        plain, clip, _ = run_case(model, cfg, fill=False)
        filled, _, _ = run_case(model, cfg, fill=True)
        n_holes = int((plain["low"] != filled["low"]).sum())
        blob = dict(frames=plain["frames"], low_res=plain["low"], low_res_filled=filled["low"],
                    obj_ptr=plain["obj_ptr"], score=plain["score"], maskmem_last=plain["maskmem_last"],
                    obj_ptr_filled=filled["obj_ptr"], score_filled=filled["score"],
                    maskmem_last_filled=filled["maskmem_last"],
                    prompt_video_res_s4=plain["prompt_video_res_s4"], seed=np.int32(cfg["seed"]),
                    num_frames=np.int32(cfg["T"]))
        if name == "t512_mask_fwd":
            with torch.inference_mode():
                bo = model.forward_image(clip[:1])
            blob["enc_feat_s0_s8"] = bo["backbone_fpn"][0][0, :, ::8, ::8].numpy()
            blob["enc_feat_s1_s4"] = bo["backbone_fpn"][1][0, :, ::4, ::4].numpy()
            blob["enc_feat_s2"] = bo["backbone_fpn"][2][0, :, ::2, ::2].numpy()
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **blob)
        print(name, "frames", plain["frames"].tolist(), "score", plain["score"].reshape(-1)[-3:],
              "pixels changed by fill-holes:", n_holes, flush=True)


if __name__ == "__main__":
    main()
