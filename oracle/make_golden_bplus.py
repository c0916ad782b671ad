"""TEST INFRASTRUCTURE -- fixture for BASELINE configs[4] / SURVEY 8f-2: sam2.1_hiera_base_plus at 1024^2 driven the way
medsam2_infer_3D_CT.py:256-283 drives it (box prompt on a key slice, forward propagation, reset_state, the same box
again, reverse propagation, union of the two passes), slices of 512 x 512 resized to 1024 x 1024 for the model.

    python oracle/make_golden_bplus.py        # -> tests/golden/bplus1024_ct_bidirectional.npz   (REAL reference, CPU)

The reference ships no B+ YAML; the model is its own classes instantiated with oracle/ref_loader.bplus_overrides().
Also stores the image encoder's three feature levels of the key slice (sub-sampled), so that the encoder is pinned on
its own and not only through the masks.
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

SEED, T, KEY = 11, 6, 2
H, W = 512, 512
BOX = np.array([190.0, 170.0, 340.0, 330.0], np.float32)  # x0, y0, x1, y1 in slice pixels
OUT = os.path.join(ROOT, "tests", "golden", "bplus1024_ct_bidirectional.npz")


def _entry(st, t):
    od = st["output_dict"]
    return od["cond_frame_outputs"].get(t) or od["non_cond_frame_outputs"][t]


def bplus_session(pred, clip):
    """The driver's sequence on any predictor with the reference API."""
    segs = np.zeros((T, H, W), np.uint8)
    rec = {}
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=torch.cuda.is_available()):
        st = pred.init_state(clip, H, W)
        for name, kw in (("fwd", {}), ("rev", dict(reverse=True))):
            _, ids, lg = pred.add_new_points_or_box(st, KEY, 1, box=BOX)
            rec[f"prompt_{name}_s2"] = lg[0, 0, ::2, ::2].float().cpu().numpy().copy()
            frames, logits, low, ptr, score = [], [], [], [], []
            for t, ids, lg in pred.propagate_in_video(st, **kw):
                frames.append(t)
                logits.append(lg[0, 0, ::2, ::2].float().cpu().numpy().copy())
                segs[t, (lg[0] > 0.0).cpu().numpy()[0]] = 1
                e = _entry(st, t)
                low.append(e["pred_masks"][0, 0].float().cpu().numpy().copy())
                ptr.append(e["obj_ptr"][0].float().cpu().numpy().copy())
                score.append(e["object_score_logits"].float().cpu().numpy().reshape(-1).copy())
            rec[f"frames_{name}"] = np.array(frames, np.int32)
            rec[f"logits_{name}_s2"] = np.stack(logits)
            rec[f"low_{name}"] = np.stack(low)
            rec[f"obj_ptr_{name}"] = np.stack(ptr)
            rec[f"score_{name}"] = np.stack(score)
            if name == "rev":
                rec["maskmem_last"] = _entry(st, frames[-1])["maskmem_features"].float().cpu().numpy()[:, :, ::2, ::2].copy()
            pred.reset_state(st)
    rec["segs"] = segs
    return rec


def main():
    from oracle.cc_ref import connected_components_ref
    from oracle.ref_loader import bplus_overrides, load_reference_predictor
    from us_video_medsam2_b200 import synth

    model = load_reference_predictor(overrides=bplus_overrides())
    import sam2.utils.misc as misc

    def patched(mask):
        lab, cnt = connected_components_ref(mask.to(torch.uint8).cpu().numpy())
        return torch.from_numpy(lab), torch.from_numpy(cnt)

    misc.get_connected_components = patched
    model.load_state_dict(synth.make_bplus_state_dict(SEED), strict=True)
    clip = synth.make_clip(T, size=1024, kind="speckle")
    with torch.inference_mode():
        out = model.forward_image(clip[KEY:KEY + 1])
        rec = bplus_session(model, clip)
    fpn = out["backbone_fpn"]  # conv_s0 / conv_s1 already applied (sam2_base.py:1220-1232)
    rec["enc_feat"] = fpn[2][0, :, ::2, ::2].float().numpy().copy()       # [256, 32, 32] of the 64 x 64 level
    rec["enc_feat_s1"] = fpn[1][0, :, ::4, ::4].float().numpy().copy()    # [64, 32, 32] of the 128 x 128 level
    rec["enc_feat_s0"] = fpn[0][0, :, ::8, ::8].float().numpy().copy()    # [32, 32, 32] of the 256 x 256 level
    np.savez_compressed(OUT, **rec)
    print({k: (v.shape, v.tolist() if v.size <= 10 else "") for k, v in rec.items()})
    print("foreground voxels per slice:", rec["segs"].reshape(T, -1).sum(1).tolist())
    print("low-res logit range:", float(rec["low_fwd"].min()), float(rec["low_fwd"].max()))


if __name__ == "__main__":
    main()
