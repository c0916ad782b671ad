"""TEST INFRASTRUCTURE -- fixture for the bidirectional 3-D driver flow (medsam2_infer_3D_CT.py:256-283):
box prompt on a key slice, forward propagation, reset_state, the same box again, reverse propagation, union of the two
passes -- on a non-square volume (slices resized to 512 x 512 for the model, masks returned at slice resolution).

    python oracle/make_golden_ct.py        # -> tests/golden/t512_ct_bidirectional.npz   (REAL reference, CPU)
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

SEED, T, KEY = 19, 9, 4
H, W = 300, 420
BOX = np.array([150.0, 90.0, 290.0, 210.0], np.float32)  # x0, y0, x1, y1 in slice pixels
OUT = os.path.join(ROOT, "tests", "golden", "t512_ct_bidirectional.npz")


def ct_session(pred, clip):
    """The driver's sequence on any predictor with the reference API; returns per-pass frame order, slice-resolution
    logits at stride 2 and the union segmentation [T, H, W] uint8."""
    segs = np.zeros((T, H, W), np.uint8)
    rec = {}
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=torch.cuda.is_available()):
        st = pred.init_state(clip, H, W)
        for name, kw in (("fwd", {}), ("rev", dict(reverse=True))):
            _, ids, lg = pred.add_new_points_or_box(st, KEY, 1, box=BOX)
            rec[f"prompt_{name}_s2"] = lg[0, 0, ::2, ::2].float().cpu().numpy().copy()
            frames, logits = [], []
            for t, ids, lg in pred.propagate_in_video(st, **kw):
                frames.append(t)
                logits.append(lg[0, 0, ::2, ::2].float().cpu().numpy().copy())
                segs[t, (lg[0] > 0.0).cpu().numpy()[0]] = 1
            rec[f"frames_{name}"] = np.array(frames, np.int32)
            rec[f"logits_{name}_s2"] = np.stack(logits)
            pred.reset_state(st)
    rec["segs"] = segs
    return rec


def main():
    from oracle.cc_ref import connected_components_ref
    from oracle.ref_loader import load_reference_predictor
    from us_video_medsam2_b200 import synth

    model = load_reference_predictor()
    import sam2.utils.misc as misc

    def patched(mask):
        lab, cnt = connected_components_ref(mask.to(torch.uint8).cpu().numpy())
        return torch.from_numpy(lab), torch.from_numpy(cnt)

    misc.get_connected_components = patched
    model.load_state_dict(synth.make_state_dict(SEED), strict=True)
    with torch.inference_mode():
        rec = ct_session(model, synth.make_clip(T, kind="speckle"))
    np.savez_compressed(OUT, **rec)
    print({k: (v.shape, v.tolist() if v.size <= 10 else "") for k, v in rec.items()})
    print("foreground voxels per slice:", rec["segs"].reshape(T, -1).sum(1).tolist())


if __name__ == "__main__":
    main()
