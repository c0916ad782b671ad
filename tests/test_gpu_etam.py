"""EfficientTAM-ti (BASELINE configs[3]) on the CUDA path: image encoder (ViT trunk + ViTDetNeck) against the
CPU oracle, and the drop-in predictor against the fixture the unmodified reference produced
(tests/golden/etam_ti_mask_fwd.npz).  Tolerances as for the Hiera model: bf16 contractions in the encoder / memory
attention / memory encoder, fp32-operand mask decoder; Dice >= 0.995 and |dlogit| <= 8e-3 on tracked frames."""
import os

import numpy as np
import pytest
import torch

from tests.golden_cases import dice
from us_video_medsam2_b200 import synth

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-6)).item()


def test_vit_encoder_matches_oracle():
    from oracle.etam_ref import EtamCfg, RefModelETAM
    from oracle.make_golden_etam import SEED
    from us_video_medsam2_b200.engine import Engine, EtamTiConfig, PackedWeights

    sd = synth.make_etam_state_dict(SEED)
    eng = Engine(PackedWeights(sd, torch.device("cuda"), EtamTiConfig))
    clip = synth.make_clip(2, kind="speckle")
    with torch.inference_mode():
        want = RefModelETAM(sd, EtamCfg).forward_image(clip[:1])["feat"][0]
        out = eng.encode_frames(clip.cuda())
        one = eng.encode_frames(clip[1:2].cuda())
    feat = out["feat"][0].float().cpu().t().reshape(256, 32, 32)
    assert _rel(feat, want) < 4e-2, _rel(feat, want)
    assert ((feat - want).abs().mean() / want.abs().mean()).item() < 1e-2
    assert float(out["feat_s0"].abs().max()) == 0.0 and out["feat_s1"].shape == (2, 4096, 64)
    assert torch.equal(one["feat"][0], out["feat"][1])  # frame-parallel: batched == per-frame


@pytest.mark.parametrize("variant", ["ti", "s"])
def test_predictor_matches_reference_fixture(golden_dir, variant):
    from efficient_track_anything.build_efficienttam import build_efficienttam_video_predictor_npz
    from oracle.make_golden_etam import SEED, SEED_S, T, T_S

    seed, T = (SEED, T) if variant == "ti" else (SEED_S, T_S)
    g = np.load(os.path.join(golden_dir, f"etam_{variant}_mask_fwd.npz"))
    pred = build_efficienttam_video_predictor_npz(f"configs/efficienttam_{variant}_512x512.yaml", device="cuda")
    pred.load_state_dict(synth.make_etam_state_dict(seed, variant), strict=True)
    clip = synth.make_clip(T, kind="speckle").cuda()
    st = pred.init_state(clip, 512, 512)
    pred.add_new_mask(st, 0, 1, synth.box_mask())
    frames, low = [], []
    od = st["output_dict"]
    for t, ids, lg in pred.propagate_in_video(st):
        assert ids == [1] and lg.shape == (1, 1, 512, 512)
        frames.append(t)
        out = od["cond_frame_outputs"].get(t) or od["non_cond_frame_outputs"][t]
        low.append(out["pred_masks"][:, 0].float().cpu().clone())
    assert frames == g["frames"].tolist()
    want = torch.from_numpy(g["low_res_filled"])
    get = lambda t: od["cond_frame_outputs"].get(t) or od["non_cond_frame_outputs"][t]
    assert float((low[0] - want[0]).abs().max()) < 1e-4  # the prompted frame reproduces the mask input exactly
    for i in range(1, len(frames)):
        same = (low[i] != 0.1) & (want[i] != 0.1)
        assert float((low[i] - want[i]).abs()[same].max()) <= 8e-3, (i, float((low[i] - want[i]).abs()[same].max()))
        assert dice(low[i], want[i]) >= 0.995, (i, dice(low[i], want[i]))
    score = np.stack([get(t)["object_score_logits"].float().cpu().numpy() for t in frames])
    assert np.sign(score).tolist() == np.sign(g["score_filled"]).tolist() and np.abs(score - g["score_filled"]).max() < 2e-2
    ptr = np.stack([get(t)["obj_ptr"].float().cpu().numpy() for t in frames])
    assert np.abs(ptr - g["obj_ptr_filled"]).max() < 5e-2


def test_two_objects_reverse_matches_per_object_reference(golden_dir):
    """Two objects (mask + clicks) prompted on the last frame, tracked in reverse: the batched CUDA frame against the
    reference's one-object-at-a-time EfficientTAM predictor."""
    from efficient_track_anything.build_efficienttam import build_efficienttam_video_predictor_npz
    from oracle.make_golden_etam import SEED, T, run_two_objects

    g = np.load(os.path.join(golden_dir, "etam_ti_two_obj_reverse.npz"))
    pred = build_efficienttam_video_predictor_npz("configs/efficienttam_ti_512x512.yaml", device="cuda")
    pred.load_state_dict(synth.make_etam_state_dict(SEED), strict=True)
    got = run_two_objects(pred, synth.make_clip(T, kind="speckle").cuda())
    assert got["frames"].tolist() == g["frames"].tolist() and got["obj_ids"].tolist() == g["obj_ids"].tolist()
    a, b = torch.from_numpy(got["video_s4"]), torch.from_numpy(g["video_s4"])
    assert a.shape == b.shape
    for i in range(a.shape[0]):
        for o in range(a.shape[1]):
            assert dice(a[i, o], b[i, o]) >= 0.99, (i, o, dice(a[i, o], b[i, o]))
            assert float((a[i, o] - b[i, o]).abs().mean()) <= 8e-4, (i, o)
