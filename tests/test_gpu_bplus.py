"""sam2.1_hiera_base_plus at 1024^2 (BASELINE configs[4]) against outputs of the unmodified reference
(tests/golden/bplus1024_ct_bidirectional.npz, made by oracle/make_golden_bplus.py): the image encoder's three feature
levels, and the 3-D CT driver's sequence (box prompt on a key slice, forward pass, reset_state, reverse pass, union) on
512 x 512 slices -- 4096 queries, up to 5 * 4096 + 24 keys, 256^2 low-res masks.  Same bars as the tiny-512 fixtures."""
import os

import numpy as np
import pytest
import torch

from us_video_medsam2_b200 import synth

pytestmark = pytest.mark.gpu
DICE_BAR = 0.995  # BASELINE.json north_star


def dice(a, b):
    a, b = a > 0, b > 0
    den = a.sum().item() + b.sum().item()
    return 1.0 if den == 0 else 2.0 * (a & b).sum().item() / den


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(os.path.dirname(__file__), "golden", "bplus1024_ct_bidirectional.npz"))


@pytest.fixture(scope="module")
def predictor():
    from oracle.make_golden_bplus import SEED
    from sam2.build_sam import build_sam2_video_predictor_npz

    pred = build_sam2_video_predictor_npz("configs/sam2.1_hiera_b+.yaml", None, device="cuda", encoder_batch=2)
    pred.load_state_dict(synth.make_bplus_state_dict(SEED), strict=True)
    return pred


def test_builder_selects_the_architecture(predictor):
    assert predictor.image_size == 1024 and predictor.cfg.feat == 64
    sd = predictor.state_dict()
    assert len(sd) == 615 and sum(v.numel() for v in sd.values()) == 80850434


def test_image_encoder_matches_reference(predictor, golden):
    """Hiera-B+ trunk (24 blocks; heads of 56 padded to 64, the 112-channel stage to 128) + FpnNeck + conv_s0 / conv_s1."""
    from oracle.make_golden_bplus import KEY, T

    clip = synth.make_clip(T, size=1024, kind="speckle")
    eng = predictor._sync_engine()
    with torch.inference_mode():
        out = eng.encode_frames(clip[KEY:KEY + 2].cuda())
        one = eng.encode_frames(clip[KEY + 1:KEY + 2].cuda())
    torch.cuda.synchronize()
    assert out["feat"].shape == (2, 4096, 256) and out["feat_s1"].shape == (2, 16384, 64)
    assert out["feat_s0"].shape == (2, 65536, 32)
    for name, lvl, side, stride in (("enc_feat", "feat", 64, 2), ("enc_feat_s1", "feat_s1", 128, 4),
                                    ("enc_feat_s0", "feat_s0", 256, 8)):
        want = torch.from_numpy(golden[name])
        got = out[lvl][0].float().cpu().t().reshape(-1, side, side)[:, ::stride, ::stride]
        rel_max = ((got - want).abs().max() / want.abs().max()).item()
        rel_mean = ((got - want).abs().mean() / want.abs().mean()).item()
        # bf16 tensor-core contractions through 24 blocks (the tiny-512 bar, tests/test_gpu_modules.py: 4e-2 / 1e-2)
        assert rel_max < 4e-2 and rel_mean < 1e-2, (name, rel_max, rel_mean)
    assert torch.equal(one["feat"][0], out["feat"][1])  # frame-parallel encoder: batching does not change a frame


def test_ct_driver_flow_matches_reference_fixture(predictor, golden):
    from oracle.make_golden_bplus import T, bplus_session

    g = golden
    got = bplus_session(predictor, synth.make_clip(T, size=1024, kind="speckle").cuda())
    for k in ("frames_fwd", "frames_rev"):
        assert got[k].tolist() == g[k].tolist()
    for name in ("fwd", "rev"):
        assert np.abs(got[f"score_{name}"] - g[f"score_{name}"]).max() < 5e-3, (got[f"score_{name}"], g[f"score_{name}"])
        for key in (f"low_{name}", f"logits_{name}_s2"):
            a, b = torch.from_numpy(got[key]), torch.from_numpy(g[key])
            for i in range(a.shape[0]):
                assert dice(a[i], b[i]) >= DICE_BAR, (key, i, dice(a[i], b[i]))
                assert float((a[i] - b[i]).abs().mean()) <= 1.5e-3, (key, i, float((a[i] - b[i]).abs().mean()))
        pa, pb = got[f"obj_ptr_{name}"], g[f"obj_ptr_{name}"]
        assert np.abs(pa - pb).max() <= 2e-2 * np.abs(pb).max(), (name, np.abs(pa - pb).max(), np.abs(pb).max())
        a, b = torch.from_numpy(got[f"prompt_{name}_s2"]), torch.from_numpy(g[f"prompt_{name}_s2"])
        assert dice(a, b) >= DICE_BAR
    mm_a, mm_b = got["maskmem_last"], g["maskmem_last"]
    assert np.quantile(np.abs(mm_a - mm_b), 0.999) < 0.25 * np.abs(mm_b).max()
    seg_a, seg_b = got["segs"].astype(bool), g["segs"].astype(bool)
    for t in range(T):
        d = 2.0 * (seg_a[t] & seg_b[t]).sum() / max(1, seg_a[t].sum() + seg_b[t].sum())
        assert d >= DICE_BAR, (t, d)


def test_two_objects_and_lockstep_sessions_equal_separate_runs(predictor):
    """Size-generic batching at 4096 queries: two objects in one session (B = 2: other split factors and kernels than
    B = 1) and two sessions tracked in lock-step (propagate_in_videos) give each object the masks it gets on its own, up
    to the bf16 noise of a different reduction order."""
    T = 5
    clip = synth.make_clip(T, size=1024, kind="speckle").cuda()
    boxes = [np.array([190.0, 170.0, 340.0, 330.0], np.float32), np.array([60.0, 250.0, 200.0, 420.0], np.float32)]

    def run(obj_boxes):
        st = predictor.init_state(clip, 512, 512)
        for i, bx in enumerate(obj_boxes):
            predictor.add_new_points_or_box(st, 0, i + 1, box=bx)
        return st, torch.stack([lg[:, 0].float().clone() for _, _, lg in predictor.propagate_in_video(st)])  # [T, B, H, W]

    with torch.inference_mode():
        alone = [run([bx])[1] for bx in boxes]
        _, both = run(boxes)
        for j in range(2):
            for t in range(T):
                assert dice(both[t, j], alone[j][t, 0]) >= 0.999, (j, t)
            assert float((both[:, j] - alone[j][:, 0]).abs().mean()) < 1e-3
        states = []
        for bx in boxes:
            st = predictor.init_state(clip, 512, 512)
            predictor.add_new_points_or_box(st, 0, 1, box=bx)
            states.append(st)
        lock = [[], []]
        for t, ids, masks in predictor.propagate_in_videos(states):
            for j, lg in enumerate(masks):
                lock[j].append(lg[:, 0].float().clone())
        for j in range(2):
            got = torch.stack(lock[j])
            assert got.shape == alone[j].shape
            for t in range(T):
                assert dice(got[t, 0], alone[j][t, 0]) >= 0.999, (j, t)
