"""Round-2 end-to-end parity on the configurations that are benchmarked: every fixture below is the output of the
UNMODIFIED reference (oracle/make_golden_r2.py, oracle/make_golden_etam.py), the CUDA path is held to it at the stated
bars (tests/test_gpu_e2e.py: Dice >= 0.995 per tracked frame and object, max |dlogit| <= 8e-3 off the hole-filling
threshold).

  * 4 objects (BASELINE configs[2] shape): the engine changes attention kernel and split factors above 2 objects;
  * the benched clip itself: 512 frames forward (and 128 in reverse from a click), sampled along the pass -- bf16 error must
    not build up through the memory bank and the pointer horizon;
  * non_overlap_masks / non_overlap_masks_for_mem_enc;
  * EfficientTAM objects prompted on different frames (per-object conditioning state)."""
import os

import numpy as np
import pytest
import torch

from tests.golden_cases import dice
from us_video_medsam2_b200 import synth

pytestmark = pytest.mark.gpu
LOGIT_TOL = 8e-3
DICE_BAR = 0.995


def _predictor(seed, **kw):
    from sam2.build_sam import build_sam2_video_predictor_npz

    p = build_sam2_video_predictor_npz("configs/sam2.1_hiera_t512.yaml", **kw)
    p.load_state_dict(synth.make_state_dict(seed), strict=True)
    return p


def _check_low(got, want, tag, prompt_exact=False):
    """got / want: [H, W] low-res logits of one object on one frame."""
    got, want = torch.as_tensor(got), torch.as_tensor(want)
    same = (got != 0.1) & (want != 0.1)  # a hole filled on one side only legitimately differs (threshold at 0)
    d = float((got - want).abs()[same].max())
    if prompt_exact:
        assert d < 1e-4, (tag, d)
        return 1.0, d
    dc = dice(got, want)
    assert d <= LOGIT_TOL, (tag, d)
    assert dc >= DICE_BAR, (tag, dc)
    return dc, d


def test_four_objects_match_reference_fixture(golden_dir):
    from oracle.make_golden_r2 import FOUR_OBJ, four_obj_session

    g = np.load(os.path.join(golden_dir, "t512_four_obj_masks.npz"))
    pred = _predictor(FOUR_OBJ["seed"], encoder_batch=4)
    got = four_obj_session(pred, synth.make_clip(FOUR_OBJ["T"], kind="speckle").cuda())
    assert got["frames"].tolist() == g["frames"].tolist()
    assert got["low"].shape == g["low"].shape == (FOUR_OBJ["T"], 4, 128, 128)
    worst = 1.0
    for i, t in enumerate(got["frames"].tolist()):
        for o in range(4):
            dc, _ = _check_low(got["low"][i, o], g["low"][i, o], (t, o), prompt_exact=(t == 0))
            worst = min(worst, dc)
    print(f"4 objects: worst Dice {worst:.5f}")
    assert np.sign(got["score"]).tolist() == np.sign(g["score"]).tolist()
    assert np.abs(got["score"] - g["score"]).max() < 2e-2
    assert np.abs(got["obj_ptr"] - g["obj_ptr"]).max() < 5e-2
    assert np.abs(got["maskmem_last"] - g["maskmem_last"]).max() < 0.25  # bf16 memory of ~N(0, 1) features


@pytest.mark.parametrize("name,reverse", [("t512_long_fwd", False), ("t512_long_rev", True)])
def test_long_clip_matches_reference_fixture(golden_dir, name, reverse):
    """The benched configuration (512 frames, one object, forward) and a 128-frame reverse pass from a click."""
    from oracle.make_golden_r2 import LONG_FWD, LONG_REV, long_session

    cfg = LONG_REV if reverse else LONG_FWD
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    pred = _predictor(cfg["seed"], encoder_batch=16, encoder_sms=64)
    got = long_session(pred, synth.make_clip(cfg["T"], kind="speckle").cuda(), cfg, reverse)
    assert got["frames"].tolist() == g["frames"].tolist() and got["kept"].tolist() == g["kept"].tolist()
    worst = 1.0
    for i, t in enumerate(got["kept"].tolist()):
        dc, _ = _check_low(got["low"][i], g["low"][i], t, prompt_exact=(i == 0 and not reverse))
        worst = min(worst, dc)
    print(f"{name}: worst Dice over {len(got['kept'])} sampled frames {worst:.5f}")
    # every frame of the pass: object score (sign = the object-present branch) and pointer norm
    assert np.sign(got["score"]).tolist() == np.sign(g["score"]).tolist()
    assert np.abs(got["score"] - g["score"]).max() < 2e-2
    assert np.abs(got["ptr_norm"] - g["ptr_norm"]).max() < 5e-2 * max(1.0, float(g["ptr_norm"].max()))


def test_non_overlap_constraints_match_reference_fixture(golden_dir):
    from oracle.make_golden_r2 import NON_OVERLAP, non_overlap_session

    g = np.load(os.path.join(golden_dir, "t512_non_overlap.npz"))
    pred = _predictor(NON_OVERLAP["seed"], non_overlap_masks=True, non_overlap_masks_for_mem_enc=True)
    assert pred.non_overlap_masks and pred.non_overlap_masks_for_mem_enc
    got = non_overlap_session(pred, synth.make_clip(NON_OVERLAP["T"], kind="speckle").cuda())
    assert got["frames"].tolist() == g["frames"].tolist()
    v = got["video_s4"]
    assert int(((v[:, 0] > 0) & (v[:, 1] > 0)).sum()) == 0  # the yielded masks never overlap
    assert int(((got["low"][:, 0] > 0) & (got["low"][:, 1] > 0)).sum()) > 0  # ... although the stored ones do
    for i, t in enumerate(got["frames"].tolist()):
        for o in range(2):
            if t == 0:
                continue  # prompt frame: mask prompt exact, box prompt decoded once (covered by t512_two_obj_mask_box)
            _check_low(got["low"][i, o], g["low"][i, o], ("low", t, o))
        # the yielded masks went through the argmax over objects at video resolution: where the two objects' (upsampled)
        # logits are closer than the logit tolerance the winner is noise -- at random init the objects see nearly the same
        # evidence -- elsewhere winner and value must agree.  The pre-constraint values are re-derived from the fixture's
        # low-res logits exactly as the reference does (bilinear to the video size, sam2_video_predictor.py:404-424).
        if t > 0:
            a, b = torch.from_numpy(got["video_s4"][i]), torch.from_numpy(g["video_s4"][i])
            up = torch.nn.functional.interpolate(torch.from_numpy(g["low"][i])[:, None], size=(512, 512), mode="bilinear",
                                                 align_corners=False)[:, 0, ::4, ::4]
            # Compared where the inputs of the constraint agree: each object's logits are within LOGIT_TOL of the
            # reference's except around a hole filled on one side only (a rewrite to 0.1 at a threshold); there the winner
            # may legitimately differ, elsewhere a margin of 4 x LOGIT_TOL between the objects decides it on both sides.
            up_mine = torch.nn.functional.interpolate(torch.from_numpy(got["low"][i])[:, None], size=(512, 512),
                                                      mode="bilinear", align_corners=False)[:, 0, ::4, ::4]
            agree = ((up_mine - up).abs() <= 2 * LOGIT_TOL).all(dim=0)
            decided = ((up[0] - up[1]).abs() > 4 * LOGIT_TOL) & agree
            assert float(agree.float().mean()) > 0.9
            for o in range(2):
                clear = decided & (up[o].abs() > 2 * LOGIT_TOL)
                assert not bool((((a[o] > 0) != (b[o] > 0)) & clear).any()), ("video", t, o)
            # ... and exactly: the yielded masks are the constraint applied to this path's own upsampled stored masks
            from us_video_medsam2_b200 import ops
            mine = ops.resize_bilinear(torch.from_numpy(got["low"][i])[:, None].cuda().contiguous(), 512, 512)
            win = torch.argmax(mine, dim=0, keepdim=True)
            keep = win == torch.arange(2, device="cuda")[:, None, None, None]
            want_mine = torch.where(keep, mine, torch.clamp(mine, max=-10.0))[:, 0, ::4, ::4].cpu()
            assert torch.equal(a, want_mine), ("yield", t)
    err = np.abs(got["maskmem_last"] - g["maskmem_last"])
    # a pixel whose winner flips inside the logit tolerance changes one input of the memory encoder from its logit to the
    # -10 clamp: the memory differs around that pixel (receptive field of the down-sampler) and nowhere else -- so the bf16
    # bar of the other fixtures holds for 99.9 % of the entries and a looser one for the rest
    assert np.quantile(err, 0.999) < 0.25 and err.max() < 1.0, (float(np.quantile(err, 0.999)), float(err.max()))
    # the constraint changes the memory: without it the last memory is clearly further from the fixture's
    plain = _predictor(NON_OVERLAP["seed"])
    ref = non_overlap_session(plain, synth.make_clip(NON_OVERLAP["T"], kind="speckle").cuda())
    off = np.abs(ref["maskmem_last"] - g["maskmem_last"])
    print(f"non-overlap memory: mean |d| with the constraint {err.mean():.2e}, without {off.mean():.2e}")
    assert off.mean() > 3 * err.mean()


def test_non_overlap_kernel_matches_torch():
    from us_video_medsam2_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn((6, 1, 64, 48), generator=g, device="cuda") * 5
    x[1] = x[0]  # ties go to the lowest index (torch.argmax)

    def ref(t):
        win = torch.argmax(t, dim=0, keepdim=True)
        keep = win == torch.arange(t.size(0), device=t.device)[:, None, None, None]
        return torch.where(keep, t, torch.clamp(t, max=-10.0))

    assert torch.equal(ops.non_overlap(x), ref(x))
    grouped = ops.non_overlap(x, group=3)
    assert torch.equal(grouped[:3], ref(x[:3])) and torch.equal(grouped[3:], ref(x[3:]))
    post = ops.non_overlap(x, 0, ops.POST_SIGMOID_AFFINE, 20.0, -10.0)
    assert torch.allclose(post, torch.sigmoid(ref(x)) * 20 - 10, atol=1e-5)


def test_many_conditioning_frames_are_refused_with_a_clear_error():
    """The device control block names at most 26 conditioning frames: more must raise before tracking, not
    assert inside a kernel wrapper."""
    pred = _predictor(19)
    T = 30
    clip = synth.make_clip(T, kind="speckle").cuda()
    st = pred.init_state(clip, 512, 512)
    for t in range(12):
        pred.add_new_mask(st, t, 1, synth.box_mask())
    out = [t for t, _, _ in pred.propagate_in_video(st)]  # 12 prompted frames (was an assert above 10): fine
    assert out == list(range(T))
    pred.reset_state(st)
    for t in range(28):
        pred.add_new_mask(st, t, 1, synth.box_mask())
    with pytest.raises(RuntimeError, match="conditioning"):
        next(pred.propagate_in_video(st))


def test_etam_objects_prompted_on_different_frames(golden_dir):
    """EfficientTAM keeps conditioning frames per object (efficienttam_video_predictor.py:489-628)."""
    from efficient_track_anything.build_efficienttam import build_efficienttam_video_predictor_npz
    from oracle.make_golden_etam import SEED, T, run_diff_frames

    g = np.load(os.path.join(golden_dir, "etam_ti_diff_frames.npz"))
    pred = build_efficienttam_video_predictor_npz("configs/efficienttam_ti_512x512.yaml", device="cuda")
    pred.load_state_dict(synth.make_etam_state_dict(SEED), strict=True)
    got = run_diff_frames(pred, synth.make_clip(T, kind="speckle").cuda())
    assert got["frames"].tolist() == g["frames"].tolist() and got["obj_ids"].tolist() == g["obj_ids"].tolist()
    assert got["cond0"].tolist() == g["cond0"].tolist() == [0] and got["cond1"].tolist() == g["cond1"].tolist() == [0, 2]
    # object 1 on frame 2 is TRACKED (score ~0.14), not a +-10 placeholder of a conditioning frame
    assert np.sign(got["score"]).tolist() == np.sign(g["score"]).tolist()
    assert np.abs(got["score"] - g["score"]).max() < 2e-2
    a, b = torch.from_numpy(got["video_s4"]), torch.from_numpy(g["video_s4"])
    for i in range(a.shape[0]):
        for o in range(a.shape[1]):
            assert dice(a[i, o], b[i, o]) >= 0.99, (i, o, dice(a[i, o], b[i, o]))
            assert float((a[i, o] - b[i, o]).abs().mean()) <= 8e-4, (i, o)


def test_etam_object_added_after_tracking_started():
    """The EfficientTAM predictor always allows new objects (reference :127-159): the frame store grows, the first
    object's results stay, and a fresh pass tracks both -- the new object alone equals a solo session of it."""
    from efficient_track_anything.build_efficienttam import build_efficienttam_video_predictor_npz
    from oracle.make_golden_etam import SEED

    T = 6
    clip = synth.make_clip(T, kind="speckle").cuda()
    pred = build_efficienttam_video_predictor_npz("configs/efficienttam_ti_512x512.yaml", device="cuda")
    pred.load_state_dict(synth.make_etam_state_dict(SEED), strict=True)
    masks = synth.multi_object_masks(2)
    st = pred.init_state(clip, 512, 512)
    pred.add_new_mask(st, 0, 1, masks[0])
    first = [lg.clone() for _, _, lg in pred.propagate_in_video(st)]
    pred.add_new_mask(st, 0, 2, masks[1])  # a new object, after tracking started
    both = [(t, lg.clone()) for t, _, lg in pred.propagate_in_video(st)]
    assert [t for t, _ in both] == list(range(T)) and both[0][1].shape == (2, 1, 512, 512)
    for t in range(T):
        assert dice(both[t][1][0].cpu(), first[t][0].cpu()) >= DICE_BAR, t  # object 1 unchanged
    solo = pred.init_state(clip, 512, 512)
    pred.add_new_mask(solo, 0, 2, masks[1])
    want = {t: lg.clone() for t, _, lg in pred.propagate_in_video(solo)}
    for t in range(T):
        assert dice(both[t][1][1].cpu(), want[t][0].cpu()) >= DICE_BAR, t


def test_lockstep_sessions_match_their_solo_runs():
    """propagate_in_videos: S sessions x Bo objects as ONE batched frame graph (per-video features indexed by
    object // Bo, one shared frame store).  Every session must reproduce its own solo propagate_in_video run at the parity
    bars (different batch size -> different attention kernel / split factors, so not bitwise), and its inference_state
    must be left as a normal, independently usable session."""
    T, S = 20, 3
    clips = [synth.make_clip(T, kind="speckle", seed=1234 + i).cuda() for i in range(S)]
    masks = synth.multi_object_masks(4)
    prompts = [(masks[0], masks[1]), (masks[2], masks[3]), (masks[1], masks[2])]
    pred = _predictor(19, encoder_batch=6)

    def new_state(i):
        st = pred.init_state(clips[i], 512, 512)
        for j, m in enumerate(prompts[i]):
            pred.add_new_mask(st, 0, 10 * i + j, m)
        return st

    solo = []
    for i in range(S):
        solo.append([lg.clone() for _, _, lg in pred.propagate_in_video(new_state(i))])
    states = [new_state(i) for i in range(S)]
    seen = []
    for t, ids, lgs in pred.propagate_in_videos(states):
        assert ids == [[10 * i, 10 * i + 1] for i in range(S)] and len(lgs) == S
        seen.append(t)
        for i in range(S):
            a, b = lgs[i].float().cpu(), solo[i][t].float().cpu()
            assert a.shape == b.shape == (2, 1, 512, 512)
            for o in range(2):
                if t == 0:
                    assert torch.equal(a[o], b[o])
                else:
                    assert dice(a[o], b[o]) >= DICE_BAR, (t, i, o, dice(a[o], b[o]))
    assert seen == list(range(T))
    assert any(k[0] == 2 * S and k[-1] == 2 for k in pred._graphs if isinstance(k[0], int))  # the batched graph was captured
    # each state is an ordinary session afterwards: stored entries are views of the shared store with reference shapes
    out = states[1]["output_dict"]["non_cond_frame_outputs"][T - 1]
    assert out["maskmem_features"].shape == (2, 64, 32, 32) and out["obj_ptr"].shape == (2, 256)
    assert states[1]["output_dict_per_obj"][1]["non_cond_frame_outputs"][5]["pred_masks"].shape == (1, 1, 128, 128)
    # ... a reverse lock-step pass over the same sessions, from fresh prompts on the last frame
    for st in states:
        pred.reset_state(st)
    for i, st in enumerate(states):
        pred.add_new_mask(st, T - 1, 7, prompts[i][0])
    rev = [t for t, _, _ in pred.propagate_in_videos(states, reverse=True, max_frame_num_to_track=6)]
    assert rev == list(range(T - 1, T - 8, -1))
    # ... and sessions that do not line up are refused
    other = pred.init_state(clips[0][:10], 512, 512)
    pred.add_new_mask(other, 0, 1, masks[0])
    with pytest.raises(ValueError, match="lock-step"):
        next(pred.propagate_in_videos([new_state(0), other]))
