"""End-to-end parity of the drop-in predictor (CUDA) against (1) the reference's own outputs committed as
golden fixtures and (2) the CPU oracle on the same seeded weights / clips.

Tolerances (stated, from the bf16-emulation envelope in DESIGN.md): tracked-frame mask Dice >= 0.995 and
max |dlogit| <= 8e-3 (logit std at random init is ~0.04-0.07); prompt frames exact; discontinuities
(argmax-IoU choice, object-score sign) must agree, which the chosen seeds guarantee with margin."""
import os

import numpy as np
import pytest
import torch

from oracle.make_golden import CASES
from tests.golden_cases import dice, replay
from us_video_medsam2_b200 import synth

pytestmark = pytest.mark.gpu
LOGIT_TOL = 8e-3
DICE_BAR = 0.995


def _predictor(seed, **kw):
    from sam2.build_sam import build_sam2_video_predictor_npz

    p = build_sam2_video_predictor_npz("configs/sam2.1_hiera_t512.yaml", **kw)
    p.load_state_dict(synth.make_state_dict(seed), strict=True)
    return p


@pytest.mark.parametrize("name", list(CASES))
def test_matches_reference_fixture(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    cfg = CASES[name]
    pred = _predictor(cfg["seed"], apply_postprocessing=cfg.get("post", True))
    out = replay(pred, name, images=synth.make_clip(cfg["T"], kind="speckle").cuda())
    assert out["frames"] == g["frames"].tolist()
    want_plain = torch.from_numpy(g["low_res"])
    want = torch.from_numpy(g["low_res_filled"])
    got = out["low"]
    prompted = {t for _, t, _, _ in cfg["prompts"]}
    for i, t in enumerate(out["frames"]):
        # compare the *unfilled* logits where neither side filled a hole (hole filling thresholds at 0 and
        # rewrites to 0.1, so a pixel that flips side legitimately differs)
        same = (got[i] != 0.1) & (want[i] != 0.1)
        d = (got[i] - want[i]).abs()[same]
        if t in prompted and cfg["prompts"][0][0] == "mask" and len(cfg["prompts"]) == 1:
            assert float(d.max()) < 1e-4
            continue
        assert float(d.max()) <= LOGIT_TOL, (t, float(d.max()))
        assert dice(got[i], want[i]) >= DICE_BAR, (t, dice(got[i], want[i]))
        # vs the stock CPU reference (which skips hole filling): identical wherever no hole was filled
        unfilled = (got[i] != 0.1)
        assert float((got[i] - want_plain[i]).abs()[unfilled & same].max()) <= LOGIT_TOL
    sfx = "_filled"
    assert np.abs(out["score"].numpy() - g["score" + sfx]).max() < 2e-2
    assert np.sign(out["score"].numpy()).tolist() == np.sign(g["score" + sfx]).tolist()
    assert np.abs(out["obj_ptr"].numpy() - g["obj_ptr" + sfx]).max() < 5e-2


@pytest.mark.parametrize("T", [16, 48])
def test_matches_oracle_on_longer_clip(T):
    """16-frame clip (BASELINE config 1 shape): memory bank fills up (7 frames) and pointers reach 15; 48 frames: three
    times the pointer horizon, the bf16 error must not build up through the memory bank."""
    from oracle.medsam2_ref import RefPredictor

    seed = 19
    clip = synth.make_clip(T, kind="speckle")
    sd = synth.make_state_dict(seed)
    ref = RefPredictor(sd, fill_holes=True)
    with torch.inference_mode():
        st = ref.init_state(clip, 512, 512)
        ref.add_new_mask(st, 0, 1, synth.box_mask())
        want = [lg.clone() for _, _, lg in ref.propagate_in_video(st)]
    pred = _predictor(seed, encoder_batch=4)
    st2 = pred.init_state(clip.cuda(), 512, 512)
    pred.add_new_mask(st2, 0, 1, synth.box_mask())
    got = [lg.float().cpu() for _, _, lg in pred.propagate_in_video(st2)]
    assert len(got) == T
    assert float((got[0] - want[0]).abs().max()) < 1e-4
    worst = 1.0
    for t in range(1, T):
        same = (got[t] != 0.1) & (want[t] != 0.1)
        assert got[t].shape == (1, 1, 512, 512)
        worst = min(worst, dice(got[t], want[t]))
    print(f"T={T}: worst Dice over tracked frames {worst:.5f}")
    assert worst >= DICE_BAR, worst


def test_video_resolution_resize_and_state_api():
    """Non-512 video size: outputs are resized to (H, W); reset_state + re-prompt works on the same state."""
    pred = _predictor(19)
    clip = synth.make_clip(3, kind="speckle").cuda()
    st = pred.init_state(clip, 360, 480)
    mask = torch.zeros((360, 480), dtype=torch.bool)
    mask[100:200, 150:300] = True
    t, ids, lg = pred.add_new_mask(st, 0, 5, mask)
    assert ids == [5] and lg.shape == (1, 1, 360, 480) and lg.dtype == torch.float32 and lg.is_cuda
    outs = list(pred.propagate_in_video(st))
    assert [o[0] for o in outs] == [0, 1, 2] and outs[1][2].shape == (1, 1, 360, 480)
    with pytest.raises(RuntimeError):
        pred.add_new_mask(st, 0, 6, mask)  # new object after tracking started
    out = st["output_dict"]["non_cond_frame_outputs"][1]
    assert out["maskmem_features"].shape == (1, 64, 32, 32) and out["maskmem_features"].dtype == torch.bfloat16
    assert out["pred_masks"].shape == (1, 1, 128, 128) and out["obj_ptr"].shape == (1, 256)
    pred.reset_state(st)
    assert st["obj_ids"] == [] and not st["tracking_has_started"]
    with pytest.raises(RuntimeError):
        next(pred.propagate_in_video(st))  # "No points are provided"
    pred.reset_state(st)  # like the reference, the failed call above already marked tracking as started
    pred.add_new_points_or_box(st, 2, 1, box=np.array([100, 80, 300, 250], np.float32))
    rev = [o[0] for o in pred.propagate_in_video(st, reverse=True)]
    assert rev == [2, 1, 0]
    with pytest.raises(ValueError):
        pred.add_new_points_or_box(st, 0, 1, points=np.zeros((1, 2), np.float32))


def test_cuda_graph_replay_equals_eager():
    """Steady-state frames replayed from the captured CUDA graph must reproduce the eager kernel sequence bit for
    bit (same kernels, same order), for 2 objects on a clip long enough to reach the steady state (frame >= 16)."""
    T = 26
    clip = synth.make_clip(T, kind="speckle").cuda()
    outs = []
    for graphs in (False, True):
        pred = _predictor(19, use_cuda_graphs=graphs, encoder_batch=4)
        st = pred.init_state(clip, 512, 512)
        for i, m in enumerate(synth.multi_object_masks(2)):
            pred.add_new_mask(st, 0, i + 1, m)
        outs.append([lg.clone() for _, _, lg in pred.propagate_in_video(st)])
        if graphs:
            assert len(pred._graphs) >= 1  # the steady state was captured
            ptr = st["output_dict"]["non_cond_frame_outputs"][T - 1]["obj_ptr"].clone()
    for a, b in zip(*outs):
        assert torch.equal(a, b)
    assert ptr.shape == (2, 256) and bool(torch.isfinite(ptr).all())


def test_encoder_on_sm_partition_matches_alternating_schedule():
    """Look-ahead encoder batches on a green-context SM partition, concurrent with the tracked frames: the encoder's
    features must be bit-identical to the alternating schedule (same kernels, only the persistent grid differs); the
    tracked frames use a different split-KV factor (fewer SMs), so masks agree to rounding, not bitwise."""
    T = 41
    clip = synth.make_clip(T, kind="speckle").cuda()
    outs, feats = [], []
    for sms in (0, 56):
        pred = _predictor(19, encoder_batch=8, encoder_sms=sms)
        st = pred.init_state(clip, 512, 512)
        pred.add_new_mask(st, 0, 1, synth.box_mask())
        seq = []
        for t, _, lg in pred.propagate_in_video(st):
            od = st["output_dict"]
            out = od["cond_frame_outputs"].get(t) or od["non_cond_frame_outputs"][t]
            seq.append(out["pred_masks"].float().cpu().clone())
            if t == 20:
                f = pred._get_image_feature(st, 20, lookahead=1)
                feats.append({k: f[k].clone() for k in ("feat", "feat_bf16", "feat_s0", "feat_s1")})
        outs.append(seq)
        if sms:
            if pred._partition_error is not None:
                pytest.skip(f"no SM partition on this driver: {pred._partition_error}")
            assert pred._partition_obj is not None and 0 < pred._partition_obj.sms < pred._partition_obj.total_sms
            assert any(k[0] == "encoder" and k[3] for k in pred._graphs)  # the encoder graph was captured on the partition
    for k in feats[0]:
        assert torch.equal(feats[0][k], feats[1][k]), k
    assert len(outs[0]) == len(outs[1]) == T
    for a, b in zip(*outs):
        assert dice(a, b) >= DICE_BAR  # two bf16 evaluation orders of the same frame: the parity bar applies
        same = (a != 0.1) & (b != 0.1)  # a hole filled on one side only legitimately differs (threshold at 0)
        assert float((a - b).abs()[same].max()) <= LOGIT_TOL


def test_partition_pipeline_reverse_multi_object_and_early_exit():
    """The look-ahead pipeline under the less common drivers of propagate_in_video: reverse tracking from a point prompt
    in the middle of the clip, two objects, a ragged tail batch, max_frame_num_to_track, and a generator abandoned half
    way followed by a fresh pass on the same predictor.  Reference behaviour = the alternating schedule (encoder_sms=0)."""
    T = 30
    clip = synth.make_clip(T, kind="speckle").cuda()

    def run(sms):
        pred = _predictor(19, encoder_batch=4, encoder_sms=sms)
        res = {}
        st = pred.init_state(clip, 512, 512)
        for i, m in enumerate(synth.multi_object_masks(2)):
            pred.add_new_mask(st, T - 1, i + 1, m)
        res["rev"] = [(t, lg.clone()) for t, _, lg in pred.propagate_in_video(st, reverse=True)]
        pred.reset_state(st)
        pred.add_new_points_or_box(st, 3, 9, points=np.array([[256.0, 250.0]], np.float32), labels=np.array([1], np.int32))
        res["fwd_limited"] = [(t, lg.clone()) for t, _, lg in pred.propagate_in_video(st, max_frame_num_to_track=14)]
        pred.reset_state(st)
        pred.add_new_mask(st, 0, 1, synth.box_mask())
        gen = pred.propagate_in_video(st)
        for _ in range(7):
            next(gen)
        gen.close()  # abandoned: the pipeline must be torn down without leaving work that corrupts the next pass
        assert "_pipeline" not in st
        pred.reset_state(st)
        pred.add_new_mask(st, 0, 1, synth.box_mask())
        res["fwd"] = [(t, lg.clone()) for t, _, lg in pred.propagate_in_video(st)]
        return res, pred

    base, _ = run(0)
    got, pred = run(48)
    if pred._partition_error is not None:
        pytest.skip(f"no SM partition on this driver: {pred._partition_error}")
    assert [t for t, _ in base["rev"]] == list(range(T - 1, -1, -1)) == [t for t, _ in got["rev"]]
    assert [t for t, _ in base["fwd_limited"]] == list(range(3, 18)) == [t for t, _ in got["fwd_limited"]]
    assert [t for t, _ in got["fwd"]] == list(range(T))
    for key in base:
        for (t, a), (_, b) in zip(base[key], got[key]):
            assert a.shape == b.shape
            assert dice(a.cpu(), b.cpu()) >= DICE_BAR, (key, t)


def test_editing_session_matches_reference_fixture(golden_dir):
    """(a16) interactive surface against the reference's own outputs: correction click on a tracked frame (previous mask
    logits + memory-conditioned features), re-propagation, clear_all_prompts_in_frame, remove_object -- same script as
    oracle/make_golden_edit.py ran on the unmodified reference."""
    from oracle.make_golden_edit import SEED, T, edit_session

    g = np.load(os.path.join(golden_dir, "t512_edit_session.npz"))
    pred = _predictor(SEED)
    got = edit_session(pred, synth.make_clip(T, kind="speckle").cuda(), synth.box_mask())
    for key in ("frames_a", "frames_b", "frames_c", "click_ids", "cond_b", "cond_after_clear", "remove_ids",
                "remove_updated_frames"):
        assert got[key].tolist() == g[key].tolist(), key
    # Tolerances: the bars of every tracked frame (Dice >= 0.995, mean |dlogit| <= 8e-4); the frame that received the
    # click has steeper logits around the point: max |dlogit| bar 4 x LOGIT_TOL there.
    worst = (1.0, ("none", -1))
    for key in ("low_a", "low_b", "low_c", "click_video_s4", "clear_video_s4", "remove_video_s4"):
        a, b = torch.from_numpy(got[key]), torch.from_numpy(g[key])
        assert a.shape == b.shape, key
        a, b = a.reshape(-1, *a.shape[-2:]), b.reshape(-1, *b.shape[-2:])
        for i in range(a.shape[0]):
            same = (a[i] != 0.1) & (b[i] != 0.1)  # hole filling rewrites to 0.1 at a threshold
            d = (a[i] - b[i]).abs()
            if key.startswith("low"):  # (the video-resolution samples interpolate across filled pixels)
                # (around the click the logits reach several units: the absolute bar plus bf16-level relative error)
                # -- relative to the local logit scale (3 x 3 neighbourhood): at the steep edge of the clicked region a
                # pixel's own value can be near zero while its neighbours are several units
                scale = torch.nn.functional.max_pool2d(b[i].abs()[None, None], 3, 1, 1)[0, 0]
                assert bool((d[same] <= 4 * LOGIT_TOL + 4e-3 * scale[same]).all()), (key, i, float(d[same].max()))
            assert float(d[same].mean()) <= 8e-4, (key, i, float(d[same].mean()))
            worst = min(worst, (dice(a[i], b[i]), (key, i)))
    print(f"editing session: worst Dice {worst[0]:.5f} at {worst[1]}")
    assert worst[0] >= DICE_BAR, worst


def test_autocast_and_half_precision_frames_are_accepted():
    """Drivers of the reference wrap the calls in torch.autocast(cuda, bfloat16) and hand over frames in any float dtype
    (medsam2_infer_3D_CT.py:256, sam2_video_predictor_npz.py:44-54): both must work and must not change the arithmetic of
    the path (the kernels take explicit dtypes; bf16 frames are widened on the fly)."""
    T = 12
    clip = synth.make_clip(T, kind="speckle").cuda()
    pred = _predictor(19, encoder_batch=4)

    def run(images, autocast):
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            st = pred.init_state(images, 512, 512)
            pred.add_new_mask(st, 0, 1, synth.box_mask())
            return [lg.clone() for _, _, lg in pred.propagate_in_video(st)]

    base = run(clip, False)
    cast = run(clip, True)
    assert all(lg.dtype == torch.float32 for lg in cast)
    for a, b in zip(base, cast):
        assert torch.equal(a, b)
    half = run(clip.to(torch.bfloat16), True)   # frames rounded to bf16 by the caller: same path, rounded input
    assert len(half) == T and half[1].shape == (1, 1, 512, 512)
    assert dice(half[-1].cpu(), base[-1].cpu()) > 0.97


def test_bidirectional_ct_driver_flow_matches_reference_fixture(golden_dir):
    """ the 3-D driver's sequence -- box prompt on a key slice under autocast, forward pass, reset_state, the
    same box, reverse pass, union -- on a 300 x 420 volume, against the reference's own outputs."""
    from oracle.make_golden_ct import SEED, T, ct_session

    g = np.load(os.path.join(golden_dir, "t512_ct_bidirectional.npz"))
    pred = _predictor(SEED, encoder_batch=4)
    got = ct_session(pred, synth.make_clip(T, kind="speckle").cuda())
    for k in ("frames_fwd", "frames_rev"):
        assert got[k].tolist() == g[k].tolist()
    for k in ("logits_fwd_s2", "logits_rev_s2"):
        a, b = torch.from_numpy(got[k]), torch.from_numpy(g[k])
        for i in range(a.shape[0]):
            assert dice(a[i], b[i]) >= DICE_BAR, (k, i, dice(a[i], b[i]))
            assert float((a[i] - b[i]).abs().mean()) <= 8e-4, (k, i)
    seg_a, seg_b = got["segs"].astype(bool), g["segs"].astype(bool)
    for t in range(T):
        d = 2.0 * (seg_a[t] & seg_b[t]).sum() / max(1, seg_a[t].sum() + seg_b[t].sum())
        assert d >= DICE_BAR, (t, d)


def test_image_predictor_matches_reference_fixture(golden_dir):
    """ SAM2ImagePredictor.set_image / predict -- a click with 3 multimask outputs, box + negative click with
    one output (stability fallback), refinement from a previous low-res mask -- against the reference's own outputs.
    Same noise floor as a tracked frame (bf16 image encoder, fp32-operand decoder); bars stated below."""
    from oracle.make_golden_image import SEED, image_session
    from sam2.sam2_image_predictor import SAM2ImagePredictor

    g = np.load(os.path.join(golden_dir, "t512_image_predictor.npz"))
    model = _predictor(SEED)
    pred = SAM2ImagePredictor(model, max_hole_area=8, max_sprinkle_area=4)
    with pytest.raises(RuntimeError):
        pred.predict(point_coords=np.zeros((1, 2), np.float32), point_labels=np.ones(1, np.int32))  # no image set yet
    got = image_session(pred)
    assert set(got) == set(g.files)
    worst = (1.0, ("none", -1))
    for k in g.files:
        a, b = got[k], g[k]
        assert a.shape == b.shape, (k, a.shape, b.shape)
        if k.endswith("_iou"):
            assert np.abs(a - b).max() < 2e-2, (k, a, b)
            assert int(np.argmax(a)) == int(np.argmax(b))
        elif k == "d_binary_s2":
            for i in range(a.shape[0]):
                dc = dice(torch.from_numpy(a[i].astype(np.float32)) - 0.5, torch.from_numpy(b[i].astype(np.float32)) - 0.5)
                worst = min(worst, (dc, (k, i)))
        else:
            for i in range(a.shape[0]):
                x, y = torch.from_numpy(a[i]), torch.from_numpy(b[i])
                if k.endswith("_low"):  # the decoder's own logits: the bf16 noise floor of the path
                    assert float((x - y).abs().mean()) <= 8e-4, (k, i, float((x - y).abs().mean()))
                    assert float((x - y).abs().max()) <= 4 * LOGIT_TOL, (k, i, float((x - y).abs().max()))
                # (the full-resolution masks went through hole / sprinkle rewriting to +-10 at a threshold and the resize
                # blends those spikes into their neighbours: only the binary agreement is comparable there)
                worst = min(worst, (dice(x, y), (k, i)))
    print(f"image predictor: worst Dice {worst[0]:.5f} at {worst[1]}")
    assert worst[0] >= DICE_BAR, worst


def test_interleaved_sessions_on_one_predictor():
    """Two sessions whose propagate_in_video generators are advanced alternately on ONE predictor: the look-ahead slots and
    the encoder graph's static outputs are shared state of the predictor, so the second session must neither steal the
    first one's pipeline nor leave stale cached features behind.  Each session must reproduce its own solo run."""
    T = 24
    clips = [synth.make_clip(T, kind="speckle", seed=1234 + i).cuda() for i in range(2)]
    for sms in (0, 48):
        pred = _predictor(19, encoder_batch=4, encoder_sms=sms)

        def solo(clip):
            st = pred.init_state(clip, 512, 512)
            pred.add_new_mask(st, 0, 1, synth.box_mask())
            return [lg.clone() for _, _, lg in pred.propagate_in_video(st)]

        want = [solo(c) for c in clips]
        states = []
        for c in clips:
            st = pred.init_state(c, 512, 512)
            pred.add_new_mask(st, 0, 1, synth.box_mask())
            states.append(st)
        gens = [pred.propagate_in_video(st) for st in states]
        got = [[], []]
        for _ in range(T):
            for i, gen in enumerate(gens):
                got[i].append(next(gen)[2].clone())
        for gen in gens:
            gen.close()
        assert pred._pipeline_owner is None
        for i in range(2):
            for t, (a, b) in enumerate(zip(got[i], want[i])):
                assert dice(a.cpu(), b.cpu()) >= DICE_BAR, (sms, i, t, dice(a.cpu(), b.cpu()))


def test_jpeg_folder_ingest_on_device(tmp_path):
    """ init_state(video_path): uint8 frames uploaded and normalised by usvm_normalize_rgb_u8 -- synchronous,
    asynchronous (decoder thread) and offloaded variants give the same frames, equal to the reference's arithmetic
    (sam2/utils/misc.py:92-101, 268-276) on the host-decoded pixels; and a session started from the folder tracks."""
    from PIL import Image

    from us_video_medsam2_b200.frames import VideoFrames, decode_jpeg_folder, load_video_frames

    rng = np.random.default_rng(5)
    T = 9
    for i in range(T):
        arr = (rng.random((96, 128, 3)) * 255).astype(np.uint8)
        Image.fromarray(arr).save(tmp_path / f"{i:05d}.jpg", quality=92)
    rgb, h, w = decode_jpeg_folder(str(tmp_path), 512)
    mean = torch.tensor(synth.IMG_MEAN)[:, None, None]
    std = torch.tensor(synth.IMG_STD)[:, None, None]
    want = torch.from_numpy(rgb.numpy() / 255.0).permute(0, 3, 1, 2).float()  # (/255 in float64, stored as fp32)
    want = (want - mean) / std
    dev = torch.device("cuda")
    sync, h1, w1 = load_video_frames(str(tmp_path), 512, False, compute_device=dev)
    assert torch.is_tensor(sync) and sync.shape == (T, 3, 512, 512) and sync.is_cuda and (h1, w1) == (96, 128)
    assert float((sync.cpu() - want).abs().max()) <= 2.4e-7 * 3  # same operations in fp32: at most an ulp of 2.6
    lazy, _, _ = load_video_frames(str(tmp_path), 512, False, async_loading_frames=True, compute_device=dev)
    assert isinstance(lazy, VideoFrames) and len(lazy) == T
    assert torch.equal(lazy[T - 1], sync[T - 1]) and torch.equal(lazy[2], sync[2])
    off, _, _ = load_video_frames(str(tmp_path), 512, True, compute_device=dev)
    assert isinstance(off, VideoFrames) and off.frames is None and torch.equal(off[4], sync[4])
    from sam2.build_sam import build_sam2_video_predictor

    pred = build_sam2_video_predictor("configs/sam2.1_hiera_t512.yaml", encoder_batch=4)
    pred.load_state_dict(synth.make_state_dict(19), strict=True)
    outs = []
    for kw in (dict(), dict(async_loading_frames=True), dict(offload_video_to_cpu=True)):
        st = pred.init_state(str(tmp_path), **kw)
        assert (st["video_height"], st["video_width"], st["num_frames"]) == (96, 128, T)
        m = torch.zeros((96, 128), dtype=torch.bool)
        m[30:70, 40:100] = True
        pred.add_new_mask(st, 0, 1, m)
        outs.append([lg.clone() for _, _, lg in pred.propagate_in_video(st)])
        assert len(outs[-1]) == T and outs[-1][3].shape == (1, 1, 96, 128)
    for a, b in zip(outs[0], outs[1]):
        assert torch.equal(a, b)
    for a, b in zip(outs[0], outs[2]):
        assert torch.equal(a, b)


def test_image_predictor_batch_equals_single_images():
    """SAM2ImagePredictor.set_image_batch / predict_batch (sam2_image_predictor.py:134-236): one batched encoder pass; every
    image's prediction must equal set_image + predict on that image alone (the encoder is frame-parallel: bit-identical)."""
    from sam2.sam2_image_predictor import SAM2ImagePredictor

    model = _predictor(26)
    pred = SAM2ImagePredictor(model, max_hole_area=8, max_sprinkle_area=4)
    rng = np.random.default_rng(11)
    images = [(rng.random((200 + 40 * i, 320, 3)) * 255).astype(np.uint8) for i in range(3)]
    pts = [np.array([[100.0 + 30 * i, 90.0]], np.float32) for i in range(3)]
    labs = [np.array([1], np.int32)] * 3
    boxes = [np.array([40.0, 30.0, 250.0, 180.0], np.float32)] * 3
    single = []
    for im, pc, pl, bx in zip(images, pts, labs, boxes):
        pred.set_image(im)
        single.append(pred.predict(point_coords=pc, point_labels=pl, box=bx, multimask_output=True, return_logits=True))
    pred.set_image_batch(images)
    with pytest.raises(AssertionError):
        pred.predict(point_coords=pts[0], point_labels=labs[0])
    masks, ious, lows = pred.predict_batch(pts, labs, box_batch=boxes, multimask_output=True, return_logits=True)
    assert len(masks) == len(ious) == len(lows) == 3
    for i in range(3):
        assert masks[i].shape == (3,) + images[i].shape[:2]
        assert np.array_equal(masks[i], single[i][0]) and np.array_equal(ious[i], single[i][1])
        assert np.array_equal(lows[i], single[i][2])
    assert pred.get_image_embedding().shape == (3, 256, 32, 32)


def test_mp4_ingest_on_device(tmp_path):
    """ init_state("clip.mp4"): frames decoded by OpenCV's FFmpeg backend, uploaded as uint8 and normalised on
    the device like the JPEG route; synchronous and asynchronous loading agree and a session started from the file tracks."""
    cv2 = pytest.importorskip("cv2")
    from us_video_medsam2_b200.frames import VideoFrames, _Mp4Source, load_video_frames

    path = str(tmp_path / "clip.mp4")
    wr = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"mp4v"), 10, (128, 96))
    if not wr.isOpened():
        pytest.skip("this OpenCV build cannot write mp4v")
    rng = np.random.default_rng(9)
    T = 7
    base = cv2.resize((rng.random((12, 16, 3)) * 255).astype(np.uint8), (128, 96), interpolation=cv2.INTER_CUBIC)
    for i in range(T):
        wr.write(np.roll(base, 3 * i, axis=1))
    wr.release()
    dev = torch.device("cuda")
    sync, h, w = load_video_frames(path, 512, False, compute_device=dev)
    assert torch.is_tensor(sync) and sync.shape == (T, 3, 512, 512) and (h, w) == (96, 128)
    src, host = _Mp4Source(path, 512), np.empty((512, 512, 3), np.uint8)
    src.decode_into(0, host)
    mean = torch.tensor(synth.IMG_MEAN)[:, None, None]
    std = torch.tensor(synth.IMG_STD)[:, None, None]
    want0 = (torch.from_numpy(host / 255.0).permute(2, 0, 1).float() - mean) / std
    assert float((sync[0].cpu() - want0).abs().max()) <= 2.4e-7 * 3
    lazy, _, _ = load_video_frames(path, 512, False, async_loading_frames=True, compute_device=dev)
    assert isinstance(lazy, VideoFrames) and len(lazy) == T and torch.equal(lazy[T - 1], sync[T - 1])
    from sam2.build_sam import build_sam2_video_predictor

    pred = build_sam2_video_predictor("configs/sam2.1_hiera_t512.yaml", encoder_batch=4)
    pred.load_state_dict(synth.make_state_dict(19), strict=True)
    st = pred.init_state(path)
    assert (st["video_height"], st["video_width"], st["num_frames"]) == (96, 128, T)
    m = torch.zeros((96, 128), dtype=torch.bool)
    m[30:70, 40:100] = True
    pred.add_new_mask(st, 0, 1, m)
    outs = [lg for _, _, lg in pred.propagate_in_video(st)]
    assert len(outs) == T and outs[3].shape == (1, 1, 96, 128)


def test_pipelined_frames_are_bit_identical():
    """pipeline_frames=True (the graph of frame t also computes frame t + 1's bank-independent attention prefix on a forked
    branch) changes the schedule, not the arithmetic: same logits, bit for bit, forward and reverse, and the state left
    behind is the same."""
    clip = synth.make_clip(40, kind="speckle").cuda()
    outs = []
    for flag in (False, True):
        pred = _predictor(19, encoder_batch=8)
        pred.pipeline_frames = flag
        st = pred.init_state(clip, 512, 512)
        pred.add_new_mask(st, 20, 1, synth.box_mask())
        got = {}
        for kw in ({}, dict(reverse=True)):
            for t, ids, lg in pred.propagate_in_video(st, **kw):
                got[(t, bool(kw))] = lg.clone()
        if flag:
            assert any(len(k) == 8 and k[6] == "pipe" for k in pred._graphs), "the pipelined graphs were never used"
        outs.append((got, st["_store"].mem.clone(), st["_store"].ptr.clone()))
    (a, mem_a, ptr_a), (b, mem_b, ptr_b) = outs
    assert a.keys() == b.keys() and len(a) == 41
    for k in a:
        assert torch.equal(a[k], b[k]), k
    assert torch.equal(mem_a, mem_b) and torch.equal(ptr_a, ptr_b)
