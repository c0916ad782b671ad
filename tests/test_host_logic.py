"""CPU-side tests: C-ABI library loads and exports every declared symbol, header/binding agreement, builder and
state-dict ABI behaviour, memory-bank selection logic, video sharding under a world_size-2 gloo group."""
import ctypes
import os
import re

import pytest
import torch

from us_video_medsam2_b200 import _lib, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_and_exports_declared_symbols():
    path = _lib.build()
    handle = ctypes.CDLL(path)
    header = open(os.path.join(ROOT, "include", "usvm2_b200.h")).read()
    declared = set(re.findall(r"^int\s+(usvm_\w+)\(", header, flags=re.M))
    assert declared == set(_lib.EXPORTED_SYMBOLS), declared ^ set(_lib.EXPORTED_SYMBOLS)
    for name in declared:
        assert hasattr(handle, name), name
    assert handle.usvm_abi_version() == 1


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libusvm2_b200.so")
    with pytest.raises(_lib.KernelLibraryError):
        _lib.lib()


def test_state_dict_abi_and_strict_checkpoint(tmp_path):
    from us_video_medsam2_b200.build_sam import _load_checkpoint
    from us_video_medsam2_b200.predictor import SAM2VideoPredictorNPZ

    m = SAM2VideoPredictorNPZ(fill_hole_area=8)
    sd = synth.make_state_dict(3)
    assert list(m.state_dict().keys()) == list(sd.keys()) and len(sd) == 471
    assert sum(v.numel() for v in sd.values()) == 38962754
    ck = tmp_path / "ck.pt"
    torch.save({"model": sd}, ck)
    _load_checkpoint(m, str(ck))
    assert torch.equal(m.state_dict()["image_encoder.trunk.pos_embed"], sd["image_encoder.trunk.pos_embed"])
    bad = dict(sd)
    bad.pop("no_obj_ptr")
    torch.save({"model": bad}, ck)
    with pytest.raises(RuntimeError):
        _load_checkpoint(m, str(ck))
    with pytest.raises(RuntimeError):  # no CPU path: the product must refuse, not fall back
        m.init_state(synth.make_clip(1), 512, 512)


def test_builder_overrides_and_errors():
    from us_video_medsam2_b200.build_sam import _load_model_kwargs, build_sam2_video_predictor

    kw = _load_model_kwargs("configs/sam2.1_hiera_t512.yaml",
                            ["++model.fill_hole_area=8", "++model.binarize_mask_from_pts_for_mem_enc=true",
                             "++model.sam_mask_decoder_extra_args.dynamic_multimask_stability_delta=0.05",
                             "++model._target_=sam2.sam2_video_predictor.SAM2VideoPredictor"])
    assert kw["fill_hole_area"] == 8 and kw["binarize_mask_from_pts_for_mem_enc"] is True
    assert kw["sam_mask_decoder_extra_args"]["dynamic_multimask_stability_delta"] == 0.05
    with pytest.raises(FileNotFoundError):
        _load_model_kwargs("configs/sam2.1_hiera_b+.yaml", [])
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            build_sam2_video_predictor("configs/sam2.1_hiera_t512.yaml")


def test_hiera_plan_and_cond_frame_selection():
    from oracle.medsam2_ref import hiera_block_plan, select_closest_cond_frames
    from us_video_medsam2_b200.engine import hiera_plan
    from us_video_medsam2_b200.predictor import _select_closest_cond_frames

    plan = hiera_plan()
    assert [p[3] for p in plan] == [8, 8, 4, 4, 14, 0, 14, 0, 14, 0, 14, 7]
    assert [p[1] for p in plan] == [96, 192, 192, 384, 384, 384, 384, 384, 384, 384, 768, 768]
    assert [i for i, p in enumerate(plan) if p[4]] == [1, 3, 10] and [i for i, p in enumerate(plan) if p[5]] == [0, 2, 9, 11]
    for a, b in zip(plan, hiera_block_plan()):
        assert a == (b["dim"], b["dim_out"], b["heads"], b["window"], b["pool"], b["emit"])
    cond = {t: t for t in (0, 5, 9, 20, 31)}
    for frame in (1, 9, 15, 40):
        for k in (-1, 2, 3, 5):
            assert _select_closest_cond_frames(frame, cond, k) == select_closest_cond_frames(frame, cond, k)


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist

    from us_video_medsam2_b200.sharding import gather_per_video, shard_videos

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    mine = shard_videos(7, rank, world)
    local = {i: torch.tensor([float(i), float(i * i)]) for i in mine}
    table = gather_per_video(local, 7)
    q.put((rank, mine, table.tolist()))
    dist.destroy_process_group()


def test_video_sharding_world_size_2_gloo():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res[0][1] == [0, 2, 4, 6] and res[1][1] == [1, 3, 5]
    want = [[float(i), float(i * i)] for i in range(7)]
    assert res[0][2] == want and res[1][2] == want
