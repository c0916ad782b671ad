#!/usr/bin/env python
"""Tiny driver for `ncu --set full`: launches the hot kernels a few times at their propagation-path shapes."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from us_video_medsam2_b200 import ops  # noqa: E402

dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(0)
rnd = lambda *s, dt=torch.bfloat16, sc=1.0: (torch.randn(s, generator=g, device=dev) * sc).to(dt)

which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "gemm_enc"):
    a, w = rnd(131072, 96), rnd(288, 96, sc=0.1)
    b = rnd(288, dt=torch.float32)
    for _ in range(4):
        ops.gemm_bf16(a, w, bias=b, bf16=True)
if which in ("all", "gemm_ma"):
    a, w = rnd(1024, 2048), rnd(256, 2048, sc=0.02)
    r = rnd(1024, 256, dt=torch.float32)
    for _ in range(4):
        ops.gemm_bf16(a, w, residual=r, f32=True)
if which in ("all", "fmha"):
    B, T, Nk, D = 1, 1024, 7232, 256
    q, kv = rnd(B * T, D), rnd(B * Nk, 4 * D)
    for _ in range(4):
        ops.fmha(q, kv, kv, B, 1, T, Nk, D, (0, T * D, D, D), (D, Nk * 4 * D, 4 * D, D), (2 * D, Nk * 4 * D, 4 * D, D),
                 num_splits=18)
if which == "encoder":  # one batched image-encoder pass (8 frames), twice: read the second pass from the launch list
    from us_video_medsam2_b200 import synth
    from us_video_medsam2_b200.build_sam import build_sam2_video_predictor_npz
    pred = build_sam2_video_predictor_npz("configs/sam2.1_hiera_t512.yaml", device=dev, encoder_batch=8)
    pred.load_state_dict(synth.make_state_dict(19), strict=True)
    eng = pred.engine()
    imgs = ops.normalize_gray_u8(synth.make_clip_u8(8, seed=1).to(dev), synth.IMG_MEAN, synth.IMG_STD)
    for _ in range(2):
        eng.encode_frames(imgs)
if which == "decoder":  # SAM heads of a tracked frame (8 tokens), three times: read the last pass from the launch list
    from us_video_medsam2_b200 import synth
    from us_video_medsam2_b200.build_sam import build_sam2_video_predictor_npz
    pred = build_sam2_video_predictor_npz("configs/sam2.1_hiera_t512.yaml", device=dev, encoder_batch=8)
    pred.load_state_dict(synth.make_state_dict(19), strict=True)
    eng = pred.engine()
    pix = rnd(1024, 256, dt=torch.float32, sc=0.5)
    s0, s1 = rnd(16384, 32, dt=torch.float32), rnd(4096, 64, dt=torch.float32)
    for _ in range(3):
        eng.sam_heads(pix, s0, s1, 1, eng.no_point_tokens(1), multimask=True)
if which == "frames":  # a short clip without CUDA graphs: every kernel of a tracked frame shows up in the launch list
    from us_video_medsam2_b200 import synth
    from us_video_medsam2_b200.build_sam import build_sam2_video_predictor_npz
    pred = build_sam2_video_predictor_npz("configs/sam2.1_hiera_t512.yaml", device=dev, encoder_batch=8, use_cuda_graphs=False)
    pred.load_state_dict(synth.make_state_dict(19), strict=True)
    clip = ops.normalize_gray_u8(synth.make_clip_u8(12, seed=1).to(dev), synth.IMG_MEAN, synth.IMG_STD)
    st = pred.init_state(clip, 512, 512)
    pred.add_new_mask(st, 0, 1, synth.box_mask())
    for _ in pred.propagate_in_video(st):
        pass
if which == "batched":  # 8 videos x 4 objects in lock-step, no CUDA graphs: every kernel of the 32-object frame is listed
    from us_video_medsam2_b200 import synth
    from us_video_medsam2_b200.build_sam import build_sam2_video_predictor_npz
    pred = build_sam2_video_predictor_npz("configs/sam2.1_hiera_t512.yaml", device=dev, encoder_batch=16, use_cuda_graphs=False)
    pred.load_state_dict(synth.make_state_dict(19), strict=True)
    clips = [ops.normalize_gray_u8(synth.make_clip_u8(20, seed=1 + i).to(dev), synth.IMG_MEAN, synth.IMG_STD) for i in range(8)]
    states = []
    for c in clips:
        st = pred.init_state(c, 512, 512)
        for j, m in enumerate(synth.multi_object_masks(4)):
            pred.add_new_mask(st, 0, j + 1, m)
        states.append(st)
    for _ in pred.propagate_in_videos(states):
        pass
torch.cuda.synchronize()
print("done")
