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
if which == "r2set":  # round-2 `ncu --set full` set: throughput GEMM shapes + the kernels VERDICT r1 asked counters for
    f32 = torch.float32

    REPS = int(os.environ.get("R2SET_REPS", "3"))

    def gemm(M, N, K, act=0, rope=None, f32out=False, res=False, reps=REPS):
        a, w, b = rnd(M, K), rnd(N, K, sc=K ** -0.5), rnd(N, dt=f32)
        r = rnd(M, N, dt=f32) if res else None
        for _ in range(reps):
            ops.gemm_bf16(a, w, bias=b, act=act, residual=r, f32=f32out, bf16=not f32out, rope=rope)

    gemm(8192, 1152, 384)                      # encoder stage-3 qkv (8 frames): persistent, bias -> bf16
    gemm(8192, 1536, 384, act=ops.ACT_GELU)    # encoder stage-3 MLP up-projection, exact-erf GELU epilogue
    gemm(131072, 384, 96, act=ops.ACT_GELU)    # encoder stage-1 MLP up-projection (HBM heavy)
    gemm(8192, 384, 1536, f32out=True, res=True)   # encoder stage-3 MLP down-projection + residual -> fp32
    gemm(32768, 2048, 256, act=ops.ACT_RELU)   # memory-attention FFN linear1 at 32 objects
    cs, sn = ops.tile_rope_table(rnd(1024, 128, dt=f32)), ops.tile_rope_table(rnd(1024, 128, dt=f32))
    gemm(32 * 7232, 1024, 64, rope=(cs, sn, 1024, 7232, 7168))  # bank key projection + RoPE at 32 objects
    gemm(1024, 768, 256, rope=(cs, sn, 512, 1024, 1024))        # gemm_bf16_tc5_kernel<32>: one-object qkv projection
    # Hiera global attention, 8 frames x 4 heads of 96 over 1024 tokens (fmha_bf16_kernel<96>)
    Fr, T, C, H = 8, 1024, 384, 4
    qkv = rnd(Fr * T, 3 * C)
    for _ in range(REPS):
        ops.fmha(qkv, qkv, qkv, Fr, H, T, T, 96, (0, T * 3 * C, 3 * C, 96), (C, T * 3 * C, 3 * C, 96),
                 (2 * C, T * 3 * C, 3 * C, 96))
    # memory-encoder depthwise 7x7 + LayerNorm at 1 and 32 objects, hole filling at 1 and 32 objects
    for B in (1, 32):
        x = rnd(B * 1024, 256, dt=f32)
        dw, db, lw, lb = rnd(49, 256, dt=f32), rnd(256, dt=f32), rnd(256, dt=f32), rnd(256, dt=f32)
        low = rnd(B, 1, 128, 128, dt=f32, sc=0.07)
        for _ in range(REPS):
            ops.dwconv7_ln(x, dw, db, lw, lb, B, 32, 32)
            ops.fill_holes(low, 8)
    # cross-attention at 32 objects (no split) -- the roofline kernel at the batched shape
    B, T, Nk, D = 32, 1024, 7232, 256
    q, kv = rnd(B * T, D), rnd(B * Nk, 4 * D)
    for _ in range(REPS):
        ops.fmha(q, kv, kv, B, 1, T, Nk, D, (0, T * D, D, D), (D, Nk * 4 * D, 4 * D, D), (2 * D, Nk * 4 * D, 4 * D, D))
if which == "r2b":  # second round-2 `ncu --set full` set: the kernels written after r2set (see profiles/r2b_kernel_set_ncu.md)
    f32 = torch.float32

    def gemm(M, N, K, act=0, f32out=False, res=False):
        a, w, b = rnd(M, K), rnd(N, K, sc=K ** -0.5), rnd(N, dt=f32)
        r = rnd(M, N, dt=f32) if res else None
        ops.gemm_bf16(a, w, bias=b, act=act, residual=r, f32=f32out, bf16=not f32out)

    # persistent GEMM, weight-stationary schedule (16 frames): stage-3 qkv / MLP up + GELU, stage-1 MLP up + GELU;
    # streaming schedule: stage-3 MLP down (K = 1536)
    gemm(16384, 1152, 384)
    gemm(16384, 1536, 384, act=ops.ACT_GELU)
    gemm(262144, 384, 96, act=ops.ACT_GELU)
    gemm(16384, 384, 1536, f32out=True, res=True)
    gemm(32768, 2048, 256, act=ops.ACT_RELU)   # memory-attention FFN linear1 at 32 objects
    # encoder attention on tcgen05: global and 14 x 14 windowed, 16 frames x 4 heads of 96
    Fr, Hh, C, heads = 16, 32, 384, 4
    qkv = rnd(Fr * Hh * Hh, 3 * C)
    bias = rnd(3 * C, dt=f32)
    ops.hiera_attn(qkv, bias, Fr, Hh, Hh, C, heads, 0)
    ops.hiera_attn(qkv, bias, Fr, Hh, Hh, C, heads, 14)
    # fused feed-forward block at one object and at 8 objects
    for M in (1024, 8192):
        x, h = rnd(M, 256, dt=f32), rnd(M, 256)
        w1, w2 = rnd(2048, 256, sc=1 / 16), rnd(256, 2048, sc=1 / 45)
        ops.ffn_fused(h, x, w1, rnd(2048, dt=f32), w2, rnd(256, dt=f32))
    # tiled hole filling at 1 and 32 objects
    for B in (1, 32):
        ops.fill_holes(rnd(B, 1, 128, 128, dt=f32, sc=0.07), 8)
    # cross-attention: one object (12-way split, as planned beside the encoder partition) and 32 objects (no split)
    T, Nk, D = 1024, 7232, 256
    for B, splits in ((1, 12), (32, 1)):
        q, kv = rnd(B * T, D), rnd(B * Nk, 4 * D)
        ops.fmha(q, kv, kv, B, 1, T, Nk, D, (0, T * D, D, D), (D, Nk * 4 * D, 4 * D, D), (2 * D, Nk * 4 * D, 4 * D, D),
                 num_splits=splits)
torch.cuda.synchronize()
print("done")
