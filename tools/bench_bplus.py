"""BASELINE configs[4]: a 3-D CT volume of 512 x 512 x T slices segmented the way medsam2_infer_3D_CT.py:256-283 does it --
slices resized to 1024 x 1024, sam2.1_hiera_base_plus, RECIST-style box prompt on the key slice, forward propagation,
reset_state, the same box, reverse propagation, union of the two passes, largest 3-D connected component -- on ONE B200.

    python tools/bench_bplus.py [--slices 256] [--steps 2] [--warmup 1]        (also: python bench.py --model hiera_b+)

Prints one JSON line in the shape of bench.py's: `value` = slices/s with the normalised 1024^2 volume resident in HBM,
`e2e` = the same with the uint8 volume in pinned host memory at the start of every step (upload, normalise, resize on the
device) and the final uint8 segmentation copied back, `components` = the image encoder alone (ms per slice, batched) and
the FLOP rate it implies.  Synthetic speckle volume, random-init weights of the named architecture (seed with the object
present), CUDA events, 2 warm-up volumes.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("TQDM_DISABLE", "1")

import numpy as np  # noqa: E402
import torch  # noqa: E402

SEED = 11                                               # tests/golden/bplus1024_ct_bidirectional.npz uses the same weights
BOX = np.array([190.0, 170.0, 340.0, 330.0], np.float32)  # x0, y0, x1, y1 in slice pixels


def encoder_gflops(cfg):
    """Analytic FLOPs of one image-encoder pass (real channel counts, no padding): patch embedding, per block the qkv /
    output / shortcut projections, QK^T and PV over the real window (or all tokens), the MLP; FpnNeck + conv_s0 / conv_s1."""
    from us_video_medsam2_b200.engine import hiera_plan

    side = cfg.image_size // 4
    fl = 2.0 * side * side * 147 * cfg.embed_dim
    levels = []
    for din, dout, heads, ws, pool, emit in hiera_plan(cfg):
        n_in = side * side
        if pool:
            side //= 2
        n_q = side * side
        keys = ws * ws if ws > 0 else n_in
        fl += 2.0 * n_in * din * 3 * dout + (2.0 * n_in * din * dout if din != dout else 0.0)   # qkv, shortcut
        fl += 4.0 * n_q * keys * dout + 2.0 * n_q * dout * dout                               # attention, output proj
        fl += 16.0 * n_q * dout * dout                                                        # MLP (4x)
        if emit:
            levels.append((n_q, dout))
    fl += sum(2.0 * n * c * 256 for n, c in levels)                                            # neck 1 x 1 convolutions
    fl += 2.0 * levels[0][0] * 256 * 32 + 2.0 * levels[1][0] * 256 * 64                        # conv_s0, conv_s1
    return fl / 1e9


def parse(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--slices", type=int, default=256)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--encoder-batch", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-slices", type=int, default=6)
    ap.add_argument("--encoder-sms", type=int, default=0,
                    help="SMs of a concurrent encoder partition; 0 (measured best for this model) = encoder batches and "
                         "tracked frames alternate on the whole device")
    return ap.parse_args(argv)


def time_cross_attention(dev, peaks, splits, iters=10):
    """The dominant kernel of a B+ tracked frame -- the memory attention's cross-attention, 4096 queries x 28 736 keys
    (7 memory frames + 16 pointers), d = 256, split over the keys as the engine does at one object -- timed alone with CUDA
    events on the launching stream, L2 flushed before every launch (K / V of one object: 59 MB per layer set)."""
    from us_video_medsam2_b200 import ops

    T, Nk, D = 4096, 7 * 4096 + 64, 256
    g = torch.Generator(device=dev).manual_seed(0)
    q = torch.randn((T, D), generator=g, device=dev).to(torch.bfloat16)
    kv = torch.randn((Nk, 4 * D), generator=g, device=dev).to(torch.bfloat16)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    fn = lambda: ops.fmha(q, kv, kv, 1, 1, T, Nk, D, (0, T * D, D, D), (D, Nk * 4 * D, 4 * D, D),
                          (2 * D, Nk * 4 * D, 4 * D, D), num_splits=splits)
    for _ in range(2):
        fn()
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    us = tot / iters * 1e3
    ach = 4.0 * T * Nk * D / us / 1e6
    traffic, src = None, None
    raw = os.path.join(ROOT, "profiles", "r2_bplus_fmha_ncu_raw.csv")  # committed `ncu --set full` capture of this launch
    if os.path.exists(raw):
        import csv

        hdr, units, vals = list(csv.reader(open(raw)))[:3]
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        traffic = sum(float(vals[i]) * scale[units[i]] for i, h in enumerate(hdr)
                      if h in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        src = "profiles/r2_bplus_fmha_ncu_raw.csv"
    return {"bound": "tensor", "achieved": ach, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
            "frac": ach / peaks["bf16_sustained"], "traffic": traffic, "traffic_source": src,
            "kernel": f"fmha_tc5_ts_kernel + fmha_combine_kernel<256> (memory-attention cross-attention, 4096 x {Nk} keys, "
                      f"d=256, {splits}-way split-KV)",
            "flops_per_launch": 4.0 * T * Nk * D, "avg_us": us, "launches_timed": iters, "peak_source": peaks["source"],
            "how": "CUDA events around the kernel pair alone, L2 flushed before every launch"}


def cpu_reference(sample):
    """The reference's algorithm for this configuration on the host cores (oracle port, CfgBPlus; pinned to the reference's
    own outputs by tests/test_oracle_pinned.py): box on the first slice of a `sample`-slice sub-volume, forward pass."""
    import time

    from oracle.medsam2_ref import CfgBPlus, RefPredictor
    from us_video_medsam2_b200 import synth

    gray = synth.make_clip_u8(sample, size=512, seed=1234).float() / 255.0
    mean = torch.tensor(synth.IMG_MEAN)[None, :, None, None]
    std = torch.tensor(synth.IMG_STD)[None, :, None, None]
    x = (gray[:, None].expand(-1, 3, -1, -1) - mean) / std
    clip = torch.nn.functional.interpolate(x, size=(1024, 1024), mode="bilinear", align_corners=False)
    pred = RefPredictor(synth.make_bplus_state_dict(SEED), cfg=CfgBPlus, fill_holes=True)
    with torch.inference_mode():
        t0 = time.perf_counter()
        st = pred.init_state(clip, 512, 512)
        pred.add_new_points_or_box(st, 0, 1, box=BOX)
        n = sum(1 for _ in pred.propagate_in_video(st))
        dt = time.perf_counter() - t0
    return n / dt, torch.get_num_threads()


def run(slices=256, steps=2, warmup=1, encoder_batch=8, encoder_sms=0, cpu_sample=6):
    from bench import ClockSampler, measured_peaks
    from sam2.build_sam import build_sam2_video_predictor_npz
    from us_video_medsam2_b200 import _lib, ops, synth

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    T, key = slices, slices // 2
    pred = build_sam2_video_predictor_npz("configs/sam2.1_hiera_b+.yaml", device=dev, encoder_batch=encoder_batch,
                                          encoder_sms=encoder_sms)
    pred.load_state_dict(synth.make_bplus_state_dict(SEED), strict=True)
    gray_host = synth.make_clip_u8(T, size=512, seed=1234).pin_memory()   # the volume as stored: uint8 [T, 512, 512]
    seg_host = torch.empty((T, 512, 512), dtype=torch.uint8).pin_memory()

    def ingest(g):
        """uint8 slices on the device -> ImageNet-normalised fp32 [T, 3, 1024, 1024] (device kernels only)."""
        x = ops.normalize_gray_u8(g, synth.IMG_MEAN, synth.IMG_STD)       # [T, 3, 512, 512]
        return ops.resize_bilinear(x.view(T * 3, 1, 512, 512), 1024, 1024).view(T, 3, 1024, 1024)

    def one_volume(images):
        seg = torch.zeros((T, 512, 512), dtype=torch.bool, device=dev)
        st = pred.init_state(images, 512, 512)
        n = 0
        for kw in ({}, dict(reverse=True)):
            pred.add_new_points_or_box(st, key, 1, box=BOX)
            for t, ids, logits in pred.propagate_in_video(st, **kw):
                seg[t] |= logits[0, 0] > 0
                n += 1
            pred.reset_state(st)
        return ops.largest_component_3d(seg), n

    vol_dev = ingest(gray_host.to(dev))

    def step_resident():
        return one_volume(vol_dev)[1]

    def step_e2e():
        seg, n = one_volume(ingest(gray_host.to(dev, non_blocking=True)))
        seg_host.copy_(seg, non_blocking=True)
        return n

    def timed(fn):
        sampler = ClockSampler(0)
        if not os.environ.get("USVM2_NO_SAMPLER"):
            sampler.start()
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        sampler.mark_begin()
        l0 = _lib.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        frames = 0
        for _ in range(steps):
            frames += fn()
        e1.record()
        torch.cuda.synchronize()
        sampler.mark_end()
        return e0.elapsed_time(e1), frames, _lib.launch_count - l0, sampler.stop()

    with torch.inference_mode():
        ms, frames, launches, clocks = timed(step_resident)
        ms_e2e, frames_e2e, _, _ = timed(step_e2e)
        fg = int(seg_host.sum())
        # the image encoder alone, batched (its share of the slice time; the rest is the sequential propagation)
        eng = pred.engine()
        x = vol_dev[:encoder_batch].contiguous()
        for _ in range(2):
            eng.encode_frames(x)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            eng.encode_frames(x)
        b.record()
        torch.cuda.synchronize()
        enc_ms = a.elapsed_time(b) / 5 / encoder_batch
    peaks = measured_peaks()
    cpu = None
    if cpu_sample > 0:
        fps, cores = cpu_reference(cpu_sample)
        cpu = {"value": fps, "unit": "slices/s", "cores": cores, "kind": "port",
               "sample": f"{cpu_sample}-slice sub-volume, box on its first slice, forward pass; oracle port of the reference "
                         f"(CfgBPlus), torch {torch.__version__} CPU fp32"}
    with torch.inference_mode():
        roofline = time_cross_attention(dev, peaks, pred.engine()._splits(1, 7 * 4096 + 64))
    GF_ENCODER_BPLUS = encoder_gflops(pred.cfg)
    enc_tf = GF_ENCODER_BPLUS / enc_ms  # GFLOP / ms = TFLOP/s
    return {
        "metric": "CT slices/sec (sam2.1_hiera_base_plus, 1024x1024, bidirectional propagation from a box on the key slice)",
        "value": T * steps / (ms / 1e3), "unit": "slices/s", "n_gpus": 1, "steps": steps, "warmup": warmup,
        "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": f"3-D CT volume 512 x 512 x {T} (uint8 slices resized to 1024^2), sam2.1_hiera_base_plus, box "
                               f"prompt on slice {key}, forward + reverse propagation ({frames // steps} tracked frames per "
                               "volume), union, largest 3-D connected component (BASELINE configs[4])",
                   "slices": T, "objects": 1, "encoder_batch": encoder_batch,
                   "encoder_sms": (pred._partition_obj.sms if pred._partition_obj is not None else 0),
                   "l2": f"inputs larger than L2: {T * 3 * 1024 * 1024 * 4 / 1e6:.0f} MB volume streamed once per pass"},
        "propagated_frames_per_s": frames / (ms / 1e3),
        "e2e": {"value": T * steps / (ms_e2e / 1e3), "unit": "slices/s", "h2d_bytes_per_step": T * 512 * 512,
                "d2h_bytes_per_step": T * 512 * 512, "propagated_frames_per_s": frames_e2e / (ms_e2e / 1e3)},
        "gpu_launches": launches, "clocks": clocks, "segmented_voxels": fg, "cpu_baseline": cpu, "roofline": roofline,
        "components": {"image_encoder": {"ms_per_slice": enc_ms, "batch": encoder_batch, "achieved": enc_tf,
                                         "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
                                         "frac": enc_tf / peaks["bf16_sustained"],
                                         "flops": f"{GF_ENCODER_BPLUS:.1f} GFLOP per slice (analytic, tools/bench_bplus.py:encoder_gflops)"}},
    }


def main(argv=None):
    a = parse(argv)
    print(json.dumps(run(a.slices, a.steps, a.warmup, a.encoder_batch, a.encoder_sms,
                         0 if a.no_cpu_baseline else a.cpu_sample_slices)))


if __name__ == "__main__":
    main()
