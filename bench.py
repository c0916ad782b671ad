#!/usr/bin/env python
"""Benchmark of the MedSAM2 per-frame propagation path (BASELINE.json metric: propagated frames/s).

  python bench.py --gpus N --steps K --warmup W          # this repo's CUDA path
  python bench.py --impl reference ...                   # the reference algorithm on the host CPU (oracle port)

Workload (BASELINE.json configs[1]): sam2.1_hiera_t512 `propagate_in_video` over a 512-frame synthetic
512x512 grayscale echo clip, 1 object, mask prompt on frame 0, 7-frame memory bank, random-init weights.
One *step* = one full propagation pass over the clip.  With N > 1 every rank tracks its own clip (independent
videos shard with no communication -> weak scaling); the value is the aggregate frames/s over all ranks, timed
on the device (CUDA events), barrier + synchronize on both sides, max over ranks.

JSON keys beyond the base contract:
  e2e          same metric through the public API from HOST buffers: per step the uint8 clip is copied from
               pinned host memory, normalised on the device, tracked, and every frame's binary mask is copied back
  roofline     dominant kernel (memory-attention cross-attention flash kernel): algorithmic FLOPs / CUDA-event
               time inside the timed region vs the measured bf16 peak
  cpu_baseline the oracle port of the reference timed on the host cores on a bounded sample of the same workload
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.environ.setdefault("TQDM_DISABLE", "1")

import torch  # noqa: E402

METRIC = "propagated frames/sec (sam2.1_hiera_t512, 512x512, {objects} object{s}, 7-frame memory bank)"


def metric_name(objects, model="hiera_t512"):
    name = METRIC.format(objects=objects, s="" if objects == 1 else "s")
    return name.replace("sam2.1_hiera_t512", MODEL_NAMES[model])
SEED = 19
ETAM_SEEDS = {"etam_ti": 18, "etam_s": 3}   # object-present branch with margin at random init
MODEL_NAMES = {"hiera_t512": "sam2.1_hiera_t512", "etam_ti": "efficienttam_ti_512x512", "etam_s": "efficienttam_s_512x512"}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=512, help="clip length T (BASELINE config 2: 512)")
    ap.add_argument("--objects", type=int, default=1)
    ap.add_argument("--encoder-batch", type=int, default=16, help="frames per batched image-encoder pass")
    ap.add_argument("--encoder-sms", type=int, default=48,
                    help="> 0: look-ahead encoder batches run concurrently with tracking on an SM partition of this size")
    ap.add_argument("--mode", default="videos", choices=["videos", "clip"],
                    help="N > 1: 'videos' = one independent clip per GPU (weak scaling, no communication); 'clip' = ONE "
                         "clip, rank 0 propagates, the other ranks run the frame-parallel encoder and send features over "
                         "NCCL point-to-point (strong scaling, bounded by the sequential propagation)")
    ap.add_argument("--model", default="hiera_t512", choices=["hiera_t512", "etam_ti", "etam_s", "hiera_b+"],
                    help="hiera_t512 = MedSAM2 sam2.1_hiera_t512 (BASELINE configs[1], the default); etam_ti / etam_s = "
                         "EfficientTAM tiny / small at 512x512 (BASELINE configs[3]) on the same clips; hiera_b+ = "
                         "sam2.1_hiera_base_plus at 1024x1024 on a 3-D CT volume, bidirectional (BASELINE configs[4]; "
                         "tools/bench_bplus.py, one GPU, --frames = slices, default 256)")
    ap.add_argument("--cpu-sample-frames", type=int, default=64,
                    help="frames of the clip the CPU arm tracks per step (the 7-frame bank is full from frame 7 on)")
    ap.add_argument("--batched-videos", type=int, default=8,
                    help="BASELINE configs[2] leg: independent videos per GPU tracked in lock-step as one batched frame "
                         "graph (0 = skip the leg)")
    ap.add_argument("--batched-objects", type=int, default=4)
    ap.add_argument("--batched-frames", type=int, default=128)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-clip-leg", action="store_true",
                    help="N > 1, videos mode: skip the extra single-clip (strong scaling) measurement")
    return ap.parse_args()


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(bf16_sustained=p.get("bf16_tflops_sustained", 1405.3), bf16_burst=p.get("bf16_tflops", 1685.0),
                    hbm=p.get("hbm_gbs", 6447.2), source="measured (MEASURED_PEAKS.json)")
    return dict(bf16_sustained=1400.0, bf16_burst=1590.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


# ------------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region, from a background thread through NVML
    (nvidia_ml_py; one-shot `nvidia-smi` calls if the module is missing).  An NVML query -- from a looping `nvidia-smi`
    child or in-process alike -- was measured to stall this process's kernel launches for 20-470 ms.  That is harmless
    while the host runs hundreds of frames ahead of the GPU, but at the start of the timed region the launch queue is
    empty and the stall turned into GPU idle time in about one run out of three.  So the first in-region sample is taken
    `lead` seconds after the region starts (the host has built its lead by then) and every `period` seconds from there."""
    NAMES = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, index, period=0.25, lead=0.4):
        self.index, self.period, self.lead, self.rows = index, period, lead, []
        self.t0 = self.t1 = None
        self._stop, self._begin = threading.Event(), threading.Event()
        self._thread, self._nvml, self._handle, self.source = None, None, None, None

    def mark_begin(self):
        self.t0 = time.time()
        self._begin.set()

    def mark_end(self):
        self.t1 = time.time()
        self._stop.set()

    def start(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            try:
                uuid = str(torch.cuda.get_device_properties(self.index).uuid)
                self._handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
            except Exception:
                self._handle = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self._nvml, self.source = pynvml, "nvml"
            self._sample()  # first query (lazy NVML state) before the warm-up
        except Exception:
            self._nvml, self.source = None, "nvidia-smi"
        self._thread = threading.Thread(target=self._loop, daemon=True)
        self._thread.start()

    def _sample(self):
        if self._nvml is not None:
            n, h = self._nvml, self._handle
            sm = n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM)
            mx = n.nvmlDeviceGetMaxClockInfo(h, n.NVML_CLOCK_SM)
            try:
                mask = n.nvmlDeviceGetCurrentClocksEventReasons(h)
            except Exception:
                mask = n.nvmlDeviceGetCurrentClocksThrottleReasons(h)
            self.rows.append((time.time(), int(sm), int(mx), int(mask)))
            return
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.active"
        out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                             capture_output=True, text=True, timeout=10).stdout.strip().split(",")
        self.rows.append((time.time(), int(out[0]), int(out[1]), int(out[2].strip(), 16)))

    def _loop(self):
        self._begin.wait()
        if self._stop.wait(self.lead):
            return
        while True:
            try:
                self._sample()
            except Exception:
                pass
            if self._stop.wait(self.period if self._nvml is not None else 1.0):
                return

    def stop(self):
        self._stop.set()
        if self._thread is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["clock sampler not started"])
        self._begin.set()
        self._thread.join(timeout=5)
        lo = self.t0 if self.t0 is not None else 0.0
        hi = self.t1 if self.t1 is not None else float("inf")
        rows = [r for r in self.rows if lo <= r[0] <= hi + 0.05]
        if not rows:  # region shorter than `lead`: one sample right after it (the GPU is still at its load clocks)
            try:
                self._sample()
            except Exception:
                pass
            rows = self.rows[-1:]
        if not rows:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["no clock samples"], source=self.source)
        sm = sorted(r[1] for r in rows)
        mask = 0
        for r in rows:
            mask |= r[3]
        reasons = sorted(name for name, bit in self.NAMES if mask & bit)
        return dict(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(r[2] for r in rows), reasons=reasons, samples=len(sm),
                    source=self.source)


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_reference_fps(frames, objects, passes=1, model="hiera_t512"):
    """frames/s of the CPU port of the reference (oracle/medsam2_ref.py) on a `frames`-long sample of the
    workload: same weights, same clip generator, same prompt, fill-holes skipped as the reference does on CPU."""
    from oracle.medsam2_ref import RefPredictor
    from us_video_medsam2_b200 import synth

    torch.set_num_threads(os.cpu_count() or 1)
    clip = synth.make_clip(frames, kind="speckle")
    masks = [synth.box_mask()] if objects == 1 else synth.multi_object_masks(objects)
    if model != "hiera_t512":
        from oracle.etam_ref import EtamCfg, EtamSCfg, etam_predictor

        v = model.split("_")[1]
        pred = etam_predictor(synth.make_etam_state_dict(ETAM_SEEDS[model], v), fill_holes=False,
                              cfg=EtamCfg if v == "ti" else EtamSCfg)
    else:
        pred = RefPredictor(synth.make_state_dict(SEED), fill_holes=False)
    best = None
    with torch.inference_mode():
        for _ in range(passes):
            st = pred.init_state(clip, 512, 512)
            for i, m in enumerate(masks):
                pred.add_new_mask(st, 0, i + 1, m)
            t0 = time.perf_counter()
            n = sum(1 for _ in pred.propagate_in_video(st))
            dt = time.perf_counter() - t0
            fps = n / dt
            best = fps if best is None else max(best, fps)
    return best, torch.get_num_threads()


def run_reference(args, rank):
    if rank != 0:
        return
    sample = max(4, min(args.frames, args.cpu_sample_frames))
    times = []
    cores = os.cpu_count() or 1
    for i in range(args.warmup + args.steps):
        fps, cores = cpu_reference_fps(sample, args.objects, model=args.model)
        if i >= args.warmup:
            times.append(sample / fps)
    ms = 1000.0 * sum(times) / len(times)
    value = sample / (ms / 1000.0)
    unit = "frames/s"
    print(json.dumps({
        "metric": metric_name(args.objects, args.model), "value": value, "unit": unit, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "impl": "reference",
        "config": {"workload": f"{MODEL_NAMES[args.model]} "
                               f"propagate_in_video, CPU port of the reference, {sample}-frame sample "
                               f"of the {args.frames}-frame synthetic echo clip, {args.objects} object(s)",
                   "frames_per_step": sample, "objects": args.objects},
        "cpu_baseline": {"value": value, "unit": unit, "cores": cores, "kind": "port",
                         "sample": f"{sample} frames per step, torch {torch.__version__} CPU fp32, fill-holes skipped "
                                   "(CUDA-only op in the reference)"},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ------------------------------------------------------------------------------------------------
# this repo's arm
# ------------------------------------------------------------------------------------------------
def time_dominant_kernel(dev, B, iters=20):
    """CUDA-event timing (on the launching stream) of the dominant kernel of the sequential part at its steady-state
    shape: memory-attention cross-attention, 1024 queries x (7 x 1024 + 64) keys, one head of 256, split over the
    key range.  The steady-state frame itself is replayed from a CUDA graph, where events cannot be placed between
    kernels, so the kernel pair is timed here on identical shapes, two ways:
      in context  -- as it runs in the frame: its operands were written microseconds earlier by the bank K / V projections
                     and the query projection.  Two CUDA graphs are timed, [producers] and [producers + attention pair],
                     each replay preceded by an L2 flush; the difference is the pair's duration (this is `avg_us`);
      cold        -- the pair alone, launched eagerly after an L2 flush (`avg_us_cold_l2`, the pessimistic figure)."""
    from us_video_medsam2_b200 import ops

    T, Nk, D = 1024, 7 * 1024 + 64, 256
    g = torch.Generator(device=dev).manual_seed(0)
    rnd = lambda *shape, scale=1.0: (torch.randn(shape, generator=g, device=dev) * scale).to(torch.bfloat16)
    q = rnd(B * T, D)
    kv = rnd(B * Nk, 4 * D)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    splits = max(1, min(max(1, 148 // (8 * B)), (Nk + 63) // 64))  # same rule as Engine._splits

    def pair(q_, k_, v_):
        return ops.fmha(q_, k_, v_, B, 1, T, Nk, D, (0, T * D, D, D), (D, Nk * 4 * D, 4 * D, D),
                        (2 * D, Nk * 4 * D, 4 * D, D), num_splits=splits)

    for _ in range(3):
        pair(q, kv, kv)
    cold = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        pair(q, kv, kv)
        e.record()
        torch.cuda.synchronize()
        cold.append(s.elapsed_time(e))
    # in context: producers = the three GEMMs that write the pair's operands in the frame
    mem_k, mem_v, h = rnd(B * Nk, 64), rnd(B * Nk, 64), rnd(B * T, D)
    wk, wv, wq = rnd(4 * D, 64, scale=0.125), rnd(4 * D, 64, scale=0.125), rnd(D, D, scale=0.0625)

    def producers():
        _, k_all = ops.gemm_bf16(mem_k, wk, bf16=True)
        _, v_all = ops.gemm_bf16(mem_v, wv, bf16=True)
        _, q_ = ops.gemm_bf16(h, wq, bf16=True)
        return q_, k_all, v_all

    def graph_ms(with_pair):
        for _ in range(2):
            r = producers()
            if with_pair:
                pair(*r)
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            r = producers()
            if with_pair:
                pair(*r)
        times = []
        for _ in range(iters):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            gr.replay()
            e.record()
            torch.cuda.synchronize()
            times.append(s.elapsed_time(e))
        return sum(times) / len(times)

    in_ctx = graph_ms(True) - graph_ms(False)
    flops = 4.0 * B * T * Nk * D
    return dict(avg_ms=in_ctx, cold_ms=sum(cold) / len(cold), flops_per_launch=flops, launches=iters, splits=splits)


def time_other_kernels(dev, B):
    """The launch list (profiles/r2b_launches_tracked_frames.csv) shows no single kernel above 19 % of a one-object frame:
    the one-tile GEMM family carries the largest share by time, the cross-attention pair by FLOPs.  For transparency the
    most frequent GEMM shape of the memory attention, its fused feed-forward block and the encoder's tcgen05 attention
    are timed too -- 20 launches captured in one CUDA graph (as they run in the frame: no host launch overhead between
    them), L2 warm (their operands are the previous kernel's output)."""
    from us_video_medsam2_b200 import ops

    g = torch.Generator(device=dev).manual_seed(1)

    def graph_us(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            for _ in range(20):
                fn()
        graph.replay()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(5):
            graph.replay()
        e.record()
        torch.cuda.synchronize()
        return s.elapsed_time(e) / 100 * 1e3

    out = []
    M, N, K = 1024 * B, 768, 256
    a = torch.randn((M, K), generator=g, device=dev).to(torch.bfloat16)
    w = (torch.randn((N, K), generator=g, device=dev) * K ** -0.5).to(torch.bfloat16)
    bias = torch.randn((N,), generator=g, device=dev)
    us = graph_us(lambda: ops.gemm_bf16(a, w, bias=bias, bf16=True))
    out.append({"kernel": "memory-attention qkv / q / out projections (gemm_bf16_tc5_kernel<32>)", "shape_mnk": [M, N, K],
                "avg_us": us, "achieved_tflops": 2.0 * M * N * K / us / 1e6,
                "bound": "latency (one 128-row tile per CTA, 0.4 GFLOP per launch)"})
    if M <= 2048:  # (the engine uses the fused kernel at one or two objects)
        x = torch.randn((M, 256), generator=g, device=dev)
        w1 = (torch.randn((2048, 256), generator=g, device=dev) / 16).to(torch.bfloat16)
        w2 = (torch.randn((256, 2048), generator=g, device=dev) / 45).to(torch.bfloat16)
        b1, b2 = torch.randn((2048,), generator=g, device=dev), torch.randn((256,), generator=g, device=dev)
        us = graph_us(lambda: ops.ffn_fused(a, x, w1, b1, w2, b2))
        out.append({"kernel": "memory-attention feed-forward block (ffn_fused_tc5_kernel: linear1 + ReLU + linear2 + residual)",
                    "shape_mnk": [M, 2048, 256], "avg_us": us, "achieved_tflops": 4.0 * M * 2048 * 256 / us / 1e6,
                    "bound": "latency (cluster of 8 CTAs per 128-row tile, reduce-scatter through distributed shared memory)"})
    Fr, C, heads = 16, 384, 4
    qkv = torch.randn((Fr * 1024, 3 * C), generator=g, device=dev).to(torch.bfloat16)
    qb = torch.randn((3 * C,), generator=g, device=dev)
    for ws, nm in ((0, "global, 1024 x 1024 per head"), (14, "14 x 14 windows, padding tokens in closed form")):
        us = graph_us(lambda: ops.hiera_attn(qkv, qb, Fr, 32, 32, C, heads, ws))
        fl = 4.0 * Fr * heads * (1024 * 1024 if ws == 0 else 9 * 196 * 196) * 96
        out.append({"kernel": f"image-encoder attention, 16 frames x 4 heads of 96 (hiera_attn_tc5_kernel: {nm})",
                    "avg_us": us, "achieved_tflops": fl / us / 1e6,
                    "bound": "tensor (bf16), per tile softmax (MUFU) bound; FLOPs of padded windows counted"})
    return out


def time_dominant_kernel_batched(dev, peaks, B=32, iters=10):
    """The roofline kernel at its THROUGHPUT shape -- the batched frame of BASELINE configs[2] on one GPU (8 videos x 4
    objects): 32 objects x 1024 queries x 7232 keys, no split-KV, no combine.  CUDA events around single launches, L2
    flushed before each (K / V of 32 objects are 237 MB: larger than L2 anyway)."""
    from us_video_medsam2_b200 import ops

    T, Nk, D = 1024, 7 * 1024 + 64, 256
    g = torch.Generator(device=dev).manual_seed(0)
    q = torch.randn((B * T, D), generator=g, device=dev).to(torch.bfloat16)
    kv = torch.randn((B * Nk, 4 * D), generator=g, device=dev).to(torch.bfloat16)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    fn = lambda: ops.fmha(q, kv, kv, B, 1, T, Nk, D, (0, T * D, D, D), (D, Nk * 4 * D, 4 * D, D),
                          (2 * D, Nk * 4 * D, 4 * D, D), num_splits=1)
    for _ in range(2):
        fn()
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        tot += s.elapsed_time(e)
    us = tot / iters * 1e3
    ach = 4.0 * B * T * Nk * D / us / 1e6
    return {"kernel": "fmha_tc5_ts_kernel (memory-attention cross-attention at 32 objects, no split)", "objects": B,
            "avg_us": us, "achieved": ach, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
            "frac": ach / peaks["bf16_sustained"], "launches_timed": iters}


GF_ENCODER, GF_PER_OBJECT = 62.9, 52.0  # GFLOP per frame (shared by objects) / per object-frame
# HBM-bound pieces, algorithmic bytes per object-frame: memory encoder = 1.0 MB mask in (fp32 512^2) + 0.5 MB
# pix_feat (bf16) + 2.8 MB weights + 0.13 MB memory out; hole filling = 64 KiB logits in + 64 KiB out
BYTES_MEM_ENCODER_WEIGHTS = 2.8e6
BYTES_MEM_ENCODER_ACT = 1.048576e6 + 0.524288e6 + 0.131072e6
BYTES_MEM_ENCODER = BYTES_MEM_ENCODER_ACT + BYTES_MEM_ENCODER_WEIGHTS
BYTES_FILL_HOLES = 2 * 128 * 128 * 4


def _event_time(fn, iters, flush=None):
    """mean ms of fn() over `iters` launches, CUDA events on the launching (current) stream, optional L2 flush before each."""
    fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        tot += s.elapsed_time(e)
    return tot / iters


def time_components(pred, dev, B, peaks, iters=10, batched_objects=0):
    """What north_star asks besides the headline: tensor-pipe utilisation of the batched image encoder and achieved HBM GB/s
    of the memory encoder and the hole-filling post-process, each timed with CUDA events at its in-run shape (the encoder
    as the 16-frame captured graph the clip replays; the other two as the launches of one tracked frame), L2 flushed
    before every repetition."""
    from us_video_medsam2_b200 import ops

    eng = pred.engine()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    out = {}
    n = pred.encoder_batch
    if pred.use_cuda_graphs and n > 1:
        graph, static_in, _, n_kernels = pred._encoder_graph(n)
        static_in.normal_()
        ms = _event_time(graph.replay, iters, flush)
        tf = GF_ENCODER * n / ms  # GFLOP / ms = TFLOP/s
        out["image_encoder"] = {"bound": "tensor", "frames_per_launch_group": n, "kernels": n_kernels, "avg_ms": ms,
                                "ms_per_frame": ms / n, "achieved": tf, "peak": peaks["bf16_sustained"],
                                "unit": "TFLOP/s", "frac": tf / peaks["bf16_sustained"],
                                "flops": f"{GF_ENCODER} GFLOP per frame (SURVEY 8d)"}
    g = torch.Generator(device=dev).manual_seed(5)
    fb = torch.randn((1024, 256), generator=g, device=dev).to(torch.bfloat16)
    # at the benched object count (latency shapes) and, with the batched leg on, at its 32 objects (throughput shapes)
    for nobj, sfx in [(B, "")] + ([(batched_objects, "_batched")] if batched_objects and batched_objects != B else []):
        low = torch.randn((nobj, 1, 128, 128), generator=g, device=dev) * 0.07
        score = torch.ones((nobj, 1), device=dev)

        def mem_enc():
            eng.encode_memory(fb, eng.mem_mask_input(low, False), score, nobj)

        for _ in range(2):
            mem_enc()
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            mem_enc()
        ms = _event_time(gr.replay, iters, flush)
        nbytes = BYTES_MEM_ENCODER_WEIGHTS + BYTES_MEM_ENCODER_ACT * nobj  # weights once per launch group
        gbs = nbytes / ms / 1e6
        out["memory_encoder" + sfx] = {
            "bound": "hbm", "objects": nobj, "avg_us": ms * 1e3, "achieved": gbs, "peak": peaks["hbm"], "unit": "GB/s",
            "frac": gbs / peaks["hbm"], "achieved_tflops": 2.9 * nobj / ms,
            "bytes": f"{BYTES_MEM_ENCODER_WEIGHTS / 1e6:.1f} MB of weights per launch group + "
                     f"{BYTES_MEM_ENCODER_ACT / 1e6:.2f} MB per object-frame (SURVEY 8d); 15 dependent launches: latency bound "
                     "at one object, a mix of bandwidth and fp32 / tensor math (2.9 GFLOP per object) at 32"}
        ms = _event_time(lambda: ops.fill_holes(low, 8), iters, flush)
        gbs = BYTES_FILL_HOLES * nobj / ms / 1e6
        out["fill_holes" + sfx] = {"bound": "hbm", "objects": nobj, "avg_us": ms * 1e3, "achieved": gbs,
                                   "peak": peaks["hbm"], "unit": "GB/s", "frac": gbs / peaks["hbm"],
                                   "bytes": "64 KiB in + 64 KiB out per object-frame"}
    return out


def ncu_traffic(kernel_substr):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the roofline kernel from the committed `ncu --set full`
    raw page (profiles/r2_roofline_kernel_ncu_raw.csv), or None when the capture is absent."""
    import csv

    path = os.path.join(ROOT, "profiles", "r2_roofline_kernel_ncu_raw.csv")
    if not os.path.exists(path):
        return None, None
    try:
        rows = list(csv.reader(open(path)))
        hdr = next(r for r in rows if "Kernel Name" in r)
        units = rows[rows.index(hdr) + 1]
        name_i = hdr.index("Kernel Name")
        rd_i, wr_i = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        vals = [float(r[rd_i].replace(",", "")) * scale.get(units[rd_i], 1.0)
                + float(r[wr_i].replace(",", "")) * scale.get(units[wr_i], 1.0)
                for r in rows[rows.index(hdr) + 2:] if len(r) > max(rd_i, wr_i) and kernel_substr in r[name_i]]
        return (sum(vals) / len(vals), os.path.relpath(path, ROOT)) if vals else (None, None)
    except Exception:
        return None, None


def run_batched_leg(pred, args, dev, rank, world, peaks):
    """BASELINE configs[2] on this GPU's share: `--batched-videos` independent videos x `--batched-objects` objects each,
    `--batched-frames` frames, tracked in lock-step by propagate_in_videos (one batched frame graph, no per-frame host
    sync).  Device-timed (CUDA events, barrier + synchronize on both sides, max over ranks), clips resident in HBM."""
    import torch.distributed as dist

    from us_video_medsam2_b200 import ops, synth

    S, Bo, T = args.batched_videos, args.batched_objects, args.batched_frames
    clips = [ops.normalize_gray_u8(synth.make_clip_u8(T, seed=4321 + rank * S + i).to(dev), synth.IMG_MEAN, synth.IMG_STD)
             for i in range(S)]
    masks = [synth.box_mask()] if Bo == 1 else synth.multi_object_masks(Bo)

    def one_pass(events=None):
        states = []
        for c in clips:
            st = pred.init_state(c, 512, 512)
            for j, m in enumerate(masks):
                pred.add_new_mask(st, 0, j + 1, m)
            states.append(st)
        # the metric is the propagation loop (frames yielded / time of the generator loop, excluding init_state
        # and prompting): events bracket exactly that
        if events is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        n = sum(1 for _ in pred.propagate_in_videos(states))
        if events is not None:
            e1.record()
            events.append((e0, e1))
        return n

    for _ in range(2):  # the batched graphs are captured at their second sighting
        one_pass()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    passes, steps, events = 3, 0, []
    for _ in range(passes):
        steps += one_pass(events)
    torch.cuda.synchronize()
    ms = torch.tensor([sum(a.elapsed_time(b) for a, b in events)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    sec = float(ms) / 1e3
    steps_s = steps / sec                      # lock-step frames per second on one GPU
    video_fps = steps_s * S * world            # propagated video-frames/s, whole job
    objf = video_fps * Bo
    tflops = (GF_ENCODER * S + GF_PER_OBJECT * S * Bo) * steps_s / 1e3  # per GPU
    return {"workload": f"{S * world} independent {T}-frame videos x {Bo} objects ({S} videos per GPU in lock-step as one "
                        f"batched frame graph of {S * Bo} objects; BASELINE configs[2] shape, no inter-GPU communication)",
            "videos_per_gpu": S, "objects_per_video": Bo, "frames": T, "n_gpus": world,
            "value": video_fps, "unit": "frames/s", "object_frames_per_s": objf, "ms_per_lockstep_frame": 1e3 / steps_s,
            "whole_path_tflops_per_gpu": tflops, "whole_path_frac": tflops / peaks["bf16_sustained"],
            "timing": "device (CUDA events around each propagate_in_videos generator loop; init_state and prompting are "
                      "outside, per SURVEY 8d), clips resident in HBM, 2 warm-up + 3 timed passes, max over ranks"}


def run_b200(args, rank, world):
    import torch.distributed as dist

    from us_video_medsam2_b200 import _lib, ops, synth
    from us_video_medsam2_b200.build_sam import build_sam2_video_predictor_npz

    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    T, B = args.frames, args.objects

    def measure(clip_mode, steps, warmup):
        """Build a predictor for the mode, run the resident and the end-to-end timed regions; returns a dict."""
        if args.model != "hiera_t512":
            from us_video_medsam2_b200.build_etam import build_efficienttam_video_predictor_npz

            v = args.model.split("_")[1]
            pred = build_efficienttam_video_predictor_npz(f"configs/efficienttam_{v}_512x512.yaml", device=dev,
                                                          encoder_batch=args.encoder_batch,
                                                          encoder_sms=0 if clip_mode else args.encoder_sms)
            pred.load_state_dict(synth.make_etam_state_dict(ETAM_SEEDS[args.model], v), strict=True)
        else:
            pred = build_sam2_video_predictor_npz("configs/sam2.1_hiera_t512.yaml", device=dev,
                                                  encoder_batch=args.encoder_batch,
                                                  encoder_sms=0 if clip_mode else args.encoder_sms)
            pred.load_state_dict(synth.make_state_dict(SEED), strict=True)
        masks = [synth.box_mask()] if B == 1 else synth.multi_object_masks(B)
        # 'clip' mode: every rank holds the SAME clip (rank 0 tracks it, the others encode their share of its frames)
        gray_host = synth.make_clip_u8(T, seed=1234 + (0 if clip_mode else rank)).pin_memory()  # [T,512,512] uint8, pinned
        if clip_mode and rank == 0:
            from us_video_medsam2_b200.pipeline import RemoteEncoders

            pred.attach_remote_encoders(RemoteEncoders(list(range(1, world)), dev))
        out_host = torch.empty((T, B, 512, 512), dtype=torch.bool).pin_memory()  # binary masks, 1 byte per pixel
        clip_dev = ops.normalize_gray_u8(gray_host.to(dev), synth.IMG_MEAN, synth.IMG_STD)  # resident copy for `value`

        copy_stream = torch.cuda.Stream()  # device -> host result copies run beside the next frame, as a driver would do

        def one_pass(images, sink=None):
            st = pred.init_state(images, 512, 512)
            for i, m in enumerate(masks):
                pred.add_new_mask(st, 0, i + 1, m)
            n = 0
            main = torch.cuda.current_stream()
            for t, ids, logits in pred.propagate_in_video(st):
                if sink is not None:
                    m = logits[:, 0] > 0            # this frame's binary masks (a fresh tensor, not the graph's static output)
                    copy_stream.wait_stream(main)
                    with torch.cuda.stream(copy_stream):
                        sink[t].copy_(m, non_blocking=True)
                    m.record_stream(copy_stream)
                n += 1
            if sink is not None:
                main.wait_stream(copy_stream)       # the step ends when its last mask is in host memory
            return n

        def step_resident():
            if clip_mode and rank > 0:  # encoder rank: serve the one plan rank 0 announces for this pass
                pred.serve_encoder(clip_dev, rank - 1, world - 1, dst=0, max_plans=1)
                return 0
            return one_pass(clip_dev)

        def step_e2e():
            g = gray_host.to(dev, non_blocking=True)
            imgs = ops.normalize_gray_u8(g, synth.IMG_MEAN, synth.IMG_STD)
            if clip_mode and rank > 0:
                pred.serve_encoder(imgs, rank - 1, world - 1, dst=0, max_plans=1)
                return 0
            return one_pass(imgs, out_host)

        def timed(fn, steps, warmup):
            sampler = ClockSampler(local)
            if rank == 0 and not os.environ.get("USVM2_NO_SAMPLER"):
                sampler.start()
            for _ in range(warmup):
                fn()
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            sampler.mark_begin()
            l0 = _lib.launch_count
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            frames = 0
            marks = []
            allocs0 = torch.cuda.memory_stats().get("num_device_alloc", 0)
            for i in range(steps):
                t_host = time.perf_counter()
                frames += fn()
                if os.environ.get("USVM2_BENCH_DEBUG"):
                    m = torch.cuda.Event(enable_timing=True)
                    m.record()
                    marks.append((m, 1e3 * (time.perf_counter() - t_host)))
            e1.record()
            torch.cuda.synchronize()
            if os.environ.get("USVM2_BENCH_DEBUG"):
                print(f"[bench debug] {fn.__name__}: cudaMalloc calls inside the timed region: "
                      f"{torch.cuda.memory_stats().get('num_device_alloc', 0) - allocs0}", file=sys.stderr, flush=True)
            prev = e0
            for i, (m, host_ms) in enumerate(marks):
                print(f"[bench debug] {fn.__name__} step {i}: host loop {host_ms:.1f} ms, device {prev.elapsed_time(m):.1f} ms",
                      file=sys.stderr, flush=True)
                prev = m
            sampler.mark_end()
            launches = _lib.launch_count - l0
            clocks = sampler.stop() if rank == 0 else None
            if world > 1:
                dist.barrier()
            ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
            fr = torch.tensor([float(frames)], device=dev)
            if world > 1:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
                dist.all_reduce(fr, op=dist.ReduceOp.SUM)
            return float(ms), float(fr), launches, clocks

        with torch.inference_mode():
            ms, frames, launches, clocks = timed(step_resident, steps, warmup)
            ms_e2e, frames_e2e, _, _ = timed(step_e2e, max(1, steps), max(1, min(warmup, 2)))
        # (encoder ranks serve exactly one plan per step -- max_plans=1 -- so no shutdown header is owed to them)
        return dict(pred=pred, ms=ms, value=frames / (ms / 1000.0), e2e=frames_e2e / (ms_e2e / 1000.0),
                    launches=launches, clocks=clocks)

    clip_mode = args.mode == "clip" and world > 1
    res = measure(clip_mode, args.steps, args.warmup)
    pred, ms, value, e2e, launches, clocks = (res[k] for k in ("pred", "ms", "value", "e2e", "launches", "clocks"))
    # Strong scaling of ONE clip (BASELINE configs[1], second half) in the same invocation: after the independent-videos
    # measurement the same ranks track a single clip -- rank 0 propagates, the others run the frame-parallel encoder and
    # send features point-to-point over NCCL.  speedup_vs_1gpu compares with one GPU's own clip rate measured above.
    clip_leg = None
    if world > 1 and not clip_mode and args.model == "hiera_t512" and not args.no_clip_leg:
        cres = measure(True, max(2, args.steps // 4), 2)
        clip_leg = {"workload": f"one {T}-frame clip, {B} object(s): propagation on rank 0, frame-parallel image encoder on "
                                f"{world - 1} rank(s), features point-to-point over NCCL (4 MiB per frame)",
                    "scaling": "strong", "n_gpus": world, "value": cres["value"], "unit": "frames/s",
                    "e2e": cres["e2e"], "ms_per_step": cres["ms"] / max(2, args.steps // 4),
                    "speedup_vs_1gpu": cres["value"] / (value / world),
                    "amdahl_bound": "(encoder + tracked frame) / tracked frame: the propagation is sequential in t",
                    "encoder_batch": args.encoder_batch}
        del cres
        torch.cuda.empty_cache()


    peaks = measured_peaks()
    ks = time_dominant_kernel(dev, B) if rank == 0 else None
    roofline = None
    if ks:
        ach = ks["flops_per_launch"] / (ks["avg_ms"] * 1e-3) / 1e12
        roofline = {"bound": "tensor", "achieved": ach, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
                    "frac": ach / peaks["bf16_sustained"],
                    "traffic": None,  # filled from the committed ncu capture below
                    "kernel": "fmha_tc5_ts_kernel + fmha_combine_kernel<256> (memory-attention cross-attention, "
                              f"1024 x 7232 keys, d=256, {ks['splits']}-way split-KV)",
                    "launches_timed": ks["launches"], "avg_us": ks["avg_ms"] * 1e3,
                    "avg_us_cold_l2": ks["cold_ms"] * 1e3, "peak_source": peaks["source"],
                    "how": "CUDA events; the kernel pair in context (graph of its producer GEMMs + the pair minus graph of "
                           "the producers alone, L2 flushed before every replay: operands as warm as in the frame, which "
                           "itself replays from a CUDA graph); avg_us_cold_l2 = the pair alone after an L2 flush"}
    if roofline is not None:
        traffic, src = ncu_traffic("fmha_tc5_ts_kernel")
        roofline["traffic"] = traffic if B == 1 else None
        roofline["traffic_source"] = src
        roofline["other_kernels"] = time_other_kernels(dev, B)
        if args.batched_videos > 0:  # the same kernel where a roofline fraction is meaningful: the batched shape
            roofline["throughput_shape"] = time_dominant_kernel_batched(dev, peaks)
        # whole path by SURVEY 8(d)'s formula: (62.9 + 52.0 * B) GFLOP per frame x the measured frames/s of one GPU
        wp = (GF_ENCODER + GF_PER_OBJECT * B) * (value / world) / 1e3
        roofline["whole_path"] = {"achieved": wp, "unit": "TFLOP/s", "frac": wp / peaks["bf16_sustained"],
                                  "formula": f"({GF_ENCODER} + {GF_PER_OBJECT} x {B}) GFLOP per frame x frames/s per GPU"}
        with torch.inference_mode():
            roofline["components"] = time_components(
                pred, dev, B, peaks, batched_objects=args.batched_videos * args.batched_objects)
    batched = None
    if args.batched_videos > 0 and not clip_mode and args.model == "hiera_t512":
        with torch.inference_mode():
            batched = run_batched_leg(pred, args, dev, rank, world, peaks)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sample = max(4, min(T, args.cpu_sample_frames))
        fps, cores = cpu_reference_fps(sample, B, model=args.model)
        cpu = {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
               "sample": f"{sample}-frame sample of the same clip, oracle port of the reference, torch "
                         f"{torch.__version__} CPU fp32"}
    if rank == 0:
        print(json.dumps({
            "metric": metric_name(B, args.model), "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if clip_mode else "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"{MODEL_NAMES[args.model]} "
                                   f"propagate_in_video, {T}-frame synthetic echo clip "
                                   f"{'(one clip for all GPUs)' if clip_mode else 'per GPU'}, "
                                   f"{B} object(s), 7-frame memory bank, mask prompt on frame 0",
                       "frames": T, "objects": B, "encoder_batch": args.encoder_batch,
                       "parallelism": (f"one clip: propagation on rank 0, frame-parallel encoder on {world - 1} rank(s), "
                                       "features sent point-to-point over NCCL") if clip_mode else f"videos x{world}",
                       "encoder_sms": (pred._partition_obj.sms if pred._partition_obj is not None else 0),
                       "precision": "bf16 tensor-core contractions (encoder / memory attention / memory encoder), "
                                    "fp32-operand mask decoder (tf32 tensor-core products on its image-side GEMMs, "
                                    "exact fp32 on the token side)",
                       "l2": f"inputs larger than L2: {T * 3 * 512 * 512 * 4 / 1e6:.0f} MB clip streamed once per step"},
            "e2e": {"value": e2e, "unit": "frames/s", "h2d_bytes_per_step": T * 512 * 512,
                    "d2h_bytes_per_step": T * B * 512 * 512},
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu, "batched": batched,
            "clip_mode": clip_leg,
        }))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if os.environ.get("USVM2_NO_GC"):
        import gc
        gc.disable()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl b200) needs a CUDA device; there is no CPU fallback")
    if args.model == "hiera_b+":
        if rank == 0:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import bench_bplus

            print(json.dumps(bench_bplus.run(256 if args.frames == 512 else args.frames, args.steps, max(1, args.warmup // 3))))
        return
    run_b200(args, rank, world)


if __name__ == "__main__":
    main()
