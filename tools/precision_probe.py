"""Precision study (GPU): per-frame error of the CUDA path against reference fixtures under the library's precision
switches.  python tools/precision_probe.py [case]   (each configuration runs in its own process)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("TQDM_DISABLE", "1")

CONFIGS = {
    "default": {},
    "fp32 split partials": {"USVM2_FMHA_PART_BF16": "0"},
    "fp32 decoder GEMMs": {"USVM2_DEC_TF32": "0"},
    "both": {"USVM2_FMHA_PART_BF16": "0", "USVM2_DEC_TF32": "0"},
    "mma.sync attention": {"USVM2_FMHA": "mma"},
    "no split-K": {"USVM2_GEMM_KSPLIT": "0"},
}


def child():
    import numpy as np
    import torch

    from tests.golden_cases import dice
    from us_video_medsam2_b200 import synth
    from sam2.build_sam import build_sam2_video_predictor_npz

    out = {}
    for seed, T, name in ((19, 8, "t512_mask_fwd"), (19, 512, "t512_long_fwd")):
        g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
        pred = build_sam2_video_predictor_npz("configs/sam2.1_hiera_t512.yaml", encoder_batch=4)
        pred.load_state_dict(synth.make_state_dict(seed), strict=True)
        clip = synth.make_clip(T, kind="speckle")[:8].cuda()
        st = pred.init_state(clip, 512, 512)
        pred.add_new_mask(st, 0, 1, synth.box_mask())
        rows = []
        want = g["low_res_filled"][:, 0] if name == "t512_mask_fwd" else g["low"]
        frames = list(range(8)) if name == "t512_mask_fwd" else g["kept"].tolist()
        for t, _, _ in pred.propagate_in_video(st):
            if t == 0 or t not in frames:
                continue
            od = st["output_dict"]["non_cond_frame_outputs"][t]
            a = od["pred_masks"][0, 0].float().cpu()
            b = torch.from_numpy(want[frames.index(t)])
            same = (a != 0.1) & (b != 0.1)
            rows.append((t, float((a - b).abs()[same].mean()), float((a - b).abs()[same].max()), dice(a, b)))
        out[name] = rows
    print("RESULT " + json.dumps(out))


def main():
    for label, env in CONFIGS.items():
        e = dict(os.environ, **env, USVM2_PROBE_CHILD="1")
        r = subprocess.run([sys.executable, __file__], env=e, capture_output=True, text=True)
        line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
        if not line:
            print(label, "FAILED", r.stderr[-500:])
            continue
        res = json.loads(line[0][7:])
        for name, rows in res.items():
            print(f"{label:22s} {name:14s} " + " ".join(f"t{t}: {m * 1e4:.2f}e-4/{d:.4f}" for t, m, mx, d in rows), flush=True)


if __name__ == "__main__":
    child() if os.environ.get("USVM2_PROBE_CHILD") else main()
