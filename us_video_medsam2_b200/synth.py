"""Deterministic synthetic weights, clips and prompts shared by tests, bench and smoke.

There is no network for checkpoints or datasets, so parity and throughput are measured on
random-init weights of the named architecture (sam2.1_hiera_t512) and synthetic grayscale
ultrasound-like clips.  Everything here is a pure function of a seed and runs on the CPU
torch generator, so the build container (where the reference produces the golden fixtures)
and the GPU box (same image, same torch build) regenerate bit-identical tensors.

Why not the reference's default init: it leaves `pos_embed*` at zero, CXBlock `gamma` at 1e-6
and every norm at weight=1 / bias=0 (hieradet.py:221-226, memory_encoder.py:97-101), which
would hide bugs in exactly those terms.  Every tensor of the state-dict ABI
(state_dict_abi.json, dumped from the reference's `state_dict()`) is therefore re-drawn.
"""
import json
import math
import os

import torch

_ABI_PATH = os.path.join(os.path.dirname(__file__), "state_dict_abi.json")

IMG_MEAN = (0.485, 0.456, 0.406)  # sam2/utils/misc.py:176-177
IMG_STD = (0.229, 0.224, 0.225)


def state_dict_abi():
    """[(name, shape)] in the reference's state_dict order (471 tensors, 38.96 M params)."""
    with open(_ABI_PATH) as f:
        return [(k, tuple(s)) for k, s in json.load(f)]


_EMBED_STD1 = (
    "iou_token.weight", "mask_tokens.weight", "obj_score_token.weight",
    "not_a_point_embed.weight", "no_mask_embed.weight",
)


def _draw(name, shape, g):
    def randn(std=1.0):
        return torch.randn(shape, generator=g, dtype=torch.float32) * std

    def uniform(bound):
        return (torch.rand(shape, generator=g, dtype=torch.float32) * 2 - 1) * bound

    leaf = name.rsplit(".", 1)[-1]
    if name.endswith("positional_encoding_gaussian_matrix"):
        return randn(1.0)
    if name.endswith(_EMBED_STD1) or ".point_embeddings." in name:
        return randn(1.0)  # nn.Embedding default N(0, 1)
    if leaf == "gamma":  # CXBlock layer scale: make the Fuser branch matter
        return 0.1 + 0.4 * torch.rand(shape, generator=g, dtype=torch.float32)
    if leaf in ("pos_embed", "pos_embed_window"):
        return randn(0.05)
    if leaf in ("maskmem_tpos_enc", "no_mem_embed", "no_mem_pos_enc", "no_obj_ptr",
                "no_obj_embed_spatial"):
        return randn(0.02).clamp_(-0.04, 0.04)
    if leaf == "weight" and len(shape) == 1:  # LayerNorm / LayerNorm2d scale
        return 1.0 + randn(0.1)
    if leaf == "bias":
        return randn(0.02)
    if leaf == "weight":
        if "output_upscaling" in name and len(shape) == 4:  # ConvTranspose2d [Cin, Cout, 2, 2]
            fan_in = shape[1] * shape[2] * shape[3]
        else:
            fan_in = int(math.prod(shape[1:]))
        return uniform(1.0 / math.sqrt(fan_in))
    raise KeyError(f"no init rule for {name} {shape}")


def etam_state_dict_abi(variant="ti"):
    """[(name, shape)] of the reference's EfficientTAM state dict: "ti" = efficienttam_ti_512x512 (455 tensors, 17.87 M
    parameters), "s" = efficienttam_s_512x512 (455 tensors, 34.06 M)."""
    with open(os.path.join(os.path.dirname(__file__), f"etam_{variant}_state_dict_abi.json")) as f:
        return [(k, tuple(s)) for k, s in json.load(f)]


def make_etam_state_dict(seed=0, variant="ti"):
    """Seeded EfficientTAM weights (same drawing rules as make_state_dict)."""
    g = torch.Generator(device="cpu")
    g.manual_seed((7000003 if variant == "ti" else 9000011) * (seed + 1))
    return {name: _draw(name, shape, g) for name, shape in etam_state_dict_abi(variant)}


def bplus_state_dict_abi():
    """[(name, shape)] of sam2.1_hiera_base_plus at 1024^2 (BASELINE configs[4]; 615 tensors, 80.85 M parameters): the
    reference ships no such YAML, so the shapes were dumped from its own classes instantiated with the `Hiera` class
    defaults (hieradet.py:174-200: stages (2, 3, 16, 3), global blocks (12, 16, 20), 14 x 14 background position table)
    and upstream's embed_dim 112 / 2 heads (oracle/ref_loader.py: bplus_overrides)."""
    with open(os.path.join(os.path.dirname(__file__), "hiera_bplus_state_dict_abi.json")) as f:
        return [(k, tuple(s)) for k, s in json.load(f)]


def make_bplus_state_dict(seed=0):
    """Seeded Hiera-B+ weights (same drawing rules as make_state_dict)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(5000011 * (seed + 1))
    return {name: _draw(name, shape, g) for name, shape in bplus_state_dict_abi()}


def make_state_dict(seed=0):
    """Full sam2.1_hiera_t512 state dict, fp32 CPU, deterministic in `seed`."""
    g = torch.Generator(device="cpu")
    g.manual_seed(1000003 * (seed + 1))
    return {name: _draw(name, shape, g) for name, shape in state_dict_abi()}


def make_clip(num_frames, size=512, seed=1234, kind="speckle"):
    """Synthetic grayscale clip -> [T, 3, size, size] fp32, ImageNet-normalised exactly as the
    reference's JPEG loader does after resize (sam2/utils/misc.py:253-276).

    kind="uniform": i.i.d. U[0,1) pixels.
    kind="speckle": low-res echo texture, bilinearly upsampled, Rayleigh-ish speckle, slowly
                    drifting over time -- closer to an ultrasound cine loop.
    """
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    if kind == "uniform":
        gray = torch.rand((num_frames, 1, size, size), generator=g)
    else:
        base = torch.rand((1, 1, 16, 16), generator=g)
        drift = torch.randn((num_frames, 1, 16, 16), generator=g) * 0.03
        tex = (base + torch.cumsum(drift, dim=0)).clamp_(0, 1)
        tex = torch.nn.functional.interpolate(tex, size=(size, size), mode="bilinear",
                                              align_corners=False)
        u = torch.rand((num_frames, 1, size, size), generator=g).clamp_(1e-6, 1 - 1e-6)
        rayleigh = torch.sqrt(-2.0 * torch.log(1 - u)) * 0.35
        gray = (tex * rayleigh).clamp_(0, 1)
        gray = torch.round(gray * 255.0) / 255.0  # 8-bit pixels like a decoded JPEG
    mean = torch.tensor(IMG_MEAN, dtype=torch.float32)[None, :, None, None]
    std = torch.tensor(IMG_STD, dtype=torch.float32)[None, :, None, None]
    return (gray.expand(-1, 3, -1, -1) - mean) / std


def make_clip_u8(num_frames, size=512, seed=1234):
    """Same speckle clip as raw uint8 grayscale [T, size, size] (the on-disk form of an echo
    clip; used by the end-to-end bench leg, which normalises on the device)."""
    clip = make_clip(num_frames, size, seed, "speckle")
    gray = clip[:, 0] * IMG_STD[0] + IMG_MEAN[0]
    return torch.round(gray * 255.0).clamp_(0, 255).to(torch.uint8)


def box_mask(size=512, y0=200, y1=300, x0=220, x1=330):
    """Rectangular mask prompt of SURVEY 8d config 1."""
    m = torch.zeros((size, size), dtype=torch.bool)
    m[y0:y1, x0:x1] = True
    return m


def multi_object_masks(num_objects, size=512):
    """`num_objects` disjoint rectangles (BASELINE config 3: 4 objects per video)."""
    out = []
    for i in range(num_objects):
        r, c = divmod(i, 2)
        y0 = 60 + r * 220
        x0 = 70 + c * 230
        out.append(box_mask(size, y0, y0 + 120 + 10 * i, x0, x0 + 140 - 10 * i))
    return out
