"""Builders of the EfficientTAM variant, mirroring efficient_track_anything/build_efficienttam.py:93-222 (same names,
arguments and post-processing overrides; Hydra is not used -- shipped configurations: `efficienttam_ti_512x512` and
`efficienttam_s_512x512`).
`vos_optimized` selects a torch.compile'd predictor in the reference; here every kernel is already native, so the flag is
accepted and has no effect."""
import logging

import torch

from .build_sam import _load_checkpoint, _parse_value
from .predictor import EfficientTAMVideoPredictor, EfficientTAMVideoPredictorNPZ

from . import synth
from .engine import EtamSConfig

_CONFIGS = {"efficienttam_ti_512x512.yaml": "ti", "efficienttam_s_512x512.yaml": "s"}


def _variant(config_file):
    name = str(config_file).replace("\\", "/").rsplit("/", 1)[-1]
    if name not in _CONFIGS:
        raise FileNotFoundError(f"unknown config {config_file!r}: this build ships {sorted(_CONFIGS)}")
    return _CONFIGS[name]


def _predictor_class(cls, variant):
    """The class attributes that fix the architecture (config constants, state-dict ABI) per shipped configuration."""
    if variant == "ti":
        return cls
    return type(cls.__name__ + "S", (cls,), dict(_config_base=EtamSConfig,
                                                 _abi=staticmethod(lambda: synth.etam_state_dict_abi("s"))))


def _kwargs(config_file, overrides, apply_postprocessing):
    _variant(config_file)
    kw = {}
    if apply_postprocessing:  # build_efficienttam.py:117-128
        kw["sam_mask_decoder_extra_args"] = dict(dynamic_multimask_via_stability=True,
                                                 dynamic_multimask_stability_delta=0.05,
                                                 dynamic_multimask_stability_thresh=0.98)
        kw["binarize_mask_from_pts_for_mem_enc"] = True
        kw["fill_hole_area"] = 8
    for ov in overrides:
        key, _, val = ov.lstrip("+").partition("=")
        parts = key.split(".")
        if parts[0] != "model" or len(parts) < 2 or parts[1] in ("_target_", "compile_image_encoder"):
            continue
        if parts[1] == "sam_mask_decoder_extra_args" and len(parts) == 3:
            kw.setdefault("sam_mask_decoder_extra_args", {})[parts[2]] = _parse_value(val)
        elif len(parts) == 2:
            kw[parts[1]] = _parse_value(val)
    return kw


def _build(cls, config_file, ckpt_path, device, mode, hydra_overrides_extra, apply_postprocessing, **kwargs):
    logging.info(f"Using device: {device}")
    model_kwargs = _kwargs(config_file, list(hydra_overrides_extra), apply_postprocessing)
    model_kwargs.update(kwargs)
    model = _predictor_class(cls, _variant(config_file))(**model_kwargs)
    _load_checkpoint(model, ckpt_path)
    model = model.to(device)
    if mode == "eval":
        model.eval()
    return model


def build_efficienttam_video_predictor(config_file, ckpt_path=None, device="cuda", mode="eval", hydra_overrides_extra=[],
                                       apply_postprocessing=True, vos_optimized=False, **kwargs):
    """reference: efficient_track_anything/build_efficienttam.py:93-153"""
    return _build(EfficientTAMVideoPredictor, config_file, ckpt_path, device, mode, hydra_overrides_extra,
                  apply_postprocessing, **kwargs)


def build_efficienttam_video_predictor_npz(config_file, ckpt_path=None, device="cuda", mode="eval",
                                           hydra_overrides_extra=[], apply_postprocessing=True, vos_optimized=False,
                                           **kwargs):
    """reference: efficient_track_anything/build_efficienttam.py:175-222"""
    return _build(EfficientTAMVideoPredictorNPZ, config_file, ckpt_path, device, mode, hydra_overrides_extra,
                  apply_postprocessing, **kwargs)
