"""Builders with the reference's signatures (sam2/build_sam.py:95-173), without Hydra.

`config_file` is accepted as the reference spells it ("configs/sam2.1_hiera_t512.yaml"); only its basename
selects the architecture.  `hydra_overrides_extra` entries of the form "++model.<key>=<value>" are applied to
the model keyword arguments the same way the reference's Hydra overrides would be.
"""
import logging
import os

import torch
import yaml

from .predictor import SAM2VideoPredictor, SAM2VideoPredictorNPZ

_CONFIG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "configs")

# flags of the flat YAML that select code paths this build implements (all fixed by the shipped config)
_REQUIRED_TRUE = ("use_mask_input_as_output_without_sam", "directly_add_no_mem_embed", "no_obj_embed_spatial",
                  "use_high_res_features_in_sam", "multimask_output_in_sam", "iou_prediction_use_sigmoid",
                  "use_obj_ptrs_in_encoder", "add_tpos_enc_to_obj_ptrs", "proj_tpos_enc_in_obj_ptrs",
                  "use_signed_tpos_enc_to_obj_ptrs", "only_obj_ptrs_in_the_past_for_eval", "pred_obj_scores",
                  "pred_obj_scores_mlp", "fixed_no_obj_ptr", "use_multimask_token_for_obj_ptr",
                  "use_mlp_for_obj_ptr_proj")


# architecture constants the kernels are specialised for (overriding them would silently be ignored otherwise)
_FIXED_VALUES = {"num_maskmem": 7, "backbone_stride": 16, "max_obj_ptrs_in_encoder": 16}
_IMAGE_SIZE = {None: 512, "hiera_b+": 1024}  # per architecture variant (the `variant` key of the shipped YAML)


def get_best_available_device():
    """CUDA or nothing: this path has no CPU / MPS implementation (reference: build_sam.py:50-60)."""
    if torch.cuda.is_available():
        return "cuda"
    raise RuntimeError("us_video_medsam2_b200 needs a CUDA device (sm_100a); there is no CPU fallback")


def _parse_value(text):
    low = text.strip().lower()
    if low in ("true", "false"):
        return low == "true"
    try:
        return int(text)
    except ValueError:
        try:
            return float(text)
        except ValueError:
            return text


def _load_model_kwargs(config_file, overrides):
    name = os.path.basename(str(config_file))
    path = os.path.join(_CONFIG_DIR, name)
    if not os.path.exists(path):
        raise FileNotFoundError(f"unknown config {config_file!r}: this build ships {os.listdir(_CONFIG_DIR)}")
    with open(path) as f:
        cfg = yaml.safe_load(f)["model"]
    flat = {k: v for k, v in cfg.items() if not isinstance(v, dict)}
    # overrides first (Hydra composes them into the config before anything is instantiated, build_sam.py:124-128) ...
    for ov in overrides:
        key, _, val = ov.lstrip("+").partition("=")
        parts = key.split(".")
        if parts[0] != "model" or len(parts) < 2:
            continue
        if parts[1] == "_target_":
            continue
        if parts[1] == "sam_mask_decoder_extra_args" and len(parts) == 3:
            flat.setdefault("sam_mask_decoder_extra_args", {})[parts[2]] = _parse_value(val)
        elif len(parts) == 2:
            flat[parts[1]] = _parse_value(val)
        else:
            raise NotImplementedError(f"override {ov!r} changes an architecture block; this build implements the "
                                      f"architecture of {name} as shipped")
    # ... then validate what the composed config asks for against what this build implements
    for flag in _REQUIRED_TRUE:
        if not flat.get(flag, False):
            raise NotImplementedError(f"config flag {flag}=false selects a code path outside this build")
    fixed = dict(_FIXED_VALUES, image_size=_IMAGE_SIZE[flat.get("variant")])
    for key, want in fixed.items():
        if key in flat and flat[key] != want:
            raise NotImplementedError(f"{key}={flat[key]!r} is outside this build (the kernels are specialised for "
                                      f"{key}={want!r}, the value in {name})")
    return {k: v for k, v in flat.items() if k not in _REQUIRED_TRUE and k not in ("backbone_stride",)}


_POSTPROCESSING_OVERRIDES = [
    "++model.sam_mask_decoder_extra_args.dynamic_multimask_via_stability=true",
    "++model.sam_mask_decoder_extra_args.dynamic_multimask_stability_delta=0.05",
    "++model.sam_mask_decoder_extra_args.dynamic_multimask_stability_thresh=0.98",
    "++model.binarize_mask_from_pts_for_mem_enc=true",
    "++model.fill_hole_area=8",
]


def _build(cls, config_file, ckpt_path, device, mode, hydra_overrides_extra, apply_postprocessing, **kwargs):
    device = device or get_best_available_device()
    logging.info(f"Using device: {device}")
    overrides = list(hydra_overrides_extra)
    if apply_postprocessing:  # appended after the caller's, so they win -- as in the reference (build_sam.py:111-123)
        overrides = overrides + _POSTPROCESSING_OVERRIDES
    model_kwargs = _load_model_kwargs(config_file, overrides)
    model_kwargs.update(kwargs)
    model = cls(**model_kwargs)
    _load_checkpoint(model, ckpt_path)
    model = model.to(device)
    if mode == "eval":
        model.eval()
    return model


def build_sam2(config_file, ckpt_path=None, device=None, mode="eval", hydra_overrides_extra=[],
               apply_postprocessing=True, **kwargs):
    """reference: sam2/build_sam.py:72-93 (returns the bare SAM2Base there).  Here the model object that owns the
    state-dict ABI and the kernel engine is the predictor class, so the same object serves `SAM2ImagePredictor(model)`."""
    extra = list(hydra_overrides_extra)
    if apply_postprocessing:  # only the stability fallback here (build_sam.py:76-83), not the video overrides
        extra = extra + _POSTPROCESSING_OVERRIDES[:3]
    return _build(SAM2VideoPredictor, config_file, ckpt_path, device, mode, extra, False, **kwargs)


def build_sam2_video_predictor(config_file, ckpt_path=None, device=None, mode="eval", hydra_overrides_extra=[],
                               apply_postprocessing=True, **kwargs):
    """reference: sam2/build_sam.py:95-133"""
    return _build(SAM2VideoPredictor, config_file, ckpt_path, device, mode, hydra_overrides_extra,
                  apply_postprocessing, **kwargs)


def build_sam2_video_predictor_npz(config_file, ckpt_path=None, device=None, mode="eval", hydra_overrides_extra=[],
                                   apply_postprocessing=True, **kwargs):
    """reference: sam2/build_sam.py:135-173"""
    return _build(SAM2VideoPredictorNPZ, config_file, ckpt_path, device, mode, hydra_overrides_extra,
                  apply_postprocessing, **kwargs)


def _load_checkpoint(model, ckpt_path):
    """Strict load of ckpt["model"] (reference: build_sam.py:197-207)."""
    if ckpt_path is not None:
        sd = torch.load(ckpt_path, map_location="cpu", weights_only=True)["model"]
        missing_keys, unexpected_keys = model.load_state_dict(sd)
        if missing_keys:
            logging.error(missing_keys)
            raise RuntimeError()
        if unexpected_keys:
            logging.error(unexpected_keys)
            raise RuntimeError()
        logging.info("Loaded checkpoint sucessfully")
