"""TEST INFRASTRUCTURE ONLY -- imports the *real* reference from /root/reference.

Only usable in the build container (the GPU box has no /root/reference).  Used by
oracle/make_golden.py to produce the committed fixtures under tests/golden/ and by
`-m "not gpu"` tests (skipped when the mount is absent) to pin oracle/medsam2_ref.py
against the reference's own PyTorch code.

hydra-core / omegaconf / iopath are not installed, so the three packages are stubbed in
sys.modules and the YAML is instantiated by a small recursive `_target_` walker that applies
the builder's overrides by hand (reference: sam2/build_sam.py:95-173).
"""
import importlib
import os
import sys
import types

import torch
import yaml

REF_ROOT = os.environ.get("USVM2_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REF_ROOT, "sam2"))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _install_shims():
    class _GlobalHydra:
        @staticmethod
        def instance():
            return types.SimpleNamespace(is_initialized=lambda: True)

    _stub("hydra", compose=None, initialize_config_module=lambda *a, **k: None)
    _stub("hydra.utils", instantiate=None)
    _stub("hydra.core")
    _stub("hydra.core.global_hydra", GlobalHydra=_GlobalHydra)
    _stub("omegaconf", OmegaConf=object)
    _stub("iopath")
    _stub("iopath.common")
    _stub("iopath.common.file_io", g_pathmgr=None)


def _instantiate(node):
    if isinstance(node, dict):
        kw = {k: _instantiate(v) for k, v in node.items() if k != "_target_"}
        if "_target_" in node:
            mod, cls = node["_target_"].rsplit(".", 1)
            return getattr(importlib.import_module(mod), cls)(**kw)
        return kw
    if isinstance(node, list):
        return [_instantiate(x) for x in node]
    if isinstance(node, str):
        try:
            return float(node)  # PyYAML reads "1e-6" as str, OmegaConf as float
        except ValueError:
            return node
    return node


def bplus_overrides():
    """sam2.1_hiera_base_plus at 1024^2 (BASELINE configs[4]) as overrides of the shipped tiny-512 YAML: the reference has
    no B+ config, its `Hiera` class defaults (hieradet.py:174-200) are the B+ stage layout; embed_dim 112 / 2 heads are
    upstream's values (80.85 M parameters, the published size of the B+ checkpoint)."""
    cfg = yaml.safe_load(open(os.path.join(REF_ROOT, "sam2/configs/sam2.1_hiera_t512.yaml")))["model"]
    enc, att = cfg["image_encoder"], cfg["memory_attention"]
    enc["trunk"].update(embed_dim=112, num_heads=2, stages=[2, 3, 16, 3], global_att_blocks=[12, 16, 20],
                        window_pos_embed_bkg_spatial_size=[14, 14])
    enc["neck"]["backbone_channel_list"] = [896, 448, 224, 112]
    att["layer"]["self_attention"]["feat_sizes"] = [64, 64]
    att["layer"]["cross_attention"]["feat_sizes"] = [64, 64]
    return dict(image_encoder=enc, memory_attention=att, image_size=1024)


def load_reference_predictor(npz=True, apply_postprocessing=True, seed=0, overrides=None):
    """Build the reference's SAM2VideoPredictor[NPZ] for sam2.1_hiera_t512 on CPU."""
    assert reference_available(), "reference mount missing"
    # the repo ships its own drop-in `sam2` package; the reference must win in this process
    for name in [k for k in sys.modules if k == "sam2" or k.startswith("sam2.")]:
        mod = sys.modules[name]
        if not getattr(mod, "__file__", "") or not str(mod.__file__).startswith(REF_ROOT):
            del sys.modules[name]
    if REF_ROOT in sys.path:
        sys.path.remove(REF_ROOT)
    sys.path.insert(0, REF_ROOT)
    _install_shims()
    cfg = yaml.safe_load(open(os.path.join(REF_ROOT, "sam2/configs/sam2.1_hiera_t512.yaml")))["model"]
    cfg["_target_"] = (
        "sam2.sam2_video_predictor_npz.SAM2VideoPredictorNPZ" if npz
        else "sam2.sam2_video_predictor.SAM2VideoPredictor"
    )
    if apply_postprocessing:
        cfg["sam_mask_decoder_extra_args"] = dict(
            dynamic_multimask_via_stability=True,
            dynamic_multimask_stability_delta=0.05,
            dynamic_multimask_stability_thresh=0.98,
        )
        cfg["binarize_mask_from_pts_for_mem_enc"] = True
        cfg["fill_hole_area"] = 8
    cfg.update(overrides or {})
    torch.manual_seed(seed)
    model = _instantiate(cfg).eval()
    return model


def load_reference_etam_predictor(apply_postprocessing=True, seed=0, config="efficienttam_ti_512x512.yaml"):
    """Build the reference's EfficientTAMVideoPredictorNPZ (efficient_track_anything/build_efficienttam.py:175-222) on CPU
    with the same shims; `compile_image_encoder` is forced off as the builder does without a capable GPU (:185-188)."""
    assert reference_available(), "reference mount missing"
    # the repo ships its own drop-in `efficient_track_anything` alias package; the reference must win in this process
    for name in [k for k in sys.modules if k == "efficient_track_anything" or k.startswith("efficient_track_anything.")]:
        mod = sys.modules[name]
        if not getattr(mod, "__file__", "") or not str(mod.__file__).startswith(REF_ROOT):
            del sys.modules[name]
    if REF_ROOT in sys.path:
        sys.path.remove(REF_ROOT)
    sys.path.insert(0, REF_ROOT)
    _install_shims()
    cfg = yaml.safe_load(open(os.path.join(REF_ROOT, "efficient_track_anything/configs", config)))["model"]
    cfg["_target_"] = "efficient_track_anything.efficienttam_video_predictor_npz.EfficientTAMVideoPredictorNPZ"
    cfg["compile_image_encoder"] = False
    if apply_postprocessing:
        cfg["sam_mask_decoder_extra_args"] = dict(
            dynamic_multimask_via_stability=True,
            dynamic_multimask_stability_delta=0.05,
            dynamic_multimask_stability_thresh=0.98,
        )
        cfg["binarize_mask_from_pts_for_mem_enc"] = True
        cfg["fill_hole_area"] = 8
    torch.manual_seed(seed)
    return _instantiate(cfg).eval()
