"""B200-native implementation of MedSAM2's per-frame video propagation path (sam2.1_hiera_t512).

    from us_video_medsam2_b200.build_sam import build_sam2_video_predictor_npz
    predictor = build_sam2_video_predictor_npz("configs/sam2.1_hiera_t512.yaml", ckpt_path)
    state = predictor.init_state(images, H, W)
    predictor.add_new_mask(state, 0, 1, mask)
    for frame_idx, obj_ids, logits in predictor.propagate_in_video(state): ...

Also here: `build_etam` (EfficientTAM-ti builders), `image_predictor` (SAM2ImagePredictor), `transforms` (SAM2Transforms),
`pipeline` (look-ahead image encoder on an SM partition or on other GPUs), `sharding` (independent videos across ranks).

Importing the package does not load the CUDA library; the first kernel call does, and raises if
csrc/libusvm2_b200.so is missing (there is no fallback path).
"""
__all__ = ["build_sam", "build_etam", "predictor", "image_predictor", "transforms", "pipeline", "sharding", "engine", "ops",
           "synth", "cc"]
