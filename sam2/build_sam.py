from us_video_medsam2_b200.build_sam import (  # noqa: F401
    build_sam2, build_sam2_video_predictor, build_sam2_video_predictor_npz, get_best_available_device, _load_checkpoint)
