from us_video_medsam2_b200.image_predictor import SAM2ImagePredictor  # noqa: F401
