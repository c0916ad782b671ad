from us_video_medsam2_b200.predictor import EfficientTAMVideoPredictor  # noqa: F401
