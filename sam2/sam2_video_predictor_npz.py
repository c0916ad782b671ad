from us_video_medsam2_b200.predictor import SAM2VideoPredictorNPZ  # noqa: F401
