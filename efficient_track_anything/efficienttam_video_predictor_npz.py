from us_video_medsam2_b200.predictor import EfficientTAMVideoPredictorNPZ  # noqa: F401
