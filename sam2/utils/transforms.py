from us_video_medsam2_b200.transforms import SAM2Transforms  # noqa: F401
