from us_video_medsam2_b200.engine import NO_OBJ_SCORE  # noqa: F401
