from us_video_medsam2_b200.build_etam import (build_efficienttam_video_predictor,  # noqa: F401
                                              build_efficienttam_video_predictor_npz)
