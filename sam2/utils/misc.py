from us_video_medsam2_b200.cc import (  # noqa: F401
    fill_holes_in_mask_scores, get_connected_components)
from us_video_medsam2_b200.frames import load_video_frames  # noqa: F401
