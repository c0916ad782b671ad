"""`from sam2 import _C; _C.get_connected_componnets(x)` -- the reference's native op
(sam2/csrc/connected_components.cu:213-289; the misspelt name is part of the contract)."""
from us_video_medsam2_b200.cc import get_connected_componnets  # noqa: F401
