"""Drop-in alias: `sam2.build_sam`, `sam2.sam2_video_predictor[_npz]`, `sam2._C`, `sam2.utils.misc` resolve to the
B200 implementation in `us_video_medsam2_b200` (same names as the reference package so its drivers import unchanged)."""
