"""Alias package: the reference's import sites for EfficientTAM (`efficient_track_anything.build_efficienttam`,
`efficient_track_anything.efficienttam_video_predictor[_npz]`) resolve to the B200 implementation."""
