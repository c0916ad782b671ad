"""Frame ingest for `init_state(video_path)`: a folder of "<frame_index>.jpg" files -> normalised fp32
[T,3,S,S] (reference semantics: sam2/utils/misc.py:92-277 -- PIL decode, RGB, resize to SxS, /255,
ImageNet mean/std).  Decode and resize stay on the host (PIL); MP4 input needs `decord`, which this
image does not have, and raises like the reference does for unsupported inputs."""
import os

import numpy as np
import torch

from .synth import IMG_MEAN, IMG_STD


def load_video_frames(video_path, image_size, offload_video_to_cpu, compute_device):
    is_str = isinstance(video_path, str)
    if isinstance(video_path, bytes) or (is_str and os.path.splitext(video_path)[-1] in (".mp4", ".MP4")):
        raise NotImplementedError("MP4 input needs the `decord` package; extract JPEG frames into a folder instead")
    if not (is_str and os.path.isdir(video_path)):
        raise NotImplementedError("Only MP4 video and JPEG folder are supported at this moment")
    from PIL import Image

    names = sorted(p for p in os.listdir(video_path) if os.path.splitext(p)[-1] in (".jpg", ".jpeg", ".JPG", ".JPEG"))
    if not names:
        raise RuntimeError(f"no images found in {video_path}")
    images = torch.zeros(len(names), 3, image_size, image_size, dtype=torch.float32)
    video_height = video_width = None
    for n, name in enumerate(names):
        img = Image.open(os.path.join(video_path, name))
        arr = np.array(img.convert("RGB").resize((image_size, image_size)))
        if arr.dtype != np.uint8:
            raise RuntimeError(f"Unknown image dtype: {arr.dtype} on {name}")
        images[n] = torch.from_numpy(arr / 255.0).permute(2, 0, 1)
        video_width, video_height = img.size
    mean = torch.tensor(IMG_MEAN, dtype=torch.float32)[:, None, None]
    std = torch.tensor(IMG_STD, dtype=torch.float32)[:, None, None]
    if not offload_video_to_cpu:
        images, mean, std = images.to(compute_device), mean.to(compute_device), std.to(compute_device)
    images -= mean
    images /= std
    return images, video_height, video_width
