"""Mirror of sam2/utils/transforms.py (SAM2Transforms): image -> model input, prompt coordinate transforms and the
mask post-processing that is the second caller of the connected-components op (transforms.py:76-115): holes of the
background up to `max_hole_area` pixels become foreground, foreground "sprinkles" up to `max_sprinkle_area` pixels become
background, then the masks are resized to the original image size.  Unlike the reference there is no
try/except-and-skip around the CUDA op: a failing kernel raises."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .cc import get_connected_components
from .synth import IMG_MEAN, IMG_STD


class SAM2Transforms(nn.Module):
    def __init__(self, resolution, mask_threshold, max_hole_area=0.0, max_sprinkle_area=0.0):
        super().__init__()
        self.resolution = resolution
        self.mask_threshold = mask_threshold
        self.max_hole_area = max_hole_area
        self.max_sprinkle_area = max_sprinkle_area
        self.mean = list(IMG_MEAN)
        self.std = list(IMG_STD)

    def __call__(self, x):
        """HWC uint8 array / PIL image -> normalised fp32 [3, S, S] (ToTensor, Resize((S, S)), Normalize of the
        reference: transforms.py:29-40)."""
        from torchvision.transforms import Normalize, Resize, ToTensor

        t = ToTensor()(x)
        return Normalize(self.mean, self.std)(Resize((self.resolution, self.resolution))(t))

    def forward_batch(self, img_list):
        return torch.stack([self(img) for img in img_list], dim=0)

    def transform_coords(self, coords, normalize=False, orig_hw=None):
        """Absolute image coordinates (normalize=True, needs orig_hw) or [0, 1] coordinates -> model pixels."""
        if normalize:
            assert orig_hw is not None
            h, w = orig_hw
            coords = coords.clone()
            coords[..., 0] = coords[..., 0] / w
            coords[..., 1] = coords[..., 1] / h
        return coords * self.resolution

    def transform_boxes(self, boxes, normalize=False, orig_hw=None):
        return self.transform_coords(boxes.reshape(-1, 2, 2), normalize, orig_hw)

    def postprocess_masks(self, masks, orig_hw):
        masks = masks.float()
        mask_flat = masks.flatten(0, 1).unsqueeze(1)  # one 1-channel image per (batch, mask)
        if self.max_hole_area > 0:
            labels, areas = get_connected_components(mask_flat <= self.mask_threshold)
            is_hole = ((labels > 0) & (areas <= self.max_hole_area)).reshape_as(masks)
            masks = torch.where(is_hole, self.mask_threshold + 10.0, masks)
        if self.max_sprinkle_area > 0:
            labels, areas = get_connected_components(mask_flat > self.mask_threshold)
            is_sprinkle = ((labels > 0) & (areas <= self.max_sprinkle_area)).reshape_as(masks)
            masks = torch.where(is_sprinkle, self.mask_threshold - 10.0, masks)
        return F.interpolate(masks, orig_hw, mode="bilinear", align_corners=False)
