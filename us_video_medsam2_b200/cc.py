"""Host-side mirror of the reference's connected-components interface.

  get_connected_componnets(inputs)      == sam2._C.get_connected_componnets
                                           (sam2/csrc/connected_components.cu:213-282)
  get_connected_components(mask)        == sam2/utils/misc.py:47-63
  fill_holes_in_mask_scores(mask, area) == sam2/utils/misc.py:312-338, but as ONE fused kernel
  get_largest_cc(segmentation)          == getLargestCC of the 3-D CT driver (medsam2_infer_3D_CT.py:76-79), on the device
"""
import torch

from . import ops


def get_connected_componnets(inputs):
    """uint8 CUDA [N,1,H,W] (H, W even) -> [labels int32, counts int32]; RuntimeError on a bad argument like the
    reference's AT_ASSERTM checks.  Asynchronous on the current stream."""
    labels, counts = ops.connected_components(inputs)
    return [labels, counts]


def get_connected_components(mask):
    labels, counts = ops.connected_components(mask.to(torch.uint8).contiguous())
    return labels, counts


def fill_holes_in_mask_scores(mask, max_area):
    """Background (score <= 0) components with area <= max_area become 0.1.  Unlike the reference there is no
    try/except-and-skip: a failing kernel raises."""
    assert max_area > 0, "max_area must be positive"
    return ops.fill_holes(mask.float(), max_area, 0.1)


def get_largest_cc(segmentation):
    """Largest 26-connected component of a binary volume [D,H,W] (CUDA tensor, any integer / bool dtype) as a bool
    tensor -- `getLargestCC` of medsam2_infer_3D_CT.py:76-79 (skimage.measure.label + bincount argmax) without the trip
    to the host.  The driver only calls it on non-empty volumes (:284); an empty one returns all False here."""
    return ops.largest_component_3d(segmentation).bool()
