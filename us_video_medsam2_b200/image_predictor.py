"""Mirror of sam2/sam2_image_predictor.py (SAM2ImagePredictor): `set_image` / `predict` on the B200 kernel path.

`sam_model` is a predictor built by `build_sam2` / `build_sam2_video_predictor` of this package (it owns the state-dict ABI and
the kernel engine).  The compute is the same as a prompted first frame of the video path: image encoder, `+ no_mem_embed`
(directly_add_no_mem_embed, sam2_image_predictor.py:117-121), prompt encoder (points / box corners as labelled points /
dense mask input), two-way mask decoder with all 3 multimask outputs or the single-mask output, then
`SAM2Transforms.postprocess_masks`.  `set_image_batch` / `predict_batch` (reference :134-236) encode all images in one
batched (frame-parallel) encoder pass and predict image by image, as the reference does.  Not mirrored: `from_pretrained`."""
import numpy as np
import torch

from . import ops
from .transforms import SAM2Transforms


class SAM2ImagePredictor:
    def __init__(self, sam_model, mask_threshold=0.0, max_hole_area=0.0, max_sprinkle_area=0.0, **kwargs):
        self.model = sam_model
        self._transforms = SAM2Transforms(resolution=self.model.image_size, mask_threshold=mask_threshold,
                                          max_hole_area=max_hole_area, max_sprinkle_area=max_sprinkle_area)
        self.mask_threshold = mask_threshold
        self.reset_predictor()

    @property
    def device(self):
        return self.model.device

    def reset_predictor(self):
        self._is_image_set = False
        self._features = None
        self._orig_hw = None
        self._is_batch = False

    @torch.no_grad()
    def set_image(self, image):
        """image: HWC uint8 numpy array (RGB) or PIL image (reference :86-131)."""
        self.reset_predictor()
        if isinstance(image, np.ndarray):
            self._orig_hw = [image.shape[:2]]
        else:
            try:
                w, h = image.size
                self._orig_hw = [(h, w)]
            except Exception:
                raise NotImplementedError("Image format not supported")
        x = self._transforms(image)[None].to(self.device).float().contiguous()
        assert x.shape[-3] == 3 and x.dim() == 4, f"input_image must be of size 1x3xHxW, got {x.shape}"
        eng = self.model._sync_engine()
        f = eng.encode_frames(x)
        T = eng.cfg.feat ** 2
        pix, _ = ops.axpby(f["feat"][0], eng.w.no_mem_embed, rows=T, x_mod=T, y_mod=1)  # + no_mem_embed
        self._features = dict(pix=pix, feat_s0=f["feat_s0"][0], feat_s1=f["feat_s1"][0])
        self._is_image_set = True

    @torch.no_grad()
    def set_image_batch(self, image_list):
        """image_list: list of HWC uint8 numpy arrays (RGB) (reference :134-175): one batched encoder pass."""
        self.reset_predictor()
        assert isinstance(image_list, list)
        self._orig_hw = []
        for image in image_list:
            assert isinstance(image, np.ndarray), "Images are expected to be an np.ndarray in RGB format, and of shape  HWC"
            self._orig_hw.append(image.shape[:2])
        x = torch.stack([self._transforms(im) for im in image_list]).to(self.device).float().contiguous()
        assert x.dim() == 4 and x.shape[1] == 3, f"img_batch must be of size Bx3xHxW, got {x.shape}"
        eng = self.model._sync_engine()
        f = eng.encode_frames(x)
        n = x.shape[0]
        T = eng.cfg.feat ** 2
        pix, _ = ops.axpby(f["feat"].view(n * T, 256), eng.w.no_mem_embed, rows=n * T, y_mod=1)  # + no_mem_embed
        self._features = [dict(pix=pix[i * T:(i + 1) * T], feat_s0=f["feat_s0"][i], feat_s1=f["feat_s1"][i])
                          for i in range(n)]
        self._is_image_set = True
        self._is_batch = True

    def predict_batch(self, point_coords_batch=None, point_labels_batch=None, box_batch=None, mask_input_batch=None,
                      multimask_output=True, return_logits=False, normalize_coords=True):
        """Per-image `predict` over the batch set by `set_image_batch` -> (list of masks, list of ious, list of low-res
        logits) (reference :177-236)."""
        assert self._is_batch, "This function should only be used when in batched mode"
        if not self._is_image_set:
            raise RuntimeError("An image must be set with .set_image_batch(...) before mask prediction.")
        all_masks, all_ious, all_low = [], [], []
        for img_idx in range(len(self._features)):
            pick = lambda batch: batch[img_idx] if batch is not None else None
            mask_input, coords, labels, ubox = self._prep_prompts(pick(point_coords_batch), pick(point_labels_batch),
                                                                   pick(box_batch), pick(mask_input_batch),
                                                                   normalize_coords, img_idx=img_idx)
            masks, iou, low = self._predict(coords, labels, ubox, mask_input, multimask_output,
                                            return_logits=return_logits, img_idx=img_idx)
            all_masks.append(masks.squeeze(0).float().cpu().numpy())
            all_ious.append(iou.squeeze(0).float().cpu().numpy())
            all_low.append(low.squeeze(0).float().cpu().numpy())
        return all_masks, all_ious, all_low

    def get_image_embedding(self):
        if not self._is_image_set:
            raise RuntimeError("An image must be set with .set_image(...) to generate an embedding.")
        fs = self.model.cfg.feat
        if self._is_batch:
            return torch.stack([f["pix"].t().reshape(256, fs, fs) for f in self._features])
        return self._features["pix"].t().reshape(1, 256, fs, fs)

    def predict(self, point_coords=None, point_labels=None, box=None, mask_input=None, multimask_output=True,
                return_logits=False, normalize_coords=True):
        """-> (masks [C,H,W], iou predictions [C], low-res logits [C,S/4,S/4]) as numpy arrays (reference :238-305)."""
        if not self._is_image_set:
            raise RuntimeError("An image must be set with .set_image(...) before mask prediction.")
        assert not self._is_batch, "use predict_batch after set_image_batch"
        mask_input, coords, labels, ubox = self._prep_prompts(point_coords, point_labels, box, mask_input, normalize_coords)
        masks, iou, low = self._predict(coords, labels, ubox, mask_input, multimask_output, return_logits=return_logits)
        return (masks.squeeze(0).float().cpu().numpy(), iou.squeeze(0).float().cpu().numpy(),
                low.squeeze(0).float().cpu().numpy())

    def _prep_prompts(self, point_coords, point_labels, box, mask_logits, normalize_coords, img_idx=-1):
        coords = labels = ubox = mask_input = None
        if point_coords is not None:
            assert point_labels is not None, "point_labels must be supplied if point_coords is supplied."
            pc = torch.as_tensor(point_coords, dtype=torch.float, device=self.device)
            coords = self._transforms.transform_coords(pc, normalize=normalize_coords, orig_hw=self._orig_hw[img_idx])
            labels = torch.as_tensor(point_labels, dtype=torch.int, device=self.device)
            if coords.dim() == 2:
                coords, labels = coords[None], labels[None]
        if box is not None:
            b = torch.as_tensor(box, dtype=torch.float, device=self.device)
            ubox = self._transforms.transform_boxes(b, normalize=normalize_coords, orig_hw=self._orig_hw[img_idx])
        if mask_logits is not None:
            mask_input = torch.as_tensor(mask_logits, dtype=torch.float, device=self.device)
            if mask_input.dim() == 3:
                mask_input = mask_input[None]
        return mask_input, coords, labels, ubox

    @torch.no_grad()
    def _predict(self, point_coords, point_labels, boxes=None, mask_input=None, multimask_output=True,
                 return_logits=False, img_idx=-1):
        if not self._is_image_set:
            raise RuntimeError("An image must be set with .set_image(...) before mask prediction.")
        eng = self.model._sync_engine()
        concat = (point_coords, point_labels) if point_coords is not None else None
        if boxes is not None:  # box corners are points with labels 2 / 3, placed first (reference :392-404)
            bc = boxes.reshape(-1, 2, 2)
            bl = torch.tensor([[2, 3]], dtype=torch.int, device=boxes.device).repeat(bc.size(0), 1)
            if concat is not None:
                concat = (torch.cat([bc, concat[0]], dim=1), torch.cat([bl, concat[1]], dim=1))
            else:
                concat = (bc, bl)
        B = 1 if concat is None else concat[0].shape[0]
        if concat is not None:
            sparse = eng.embed_points(concat[0], concat[1])
        else:  # PromptEncoder with no points: an empty sparse embedding -> only the 6 output tokens
            sparse = torch.zeros((B, 0, 256), dtype=torch.float32, device=self.device)
        dense = None
        if mask_input is not None:
            m = mask_input.float()
            L = self.model.image_size // 4
            if tuple(m.shape[-2:]) != (L, L):
                m = ops.resize_bilinear_aa(m.contiguous(), L, L)
            dense = eng.embed_mask_prompt(m.contiguous(), B)
        f = self._features[img_idx] if self._is_batch else self._features
        pix = f["pix"] if B == 1 else f["pix"].repeat(B, 1)
        o = eng.sam_heads(pix, f["feat_s0"], f["feat_s1"], B, sparse, dense=dense, multimask=multimask_output)
        if multimask_output:
            low = o["masks"][:, 1:4]
            iou = torch.sigmoid(o["iou_logits"][:, 1:4])
        else:
            # mask token 0, or -- when the model was built with dynamic_multimask_via_stability -- the best multimask
            # output if token 0's mask is not stable (mask_decoder.py:160-166, 247-295).  This is the decoder's own output,
            # NOT gated by the object score (that gate belongs to the tracking heads, whose kernel also reports the
            # pointer token rather than the chosen mask), so the choice is re-derived here from the four masks.
            cfg = self.model.cfg
            rows = torch.arange(B, device=self.device)
            idx = torch.zeros(B, dtype=torch.long, device=self.device)
            if cfg.dynamic_multimask_via_stability:
                m0 = o["masks"][:, 0].flatten(1)
                d = cfg.dynamic_multimask_stability_delta
                area_i, area_u = (m0 > d).sum(-1).float(), (m0 > -d).sum(-1).float()
                stability = torch.where(area_u > 0, area_i / area_u, torch.ones_like(area_u))
                best = torch.argmax(o["iou_logits"][:, 1:4], dim=-1) + 1
                idx = torch.where(stability >= cfg.dynamic_multimask_stability_thresh, idx, best)
            low = o["masks"][rows, idx][:, None]
            iou = torch.sigmoid(o["iou_logits"])[rows, idx][:, None]
        masks = self._transforms.postprocess_masks(low, self._orig_hw[img_idx])
        low = torch.clamp(low, -32.0, 32.0)
        if not return_logits:
            masks = masks > self.mask_threshold
        return masks, iou, low
