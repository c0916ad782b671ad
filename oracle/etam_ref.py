"""TEST INFRASTRUCTURE -- CPU restatement of the EfficientTAM variant of the path (BASELINE configs[3]):
`efficienttam_ti_512x512.yaml` = plain ViT-tiny trunk (12 blocks x 192-d, 3 heads of 64, 14 x 14 windows except the global
blocks 2 / 5 / 8 / 11) + ViTDetNeck, feeding the SAME memory attention / memory encoder / mask decoder as MedSAM2 with
three switches off: no high-resolution decoder features, no temporal encoding on object pointers, no
`no_obj_embed_spatial`.  Everything that is shared is inherited from oracle/medsam2_ref.py; only the differences are
restated here, each citing the reference file:line it follows (paths relative to /root/reference).

Pinned to the reference's own outputs by tests/test_oracle_pinned.py (fixture tests/golden/etam_ti_mask_fwd.npz, produced
by oracle/make_golden_etam.py from the unmodified reference).  The CUDA path for this variant is not built yet: this
file is the checker it will be held to."""
import json
import os

import torch
import torch.nn.functional as F

from oracle.medsam2_ref import Cfg, RefModel, RefPredictor, conv2d, layer_norm_2d, linear, sdpa



class EtamCfg(Cfg):
    """efficient_track_anything/configs/efficienttam_ti_512x512.yaml (+ builder overrides,
    build_efficienttam.py:117-128)."""
    patch = 16
    vit_dim = 192
    vit_depth = 12
    vit_heads = 3
    vit_window = 14
    vit_window_blocks = (0, 1, 3, 4, 6, 7, 9, 10)
    use_high_res_features_in_sam = False
    add_tpos_enc_to_obj_ptrs = False
    no_obj_embed_spatial = False


class EtamSCfg(EtamCfg):
    """efficient_track_anything/configs/efficienttam_s_512x512.yaml: ViT-small trunk (384-d, 6 heads of 64)."""
    vit_dim = 384
    vit_heads = 6


def etam_state_dict_abi(variant="ti"):
    from us_video_medsam2_b200.synth import etam_state_dict_abi as abi

    return abi(variant)


def make_etam_state_dict(seed=0, variant="ti"):
    """Seeded weights for every tensor of the ABI (same drawing rules as synth.make_state_dict)."""
    from us_video_medsam2_b200.synth import make_etam_state_dict as make

    return make(seed, variant)


class RefModelETAM(RefModel):
    """EfficientTAMBase (modeling/efficienttam_base.py) restated as a delta on RefModel."""

    # ---------------- image encoder: ViT trunk + ViTDetNeck ----------------
    def _abs_pos(self, h, w):
        """get_abs_pos (backbones/utils.py:97-128): drop the cls token, bicubic-resize the 14 x 14 grid."""
        pe = self.p("image_encoder.trunk.pos_embed")[:, 1:]
        size = int(round(pe.shape[1] ** 0.5))
        if (size, size) != (h, w):
            pe = F.interpolate(pe.reshape(1, size, size, -1).permute(0, 3, 1, 2), size=(h, w), mode="bicubic",
                               align_corners=False).permute(0, 2, 3, 1)
            return pe
        return pe.reshape(1, h, w, -1)

    def vit(self, img):
        """ViT.forward + Block.forward + Attention.forward (backbones/vitdet.py:282-299, 148-163, 57-79)."""
        cfg = self.cfg
        t = "image_encoder.trunk."
        x = conv2d(img, self.p(t + "patch_embed.proj.weight"), self.p(t + "patch_embed.proj.bias"), stride=cfg.patch)
        x = x.permute(0, 2, 3, 1)
        x = x + self._abs_pos(x.shape[1], x.shape[2])
        heads = cfg.vit_heads
        for i in range(cfg.vit_depth):
            b = t + f"blocks.{i}."
            shortcut = x
            h = F.layer_norm(x, (x.shape[-1],), self.p(b + "norm1.weight"), self.p(b + "norm1.bias"), 1e-6)
            H, W = h.shape[1], h.shape[2]
            ws = cfg.vit_window if i in cfg.vit_window_blocks else 0
            if ws:
                h, pad_hw = self._to_windows(h, ws)
            Bw, Hh, Ww, C = h.shape
            qkv = linear(h, self.p(b + "attn.qkv.weight"), self.p(b + "attn.qkv.bias"))
            qkv = qkv.reshape(Bw, Hh * Ww, 3, heads, C // heads).permute(2, 0, 3, 1, 4)
            o = sdpa(qkv[0], qkv[1], qkv[2])
            o = o.permute(0, 2, 1, 3).reshape(Bw, Hh, Ww, C)
            o = linear(o, self.p(b + "attn.proj.weight"), self.p(b + "attn.proj.bias"))
            if ws:
                o = self._from_windows(o, ws, pad_hw, (H, W))
            x = shortcut + o
            h = F.layer_norm(x, (x.shape[-1],), self.p(b + "norm2.weight"), self.p(b + "norm2.bias"), 1e-6)
            h = F.gelu(linear(h, self.p(b + "mlp.layers.0.weight"), self.p(b + "mlp.layers.0.bias")))
            x = x + linear(h, self.p(b + "mlp.layers.1.weight"), self.p(b + "mlp.layers.1.bias"))
        return x.permute(0, 3, 1, 2)

    def forward_image(self, img):
        """ImageEncoder.forward + ViTDetNeck.forward (backbones/image_encoder.py:31-44, 93-108): 1x1 conv (no bias) ->
        LayerNorm2d -> 3x3 conv (no bias) -> LayerNorm2d; one feature level, no conv_s0 / conv_s1
        (use_high_res_features_in_sam: false)."""
        x = self.vit(img.float())
        n = "image_encoder.neck.convs.0."
        x = conv2d(x, self.p(n + "conv_1x1.weight"), None)
        x = layer_norm_2d(x, self.p(n + "norm_0.weight"), self.p(n + "norm_0.bias"))
        x = conv2d(x, self.p(n + "conv_3x3.weight"), None, padding=1)
        x = layer_norm_2d(x, self.p(n + "norm_1.weight"), self.p(n + "norm_1.bias"))
        B = x.shape[0]
        zero = lambda c, s: torch.zeros((B, c, s, s))
        # without high-res features the decoder's upscaling is dc1 -> LN -> act -> dc2 -> act (mask_decoder.py:221-227):
        # the restated decoder adds feat_s1 / feat_s0, so zeros reproduce it exactly
        return dict(feat_s0=zero(32, 4 * x.shape[2]), feat_s1=zero(64, 2 * x.shape[2]), feat=x,
                    pos=self.sine_pos(x.shape[2], x.shape[3], 256))


def etam_predictor(state_dict, fill_holes=True, cfg=EtamCfg):
    """RefPredictor over the EfficientTAM model.  The reference's EfficientTAM predictor keeps per-object state and runs
    objects one at a time (efficienttam_video_predictor.py:592-628); objects are independent on this path, so the batched
    session logic of RefPredictor yields the same masks per object."""
    return RefPredictor(state_dict, cfg=cfg, fill_holes=fill_holes, model_cls=RefModelETAM)
