"""Device-side model step of the propagation path, composed from the kernels in csrc/.

Mirrors the reference's SAM2Base methods (sam2/modeling/sam2_base.py:764-1681) for the shipped
configuration sam2.1_hiera_t512, but on packed weights and token-major activations:

  forward_image            -> Engine.encode_frames      (Hiera trunk + FpnNeck + conv_s0/s1)
  _prepare_memory_conditioned_features / MemoryAttention -> Engine.memory_attention
  _forward_sam_heads / _use_mask_as_output               -> Engine.sam_heads / Engine.mask_as_output
  _encode_new_memory / MemoryEncoder                     -> Engine.encode_memory

Precision plan (calibrated with oracle.medsam2_ref.set_matmul_emulation): the image
encoder, memory attention and memory encoder run their contractions on bf16 tensor cores with fp32
accumulation, fp32 residual stream, fp32 LayerNorm / softmax statistics; the SAM mask decoder and
everything after it (mask logits, IoU / object scores, object pointers) stays in fp32.
"""
import math
import os

import torch
import torch.nn.functional as F

from . import ops
from .ops import ACT_GELU, ACT_RELU, BF16, F32

NO_OBJ_SCORE = -1024.0
# precision study switch: 0 = the decoder's image-side GEMMs as exact fp32 products (SIMT kernel) instead of tf32
_DEC_TF32 = os.environ.get("USVM2_DEC_TF32", "1") != "0"


class ModelConfig:
    """Hyper-parameters of sam2/configs/sam2.1_hiera_t512.yaml (+ builder overrides, build_sam.py:108-122, i.e. the
    apply_postprocessing=True configuration; the predictor resets the three post-processing switches to the
    reference's class defaults when the builder does not pass them)."""
    image_size = 512
    embed_dim = 96
    num_heads = 1
    stages = (1, 2, 7, 2)
    global_att_blocks = (5, 7, 9)
    window_spec = (8, 4, 14, 7)
    q_pool = 3
    d_model = 256
    mem_dim = 64
    num_maskmem = 7
    max_obj_ptrs_in_encoder = 16
    max_cond_frames_in_attn = -1
    memory_temporal_stride_for_eval = 1
    sigmoid_scale_for_mem_enc = 20.0
    sigmoid_bias_for_mem_enc = -10.0
    multimask_output_in_sam = True
    multimask_output_for_tracking = True
    multimask_min_pt_num = 0
    multimask_max_pt_num = 1
    dynamic_multimask_via_stability = True
    dynamic_multimask_stability_delta = 0.05
    dynamic_multimask_stability_thresh = 0.98
    binarize_mask_from_pts_for_mem_enc = True
    non_overlap_masks_for_mem_enc = False
    arch = "hiera"
    use_high_res_features_in_sam = True
    add_tpos_enc_to_obj_ptrs = True
    has_no_obj_embed_spatial = True
    fill_hole_area = 8
    feat = 32  # image_size / 16: side of the stride-16 feature map the propagation tail runs on


class HieraBPlusConfig(ModelConfig):
    """sam2.1_hiera_base_plus at 1024^2 (BASELINE configs[4], the 3-D CT driver's model): the `Hiera` class defaults
    (hieradet.py:174-200) + upstream's embed_dim / num_heads; the propagation tail is the shared one on a 64 x 64 feature
    map (4096 queries, <= 7 * 4096 + 64 keys, 256^2 low-res masks).  Heads of 56 are zero-padded to 64 and the 112-channel
    stage to 128 when the weights are packed (PackedWeights._pack_hiera), so every kernel sees multiples of 32."""
    image_size = 1024
    embed_dim = 112
    num_heads = 2
    stages = (2, 3, 16, 3)
    global_att_blocks = (12, 16, 20)
    feat = 64


class EtamTiConfig(ModelConfig):
    """efficient_track_anything/configs/efficienttam_ti_512x512.yaml: plain ViT-tiny trunk + ViTDetNeck; the propagation
    tail is the shared one with three switches off (efficienttam_base.py)."""
    arch = "vit"
    patch = 16
    vit_dim = 192
    vit_depth = 12
    vit_heads = 3
    vit_window = 14
    vit_window_blocks = (0, 1, 3, 4, 6, 7, 9, 10)
    use_high_res_features_in_sam = False
    add_tpos_enc_to_obj_ptrs = False
    has_no_obj_embed_spatial = False


class EtamSConfig(EtamTiConfig):
    """efficient_track_anything/configs/efficienttam_s_512x512.yaml: ViT-small trunk (384-d, 6 heads of 64)."""
    vit_dim = 384
    vit_heads = 6


def hiera_plan(cfg=ModelConfig):
    """(dim_in, dim_out, heads, window, pool, emit) per block; window size lags one stage
    (hieradet.py:201-256)."""
    ends = [sum(cfg.stages[:i]) - 1 for i in range(1, len(cfg.stages) + 1)]
    pool_blocks = [e + 1 for e in ends[:-1]][: cfg.q_pool]
    plan, dim, heads, stage = [], cfg.embed_dim, cfg.num_heads, 1
    for i in range(sum(cfg.stages)):
        window = 0 if i in cfg.global_att_blocks else cfg.window_spec[stage - 1]
        dim_out = dim
        if i - 1 in ends:
            dim_out, heads, stage = dim * 2, heads * 2, stage + 1
        plan.append((dim, dim_out, heads, window, i in pool_blocks, i in ends))
        dim = dim_out
    return plan


def _cpad(c):
    """Channel counts of the Hiera residual stream as the kernels see them: multiples of 32."""
    return (c + 31) // 32 * 32


def _head_pad(hd):
    """Head widths the encoder attention kernels are built for: 64 and 96."""
    if hd > 96:
        raise NotImplementedError(f"encoder heads of {hd} channels")
    return hd if hd in (64, 96) else (64 if hd < 64 else 96)


# ------------------------------------------------------------------------------------------------
# constant tables (computed once on the host at pack time)
# ------------------------------------------------------------------------------------------------
def _sine_pos_2d(h, w, feats, temperature=10000.0):
    """PositionEmbeddingSine (position_encoding.py:79-112), token-major [h*w, feats]."""
    half = feats // 2
    y = torch.arange(1, h + 1, dtype=F32)
    x = torch.arange(1, w + 1, dtype=F32)
    y = y / (y[-1] + 1e-6) * (2 * math.pi)
    x = x / (x[-1] + 1e-6) * (2 * math.pi)
    dim_t = temperature ** (2 * torch.div(torch.arange(half, dtype=F32), 2, rounding_mode="floor") / half)
    px, py = x[:, None] / dim_t, y[:, None] / dim_t
    px = torch.stack((px[:, 0::2].sin(), px[:, 1::2].cos()), dim=2).flatten(1)
    py = torch.stack((py[:, 0::2].sin(), py[:, 1::2].cos()), dim=2).flatten(1)
    pos = torch.cat((py[:, None, :].expand(h, w, half), px[None, :, :].expand(h, w, half)), dim=2)
    return pos.reshape(h * w, feats).contiguous()


def _rope_tables(dim, end_x, end_y, theta=10000.0):
    """compute_axial_cis (position_encoding.py:174-183) as cos / sin [end_x*end_y, dim/2]."""
    freqs = 1.0 / (theta ** (torch.arange(0, dim, 4)[: dim // 4].float() / dim))
    t = torch.arange(end_x * end_y, dtype=F32)
    ang = torch.cat([torch.outer((t % end_x).float(), freqs),
                     torch.outer(torch.div(t, end_x, rounding_mode="floor").float(), freqs)], dim=-1)
    return torch.cos(ang).contiguous(), torch.sin(ang).contiguous()


class PackedWeights:
    """Reference state-dict tensors re-laid-out for the kernels (bf16 [N,K] GEMM weights, fused QKV,
    conv weights as [k][k][Cin][Cout], ConvTranspose as pixel-shuffle GEMMs, constant tables)."""

    def __init__(self, sd, device, cfg=ModelConfig):
        self.cfg = cfg
        self.device = device
        g = lambda k: sd[k].detach().to("cpu", F32)  # pack on the host, upload once
        dev = lambda t, dt=F32: t.to(dtype=dt).contiguous().to(device)
        w16 = lambda k: dev(g(k).reshape(g(k).shape[0], -1), BF16)
        f32 = lambda k: dev(g(k))
        self.plan = hiera_plan(cfg)

        # ---- image encoder ----
        if cfg.arch == "vit":
            self._pack_vit(g, dev, w16, f32)
        else:
            self._pack_hiera(g, dev, w16, f32)
        d = "sam_mask_decoder."
        fs = cfg.feat         # side of the stride-16 feature map (32 at 512^2, 64 at 1024^2)
        T = fs * fs
        self.feat_pos = dev(_sine_pos_2d(fs, fs, 256))  # vision_pos_enc of the stride-16 level, [T, 256]

        # ---- memory attention ----
        self.ma_layers = []
        for l in range(4):
            p = f"memory_attention.layers.{l}."
            sa, ca = p + "self_attn.", p + "cross_attn_image."
            self.ma_layers.append(dict(
                n1=(f32(p + "norm1.weight"), f32(p + "norm1.bias")),
                n2=(f32(p + "norm2.weight"), f32(p + "norm2.bias")),
                n3=(f32(p + "norm3.weight"), f32(p + "norm3.bias")),
                sa_qkv_w=dev(torch.cat([g(sa + "q_proj.weight"), g(sa + "k_proj.weight"), g(sa + "v_proj.weight")]), BF16),
                sa_qkv_b=dev(torch.cat([g(sa + "q_proj.bias"), g(sa + "k_proj.bias"), g(sa + "v_proj.bias")])),
                sa_o=(w16(sa + "out_proj.weight"), f32(sa + "out_proj.bias")),
                ca_q=(w16(ca + "q_proj.weight"), f32(ca + "q_proj.bias")),
                ca_k=(w16(ca + "k_proj.weight"), f32(ca + "k_proj.bias")),
                ca_v=(w16(ca + "v_proj.weight"), f32(ca + "v_proj.bias")),
                ca_o=(w16(ca + "out_proj.weight"), f32(ca + "out_proj.bias")),
                l1=(w16(p + "linear1.weight"), f32(p + "linear1.bias")),
                l2=(w16(p + "linear2.weight"), f32(p + "linear2.bias"))))
        self.ma_norm = (f32("memory_attention.norm.weight"), f32("memory_attention.norm.bias"))
        # tracked frames feed the decoder `norm(x) + no_mask_embed` (dense prompt = no_mask_embed broadcast,
        # prompt_encoder.py:176-180): the constant row is folded into the final norm's bias
        self.ma_norm_nomask = (self.ma_norm[0], dev(g("memory_attention.norm.bias")
                                                    + g("sam_prompt_encoder.no_mask_embed.weight").reshape(256)))
        # the cross-attention key / value inputs are the same for all 4 layers: project them with ONE GEMM each
        ca = "memory_attention.layers.{}.cross_attn_image."
        self.ca_k_all = (dev(torch.cat([g(ca.format(l) + "k_proj.weight") for l in range(4)]), BF16),
                         dev(torch.cat([g(ca.format(l) + "k_proj.bias") for l in range(4)])))
        self.ca_v_all = (dev(torch.cat([g(ca.format(l) + "v_proj.weight") for l in range(4)]), BF16),
                         dev(torch.cat([g(ca.format(l) + "v_proj.bias") for l in range(4)])))
        c, s = _rope_tables(256, fs, fs)
        # (tiled for the GEMM epilogues: 32 consecutive positions x 4 columns contiguous, see ops.tile_rope_table)
        self.rope_cos, self.rope_sin = ops.tile_rope_table(dev(c)), ops.tile_rope_table(dev(s))

        # ---- memory encoder ----
        e = "memory_encoder."
        ds = e + "mask_downsampler.encoder."
        kkio = lambda k: dev(g(k).permute(2, 3, 1, 0))  # [Cout,Cin,k,k] -> [k][k][Cin][Cout]
        self.md_convs = [(kkio(ds + f"{3 * i}.weight"), f32(ds + f"{3 * i}.bias"),
                          f32(ds + f"{3 * i + 1}.weight"), f32(ds + f"{3 * i + 1}.bias")) for i in range(3)]
        self.md_conv3_w = dev(g(ds + "9.weight").permute(0, 2, 3, 1).reshape(256, 576), BF16)  # k = (ky,kx,ci)
        self.md_conv3_b = f32(ds + "9.bias")
        self.md_ln3 = (f32(ds + "10.weight"), f32(ds + "10.bias"))
        self.md_out = (w16(ds + "12.weight"), f32(ds + "12.bias"))
        self.pix_proj = (w16(e + "pix_feat_proj.weight"), f32(e + "pix_feat_proj.bias"))
        self.fuser = []
        for l in range(2):
            p = e + f"fuser.layers.{l}."
            self.fuser.append(dict(dw_w=dev(g(p + "dwconv.weight").reshape(256, 49).t()), dw_b=f32(p + "dwconv.bias"),
                                   ln=(f32(p + "norm.weight"), f32(p + "norm.bias")),
                                   pw1=(w16(p + "pwconv1.weight"), f32(p + "pwconv1.bias")),
                                   pw2=(w16(p + "pwconv2.weight"), f32(p + "pwconv2.bias")),
                                   gamma=f32(p + "gamma")))
        self.mem_out = (w16(e + "out_proj.weight"), f32(e + "out_proj.bias"))
        self.mem_pos = dev(_sine_pos_2d(fs, fs, 64))  # [T, 64]
        self.maskmem_tpos = dev(g("maskmem_tpos_enc").reshape(cfg.num_maskmem, cfg.mem_dim))
        # (EfficientTAM has no such parameter: a zero vector makes the shared epilogue a no-op)
        self.no_obj_embed_spatial = (dev(g("no_obj_embed_spatial").reshape(cfg.mem_dim)) if cfg.has_no_obj_embed_spatial
                                     else dev(torch.zeros(cfg.mem_dim)))
        self.no_mem_embed = dev(g("no_mem_embed").reshape(1, 256))
        self.no_obj_ptr = dev(g("no_obj_ptr").reshape(256))
        # add_tpos_enc_to_obj_ptrs=false (EfficientTAM): pointer tokens get a zero position encoding and
        # obj_ptr_tpos_proj is an Identity without parameters (sam2_base.py:1393-1404)
        self.tpos_proj = ((f32("obj_ptr_tpos_proj.weight"), f32("obj_ptr_tpos_proj.bias"))
                          if cfg.add_tpos_enc_to_obj_ptrs else None)
        self.zero_ptr_pos = dev(torch.zeros(4 * ops._lib.MAX_PTRS, cfg.mem_dim))

        # ---- prompt encoder ----
        pe_ = "sam_prompt_encoder."
        self.pe_gauss = f32(pe_ + "pe_layer.positional_encoding_gaussian_matrix")
        self.point_table = dev(torch.cat([g(pe_ + f"point_embeddings.{j}.weight") for j in range(4)]
                                         + [g(pe_ + "not_a_point_embed.weight")]))
        self.no_mask_embed = dev(g(pe_ + "no_mask_embed.weight").reshape(1, 256))
        c01 = (torch.arange(fs, dtype=F32) + 0.5) / fs
        grid = torch.stack([c01[None, :].expand(fs, fs), c01[:, None].expand(fs, fs)], dim=-1).reshape(T, 2)
        ang = 2 * math.pi * ((2 * grid - 1) @ g(pe_ + "pe_layer.positional_encoding_gaussian_matrix"))
        self.dense_pe = dev(torch.cat([ang.sin(), ang.cos()], dim=-1))  # [T, 256]
        md = pe_ + "mask_downscaling."
        self.pm_convs = [(kkio(md + "0.weight"), f32(md + "0.bias"), f32(md + "1.weight"), f32(md + "1.bias")),
                         (kkio(md + "3.weight"), f32(md + "3.bias"), f32(md + "4.weight"), f32(md + "4.bias"))]
        self.pm_out = (dev(g(md + "6.weight").reshape(256, 16)), f32(md + "6.bias"))
        self.mask_downsample = (kkio("mask_downsample.weight"), f32("mask_downsample.bias"))

        # ---- mask decoder (fp32) ----
        tr = d + "transformer."
        lin = lambda p: (f32(p + "weight"), f32(p + "bias"))

        def attn(p):
            return dict(q=lin(p + "q_proj."), k=lin(p + "k_proj."), v=lin(p + "v_proj."), o=lin(p + "out_proj."))

        self.dec_layers = []
        for l in range(2):
            p = tr + f"layers.{l}."
            sa = attn(p + "self_attn.")
            sa["qkv_w"] = dev(torch.cat([g(p + "self_attn.q_proj.weight"), g(p + "self_attn.k_proj.weight"),
                                         g(p + "self_attn.v_proj.weight")]))
            sa["qkv_b"] = dev(torch.cat([g(p + "self_attn.q_proj.bias"), g(p + "self_attn.k_proj.bias"),
                                         g(p + "self_attn.v_proj.bias")]))
            self.dec_layers.append(dict(sa=sa, t2i=attn(p + "cross_attn_token_to_image."),
                                        i2t=attn(p + "cross_attn_image_to_token."),
                                        mlp=(lin(p + "mlp.layers.0."), lin(p + "mlp.layers.1.")),
                                        norms=[lin(p + f"norm{j}.") for j in (1, 2, 3, 4)]))
        for Lyr in self.dec_layers:  # image->token attention: k and v projections of the tokens as one [256, 256] matrix
            Lyr["i2t"]["kv_w"] = torch.cat([Lyr["i2t"]["k"][0], Lyr["i2t"]["v"][0]]).contiguous()
            Lyr["i2t"]["kv_b"] = torch.cat([Lyr["i2t"]["k"][1], Lyr["i2t"]["v"][1]]).contiguous()
        self.dec_final = attn(tr + "final_attn_token_to_image.")
        self.dec_final_norm = lin(tr + "norm_final_attn.")
        # image-side projections of one layer share their input (`keys`): fuse [t2i.k | t2i.v | i2t.q] into one
        # GEMM; the `+ key_pe` of k and q becomes a constant additive table  dense_pe @ W^T  (linearity)
        dpe = self.dense_pe.detach().cpu().double()

        def pe_table(key, zero_cols=0):
            return (dpe @ g(key).double().t()).float()

        for l, Lyr in enumerate(self.dec_layers):
            p = tr + f"layers.{l}."
            t2i, i2t = p + "cross_attn_token_to_image.", p + "cross_attn_image_to_token."
            Lyr["img_w"] = dev(torch.cat([g(t2i + "k_proj.weight"), g(t2i + "v_proj.weight"), g(i2t + "q_proj.weight")]))
            Lyr["img_b"] = dev(torch.cat([g(t2i + "k_proj.bias"), g(t2i + "v_proj.bias"), g(i2t + "q_proj.bias")]))
            Lyr["img_pe"] = dev(torch.cat([pe_table(t2i + "k_proj.weight"), torch.zeros(T, 128),
                                           pe_table(i2t + "q_proj.weight")], dim=1))
        fin = tr + "final_attn_token_to_image."
        self.dec_final["img_w"] = dev(torch.cat([g(fin + "k_proj.weight"), g(fin + "v_proj.weight")]))
        self.dec_final["img_b"] = dev(torch.cat([g(fin + "k_proj.bias"), g(fin + "v_proj.bias")]))
        self.dec_final["img_pe"] = dev(torch.cat([pe_table(fin + "k_proj.weight"), torch.zeros(T, 128)], dim=1))
        self.out_tokens = dev(torch.cat([g(d + "obj_score_token.weight"), g(d + "iou_token.weight"),
                                         g(d + "mask_tokens.weight")]))  # [6, 256]
        up = d + "output_upscaling."
        w = g(up + "0.weight")  # ConvTranspose2d [Cin=256, Cout=64, 2, 2] -> rows (dy, dx, co)
        self.up1_w = dev(w.permute(2, 3, 1, 0).reshape(256, 256))
        self.up1_b = dev(g(up + "0.bias").repeat(4))
        self.up1_ln = lin(up + "1.")
        w = g(up + "3.weight")  # [64, 32, 2, 2]
        self.up2_w = dev(w.permute(2, 3, 1, 0).reshape(128, 64))
        self.up2_b = dev(g(up + "3.bias").repeat(4))

        def mlp3(prefixes):
            return tuple(dev(torch.stack([g(p + f"layers.{j}.{kind}") for p in prefixes]))
                         for j in range(3) for kind in ("weight", "bias"))

        self.hyper = mlp3([d + f"output_hypernetworks_mlps.{j}." for j in range(4)])
        # the six token heads [object score, IoU, hyper 0..3] read token rows 0..5 in that order: one stacked MLP
        # (last layer zero-padded to 32 outputs)
        heads = [d + "pred_obj_score_head.", d + "iou_prediction_head."] + [d + f"output_hypernetworks_mlps.{j}."
                                                                            for j in range(4)]

        def pad32(t):
            out = torch.zeros((32,) + tuple(t.shape[1:]))
            out[: t.shape[0]] = t
            return out

        self.heads6 = [dev(torch.stack([g(p + "layers.0.weight") for p in heads])),
                       dev(torch.stack([g(p + "layers.0.bias") for p in heads])),
                       dev(torch.stack([g(p + "layers.1.weight") for p in heads])),
                       dev(torch.stack([g(p + "layers.1.bias") for p in heads])),
                       dev(torch.stack([pad32(g(p + "layers.2.weight")) for p in heads])),
                       dev(torch.stack([pad32(g(p + "layers.2.bias")) for p in heads]))]
        self.iou_head = mlp3([d + "iou_prediction_head."])
        self.score_head = mlp3([d + "pred_obj_score_head."])
        self.obj_ptr_proj = mlp3(["obj_ptr_proj."])


    def _pack_hiera(self, g, dev, w16, f32):
        """Hiera trunk + FpnNeck.  Channel counts that are not multiples of 32 (B+: 112) are zero-padded to the next one
        and heads whose width no attention kernel is built for (B+: 56) to 64 -- zero weight rows / columns, so the padded
        channels of the residual stream stay exactly zero and the padded head columns contribute nothing; LayerNorm runs
        over the real channels only (ops.layernorm(valid=...)).  The softmax scale of a padded head (kernels use
        1 / sqrt(padded width)) is corrected by scaling the query rows."""
        cfg = self.cfg
        t = "image_encoder.trunk."
        E, Ep, S4 = cfg.embed_dim, _cpad(cfg.embed_dim), cfg.image_size // 4
        pw = g(t + "patch_embed.proj.weight").reshape(E, 147)
        self.patch_w = dev(F.pad(pw, (0, 13, 0, Ep - E)), BF16)  # K 147 -> 160 (TMA row pitch must be 16 B aligned)
        self.patch_b = dev(F.pad(g(t + "patch_embed.proj.bias"), (0, Ep - E)))
        pe = F.interpolate(g(t + "pos_embed"), size=(S4, S4), mode="bicubic")
        win = g(t + "pos_embed_window")
        pe = pe + win.tile(1, 1, S4 // win.shape[2], S4 // win.shape[3])
        self.hiera_pos = dev(F.pad(pe[0].permute(1, 2, 0).reshape(S4 * S4, E), (0, Ep - E)))
        self.blocks = []
        for i, (din, dout, heads, ws, pool, emit) in enumerate(self.plan):
            p = t + f"blocks.{i}."
            hd = dout // heads
            hdp = _head_pad(hd)
            dinp, doutp, da = _cpad(din), _cpad(dout), heads * hdp
            qw = F.pad(g(p + "attn.qkv.weight").reshape(3, heads, hd, din), (0, dinp - din, 0, hdp - hd))
            qb = F.pad(g(p + "attn.qkv.bias").reshape(3, heads, hd), (0, hdp - hd))
            if hdp != hd:
                qw[0] *= math.sqrt(hdp / hd)
                qb[0] *= math.sqrt(hdp / hd)
            pj = F.pad(g(p + "attn.proj.weight").reshape(dout, heads, hd), (0, hdp - hd, 0, 0, 0, doutp - dout))
            padv = lambda k, n, n_pad: dev(F.pad(g(k), (0, n_pad - n)))
            blk = dict(n1=(f32(p + "norm1.weight"), f32(p + "norm1.bias")),
                       n2=(f32(p + "norm2.weight"), f32(p + "norm2.bias")),
                       qkv_w=dev(qw.reshape(3 * da, dinp), BF16), qkv_b=dev(qb.reshape(3 * da)),
                       proj_w=dev(pj.reshape(doutp, da), BF16), proj_b=padv(p + "attn.proj.bias", dout, doutp),
                       w1=dev(F.pad(g(p + "mlp.layers.0.weight"), (0, doutp - dout)), BF16), b1=f32(p + "mlp.layers.0.bias"),
                       w2=dev(F.pad(g(p + "mlp.layers.1.weight"), (0, 0, 0, doutp - dout)), BF16),
                       b2=padv(p + "mlp.layers.1.bias", dout, doutp),
                       dims=(din, dinp, dout, doutp, da, hdp))
            if din != dout:
                blk["sc_w"] = dev(F.pad(g(p + "proj.weight"), (0, dinp - din, 0, doutp - dout)), BF16)
                blk["sc_b"] = padv(p + "proj.bias", dout, doutp)
            self.blocks.append(blk)
        n = "image_encoder.neck.convs."
        self.neck = []
        for j in range(4):
            wj = g(n + f"{j}.conv.weight")
            wj = wj.reshape(wj.shape[0], -1)
            self.neck.append((dev(F.pad(wj, (0, _cpad(wj.shape[1]) - wj.shape[1])), BF16), f32(n + f"{j}.conv.bias")))
        d = "sam_mask_decoder."
        self.conv_s0 = (w16(d + "conv_s0.weight"), f32(d + "conv_s0.bias"))
        self.conv_s1 = (w16(d + "conv_s1.weight"), f32(d + "conv_s1.bias"))

    def _pack_vit(self, g, dev, w16, f32):
        """ViT trunk + ViTDetNeck (efficient_track_anything/modeling/backbones/vitdet.py, image_encoder.py:47-108)."""
        cfg = self.cfg
        t = "image_encoder.trunk."
        self.patch_w = w16(t + "patch_embed.proj.weight")  # [192, 3*16*16], k = (c, ky, kx)
        self.patch_b = f32(t + "patch_embed.proj.bias")
        pe = g(t + "pos_embed")[:, 1:]  # drop the cls token (get_abs_pos, backbones/utils.py:97-128)
        size = int(round(pe.shape[1] ** 0.5))
        n = cfg.image_size // cfg.patch
        pe = F.interpolate(pe.reshape(1, size, size, -1).permute(0, 3, 1, 2), size=(n, n), mode="bicubic",
                           align_corners=False)
        self.vit_pos = dev(pe[0].permute(1, 2, 0).reshape(n * n, cfg.vit_dim))
        self.blocks = []
        for i in range(cfg.vit_depth):
            p = t + f"blocks.{i}."
            self.blocks.append(dict(n1=(f32(p + "norm1.weight"), f32(p + "norm1.bias")),
                                    n2=(f32(p + "norm2.weight"), f32(p + "norm2.bias")),
                                    qkv_w=w16(p + "attn.qkv.weight"), qkv_b=f32(p + "attn.qkv.bias"),
                                    proj_w=w16(p + "attn.proj.weight"), proj_b=f32(p + "attn.proj.bias"),
                                    w1=w16(p + "mlp.layers.0.weight"), b1=f32(p + "mlp.layers.0.bias"),
                                    w2=w16(p + "mlp.layers.1.weight"), b2=f32(p + "mlp.layers.1.bias")))
        nk = "image_encoder.neck.convs.0."
        self.neck_1x1 = w16(nk + "conv_1x1.weight")
        self.neck_ln0 = (f32(nk + "norm_0.weight"), f32(nk + "norm_0.bias"))
        self.neck_3x3 = dev(g(nk + "conv_3x3.weight").permute(0, 2, 3, 1).reshape(256, 9 * 256), BF16)  # k = (ky, kx, ci)
        self.neck_ln1 = (f32(nk + "norm_1.weight"), f32(nk + "norm_1.bias"))


class Engine:
    def __init__(self, weights: PackedWeights):
        self.w = weights
        self.cfg = weights.cfg
        self.sm_budget = 148  # SMs the tracked frame may count on (fewer while the encoder owns an SM partition)
        self._tail_stream = None
        self._side_streams = {}
        self._tok_const = {}
        # independent sub-chains of a frame (memory-bank projections || first self-attention block; token side || image
        # side of the two-way decoder; pix_feat projection || mask down-sampler) run on forked streams, i.e. as parallel
        # branches of the captured frame graph -- the frame is a latency chain of ~125 small kernels, not throughput bound
        self.fork_branches = os.environ.get("USVM2_FORK", "1") != "0"
        self.fused_windows = os.environ.get("USVM2_FUSED_WINDOWS", "1") != "0"
        # global / 14 x 14 windowed encoder attention on the tcgen05 kernel (0: the mma.sync kernels, kept for comparison)
        self.tc5_encoder_attn = os.environ.get("USVM2_ENCODER_ATTN_TC5", "1") != "0"
        # memory-attention feed-forward block as one cluster kernel at one or two objects (0: two GEMM launches)
        self.fused_ffn = os.environ.get("USVM2_FUSED_FFN", "1") != "0"
        # tracked frames: the mask's 512^2 upsampling + sigmoid inside the first down-sampler convolution (0: separate pass)
        self.lazy_mask_upsample = os.environ.get("USVM2_LAZY_MASK_UPSAMPLE", "1") != "0"

    # ---------------------------------------------------------------- forked branches
    def _side(self, i):
        st = self._side_streams.get(i)
        if st is None:
            st = self._side_streams[i] = torch.cuda.Stream()
        return st

    @staticmethod
    def _handoff(dst, src, *tensors):
        """Make stream `dst` wait for everything enqueued on `src`; the listed tensors (allocated on one of the two) are
        from now on used on `dst` as well, which the caching allocator must know before it recycles their blocks."""
        dst.wait_stream(src)
        for t in tensors:
            if t is not None:
                t.record_stream(dst)

    # ---------------------------------------------------------------- image encoder
    def encode_frames(self, imgs):
        """imgs fp32 [F,3,512,512] -> dict(feat [F,1024,256] fp32 (+ bf16 copy), feat_s1 [F,4096,64],
        feat_s0 [F,16384,32]) -- Hiera.forward + FpnNeck.forward + forward_image
        (hieradet.py:283-299, image_encoder.py:104-136, sam2_base.py:1220-1232)."""
        w = self.w
        if self.cfg.arch == "vit":
            return self._encode_frames_vit(imgs)
        Fr = imgs.shape[0]
        S4, fs = self.cfg.image_size // 4, self.cfg.feat
        A = ops.im2col_patch(imgs.contiguous())
        x, _ = ops.gemm_bf16(A, w.patch_w, bias=w.patch_b, residual=w.hiera_pos, res_mod=S4 * S4, f32=True)
        H = W = S4
        stage_out = []
        for blk, (_, _, heads, ws, pool, emit) in zip(w.blocks, w.plan):
            # (din / dout: real channel counts = LayerNorm widths; dinp / doutp: padded stream widths; da, hd: width of the
            # attention tensors and of one head after head padding -- all equal to the real ones for Hiera-tiny)
            din, dinp, dout, doutp, da, hd = blk["dims"]
            _, h = ops.layernorm(x, *blk["n1"], 1e-6, bf16=True, valid=din)
            if din != dout:
                sc, _ = ops.gemm_bf16(h, blk["sc_w"], bias=blk["sc_b"], f32=True)
                shortcut = ops.maxpool2(sc, Fr, H, W, doutp)
            else:
                shortcut = x
            _, qkv = ops.gemm_bf16(h, blk["qkv_w"], bias=blk["qkv_b"], bf16=True)
            Ho, Wo = (H // 2, W // 2) if pool else (H, W)
            # windows of <= 64 keys: one fused kernel.  (It also handles the 14 x 14 windows of stage 3, but one CTA per
            # (window, head) walking 13 query slabs is slower there than gather + flash kernel + scatter: measured.)
            if self.tc5_encoder_attn and (ws in (7, 14) or (ws == 0 and (H * W) % 128 == 0)) and (not pool or ws == 14):
                # stage-3 / 4 blocks (global, 14 x 14 and 7 x 7 windows, the q-pool block between them): tcgen05 kernel,
                # window (un)partition = TMA coordinates
                att = ops.hiera_attn(qkv, blk["qkv_b"], Fr, H, W, da, heads, window=ws, pool=pool)
            elif 0 < ws * ws <= 64 and self.fused_windows:
                # stage-1 / 2 windows (64 and 16 keys): a 128-row tcgen05 tile would be >= 50 % masking; fused mma.sync kernel
                att = ops.window_attn(qkv, blk["qkv_b"], Fr, H, W, ws, pool, da, heads)
            elif ws > 0:
                Qw, Kw, Vw, nw, nq, nk = ops.window_gather(qkv, blk["qkv_b"], Fr, H, W, ws, pool, da)
                Ow = ops.fmha(Qw, Kw, Vw, Fr * nw, heads, nq, nk, hd, (0, nq * da, da, hd),
                              (0, nk * da, da, hd), (0, nk * da, da, hd))
                att = ops.window_scatter(Ow, Fr, Ho, Wo, ws // 2 if pool else ws, da)
            else:
                T = H * W
                att = ops.fmha(qkv, qkv, qkv, Fr, heads, T, T, hd, (0, T * 3 * da, 3 * da, hd),
                               (da, T * 3 * da, 3 * da, hd), (2 * da, T * 3 * da, 3 * da, hd))
                att = att.reshape(Fr * T, da)
            H, W = Ho, Wo
            x, _ = ops.gemm_bf16(att, blk["proj_w"], bias=blk["proj_b"], residual=shortcut, f32=True)
            _, h2 = ops.layernorm(x, *blk["n2"], 1e-6, bf16=True, valid=dout)
            _, m = ops.gemm_bf16(h2, blk["w1"], bias=blk["b1"], act=ACT_GELU, bf16=True)
            x, xb = ops.gemm_bf16(m, blk["w2"], bias=blk["b2"], residual=x, f32=True, bf16=emit)
            if emit:
                stage_out.append(xb)
        s0, s1, s2, s3 = stage_out  # tiny: 128^2 x 96, 64^2 x 192, 32^2 x 384, 16^2 x 768
        lat3, _ = ops.gemm_bf16(s3, *self._wb(w.neck[0]), f32=True)
        lat2, _ = ops.gemm_bf16(s2, *self._wb(w.neck[1]), f32=True)
        lat2_b = ops.upsample2_add_(lat2, lat3, Fr, fs, fs, 256, bf16=True)
        _, lat1_b = ops.gemm_bf16(s1, *self._wb(w.neck[2]), bf16=True)
        _, lat0_b = ops.gemm_bf16(s0, *self._wb(w.neck[3]), bf16=True)
        feat_s1, _ = ops.gemm_bf16(lat1_b, *self._wb(w.conv_s1), f32=True)
        feat_s0, _ = ops.gemm_bf16(lat0_b, *self._wb(w.conv_s0), f32=True)
        T = fs * fs
        return dict(feat=lat2.view(Fr, T, 256), feat_bf16=lat2_b.view(Fr, T, 256),
                    feat_s1=feat_s1.view(Fr, 4 * T, 64), feat_s0=feat_s0.view(Fr, 16 * T, 32))

    @staticmethod
    def _wb(pair):
        return pair[0], pair[1]

    def _encode_frames_vit(self, imgs):
        """EfficientTAM image encoder: ViT.forward + ViTDetNeck.forward (efficient_track_anything/modeling/backbones/
        vitdet.py:282-299, image_encoder.py:93-108).  imgs fp32 [F,3,512,512] -> same dict as the Hiera path; the two
        high-resolution levels do not exist in this variant and are constant zero maps (the shared decoder kernels add
        them), the patch embedding is a GEMM over non-overlapping 16 x 16 patches."""
        w, cfg = self.w, self.cfg
        Fr, P = imgs.shape[0], cfg.patch
        n = cfg.image_size // P
        T, C, heads = n * n, cfg.vit_dim, cfg.vit_heads
        hd = C // heads
        A = ops.im2col_patch_grid(imgs.contiguous(), P)  # [F*T, 3*P*P] bf16, k = (c, ky, kx)
        x, _ = ops.gemm_bf16(A, w.patch_w, bias=w.patch_b, residual=w.vit_pos, res_mod=T, f32=True)
        for i, blk in enumerate(w.blocks):
            _, h = ops.layernorm(x, *blk["n1"], 1e-6, bf16=True)
            _, qkv = ops.gemm_bf16(h, blk["qkv_w"], bias=blk["qkv_b"], bf16=True)
            if self.tc5_encoder_attn and hd in (64, 96) and T % 128 == 0 and (
                    i not in cfg.vit_window_blocks or cfg.vit_window == 14):
                att = ops.hiera_attn(qkv, blk["qkv_b"], Fr, n, n, C, heads,
                                     window=cfg.vit_window if i in cfg.vit_window_blocks else 0)
            elif i in cfg.vit_window_blocks:
                ws = cfg.vit_window
                Qw, Kw, Vw, nw, nq, nk = ops.window_gather(qkv, blk["qkv_b"], Fr, n, n, ws, False, C)
                Ow = ops.fmha(Qw, Kw, Vw, Fr * nw, heads, nq, nk, hd, (0, nq * C, C, hd), (0, nk * C, C, hd),
                              (0, nk * C, C, hd))
                att = ops.window_scatter(Ow, Fr, n, n, ws, C)
            else:
                att = ops.fmha(qkv, qkv, qkv, Fr, heads, T, T, hd, (0, T * 3 * C, 3 * C, hd),
                               (C, T * 3 * C, 3 * C, hd), (2 * C, T * 3 * C, 3 * C, hd)).reshape(Fr * T, C)
            x, _ = ops.gemm_bf16(att, blk["proj_w"], bias=blk["proj_b"], residual=x, f32=True)
            _, h2 = ops.layernorm(x, *blk["n2"], 1e-6, bf16=True)
            _, m = ops.gemm_bf16(h2, blk["w1"], bias=blk["b1"], act=ACT_GELU, bf16=True)
            last = i == len(w.blocks) - 1
            x, xb = ops.gemm_bf16(m, blk["w2"], bias=blk["b2"], residual=x, f32=True, bf16=last)
        y, _ = ops.gemm_bf16(xb, w.neck_1x1, f32=True)                       # 1x1 conv, no bias
        y, _ = ops.layernorm(y, *w.neck_ln0, 1e-6, f32=True)                 # LayerNorm2d = per-pixel norm over channels
        A3 = ops.im2col_nhwc(y, Fr, n, n, 256, 3, 1, 1)
        z, _ = ops.gemm_bf16(A3, w.neck_3x3, f32=True)                       # 3x3 conv, no bias
        feat, feat_b = ops.layernorm(z, *w.neck_ln1, 1e-6, f32=True, bf16=True)
        zeros = lambda rows, ch: torch.zeros((Fr, rows, ch), dtype=F32, device=imgs.device)
        return dict(feat=feat.view(Fr, T, 256), feat_bf16=feat_b.view(Fr, T, 256), feat_s1=zeros(4 * T, 64),
                    feat_s0=zeros(16 * T, 32))

    # ---------------------------------------------------------------- memory attention
    def attention_prefix(self, feat, B, group=0, out=None):
        """The part of the memory attention that does not depend on the memory bank: x = feat + 0.1 pos, and of layer 0 the
        self-attention block (norm1, q / k / v projection + RoPE, attention, output projection + residual), norm2 and the
        cross-attention's query projection + RoPE.  Returns (x fp32 [B*T,256], q bf16 [B*T,256]) -- with `out` = (x, q)
        buffers the two are written there (the frame graph of frame t computes this for frame t + 1, see
        SAM2VideoPredictor._run_tracked_frame)."""
        w = self.w
        T = self.cfg.feat ** 2
        cs, sn = w.rope_cos, w.rope_sin
        L = w.ma_layers[0]
        x, _ = ops.axpby(feat, w.feat_pos, 1.0, 0.1, rows=B * T, x_mod=T, y_mod=T, x_div=group * T)
        _, h = ops.layernorm(x, *L["n1"], 1e-5, bf16=True)
        _, qkv = ops.gemm_bf16(h, L["sa_qkv_w"], bias=L["sa_qkv_b"], bf16=True, rope=(cs, sn, 512, T, T))
        sa_impl, sa_splits = ("mma", max(1, 8 // B)) if B <= 2 and T <= 1024 else (None, self._splits(B, T))
        o = ops.fmha(qkv, qkv, qkv, B, 1, T, T, 256, (0, T * 768, 768, 256), (256, T * 768, 768, 256),
                     (512, T * 768, 768, 256), num_splits=sa_splits, impl=sa_impl)
        x, h = ops.gemm_bf16(o.view(B * T, 256), L["sa_o"][0], bias=L["sa_o"][1], residual=x, f32=True,
                             ln=(L["n2"][0], L["n2"][1], 1e-5), out_f32=None if out is None else out[0])
        _, q = ops.gemm_bf16(h, L["ca_q"][0], bias=L["ca_q"][1], bf16=True, rope=(cs, sn, 256, T, T),
                             out_bf16=None if out is None else out[1])
        return x, q

    def memory_attention(self, feat, k_in, v_in, Nk, n_ptr_tok, B, bank=None, fold_no_mask=False, group=0, prefix=None):
        """feat fp32 [1024,256] (one frame, shared by the B objects; with group > 0: [B/group, 1024, 256], one frame per
        `group` consecutive objects -- several videos batched into one launch); k_in / v_in bf16 [B, Nk, 64] assembled memory
        (k_in already carries the position encodings), the last n_ptr_tok rows are object-pointer tokens (no RoPE).
        Returns fp32 [B*1024, 256]  (MemoryAttention.forward, memory_attention.py:119-169; RoPEAttention,
        sam/transformer.py:311-360).  RoPE is fused into the q / k projection epilogues; the key / value
        projections of the 4 layers are batched into one GEMM each (`project_memory`).
        `bank`: callable returning (k_all, v_all) -- lets the caller produce the projected bank on a forked stream; it is
        called right before the first cross-attention, after the first self-attention block has been enqueued.
        fold_no_mask: return norm(x) + no_mask_embed, the decoder's `src` on tracked frames (sam_heads(src_ready=True)).
        prefix: (x, q) from attention_prefix() for this frame, computed earlier -- layer 0 then starts at its cross-attention."""
        w = self.w
        T = self.cfg.feat ** 2
        cs, sn = w.rope_cos, w.rope_sin
        if prefix is None:
            x, _ = ops.axpby(feat, w.feat_pos, 1.0, 0.1, rows=B * T, x_mod=T, y_mod=T, x_div=group * T)
        k_all = v_all = None
        if bank is None:
            k_all, v_all = self.project_memory(k_in, v_in, Nk, n_ptr_tok, B)
        for li, L in enumerate(w.ma_layers):
            if li == 0 and prefix is not None:
                x, q = prefix
            else:
                _, h = ops.layernorm(x, *L["n1"], 1e-5, bf16=True)
                _, qkv = ops.gemm_bf16(h, L["sa_qkv_w"], bias=L["sa_qkv_b"], bf16=True, rope=(cs, sn, 512, T, T))
                # 1024 keys are 16 key tiles: with one or two objects the mma.sync kernel (64-row query tiles, so twice the
                # CTAs per split) beats the tcgen05 kernel, whose fixed prologue / epilogue dominates such short ranges
                sa_impl, sa_splits = ("mma", max(1, 8 // B)) if B <= 2 and T <= 1024 else (None, self._splits(B, T))
                o = ops.fmha(qkv, qkv, qkv, B, 1, T, T, 256, (0, T * 768, 768, 256), (256, T * 768, 768, 256),
                             (512, T * 768, 768, 256), num_splits=sa_splits, impl=sa_impl)
                # out-projection + residual, then the next sub-block's LayerNorm
                x, h = ops.gemm_bf16(o.view(B * T, 256), L["sa_o"][0], bias=L["sa_o"][1], residual=x, f32=True,
                                     ln=(L["n2"][0], L["n2"][1], 1e-5))
                _, q = ops.gemm_bf16(h, L["ca_q"][0], bias=L["ca_q"][1], bf16=True, rope=(cs, sn, 256, T, T))
            if k_all is None:
                k_all, v_all = bank()
            o = ops.fmha(q, k_all, v_all, B, 1, T, Nk, 256, (0, T * 256, 256, 256), (li * 256, Nk * 1024, 1024, 256),
                         (li * 256, Nk * 1024, 1024, 256), num_splits=self._splits(B, Nk))
            x, h = ops.gemm_bf16(o.view(B * T, 256), L["ca_o"][0], bias=L["ca_o"][1], residual=x, f32=True,
                                 ln=(L["n3"][0], L["n3"][1], 1e-5))
            if self.fused_ffn and B * T <= 2048:  # (one wave of clusters: 8 CTAs per 128 rows; measured slower from 4 objects on)
                # linear1 + ReLU + linear2 + residual as one cluster kernel: the hidden activations stay on the SM
                x = ops.ffn_fused(h, x, L["l1"][0], L["l1"][1], L["l2"][0], L["l2"][1])
            else:
                _, m = ops.gemm_bf16(h, L["l1"][0], bias=L["l1"][1], act=ACT_RELU, bf16=True)
                x, _ = ops.gemm_bf16(m, L["l2"][0], bias=L["l2"][1], residual=x, f32=True)
        out, _ = ops.layernorm(x, *(w.ma_norm_nomask if fold_no_mask else w.ma_norm), 1e-5, f32=True)
        return out

    def project_memory(self, k_in, v_in, Nk, n_ptr_tok, B):
        """Key / value projections of the assembled bank for all 4 layers at once: ([B*Nk, 4*256] rotated keys, values)."""
        return self.project_keys(k_in, Nk, n_ptr_tok, B), self.project_values(v_in, Nk, B)

    def project_keys(self, k_in, Nk, n_ptr_tok, B):
        w = self.w
        return ops.gemm_bf16(k_in.view(B * Nk, 64), w.ca_k_all[0], bias=w.ca_k_all[1], bf16=True,
                             rope=(w.rope_cos, w.rope_sin, 1024, Nk, Nk - n_ptr_tok))[1]

    def project_values(self, v_in, Nk, B):
        w = self.w
        return ops.gemm_bf16(v_in.view(B * Nk, 64), w.ca_v_all[0], bias=w.ca_v_all[1], bf16=True)[1]

    def _splits(self, B, Nk):
        """Split-KV factor: the tcgen05 kernel runs T / 128 query tiles per object (8 at 512^2); fill the 148 SMs."""
        tiles = (Nk + 63) // 64
        want = max(1, self.sm_budget // (self.cfg.feat ** 2 // 128 * B))
        return max(1, min(want, tiles))

    def assemble_memory(self, ctrl, B, n_mem, n_ptr):
        """Memory-bank assembly from the frame store named by the device control block
        (sam2_base.py:1344-1437): returns (k_in, v_in, Nk, n_ptr_tokens)."""
        w = self.w
        if n_ptr > 0:
            ptr_pos = ops.ptr_tpos(ctrl, *w.tpos_proj, n_ptr) if w.tpos_proj is not None else w.zero_ptr_pos
        else:
            ptr_pos = None
        k_in, v_in, Nk = ops.build_memory_store(ctrl, w.mem_pos, w.maskmem_tpos, ptr_pos, B, n_mem, n_ptr,
                                                T=self.cfg.feat ** 2)
        return k_in, v_in, Nk, 4 * n_ptr

    def track_frame(self, f, ctrl, B, n_mem, n_ptr, video_hw, fill_hole_area, group=0, prefix=None, next_feat=None,
                    next_prefix_out=None):
        """One steady-state tracked frame, entirely on the device and free of host-dependent control flow (so it can
        be captured in a CUDA graph): memory attention over the bank named by `ctrl`, SAM heads with multimask
        output, memory encoder, hole filling, video-resolution resize.  Everything the frame leaves behind is
        written into slot ctrl->cur_frame of the frame store.  f: dict(feat, feat_bf16, feat_s0, feat_s1) of this
        frame.  group > 0: the batch holds B / group videos of `group` objects each in lock-step (same frame index, same
        bank layout); f then carries one frame per video ([B/group, ...]) and every per-video operand is indexed by
        object // group.  Returns (video_res logits [B,1,H,W], hole-filled low-res logits [B,1,128,128]).
        Frame pipelining: `prefix` = this frame's attention_prefix() computed during the previous frame; `next_feat`
        (+ `next_prefix_out` buffers) = the next frame's features, whose prefix is computed here on a forked branch beside
        the decoder and the memory encoder -- neither depends on the other, and the frame is a latency chain."""
        main = torch.cuda.current_stream()
        if self.fork_branches:
            # the bank (gather + temporal encodings + K/V projections of ~7 k rows) does not depend on this frame's
            # features, the first self-attention block does not depend on the bank: two branches, joined before the
            # first cross-attention
            side, side_v = self._side(0), self._side(2)
            side.wait_stream(main)
            side_v.wait_stream(main)
            with torch.cuda.stream(side_v):
                # pix_feat_proj of the memory encoder only reads this frame's features: off the critical path here
                pp, _ = ops.gemm_bf16(f["feat_bf16"].view(-1, 256), *self.w.pix_proj, f32=True)
            with torch.cuda.stream(side):
                k_in, v_in, Nk, n_tok = self.assemble_memory(ctrl, B, n_mem, n_ptr)
            self._handoff(side_v, side, v_in)
            with torch.cuda.stream(side):
                k_all = self.project_keys(k_in, Nk, n_tok, B)
            with torch.cuda.stream(side_v):  # keys and values are projected side by side
                v_all = self.project_values(v_in, Nk, B)
            kv = (k_all, v_all)

            def bank():
                self._handoff(main, side, k_all)
                self._handoff(main, side_v, v_all, pp)
                return kv

            pix = self.memory_attention(f["feat"], None, None, Nk, n_tok, B, bank=bank, fold_no_mask=True, group=group,
                                        prefix=prefix)
        else:
            pp = None
            k_in, v_in, Nk, n_tok = self.assemble_memory(ctrl, B, n_mem, n_ptr)
            pix = self.memory_attention(f["feat"], k_in, v_in, Nk, n_tok, B, fold_no_mask=True, group=group, prefix=prefix)
        pre_s = None
        if next_feat is not None:
            # the next frame's bank-independent prefix: a parallel branch from here (the cross-attentions above want the
            # whole device; the decoder and the memory encoder below are chains of small kernels)
            pre_s = self._side(3)
            pre_s.wait_stream(main)
            with torch.cuda.stream(pre_s):
                self.attention_prefix(next_feat, B, group=group, out=next_prefix_out)
        o = self.sam_heads(pix, f["feat_s0"], f["feat_s1"], B, None, multimask=True, src_ready=True, defer_ptr=True,
                           group=group)
        # The user-facing tail (single-CTA hole filling, store write, video-resolution resize) is independent of the
        # memory encoder: it runs on a forked stream -- a parallel branch of the captured graph -- and joins at the end.
        if self._tail_stream is None:
            self._tail_stream = torch.cuda.Stream()
        tail = self._tail_stream
        self._handoff(tail, main, o["low"], o["score"], *o.get("_keep", ()))
        with torch.cuda.stream(tail):
            if o["obj_ptr"] is None:
                o["obj_ptr"] = o["finish_ptr"]()
            pm = ops.fill_holes(o["low"], fill_hole_area) if fill_hole_area > 0 else o["low"]
            ops.store_outputs(ctrl, o["obj_ptr"], o["score"], pm)
            vh, vw = video_hw
            video = pm if (vh, vw) == tuple(pm.shape[-2:]) else ops.resize_bilinear(pm, vh, vw)
        # (non_overlap_masks_for_mem_enc applies on every frame in eval, sam2_base.py:1466-1471: device-side, so the frame
        # stays one capturable graph)
        mask_in = self.mem_mask_input(o["low"], False, non_overlap=self.cfg.non_overlap_masks_for_mem_enc, group=group,
                                      lazy=True)
        self.encode_memory(f["feat_bf16"], mask_in, o["score"], B, ctrl=ctrl, pix_proj=pp, group=group)
        self._handoff(main, tail, video, pm)
        if pre_s is not None:
            main.wait_stream(pre_s)
        return video, pm

    # ---------------------------------------------------------------- SAM heads
    def token_constants(self, B):
        """Tracked frames carry no prompt: their 8 decoder tokens (6 output tokens + 2 padding points) are constants of the
        model, and so is everything the token side computes before it first meets the image -- layer 0's self-attention,
        norm1 and the token->image query projection (transformer.py:181-199).  Computed once per object count with the
        same kernels the general path uses; returns (tokens, queries after norm1, layer-0 token->image query)."""
        hit = self._tok_const.get(B)
        if hit is not None:
            return hit
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("token constants must be computed before the frame graph is captured")
        w = self.w
        sk = ops.gemm_skinny
        Nt = 8
        tokens = torch.cat([w.out_tokens[None].expand(B, -1, -1), self.no_point_tokens(B)], dim=1)
        tokens = tokens.reshape(B * Nt, 256).contiguous()
        L0 = w.dec_layers[0]
        qkv = sk(tokens, L0["sa"]["qkv_w"], L0["sa"]["qkv_b"])
        o = ops.attn_small(qkv[:, 0:256], qkv[:, 256:512], qkv[:, 512:768], B, 8, Nt, Nt, 32)
        queries = ops.layernorm(sk(o, *L0["sa"]["o"]), L0["norms"][0][0], L0["norms"][0][1], 1e-5, f32=True)[0]
        q = sk(queries, *L0["t2i"]["q"], x2=tokens)
        torch.cuda.current_stream().synchronize()  # one-off: later uses come from other streams / captured graphs
        self._tok_const[B] = (tokens, queries, q)
        return self._tok_const[B]

    def sam_heads(self, pix_feat, feat_s0, feat_s1, B, sparse, dense=None, multimask=True, feat_shared=True,
                  src_ready=False, defer_ptr=False, group=0):
        """pix_feat fp32 [B*1024,256]; sparse fp32 [B,P,256] prompt tokens; dense fp32 [B*1024,256] or None
        (-> no_mask_embed).  Returns dict(low [B,1,128,128], obj_ptr [B,256], score [B,1], iou [B,1]).
        (_forward_sam_heads sam2_base.py:1010-1166 + MaskDecoder mask_decoder.py:110-295 +
        TwoWayTransformer transformer.py:90-212).  Token-side linears (8 rows per object) run on the skinny
        GEMM, image-side projections of a layer are one fused GEMM, attention on the t2i / i2t kernels.
        feat_s0 / feat_s1: one frame shared by all B objects (feat_shared), one per `group` consecutive objects
        (group > 0), or one per object."""
        w = self.w
        fs = self.cfg.feat
        T = fs * fs
        feat_group = group if group > 0 else (B if feat_shared else 0)
        if src_ready:  # pix_feat already carries the dense prompt embedding (memory_attention(fold_no_mask=True))
            assert dense is None
            src = pix_feat
        elif dense is None:
            src, _ = ops.axpby(pix_feat, w.no_mask_embed, rows=B * T, y_mod=1)
        else:
            src, _ = ops.axpby(pix_feat, dense, rows=B * T)
        const0 = None
        if sparse is None:  # no prompt (tracked frame): constant tokens, token-side prefix precomputed
            Nt = 8
            tokens, q_norm1, q_t2i = self.token_constants(B)
            const0 = (q_norm1, q_t2i)
        else:
            P = sparse.shape[1]
            Nt = 6 + P
            tokens = torch.cat([w.out_tokens[None].expand(B, -1, -1), sparse], dim=1).reshape(B * Nt, 256).contiguous()
        queries, keys = tokens, src
        ln = lambda x, nb: ops.layernorm(x, nb[0], nb[1], 1e-5, f32=True)[0]
        sk = ops.gemm_skinny
        tok = torch.cuda.current_stream()
        img_s = self._side(1) if self.fork_branches else None
        hs_pre, keys = self._two_way_transformer(tokens, keys, B, Nt, img_stream=img_s, const0=const0,
                                                 defer_final_norm=True)
        # output upscaling (image side) || the six token heads (token side); they meet in the mask product
        with torch.cuda.stream(img_s if img_s is not None else tok):
            g1 = ops.gemm_f32(keys, w.up1_w, w.up1_b, tf32=_DEC_TF32)
            u1 = ops.upscale1_ln_gelu(g1, feat_s1, w.up1_ln[0], w.up1_ln[1], B, fs, fs, feat_group)
            g2 = ops.gemm_f32(u1, w.up2_w, w.up2_b, tf32=_DEC_TF32)
        # six stacked heads on token rows 0..5: [object score, IoU, hyper-network 0..3]
        # norm_final_attn runs on load in the first head layer, which also leaves the normalised token rows 0..5 in hs
        W1, b1, W2, b2, W3, b3 = w.heads6
        hs = torch.empty_like(hs_pre)
        h1 = sk(None, W1, b1, M=B, x_ptr=hs_pre.data_ptr(), x_rs=Nt * 256, x_is=256, act=ACT_RELU, instances=6,
                ln=(w.dec_final_norm[0], w.dec_final_norm[1], 1e-5), ln_out=hs, ln_rs=Nt * 256, ln_is=256)
        h2 = sk(h1, W2, b2, M=B, x_rs=6 * 256, x_is=256, act=ACT_RELU, instances=6)
        y = sk(h2, W3, b3, M=B, x_rs=6 * 256, x_is=256, instances=6)  # [B, 6*32]
        if img_s is not None:
            self._handoff(tok, img_s, g2, keys)
        masks = ops.upscale2_masks(g2, feat_s0, y[:, 64:], B, 2 * fs, 2 * fs, feat_group, hyper_bs=192)
        # single-mask output without the stability fallback (apply_postprocessing=False, mask_decoder.py:160-166): a
        # threshold no stability score can miss keeps mask token 0
        stab_thresh = self.cfg.dynamic_multimask_stability_thresh if self.cfg.dynamic_multimask_via_stability else -1.0
        low, idx, iou_sel = ops.sam_select(masks, y[:, 32:], y, multimask, self.cfg.dynamic_multimask_stability_delta,
                                           stab_thresh, NO_OBJ_SCORE, iou_stride=192,
                                           score_stride=192, iou_is_logit=True)
        score = y[:, 0:1].contiguous()
        P1, pb1, P2, pb2, P3, pb3 = w.obj_ptr_proj
        if defer_ptr:
            # the pointer (3-layer MLP on the selected mask token + no-object mix) is not needed by the memory encoder: the
            # caller runs it on its forked tail branch
            def finish_ptr():
                t1 = sk(None, P1[0], pb1[0], M=B, x_ptr=hs.data_ptr() + 4 * 2 * 256, x_rs=Nt * 256, row_select=idx,
                        x_sel_stride=256, act=ACT_RELU)
                ptr = sk(sk(t1, P2[0], pb2[0], act=ACT_RELU), P3[0], pb3[0])
                return ops.objptr_mix_(ptr, score, w.no_obj_ptr)

            return dict(low=low, obj_ptr=None, score=score, iou=iou_sel, masks=masks, iou_logits=y[:, 32:36], idx=idx,
                        finish_ptr=finish_ptr, _keep=(hs, idx))
        t1 = sk(None, P1[0], pb1[0], M=B, x_ptr=hs.data_ptr() + 4 * 2 * 256, x_rs=Nt * 256, row_select=idx,
                x_sel_stride=256, act=ACT_RELU)
        t2 = sk(t1, P2[0], pb2[0], act=ACT_RELU)
        ptr = sk(t2, P3[0], pb3[0])
        ops.objptr_mix_(ptr, score, w.no_obj_ptr)
        return dict(low=low, obj_ptr=ptr, score=score, iou=iou_sel, masks=masks, iou_logits=y[:, 32:36], idx=idx)

    def _two_way_transformer(self, tokens, keys, B, Nt, img_stream=None, const0=None, defer_final_norm=False):
        """TwoWayTransformer (sam/transformer.py:90-135) with one launch per token-side layer: any token count.
        With `img_stream` the image-side work (the fused k / v / q projections of `keys`, image->token attention and its
        output projection + norm) is enqueued there and only meets the token side where the data does: the projection of
        layer l overlaps the token self-attention chain of layer l, the image->token block of layer l overlaps the token
        self-attention of layer l + 1.  On return `keys` is still owned by `img_stream` (the caller continues there)."""
        w = self.w
        T = self.cfg.feat ** 2
        queries = tokens
        ln = lambda x, nb: ops.layernorm(x, nb[0], nb[1], 1e-5, f32=True)[0]
        lnp = lambda nb: (nb[0], nb[1], 1e-5)
        sk = ops.gemm_skinny
        tok = torch.cuda.current_stream()
        par = img_stream is not None
        img_s = img_stream if par else tok
        if par:
            self._handoff(img_s, tok, keys, tokens)
        for l, Lyr in enumerate(w.dec_layers):
            sa, t2i, i2t = Lyr["sa"], Lyr["t2i"], Lyr["i2t"]
            with torch.cuda.stream(img_s):
                img = ops.gemm_f32(keys, Lyr["img_w"], Lyr["img_b"], residual=Lyr["img_pe"], res_mod=T, tf32=_DEC_TF32)  # [B*T, 384]
            if l == 0 and const0 is not None:
                queries, q = const0  # constants of the model on prompt-free frames (Engine.token_constants)
            else:
                if l == 0:
                    qkv = sk(queries, sa["qkv_w"], sa["qkv_b"])
                    o = ops.attn_small(qkv[:, 0:256], qkv[:, 256:512], qkv[:, 512:768], B, 8, Nt, Nt, 32)
                    pre = sk(o, *sa["o"])
                else:
                    # q, k read queries + token_pe, v reads queries: one launch, the positional add limited to 512 columns
                    qkv = sk(queries, sa["qkv_w"], sa["qkv_b"], x2=tokens, x2_cols=512)
                    o = ops.attn_small(qkv[:, 0:256], qkv[:, 256:512], qkv[:, 512:768], B, 8, Nt, Nt, 32)
                    pre = sk(o, *sa["o"], residual=queries)
                # norm1 runs inside the projection that consumes it (LayerNorm on load), which also leaves `queries`
                queries = torch.empty_like(pre)
                q = sk(pre, *t2i["q"], x2=tokens, ln=lnp(Lyr["norms"][0]), ln_out=queries)
            if par:
                self._handoff(tok, img_s, img)
            o = ops.attn_t2i(q, img[:, 0:128], img[:, 128:256], B, Nt, T)
            pre = sk(o, *t2i["o"], residual=queries)
            if B * Nt >= 64:
                # many objects (lock-step videos): the skinny kernel re-streams the 2 x 2 MB of MLP weights once per 8
                # rows; from 64 rows on the tiled fp32 GEMM (weights read once per 64-row tile) is the better kernel
                queries = ln(pre, Lyr["norms"][1])
                m = ops.gemm_f32(queries, *Lyr["mlp"][0], act=ACT_RELU)
                # (K = 2048 on 16 tiles of the SIMT kernel is a 110 us serial reduction: tf32 tensor-core products here,
                # like the decoder's image side)
                pre = ops.gemm_f32(m, *Lyr["mlp"][1], residual=queries, tf32=_DEC_TF32)
            else:
                queries = torch.empty_like(pre)
                m = sk(pre, *Lyr["mlp"][0], act=ACT_RELU, ln=lnp(Lyr["norms"][1]), ln_out=queries)       # norm2 on load
                pre = sk(m, *Lyr["mlp"][1], residual=queries)
            queries = torch.empty_like(pre)
            kv2 = sk(pre, i2t["kv_w"], i2t["kv_b"], x2=tokens, x2_cols=128, ln=lnp(Lyr["norms"][2]),  # norm3 on load
                     ln_out=queries)  # [k (with pe) | v]
            if par:
                self._handoff(img_s, tok, kv2)
            with torch.cuda.stream(img_s):
                o2 = ops.attn_i2t(img[:, 256:384], kv2[:, 0:128], kv2[:, 128:256], B, T, Nt)
                keys = ln(ops.gemm_f32(o2, *i2t["o"], residual=keys, tf32=_DEC_TF32), Lyr["norms"][3])
        fin = w.dec_final
        with torch.cuda.stream(img_s):
            img = ops.gemm_f32(keys, fin["img_w"], fin["img_b"], residual=fin["img_pe"], res_mod=T, tf32=_DEC_TF32)  # [B*T, 256]
        q = sk(queries, *fin["q"], x2=tokens)
        if par:
            self._handoff(tok, img_s, img)
        o = ops.attn_t2i(q, img[:, 0:128], img[:, 128:256], B, Nt, T)
        hs_pre = sk(o, *fin["o"], residual=queries)  # [B*Nt, 256]; norm_final_attn is applied by the consumer
        if not defer_final_norm:
            return ln(hs_pre, w.dec_final_norm), keys
        return hs_pre, keys

    def embed_points(self, coords, labels):
        """PromptEncoder._embed_points with pad=True (prompt_encoder.py:79-103): coords [B,P,2] model pixels."""
        B = coords.shape[0]
        dev = self.w.device
        c = torch.cat([coords.to(dev, F32) + 0.5, torch.zeros(B, 1, 2, device=dev)], dim=1).contiguous()
        lab = torch.cat([labels.to(dev, torch.int32), -torch.ones(B, 1, dtype=torch.int32, device=dev)], dim=1)
        return ops.point_embed(c, lab.contiguous(), self.w.pe_gauss, self.w.point_table, self.cfg.image_size)

    def no_point_tokens(self, B):
        """The two padding tokens used on tracked frames (label -1 point + pad): both not_a_point_embed."""
        return self.w.point_table[4][None, None].expand(B, 2, 256)

    def embed_mask_prompt(self, m128, B):
        """PromptEncoder.mask_downscaling on a [B,1,L,L] dense prompt (L = image_size / 4) -> fp32 [B*T, 256]."""
        w = self.w
        L = self.cfg.image_size // 4
        x = m128.reshape(B, L, L, 1).contiguous()
        (w0, b0, lw0, lb0), (w1, b1, lw1, lb1) = w.pm_convs
        x, H, W = ops.conv2d_small(x, w0, b0, B, L, L, 1, 4, 2, 2, 0, ln=(lw0, lb0), gelu=True)
        x, H, W = ops.conv2d_small(x, w1, b1, B, H, W, 4, 16, 2, 2, 0, ln=(lw1, lb1), gelu=True)
        return ops.gemm_f32(x, *w.pm_out)

    def mask_as_output(self, feat, feat_s0, feat_s1, mask512, B):
        """SAM2Base._use_mask_as_output (sam2_base.py:1168-1218); mask512 fp32 {0,1} [B,1,512,512]."""
        w = self.w
        S, T = self.cfg.image_size, self.cfg.feat ** 2
        L = S // 4
        high = mask512 * 20.0 - 10.0
        low = ops.resize_bilinear_aa(high, L, L)
        md, _, _ = ops.conv2d_small(mask512.reshape(B, S, S, 1).contiguous(), *w.mask_downsample, B, S, S, 1,
                                    1, 4, 4, 0)
        dense = self.embed_mask_prompt(md.view(B, 1, L, L), B)
        pix, _ = ops.axpby(feat, None, rows=B * T, x_mod=T)
        o = self.sam_heads(pix, feat_s0, feat_s1, B, self.no_point_tokens(B), dense=dense, multimask=False)
        present = (mask512.flatten(1) > 0).any(dim=1, keepdim=True)
        score = present.to(F32) * 20.0 - 10.0
        ptr = torch.where(present, o["obj_ptr"], w.no_obj_ptr[None].expand(B, -1))
        return dict(low=low, high=high, obj_ptr=ptr, score=score)

    # ---------------------------------------------------------------- memory encoder
    def encode_memory(self, feat_bf16, mask_in512, score, B, ctrl=None, pix_proj=None, group=0):
        """mask_in512 fp32 [B,512,512]: already sigmoid*20-10 / binarised (see Engine.mem_mask_input);
        feat_bf16 [1024,256] raw frame features.  Returns bf16 token-major memory [B,1024,64], or writes it into the
        frame store slot named by `ctrl`
        (_encode_new_memory sam2_base.py:1450-1498, MemoryEncoder memory_encoder.py:158-181)."""
        w = self.w
        S, fs = self.cfg.image_size, self.cfg.feat
        T = fs * fs
        H, W, Cin = S, S, 1
        for si, ((cw, cb, lw, lb), Cout) in enumerate(zip(w.md_convs, (4, 16, 64))):
            if si == 0 and isinstance(mask_in512, tuple):
                # (low-res logits, post mode): the 512^2 upsampling + sigmoid / binarise is evaluated inside the first
                # convolution's footprint load -- the upsampled mask is never written (see mem_mask_input(..., lazy=True))
                low, post = mask_in512
                x, H, W = ops.conv2d_mask_first(low, post, self.cfg.sigmoid_scale_for_mem_enc,
                                                self.cfg.sigmoid_bias_for_mem_enc, cw, cb, B, H, W, 3, 2, 1, ln=(lw, lb),
                                                gelu=True)
            else:
                if si == 0:
                    x = mask_in512.reshape(B, S, S, 1)
                x, H, W = ops.conv2d_small(x, cw, cb, B, H, W, Cin, Cout, 3, 2, 1, ln=(lw, lb), gelu=True)
            Cin = Cout
        A4 = ops.im2col_nhwc(x, B, 2 * fs, 2 * fs, 64, 3, 2, 1)
        _, c4n = ops.gemm_bf16(A4, w.md_conv3_w, bias=w.md_conv3_b, ln=(w.md_ln3[0], w.md_ln3[1], 1e-6, True))
        pp = pix_proj  # pix_feat_proj, shared by all objects (precomputed by the caller on a forked branch, or here)
        if pp is None:
            pp, _ = ops.gemm_bf16(feat_bf16.view(-1, 256), *w.pix_proj, f32=True)
        x, _ = ops.gemm_bf16(c4n, w.md_out[0], bias=w.md_out[1], residual=pp, res_mod=T, f32=True,
                             res_div=group * T)
        xb = None
        for i, Lf in enumerate(w.fuser):
            h = ops.dwconv7_ln(x, Lf["dw_w"], Lf["dw_b"], Lf["ln"][0], Lf["ln"][1], B, fs, fs)
            _, t = ops.gemm_bf16(h, Lf["pw1"][0], bias=Lf["pw1"][1], act=ACT_GELU, bf16=True)
            x, xb = ops.gemm_bf16(t, Lf["pw2"][0], bias=Lf["pw2"][1], col_scale=Lf["gamma"], residual=x, f32=True,
                                  bf16=(i == len(w.fuser) - 1))
        out, _ = ops.gemm_bf16(xb, w.mem_out[0], bias=w.mem_out[1], f32=True)
        return ops.finalize_memory(out, score, w.no_obj_embed_spatial, B, T=T, ctrl=ctrl)

    def mem_mask_input(self, masks, binarize, non_overlap=False, group=0, lazy=False):
        """Upsample low-res logits [B,1,h,w] to 512^2 (if needed) fused with sigmoid*20-10 or (x>0)*20-10.
        non_overlap: the non-overlapping constraint (sam2_base.py:1466-1471, 1663-1681) is applied to the 512^2 logits
        first, per group of `group` objects (0 = all: the objects of one video).
        lazy: return (low-res logits, post mode) for encode_memory to upsample inside its first convolution, when nothing
        needs the materialised 512^2 mask (no non-overlap constraint)."""
        cfg = self.cfg
        S = cfg.image_size
        post = ops.POST_BINARIZE_AFFINE if binarize else ops.POST_SIGMOID_AFFINE
        if lazy and self.lazy_mask_upsample and not (non_overlap and masks.shape[0] > 1) \
                and tuple(masks.shape[-2:]) != (S, S) and masks.dtype == F32 and masks.is_contiguous():
            return (masks, post)
        if non_overlap and masks.shape[0] > 1:
            hi = masks if tuple(masks.shape[-2:]) == (S, S) else ops.resize_bilinear(masks, S, S)
            return ops.non_overlap(hi, group, post, cfg.sigmoid_scale_for_mem_enc, cfg.sigmoid_bias_for_mem_enc)
        return ops.resize_bilinear(masks, S, S, post, cfg.sigmoid_scale_for_mem_enc, cfg.sigmoid_bias_for_mem_enc)

    def memory_attention_from_tensors(self, feat, mem_frames, tpos_rows, ptrs, ptr_pos, B):
        """Convenience for tests / one-off calls: assemble the bank from explicit tensors (usvm_build_memory)."""
        w = self.w
        k_in, v_in, Nk = ops.build_memory(mem_frames, tpos_rows, w.mem_pos, w.maskmem_tpos, ptrs, ptr_pos, B)
        return self.memory_attention(feat, k_in, v_in, Nk, 0 if ptrs is None else ptrs.shape[1], B)

    # ---------------------------------------------------------------- object pointers
    def obj_ptr_tokens(self, pos_list, ptr_list, max_ptrs, B):
        """Object-pointer tokens + temporal pos enc (sam2_base.py:1361-1417).  ptr_list: fp32 [B,256] each.
        Returns (ptrs fp32 [B, P*4, 64], ptr_pos fp32 [P*4, 64])."""
        dev = self.w.device
        P = len(pos_list)
        ptrs = torch.stack(ptr_list, dim=1).reshape(B, P * 4, 64).contiguous()
        rel = torch.tensor(pos_list, dtype=F32) / float(max_ptrs - 1)
        pe_dim = 128
        dim_t = 10000.0 ** (2 * torch.div(torch.arange(pe_dim, dtype=F32), 2, rounding_mode="floor") / pe_dim)
        e = rel[:, None] / dim_t
        tp = torch.cat([e.sin(), e.cos()], dim=-1).to(dev)  # get_1d_sine_pe (sam2_utils.py:64-74), [P, 256]
        if self.w.tpos_proj is None:
            return ptrs, torch.zeros((P * 4, 64), dtype=F32, device=dev)
        tp = ops.gemm_f32(tp.contiguous(), *self.w.tpos_proj)  # obj_ptr_tpos_proj, [P, 64]
        return ptrs, tp.repeat_interleave(4, dim=0).contiguous()
