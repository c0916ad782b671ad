"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Never imported by the product package.

CPU restatement (plain PyTorch, fp32) of the reference's per-frame video propagation path for
`sam2.1_hiera_t512`, written functionally over the reference's own state-dict keys.  Only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may
import this file.

Parity pinning: the reference ships no tests or golden vectors for this path, so the
oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF: `oracle/make_golden.py` imports the
real reference from /root/reference (build container only), runs it on seeded synthetic clips
and seeded weights and commits the results under tests/golden/; `tests/test_oracle_pinned.py`
checks this file against those fixtures (and, when the mount is present, against the live
reference module by module).

Each function cites the reference file:line it restates (paths relative to /root/reference).

`set_matmul_emulation("bf16"|"tf32"|None)` rounds the *inputs* of every contraction (linear,
conv, attention matmuls) to the given format while accumulating in fp32 -- a CPU model of
what a tensor-core path does, used to calibrate the tolerance the GPU tests state.
"""
import math
from collections import OrderedDict

import torch
import torch.nn.functional as F

NO_OBJ_SCORE = -1024.0  # sam2/modeling/sam2_base.py:22

# --------------------------------------------------------------------------------------
# precision emulation
# --------------------------------------------------------------------------------------
_EMU = {"mode": None}


def set_matmul_emulation(mode):
    assert mode in (None, "bf16", "tf32")
    _EMU["mode"] = mode


def _q(x):
    m = _EMU["mode"]
    if m is None:
        return x
    if m == "bf16":
        return x.to(torch.bfloat16).to(torch.float32)
    # tf32: keep 10 explicit mantissa bits, round-to-nearest-even on the dropped 13
    i = x.contiguous().view(torch.int32)
    lsb = (i >> 13) & 1
    i = (i + 0x0FFF + lsb) & ~0x1FFF
    return i.view(torch.float32)


def linear(x, w, b=None):
    return F.linear(_q(x), _q(w), b)


def conv2d(x, w, b=None, **kw):
    return F.conv2d(_q(x), _q(w), b, **kw)


def sdpa(q, k, v):
    """softmax(q k^T / sqrt(d)) v, fp32 softmax (F.scaled_dot_product_attention semantics,
    hieradet.py:70, transformer.py:270/344)."""
    s = torch.matmul(_q(q), _q(k).transpose(-1, -2)) / math.sqrt(q.shape[-1])
    p = torch.softmax(s, dim=-1)
    return torch.matmul(_q(p), _q(v))


def layer_norm(x, w, b, eps):
    return F.layer_norm(x, (x.shape[-1],), w, b, eps)


def layer_norm_2d(x, w, b, eps=1e-6):
    """Per-pixel LayerNorm over channels of an NCHW map (sam2_utils.py:141-153)."""
    u = x.mean(1, keepdim=True)
    s = (x - u).pow(2).mean(1, keepdim=True)
    x = (x - u) / torch.sqrt(s + eps)
    return w[None, :, None, None] * x + b[None, :, None, None]


# --------------------------------------------------------------------------------------
# model configuration (sam2/configs/sam2.1_hiera_t512.yaml + builder overrides,
# sam2/build_sam.py:108-122)
# --------------------------------------------------------------------------------------
class Cfg:
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
    multimask_min_pt_num = 0
    multimask_max_pt_num = 1
    fill_hole_area = 8
    binarize_mask_from_pts_for_mem_enc = True
    dynamic_multimask_via_stability = True  # builder override (build_sam.py:108-122); False with apply_postprocessing=False
    dynamic_multimask_stability_delta = 0.05
    dynamic_multimask_stability_thresh = 0.98
    non_overlap_masks = False
    non_overlap_masks_for_mem_enc = False


class CfgBPlus(Cfg):
    """sam2.1_hiera_base_plus at 1024^2 (BASELINE configs[4]): the `Hiera` class defaults (hieradet.py:174-200) with
    upstream's embed_dim 112 / 2 heads; feature map 64 x 64 (RoPEAttention feat_sizes), 256^2 low-res masks.  Pinned
    against tests/golden/bplus1024_ct_bidirectional.npz (outputs of the reference's own classes)."""
    image_size = 1024
    embed_dim = 112
    num_heads = 2
    stages = (2, 3, 16, 3)
    global_att_blocks = (12, 16, 20)


class CfgNoPost(Cfg):
    """build_sam2_video_predictor(..., apply_postprocessing=False): none of the builder's overrides
    (build_sam.py:108-122), i.e. the class defaults of SAM2Base / MaskDecoder / SAM2VideoPredictor."""
    dynamic_multimask_via_stability = False
    binarize_mask_from_pts_for_mem_enc = False
    fill_hole_area = 0


def hiera_block_plan(cfg=Cfg):
    """Per-block (dim_in, dim_out, heads, window, pool) following the lagged-window rule of
    Hiera.__init__ (hieradet.py:201-256)."""
    stage_ends = [sum(cfg.stages[:i]) - 1 for i in range(1, len(cfg.stages) + 1)]
    pool_blocks = [e + 1 for e in stage_ends[:-1]][: cfg.q_pool]
    plan, dim, heads, stage = [], cfg.embed_dim, getattr(cfg, "num_heads", 1), 1
    for i in range(sum(cfg.stages)):
        window = cfg.window_spec[stage - 1]
        if i in cfg.global_att_blocks:
            window = 0
        dim_out = dim
        if i - 1 in stage_ends:
            dim_out, heads, stage = dim * 2, heads * 2, stage + 1
        plan.append(dict(dim=dim, dim_out=dim_out, heads=heads, window=window,
                         pool=i in pool_blocks, emit=i in stage_ends))
        dim = dim_out
    return plan


# --------------------------------------------------------------------------------------
# position encodings
# --------------------------------------------------------------------------------------
def sine_pos_2d(h, w, num_pos_feats, temperature=10000.0):
    """PositionEmbeddingSine.forward with normalize=True, scale=2*pi
    (position_encoding.py:79-112) -> [num_pos_feats, h, w]."""
    half = num_pos_feats // 2
    eps, scale = 1e-6, 2 * math.pi
    y = torch.arange(1, h + 1, dtype=torch.float32)
    x = torch.arange(1, w + 1, dtype=torch.float32)
    y = y / (y[-1] + eps) * scale
    x = x / (x[-1] + eps) * scale
    idx = torch.arange(half, dtype=torch.float32)
    dim_t = temperature ** (2 * torch.div(idx, 2, rounding_mode="floor") / half)
    px = x[:, None] / dim_t  # [w, half]
    py = y[:, None] / dim_t  # [h, half]
    px = torch.stack((px[:, 0::2].sin(), px[:, 1::2].cos()), dim=2).flatten(1)
    py = torch.stack((py[:, 0::2].sin(), py[:, 1::2].cos()), dim=2).flatten(1)
    pos = torch.cat((py[:, None, :].expand(h, w, half), px[None, :, :].expand(h, w, half)), dim=2)
    return pos.permute(2, 0, 1).contiguous()


def sine_pos_1d(pos_inds, dim, temperature=10000.0):
    """get_1d_sine_pe (sam2_utils.py:64-74)."""
    pe_dim = dim // 2
    idx = torch.arange(pe_dim, dtype=torch.float32)
    dim_t = temperature ** (2 * torch.div(idx, 2, rounding_mode="floor") / pe_dim)
    e = pos_inds.unsqueeze(-1) / dim_t
    return torch.cat([e.sin(), e.cos()], dim=-1)


def axial_rope_table(dim, end_x, end_y, theta=10000.0):
    """compute_axial_cis (position_encoding.py:174-183) as separate cos / sin tables
    [end_x*end_y, dim/2]; pair j<dim/4 rotates with the x coordinate, the rest with y."""
    freqs = 1.0 / (theta ** (torch.arange(0, dim, 4)[: dim // 4].float() / dim))
    t = torch.arange(end_x * end_y, dtype=torch.float32)
    tx = (t % end_x).float()
    ty = torch.div(t, end_x, rounding_mode="floor").float()
    ang = torch.cat([torch.outer(tx, freqs), torch.outer(ty, freqs)], dim=-1)
    return torch.cos(ang), torch.sin(ang)


def apply_rope(x, cos, sin):
    """Complex multiply of adjacent channel pairs (apply_rotary_enc,
    position_encoding.py:194-221).  x [..., N, d]; cos/sin [N, d/2]."""
    xr = x.float().reshape(*x.shape[:-1], -1, 2)
    a, b = xr[..., 0], xr[..., 1]
    out = torch.stack((a * cos - b * sin, a * sin + b * cos), dim=-1)
    return out.flatten(-2)


def random_fourier_pe(coords01, gauss):
    """PositionEmbeddingRandom._pe_encoding (position_encoding.py:127-134)."""
    c = (2 * coords01 - 1) @ gauss
    c = 2 * math.pi * c
    return torch.cat([torch.sin(c), torch.cos(c)], dim=-1)


# --------------------------------------------------------------------------------------
# the model
# --------------------------------------------------------------------------------------
class RefModel:
    """Functional restatement of SAM2Base (sam2_base.py:764-1681) over a state dict."""

    def __init__(self, state_dict, cfg=Cfg):
        self.sd = {k: v.detach().to(torch.float32).cpu() for k, v in state_dict.items()}
        self.cfg = cfg
        self.plan = hiera_block_plan(cfg)
        self.fs = cfg.image_size // 16  # side of the stride-16 feature map
        self.rope_cos, self.rope_sin = axial_rope_table(cfg.d_model, self.fs, self.fs)
        self._sine_cache = {}

    def p(self, name):
        return self.sd[name]

    # ---------------- image encoder ----------------
    def _hiera_pos_embed(self, h, w):
        """Hiera._get_pos_embed (hieradet.py:273-281) -> [1, h, w, C]."""
        pe = F.interpolate(self.p("image_encoder.trunk.pos_embed"), size=(h, w), mode="bicubic")
        win = self.p("image_encoder.trunk.pos_embed_window")
        pe = pe + win.tile(1, 1, h // win.shape[2], w // win.shape[3])
        return pe.permute(0, 2, 3, 1)

    @staticmethod
    def _to_windows(x, ws):
        """window_partition (backbones/utils.py:17-37); zero padding on the bottom/right."""
        B, H, W, C = x.shape
        ph, pw = (ws - H % ws) % ws, (ws - W % ws) % ws
        if ph or pw:
            x = F.pad(x, (0, 0, 0, pw, 0, ph))
        Hp, Wp = H + ph, W + pw
        x = x.view(B, Hp // ws, ws, Wp // ws, ws, C).permute(0, 1, 3, 2, 4, 5)
        return x.reshape(-1, ws, ws, C), (Hp, Wp)

    @staticmethod
    def _from_windows(xw, ws, pad_hw, hw):
        """window_unpartition (backbones/utils.py:40-61)."""
        Hp, Wp = pad_hw
        H, W = hw
        B = xw.shape[0] // ((Hp // ws) * (Wp // ws))
        x = xw.reshape(B, Hp // ws, Wp // ws, ws, ws, -1).permute(0, 1, 3, 2, 4, 5)
        return x.reshape(B, Hp, Wp, -1)[:, :H, :W, :]

    @staticmethod
    def _maxpool2(x):  # NHWC 2x2 / stride 2 (do_pool, hieradet.py:25-36)
        return F.max_pool2d(x.permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1)

    def _hiera_block(self, i, x):
        """MultiScaleBlock.forward + MultiScaleAttention.forward (hieradet.py:56-81,134-166)."""
        b = self.plan[i]
        pre = f"image_encoder.trunk.blocks.{i}."
        shortcut = x
        x = layer_norm(x, self.p(pre + "norm1.weight"), self.p(pre + "norm1.bias"), 1e-6)
        if b["dim"] != b["dim_out"]:
            shortcut = self._maxpool2(linear(x, self.p(pre + "proj.weight"), self.p(pre + "proj.bias")))
        ws = b["window"]
        H, W = x.shape[1:3]
        if ws > 0:
            x, pad_hw = self._to_windows(x, ws)
        Bw, Hw, Ww, _ = x.shape
        heads = b["heads"]
        qkv = linear(x, self.p(pre + "attn.qkv.weight"), self.p(pre + "attn.qkv.bias"))
        qkv = qkv.reshape(Bw, Hw * Ww, 3, heads, -1)
        q, k, v = qkv.unbind(2)
        if b["pool"]:
            q = self._maxpool2(q.reshape(Bw, Hw, Ww, -1))
            Hw, Ww = q.shape[1:3]
            q = q.reshape(Bw, Hw * Ww, heads, -1)
        o = sdpa(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2)).transpose(1, 2)
        o = o.reshape(Bw, Hw, Ww, -1)
        o = linear(o, self.p(pre + "attn.proj.weight"), self.p(pre + "attn.proj.bias"))
        if b["pool"]:
            ws = ws // 2
            H, W = shortcut.shape[1:3]
            pad_hw = (H + (ws - H % ws) % ws, W + (ws - W % ws) % ws) if ws > 0 else (H, W)
        if b["window"] > 0:
            o = self._from_windows(o, ws, pad_hw, (H, W))
        x = shortcut + o
        h = layer_norm(x, self.p(pre + "norm2.weight"), self.p(pre + "norm2.bias"), 1e-6)
        h = F.gelu(linear(h, self.p(pre + "mlp.layers.0.weight"), self.p(pre + "mlp.layers.0.bias")))
        h = linear(h, self.p(pre + "mlp.layers.1.weight"), self.p(pre + "mlp.layers.1.bias"))
        return x + h

    def hiera(self, img):
        """PatchEmbed + Hiera.forward (backbones/utils.py:90-94, hieradet.py:283-299);
        returns the four stage outputs as NHWC maps."""
        x = conv2d(img, self.p("image_encoder.trunk.patch_embed.proj.weight"),
                   self.p("image_encoder.trunk.patch_embed.proj.bias"), stride=4, padding=3)
        x = x.permute(0, 2, 3, 1)
        x = x + self._hiera_pos_embed(x.shape[1], x.shape[2])
        outs = []
        for i, b in enumerate(self.plan):
            x = self._hiera_block(i, x)
            if b["emit"]:
                outs.append(x)
        return outs

    def sine_pos(self, h, w, feats):
        key = (h, w, feats)
        if key not in self._sine_cache:
            self._sine_cache[key] = sine_pos_2d(h, w, feats)
        return self._sine_cache[key]

    def forward_image(self, img):
        """ImageEncoder.forward + FpnNeck.forward + SAM2Base.forward_image
        (image_encoder.py:31-44,104-136; sam2_base.py:1220-1232).

        Returns feat_s0 [B,32,128,128], feat_s1 [B,64,64,64], feat [B,256,32,32] and the sine
        position encoding of the 32x32 level [256,32,32]."""
        xs = [t.permute(0, 3, 1, 2) for t in self.hiera(img.float())]
        n = len(xs) - 1
        lat = [None] * (n + 1)
        for lvl in range(n + 1):
            pre = f"image_encoder.neck.convs.{n - lvl}.conv."
            lat[lvl] = conv2d(xs[lvl], self.p(pre + "weight"), self.p(pre + "bias"))
        # top-down only into level 2 (fpn_top_down_levels [2, 3]; level 3 has no parent)
        lvl2 = lat[2] + F.interpolate(lat[3], scale_factor=2.0, mode="nearest")
        d = "sam_mask_decoder."
        feat_s0 = conv2d(lat[0], self.p(d + "conv_s0.weight"), self.p(d + "conv_s0.bias"))
        feat_s1 = conv2d(lat[1], self.p(d + "conv_s1.weight"), self.p(d + "conv_s1.bias"))
        return dict(feat_s0=feat_s0, feat_s1=feat_s1, feat=lvl2,
                    pos=self.sine_pos(lvl2.shape[2], lvl2.shape[3], 256))

    # ---------------- memory attention ----------------
    def _attn_proj(self, pre, q, k, v, heads):
        q = linear(q, self.p(pre + "q_proj.weight"), self.p(pre + "q_proj.bias"))
        k = linear(k, self.p(pre + "k_proj.weight"), self.p(pre + "k_proj.bias"))
        v = linear(v, self.p(pre + "v_proj.weight"), self.p(pre + "v_proj.bias"))
        split = lambda t: t.reshape(t.shape[0], t.shape[1], heads, -1).transpose(1, 2)
        return split(q), split(k), split(v)

    def _attn_out(self, pre, o):
        o = o.transpose(1, 2).reshape(o.shape[0], o.shape[2], -1)
        return linear(o, self.p(pre + "out_proj.weight"), self.p(pre + "out_proj.bias"))

    def _rope_attention(self, pre, q, k, v, num_k_exclude_rope=0, repeat_k=False):
        """RoPEAttention.forward (transformer.py:311-360), 1 head of 256."""
        q, k, v = self._attn_proj(pre, q, k, v, 1)
        q = apply_rope(q, self.rope_cos, self.rope_sin)
        n_rope = k.shape[-2] - num_k_exclude_rope
        if n_rope > 0:
            r = n_rope // q.shape[-2] if repeat_k else 1
            k = torch.cat([apply_rope(k[:, :, :n_rope], self.rope_cos.repeat(r, 1),
                                      self.rope_sin.repeat(r, 1)), k[:, :, n_rope:]], dim=2)
        return self._attn_out(pre, sdpa(q, k, v))

    def memory_attention(self, curr, curr_pos, memory, memory_pos, num_obj_ptr_tokens):
        """MemoryAttention.forward + MemoryAttentionLayer.forward
        (memory_attention.py:58-99,119-169); batch-first [B, N, C] tensors."""
        x = curr + 0.1 * curr_pos
        for l in range(4):
            pre = f"memory_attention.layers.{l}."
            h = layer_norm(x, self.p(pre + "norm1.weight"), self.p(pre + "norm1.bias"), 1e-5)
            x = x + self._rope_attention(pre + "self_attn.", h, h, h)
            h = layer_norm(x, self.p(pre + "norm2.weight"), self.p(pre + "norm2.bias"), 1e-5)
            x = x + self._rope_attention(pre + "cross_attn_image.", h, memory + memory_pos, memory,
                                         num_k_exclude_rope=num_obj_ptr_tokens, repeat_k=True)
            h = layer_norm(x, self.p(pre + "norm3.weight"), self.p(pre + "norm3.bias"), 1e-5)
            h = F.relu(linear(h, self.p(pre + "linear1.weight"), self.p(pre + "linear1.bias")))
            x = x + linear(h, self.p(pre + "linear2.weight"), self.p(pre + "linear2.bias"))
        return layer_norm(x, self.p("memory_attention.norm.weight"), self.p("memory_attention.norm.bias"), 1e-5)

    # ---------------- prompt encoder ----------------
    def dense_pe(self):
        """PromptEncoder.get_dense_pe (prompt_encoder.py:68-77, position_encoding.py:136-148)."""
        g = self.p("sam_prompt_encoder.pe_layer.positional_encoding_gaussian_matrix")
        fs = self.fs
        c = (torch.arange(fs, dtype=torch.float32) + 0.5) / fs
        grid = torch.stack([c[None, :].expand(fs, fs), c[:, None].expand(fs, fs)], dim=-1)
        return random_fourier_pe(grid, g).permute(2, 0, 1)[None]

    def embed_points(self, coords, labels):
        """PromptEncoder._embed_points with pad=True (prompt_encoder.py:79-103)."""
        pe = "sam_prompt_encoder."
        B = coords.shape[0]
        coords = torch.cat([coords + 0.5, torch.zeros(B, 1, 2)], dim=1)
        labels = torch.cat([labels, -torch.ones(B, 1, dtype=labels.dtype)], dim=1)
        e = random_fourier_pe(coords / self.cfg.image_size,
                              self.p(pe + "pe_layer.positional_encoding_gaussian_matrix"))
        e = torch.where((labels == -1)[..., None], self.p(pe + "not_a_point_embed.weight").expand_as(e), e)
        for lab in range(4):
            e = e + (labels == lab)[..., None].float() * self.p(pe + f"point_embeddings.{lab}.weight")
        return e

    def embed_mask(self, m):
        """PromptEncoder.mask_downscaling (prompt_encoder.py:57-66)."""
        pre = "sam_prompt_encoder.mask_downscaling."
        x = conv2d(m, self.p(pre + "0.weight"), self.p(pre + "0.bias"), stride=2)
        x = F.gelu(layer_norm_2d(x, self.p(pre + "1.weight"), self.p(pre + "1.bias")))
        x = conv2d(x, self.p(pre + "3.weight"), self.p(pre + "3.bias"), stride=2)
        x = F.gelu(layer_norm_2d(x, self.p(pre + "4.weight"), self.p(pre + "4.bias")))
        return conv2d(x, self.p(pre + "6.weight"), self.p(pre + "6.bias"))

    # ---------------- mask decoder ----------------
    def _attention(self, pre, q, k, v, heads=8):
        """Attention.forward (transformer.py:257-286)."""
        q, k, v = self._attn_proj(pre, q, k, v, heads)
        return self._attn_out(pre, sdpa(q, k, v))

    def _mlp(self, pre, x, n, act=F.relu, sigmoid=False):
        """MLP.forward (sam2_utils.py:131-136)."""
        for i in range(n):
            x = linear(x, self.p(pre + f"layers.{i}.weight"), self.p(pre + f"layers.{i}.bias"))
            if i < n - 1:
                x = act(x)
        return torch.sigmoid(x) if sigmoid else x

    def two_way_transformer(self, src, pos, tokens):
        """TwoWayTransformer.forward / TwoWayAttentionBlock.forward (transformer.py:90-212)."""
        t = "sam_mask_decoder.transformer."
        keys = src.flatten(2).permute(0, 2, 1)
        key_pe = pos.flatten(2).permute(0, 2, 1)
        queries, query_pe = tokens, tokens
        ln = lambda x, n: layer_norm(x, self.p(n + ".weight"), self.p(n + ".bias"), 1e-5)
        for l in range(2):
            pre = t + f"layers.{l}."
            if l == 0:
                queries = self._attention(pre + "self_attn.", queries, queries, queries)
            else:
                q = queries + query_pe
                queries = queries + self._attention(pre + "self_attn.", q, q, queries)
            queries = ln(queries, pre + "norm1")
            q, k = queries + query_pe, keys + key_pe
            queries = ln(queries + self._attention(pre + "cross_attn_token_to_image.", q, k, keys), pre + "norm2")
            queries = ln(queries + self._mlp(pre + "mlp.", queries, 2), pre + "norm3")
            q, k = queries + query_pe, keys + key_pe
            keys = ln(keys + self._attention(pre + "cross_attn_image_to_token.", k, q, queries), pre + "norm4")
        q, k = queries + query_pe, keys + key_pe
        queries = ln(queries + self._attention(t + "final_attn_token_to_image.", q, k, keys), t + "norm_final_attn")
        return queries, keys

    def mask_decoder(self, feat, sparse, dense, feat_s0, feat_s1, multimask_output):
        """MaskDecoder.forward / predict_masks (mask_decoder.py:110-245)."""
        d = "sam_mask_decoder."
        B = sparse.shape[0]
        out_tokens = torch.cat([self.p(d + "obj_score_token.weight"), self.p(d + "iou_token.weight"),
                                self.p(d + "mask_tokens.weight")], dim=0)
        tokens = torch.cat([out_tokens[None].expand(B, -1, -1), sparse], dim=1)
        src = feat + dense
        pos = self.dense_pe().expand(B, -1, -1, -1)
        b, c, h, w = src.shape
        hs, src = self.two_way_transformer(src, pos, tokens)
        iou_token_out = hs[:, 1]
        mask_tokens_out = hs[:, 2:6]
        src = src.transpose(1, 2).reshape(b, c, h, w)
        up = d + "output_upscaling."
        x = F.conv_transpose2d(_q(src), _q(self.p(up + "0.weight")), self.p(up + "0.bias"), stride=2) + feat_s1
        x = F.gelu(layer_norm_2d(x, self.p(up + "1.weight"), self.p(up + "1.bias")))
        x = F.gelu(F.conv_transpose2d(_q(x), _q(self.p(up + "3.weight")), self.p(up + "3.bias"), stride=2) + feat_s0)
        hyper = torch.stack([self._mlp(d + f"output_hypernetworks_mlps.{i}.", mask_tokens_out[:, i], 3)
                             for i in range(4)], dim=1)
        b, c, h, w = x.shape
        masks = torch.matmul(_q(hyper), _q(x.reshape(b, c, h * w))).reshape(b, -1, h, w)
        iou = self._mlp(d + "iou_prediction_head.", iou_token_out, 3, sigmoid=True)
        score = self._mlp(d + "pred_obj_score_head.", hs[:, 0], 3)
        if multimask_output:
            return masks[:, 1:], iou[:, 1:], mask_tokens_out[:, 1:], score
        if self.cfg.dynamic_multimask_via_stability:  # mask_decoder.py:160-166
            masks, iou = self._stability_select(masks, iou)
        else:
            masks, iou = masks[:, 0:1], iou[:, 0:1]
        return masks, iou, mask_tokens_out[:, 0:1], score

    def _stability_select(self, masks, iou):
        """MaskDecoder._dynamic_multimask_via_stability (mask_decoder.py:247-295)."""
        delta, thresh = self.cfg.dynamic_multimask_stability_delta, self.cfg.dynamic_multimask_stability_thresh
        best = torch.argmax(iou[:, 1:], dim=-1)
        bi = torch.arange(masks.shape[0])
        best_mask, best_iou = masks[:, 1:][bi, best][:, None], iou[:, 1:][bi, best][:, None]
        single = masks[:, 0:1].flatten(-2)
        area_i = (single > delta).sum(-1).float()
        area_u = (single > -delta).sum(-1).float()
        stable = torch.where(area_u > 0, area_i / area_u, torch.ones_like(area_u)) >= thresh
        return (torch.where(stable[..., None, None], masks[:, 0:1], best_mask),
                torch.where(stable, iou[:, 0:1], best_iou))

    def sam_heads(self, pix_feat, feat_s0, feat_s1, point_inputs=None, mask_inputs=None,
                  multimask_output=False):
        """SAM2Base._forward_sam_heads (sam2_base.py:1010-1166)."""
        B = pix_feat.shape[0]
        if point_inputs is not None:
            coords, labels = point_inputs["point_coords"], point_inputs["point_labels"]
        else:
            coords, labels = torch.zeros(B, 1, 2), -torch.ones(B, 1, dtype=torch.int32)
        sparse = self.embed_points(coords.float(), labels)
        if mask_inputs is not None:
            low_hw = (self.cfg.image_size // 4, self.cfg.image_size // 4)
            if tuple(mask_inputs.shape[-2:]) != low_hw:
                mask_inputs = F.interpolate(mask_inputs.float(), size=low_hw, mode="bilinear",
                                            align_corners=False, antialias=True)
            dense = self.embed_mask(mask_inputs)
        else:
            dense = self.p("sam_prompt_encoder.no_mask_embed.weight").reshape(1, -1, 1, 1).expand(B, -1, self.fs, self.fs)
        low_multi, ious, tokens, score = self.mask_decoder(pix_feat, sparse, dense, feat_s0, feat_s1,
                                                           multimask_output)
        appearing = score > 0
        low_multi = torch.where(appearing[:, None, None], low_multi, torch.full_like(low_multi, NO_OBJ_SCORE))
        high_multi = F.interpolate(low_multi, size=(self.cfg.image_size, self.cfg.image_size), mode="bilinear",
                                   align_corners=False)
        token = tokens[:, 0]
        if multimask_output:
            best = torch.argmax(ious, dim=-1)
            bi = torch.arange(B)
            low, high = low_multi[bi, best][:, None], high_multi[bi, best][:, None]
            token = tokens[bi, best]
        else:
            low, high = low_multi, high_multi
        obj_ptr = self._mlp("obj_ptr_proj.", token, 3)
        lam = appearing.float()
        obj_ptr = lam * obj_ptr + (1 - lam) * self.p("no_obj_ptr")
        return dict(low_multi=low_multi, ious=ious, low=low, high=high, obj_ptr=obj_ptr, score=score)

    def mask_as_output(self, pix_feat, feat_s0, feat_s1, mask_inputs):
        """SAM2Base._use_mask_as_output (sam2_base.py:1168-1218)."""
        m = mask_inputs.float()
        high = m * 20.0 - 10.0
        low = F.interpolate(high, size=(high.shape[-2] // 4, high.shape[-1] // 4), mode="bilinear",
                            align_corners=False, antialias=True)
        md = conv2d(m, self.p("mask_downsample.weight"), self.p("mask_downsample.bias"), stride=4)
        ptr = self.sam_heads(pix_feat, feat_s0, feat_s1, mask_inputs=md)["obj_ptr"]
        lam = (m.flatten(1) > 0).any(dim=1)[:, None].float()
        score = 20.0 * lam - 10.0
        ptr = lam * ptr + (1 - lam) * self.p("no_obj_ptr")
        return dict(low=low, high=high, obj_ptr=ptr, score=score, ious=torch.ones(m.shape[0], 1))

    # ---------------- memory encoder ----------------
    def encode_memory(self, pix_feat, high_res_masks, score, is_mask_from_pts):
        """SAM2Base._encode_new_memory + MemoryEncoder.forward
        (sam2_base.py:1450-1498, memory_encoder.py:17-181) -> fp32 [B,64,32,32]."""
        cfg = self.cfg
        if cfg.non_overlap_masks_for_mem_enc:
            high_res_masks = non_overlapping(high_res_masks)
        if cfg.binarize_mask_from_pts_for_mem_enc and is_mask_from_pts:
            m = (high_res_masks > 0).float()
        else:
            m = torch.sigmoid(high_res_masks)
        m = m * cfg.sigmoid_scale_for_mem_enc + cfg.sigmoid_bias_for_mem_enc
        e = "memory_encoder."
        ds = e + "mask_downsampler.encoder."
        for i in range(4):
            m = conv2d(m, self.p(ds + f"{3 * i}.weight"), self.p(ds + f"{3 * i}.bias"), stride=2, padding=1)
            m = F.gelu(layer_norm_2d(m, self.p(ds + f"{3 * i + 1}.weight"), self.p(ds + f"{3 * i + 1}.bias")))
        m = conv2d(m, self.p(ds + "12.weight"), self.p(ds + "12.bias"))
        x = conv2d(pix_feat, self.p(e + "pix_feat_proj.weight"), self.p(e + "pix_feat_proj.bias")) + m
        for l in range(2):
            f = e + f"fuser.layers.{l}."
            h = F.conv2d(_q(x), _q(self.p(f + "dwconv.weight")), self.p(f + "dwconv.bias"), padding=3,
                         groups=x.shape[1])
            h = layer_norm_2d(h, self.p(f + "norm.weight"), self.p(f + "norm.bias")).permute(0, 2, 3, 1)
            h = F.gelu(linear(h, self.p(f + "pwconv1.weight"), self.p(f + "pwconv1.bias")))
            h = linear(h, self.p(f + "pwconv2.weight"), self.p(f + "pwconv2.bias")) * self.p(f + "gamma")
            x = x + h.permute(0, 3, 1, 2)
        x = conv2d(x, self.p(e + "out_proj.weight"), self.p(e + "out_proj.bias"))
        if not getattr(self.cfg, "no_obj_embed_spatial", True):  # EfficientTAM: the parameter does not exist
            return x
        absent = 1 - (score > 0).float()
        return x + absent[..., None, None] * self.p("no_obj_embed_spatial")[..., None, None]

    # ---------------- memory bank assembly ----------------
    def condition_on_memory(self, frame_idx, is_init_cond_frame, feat, pos, output_dict, num_frames,
                            reverse=False):
        """SAM2Base._prepare_memory_conditioned_features (sam2_base.py:1271-1448).
        feat [B,256,32,32]; returns the memory-conditioned map of the same shape."""
        cfg = self.cfg
        B, C, H, W = feat.shape
        if is_init_cond_frame:  # directly_add_no_mem_embed
            return feat + self.p("no_mem_embed").reshape(1, C, 1, 1)
        cond = output_dict["cond_frame_outputs"]
        sel, unsel = select_closest_cond_frames(frame_idx, cond, cfg.max_cond_frames_in_attn)
        entries = [(0, o) for o in sel.values()]
        r = cfg.memory_temporal_stride_for_eval
        for t_pos in range(1, cfg.num_maskmem):
            t_rel = cfg.num_maskmem - t_pos
            if t_rel == 1:
                prev = frame_idx + t_rel if reverse else frame_idx - t_rel
            elif not reverse:
                prev = ((frame_idx - 2) // r) * r - (t_rel - 2) * r
            else:
                prev = -(-(frame_idx + 2) // r) * r + (t_rel - 2) * r
            o = output_dict["non_cond_frame_outputs"].get(prev, None)
            if o is None:
                o = unsel.get(prev, None)
            entries.append((t_pos, o))
        mem, mem_pos = [], []
        tpos = self.p("maskmem_tpos_enc")
        for t_pos, o in entries:
            if o is None:
                continue
            mem.append(o["maskmem_features"].float().flatten(2).permute(0, 2, 1))
            enc = o["maskmem_pos_enc"][-1].flatten(2).permute(0, 2, 1)
            mem_pos.append(enc + tpos[cfg.num_maskmem - t_pos - 1].reshape(1, 1, -1))
        # object pointers
        max_ptrs = min(num_frames, cfg.max_obj_ptrs_in_encoder)
        sign = -1 if reverse else 1
        pos_ptrs = [((frame_idx - t) * sign, o["obj_ptr"]) for t, o in sel.items()
                    if (t >= frame_idx if reverse else t <= frame_idx)]
        for t_diff in range(1, max_ptrs):
            t = frame_idx + t_diff if reverse else frame_idx - t_diff
            if t < 0 or t >= num_frames:
                break
            o = output_dict["non_cond_frame_outputs"].get(t, unsel.get(t, None))
            if o is not None:
                pos_ptrs.append((t_diff, o["obj_ptr"]))
        n_ptr_tokens = 0
        if pos_ptrs:
            plist, ptrs = zip(*pos_ptrs)
            ptrs = torch.stack(ptrs, dim=1)  # [B, P, C]
            if getattr(cfg, "add_tpos_enc_to_obj_ptrs", True):
                tp = sine_pos_1d(torch.tensor(plist, dtype=torch.float32) / (max_ptrs - 1), C)
                tp = linear(tp, self.p("obj_ptr_tpos_proj.weight"), self.p("obj_ptr_tpos_proj.bias"))
            else:  # sam2_base.py:1403-1404 / efficienttam_base.py: zero position encoding for the pointers
                tp = torch.zeros((len(plist), cfg.mem_dim))
            split = C // cfg.mem_dim
            mem.append(ptrs.reshape(B, -1, cfg.mem_dim))
            mem_pos.append(tp.repeat_interleave(split, dim=0)[None].expand(B, -1, -1))
            n_ptr_tokens = ptrs.shape[1] * split
        memory, memory_pos = torch.cat(mem, dim=1), torch.cat(mem_pos, dim=1)
        curr = feat.flatten(2).permute(0, 2, 1)
        curr_pos = pos.flatten(1).permute(1, 0)[None].expand(B, -1, -1)
        out = self.memory_attention(curr, curr_pos, memory, memory_pos, n_ptr_tokens)
        return out.permute(0, 2, 1).reshape(B, C, H, W)

    # ---------------- one frame ----------------
    def track_step(self, frame_idx, is_init_cond_frame, feats, B, point_inputs, mask_inputs, output_dict,
                   num_frames, reverse=False, run_mem_encoder=True, prev_sam_mask_logits=None):
        """SAM2Base.track_step / _track_step (sam2_base.py:1500-1651)."""
        ex = lambda t: t.expand(B, -1, -1, -1)
        feat, s0, s1 = ex(feats["feat"]), ex(feats["feat_s0"]), ex(feats["feat_s1"])
        if mask_inputs is not None:
            o = self.mask_as_output(feat, s0, s1, mask_inputs)
        else:
            pix = self.condition_on_memory(frame_idx, is_init_cond_frame, feat, feats["pos"], output_dict,
                                           num_frames, reverse)
            dense_in = prev_sam_mask_logits if prev_sam_mask_logits is not None else None
            n_pts = 0 if point_inputs is None else point_inputs["point_labels"].shape[1]
            multimask = self.cfg.multimask_min_pt_num <= n_pts <= self.cfg.multimask_max_pt_num
            o = self.sam_heads(pix, s0, s1, point_inputs=point_inputs, mask_inputs=dense_in,
                               multimask_output=multimask)
        out = dict(pred_masks=o["low"], pred_masks_high_res=o["high"], obj_ptr=o["obj_ptr"],
                   object_score_logits=o["score"], ious=o["ious"], maskmem_features=None, maskmem_pos_enc=None)
        if run_mem_encoder:
            mf = self.encode_memory(feat, o["high"], o["score"], is_mask_from_pts=point_inputs is not None)
            out["maskmem_features"] = mf
            out["maskmem_pos_enc"] = [self.sine_pos(self.fs, self.fs, 64)[None].expand(B, -1, -1, -1)]
        return out


def select_closest_cond_frames(frame_idx, cond, max_num):
    """sam2_utils.py:19-61."""
    if max_num == -1 or len(cond) <= max_num:
        return cond, {}
    chosen = {}
    before = max((t for t in cond if t < frame_idx), default=None)
    if before is not None:
        chosen[before] = cond[before]
    after = min((t for t in cond if t >= frame_idx), default=None)
    if after is not None:
        chosen[after] = cond[after]
    rest = sorted((t for t in cond if t not in chosen), key=lambda t: abs(t - frame_idx))
    for t in rest[: max_num - len(chosen)]:
        chosen[t] = cond[t]
    return chosen, {t: v for t, v in cond.items() if t not in chosen}


def non_overlapping(pred_masks):
    """SAM2Base._apply_non_overlapping_constraints (sam2_base.py:1663-1681)."""
    if pred_masks.shape[0] == 1:
        return pred_masks
    winner = torch.argmax(pred_masks, dim=0, keepdim=True)
    keep = winner == torch.arange(pred_masks.shape[0])[:, None, None, None]
    return torch.where(keep, pred_masks, torch.clamp(pred_masks, max=-10.0))


def fill_holes(mask, max_area):
    """fill_holes_in_mask_scores (sam2/utils/misc.py:312-338) with the CC restatement of
    oracle/cc_ref.py in place of the CUDA-only sam2._C op."""
    from oracle.cc_ref import connected_components_ref

    labels, areas = connected_components_ref((mask <= 0).to(torch.uint8).numpy())
    hole = torch.from_numpy((labels > 0) & (areas <= max_area))
    return torch.where(hole, torch.full_like(mask, 0.1), mask)


# --------------------------------------------------------------------------------------
# the predictor (session state + frame loop)
# --------------------------------------------------------------------------------------
class RefPredictor:
    """Restatement of SAM2VideoPredictor[NPZ] (sam2_video_predictor.py:18-1172): same public
    methods, same `inference_state` keys, CPU fp32.  `fill_holes=False` reproduces the
    reference's CPU behaviour (the CUDA-only CC op raises and the step is skipped,
    misc.py:321-336); `True` reproduces its GPU behaviour."""

    def __init__(self, state_dict, cfg=Cfg, fill_holes=True, model_cls=None):
        self.model = (model_cls or RefModel)(state_dict, cfg)
        self.cfg = cfg
        self.image_size = cfg.image_size
        self.fill_holes = fill_holes

    # -- state --
    def init_state(self, images, video_height, video_width):
        st = dict(images=images, num_frames=len(images), video_height=video_height, video_width=video_width,
                  point_inputs_per_obj={}, mask_inputs_per_obj={}, cached_features={}, constants={},
                  obj_id_to_idx=OrderedDict(), obj_idx_to_id=OrderedDict(), obj_ids=[],
                  output_dict=dict(cond_frame_outputs={}, non_cond_frame_outputs={}),
                  output_dict_per_obj={}, temp_output_dict_per_obj={},
                  consolidated_frame_inds=dict(cond_frame_outputs=set(), non_cond_frame_outputs=set()),
                  tracking_has_started=False, frames_already_tracked={})
        self._features(st, 0)
        return st

    def _obj_idx(self, st, obj_id):
        idx = st["obj_id_to_idx"].get(obj_id)
        if idx is not None:
            return idx
        if st["tracking_has_started"]:
            raise RuntimeError(f"Cannot add new object id {obj_id} after tracking starts.")
        idx = len(st["obj_id_to_idx"])
        st["obj_id_to_idx"][obj_id] = idx
        st["obj_idx_to_id"][idx] = obj_id
        st["obj_ids"] = list(st["obj_id_to_idx"])
        for key in ("point_inputs_per_obj", "mask_inputs_per_obj"):
            st[key][idx] = {}
        for key in ("output_dict_per_obj", "temp_output_dict_per_obj"):
            st[key][idx] = dict(cond_frame_outputs={}, non_cond_frame_outputs={})
        return idx

    def _features(self, st, t):
        hit = st["cached_features"].get(t)
        if hit is None:
            hit = self.model.forward_image(st["images"][t].float()[None])
            st["cached_features"] = {t: hit}
        return hit

    # -- one frame --
    def _single_frame(self, st, output_dict, t, B, is_init, point_inputs, mask_inputs, reverse, run_mem,
                      prev_logits=None):
        """_run_single_frame_inference (sam2_video_predictor.py:912-978)."""
        out = self.model.track_step(t, is_init, self._features(st, t), B, point_inputs, mask_inputs,
                                    output_dict, st["num_frames"], reverse, run_mem, prev_logits)
        mf = out["maskmem_features"]
        if mf is not None:
            mf = mf.to(torch.bfloat16)
        pm = out["pred_masks"]
        if self.fill_holes and self.cfg.fill_hole_area > 0:
            pm = fill_holes(pm, self.cfg.fill_hole_area)
        compact = dict(maskmem_features=mf, maskmem_pos_enc=out["maskmem_pos_enc"], pred_masks=pm,
                       obj_ptr=out["obj_ptr"], object_score_logits=out["object_score_logits"],
                       ious=out["ious"])
        return compact, pm

    def _video_res(self, st, masks):
        """_get_orig_video_res_output (sam2_video_predictor.py:404-424)."""
        hw = (st["video_height"], st["video_width"])
        if tuple(masks.shape[-2:]) != hw:
            masks = F.interpolate(masks, size=hw, mode="bilinear", align_corners=False)
        return non_overlapping(masks) if self.cfg.non_overlap_masks else masks

    def _prompt_common(self, st, t, obj_id):
        idx = self._obj_idx(st, obj_id)
        is_init = t not in st["frames_already_tracked"]
        reverse = False if is_init else st["frames_already_tracked"][t]["reverse"]
        key = "cond_frame_outputs" if is_init else "non_cond_frame_outputs"
        return idx, is_init, reverse, key

    def add_new_mask(self, st, frame_idx, obj_id, mask):
        """sam2_video_predictor.py:321-402."""
        idx, is_init, reverse, key = self._prompt_common(st, frame_idx, obj_id)
        m = torch.as_tensor(mask).to(torch.bool)[None, None].float()
        if tuple(m.shape[-2:]) != (self.image_size, self.image_size):
            m = F.interpolate(m, size=(self.image_size, self.image_size), mode="bilinear",
                              align_corners=False, antialias=True)
            m = (m >= 0.5).float()
        st["mask_inputs_per_obj"][idx][frame_idx] = m
        st["point_inputs_per_obj"][idx].pop(frame_idx, None)
        out, _ = self._single_frame(st, st["output_dict_per_obj"][idx], frame_idx, 1, is_init, None, m,
                                    reverse, False)
        st["temp_output_dict_per_obj"][idx][key][frame_idx] = out
        cons = self._consolidate(st, frame_idx, key == "cond_frame_outputs", False, True)
        return frame_idx, st["obj_ids"], self._video_res(st, cons["pred_masks_video_res"])

    def add_new_points_or_box(self, st, frame_idx, obj_id, points=None, labels=None, clear_old_points=True,
                              normalize_coords=True, box=None):
        """sam2_video_predictor.py:173-314."""
        idx, is_init, reverse, key = self._prompt_common(st, frame_idx, obj_id)
        if (points is not None) != (labels is not None):
            raise ValueError("points and labels must be provided together")
        if points is None and box is None:
            raise ValueError("at least one of points or box must be provided as input")
        pts = torch.zeros(0, 2) if points is None else torch.as_tensor(points, dtype=torch.float32)
        lab = torch.zeros(0, dtype=torch.int32) if labels is None else torch.as_tensor(labels, dtype=torch.int32)
        pts, lab = pts.reshape(1, -1, 2), lab.reshape(1, -1)
        if box is not None:
            if not clear_old_points:
                raise ValueError("cannot add box without clearing old points")
            pts = torch.cat([torch.as_tensor(box, dtype=torch.float32).reshape(1, 2, 2), pts], dim=1)
            lab = torch.cat([torch.tensor([[2, 3]], dtype=torch.int32), lab], dim=1)
        if normalize_coords:
            pts = pts / torch.tensor([st["video_width"], st["video_height"]], dtype=torch.float32)
        pts = pts * self.image_size
        old = None if clear_old_points else st["point_inputs_per_obj"][idx].get(frame_idx)
        if old is not None:
            pts = torch.cat([old["point_coords"], pts], dim=1)
            lab = torch.cat([old["point_labels"], lab], dim=1)
        pin = dict(point_coords=pts, point_labels=lab)
        st["point_inputs_per_obj"][idx][frame_idx] = pin
        st["mask_inputs_per_obj"][idx].pop(frame_idx, None)
        od, tmp = st["output_dict_per_obj"][idx], st["temp_output_dict_per_obj"][idx]
        prev = tmp[key].get(frame_idx) or od["cond_frame_outputs"].get(frame_idx) \
            or od["non_cond_frame_outputs"].get(frame_idx)
        prev_logits = None
        if prev is not None and prev["pred_masks"] is not None:
            prev_logits = prev["pred_masks"].clamp(-32.0, 32.0)
        out, _ = self._single_frame(st, od, frame_idx, 1, is_init, pin, None, reverse, False, prev_logits)
        tmp[key][frame_idx] = out
        cons = self._consolidate(st, frame_idx, key == "cond_frame_outputs", False, True)
        return frame_idx, st["obj_ids"], self._video_res(st, cons["pred_masks_video_res"])

    def _consolidate(self, st, t, is_cond, run_mem, at_video_res=False):
        """_consolidate_temp_output_across_obj (sam2_video_predictor.py:426-554)."""
        B = len(st["obj_idx_to_id"])
        key = "cond_frame_outputs" if is_cond else "non_cond_frame_outputs"
        if at_video_res:
            hw, mkey = (st["video_height"], st["video_width"]), "pred_masks_video_res"
        else:
            hw, mkey = (self.image_size // 4,) * 2, "pred_masks"
        cons = {"maskmem_features": None, "maskmem_pos_enc": None,
                mkey: torch.full((B, 1, *hw), NO_OBJ_SCORE),
                "obj_ptr": torch.full((B, self.cfg.d_model), NO_OBJ_SCORE),
                "object_score_logits": torch.full((B, 1), 10.0)}
        empty_ptr = None
        for i in range(B):
            out = st["temp_output_dict_per_obj"][i][key].get(t)
            if out is None:
                out = st["output_dict_per_obj"][i]["cond_frame_outputs"].get(t)
            if out is None:
                out = st["output_dict_per_obj"][i]["non_cond_frame_outputs"].get(t)
            if out is None:
                if run_mem:
                    if empty_ptr is None:
                        z = torch.zeros(1, 1, self.image_size, self.image_size)
                        empty_ptr = self.model.track_step(t, True, self._features(st, t), 1, None, z, {},
                                                          st["num_frames"], False, False)["obj_ptr"]
                    cons["obj_ptr"][i:i + 1] = empty_ptr
                continue
            m = out["pred_masks"]
            if tuple(m.shape[-2:]) != hw:
                m = F.interpolate(m, size=hw, mode="bilinear", align_corners=False)
            cons[mkey][i:i + 1] = m
            cons["obj_ptr"][i:i + 1] = out["obj_ptr"]
            cons["object_score_logits"][i:i + 1] = out["object_score_logits"]
        if run_mem:
            high = F.interpolate(cons["pred_masks"], size=(self.image_size,) * 2, mode="bilinear",
                                 align_corners=False)
            feats = self._features(st, t)
            mf = self.model.encode_memory(feats["feat"].expand(B, -1, -1, -1), high,
                                          cons["object_score_logits"], True)
            cons["maskmem_features"] = mf.to(torch.bfloat16)
            cons["maskmem_pos_enc"] = [self.model.sine_pos(self.model.fs, self.model.fs, 64)[None].expand(B, -1, -1, -1)]
        return cons

    def _per_object(self, st, t, out, key):
        for i, od in st["output_dict_per_obj"].items():
            s = slice(i, i + 1)
            o = dict(maskmem_features=None, maskmem_pos_enc=None, pred_masks=out["pred_masks"][s],
                     obj_ptr=out["obj_ptr"][s], object_score_logits=out["object_score_logits"][s])
            if out["maskmem_features"] is not None:
                o["maskmem_features"] = out["maskmem_features"][s]
                o["maskmem_pos_enc"] = [x[s] for x in out["maskmem_pos_enc"]]
            od[key][t] = o

    def _preflight(self, st):
        """propagate_in_video_preflight (sam2_video_predictor.py:593-660)."""
        st["tracking_has_started"] = True
        od, cfi = st["output_dict"], st["consolidated_frame_inds"]
        for is_cond in (False, True):
            key = "cond_frame_outputs" if is_cond else "non_cond_frame_outputs"
            frames = set()
            for tmp in st["temp_output_dict_per_obj"].values():
                frames.update(tmp[key].keys())
            cfi[key].update(frames)
            for t in frames:
                cons = self._consolidate(st, t, is_cond, True)
                od[key][t] = cons
                self._per_object(st, t, cons, key)
            for tmp in st["temp_output_dict_per_obj"].values():
                tmp[key].clear()
        for t in od["cond_frame_outputs"]:
            od["non_cond_frame_outputs"].pop(t, None)
        for o in st["output_dict_per_obj"].values():
            for t in o["cond_frame_outputs"]:
                o["non_cond_frame_outputs"].pop(t, None)
        for t in cfi["cond_frame_outputs"]:
            cfi["non_cond_frame_outputs"].discard(t)

    def propagate_in_video(self, st, start_frame_idx=None, max_frame_num_to_track=None, reverse=False):
        """sam2_video_predictor.py:663-745 (generator)."""
        self._preflight(st)
        od, cfi = st["output_dict"], st["consolidated_frame_inds"]
        T, B = st["num_frames"], len(st["obj_idx_to_id"])
        if not od["cond_frame_outputs"]:
            raise RuntimeError("No points are provided; please add points first")
        if start_frame_idx is None:
            start_frame_idx = min(od["cond_frame_outputs"])
        if max_frame_num_to_track is None:
            max_frame_num_to_track = T
        if reverse:
            end = max(start_frame_idx - max_frame_num_to_track, 0)
            order = range(start_frame_idx, end - 1, -1) if start_frame_idx > 0 else []
        else:
            end = min(start_frame_idx + max_frame_num_to_track, T - 1)
            order = range(start_frame_idx, end + 1)
        for t in order:
            if t in cfi["cond_frame_outputs"]:
                key = "cond_frame_outputs"
                out = od[key][t]
                pm = out["pred_masks"]
            elif t in cfi["non_cond_frame_outputs"]:
                key = "non_cond_frame_outputs"
                out = od[key][t]
                pm = out["pred_masks"]
            else:
                key = "non_cond_frame_outputs"
                out, pm = self._single_frame(st, od, t, B, False, None, None, reverse, True)
                od[key][t] = out
            self._per_object(st, t, out, key)
            st["frames_already_tracked"][t] = {"reverse": reverse}
            yield t, st["obj_ids"], self._video_res(st, pm)

    def _reset_tracking_results(self, st):
        """sam2_video_predictor.py:860-876: every prompt and tracking result goes, the object ids stay."""
        for k in ("point_inputs_per_obj", "mask_inputs_per_obj"):
            for v in st[k].values():
                v.clear()
        for k in ("output_dict_per_obj", "temp_output_dict_per_obj"):
            for v in st[k].values():
                v["cond_frame_outputs"].clear()
                v["non_cond_frame_outputs"].clear()
        for k in ("cond_frame_outputs", "non_cond_frame_outputs"):
            st["output_dict"][k].clear()
            st["consolidated_frame_inds"][k].clear()
        st["tracking_has_started"] = False
        st["frames_already_tracked"].clear()

    def clear_all_prompts_in_frame(self, st, frame_idx, obj_id, need_output=True):
        """sam2_video_predictor.py:777-846."""
        idx = self._obj_idx(st, obj_id)
        st["point_inputs_per_obj"][idx].pop(frame_idx, None)
        st["mask_inputs_per_obj"][idx].pop(frame_idx, None)
        tmp = st["temp_output_dict_per_obj"]
        tmp[idx]["cond_frame_outputs"].pop(frame_idx, None)
        tmp[idx]["non_cond_frame_outputs"].pop(frame_idx, None)
        B = len(st["obj_idx_to_id"])
        has_input = any(frame_idx in st["point_inputs_per_obj"][i] or frame_idx in st["mask_inputs_per_obj"][i]
                        for i in range(B))
        if not has_input:  # the frame stops being a conditioning frame: its output is downgraded
            od, cfi = st["output_dict"], st["consolidated_frame_inds"]
            cfi["cond_frame_outputs"].discard(frame_idx)
            cfi["non_cond_frame_outputs"].discard(frame_idx)
            out = od["cond_frame_outputs"].pop(frame_idx, None)
            if out is not None:
                od["non_cond_frame_outputs"][frame_idx] = out
                st["frames_already_tracked"].pop(frame_idx, None)
            for i in range(B):
                pod = st["output_dict_per_obj"][i]
                o = pod["cond_frame_outputs"].pop(frame_idx, None)
                if o is not None:
                    pod["non_cond_frame_outputs"][frame_idx] = o
            if len(od["cond_frame_outputs"]) == 0:
                self._reset_tracking_results(st)
        if not need_output:
            return None
        is_cond = any(frame_idx in t["cond_frame_outputs"] for t in tmp.values())
        cons = self._consolidate(st, frame_idx, is_cond, False, True)
        return frame_idx, st["obj_ids"], self._video_res(st, cons["pred_masks_video_res"])

    def remove_object(self, st, obj_id, strict=False, need_output=True):
        """sam2_video_predictor.py:1042-1152."""
        rm = st["obj_id_to_idx"].get(obj_id)
        updated = []
        if rm is None:
            if not strict:
                return st["obj_ids"], updated
            raise RuntimeError(f"Cannot remove object id {obj_id} as it doesn't exist.")
        if len(st["obj_id_to_idx"]) == 1:
            self.reset_state(st)
            return st["obj_ids"], updated
        input_frames = set(st["point_inputs_per_obj"][rm]) | set(st["mask_inputs_per_obj"][rm])
        for t in input_frames:
            self.clear_all_prompts_in_frame(st, t, obj_id, need_output=False)
        old_ids = st["obj_ids"]
        old_inds = list(range(len(old_ids)))
        keep = [i for i in old_inds if i != rm]
        new_ids = [old_ids[i] for i in keep]
        remap = {o: n for n, o in enumerate(keep)}
        st["obj_id_to_idx"] = OrderedDict((oid, n) for n, oid in enumerate(new_ids))
        st["obj_idx_to_id"] = OrderedDict((n, oid) for n, oid in enumerate(new_ids))
        st["obj_ids"] = new_ids
        for k in ("point_inputs_per_obj", "mask_inputs_per_obj", "output_dict_per_obj", "temp_output_dict_per_obj"):
            c = st[k]
            vals = {i: c.pop(i) for i in old_inds}
            c.update({remap[i]: v for i, v in vals.items() if i in remap})
        for key in ("cond_frame_outputs", "non_cond_frame_outputs"):
            for t, out in st["output_dict"][key].items():
                out["maskmem_features"] = out["maskmem_features"][keep]
                out["maskmem_pos_enc"] = [x[keep] for x in out["maskmem_pos_enc"]]
                out["pred_masks"] = out["pred_masks"][keep]
                out["obj_ptr"] = out["obj_ptr"][keep]
                out["object_score_logits"] = out["object_score_logits"][keep]
                self._per_object(st, t, out, key)
        if need_output:
            tmp = st["temp_output_dict_per_obj"]
            for t in input_frames:
                is_cond = any(t in d["cond_frame_outputs"] for d in tmp.values())
                cons = self._consolidate(st, t, is_cond, False, True)
                updated.append((t, self._video_res(st, cons["pred_masks_video_res"])))
        return st["obj_ids"], updated

    def reset_state(self, st):
        """sam2_video_predictor.py:848-876."""
        for k in ("obj_id_to_idx", "obj_idx_to_id", "point_inputs_per_obj", "mask_inputs_per_obj",
                  "output_dict_per_obj", "temp_output_dict_per_obj", "frames_already_tracked"):
            st[k].clear()
        st["obj_ids"] = []
        for k in ("cond_frame_outputs", "non_cond_frame_outputs"):
            st["output_dict"][k].clear()
            st["consolidated_frame_inds"][k].clear()
        st["tracking_has_started"] = False


# --------------------------------------------------------------------------------------
# image predictor (sam2/sam2_image_predictor.py) -- checker for us_video_medsam2_b200/image_predictor.py
# --------------------------------------------------------------------------------------
class RefImagePredictor:
    """Restatement of SAM2ImagePredictor.set_image / predict (sam2_image_predictor.py:86-131, 238-430) and of
    SAM2Transforms.postprocess_masks (sam2/utils/transforms.py:76-115) with the CC restatement in place of sam2._C."""

    def __init__(self, state_dict, cfg=Cfg, mask_threshold=0.0, max_hole_area=0.0, max_sprinkle_area=0.0):
        self.model = RefModel(state_dict, cfg)
        self.cfg = cfg
        self.mask_threshold, self.max_hole_area, self.max_sprinkle_area = mask_threshold, max_hole_area, max_sprinkle_area
        self._feats = None

    def set_image(self, image):
        from torchvision.transforms import Normalize, Resize, ToTensor  # the reference's own pre-processing ops

        self._orig_hw = tuple(image.shape[:2])
        s = self.cfg.image_size
        x = Normalize([0.485, 0.456, 0.406], [0.229, 0.224, 0.225])(Resize((s, s))(ToTensor()(image)))[None]
        f = self.model.forward_image(x)
        f["feat"] = f["feat"] + self.model.p("no_mem_embed").reshape(1, -1, 1, 1)  # directly_add_no_mem_embed (:117-121)
        self._feats = f

    def _post(self, masks):
        from oracle.cc_ref import connected_components_ref

        thr = self.mask_threshold
        flat = masks.flatten(0, 1).unsqueeze(1)
        if self.max_hole_area > 0:
            lab, area = connected_components_ref((flat <= thr).to(torch.uint8).numpy())
            hole = torch.from_numpy((lab > 0) & (area <= self.max_hole_area)).reshape_as(masks)
            masks = torch.where(hole, torch.full_like(masks, thr + 10.0), masks)
        if self.max_sprinkle_area > 0:
            lab, area = connected_components_ref((flat > thr).to(torch.uint8).numpy())
            spr = torch.from_numpy((lab > 0) & (area <= self.max_sprinkle_area)).reshape_as(masks)
            masks = torch.where(spr, torch.full_like(masks, thr - 10.0), masks)
        return F.interpolate(masks, self._orig_hw, mode="bilinear", align_corners=False)

    def predict(self, point_coords=None, point_labels=None, box=None, mask_input=None, multimask_output=True,
                return_logits=False, normalize_coords=True):
        if self._feats is None:
            raise RuntimeError("An image must be set with .set_image(...) before mask prediction.")
        m, s = self.model, float(self.cfg.image_size)
        h, w = self._orig_hw
        scale = torch.tensor([s / w, s / h]) if normalize_coords else torch.tensor([s, s])
        coords = labels = None
        if point_coords is not None:
            coords = torch.as_tensor(point_coords, dtype=torch.float32).reshape(1, -1, 2) * scale
            labels = torch.as_tensor(point_labels, dtype=torch.int32).reshape(1, -1)
        if box is not None:  # corners as points labelled 2 / 3, placed first (:392-404)
            bc = torch.as_tensor(box, dtype=torch.float32).reshape(1, 2, 2) * scale
            bl = torch.tensor([[2, 3]], dtype=torch.int32)
            coords = bc if coords is None else torch.cat([bc, coords], dim=1)
            labels = bl if labels is None else torch.cat([bl, labels], dim=1)
        f = self._feats
        sparse = m.embed_points(coords, labels) if coords is not None else torch.zeros(1, 0, 256)
        if mask_input is not None:
            dense = m.embed_mask(torch.as_tensor(mask_input, dtype=torch.float32).reshape(1, 1, int(s) // 4, int(s) // 4))
        else:
            dense = m.p("sam_prompt_encoder.no_mask_embed.weight").reshape(1, -1, 1, 1).expand(1, -1, m.fs, m.fs)
        low, iou, _, _ = m.mask_decoder(f["feat"], sparse, dense, f["feat_s0"], f["feat_s1"], multimask_output)
        masks = self._post(low.float())
        low = low.clamp(-32.0, 32.0)
        if not return_logits:
            masks = masks > self.mask_threshold
        return masks[0].float().numpy() if return_logits else masks[0].numpy(), iou[0].numpy(), low[0].numpy()
