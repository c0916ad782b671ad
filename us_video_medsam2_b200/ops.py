"""Thin torch-tensor wrappers over the C ABI (include/usvm2_b200.h).

PyTorch is used for device memory and streams only: every function here takes CUDA tensors, passes
raw pointers to a hand-written kernel on torch's current stream and returns tensors.  There is no
fallback path: a non-CUDA tensor or a missing library raises.
"""
import ctypes as C
import math
import os

import torch

from . import _lib
from ._lib import (ACT_GELU, ACT_NONE, ACT_RELU, HieraAttnParams, POST_BINARIZE_AFFINE, POST_NONE, POST_SIGMOID_AFFINE,  # noqa: F401
                   FmhaParams, FrameCtrl, GemmEpilogue, MemoryFrames, SkinnyParams, call)

BF16 = torch.bfloat16
F32 = torch.float32

# bring-up switch: route the bf16 GEMMs through the SIMT kernel (still CUDA, still this library)
_FORCE_SIMT = os.environ.get("USVM2_GEMM", "") == "simt"


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream():
    """cudaStream_t of torch's current stream on the current device (raw handle: the Python Stream wrapper costs
    ~15 us per call, more than most of the kernels launched here)."""
    if _raw_stream is not None:
        return _raw_stream(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _chk(t, dtype, name):
    if not (t.is_cuda and t.dtype == dtype):
        raise TypeError(f"{name}: expected a CUDA {dtype} tensor, got {t.device} {t.dtype}")


def set_sm_budget(sms):
    """SMs that persistent kernels launched (or captured) from now on may occupy; 0 = the whole device."""
    _lib.lib().usvm_set_sm_budget(int(sms))


def empty(shape, dtype, like):
    return torch.empty(shape, dtype=dtype, device=like.device)


# ------------------------------------------------------------------------------------------------
# GEMM
# ------------------------------------------------------------------------------------------------
def _epilogue(M, N, bias, act, col_scale, residual, res_mod, want_f32, want_bf16, like, out_f32, out_bf16, rope=None,
              ln=None, res_div=0):
    ep = GemmEpilogue()
    ep.res_div = res_div
    if ln is not None:  # (weight, bias, eps[, gelu])
        ep.ln_w, ep.ln_b, ep.ln_eps = ln[0].data_ptr(), ln[1].data_ptr(), ln[2]
        ep.ln_gelu = int(len(ln) > 3 and bool(ln[3]))
    if rope is not None:  # (cos, sin, cols, rows_per_batch, n_rope); tables in the tiled order of tile_rope_table()
        if rope[0].dim() != 4 or rope[1].dim() != 4:
            raise ValueError("fused RoPE takes tables laid out by ops.tile_rope_table()")
        ep.rope_cos, ep.rope_sin = rope[0].data_ptr(), rope[1].data_ptr()
        ep.rope_cols, ep.rope_rows_per_batch, ep.rope_n_rope = rope[2], rope[3], rope[4]
        ep.rope_table_rows = rope[0].numel() // 128
    ep.bias, ep.col_scale = _ptr(bias), _ptr(col_scale)
    ep.residual = _ptr(residual)
    ep.ldr = residual.stride(0) if residual is not None else 0
    ep.res_mod, ep.act = res_mod, act
    if want_f32 and out_f32 is None:
        out_f32 = empty((M, N), F32, like)
    if want_bf16 and out_bf16 is None:
        out_bf16 = empty((M, N), BF16, like)
    ep.out_f32, ep.ldo_f32 = _ptr(out_f32), (out_f32.stride(0) if out_f32 is not None else 0)
    ep.out_bf16, ep.ldo_bf16 = _ptr(out_bf16), (out_bf16.stride(0) if out_bf16 is not None else 0)
    return ep, out_f32, out_bf16


def tile_rope_table(t):
    """Row-major rotary table fp32 [R, 128] (R % 32 == 0) -> the tiled order the GEMM epilogue reads,
    [R / 32][128 / 4][32][4]: element (pos, tc) at ((pos // 32) * 32 + tc // 4) * 128 + (pos % 32) * 4 + tc % 4."""
    R, Cc = t.shape
    assert Cc == 128 and R % 32 == 0
    return t.view(R // 32, 32, 32, 4).permute(0, 2, 1, 3).contiguous()


def gemm_bf16(a, w, bias=None, act=ACT_NONE, col_scale=None, residual=None, res_mod=0, f32=False, bf16=False,
              out_f32=None, out_bf16=None, block_n=0, simt=None, rope=None, ln=None, ln_fused=None, res_div=0):
    """epi(a[M,K] @ w[N,K]^T) on the tcgen05 kernel; returns (fp32 out or None, bf16 out or None).
    rope = (cos, sin, cols, rows_per_batch, n_rope): fused rotary encoding of output columns [0, cols).
    ln = (weight, bias, eps[, gelu]): the bf16 output becomes LayerNorm(row) (then GELU) of the result.  With N == 256
    the norm can run inside the GEMM epilogue (one 128 x 256 tile per CTA owns whole rows, ln_fused=True).  Measured with
    the current kernels (tools/bench_lnfuse.py): a fused tile takes ~20 us whatever M is (two passes over a 128 x 256
    accumulator by four epilogue warps), GEMM + LayerNorm launches 7.7 us at M = 4096, 18 us at 16384, 35 us against 50 at
    32768 -- the two launches win everywhere, so ln_fused=None means not fused."""
    _chk(a, BF16, "a"), _chk(w, BF16, "w")
    M, K = a.shape
    N = w.shape[0]
    assert w.shape[1] == K and a.stride(1) == 1 and w.stride(1) == 1
    if ln is not None and ((_FORCE_SIMT if simt is None else simt) or N != 256 or not ln_fused):  # GEMM, then LayerNorm
        o32, _ = gemm_bf16(a, w, bias, act, col_scale, residual, res_mod, f32=True, out_f32=out_f32, block_n=block_n,
                           simt=simt, rope=rope, res_div=res_div)
        _, o16 = layernorm(o32, ln[0], ln[1], ln[2], bf16=True, gelu=len(ln) > 3 and bool(ln[3]))
        return o32, o16
    ep, o32, o16 = _epilogue(M, N, bias, act, col_scale, residual, res_mod, f32 or out_f32 is not None,
                             bf16 or out_bf16 is not None or ln is not None, a, out_f32, out_bf16, rope, ln, res_div)
    if _FORCE_SIMT if simt is None else simt:
        call("usvm_gemm_simt", a.data_ptr(), 1, a.stride(0), w.data_ptr(), 1, w.stride(0), C.byref(ep), M, N, K,
             _stream())
    else:
        call("usvm_gemm_bf16_tc5", a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0), C.byref(ep), M, N, K,
             block_n, _stream())
    return o32, o16


def gemm_f32(a, w, bias=None, act=ACT_NONE, col_scale=None, residual=None, res_mod=0, out=None, tf32=False,
             block_n=0):
    """fp32-operand GEMM (decoder tail): epi(a[M,K] @ w[N,K]^T) -> fp32 [M,N].  tf32=True runs the tcgen05 kernel with
    tf32 products (image-side projections, >= 128 rows); otherwise the SIMT kernel with exact fp32 products."""
    _chk(a, F32, "a"), _chk(w, F32, "w")
    M, K = a.shape
    N = w.shape[0]
    assert w.shape[1] == K and a.stride(1) == 1 and w.stride(1) == 1
    ep, o32, _ = _epilogue(M, N, bias, act, col_scale, residual, res_mod, True, False, a, out, None)
    if tf32 and not _FORCE_SIMT:
        call("usvm_gemm_tf32_tc5", a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0), C.byref(ep), M, N, K, block_n,
             _stream())
        return o32
    call("usvm_gemm_simt", a.data_ptr(), 0, a.stride(0), w.data_ptr(), 0, w.stride(0), C.byref(ep), M, N, K,
         _stream())
    return o32


# ------------------------------------------------------------------------------------------------
# attention
# ------------------------------------------------------------------------------------------------
_FMHA_PART_BF16 = os.environ.get("USVM2_FMHA_PART_BF16", "1") != "0"
_FMHA_IMPL = os.environ.get("USVM2_FMHA", "tc5")  # "tc5" (tcgen05 where the shape allows) | "mma" (mma.sync only)


def ffn_fused(h, x, w1, b1, w2, b2):
    """x + relu(h @ w1^T + b1) @ w2^T + b2 as one cluster kernel (memory-attention feed-forward block):
    h bf16 [M,256] = LayerNorm(x), x fp32 [M,256]; returns a new fp32 [M,256]."""
    _chk(h, BF16, "h"), _chk(x, F32, "x"), _chk(w1, BF16, "w1"), _chk(w2, BF16, "w2")
    M, D = x.shape
    Hd = w1.shape[0]
    out = empty((M, D), F32, x)
    call("usvm_ffn_fused_tc5", h.data_ptr(), x.data_ptr(), w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(),
         out.data_ptr(), M, D, Hd, _stream())
    return out


def fmha(q, k, v, B, H, Nq, Nk, head_dim, q_addr, k_addr, v_addr, out=None, num_splits=1, impl=None):
    """q/k/v: bf16 tensors; *_addr = (element offset, batch stride, row stride, head stride).
    Returns bf16 [B, Nq, H*head_dim].  head_dim 256 / one head / Nq % 128 == 0 / contiguous batches run on the
    tcgen05 kernel (usvm_fmha_tc5), everything else on the mma.sync kernel (usvm_fmha_bf16)."""
    for t in (q, k, v):
        _chk(t, BF16, "qkv")
    if out is None:
        out = empty((B, Nq, H * head_dim), BF16, q)
    p = FmhaParams()
    p.q, p.k, p.v = q.data_ptr() + 2 * q_addr[0], k.data_ptr() + 2 * k_addr[0], v.data_ptr() + 2 * v_addr[0]
    p.o = out.data_ptr()
    p.q_bs, p.q_rs, p.q_hs = q_addr[1:]
    p.k_bs, p.k_rs, p.k_hs = k_addr[1:]
    p.v_bs, p.v_rs, p.v_hs = v_addr[1:]
    p.o_bs, p.o_rs, p.o_hs = Nq * H * head_dim, H * head_dim, head_dim
    p.B, p.H, p.Nq, p.Nk, p.head_dim = B, H, Nq, Nk, head_dim
    num_splits = max(1, min(num_splits, (Nk + 63) // 64))
    p.num_splits = num_splits
    p.scale = 1.0 / math.sqrt(head_dim)
    which = impl or _FMHA_IMPL
    use_tc5 = (which == "tc5" and head_dim == 256 and H == 1 and Nq % 128 == 0
               and q_addr[1] == Nq * q_addr[2] and k_addr[1] == Nk * k_addr[2] and v_addr[1] == Nk * v_addr[2])
    if num_splits > 1:
        p.part_bf16 = int(use_tc5 and _FMHA_PART_BF16)  # split partials in bf16 (tcgen05 kernel only)
        o_part = empty((num_splits, B * H, Nq, head_dim), BF16 if p.part_bf16 else F32, q)
        ml_part = empty((num_splits, B * H, Nq, 2), F32, q)
        p.o_part, p.ml_part = o_part.data_ptr(), ml_part.data_ptr()
    if use_tc5:
        call("usvm_fmha_tc5", C.byref(p), _stream())
        if num_splits > 1:
            call("usvm_fmha_combine", C.byref(p), _stream())
    else:
        call("usvm_fmha_bf16", C.byref(p), _stream())
    return out


def hiera_attn(qkv, qkv_bias, Fr, H, W, dim, heads, window=0, pool=False):
    """Image-encoder attention on the tcgen05 kernel, straight from the projection's raster-order output:
    qkv bf16 [Fr*H*W, 3*dim] -> bf16 [Fr*H*W, dim].  window 0 = global, 14 / 7 = the windows of Hiera stages 3 / 4 and of
    the ViT trunk (padding tokens enter in closed form through qkv_bias); pool: 2 x 2 max-pooled queries (window 14),
    output on the pooled grid [Fr*(H/2)*(W/2), dim]."""
    _chk(qkv, BF16, "qkv")
    _chk(qkv_bias, F32, "qkv_bias")
    if qkv.shape != (Fr * H * W, 3 * dim):
        raise RuntimeError(f"hiera_attn: qkv {tuple(qkv.shape)} != {(Fr * H * W, 3 * dim)}")
    out = empty((Fr * (H // 2) * (W // 2) if pool else Fr * H * W, dim), BF16, qkv)
    p = HieraAttnParams()
    p.qkv, p.out, p.qkv_bias = qkv.data_ptr(), out.data_ptr(), qkv_bias.data_ptr()
    p.F, p.H, p.W, p.dim, p.heads, p.window, p.pool = Fr, H, W, dim, heads, window, int(bool(pool))
    p.scale = 1.0 / math.sqrt(dim // heads)
    call("usvm_hiera_attn_tc5", C.byref(p), _stream())
    return out


def attn_small(q, k, v, B, H, Nq, Nk, head_dim):
    """fp32 [B*Nq, H*dh] x [B*Nk, H*dh] -> fp32 [B*Nq, H*dh] (row strides taken from the tensors)."""
    out = empty((B * Nq, H * head_dim), F32, q)
    call("usvm_attn_small_f32", q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), B, H, Nq, Nk, head_dim,
         q.stride(0), k.stride(0), v.stride(0), out.stride(0), 1.0 / math.sqrt(head_dim), _stream())
    return out


# ------------------------------------------------------------------------------------------------
# normalisation / layout
# ------------------------------------------------------------------------------------------------
def layernorm(x, w, b, eps, f32=False, bf16=False, gelu=False, valid=None):
    """valid < x.shape[1]: the rows carry `valid` real channels followed by zero padding (a channel count that is not a
    multiple of 32, see engine._cpad); the norm runs over the real ones and the padding of the outputs is zero."""
    _chk(x, F32, "x")
    rows, Cc = x.shape
    Cv = Cc if valid is None else valid
    alloc = empty if Cv == Cc else (lambda shape, dt, like: torch.zeros(shape, dtype=dt, device=like.device))
    o32 = alloc((rows, Cc), F32, x) if f32 else None
    o16 = alloc((rows, Cc), BF16, x) if bf16 else None
    call("usvm_layernorm", x.data_ptr(), x.stride(0), w.data_ptr(), b.data_ptr(), eps, int(gelu), _ptr(o32), Cc,
         _ptr(o16), Cc, rows, Cv, _stream())
    return o32, o16


def axpby(x, y, alpha=1.0, beta=1.0, rows=None, x_mod=0, y_mod=0, f32=True, bf16=False, x_div=0):
    """alpha * x[r % x_mod] + beta * y[r % y_mod] over `rows` rows of C channels; with x_div > 0 the x row is
    (r // x_div) * x_mod + r % x_mod (groups of x_div output rows share one x_mod-row block of x)."""
    Cc = x.shape[-1]
    rows = rows if rows is not None else x.numel() // Cc
    o32 = empty((rows, Cc), F32, x) if f32 else None
    o16 = empty((rows, Cc), BF16, x) if bf16 else None
    call("usvm_axpby_rows", x.data_ptr(), _ptr(y), alpha, beta, x_mod, y_mod, x_div, _ptr(o32), _ptr(o16), rows, Cc,
         _stream())
    return o32, o16


def cast_bf16(x):
    out = empty(x.shape, BF16, x)
    call("usvm_cast_f32_bf16", x.data_ptr(), out.data_ptr(), x.numel(), _stream())
    return out


def rope(x, col0, cos_t, sin_t, rows_per_batch, n_rope, dim=256):
    """Rotate columns [col0, col0+dim) of fp32 x [rows, ld] -> bf16 [rows, dim]."""
    rows = x.shape[0]
    out = empty((rows, dim), BF16, x)
    call("usvm_rope_bf16", x.data_ptr() + 4 * col0, x.stride(0), cos_t.data_ptr(), sin_t.data_ptr(), out.data_ptr(),
         dim, rows, rows_per_batch, n_rope, cos_t.shape[0], dim, _stream())
    return out


def window_gather(qkv, qkv_bias, F, Hg, Wg, ws, pool, Cc):
    nw = -(-Hg // ws) * -(-Wg // ws)
    nk = ws * ws
    nq = (ws // 2) ** 2 if pool else nk
    Qw, Kw, Vw = (empty((F * nw, n, Cc), BF16, qkv) for n in (nq, nk, nk))
    call("usvm_window_gather", qkv.data_ptr(), qkv_bias.data_ptr(), Qw.data_ptr(), Kw.data_ptr(), Vw.data_ptr(), F,
         Hg, Wg, ws, int(pool), Cc, _stream())
    return Qw, Kw, Vw, nw, nq, nk


def window_attn(qkv, qkv_bias, F, Hg, Wg, ws, pool, Cc, heads):
    """Fused Hiera window attention (windows of <= 256 keys): qkv bf16 [F*Hg*Wg, 3*Cc] raster order -> bf16
    [F*Ho*Wo, Cc] raster order; partition, q max-pool, attention and un-partition in one kernel."""
    Ho, Wo = (Hg // 2, Wg // 2) if pool else (Hg, Wg)
    out = empty((F * Ho * Wo, Cc), BF16, qkv)
    call("usvm_window_attn_bf16", qkv.data_ptr(), qkv_bias.data_ptr(), out.data_ptr(), F, Hg, Wg, ws, int(pool), Cc,
         heads, _stream())
    return out


def window_scatter(Ow, F, Ho, Wo, wq, Cc):
    out = empty((F * Ho * Wo, Cc), BF16, Ow)
    call("usvm_window_scatter", Ow.data_ptr(), out.data_ptr(), F, Ho, Wo, wq, Cc, _stream())
    return out


def maxpool2(x, F, H, W, Cc):
    out = empty((F * (H // 2) * (W // 2), Cc), F32, x)
    call("usvm_maxpool2_nhwc", x.data_ptr(), out.data_ptr(), F, H, W, Cc, _stream())
    return out


def upsample2_add_(fine, coarse, F, H, W, Cc, bf16=False):
    o16 = empty((F * H * W, Cc), BF16, fine) if bf16 else None
    call("usvm_upsample2_add", fine.data_ptr(), coarse.data_ptr(), _ptr(o16), F, H, W, Cc, _stream())
    return o16


def im2col_patch(img, KP=160):
    _chk(img, F32, "img")
    F, _, S, _ = img.shape
    assert img.is_contiguous()
    A = empty((F * (S // 4) ** 2, KP), BF16, img)
    call("usvm_im2col_patch", img.data_ptr(), A.data_ptr(), F, S, KP, _stream())
    return A


def im2col_patch_grid(imgs, P):
    """fp32 [F,3,S,S] -> bf16 [F*(S/P)^2, 3*P*P]: non-overlapping P x P patches, columns (c, ky, kx)."""
    _chk(imgs, F32, "imgs")
    Fr, _, S, _ = imgs.shape
    A = empty((Fr * (S // P) * (S // P), 3 * P * P), BF16, imgs)
    call("usvm_im2col_patch_grid", imgs.data_ptr(), A.data_ptr(), Fr, S, P, _stream())
    return A


def normalize_gray_u8(gray, mean, std):
    F, H, W = gray.shape
    out = empty((F, 3, H, W), F32, gray)
    m = (C.c_float * 3)(*mean)
    s = (C.c_float * 3)(*std)
    call("usvm_normalize_gray_u8", gray.data_ptr(), out.data_ptr(), F, H, W, m, s, _stream())
    return out


def normalize_rgb_u8(rgb, mean, std, out=None):
    """uint8 CUDA [F,H,W,3] -> normalised fp32 [F,3,H,W] (into `out` when given)."""
    if not (rgb.is_cuda and rgb.dtype == torch.uint8 and rgb.dim() == 4 and rgb.shape[-1] == 3 and rgb.is_contiguous()):
        raise TypeError("normalize_rgb_u8: expected a contiguous CUDA uint8 tensor [F,H,W,3]")
    F, H, W, _ = rgb.shape
    if out is None:
        out = empty((F, 3, H, W), F32, rgb)
    m = (C.c_float * 3)(*mean)
    s = (C.c_float * 3)(*std)
    call("usvm_normalize_rgb_u8", rgb.data_ptr(), out.data_ptr(), F, H, W, m, s, _stream())
    return out


def build_memory(frames, tpos_rows, pos, tpos, ptrs, ptr_pos, B, T=1024, Cm=64):
    """frames: list of bf16 [B,T,Cm] tensors; returns (k_in, v_in) bf16 [B, Nk, Cm], Nk."""
    n_ptr = 0 if ptrs is None else ptrs.shape[1]
    Nk = len(frames) * T + n_ptr
    k_in = empty((B, Nk, Cm), BF16, pos)
    v_in = empty((B, Nk, Cm), BF16, pos)
    off = 0
    chunks = [frames[i:i + _lib.MAX_MEMORY_FRAMES] for i in range(0, len(frames), _lib.MAX_MEMORY_FRAMES)] or [[]]
    done = 0
    for ci, chunk in enumerate(chunks):
        mf = MemoryFrames()
        for i, f in enumerate(chunk):
            _chk(f, BF16, "memory frame")
            assert f.is_contiguous() and f.shape == (B, T, Cm)
            mf.mem[i] = f.data_ptr()
            mf.tpos_index[i] = tpos_rows[done + i]
        mf.count = len(chunk)
        last = ci == len(chunks) - 1
        np_here = n_ptr if last else 0
        if len(chunk) * T + np_here > 0:
            call("usvm_build_memory", C.byref(mf), pos.data_ptr(), tpos.data_ptr(), _ptr(ptrs) if last else 0,
                 _ptr(ptr_pos) if last else 0, k_in.data_ptr(), v_in.data_ptr(), B, T, Cm, np_here, Nk, off, _stream())
        off += len(chunk) * T
        done += len(chunk)
    return k_in, v_in, Nk


class FrameStore:
    """Per-session device store of everything a tracked frame leaves behind, indexed by frame number:
    spatial memory (bf16 token-major), object pointer, object score, hole-filled low-res mask logits."""

    def __init__(self, num_frames, B, device, T=1024, Cm=64, ptr_dim=256, low=128):
        self.num_frames, self.B = num_frames, B
        self.mem = torch.zeros((num_frames, B, T, Cm), dtype=BF16, device=device)
        self.ptr = torch.zeros((num_frames, B, ptr_dim), dtype=F32, device=device)
        self.score = torch.zeros((num_frames, B, 1), dtype=F32, device=device)
        self.masks = torch.zeros((num_frames, B, 1, low, low), dtype=F32, device=device)
        self.shared, self.column0 = None, 0  # set on the column views of a store shared by lock-step sessions

    def columns(self, lo, n):
        """The store of objects [lo, lo + n) as a VIEW of this one (per-session face of a store shared by several
        sessions tracked in lock-step); slot strides stay those of the shared store."""
        v = FrameStore.__new__(FrameStore)
        v.num_frames, v.B = self.num_frames, n
        v.mem, v.ptr = self.mem[:, lo:lo + n], self.ptr[:, lo:lo + n]
        v.score, v.masks = self.score[:, lo:lo + n], self.masks[:, lo:lo + n]
        v.shared, v.column0 = self, lo
        return v

    def grow(self, B):
        """Append zero-filled columns for objects added later; the existing objects keep their results."""
        def wider(t):
            out = torch.zeros((t.shape[0], B) + tuple(t.shape[2:]), dtype=t.dtype, device=t.device)
            out[:, : self.B].copy_(t)
            return out

        self.mem, self.ptr, self.score, self.masks = wider(self.mem), wider(self.ptr), wider(self.score), wider(self.masks)
        self.B, self.shared = B, None  # (no longer a column view of a shared store)

    def select_objects(self, keep):
        self.mem, self.ptr = self.mem[:, keep].contiguous(), self.ptr[:, keep].contiguous()
        self.score, self.masks = self.score[:, keep].contiguous(), self.masks[:, keep].contiguous()
        self.B, self.shared = len(keep), None


def new_frame_ctrl(device):
    """Device buffer holding one usvm_frame_ctrl."""
    return torch.zeros(C.sizeof(FrameCtrl) // 8 + 1, dtype=torch.int64, device=device)


def set_frame_ctrl(ctrl_dev, store, obj0, cur_frame, mem_frames, mem_tpos, ptr_frames, ptr_rel, copies=None):
    """Fill the device control block through a by-value kernel parameter (async, no staging copy).  `obj0` offsets the
    store bases so that a per-object run (B = 1) addresses object obj0 of a multi-object store.
    copies: up to 4 (src, dst) tensor pairs of equal size copied by the same launch (the frame's features into the static
    inputs of the captured frame graph)."""
    c = FrameCtrl()
    c.mem_store = store.mem.data_ptr() + 2 * obj0 * store.mem.stride(1)
    c.ptr_store = store.ptr.data_ptr() + 4 * obj0 * store.ptr.stride(1)
    c.score_store = store.score.data_ptr() + 4 * obj0 * store.score.stride(1)
    c.mask_store = store.masks.data_ptr() + 4 * obj0 * store.masks.stride(1)
    c.mem_slot_stride, c.ptr_slot_stride = store.mem.stride(0), store.ptr.stride(0)
    c.score_slot_stride, c.mask_slot_stride = store.score.stride(0), store.masks.stride(0)
    c.cur_frame, c.n_mem, c.n_ptr = cur_frame, len(mem_frames), len(ptr_frames)
    if len(mem_frames) > _lib.MAX_MEMORY_FRAMES or len(ptr_frames) > _lib.MAX_PTRS:
        raise RuntimeError(f"memory bank of {len(mem_frames)} frames / {len(ptr_frames)} object pointers exceeds the frame "
                           f"control block ({_lib.MAX_MEMORY_FRAMES} / {_lib.MAX_PTRS}): too many conditioning frames")
    for i, (f, t) in enumerate(zip(mem_frames, mem_tpos)):
        c.mem_frame[i], c.mem_tpos[i] = f, t
    for i, (f, r) in enumerate(zip(ptr_frames, ptr_rel)):
        c.ptr_frame[i], c.ptr_rel[i] = f, r
    if not copies:
        call("usvm_set_frame_ctrl", ctrl_dev.data_ptr(), C.byref(c), _stream())
        return
    n = len(copies)
    src, dst, nbytes = (C.c_void_p * n)(), (C.c_void_p * n)(), (C.c_longlong * n)()
    for i, (a, b) in enumerate(copies):
        if not (a.is_contiguous() and b.is_contiguous() and a.dtype == b.dtype and a.numel() == b.numel()):
            raise ValueError("frame prologue copies need contiguous tensors of equal type and size")
        src[i], dst[i], nbytes[i] = a.data_ptr(), b.data_ptr(), a.numel() * a.element_size()
    call("usvm_frame_prologue", ctrl_dev.data_ptr(), C.byref(c), src, dst, nbytes, n, _stream())


def finalize_memory(x, score, no_obj_embed, B, T=1024, Cm=64, ctrl=None):
    """-> bf16 [B,T,Cm]; with a device control block the result is written in place into slot ctrl->cur_frame of the
    frame store (returns None).  `score`: one value per object (any stride)."""
    sstride = score.stride(0) if score.dim() > 1 else 1
    if ctrl is None:
        mem = empty((B, T, Cm), BF16, x)
        call("usvm_finalize_memory", x.data_ptr(), score.data_ptr(), sstride, no_obj_embed.data_ptr(), mem.data_ptr(),
             B, T, Cm, 0, _stream())
        return mem
    call("usvm_finalize_memory", x.data_ptr(), score.data_ptr(), sstride, no_obj_embed.data_ptr(), 0, B, T, Cm,
         ctrl.data_ptr(), _stream())
    return None


def ptr_tpos(ctrl, W, bias, n_ptr):
    out = empty((n_ptr * 4, 64), F32, W)
    call("usvm_ptr_tpos", ctrl.data_ptr(), W.data_ptr(), bias.data_ptr(), out.data_ptr(), n_ptr, _stream())
    return out


def build_memory_store(ctrl, pos, tpos, ptr_pos, B, n_mem, n_ptr, T=1024, Cm=64):
    """k_in / v_in bf16 [B, Nk, Cm] gathered from the frame store named by the control block."""
    Nk = n_mem * T + n_ptr * 4
    k_in = empty((B, Nk, Cm), BF16, pos)
    v_in = empty((B, Nk, Cm), BF16, pos)
    call("usvm_build_memory_store", ctrl.data_ptr(), pos.data_ptr(), tpos.data_ptr(), _ptr(ptr_pos), k_in.data_ptr(),
         v_in.data_ptr(), B, T, Cm, n_mem, n_ptr, _stream())
    return k_in, v_in, Nk


def store_outputs(ctrl, obj_ptr, score, masks):
    """slot ctrl->cur_frame of the pointer / score / mask stores <- obj_ptr [B,256], score [B,(1)], masks [B,1,h,w]."""
    B = obj_ptr.shape[0]
    call("usvm_store_outputs", ctrl.data_ptr(), obj_ptr.data_ptr(), score.data_ptr(),
         score.stride(0) if score.dim() > 1 else 1, masks.data_ptr(), B, obj_ptr.shape[1], masks.numel() // B, _stream())


# ------------------------------------------------------------------------------------------------
# convs / resize
# ------------------------------------------------------------------------------------------------
def conv2d_small(x, w_kkio, bias, B, H, W, Cin, Cout, k, stride, pad, ln=None, eps=1e-6, gelu=False, bf16=False):
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    out = empty((B * Ho * Wo, Cout), BF16 if bf16 else F32, x)
    call("usvm_conv2d_small", x.data_ptr(), w_kkio.data_ptr(), bias.data_ptr(), _ptr(ln[0]) if ln else 0,
         _ptr(ln[1]) if ln else 0, eps, int(gelu), 0 if bf16 else out.data_ptr(), out.data_ptr() if bf16 else 0, B, H,
         W, Cin, Cout, k, stride, pad, _stream())
    return out, Ho, Wo


def conv2d_mask_first(low, post_mode, post_scale, post_bias, w_kkio, bias, B, H, W, k, stride, pad, ln=None, eps=1e-6,
                      gelu=False):
    """First MaskDownSampler stage on the virtual H x W upsampling of low [B,1,h,w] (bilinear + post_mode fused into the
    footprint load): -> (fp32 [B*Ho*Wo, 4], Ho, Wo)."""
    _chk(low, F32, "low")
    hi, wi = low.shape[-2:]
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    out = empty((B * Ho * Wo, 4), F32, low)
    call("usvm_conv2d_mask_first", low.data_ptr(), hi, wi, post_mode, post_scale, post_bias, w_kkio.data_ptr(),
         bias.data_ptr(), _ptr(ln[0]) if ln else 0, _ptr(ln[1]) if ln else 0, eps, int(gelu), out.data_ptr(), B, H, W, k,
         stride, pad, _stream())
    return out, Ho, Wo


def im2col_nhwc(x, B, H, W, Cc, k, stride, pad):
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    A = empty((B * Ho * Wo, k * k * Cc), BF16, x)
    call("usvm_im2col_nhwc", x.data_ptr(), A.data_ptr(), B, H, W, Cc, k, stride, pad, _stream())
    return A


def dwconv7_ln(x, w_49c, bias, ln_w, ln_b, B, H, W, Cc=256, eps=1e-6):
    out = empty((B * H * W, Cc), BF16, x)
    call("usvm_dwconv7_ln", x.data_ptr(), w_49c.data_ptr(), bias.data_ptr(), ln_w.data_ptr(), ln_b.data_ptr(), eps,
         out.data_ptr(), B, H, W, Cc, _stream())
    return out


def resize_bilinear(x, Ho, Wo, post=POST_NONE, scale=1.0, bias=0.0):
    """x fp32 [..., Hi, Wi] -> [..., Ho, Wo] (align_corners=False)."""
    _chk(x, F32, "x")
    x = x.contiguous()
    Hi, Wi = x.shape[-2:]
    out = empty((*x.shape[:-2], Ho, Wo), F32, x)
    call("usvm_resize_bilinear", x.data_ptr(), out.data_ptr(), x.numel() // (Hi * Wi), Hi, Wi, Ho, Wo, post, scale,
         bias, _stream())
    return out


def non_overlap(x, group=0, post=POST_NONE, scale=1.0, bias=0.0):
    """_apply_non_overlapping_constraints over the objects (dim 0) of x fp32 [B, ..., H, W], per group of `group`
    consecutive objects (0: all of them), fused with the post transform of resize_bilinear; returns a new tensor."""
    _chk(x, F32, "x")
    x = x.contiguous()
    B = x.shape[0]
    out = torch.empty_like(x)
    call("usvm_non_overlap_f32", x.data_ptr(), out.data_ptr(), B, x.numel() // B, group, post, scale, bias, _stream())
    return out


def resize_bilinear_aa(x, Ho, Wo, binarize_half=False):
    _chk(x, F32, "x")
    x = x.contiguous()
    Hi, Wi = x.shape[-2:]
    out = empty((*x.shape[:-2], Ho, Wo), F32, x)
    call("usvm_resize_bilinear_aa", x.data_ptr(), out.data_ptr(), x.numel() // (Hi * Wi), Hi, Wi, Ho, Wo,
         int(binarize_half), _stream())
    return out


# ------------------------------------------------------------------------------------------------
# decoder tail
# ------------------------------------------------------------------------------------------------
def upscale1_ln_gelu(g1, feat_s1, ln_w, ln_b, B, Hc, Wc, feat_group, eps=1e-6):
    """feat_group: consecutive objects that share one frame of feat_s1 (0: every object has its own)."""
    out = empty((B * 4 * Hc * Wc, 64), F32, g1)
    call("usvm_upscale1_ln_gelu", g1.data_ptr(), feat_s1.data_ptr(), ln_w.data_ptr(), ln_b.data_ptr(), eps,
         out.data_ptr(), B, Hc, Wc, 64, int(feat_group), _stream())
    return out


def upscale2_masks(g2, feat_s0, hyper, B, Hc, Wc, feat_group, hyper_bs=128):
    """hyper: [B, 4, 32] hyper-network outputs, object stride hyper_bs elements; feat_group as in upscale1_ln_gelu."""
    masks = empty((B, 4, 2 * Hc, 2 * Wc), F32, g2)
    call("usvm_upscale2_masks", g2.data_ptr(), feat_s0.data_ptr(), hyper.data_ptr(), hyper_bs, masks.data_ptr(), B, Hc,
         Wc, int(feat_group), _stream())
    return masks


def small_mlp3(x_ptr, x_row_stride, x_inst_stride, row_select, mlp, out_dim, rows, instances, like, sigmoid=False):
    """mlp = (w1, b1, w2, b2, w3, b3) fp32 stacked over instances; returns fp32 [rows, instances, out_dim]."""
    y = empty((rows, instances, out_dim), F32, like)
    call("usvm_small_mlp3", x_ptr, x_row_stride, x_inst_stride, _ptr(row_select), *(t.data_ptr() for t in mlp),
         out_dim, int(sigmoid), y.data_ptr(), instances * out_dim, out_dim, rows, instances, _stream())
    return y


def gemm_skinny(x, w, bias=None, M=None, x_rs=None, x_is=0, x2=None, x2_rs=None, x2_is=0, act=ACT_NONE,
                residual=None, r_rs=None, r_is=0, instances=1, row_select=None, x_sel_stride=0, x_ptr=None, out=None,
                x2_cols=0, ln=None, ln_out=None, ln_rs=None, ln_is=0):
    """fp32 GEMM for a few rows: out[i, m, :] = act((x + x2)[i, m] @ w[i].T + bias[i]) + residual[i, m].
    x: tensor [M, K] (or raw address via x_ptr with explicit strides); w: [N, K] or [instances, N, K].
    ln = (weight, bias, eps): the input rows are LayerNorm-ed on load (K == 256; x2 is added afterwards) and the normalised
    rows are also written to ln_out -- the same values a separate usvm_layernorm launch would produce."""
    N, K = w.shape[-2], w.shape[-1]
    if M is None:
        M = x.shape[0]
    p = SkinnyParams()
    p.x = x_ptr if x_ptr is not None else x.data_ptr()
    p.x_rs = x_rs if x_rs is not None else x.stride(0)
    p.x_is = x_is
    p.x2 = _ptr(x2)
    p.x2_rs = (x2_rs if x2_rs is not None else x2.stride(0)) if x2 is not None else 0
    p.x2_is = x2_is
    p.row_select, p.x_sel_stride = _ptr(row_select), x_sel_stride
    p.w, p.w_is = w.data_ptr(), (N * K if instances > 1 else 0)
    p.bias, p.b_is = _ptr(bias), (N if instances > 1 else 0)
    if out is None:
        out = empty((M, instances * N) if instances > 1 else (M, N), F32, w)
    p.residual = _ptr(residual)
    p.r_rs = (r_rs if r_rs is not None else residual.stride(0)) if residual is not None else 0
    p.r_is = r_is
    p.out, p.o_rs, p.o_is = out.data_ptr(), out.stride(0), (N if instances > 1 else 0)
    p.M, p.N, p.K, p.instances, p.act = M, N, K, instances, act
    p.x2_cols = x2_cols  # x2 only for output columns < x2_cols (0: all)
    if ln is not None:
        p.ln_w, p.ln_b, p.ln_eps = ln[0].data_ptr(), ln[1].data_ptr(), ln[2]
        p.ln_out = _ptr(ln_out)
        p.ln_rs = (ln_rs if ln_rs is not None else ln_out.stride(0)) if ln_out is not None else 0
        p.ln_is = ln_is
    call("usvm_gemm_skinny_f32", C.byref(p), _stream())
    return out


_T2I_COUNTERS = {}


def attn_t2i(q, k, v, B, Nt, Nk, H=8):
    """token->image attention, head_dim 16: q [B*Nt, H*16]; k, v column views of an image-side buffer."""
    out = empty((B * Nt, H * 16), F32, q)
    assert k.stride(0) == v.stride(0)
    splits = (Nk + 127) // 128
    if 1 < splits <= 64:  # key-split kernel: B*H*splits CTAs, the last one per head merges the partials
        key = (q.device, B * H)
        if key not in _T2I_COUNTERS:
            _T2I_COUNTERS[key] = torch.zeros(B * H, dtype=torch.int32, device=q.device)
        part = empty((B * H * splits * Nt * 18,), F32, q)
        call("usvm_attn_t2i_split_f32", q.data_ptr(), q.stride(0), k.data_ptr(), v.data_ptr(), k.stride(0),
             out.data_ptr(), out.stride(0), B, H, Nt, Nk, 0.25, part.data_ptr(), _T2I_COUNTERS[key].data_ptr(), _stream())
        return out
    call("usvm_attn_t2i_f32", q.data_ptr(), q.stride(0), k.data_ptr(), v.data_ptr(), k.stride(0), out.data_ptr(),
         out.stride(0), B, H, Nt, Nk, 0.25, _stream())
    return out


def attn_i2t(q, k, v, B, Nq, Nt, H=8):
    """image->token attention, head_dim 16: q [B*Nq, .] (column view), k, v [B*Nt, H*16]."""
    out = empty((B * Nq, H * 16), F32, q)
    assert k.stride(0) == v.stride(0)
    call("usvm_attn_i2t_f32", q.data_ptr(), q.stride(0), k.data_ptr(), v.data_ptr(), k.stride(0), out.data_ptr(),
         out.stride(0), B, H, Nq, Nt, 0.25, _stream())
    return out


def sam_select(masks, iou, score, multimask, delta, thresh, no_obj_score, iou_stride=4, score_stride=1,
               iou_is_logit=False):
    """iou: 4 values per object at stride iou_stride (sigmoid applied in-kernel when iou_is_logit);
    score: one value per object at stride score_stride."""
    B, _, H, W = masks.shape
    low = empty((B, 1, H, W), F32, masks)
    idx = empty((B,), torch.int32, masks)
    iou_sel = empty((B, 1), F32, masks)
    call("usvm_sam_select", masks.data_ptr(), iou.data_ptr(), iou_stride, int(iou_is_logit), score.data_ptr(),
         score_stride, int(multimask), delta, thresh, no_obj_score, low.data_ptr(), idx.data_ptr(), iou_sel.data_ptr(),
         B, H * W, _stream())
    return low, idx, iou_sel


def objptr_mix_(ptr, score, no_obj_ptr):
    call("usvm_objptr_mix", ptr.data_ptr(), score.data_ptr(), score.stride(0), no_obj_ptr.data_ptr(), ptr.shape[0],
         ptr.shape[1], _stream())
    return ptr


def point_embed(coords, labels, gauss, table, image_size):
    n = coords.shape[0] * coords.shape[1]
    out = empty((*coords.shape[:2], 256), F32, coords)
    call("usvm_point_embed", coords.data_ptr(), labels.data_ptr(), gauss.data_ptr(), table.data_ptr(),
         float(image_size), out.data_ptr(), n, _stream())
    return out


# ------------------------------------------------------------------------------------------------
# connected components / hole filling
# ------------------------------------------------------------------------------------------------
def connected_components(mask_u8):
    """uint8 CUDA [N,1,H,W] -> (labels int32, counts int32), the sam2._C contract."""
    if not mask_u8.is_cuda:
        raise RuntimeError("inputs must be a CUDA tensor")
    if mask_u8.dim() != 4 or mask_u8.shape[1] != 1:
        raise RuntimeError("inputs must be [N, 1, H, W] shape")
    if mask_u8.dtype != torch.uint8:
        raise RuntimeError("inputs must be a uint8 type")
    N, _, H, W = mask_u8.shape
    if H % 2 or W % 2:
        raise RuntimeError("height and width must be even numbers")
    mask_u8 = mask_u8.contiguous()
    labels = torch.empty((N, 1, H, W), dtype=torch.int32, device=mask_u8.device)
    counts = torch.empty_like(labels)
    if N:
        call("usvm_cc2d_label_u8", mask_u8.data_ptr(), labels.data_ptr(), counts.data_ptr(), N, H, W, _stream())
    return labels, counts


def largest_component_3d(vol):
    """uint8 / bool CUDA [D,H,W] -> uint8 [D,H,W] mask of the largest 26-connected component (all zero if empty)."""
    if not vol.is_cuda or vol.dim() != 3:
        raise RuntimeError("largest_component_3d expects a CUDA tensor [D, H, W]")
    v = vol.to(torch.uint8).contiguous()
    D, H, W = v.shape
    out = torch.empty_like(v)
    par = torch.empty((D, H, W), dtype=torch.int32, device=v.device)
    cnt = torch.empty_like(par)
    best = torch.empty(1, dtype=torch.int64, device=v.device)
    call("usvm_cc3d_largest_u8", v.data_ptr(), out.data_ptr(), par.data_ptr(), cnt.data_ptr(), best.data_ptr(), D, H, W,
         _stream())
    return out


def fill_holes(scores, max_area, fill_value=0.1):
    """Fused fill_holes_in_mask_scores: fp32 CUDA [N,1,H,W] -> new tensor."""
    _chk(scores, F32, "scores")
    scores = scores.contiguous()
    N, _, H, W = scores.shape
    out = torch.empty_like(scores)
    sl = sc = None
    if (H // 2) * (W // 2) * 8 + H * W > 200 * 1024:
        sl = torch.empty((N, H, W), dtype=torch.int32, device=scores.device)
        sc = torch.empty_like(sl)
    call("usvm_fill_holes_f32", scores.data_ptr(), out.data_ptr(), _ptr(sl), _ptr(sc), N, H, W, int(max_area),
         float(fill_value), _stream())
    return out
