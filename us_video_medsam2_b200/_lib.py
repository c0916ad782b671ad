"""ctypes binding of csrc/libusvm2_b200.so (C ABI declared in include/usvm2_b200.h).

The product path has NO fallback: if the shared library is missing or a kernel call fails, an
exception is raised.  `build()` compiles the library in-tree with nvcc for sm_100a.
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(CSRC, "libusvm2_b200.so")

MAX_MEMORY_FRAMES = 32

ACT_NONE, ACT_RELU, ACT_GELU = 0, 1, 2
POST_NONE, POST_SIGMOID_AFFINE, POST_BINARIZE_AFFINE = 0, 1, 2


class GemmEpilogue(C.Structure):
    _fields_ = [("bias", C.c_void_p), ("col_scale", C.c_void_p), ("residual", C.c_void_p), ("ldr", C.c_int),
                ("res_mod", C.c_int), ("act", C.c_int), ("out_f32", C.c_void_p), ("ldo_f32", C.c_int),
                ("out_bf16", C.c_void_p), ("ldo_bf16", C.c_int),
                ("rope_cos", C.c_void_p), ("rope_sin", C.c_void_p), ("rope_cols", C.c_int),
                ("rope_rows_per_batch", C.c_int), ("rope_n_rope", C.c_int), ("rope_table_rows", C.c_int),
                ("ln_w", C.c_void_p), ("ln_b", C.c_void_p), ("ln_eps", C.c_float), ("ln_gelu", C.c_int),
                ("res_div", C.c_int)]


class FmhaParams(C.Structure):
    _fields_ = [("q", C.c_void_p), ("k", C.c_void_p), ("v", C.c_void_p), ("o", C.c_void_p),
                ("q_bs", C.c_longlong), ("k_bs", C.c_longlong), ("v_bs", C.c_longlong), ("o_bs", C.c_longlong),
                ("q_rs", C.c_int), ("k_rs", C.c_int), ("v_rs", C.c_int), ("o_rs", C.c_int),
                ("q_hs", C.c_int), ("k_hs", C.c_int), ("v_hs", C.c_int), ("o_hs", C.c_int),
                ("B", C.c_int), ("H", C.c_int), ("Nq", C.c_int), ("Nk", C.c_int), ("head_dim", C.c_int),
                ("num_splits", C.c_int), ("o_part", C.c_void_p), ("ml_part", C.c_void_p), ("scale", C.c_float),
                ("part_bf16", C.c_int)]


class HieraAttnParams(C.Structure):
    _fields_ = [("qkv", C.c_void_p), ("out", C.c_void_p), ("qkv_bias", C.c_void_p), ("F", C.c_int), ("H", C.c_int),
                ("W", C.c_int), ("dim", C.c_int), ("heads", C.c_int), ("window", C.c_int), ("scale", C.c_float),
                ("pool", C.c_int)]


MAX_PTRS = 48


class SkinnyParams(C.Structure):
    _fields_ = [("x", C.c_void_p), ("x_is", C.c_longlong), ("x_rs", C.c_longlong),
                ("x2", C.c_void_p), ("x2_is", C.c_longlong), ("x2_rs", C.c_longlong),
                ("row_select", C.c_void_p), ("x_sel_stride", C.c_longlong),
                ("w", C.c_void_p), ("w_is", C.c_longlong), ("bias", C.c_void_p), ("b_is", C.c_longlong),
                ("residual", C.c_void_p), ("r_is", C.c_longlong), ("r_rs", C.c_longlong),
                ("out", C.c_void_p), ("o_is", C.c_longlong), ("o_rs", C.c_longlong),
                ("M", C.c_int), ("N", C.c_int), ("K", C.c_int), ("instances", C.c_int), ("act", C.c_int),
                ("x2_cols", C.c_int),
                ("ln_w", C.c_void_p), ("ln_b", C.c_void_p), ("ln_out", C.c_void_p), ("ln_is", C.c_longlong),
                ("ln_rs", C.c_longlong), ("ln_eps", C.c_float)]


class FrameCtrl(C.Structure):
    _fields_ = [("mem_store", C.c_void_p), ("ptr_store", C.c_void_p), ("score_store", C.c_void_p),
                ("mask_store", C.c_void_p), ("mem_slot_stride", C.c_longlong), ("ptr_slot_stride", C.c_longlong),
                ("score_slot_stride", C.c_longlong), ("mask_slot_stride", C.c_longlong),
                ("cur_frame", C.c_int), ("n_mem", C.c_int), ("n_ptr", C.c_int), ("reserved", C.c_int),
                ("mem_frame", C.c_int * MAX_MEMORY_FRAMES), ("mem_tpos", C.c_int * MAX_MEMORY_FRAMES),
                ("ptr_frame", C.c_int * MAX_PTRS), ("ptr_rel", C.c_float * MAX_PTRS)]


class MemoryFrames(C.Structure):
    _fields_ = [("mem", C.c_void_p * MAX_MEMORY_FRAMES), ("tpos_index", C.c_int * MAX_MEMORY_FRAMES),
                ("count", C.c_int)]


_P, _I, _F, _LL = C.c_void_p, C.c_int, C.c_float, C.c_longlong

# name -> argtypes (every function returns int)
_SIGNATURES = {
    "usvm_abi_version": [],
    "usvm_device_sm": [],
    "usvm_set_sm_budget": [_I],
    "usvm_cc2d_label_u8": [_P, _P, _P, _I, _I, _I, _P],
    "usvm_cc3d_largest_u8": [_P, _P, _P, _P, _P, _I, _I, _I, _P],
    "usvm_fill_holes_f32": [_P, _P, _P, _P, _I, _I, _I, _I, _F, _P],
    "usvm_gemm_bf16_tc5": [_P, _I, _P, _I, C.POINTER(GemmEpilogue), _I, _I, _I, _I, _P],
    "usvm_gemm_tf32_tc5": [_P, _I, _P, _I, C.POINTER(GemmEpilogue), _I, _I, _I, _I, _P],
    "usvm_debug_pgemm_profile": [_P],
    "usvm_ffn_fused_tc5": [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P],
    "usvm_gemm_simt": [_P, _I, _I, _P, _I, _I, C.POINTER(GemmEpilogue), _I, _I, _I, _P],
    "usvm_fmha_bf16": [C.POINTER(FmhaParams), _P],
    "usvm_fmha_tc5": [C.POINTER(FmhaParams), _P],
    "usvm_fmha_combine": [C.POINTER(FmhaParams), _P],
    "usvm_hiera_attn_tc5": [C.POINTER(HieraAttnParams), _P],
    "usvm_attn_small_f32": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _F, _P],
    "usvm_layernorm": [_P, _I, _P, _P, _F, _I, _P, _I, _P, _I, _I, _I, _P],
    "usvm_axpby_rows": [_P, _P, _F, _F, _I, _I, _I, _P, _P, _LL, _I, _P],
    "usvm_cast_f32_bf16": [_P, _P, _LL, _P],
    "usvm_rope_bf16": [_P, _I, _P, _P, _P, _I, _LL, _I, _I, _I, _I, _P],
    "usvm_window_gather": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "usvm_window_scatter": [_P, _P, _I, _I, _I, _I, _I, _P],
    "usvm_maxpool2_nhwc": [_P, _P, _I, _I, _I, _I, _P],
    "usvm_upsample2_add": [_P, _P, _P, _I, _I, _I, _I, _P],
    "usvm_im2col_patch": [_P, _P, _I, _I, _I, _P],
    "usvm_im2col_patch_grid": [_P, _P, _I, _I, _I, _P],
    "usvm_normalize_gray_u8": [_P, _P, _I, _I, _I, C.POINTER(C.c_float), C.POINTER(C.c_float), _P],
    "usvm_normalize_rgb_u8": [_P, _P, _I, _I, _I, C.POINTER(C.c_float), C.POINTER(C.c_float), _P],
    "usvm_build_memory": [C.POINTER(MemoryFrames), _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "usvm_conv2d_small": [_P, _P, _P, _P, _P, _F, _I, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    "usvm_conv2d_mask_first": [_P, _I, _I, _I, _F, _F, _P, _P, _P, _P, _F, _I, _P, _I, _I, _I, _I, _I, _I, _P],
    "usvm_im2col_nhwc": [_P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "usvm_dwconv7_ln": [_P, _P, _P, _P, _P, _F, _P, _I, _I, _I, _I, _P],
    "usvm_resize_bilinear": [_P, _P, _LL, _I, _I, _I, _I, _I, _F, _F, _P],
    "usvm_non_overlap_f32": [_P, _P, _I, _LL, _I, _I, _F, _F, _P],
    "usvm_resize_bilinear_aa": [_P, _P, _LL, _I, _I, _I, _I, _I, _P],
    "usvm_upscale1_ln_gelu": [_P, _P, _P, _P, _F, _P, _I, _I, _I, _I, _I, _P],
    "usvm_upscale2_masks": [_P, _P, _P, _I, _P, _I, _I, _I, _I, _P],
    "usvm_finalize_memory": [_P, _P, _I, _P, _P, _I, _I, _I, _P, _P],
    "usvm_set_frame_ctrl": [_P, C.POINTER(FrameCtrl), _P],
    "usvm_frame_prologue": [_P, C.POINTER(FrameCtrl), C.POINTER(_P), C.POINTER(_P), C.POINTER(_LL), _I, _P],
    "usvm_ptr_tpos": [_P, _P, _P, _P, _I, _P],
    "usvm_build_memory_store": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    "usvm_store_outputs": [_P, _P, _P, _I, _P, _I, _I, _I, _P],
    "usvm_small_mlp3": [_P, _LL, _LL, _P, _P, _P, _P, _P, _P, _P, _I, _I, _P, _LL, _LL, _I, _I, _P],
    "usvm_window_attn_bf16": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "usvm_gemm_skinny_f32": [C.POINTER(SkinnyParams), _P],
    "usvm_attn_t2i_f32": [_P, _I, _P, _P, _I, _P, _I, _I, _I, _I, _I, _F, _P],
    "usvm_attn_t2i_split_f32": [_P, _I, _P, _P, _I, _P, _I, _I, _I, _I, _I, _F, _P, _P, _P],
    "usvm_attn_i2t_f32": [_P, _I, _P, _P, _I, _P, _I, _I, _I, _I, _I, _F, _P],
    "usvm_sam_select": [_P, _P, _I, _I, _P, _I, _I, _F, _F, _F, _P, _P, _P, _I, _I, _P],
    "usvm_objptr_mix": [_P, _P, _I, _P, _I, _I, _P],
    "usvm_point_embed": [_P, _P, _P, _P, _F, _P, _I, _P],
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)


class KernelLibraryError(RuntimeError):
    pass


def build(verbose=False):
    """Compile every CUDA source for sm_100a into csrc/libusvm2_b200.so (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", CSRC, "-j8"], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout[-4000:], r.stderr[-4000:])
    if r.returncode != 0:
        raise KernelLibraryError("nvcc build of libusvm2_b200.so failed")
    return LIB_PATH


_lib = None
launch_count = 0  # kernels-launching calls made through this binding (bench.py reports it as gpu_launches)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise KernelLibraryError(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU or PyTorch fallback for this path)")
        handle = C.CDLL(LIB_PATH)
        for name, argtypes in _SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the .so is stale
            fn.argtypes = argtypes
            fn.restype = C.c_int
        if handle.usvm_abi_version() != 1:
            raise KernelLibraryError("libusvm2_b200.so ABI version mismatch; rebuild")
        _lib = handle
    return _lib


_bound_device = None


def bind_device(index):
    """One process drives ONE GPU (the layout of bench.py / sharding.py: one rank per device).  The kernels' opt-in
    shared-memory sizes (cudaFuncSetAttribute) are configured once per process on first launch, and that attribute is
    per device -- so a second device in the same process is refused here, loudly, instead of failing at launch."""
    global _bound_device
    if _bound_device is None:
        _bound_device = int(index)
    elif _bound_device != int(index):
        raise KernelLibraryError(f"this process already runs the kernel library on cuda:{_bound_device}; use one "
                                 f"process per GPU (asked for cuda:{int(index)})")


def call(name, *args):
    """Invoke a kernel entry point; raises on a non-zero status."""
    global launch_count
    rc = getattr(lib(), name)(*args)
    if rc != 0:
        raise KernelLibraryError(f"{name} failed with status {rc} "
                                 "(-1 bad argument, -2 CUDA launch error, -3 driver entry point)")
    launch_count += 1
