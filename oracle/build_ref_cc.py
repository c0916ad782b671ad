"""TEST INFRASTRUCTURE -- builds the REFERENCE's own connected-components CUDA extension for sm_100a.

    python oracle/build_ref_cc.py        # -> oracle/_ref/ref_sam2_C.so   (git-ignored, travels to the GPU box)

The source is compiled from where it lies (/root/reference/sam2/csrc/connected_components.cu, nothing is copied into
the repo), with the same recipe as the reference's setup.py:96-106 (`CUDAExtension("sam2._C", ...)`) but
`-gencode arch=compute_100a,code=sm_100a`.  tests/test_gpu_cc.py loads the result on the GPU box (where /root/reference
does not exist) and holds this repo's kernel to the reference kernel's labels and counts bit for bit.
Only tests / smoke / bench may touch anything under oracle/.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SRC = "/root/reference/sam2/csrc/connected_components.cu"
OUT_DIR = os.path.join(ROOT, "oracle", "_ref")
NAME = "ref_sam2_C"


def so_path():
    return os.path.join(OUT_DIR, NAME + ".so")


def build(verbose=False):
    """Returns the path of the built module, or None when the reference tree is not mounted (GPU box)."""
    if not os.path.exists(REF_SRC):
        return so_path() if os.path.exists(so_path()) else None
    if os.path.exists(so_path()) and os.path.getmtime(so_path()) >= os.path.getmtime(REF_SRC):
        return so_path()
    os.makedirs(OUT_DIR, exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    os.environ.setdefault("MAX_JOBS", "4")
    from torch.utils import cpp_extension

    cpp_extension.load(name=NAME, sources=[REF_SRC], build_directory=OUT_DIR, is_python_module=False, verbose=verbose,
                       extra_cuda_cflags=["-DCUDA_HAS_FP16=1", "-D__CUDA_NO_HALF_OPERATORS__",
                                          "-D__CUDA_NO_HALF_CONVERSIONS__", "-D__CUDA_NO_HALF2_OPERATORS__",
                                          "-gencode", "arch=compute_100a,code=sm_100a"])
    return so_path()


def load():
    """Import the built reference op: returns a callable get_connected_componnets(uint8 CUDA [N,1,H,W]) or None."""
    path = so_path()
    if not os.path.exists(path):
        return None
    import importlib.util

    import torch  # noqa: F401  (libtorch must be loaded first)

    spec = importlib.util.spec_from_file_location(NAME, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.get_connected_componnets


if __name__ == "__main__":
    p = build(verbose="-v" in sys.argv)
    print("built" if p else "reference tree not mounted and no prebuilt module", p)
