"""ctypes binding of libvpf.so (include/vpf.h).  There is no CPU or PyTorch fallback: if the
CUDA library is missing or fails to load, importing a compute entry point raises."""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

_c_void_p, _c_int, _c_float, _c_double = C.c_void_p, C.c_int, C.c_float, C.c_double
_u64, _u32, _size_t = C.c_uint64, C.c_uint32, C.c_size_t


class VitConfig(C.Structure):
    _fields_ = [("dim", C.c_int32), ("depth", C.c_int32), ("heads", C.c_int32),
                ("max_crops", C.c_int32), ("dtype", C.c_int32), ("crop_size", C.c_int32),
                ("residual_dtype", C.c_int32)]


DTYPE_BF16, DTYPE_F16 = 0, 1
RESIDUAL_F32, RESIDUAL_F16 = 0, 1
ABI_VERSION = 7          # VPF_ABI_VERSION in include/vpf.h


# name -> (restype, argtypes); mirrors include/vpf.h one to one
SIGNATURES = {
    "vpf_abi_version": (_c_int, []),
    "vpf_last_error": (C.c_char_p, []),
    "vpf_init_particles": (_c_int, [_c_void_p, _c_int, _c_int, _c_float, _c_float, _c_float, _u64,
                                    _u32, _u32, _c_void_p]),
    "vpf_propagate": (_c_int, [_c_void_p, _c_int, _c_int, _c_float, _c_float, _c_float, _u64, _u32,
                               _u32, _u32, _c_void_p]),
    "vpf_frame_advance": (_c_int, [_c_void_p, _c_void_p]),
    "vpf_propagate_dev": (_c_int, [_c_void_p, _c_int, _c_int, _c_float, _c_float, _c_float, _u64,
                                   _c_void_p, _u32, _u32, _c_void_p]),
    "vpf_resample_dev": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_int, _u32, _u64, _u32, _c_void_p,
                                  _c_void_p, _c_void_p, _c_int, _c_void_p, _c_void_p]),
    "vpf_philox_fill": (_c_int, [_c_void_p, _c_int, _u32, _u32, _u32, _u32, _u64, _c_void_p]),
    "vpf_normals_fill": (_c_int, [_c_void_p, _c_int, _u32, _u32, _u32, _u64, _c_void_p]),
    "vpf_resample_uniform": (_c_double, [_u64, _u32, _u32]),
    "vpf_crop_state": (_c_int, [_c_void_p, _c_int, _c_int, _c_void_p, _c_int, _c_int, _c_float,
                                _c_float, _c_void_p, _c_int, _c_int, _c_void_p]),
    "vpf_crop_boxes": (_c_int, [_c_void_p, _c_int, _c_int, _c_void_p, _c_int, _c_void_p, _c_int,
                                _c_int, _c_void_p]),
    "vpf_vit_create": (_c_int, [C.POINTER(_c_void_p), C.POINTER(VitConfig)]),
    "vpf_vit_destroy": (_c_int, [_c_void_p]),
    "vpf_vit_workspace_bytes": (_size_t, [C.POINTER(VitConfig)]),
    "vpf_vit_set_weights": (_c_int, [_c_void_p, C.POINTER(_c_void_p), _c_int]),
    "vpf_vit_set_workspace": (_c_int, [_c_void_p, _c_void_p, _size_t]),
    "vpf_vit_forward": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_void_p, _c_void_p]),
    "vpf_gemm_bf16": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int,
                               _c_int, _c_int, _c_int, _c_int, _c_int, _c_void_p]),
    "vpf_layernorm": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int,
                               _c_void_p]),
    "vpf_attention": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_void_p]),
    "vpf_vit_set_ln_fold": (_c_int, [_c_void_p, C.POINTER(_c_void_p), _c_int]),
    "vpf_vit_ln_fold": (_c_int, [_c_void_p]),
    "vpf_ln_stats_parts": (_c_int, [_c_int]),
    "vpf_rowstats16": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_int, _c_void_p]),
    "vpf_gemm_ln": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_void_p,
                             _c_int, _c_int, _c_int, _c_int, _c_int, _c_void_p]),
    "vpf_gemm_residual_stats": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int,
                                         _c_int, _c_int, _c_int, _c_void_p]),
    "vpf_vit_launches": (_c_int, [_c_void_p]),
    "vpf_vit_set_timing": (_c_int, [_c_void_p, _c_int]),
    "vpf_vit_timing": (_c_int, [_c_void_p, _c_void_p, _c_int]),
    "vpf_loglik": (_c_int, [_c_void_p, _c_int, _c_void_p, _c_int, _c_float, _c_void_p, _c_void_p,
                            _c_void_p]),
    "vpf_sumexp": (_c_int, [_c_void_p, _c_int, _c_void_p, _c_int, _c_void_p, _c_void_p]),
    "vpf_weights": (_c_int, [_c_void_p, _c_int, _c_void_p, _c_void_p, _c_int, _c_void_p, _c_int,
                             _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_void_p]),
    "vpf_resample": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_int, _u32, _c_double, _c_void_p,
                              _c_void_p, _c_int, _c_void_p, _c_void_p]),
    "vpf_resample_records": (_c_int, [_c_void_p, C.c_longlong, _c_int, _c_void_p, _c_int, _c_int,
                                      _u32, _c_double, _c_void_p, _u64, _u32, _c_void_p, _c_int,
                                      _c_void_p, _c_void_p, _c_void_p]),
}

_lib = None


def lib_path() -> str:
    """libvpf.so of this tree; VPF_LIB_PATH points at another build of the same ABI (same-box A/B
    runs of two kernel versions: tools/ab_bench.sh)."""
    return os.environ.get("VPF_LIB_PATH") or _build.LIB


def load(build_if_missing: bool = True):
    """Loads libvpf.so (building it with nvcc first if needed).  Raises if that is impossible."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if build_if_missing and path == _build.LIB and (not os.path.exists(path) or _build.needs_build()):
        _build.build()
    if not os.path.exists(path):
        raise RuntimeError(f"{path} not found: the CUDA extension is required (no CPU fallback)")
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError here == ABI mismatch with include/vpf.h
        fn.restype = res
        fn.argtypes = args
    if lib.vpf_abi_version() != ABI_VERSION:
        raise RuntimeError("libvpf.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().vpf_last_error().decode(errors="replace")
        raise RuntimeError(f"libvpf {what} failed (rc={rc}): {msg}")


def ptr(t) -> int:
    """Device pointer of a torch tensor (must be contiguous) or None -> NULL."""
    if t is None:
        return None
    assert t.is_contiguous(), "libvpf needs contiguous tensors"
    return t.data_ptr()
