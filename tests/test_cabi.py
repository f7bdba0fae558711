"""CPU checks of the drop-in boundary: libvpf.so builds, loads without a GPU/driver, and exports
every symbol include/vpf.h declares; host-only entry points agree with the oracle."""
import os
import re

import pytest


def _declared_symbols():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, "include", "vpf.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vpf_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from vitparticlefiltertracker_b200 import _lib
    lib = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/vpf.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature"
    assert set(_lib.SIGNATURES) == set(names)
    assert lib.vpf_abi_version() == 7


def test_host_only_entry_points():
    from vitparticlefiltertracker_b200 import _lib
    from oracle.philox import philox_uniform53
    lib = _lib.load()
    for seed, t, s in [(0, 1, 0), (123456789012345, 77, 3), (2 ** 64 - 1, 2 ** 32 - 1, 9)]:
        assert lib.vpf_resample_uniform(seed, t, s) == philox_uniform53(t, s, seed)
    cfg = _lib.VitConfig(768, 12, 12, 4096, 0)
    import ctypes as C
    need = lib.vpf_vit_workspace_bytes(C.byref(cfg))
    rows = 4096 * 197
    assert need >= rows * 768 * (4 + 2 + 8)
    # argument validation happens before any CUDA call
    assert lib.vpf_vit_create(None, C.byref(cfg)) == -1
    assert b"null" in lib.vpf_last_error()
    h = C.c_void_p()
    bad = _lib.VitConfig(100, 12, 12, 16, 0)
    assert lib.vpf_vit_create(C.byref(h), C.byref(bad)) == -1
    assert lib.vpf_vit_create(C.byref(h), C.byref(cfg)) == 0 and h.value
    assert lib.vpf_vit_launches(h) == 2 + 7 * 11 + 10        # CLS-only last block
    assert lib.vpf_vit_forward(h, 1, 1, 1, None) == -3          # weights / workspace missing
    assert lib.vpf_vit_destroy(h) == 0


def test_product_package_does_not_import_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "vitparticlefiltertracker_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f


def test_tracker_requires_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    from vitparticlefiltertracker_b200 import Tracker
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Tracker("vit_tiny_patch16_224", 16)
