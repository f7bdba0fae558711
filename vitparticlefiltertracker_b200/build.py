"""Builds libvpf.so (hand-written sm_100a CUDA behind the C ABI of include/vpf.h) in-tree.

nvcc cross-compiles without a GPU; the static CUDA runtime is linked so the library loads on
driver-less machines (symbol / ABI tests), and the driver is only touched at run time.
"""
from __future__ import annotations

import fcntl
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libvpf.so")
SOURCES = ["pf.cu", "crop.cu", "vit.cu"]
# crop.cu reproduces the oracle's fp32 roundings operation by operation: no FMA contraction of its
# plain a*b+c expressions (explicit .rn intrinsics carry the rest)
EXTRA_FLAGS = {"crop.cu": ["-fmad=false"]}
HEADERS = sorted(f for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))) + ["../../include/vpf.h"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=default", "--expt-relaxed-constexpr",
    "-cudart", "static",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libvpf.so cannot be built (there is no CPU fallback)")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps)


def build_variant(out_path: str, defines) -> str:
    """Builds a second library with extra -D flags (same sources, same ABI) for same-box A/B runs
    (tools/ab_bench.sh, VPF_LIB_PATH); objects go to a scratch directory."""
    import tempfile
    nvcc = _nvcc()
    with tempfile.TemporaryDirectory() as tmp:
        objs = []
        for src in SOURCES:
            obj = os.path.join(tmp, src.replace(".cu", ".o"))
            subprocess.run([nvcc, *NVCC_FLAGS, *EXTRA_FLAGS.get(src, []), *[f"-D{d}" for d in defines],
                            "-c", os.path.join(CSRC, src), "-o", obj], check=True)
            objs.append(obj)
        subprocess.run([nvcc, "-shared", "-cudart", "static", "-gencode",
                        "arch=compute_100a,code=sm_100a", "-o", out_path, *objs], check=True)
    return out_path


def build(force: bool = False, verbose: bool = False) -> str:
    """Compiles csrc/*.cu and links libvpf.so.  Safe under torchrun: ranks serialise on a file
    lock (the first one builds, the others find the library up to date), and the library is
    linked under a temporary name and renamed into place, so nobody dlopens a half-written file."""
    if not force and not needs_build():
        return LIB
    with open(os.path.join(HERE, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not needs_build():      # another process built it while we waited
                return LIB
            return _build_locked(verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(verbose: bool) -> str:
    nvcc = _nvcc()
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(CSRC, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, *EXTRA_FLAGS.get(src, []), "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out.decode())
        if p.returncode:
            raise RuntimeError(f"nvcc failed on {src}")
    tmp = f"{LIB}.{os.getpid()}.tmp"
    cmd = [nvcc, "-shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a",
           "-o", tmp, *objs]
    subprocess.run(cmd, check=True)
    os.replace(tmp, LIB)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
