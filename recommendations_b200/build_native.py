"""Builds recommendations_b200/lib/librecemb_b200.so with nvcc for sm_100a (in-tree).

The library has no torch / Python dependency: plain C ABI (include/recemb_b200.h),
static cudart.  `python -m recommendations_b200.build_native [--force]`.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIBDIR = PKG / "lib"
LIB = LIBDIR / "librecemb_b200.so"
OBJDIR = PKG / "build"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-O3",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: the CUDA toolkit is required to build recemb_b200")


def _digest(src: Path) -> str:
    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS).encode())
    for p in [src, *sorted(CSRC.glob("*.cuh")), PKG.parent / "include" / "recemb_b200.h"]:
        h.update(p.read_bytes())
    return h.hexdigest()


_compiled = []  # translation units nvcc actually compiled in this process (the rest were reused)


def _compile(src: Path, force: bool) -> Path:
    obj = OBJDIR / (src.stem + ".o")
    stamp = OBJDIR / (src.stem + ".sha")
    dig = _digest(src)
    if not force and obj.exists() and stamp.exists() and stamp.read_text() == dig:
        return obj
    _compiled.append(src.name)
    cmd = [_nvcc(), *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    (OBJDIR / (src.stem + ".ptxas.log")).write_text(res.stderr)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError(f"nvcc failed for {src.name}")
    stamp.write_text(dig)
    return obj


def build(force: bool = False, verbose: bool = True) -> Path:
    del _compiled[:]
    OBJDIR.mkdir(exist_ok=True)
    LIBDIR.mkdir(exist_ok=True)
    sources = sorted(CSRC.glob("*.cu"))
    with ThreadPoolExecutor(max_workers=min(len(sources), os.cpu_count() or 1)) as ex:
        objs = list(ex.map(lambda s: _compile(s, force), sources))
    newest = max(o.stat().st_mtime for o in objs)
    if force or not LIB.exists() or LIB.stat().st_mtime < newest:
        cmd = [_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a",
               "-Xcompiler", "-fPIC", "-cudart", "static", "-o", str(LIB), *map(str, objs)]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError("link failed")
    if verbose:
        print(f"[recemb_b200] built {LIB} ({LIB.stat().st_size >> 10} KiB): compiled {len(_compiled)} / "
              f"reused {len(sources) - len(_compiled)} of {len(sources)} translation units"
              + (f" ({', '.join(sorted(_compiled))})" if _compiled else " (source digests unchanged)"))
    return LIB


TORCH_SRC = PKG / "csrc_torch" / "torch_ops.cpp"
TORCH_LIB = LIBDIR / "librecemb_torch_ops.so"


def build_torch_ops(force: bool = False, verbose: bool = True) -> Path:
    """lib/librecemb_torch_ops.so: TORCH_LIBRARY registration of the forward lookups (TorchScript export,
    recommendations_b200/export.py).  Plain g++ against the torch headers; links librecemb_b200.so
    (rpath $ORIGIN) -- the kernels stay behind the C ABI."""
    import torch
    tl = Path(torch.__file__).resolve().parent
    h = hashlib.sha256()
    for p in (TORCH_SRC, PKG.parent / "include" / "recemb_b200.h"):
        h.update(p.read_bytes())
    h.update(torch.__version__.encode())
    stamp = OBJDIR / "torch_ops.sha"
    OBJDIR.mkdir(exist_ok=True)
    if not force and TORCH_LIB.exists() and stamp.exists() and stamp.read_text() == h.hexdigest():
        if verbose:
            print(f"[recemb_b200] {TORCH_LIB.name}: reused (source digest unchanged)")
        return TORCH_LIB
    cmd = [shutil.which("g++") or "g++", "-O2", "-std=c++17", "-fPIC", "-shared",
           f"-D_GLIBCXX_USE_CXX11_ABI={int(torch.compiled_with_cxx11_abi())}",
           f"-I{tl / 'include'}", f"-I{tl / 'include' / 'torch' / 'csrc' / 'api' / 'include'}",
           "-I/usr/local/cuda/include", str(TORCH_SRC), "-o", str(TORCH_LIB),
           f"-L{tl / 'lib'}", "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch", f"-L{LIBDIR}", "-lrecemb_b200",
           "-Wl,-rpath,$ORIGIN"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("g++ failed for torch_ops.cpp")
    stamp.write_text(h.hexdigest())
    if verbose:
        print(f"[recemb_b200] built {TORCH_LIB} ({TORCH_LIB.stat().st_size >> 10} KiB)")
    return TORCH_LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
    build_torch_ops(force="--force" in sys.argv)
