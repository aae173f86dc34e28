"""In-tree build of the C-ABI CUDA library ``tair_b200/libtair_b200.so``.

nvcc cross-compiles for sm_100a without a GPU.  Objects are cached under ``tair_b200/csrc/_build`` keyed by a SHA-256 of
the source, every header it can include and the compiler flags; the linked library carries a stamp file
(``libtair_b200.so.stamp``) with the hash of ALL sources.  ``ensure()`` — called by ``_lib.lib()`` before the library is
loaded — compares that stamp with the sources on disk and rebuilds on mismatch, so a prebuilt ``.so`` that travelled to
the GPU box can never silently run stale code.  Run as ``python -m tair_b200.build`` or through
``__graft_entry__.build()``.
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJDIR = os.path.join(CSRC, "_build")
LIB = os.path.join(HERE, "libtair_b200.so")
STAMP = LIB + ".stamp"
HEADER = os.path.join(HERE, "..", "include", "tair_b200.h")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; the tair_b200 CUDA library cannot be built")


def _sources() -> list[str]:
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers() -> list[str]:
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))) + [HEADER]


def _digest(paths) -> str:
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for p in paths:
        h.update(os.path.basename(p).encode())
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def source_hash() -> str:
    """Hash of everything the library is built from (csrc/*.cu, csrc/*.cuh, include/tair_b200.h, flags)."""
    return _digest([os.path.join(CSRC, s) for s in _sources()] + _headers())


def _compile(src: str, verbose: bool) -> str:
    obj = os.path.join(OBJDIR, src[:-3] + ".o")
    spath = os.path.join(CSRC, src)
    want = _digest([spath] + _headers())
    tag = obj + ".sha"
    if os.path.exists(obj) and os.path.exists(tag) and open(tag).read().strip() == want:
        return obj
    cmd = [_nvcc(), *NVCC_FLAGS, "-c", spath, "-o", obj]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{res.stdout}\n{res.stderr}")
    if verbose:
        sys.stderr.write(res.stderr)
    with open(tag, "w") as f:
        f.write(want)
    return obj


def is_current() -> bool:
    return os.path.exists(LIB) and os.path.exists(STAMP) and open(STAMP).read().strip() == source_hash()


def build(verbose: bool = False, force: bool = False) -> str:
    os.makedirs(OBJDIR, exist_ok=True)
    if force:
        for f in os.listdir(OBJDIR):
            os.remove(os.path.join(OBJDIR, f))
    if not force and not verbose and is_current():
        return LIB
    srcs = _sources()
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile(s, verbose), srcs))
    stale = {os.path.join(OBJDIR, f) for f in os.listdir(OBJDIR) if f.endswith(".o")} - set(objs)
    for f in stale:   # object of a deleted source
        os.remove(f)
    cmd = [_nvcc(), "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a",
           "-Xcompiler", "-fPIC", "-cudart", "static"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    with open(STAMP, "w") as f:
        f.write(source_hash())
    return LIB


def ensure() -> str:
    """Library path, rebuilt first when missing or when its stamp does not match the sources on disk."""
    if is_current():
        return LIB
    return build()


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
