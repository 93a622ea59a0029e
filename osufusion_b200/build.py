"""Builds libosufusion_sm100.so (hand-written sm_100a CUDA kernels behind a C-ABI) in-tree with nvcc.

The library is compiled ONLY for sm_100a (`-gencode arch=compute_100a,code=sm_100a`); there is no
fallback architecture and no CPU path.  nvcc cross-compiles on a GPU-less box.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
LIB_DIR = PKG_DIR / "lib"
LIB_PATH = LIB_DIR / "libosufusion_sm100.so"
INCLUDE = PKG_DIR.parent / "include"

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-I", str(INCLUDE),
]


def _sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def _digest() -> str:
    h = hashlib.sha256()
    for p in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + [INCLUDE / "osufusion_b200.h"]):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build_native(force: bool = False, verbose: bool = False) -> Path:
    """Build (if stale) under an exclusive file lock: under torchrun every rank may get here at once; the first one builds into a
    temporary file and renames it into place atomically, the others wait for the lock and find the library up to date."""
    import fcntl
    LIB_DIR.mkdir(exist_ok=True)
    with open(LIB_DIR / ".build.lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            return _build_locked(force, verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(force: bool, verbose: bool) -> Path:
    stamp = LIB_DIR / "build.stamp"
    digest = _digest()
    if not force and LIB_PATH.exists() and stamp.exists() and stamp.read_text() == digest:
        return LIB_PATH
    if not Path(NVCC).exists():
        raise RuntimeError(f"nvcc not found at {NVCC}; cannot build {LIB_PATH.name}")
    obj_dir = LIB_DIR / "obj"
    obj_dir.mkdir(exist_ok=True)

    def compile_one(src: Path) -> Path:
        obj = obj_dir / (src.stem + ".o")
        cmd = [NVCC, *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas")
            cmd.insert(2, "-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, _sources()))
    tmp = LIB_PATH.with_suffix(f".tmp{os.getpid()}.so")
    cmd = [NVCC, "-shared", "-o", str(tmp), *map(str, objs), "-cudart", "static",
           "-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    os.replace(tmp, LIB_PATH)
    stamp.write_text(digest)
    return LIB_PATH


if __name__ == "__main__":
    p = build_native(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(p)
