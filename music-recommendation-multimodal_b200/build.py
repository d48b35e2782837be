"""In-tree build of libtt_b200.so (sm_100a only).

nvcc cross-compiles without a GPU, so this runs in the CPU-only dev container; the
resulting .so sits next to the sources and travels to the GPU box with the repo
snapshot. Objects are cached by mtime under csrc/build/.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
BUILD = CSRC / "build"
LIB = PKG_DIR / "libtt_b200.so"
REPO = PKG_DIR.parent

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def _deps_mtime() -> float:
    hdrs = list(CSRC.glob("*.cuh")) + list((REPO / "include").glob("*.h")) + [Path(__file__)]
    return max(h.stat().st_mtime for h in hdrs)


def _compile(src: Path, hdr_mtime: float, verbose: bool) -> Path:
    obj = BUILD / (src.stem + ".o")
    if obj.exists() and obj.stat().st_mtime > max(src.stat().st_mtime, hdr_mtime):
        return obj
    cmd = [NVCC, *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = BUILD / (src.stem + ".log")
    log.write_text(res.stdout + res.stderr)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError(f"nvcc failed on {src.name}")
    if verbose:
        sys.stderr.write(f"[build] {src.name} ok\n")
    return obj


def build(verbose: bool = False, force: bool = False) -> Path:
    BUILD.mkdir(exist_ok=True)
    srcs = sorted(CSRC.glob("*.cu"))
    hdr_mtime = _deps_mtime()
    if force:
        for o in BUILD.glob("*.o"):
            o.unlink()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile(s, hdr_mtime, verbose), srcs))
    newest = max(o.stat().st_mtime for o in objs)
    if not LIB.exists() or LIB.stat().st_mtime < newest:
        cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a",
               "-o", str(LIB), *map(str, objs)]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError("link failed")
        if verbose:
            sys.stderr.write(f"[build] linked {LIB.name}\n")
    return LIB


if __name__ == "__main__":
    build(verbose=True, force="--force" in sys.argv)
    print(LIB)
