"""Builds libnt_b200.so (CUDA kernels + C-ABI + native engine) in-tree with nvcc for sm_100a.

The .so is git-ignored but travels to the GPU box with the gpurun snapshot. Incremental: a source is
recompiled only when it (or a header) is newer than its object file.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
OBJ = PKG / "_build"
LIB = PKG / "libnt_b200.so"
CLI = PKG / "ntransformer"

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-std=c++17", "-O3", "-lineinfo", "--use_fast_math", "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function",
          "-ccbin", shutil.which("g++") or "g++"]


def _nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: the CUDA extension cannot be built (there is no CPU fallback)")
    return nvcc


def sources() -> list[Path]:
    srcs = sorted(CSRC.glob("*.cu")) + sorted(CSRC.glob("*.cpp"))
    eng = CSRC / "engine"
    if eng.is_dir():
        srcs += sorted(p for p in eng.glob("*.cpp") if p.name != "main.cpp") + sorted(eng.glob("*.cu"))
    return srcs


def _headers_mtime() -> float:
    hs = list(CSRC.rglob("*.h")) + list(CSRC.rglob("*.cuh")) + list((PKG.parent / "include").glob("*.h"))
    return max((h.stat().st_mtime for h in hs), default=0.0)


def _compile(src: Path, hdr_m: float, verbose: bool) -> Path:
    obj = OBJ / (src.relative_to(CSRC).as_posix().replace("/", "__") + ".o")
    if obj.exists() and obj.stat().st_mtime > max(src.stat().st_mtime, hdr_m):
        return obj
    cmd = [_nvcc(), *ARCH, *COMMON, "-I", str(CSRC), "-I", str(PKG.parent / "include"), "-x", "cu", "-c", str(src), "-o", str(obj)]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src}")
    return obj


def build(verbose: bool = False, force: bool = False) -> Path:
    OBJ.mkdir(exist_ok=True)
    if force:
        for o in OBJ.glob("*.o"):
            o.unlink()
    hdr_m = _headers_mtime()
    srcs = sources()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile(s, hdr_m, verbose), srcs))
    newest = max(o.stat().st_mtime for o in objs)
    if not LIB.exists() or LIB.stat().st_mtime < newest:
        cmd = [_nvcc(), "-shared", "-o", str(LIB), *map(str, objs), "-lcudart", "-lpthread"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    main = CSRC / "engine" / "main.cpp"
    if main.exists() and (not CLI.exists() or CLI.stat().st_mtime < max(LIB.stat().st_mtime, main.stat().st_mtime)):
        cmd = [_nvcc(), *COMMON, "-I", str(CSRC), "-I", str(PKG.parent / "include"), "-o", str(CLI), str(main),
               "-L", str(PKG), "-lnt_b200", "-Xlinker", "-rpath=$ORIGIN", "-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("CLI link failed")
    return LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
