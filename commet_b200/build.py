"""Build recipe: nvcc for sm_100a, everything in-tree.

    python -m commet_b200.build          # library + tools
    commet_b200/lib/libcommet_b200.so    # CUDA kernels + C-ABI (include/commet_b200.h)
    commet_b200/bin/{index_and_search,filter_reads,bvop,extract_reads}   # drop-in executables (+ commet_nxn)
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
LIB = PKG / "lib" / "libcommet_b200.so"
BIN = PKG / "bin"
TOOLS = ("index_and_search", "filter_reads", "bvop", "commet_nxn", "extract_reads", "compare_reads")
HOST_ONLY = ("extract_reads",)          # pure I/O tools: no CUDA library behind them

NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC,-Wall,-Wno-unknown-pragmas", "-shared"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: the CUDA extension cannot be built")


def _stale(target: Path, sources) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(s).stat().st_mtime > t for s in sources)


def build_lib(force: bool = False, verbose: bool = False) -> Path:
    srcs = [CSRC / "capi.cu", *sorted((CSRC / "capi").glob("*.inl")), CSRC / "kernels.cuh", *sorted((CSRC / "kernels").glob("*.cuh")),
            ROOT / "include" / "commet_b200.h"]
    if force or _stale(LIB, srcs):
        LIB.parent.mkdir(parents=True, exist_ok=True)
        cmd = [_nvcc(), *NVCC_FLAGS, "-o", str(LIB), str(CSRC / "capi.cu")]
        if verbose:
            cmd[1:1] = ["-Xptxas", "-v"]
            print(" ".join(cmd))
        subprocess.run(cmd, check=True)
    return LIB


def build_tools(force: bool = False) -> list[Path]:
    """The drop-in executables: plain C++17 hosts over the C-ABI (rpath = ../lib)."""
    build_lib()
    BIN.mkdir(parents=True, exist_ok=True)
    out = []
    hdrs = sorted((CSRC / "host").glob("*.hpp")) + [ROOT / "include" / "commet_b200.h"]
    for t in TOOLS:
        src = CSRC / "tools" / f"{t}.cpp"
        if not src.exists():
            continue
        exe = BIN / t
        if t in HOST_ONLY:
            if force or _stale(exe, [src, *hdrs]):
                subprocess.run(["g++", "-O2", "-std=c++17", "-Wall", "-pthread", "-I", str(CSRC / "host"), "-o", str(exe),
                                str(src), "-lz"], check=True)
        elif force or _stale(exe, [src, *hdrs, LIB]):
            subprocess.run(["g++", "-O2", "-std=c++17", "-Wall", "-I", str(ROOT / "include"), "-I", str(CSRC / "host"),
                            "-o", str(exe), str(src), "-L", str(LIB.parent), "-lcommet_b200",
                            "-Wl,-rpath,$ORIGIN/../lib", "-lz", "-lpthread"], check=True)
        out.append(exe)
    return out


def build_all(force: bool = False, verbose: bool = False):
    build_lib(force, verbose)
    build_tools(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
