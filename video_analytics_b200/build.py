"""In-tree build of libva_b200.so (hand-written sm_100a CUDA + the C-ABI of include/va_b200.h).

`python -m video_analytics_b200.build` or `__graft_entry__.build()`.  nvcc cross-compiles without a GPU.
The .so is git-ignored but travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
LIB_PATH = PKG_DIR / "libva_b200.so"
BUILD_DIR = PKG_DIR / "build"

SOURCES = ["va_api.cu", "va_conv_tc.cu", "va_conv1_fused.cu", "va_small_kernels.cu", "va_train_kernels.cu", "va_wgrad_tc.cu", "va_jpeg.cu", "va_svm_fit.cu", "va_tvl1.cu", "va_allreduce.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _digest() -> str:
    h = hashlib.sha256()
    for p in sorted(CSRC.glob("*")) + [PKG_DIR.parent / "include" / "va_b200.h"]:
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build_library(force: bool = False, verbose: bool = False) -> Path:
    """Compile every .cu for sm_100a and link libva_b200.so; no-op when sources are unchanged."""
    BUILD_DIR.mkdir(exist_ok=True)
    stamp = BUILD_DIR / "digest.txt"
    digest = _digest()
    if not force and LIB_PATH.exists() and stamp.exists() and stamp.read_text() == digest:
        return LIB_PATH
    nvcc = _nvcc()
    objs = []
    procs = []
    for src in SOURCES:
        obj = BUILD_DIR / (src + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-c", str(CSRC / src), "-o", str(obj)]
        procs.append((src, cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(str(obj))
    log = []
    for src, cmd, pr in procs:
        out, _ = pr.communicate()
        log.append(f"$ {' '.join(cmd)}\n{out}")
        if pr.returncode != 0:
            sys.stderr.write(log[-1])
            raise RuntimeError(f"nvcc failed on {src}")
    link = [nvcc, "-shared", "-o", str(LIB_PATH), *objs, "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a"]
    pr = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    log.append(f"$ {' '.join(link)}\n{pr.stdout}")
    if pr.returncode != 0:
        sys.stderr.write(log[-1])
        raise RuntimeError("link failed")
    (BUILD_DIR / "build.log").write_text("\n".join(log))
    stamp.write_text(digest)
    if verbose:
        print("\n".join(log))
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
