"""Builds hvqm4_b200/libhvqm4_b200.so in-tree with nvcc for sm_100a (no GPU needed to compile)."""
from __future__ import annotations

import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB = os.path.join(_HERE, "libhvqm4_b200.so")
SOURCES = ["recon.cu", "sweep.cu", "row.cu", "rgb.cu", "audio.cu", "entropy_dev.cu", "api.cpp", "player.cpp", "entropy.c"]
HEADERS = ["recon.h", "recon_core.h", "recon_dev.cuh", "sweep_core.h", "row_core.h", "symbuf.h", "entropy.h", "entropy_dev.h", os.path.join("..", "..", "include", "hvqm4.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build_native(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    objs = []
    common = ["-O3", "-lineinfo", "-Xcompiler", "-fPIC,-fvisibility=hidden,-Wall,-Wno-unknown-pragmas"]
    common += os.environ.get("HVQM4_EXTRA_CFLAGS", "").split()   # tuning experiments, e.g. -DHVQM4_BAND_WARPS=16
    for src in SOURCES:
        obj = os.path.join(CSRC, os.path.splitext(src)[0] + ".o")
        cmd = [NVCC] + ARCH + common
        if src.endswith(".c"):
            # host serial stage: plain C, handed to gcc by nvcc
            cmd += ["-x", "c", "-Xcompiler", "-std=gnu11,-O3"]
        else:
            cmd += ["-std=c++17"]
            if src.endswith(".cu") and verbose:
                cmd += ["-Xptxas", "-v"]
        cmd += ["-c", os.path.join(CSRC, src), "-o", obj]
        subprocess.check_call(cmd)
        objs.append(obj)
    subprocess.check_call([NVCC] + ARCH + ["-shared", "-o", LIB] + objs + ["-lpthread"])
    return LIB


if __name__ == "__main__":
    print(build_native(force=True, verbose=True))
