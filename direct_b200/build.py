"""Build recipe of libdirect_ddp_b200.so (hand-written CUDA for sm_100a + the C-ABI), in-tree."""
from __future__ import annotations

import glob
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libdirect_ddp_b200.so")
SOURCES = [os.path.join(HERE, "csrc", "direct_ddp.cu"), os.path.join(HERE, "host", "corridor_replay.cpp")]
DEPS = SOURCES + sorted(glob.glob(os.path.join(HERE, "csrc", "*.h")) + glob.glob(os.path.join(HERE, "csrc", "*.cuh")) +
                        glob.glob(os.path.join(HERE, "..", "include", "*.h")))
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--shared", "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def nvcc_path() -> str:
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found; libdirect_ddp_b200.so cannot be built")
    return p


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA extension for sm_100a (cross-compiles without a GPU). Returns the .so path."""
    if force or is_stale():
        cmd = [nvcc_path()] + NVCC_FLAGS + ["-o", LIB] + SOURCES
        res = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or res.returncode != 0:
            print(res.stdout)
            print(res.stderr)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed building libdirect_ddp_b200.so")
        with open(os.path.join(HERE, "ptxas_info.txt"), "w") as f:
            f.write(res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
