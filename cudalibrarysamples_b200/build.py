"""Build libb200spmv.so (the sm_100a kernels + the cuSPARSE-symbol shim) in-tree with nvcc.

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.  nvcc cross-compiles for sm_100a
without a GPU, so this runs in the CPU-only build container too.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libb200spmv.so")
SOURCES = ["spmv_csr.cu", "spmv_coo_sell.cu", "workload_gen.cu", "cusparse_shim.cpp"]
HEADERS = ["spmv_common.cuh", os.path.join("..", "..", "include", "b200spmv.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=default",
    "-Xptxas", "-v",
]


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: cannot build libb200spmv.so")
    return exe


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build_native(force: bool = False, verbose: bool = False, extra_flags=(), out_path: str | None = None,
                 tag: str = "") -> str:
    """Compile every CUDA source for sm_100a into cudalibrarysamples_b200/libb200spmv.so.

    extra_flags / out_path / tag build a tuning variant (scripts/sweep.py) next to the default library."""
    variant = bool(extra_flags) or out_path is not None
    if not variant and not force and not needs_build():
        return LIB_PATH
    target = out_path or LIB_PATH
    objs = []
    build_dir = os.path.join(PKG_DIR, "build", tag) if tag else os.path.join(PKG_DIR, "build")
    os.makedirs(build_dir, exist_ok=True)
    log = []
    for s in SOURCES:
        o = os.path.join(build_dir, s + ".o")
        cmd = [nvcc(), *NVCC_FLAGS, *extra_flags, "-c", os.path.join(CSRC, s), "-o", o]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log.append("$ " + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode != 0:
            sys.stderr.write(log[-1])
            raise RuntimeError(f"nvcc failed on {s}")
        objs.append(o)
    cmd = [nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", target, *objs, "-ldl", "-lpthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log.append("$ " + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode != 0:
        sys.stderr.write(log[-1])
        raise RuntimeError("link failed")
    with open(os.path.join(build_dir, "build.log"), "w") as f:
        f.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    return target


if __name__ == "__main__":
    print(build_native(force="--force" in sys.argv, verbose=True))
