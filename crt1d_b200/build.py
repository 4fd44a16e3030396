"""Build libcrt1d_b200.so in-tree with nvcc for sm_100a (B200).  No torch, no JIT cache: the .so sits
next to the package so it travels with the repo snapshot to the GPU box.

    python -m crt1d_b200.build [--force] [--verbose]
"""
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libcrt1d_b200.so")
SOURCES = ["crt_kernels.cu", "crt_abi.cu"]
HEADERS = ["crt_core.cuh", "crt_scheme.cuh", "crt_leafangle.cuh", "crt_internal.h",
           os.path.join("..", "..", "include", "crt1d_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-DCRT1D_BUILD",  # precise math throughout: fp64 parity, never --use_fast_math
    "-Xcompiler", "-fPIC,-O3,-fvisibility=hidden",
    "-Xptxas", "-v",
    "-shared", "-cudart", "static",
]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc); the CUDA library cannot be built")


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, defines=(), out=None):
    """Compile the shared library; returns its path.  Raises on failure (no fallback).
    `defines` / `out` build tuning variants (e.g. ("CRT_BLOCK=256",)) next to the default library."""
    out = out or LIB_PATH
    if not force and out == LIB_PATH and not needs_build():
        return LIB_PATH
    cmd = [nvcc_path()] + NVCC_FLAGS + [f"-D{d}" for d in defines] + [os.path.join(CSRC, s) for s in SOURCES]
    cmd += ["-o", out + ".tmp"]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or proc.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + proc.stdout + proc.stderr)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed building libcrt1d_b200.so (see output above)")
    os.replace(out + ".tmp", out)
    with open(os.path.join(PKG_DIR, "csrc", "ptxas_info.txt" if out == LIB_PATH else os.path.basename(out) + ".ptxas.txt"), "w") as f:
        f.write(proc.stderr)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
