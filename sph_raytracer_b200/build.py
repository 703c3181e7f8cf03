"""In-tree build of the CUDA extension (libsphrt.so) with nvcc for sm_100a.

The shared library sits next to this file so that it travels with the source tree
(``gpurun`` snapshots) and is never installed into site-packages.
"""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SOURCES = [os.path.join(HERE, "csrc", "sphrt_kernels.cu")]
HEADERS = [os.path.join(HERE, "csrc", "sphrt_walker.cuh"), os.path.join(HERE, "csrc", "sphrt_fast.cuh"),
           os.path.join(HERE, "csrc", "sphrt_coop.cuh"),
           os.path.join(HERE, "..", "include", "sphrt.h")]
LIB = os.path.join(HERE, "libsphrt.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    # the walker reproduces the reference's float64 rounding: no implicit contraction
    "-fmad=false",
    "-Xcompiler", "-fPIC", "-shared", "-diag-suppress", "177",
]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; the CUDA extension cannot be built")
    return exe


def is_stale():
    if not os.path.exists(LIB):
        return True
    built = os.path.getmtime(LIB)
    return any(os.path.getmtime(p) > built for p in SOURCES + HEADERS)


def build(force=False, verbose=False):
    """Compile libsphrt.so if it is missing or older than its sources."""
    if not force and not is_stale():
        return LIB
    cmd = [_nvcc(), *NVCC_FLAGS, "-o", LIB, *SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
