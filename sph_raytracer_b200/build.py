"""In-tree build of the CUDA extension (libsphrt.so) with nvcc for sm_100a.

The shared library sits next to this file so that it travels with the source tree
(``gpurun`` snapshots) and is never installed into site-packages.  The three translation units
(forward kernels, adjoint kernels, the fused normal step, everything else) are compiled in parallel and linked.
"""
import os
import shutil
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SOURCES = [os.path.join(CSRC, n) for n in ("sphrt_forward.cu", "sphrt_adjoint.cu", "sphrt_api.cu", "sphrt_normal.cu")]
HEADERS = [os.path.join(CSRC, n) for n in ("sphrt_kernels.cuh", "sphrt_walker.cuh", "sphrt_fast.cuh",
                                           "sphrt_coop.cuh", "sphrt_warp.cuh")] + [os.path.join(HERE, "..", "include", "sphrt.h")]
LIB = os.path.join(HERE, "libsphrt.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    # the walker reproduces the reference's float64 rounding: no implicit contraction
    "-fmad=false",
    "-Xcompiler", "-fPIC", "-diag-suppress", "177",
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


def build(force=False, verbose=False, extra_flags=(), out=None):
    """Compile libsphrt.so if it is missing or older than its sources."""
    out = out or LIB
    if not force and out == LIB and not is_stale():
        return LIB
    nvcc = _nvcc()
    with tempfile.TemporaryDirectory(prefix="sphrt_build_") as tmp:
        jobs = []
        for src in SOURCES:
            obj = os.path.join(tmp, os.path.basename(src)[:-3] + ".o")
            cmd = [nvcc, *NVCC_FLAGS, *extra_flags, "-c", "-o", obj, src]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            jobs.append((obj, cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)))
        log = []
        for obj, cmd, proc in jobs:
            so, se = proc.communicate()
            log.append(se)
            if proc.returncode != 0:
                raise RuntimeError("nvcc failed: " + " ".join(cmd) + "\n" + so + se)
        link = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", out, *[j[0] for j in jobs]]
        res = subprocess.run(link, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("link failed:\n" + res.stdout + res.stderr)
    if verbose:
        print("\n".join(log))
    return out


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
