"""Build the sm_100a CUDA library (libsmpl_b200.so) in-tree with nvcc.

The library is plain C ABI (include/smpl_b200.h), loaded with ctypes; there is no torch or pybind dependency in
the native code, so one nvcc invocation per translation unit plus a link step is the whole build.  The .so is
git-ignored but travels to the GPU box with the repository snapshot.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libsmpl_b200.so")
SOURCES = ["api.cu", "pose_kernels.cu", "blend_kernels.cu", "lbs_kernels.cu", "mask_kernels.cu", "seg_kernels.cu",
           "sil_kernels.cu", "tc_gemm.cu", "loss_kernels.cu", "dense_kernels.cu", "render_kernels.cu"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
# No --use_fast_math: sqrt, division and exp must keep their IEEE / documented-ulp behaviour for parity.
NVCC_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; this package has no CPU path and cannot be built without the CUDA toolkit")


def _digest() -> str:
    h = hashlib.sha256()
    files = sorted(os.listdir(CSRC)) + [os.path.join("..", "..", "include", "smpl_b200.h")]
    for name in files:
        path = os.path.join(CSRC, name)
        if os.path.isfile(path):
            h.update(name.encode())
            with open(path, "rb") as f:
                h.update(f.read())
    h.update(" ".join(ARCH + NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu for sm_100a and link libsmpl_b200.so.  Returns the library path."""
    os.makedirs(BUILD, exist_ok=True)
    stamp = os.path.join(BUILD, "digest.txt")
    digest = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == digest:
        return LIB
    nvcc = _nvcc()

    def compile_one(src: str) -> str:
        obj = os.path.join(BUILD, src.replace(".cu", ".o"))
        cmd = [nvcc, *ARCH, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        res = subprocess.run(cmd, capture_output=True, text=True)
        log = os.path.join(BUILD, src.replace(".cu", ".ptxas.log"))
        with open(log, "w") as f:
            f.write(res.stdout + res.stderr)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed on %s:\n%s" % (src, res.stdout + res.stderr))
        if verbose:
            sys.stderr.write(res.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [nvcc, *ARCH, "-shared", "-Xcompiler", "-fPIC", "-o", LIB, *objs]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + res.stdout + res.stderr)
    with open(stamp, "w") as f:
        f.write(digest)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
