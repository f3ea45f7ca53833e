"""Build libcudavec.so in-tree for sm_100a (B200).  `python -m eigensolvers_b200.build`.

nvcc cross-compiles without a GPU; the resulting .so is git-ignored but travels with the
working tree to the GPU box.  No other architecture is built: this library is sm_100a only.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libcudavec.so")
SOURCES = ["cudavec.cu"]
DEPS = ["api.cu", "comm.cu", "peer.cu", "solvers.cu", "common.cuh", "internal.h", "kernels_vec.cuh",
        "kernels_spmv.cuh", "kernels_dia.cuh", "kernels_kron.cuh", "kernels_orth.cuh", "kernels_batch.cuh", os.path.join("..", "..", "include", "cudavec.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xptxas=-v",
    "-Xcompiler", "-fPIC", "-shared",
    "-cudart", "static",
]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    for f in SOURCES + DEPS:
        if os.path.getmtime(os.path.join(CSRC, f)) > t:
            return True
    return False


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cmd = [nvcc] + NVCC_FLAGS + ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES] + ["-ldl"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libcudavec.so (see output above)")
    logdir = os.path.join(os.path.dirname(HERE), "build")   # git-ignored
    os.makedirs(logdir, exist_ok=True)
    with open(os.path.join(logdir, "ptxas.log"), "w") as fh:
        fh.write(res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
