"""Build the C-ABI shared library in-tree: csrc/libscl_b200.so (sm_100a only).

    python -m stanford_compression_library_b200.build
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(CSRC, "libscl_b200.so")
SOURCES = ["scl_kernels.cu"]


def _deps():
    """every source the library is built from: all of csrc/ plus the public header"""
    files = [f for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h", ".hpp"))]
    return files + [os.path.join("..", "..", "include", "scl_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-shared",
              "-ccbin", "/usr/bin/g++"]


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in _deps())


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT] + SOURCES
    subprocess.check_call(cmd, cwd=CSRC)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
