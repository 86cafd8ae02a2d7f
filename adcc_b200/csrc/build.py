"""Build libadccb.so in-tree with nvcc for sm_100a (no GPU needed: cross-compiles)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "libadccb.so")
SOURCES = ["cabi.cu", "dgemm.cu", "tensor_ops.cu", "packed_ops.cu"]
HEADERS = ["common.cuh", os.path.join("..", "..", "include", "adccb.h")]


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(os.path.join(HERE, f)) > t for f in SOURCES + HEADERS)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
           "-Xcompiler", "-fPIC", "-shared", "-o", OUT] + SOURCES + ["-lcudart"]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    # compiler chatter goes to stderr: bench.py must print exactly one JSON line on stdout
    subprocess.run(cmd, cwd=HERE, check=True, stdout=sys.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
