"""Build the INT8 slice-GEMM study library (tools/studies/int8_gemm/libstudy_i8.so) from the CUTLASS
headers vendored with flashinfer.  Cross-compiles without a GPU (about 45 s)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "libstudy_i8.so")


def cutlass_root():
    import flashinfer
    return os.path.join(os.path.dirname(flashinfer.__file__), "data", "cutlass")


def build(force=False):
    srcs = [os.path.join(HERE, "cutlass_i8_gemm.cu"), os.path.join(HERE, "slices.cu")]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) > os.path.getmtime(f) for f in srcs):
        return OUT
    root = cutlass_root()
    cmd = [os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc"), "-gencode", "arch=compute_100a,code=sm_100a",
           "-std=c++17", "-O3", "--expt-relaxed-constexpr", "-lineinfo",
           "-I", os.path.join(root, "include"), "-I", os.path.join(root, "tools", "util", "include"),
           "-Xcompiler", "-fPIC", "-shared", "-o", OUT] + srcs
    subprocess.run(cmd, check=True, stdout=sys.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
