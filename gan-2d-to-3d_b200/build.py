"""Builds csrc/g2s_kernels.cu into csrc/libg2s_b200.so for sm_100a (in-tree, so the .so travels to the GPU box).

    python gan-2d-to-3d_b200/build.py [--force]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(HERE)
SOURCES = [os.path.join(CSRC, "g2s_kernels.cu")]
DEPS = SOURCES + [os.path.join(CSRC, "g2s_math.cuh"), os.path.join(CSRC, "g2s_raster.cuh"),
                  os.path.join(CSRC, "g2s_splat.cuh"), os.path.join(CSRC, "g2s_tile.cuh"),
                  os.path.join(CSRC, "g2s_bigface.cuh"), os.path.join(CSRC, "g2s_callers.cuh"),
                  os.path.join(ROOT, "include", "g2s_b200.h")]
LIB = os.path.join(CSRC, "libg2s_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared",
         "-Xcompiler", "-fPIC", "-cudart", "shared"]


def needs_build():
    if not os.path.exists(LIB):
        return True
    if not os.path.exists(NVCC):
        return False          # prebuilt library, no compiler on this machine
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + SOURCES
    subprocess.check_call(cmd)
    return LIB


def build_variant(out, defines):
    """Kernel A/B experiments: the same sources with extra -D flags into `out` (selected at run time with G2S_LIB=...)."""
    cmd = [NVCC] + FLAGS + ["-D" + d for d in defines] + ["-o", out] + SOURCES
    subprocess.check_call(cmd)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
