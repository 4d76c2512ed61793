"""Build libnda_b200.so in-tree with nvcc for sm_100a (no other architecture, no JIT cache).

    python -m namdanalyzer_b200.build [--force]
"""
import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
SRC_DIR = os.path.join(_HERE, "csrc")
INC_DIR = os.path.join(os.path.dirname(_HERE), "include")
OUT = os.path.join(_HERE, "lib", "libnda_b200.so")
SOURCES = ["nda_b200.cu"]
DEPS = ["nda_b200.cu", "nda_kernels.cuh", "nda_device.cuh", "nda_surface.cuh", "nda_tc.cuh"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "--shared", "-Xcompiler", "-fPIC", "-Xcompiler", "-O2", "-Xcompiler", "-fopenmp", "-lgomp", "-ldl"]


def _stale():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(SRC_DIR, d) for d in DEPS] + [os.path.join(INC_DIR, "nda_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not _stale():
        return OUT
    nvcc = os.environ.get("NVCC", "nvcc")
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    # Only ONE configuration is ever built here: no environment variable or flag can change what the product
    # library computes.  Trace / pipeline-floor experiment builds live in tools/build_variants.py and are
    # written to tools/variants/ under their own names.
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
          ["-o", OUT] + [os.path.join(SRC_DIR, s) for s in SOURCES]
    subprocess.check_call(cmd)
    return OUT


def build_variant(tag, defines, out_dir):
    """tools/build_variants.py: the same sources with extra -D flags into ``out_dir/libnda_b200_<tag>.so``."""
    nvcc = os.environ.get("NVCC", "nvcc")
    os.makedirs(out_dir, exist_ok=True)
    out = os.path.join(out_dir, "libnda_b200_%s.so" % tag)
    subprocess.check_call([nvcc] + NVCC_FLAGS + ["-D" + d for d in defines] + ["-o", out] +
                          [os.path.join(SRC_DIR, s) for s in SOURCES])
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
