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
              "--shared", "-Xcompiler", "-fPIC", "-Xcompiler", "-O2", "-Xcompiler", "-fopenmp", "-lgomp"]


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
    extra = ["-DNDA_TC_TRACE_BUILD"] if os.environ.get("NDA_TC_TRACE_BUILD") == "1" else []   # tools/tc_trace.py
    if os.environ.get("NDA_TC_EXPERIMENT"):       # pipeline-floor experiments (1: no reduction, 2: no TMEM load); WRONG results
        extra.append("-DNDA_TC_EXPERIMENT=" + os.environ["NDA_TC_EXPERIMENT"])
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + \
          ["-o", OUT] + [os.path.join(SRC_DIR, s) for s in SOURCES]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
