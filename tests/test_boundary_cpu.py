"""CPU: the drop-in boundary -- the C-ABI library loads and exports every symbol declared in
include/nda_b200.h, the Python shim mirrors NAMDAnalyzer.lib.pylibFuncs' callables and Cython's
argument errors, and the product never computes without its CUDA extension."""
import os
import re
import ctypes
import numpy as np
import pytest

from conftest import ROOT
from namdanalyzer_b200 import _capi
from namdanalyzer_b200.lib import pylibFuncs as plf

HEADER = os.path.join(ROOT, "include", "nda_b200.h")


def declared_symbols():
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(nda_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    syms = declared_symbols()
    assert len(syms) >= 20
    lib = ctypes.CDLL(_capi.LIB_PATH)
    for s in syms:
        assert hasattr(lib, s), "libnda_b200.so lacks %s" % s
    # and the ctypes table binds exactly the header's symbols
    assert sorted(_capi.SIGNATURES) == syms


def test_signatures_are_plain_c():
    txt = open(HEADER).read()
    assert 'extern "C"' in txt
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    for bad in ("torch", "at::", "std::", "Tensor"):
        assert bad not in txt


def test_parallel_backend_is_2_like_the_reference_cuda_build():
    # lib/cuda/src/getParallelBackend.cpp:4-7; routes NAMDDCD.getWithin to the native branch
    assert plf.py_getParallelBackend() == 2


def test_reference_operator_names_exist():
    # the names dataParsers/dcdParser.py:22-31 and dataAnalysis/RadialDensity.py:21-25 import
    for n in ("py_getWithin", "py_getDistances", "py_getRadialNbrDensity", "py_getParallelBackend"):
        assert callable(getattr(plf, n))


def _args():
    xyz = np.zeros((4, 2, 3), np.float32)
    cell = np.ones((2, 3), np.float32)
    return xyz, cell


def test_cython_like_argument_errors():
    xyz, cell = _args()
    out = np.zeros(5, np.float32)
    with pytest.raises(TypeError):
        plf.py_getRadialNbrDensity(None, xyz, out, cell, 0, 1.0, 0.1)
    with pytest.raises(TypeError):
        plf.py_getRadialNbrDensity(xyz.tolist(), xyz, out, cell, 0, 1.0, 0.1)
    with pytest.raises(ValueError, match="dtype mismatch"):
        plf.py_getRadialNbrDensity(xyz.astype(np.float64), xyz, out, cell, 0, 1.0, 0.1)
    with pytest.raises(ValueError, match="wrong number of dimensions"):
        plf.py_getRadialNbrDensity(xyz[:, 0], xyz, out, cell, 0, 1.0, 0.1)
    with pytest.raises(ValueError, match="not C-contiguous"):
        plf.py_getRadialNbrDensity(xyz[:, ::2], xyz, out, cell[:1], 0, 1.0, 0.1)
    with pytest.raises(ValueError, match="dtype mismatch"):
        plf.py_getWithin(xyz, np.arange(2), np.arange(2, dtype=np.int32), np.zeros((4, 2), np.int32), cell, 3.0)
    with pytest.raises(ValueError):
        plf.py_getWithin(xyz, np.arange(2, dtype=np.int32), np.arange(2, dtype=np.int32),
                         np.zeros((4, 3), np.int32), cell, 3.0)          # out shape cross-check
    with pytest.raises(ValueError):
        plf.py_getDistances(xyz, xyz, np.zeros((4, 4), np.float32), np.ones((3, 3), np.float32), 0)


def test_no_cpu_fallback_without_a_gpu():
    """Without a device the product raises; it never routes through oracle/ or numpy."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    xyz, cell = _args()
    with pytest.raises(RuntimeError, match="libnda_b200"):
        plf.py_getRadialNbrDensity(xyz, xyz, np.zeros(5, np.float32), cell, 1, 1.0, 0.1)
    with pytest.raises(RuntimeError, match="libnda_b200"):
        plf.py_getWithin(xyz, np.arange(2, dtype=np.int32), np.arange(4, dtype=np.int32),
                         np.zeros((4, 2), np.int32), cell, 3.0)
    with pytest.raises(RuntimeError, match="libnda_b200"):
        plf.py_getDistances(xyz, xyz, np.zeros((4, 4), np.float32), cell, 0)


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "namdanalyzer_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "minimage_port" not in src and "libref_" not in src, f


def test_options_and_device_list_without_a_gpu():
    """nda_set_option needs no GPU (the switches are plain host state, the environment is read once); an unknown name is
    an error; the device-list calls fail cleanly when there is no device instead of falling back to anything"""
    lib = _capi.load()
    assert lib.nda_set_option(b"NDA_FILTER", b"fold2") == 0
    assert lib.nda_set_option(b"NDA_FILTER", None) == 0
    assert lib.nda_set_option(b"NDA_NO_SUCH_SWITCH", b"1") != 0
    assert b"unknown option" in lib.nda_last_error()
    if lib.nda_device_count() == 0:
        arr = (ctypes.c_int * 2)(0, 1)
        assert lib.nda_set_devices(arr, 2) != 0
        with pytest.raises(RuntimeError):
            _capi.set_devices([0])
