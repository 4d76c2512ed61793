"""namdanalyzer_b200 -- the minimum-image pair path of kpounot/NAMDAnalyzer (``within`` selections,
``getDistances``, radial number densities) as hand-written sm_100a CUDA behind the reference's own
operator names.

    from namdanalyzer_b200.lib.pylibFuncs import py_getWithin, py_getDistances, \\
        py_getRadialNbrDensity, py_getParallelBackend         # == NAMDAnalyzer.lib.pylibFuncs
    from namdanalyzer_b200.dataAnalysis.RadialDensity import RadialNumberDensity, ResidueWiseWaterDensity
    namdanalyzer_b200.install()    # route an installed NAMDAnalyzer's pair operators through this library

There is no CPU fallback: without ``lib/libnda_b200.so`` or without a B200 every operator raises.
"""
import sys
import types

__version__ = "0.2.0"

_PAIR_OPS = ("py_getWithin", "py_getDistances", "py_getRadialNbrDensity", "py_getParallelBackend",
             "py_setWaterDistPBC", "py_waterOrientAtSurface", "py_getHBCorr", "py_getDCDCoor", "py_getDCDCell")


def pin(array):
    """Page-lock a numpy array in place (``cudaHostRegister``) and return it.  The host operators then DMA
    straight from the array instead of gathering its rows through pinned staging with host threads: about
    1.4x faster end to end on one GPU, and the only way the frames of one call pull at link rate from several
    GPUs at once (``_capi.set_devices``).  Costs about 23 ms per 300 MB once (unpin: 3 ms), so it pays after a few
    dozen calls on one GPU.  Keep the array alive until ``unpin(array)``."""
    from . import _capi
    _capi.check(_capi.load().nda_host_register(array.ctypes.data, array.nbytes))
    return array


def unpin(array):
    from . import _capi
    _capi.check(_capi.load().nda_host_unregister(array.ctypes.data))


def install():
    """Make ``from NAMDAnalyzer.lib.pylibFuncs import py_getWithin, ...`` resolve to this library.

    The reference imports its operators by name at module import time
    (NAMDAnalyzer/dataParsers/dcdParser.py:22-31, dataAnalysis/RadialDensity.py:21-25), so call this
    BEFORE importing NAMDAnalyzer.  If the reference's compiled ``pylibFuncs`` exists its other
    entry points (scattering functions, ...) are kept; the operators in ``_PAIR_OPS`` are replaced: the
    four pair operators, the three SURVEY section 8 f-3 / f-4 operators and the two DCD readers of row
    f-2 (the reference's own reader does not compile outside MSVC).  When ``NAMDAnalyzer.selection.selParser``
    is importable, ``SelParser.process`` is patched as well (``namdanalyzer_b200.selection``): the per-frame
    Python loop behind ``selection('... within ... frame a:b')`` becomes one vectorised decode (row f-1).
    Returns the module object that was installed."""
    from .lib import pylibFuncs as ours
    name = "NAMDAnalyzer.lib.pylibFuncs"
    try:
        import importlib
        mod = importlib.import_module(name)
    except Exception:
        mod = types.ModuleType(name)
        mod.__doc__ = "namdanalyzer_b200 stand-in for the reference's Cython module (pair operators only)"
        if "NAMDAnalyzer" not in sys.modules:
            try:
                import NAMDAnalyzer  # noqa: F401
            except Exception:
                pkg = types.ModuleType("NAMDAnalyzer")
                pkg.__path__ = []
                sys.modules["NAMDAnalyzer"] = pkg
        if "NAMDAnalyzer.lib" not in sys.modules:
            lib = types.ModuleType("NAMDAnalyzer.lib")
            lib.__path__ = []
            sys.modules["NAMDAnalyzer.lib"] = lib
        sys.modules[name] = mod
    for op in _PAIR_OPS:
        setattr(mod, op, getattr(ours, op))
    try:                                                    # row f-1 for the reference's own callers
        import importlib
        from .selection import patch_selparser
        patch_selparser(importlib.import_module("NAMDAnalyzer.selection.selParser").SelParser)
    except Exception:
        pass                                                # reference not importable here: nothing to patch
    return mod
