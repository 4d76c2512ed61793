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

__version__ = "0.1.0"

_PAIR_OPS = ("py_getWithin", "py_getDistances", "py_getRadialNbrDensity", "py_getParallelBackend",
             "py_setWaterDistPBC", "py_waterOrientAtSurface", "py_getHBCorr", "py_getDCDCoor", "py_getDCDCell")


def install():
    """Make ``from NAMDAnalyzer.lib.pylibFuncs import py_getWithin, ...`` resolve to this library.

    The reference imports its operators by name at module import time
    (NAMDAnalyzer/dataParsers/dcdParser.py:22-31, dataAnalysis/RadialDensity.py:21-25), so call this
    BEFORE importing NAMDAnalyzer.  If the reference's compiled ``pylibFuncs`` exists its other
    entry points (DCD readers, scattering functions, ...) are kept; only the four pair operators
    are replaced.  Returns the module object that was installed."""
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
    return mod
