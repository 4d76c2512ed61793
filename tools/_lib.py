"""Tools only: ``use_variant()`` points the ctypes loader at an experiment build (tools/build_variants.py)
when ``--lib PATH`` is on the command line or NDA_TOOL_LIB is set.  The package itself never looks at either."""
import os
import sys


def use_variant():
    path = os.environ.get("NDA_TOOL_LIB")
    if "--lib" in sys.argv:
        i = sys.argv.index("--lib")
        path = sys.argv[i + 1]
        del sys.argv[i:i + 2]
    if path:
        from namdanalyzer_b200 import _capi
        _capi.LIB_PATH = os.path.abspath(path)
        sys.stderr.write("[tools] using library variant %s\n" % _capi.LIB_PATH)
    return path
