#!/usr/bin/env python
"""Experiment builds of the library -- NEVER the product: each variant goes to tools/variants/
libnda_b200_<tag>.so, and a tool has to opt in with ``--lib`` / ``NDA_TOOL_LIB`` (tools/_lib.py).

    python tools/build_variants.py trace            # -DNDA_TC_TRACE_BUILD: clock64 stamps (tools/tc_trace.py)
    python tools/build_variants.py floor1 floor2    # -DNDA_TC_EXPERIMENT=1|2: pipeline floor, WRONG results
    python tools/build_variants.py name:-DFOO=1,-DBAR   # anything else

The experiment variants return wrong results by design (the epilogue skips its reduction / its TMEM load).
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from namdanalyzer_b200 import build as _b  # noqa: E402

KNOWN = {"trace": ["NDA_TC_TRACE_BUILD"], "floor1": ["NDA_TC_EXPERIMENT=1"], "floor2": ["NDA_TC_EXPERIMENT=2"]}
OUT_DIR = os.path.join(ROOT, "tools", "variants")

if __name__ == "__main__":
    for arg in sys.argv[1:]:
        if ":" in arg:
            tag, d = arg.split(":", 1)
            defines = [x[2:] if x.startswith("-D") else x for x in d.split(",") if x]
        else:
            tag, defines = arg, KNOWN[arg]
        print(_b.build_variant(tag, defines, OUT_DIR))
