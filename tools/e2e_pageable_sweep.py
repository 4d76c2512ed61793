#!/usr/bin/env python
"""py_getWithin on config 2 from PAGEABLE numpy arrays: host threads x upload chunks, and what the mask scatter costs."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from namdanalyzer_b200 import synth, _capi
from namdanalyzer_b200.lib import pylibFuncs as plf
c = synth.config2(n_frames=1000)
xyz = c["allAtoms"]
out = np.zeros(xyz.shape[:2], np.int32)
print("cores", os.cpu_count())


def bench(bits_only, n=8):
    def call():
        if bits_only:
            plf.py_getWithinBits(xyz, c["refSel"], c["outSel"], c["cellDims"], 3.0)
        else:
            plf.py_getWithin(xyz, c["refSel"], c["outSel"], out, c["cellDims"], 3.0)
    for _ in range(3):
        call()
    t0 = time.perf_counter()
    for _ in range(n):
        call()
    return (time.perf_counter() - t0) / n * 1e3


for th in ("4", "8", "12", "16", "24", "32"):
    if int(th) > (os.cpu_count() or 1):
        continue
    _capi.set_option("NDA_HOST_THREADS", th)
    row = []
    for ch in (None, "3", "4", "6", "8", "12"):
        _capi.set_option("NDA_WITHIN_CHUNKS", ch)
        row.append("%s:%.2f" % (ch or "auto", bench(False)))
    _capi.set_option("NDA_WITHIN_CHUNKS", None)
    print("threads %2s  ms/call by chunks  %s   bits only (auto) %.2f" % (th, "  ".join(row), bench(True)))
_capi.set_option("NDA_HOST_THREADS", None)
_capi.set_option("NDA_DEBUG", "2")
plf.py_getWithin(xyz, c["refSel"], c["outSel"], out, c["cellDims"], 3.0)
