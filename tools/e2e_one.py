#!/usr/bin/env python
"""py_getWithin end to end on config 2 (pinned host array): ms per call, best and median of N.
Usage: e2e_one.py [nFrames] [calls]; with NDA_DEBUG=2 the library prints its host phase clock for the last call."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from namdanalyzer_b200 import synth
from namdanalyzer_b200.lib import pylibFuncs as plf
nF = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
c = synth.config2(n_frames=nF)
xyz = torch.from_numpy(c["allAtoms"]).pin_memory().numpy()
out = np.zeros(xyz.shape[:2], np.int32)
for _ in range(4):
    plf.py_getWithin(xyz, c["refSel"], c["outSel"], out, c["cellDims"], 3.0)
ts = []
for _ in range(n):
    t0 = time.perf_counter()
    plf.py_getWithin(xyz, c["refSel"], c["outSel"], out, c["cellDims"], 3.0)
    ts.append((time.perf_counter() - t0) * 1e3)
ts.sort()
print("frames %d: best %.3f  median %.3f ms/call  (%d set)" % (nF, ts[0], ts[len(ts) // 2], int(out.sum())))
