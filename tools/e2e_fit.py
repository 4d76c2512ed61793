#!/usr/bin/env python
"""py_getWithin end to end (pinned host arrays) for several frame counts: fixed cost vs per-frame cost."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from namdanalyzer_b200 import synth
from namdanalyzer_b200.lib import pylibFuncs as plf
res = []
for nF in (125, 250, 500, 1000, 2000):
    c = synth.config2(n_frames=nF)
    xyz = torch.from_numpy(c["allAtoms"]).pin_memory().numpy()
    out = np.zeros(xyz.shape[:2], np.int32)
    for _ in range(3):
        plf.py_getWithin(xyz, c["refSel"], c["outSel"], out, c["cellDims"], 3.0)
    n = 10
    t0 = time.perf_counter()
    for _ in range(n):
        plf.py_getWithin(xyz, c["refSel"], c["outSel"], out, c["cellDims"], 3.0)
    dt = (time.perf_counter() - t0) / n
    mb = np.union1d(c["refSel"], c["outSel"]).size * nF * 12 / 1e6
    res.append((nF, dt, mb))
    print("frames %5d: %.3f ms/call  %.1f MB needed rows  -> %.1f GB/s overall" % (nF, dt * 1e3, mb, mb / dt / 1e3))
(n1, t1, m1), (n2, t2, m2) = res[-2], res[-1]
b = (t2 - t1) / (m2 - m1)
print("marginal rate %.1f GB/s, fixed cost %.3f ms" % (1e-3 / b, (t1 - b * m1) * 1e3))
