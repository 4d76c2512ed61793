#!/usr/bin/env python
"""Wall time of the host entry point py_getWithin on config 2 for several upload chunkings."""
import os, sys, time, subprocess
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    from namdanalyzer_b200 import synth
    from namdanalyzer_b200.lib import pylibFuncs as plf
    c = synth.config2(n_frames=1000)
    pin = os.environ.get("PIN", "1") == "1"
    xyz = torch.from_numpy(c["allAtoms"])
    if pin:
        xyz = xyz.pin_memory()
    xyz = xyz.numpy()
    out = np.zeros(xyz.shape[:2], np.int32)
    bits_only = os.environ.get("BITS", "0") == "1"
    def call():
        if bits_only:
            plf.py_getWithinBits(xyz, c["refSel"], c["outSel"], c["cellDims"], 3.0)
        else:
            plf.py_getWithin(xyz, c["refSel"], c["outSel"], out, c["cellDims"], 3.0)
    for _ in range(3):
        call()
    t0 = time.perf_counter()
    n = 10
    for _ in range(n):
        call()
    dt = (time.perf_counter() - t0) / n
    print("bits_only=%d chunks=%s pinned=%d: %.2f ms/call, %.3e pairs/s" % (bits_only, os.environ.get("NDA_WITHIN_CHUNKS", "auto"), pin, dt * 1e3, 1231 * 7923 * 1000 / dt))
else:
    for pin in ("1", "0"):
        for ch in ("1", "2", "4", "6", "8", "12", "16"):
            env = dict(os.environ, NDA_WITHIN_CHUNKS=ch, PIN=pin)
            subprocess.run([sys.executable, __file__, "child"], env=env)
