import time, numpy as np, sys
sys.path.insert(0, ".")
import namdanalyzer_b200 as nda
from namdanalyzer_b200 import synth, _capi
from namdanalyzer_b200.lib import pylibFuncs as plf
c = synth.config2(n_frames=1000)
xyz = c["allAtoms"]
out = np.zeros(xyz.shape[:2], np.int32)
plf.py_getWithin(xyz, c["refSel"], c["outSel"], out, c["cellDims"], 3.0)
for rep in range(3):
    t0 = time.perf_counter(); nda.pin(xyz); t1 = time.perf_counter(); nda.unpin(xyz); t2 = time.perf_counter()
    print("pin %.1f ms, unpin %.1f ms for %.0f MB" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, xyz.nbytes / 1e6))
