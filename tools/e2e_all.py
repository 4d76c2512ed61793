#!/usr/bin/env python
"""End-to-end wall time of the reference-shaped operators on the five BASELINE configs (host numpy
arrays, pageable unless PIN=1), with the pair-kernel device time of the same call."""
import ctypes, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from namdanalyzer_b200 import synth, _capi
from namdanalyzer_b200.lib import pylibFuncs as plf

lib = _capi.load()
PIN = os.environ.get("PIN", "0") == "1"


def pin(a):
    return torch.from_numpy(a).pin_memory().numpy() if PIN else a


def kms():
    v = ctypes.c_double(0); n = ctypes.c_int(0)
    _capi.check(lib.nda_last_kernel_ms(ctypes.byref(v), ctypes.byref(n)))
    return v.value, n.value


def timeit(fn, reps=5):
    fn(); fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    dt = (time.perf_counter() - t0) / reps
    return dt


def report(name, pairs, dt):
    k, n = kms()
    print("%-44s pairs %.3e  e2e %8.2f ms %.3e pairs/s | pair kernel %8.3f ms (%d launches) %.3e pairs/s"
          % (name, pairs, dt * 1e3, pairs / dt, k, n, pairs / (k * 1e-3) if k else 0), flush=True)


c = synth.config1()
s1 = pin(c["sel1"]); out = np.zeros(149, np.float32)
dt = timeit(lambda: plf.py_getRadialNbrDensity(s1, s1, out, c["cellDims"], 1, 15.0, 0.1))
report("cfg1 RDF O-O 1000 atoms x 100 frames", 1000 * 999 / 2 * 100, dt)

c = synth.config2(n_frames=1000)
x = pin(c["allAtoms"]); o = np.zeros(x.shape[:2], np.int32)
dt = timeit(lambda: plf.py_getWithin(x, c["refSel"], c["outSel"], o, c["cellDims"], 3.0))
report("cfg2 within 3 A, 25000 atoms x 1000 frames", 1231.0 * 7923 * 1000, dt)
del x, o

nF3 = int(os.environ.get("CFG3_FRAMES", "500"))
c = synth.config3(n_frames=nF3)
w, p_ = pin(c["waters"]), pin(c["protein"]); og = np.zeros((76, 149), np.float32)
dt = timeit(lambda: plf.py_getRadialNbrDensityGrouped(w, p_, c["groupId"], og, c["cellDims"], 0, 15.0, 0.1), reps=3)
report("cfg3 grouped RDF 20000 x 1231 x %d frames" % nF3, 20000.0 * 1231 * nF3, dt)
# the reference's call pattern: one call per residue (first 8 residues timed)
def per_res():
    for r in range(8):
        sel = np.ascontiguousarray(c["protein"][c["groupId"] == r])
        o1 = np.zeros(149, np.float32)
        plf.py_getRadialNbrDensity(w, sel, o1, c["cellDims"], 0, 15.0, 0.1)
dt = timeit(per_res, reps=2)
npair = sum(int((c["groupId"] == r).sum()) for r in range(8)) * 20000.0 * nF3
print("%-44s pairs %.3e  e2e %8.2f ms %.3e pairs/s (8 calls, waters re-uploaded each call)" % ("cfg3 call-by-call, residues 0-7", npair, dt * 1e3, npair / dt))
del w, p_

c = synth.config4(n_frames=10000)
a, b = pin(c["sel1"]), pin(c["sel2"]); od = np.zeros((7, 1231), np.float32)
dt = timeit(lambda: plf.py_getDistances(a, b, od, c["cellDims"], 0))
report("cfg4 distances 7 x 1231 x 10000 frames", 7.0 * 1231 * 10000, dt)
print("    cfg4 ingest: %.1f MB -> %.1f GB/s end to end" % ((a.nbytes + b.nbytes) / 1e6, (a.nbytes + b.nbytes) / dt / 1e9))

c = synth.config5(n_side=69, n_frames=2)
s5 = pin(c["sel1"]); out = np.zeros(149, np.float32)
dt = timeit(lambda: plf.py_getRadialNbrDensity(s5, s5, out, c["cellDims"], 1, 15.0, 0.1), reps=2)
n = s5.shape[0]
report("cfg5 RDF O-O %d atoms x 2 frames" % n, 0.5 * n * (n - 1) * 2, dt)
