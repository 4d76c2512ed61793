#!/usr/bin/env python
"""SURVEY row f-3 measurement: setWaterDistPBC / waterOrientAtSurface on the config-2 geometry
(7923 water molecules, 1231 protein atoms), GPU end to end + kernel, and the reference's own CPU
implementation (oracle/_ref/libref_fast.so) on a bounded sample of the same frames."""
import ctypes, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from namdanalyzer_b200 import synth, _capi
from namdanalyzer_b200.lib import pylibFuncs as plf
from oracle import ref

lib = _capi.load()
nF = int(sys.argv[1]) if len(sys.argv) > 1 else 200
c = synth.config2(n_frames=nF)
xyz, cell = c["allAtoms"], c["cellDims"]
prot = np.ascontiguousarray(xyz[:1231])
water = np.ascontiguousarray(xyz[1231:])            # 7923 molecules x 3 atoms, O first
wO = np.ascontiguousarray(water[0::3])
vec = np.ascontiguousarray(((wO - water[1::3]) + (wO - water[2::3])) / np.float32(2))
pairs = 7923.0 * 1231 * nF


def kms():
    v = ctypes.c_double(0); n = ctypes.c_int(0)
    _capi.check(lib.nda_last_kernel_ms(ctypes.byref(v), ctypes.byref(n)))
    return v.value


for name, fn in (("setWaterDistPBC", lambda: plf.py_setWaterDistPBC(water.copy(), prot, cell, 3)),
                 ("waterOrientAtSurface", lambda: plf.py_waterOrientAtSurface(wO.copy(), vec, prot, np.zeros((1, 1), np.float32),
                                                                              cell, 0.0, 5.0, 5))):
    fn(); fn()
    t0 = time.perf_counter()
    for _ in range(3):
        fn()
    dt = (time.perf_counter() - t0) / 3
    k = kms()
    print("%-22s %d frames: e2e %.2f ms (%.3e pairs/s) | kernel %.3f ms (%.3e pairs/s)" % (name, nF, dt * 1e3, pairs / dt, k, pairs / (k * 1e-3)))

if ref.available("fast"):
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    R = ref.RefLib("fast")
    s = min(nF, 4)
    t0 = time.perf_counter()
    R.set_water_dist_pbc_literal(water[:, :s], prot[:, :s], cell[:s], 3)
    dt = time.perf_counter() - t0
    print("reference CPU setWaterDistPBC (serial in the reference), %d frames: %.3e pairs/s" % (s, 7923.0 * 1231 * s / dt))
    t0 = time.perf_counter()
    R.water_orient_at_surface(wO[:, :s], vec[:, :s], prot[:, :s], cell[:s], 0.0, 5.0, 5)
    dt = time.perf_counter() - t0
    print("reference CPU waterOrientAtSurface (%d threads), %d frames: %.3e pairs/s" % (os.cpu_count(), s, 7923.0 * 1231 * s / dt))

# ---- row f-4: H-bond autocorrelation, 4000 water-like triples x 200 frames
from oracle import port
hb = synth.hbond_case(n_water=4000, n_frames=200, L=49.0)
args = (hb["acceptors"], hb["donors"], hb["hydrogens"], hb["cellDims"], hb["maxR"], hb["minAngle"], 1)
plf.py_getHBCounts(*args)
t0 = time.perf_counter()
cnt = plf.py_getHBCounts(*args)
dt = time.perf_counter() - t0
st = (ctypes.c_double * 4)()
lib.nda_last_stats(st)
print("getHBCorr continuous, 4000 x 4000 / 2 pairs at t0, %d bonded, 200 frames: e2e %.2f ms, walk kernel %.3f ms" % (int(st[1]), dt * 1e3, kms()))
t0 = time.perf_counter()
w = port.hb_corr(*args)
print("  C restatement on the CPU (%d threads): %.1f ms; equal: %s" % (port.num_threads(), (time.perf_counter() - t0) * 1e3, np.array_equal(w, cnt)))
