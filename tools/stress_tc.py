#!/usr/bin/env python
"""Stress of the tensor-core kernels: many calls with random shapes / frame counts / ranges, checked
against each other (tensor filter vs packed-FP32 filter), to shake out rare hand-over deadlocks or
races.  Run under `timeout`."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from namdanalyzer_b200 import _capi
from namdanalyzer_b200.lib import pylibFuncs as plf
lib = _capi.load()
rng = np.random.default_rng(int(os.environ.get("SEED", "1")))
n_iter = int(os.environ.get("ITERS", "400"))
_capi.set_option("NDA_TC_ALLOW_PADDING", "1")
t0 = time.time()
n_tc = 0
for it in range(n_iter):
    nF = int(rng.integers(1, 40))
    n1, n2 = int(rng.integers(1, 3000)), int(rng.integers(1, 1500))
    L = float(rng.uniform(40, 90))
    cell = (L * (1 + 0.01 * rng.standard_normal((nF, 3)))).astype(np.float32)
    xyz = rng.uniform(-L, 2 * L, (n1 + n2, nF, 3)).astype(np.float32)
    r = np.arange(0, n2, dtype=np.int32)
    o = np.arange(n2, n1 + n2, dtype=np.int32)
    dist = float(rng.uniform(1.0, 0.25 * L))
    res = []
    for filt in ("tc", "fold2"):
        _capi.set_option("NDA_FILTER", filt)
        out = np.zeros(xyz.shape[:2], np.int32)
        plf.py_getWithin(xyz, r, o, out, cell, dist)
        if filt == "tc":
            n_tc += lib.nda_last_filter() == 100
        a, b = np.ascontiguousarray(xyz[n2:]), np.ascontiguousarray(xyz[:n2])
        nb = int(rng.integers(1, 120)) if filt == "tc" else nb
        dr = float(dist / nb) if filt == "tc" else dr
        c1 = plf.py_getRadialNbrCounts(a, b, cell, 0, nb, dr)
        c2 = plf.py_getRadialNbrCounts(a, a, cell, 1, nb, dr)
        res.append((out, c1, c2))
    assert np.array_equal(res[0][0], res[1][0]), "within differs at iteration %d" % it
    assert np.array_equal(res[0][1], res[1][1]) and np.array_equal(res[0][2], res[1][2]), "rdf differs at iteration %d" % it
print("stress ok: %d iterations, %d on the tensor path, %.1f s" % (n_iter, n_tc, time.time() - t0))
