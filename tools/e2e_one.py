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
flags0 = sys.argv[3] if len(sys.argv) > 3 else ""
xyz = c["allAtoms"] if "u" in flags0 else torch.from_numpy(c["allAtoms"]).pin_memory().numpy()   # u: pageable input
out = np.zeros(xyz.shape[:2], np.int32)
flags = sys.argv[3] if len(sys.argv) > 3 else ""     # emulate bench.py conditions: p d s r c
if "p" in flags:
    out = torch.zeros(xyz.shape[:2], dtype=torch.int32).pin_memory().numpy()
if "c" in flags:
    c["cellDims"] = torch.from_numpy(c["cellDims"]).pin_memory().numpy()
keep = []
if "d" in flags:
    keep.append(torch.from_numpy(xyz).to("cuda"))
    keep.append(torch.zeros(xyz.shape[:2], dtype=torch.int32, device="cuda"))
if "s" in flags:
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
if "r" in flags:
    import ctypes
    from namdanalyzer_b200 import _capi
    lib = _capi.load()
    d_xyz = torch.from_numpy(xyz).to("cuda"); d_cell = torch.from_numpy(c["cellDims"]).to("cuda")
    d_rs = torch.from_numpy(c["refSel"]).to("cuda"); d_os = torch.from_numpy(c["outSel"]).to("cuda")
    d_bits = torch.zeros((nF, (c["outSel"].size + 31) // 32), dtype=torch.int32, device="cuda")
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    for _ in range(5):
        _capi.check(lib.nda_dev_within(d_xyz.data_ptr(), xyz.shape[0], nF, d_rs.data_ptr(), c["refSel"].size, d_os.data_ptr(),
                                       c["outSel"].size, d_bits.data_ptr(), None, d_cell.data_ptr(), 3.0, 0, nF, st))
    torch.cuda.synchronize()
for _ in range(4):
    plf.py_getWithin(xyz, c["refSel"], c["outSel"], out, c["cellDims"], 3.0)
ts = []
for _ in range(n):
    t0 = time.perf_counter()
    plf.py_getWithin(xyz, c["refSel"], c["outSel"], out, c["cellDims"], 3.0)
    ts.append((time.perf_counter() - t0) * 1e3)
ts.sort()
print(flags, "frames %d: best %.3f  median %.3f ms/call  (%d set)" % (nF, ts[0], ts[len(ts) // 2], int(out.sum())))
