#!/usr/bin/env python
"""Pipeline timeline of the tensor-core within kernel (CTA 0): NDA_TC_TRACE=1 python tools/tc_trace.py

Per B block g: when each MMA warp got its TMEM half back and had committed, when epilogue warps 0 and 15
saw the half and had loaded it.  Prints intervals in clocks."""
import ctypes, os, sys
import numpy as np
import torch
os.environ["NDA_TC_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from namdanalyzer_b200 import synth, _capi  # noqa: E402
lib = _capi.load()
dev = torch.device("cuda", 0)
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
nF = 200
c = synth.config2(n_frames=nF)
xyz = torch.from_numpy(c["allAtoms"]).to(dev)
cell = torch.from_numpy(c["cellDims"]).to(dev)
rs, os_ = torch.from_numpy(c["refSel"]).to(dev), torch.from_numpy(c["outSel"]).to(dev)
bits = torch.zeros((nF, (os_.numel() + 31) // 32), dtype=torch.int32, device=dev)
for _ in range(2):
    _capi.check(lib.nda_dev_within(xyz.data_ptr(), xyz.shape[0], nF, rs.data_ptr(), rs.numel(), os_.data_ptr(),
                                   os_.numel(), bits.data_ptr(), None, cell.data_ptr(), 3.0, 0, nF, st))
torch.cuda.synchronize()
tr = np.zeros(256 * 12, dtype=np.int64)
_capi.check(lib.nda_debug_tc_trace(tr.ctypes.data_as(ctypes.POINTER(ctypes.c_longlong))))
tr = tr.reshape(256, 12)
t0 = tr[tr > 0].min()
names = ["m0.pass", "m0.done", "m1.pass", "m1.done", "e0.h0.see", "e0.h0.ld", "e0.h1.see", "e0.h1.ld",
         "e15.h0.see", "e15.h0.ld", "e15.h1.see", "e15.h1.ld"]
print("g  " + " ".join("%10s" % n for n in names))
for g in range(40, 64):
    print("%3d " % g + " ".join("%10d" % (v - t0) for v in tr[g]))
d = np.diff(tr[20:200, 0])
print("block period (m0.pass to m0.pass): mean %.0f  min %d  max %d" % (d.mean(), d.min(), d.max()))
for a, b, lbl in [(0, 1, "MMA warp 0: pass tempty -> issued+committed"), (1, 4, "m0 committed -> e0 sees tfull[0]"),
                  (4, 5, "e0: tfull[0] seen -> loaded"), (5, 6, "e0: h0 loaded -> sees tfull[1]"), (6, 7, "e0: tfull[1] seen -> loaded"),
                  (8, 9, "e15: tfull[0] seen -> loaded"), (3, 6, "m1 committed -> e0 sees tfull[1]")]:
    x = tr[20:200, b] - tr[20:200, a]
    print("%-48s mean %6.0f  min %6d  max %6d" % (lbl, x.mean(), x.min(), x.max()))
x = tr[21:200, 0] - tr[20:199, 5]
print("%-48s mean %6.0f  min %6d  max %6d" % ("e0 loaded h0(g) -> m0 passes tempty for g+1", x.mean(), x.min(), x.max()))
x = tr[21:200, 0] - tr[20:199, 9]
print("%-48s mean %6.0f  min %6d  max %6d" % ("e15 loaded h0(g) -> m0 passes tempty for g+1", x.mean(), x.min(), x.max()))
