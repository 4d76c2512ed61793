#!/usr/bin/env python
"""Pipeline timeline of the tensor-core within kernel (CTA 0).  Needs a trace build of the library:

    python tools/build_variants.py trace && python tools/tc_trace.py --lib tools/variants/libnda_b200_trace.so


Per step g: when the MMA warp got its TMEM buffer back and had committed, when epilogue warps 0 and 15
saw the buffer and had loaded it.  Prints intervals in clocks."""
import ctypes, os, sys
import numpy as np
import torch
os.environ["NDA_TC_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from namdanalyzer_b200 import synth, _capi  # noqa: E402
from tools import _lib  # noqa: E402
_lib.use_variant()
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
tr = np.zeros(256 * 32, dtype=np.int64)
_capi.check(lib.nda_debug_tc_trace(tr.ctypes.data_as(ctypes.POINTER(ctypes.c_longlong))))
tr = tr.reshape(256, 32)
if os.path.isdir("gpurun_out"):
    np.save("gpurun_out/trace_raw_%s.npy" % os.environ.get("TRACE_TAG", "x"), tr)
t0 = tr[tr > 0].min()
names = {0: "mma.pass", 1: "mma.done", 4: "e0.see", 5: "e0.ld", 8: "e15.see", 9: "e15.ld"}
cols = sorted(names)
print("g   " + " ".join("%10s" % names[k] for k in cols))
for g in range(64, 96):
    print("%3d " % g + " ".join("%10d" % (tr[g][k] - t0) for k in cols))
d = np.diff(tr[20:250, 0])
print("step period (mma.pass to mma.pass): mean %.0f  median %.0f  min %d  max %d" % (d.mean(), np.median(d), d.min(), d.max()))
for a, b, lbl in [(0, 1, "MMA warp: buffer back -> issued + committed"), (1, 4, "committed -> e0 sees tfull"),
                  (4, 5, "e0: tfull seen -> loaded"), (8, 9, "e15: tfull seen -> loaded")]:
    x = tr[20:250, b] - tr[20:250, a]
    print("%-48s mean %6.0f  median %6.0f  min %6d  max %6d" % (lbl, x.mean(), np.median(x), x.min(), x.max()))
for k, lbl in [(5, "e0"), (9, "e15")]:
    x = tr[22:250, 0] - tr[20:248, k]
    print("%-48s mean %6.0f  median %6.0f  min %6d  max %6d" % (lbl + " loaded step g -> MMA warp gets buffer for g+2", x.mean(), np.median(x), x.min(), x.max()))
x = tr[21:250, 4] - tr[20:249, 5]
print("%-48s mean %6.0f  median %6.0f  min %6d  max %6d" % ("e0: loaded step g -> sees step g+1", x.mean(), np.median(x), x.min(), x.max()))

d = np.diff(tr[8:250, 0])
print("histogram of step periods (clocks):", np.histogram(d, bins=[0, 300, 400, 500, 600, 800, 1000, 1500, 2500, 10000])[0].tolist(),
      "edges 0,300,400,500,600,800,1000,1500,2500,10000")
big = np.flatnonzero(d > 1200) + 8
print("steps followed by a gap > 1200 clocks:", big.tolist())

ld = tr[30:250, 16:32]                                      # every epilogue warp: when it had loaded step g
first, last = ld.min(1), ld.max(1)
print("spread of the 16 warps' loads per step: mean %.0f  median %.0f  max %d" % ((last - first).mean(), np.median(last - first), (last - first).max()))
print("slowest warp per step (histogram over warps 0..15):", np.bincount(ld.argmax(1), minlength=16).tolist())
x = tr[32:250, 0] - last[:218]
print("last warp loaded step g -> MMA warp gets the buffer for g+2: mean %.0f  median %.0f  min %d" % (x.mean(), np.median(x), x.min()))
x = first - tr[30:250, 1]
print("MMA committed -> first warp has loaded: mean %.0f  median %.0f" % (x.mean(), np.median(x)))

# where the MMA warp spends its step: stage wait (first block of a stage only), TMEM-buffer wait, issue
sw = tr[20:250, 3] - tr[20:250, 2]
has = tr[20:250, 2] > 0
print("MMA warp, steps that open a stage (%d of %d): wait for the stage  mean %.0f  median %.0f  max %d" %
      (has.sum(), has.size, sw[has].mean(), np.median(sw[has]), sw[has].max()))
bw = tr[20:250, 0] - tr[20:250, 6]
print("MMA warp: wait for the TMEM buffer (incl. fence)                 mean %.0f  median %.0f  min %d  max %d" % (bw.mean(), np.median(bw), bw.min(), bw.max()))
ov = tr[21:250, 6] - tr[20:249, 1]
print("MMA warp: committed step g -> at the buffer wait of step g+1     mean %.0f  median %.0f  min %d  max %d" % (ov.mean(), np.median(ov), ov.min(), ov.max()))
