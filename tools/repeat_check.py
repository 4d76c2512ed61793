#!/usr/bin/env python
"""Repeat one large within call many times and compare the bitmasks (determinism / race check)."""
import sys, os, ctypes
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from namdanalyzer_b200 import synth, _capi
lib = _capi.load()
dev = torch.device("cuda", 0)
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
nF = int(os.environ.get("NF", "500"))
c = synth.config2(n_frames=nF)
xyz = torch.from_numpy(c["allAtoms"]).to(dev); cell = torch.from_numpy(c["cellDims"]).to(dev)
rs, os_ = torch.from_numpy(c["refSel"]).to(dev), torch.from_numpy(c["outSel"]).to(dev)
words = (os_.numel() + 31) // 32
bits = torch.zeros((nF, words), dtype=torch.int32, device=dev)
ref = None
bad = 0
for i in range(int(os.environ.get("REPS", "200"))):
    _capi.check(lib.nda_dev_within(xyz.data_ptr(), xyz.shape[0], nF, rs.data_ptr(), rs.numel(), os_.data_ptr(), os_.numel(),
                                   bits.data_ptr(), None, cell.data_ptr(), 3.0, 0, nF, st))
    torch.cuda.synchronize()
    h = bits.cpu().numpy().view(np.uint32).copy()
    if ref is None:
        ref = h
        print("filter", lib.nda_last_filter(), "popcount", int(np.unpackbits(ref.view(np.uint8)).sum()))
        continue
    d = ref ^ h
    if d.any():
        bad += 1
        fr, wd = np.nonzero(d)
        print("call %d: %d words differ; frames %s words %s; ref bits %s new bits %s" %
              (i, len(fr), fr[:6].tolist(), wd[:6].tolist(), [hex(int(x)) for x in ref[fr[:3], wd[:3]]], [hex(int(x)) for x in h[fr[:3], wd[:3]]]))
print("calls with differences:", bad)
