#!/usr/bin/env python
"""Tensor-core filter vs packed-FP32 filter as a function of the fraction of pairs in range
(config-3 geometry: 20000 waters x 1231 protein atoms, 85 A cell, 60 frames, grouped histogram)."""
import ctypes, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from namdanalyzer_b200 import synth, _capi
lib = _capi.load()
dev = torch.device("cuda", 0)
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
nF = 60
c = synth.config3(n_frames=nF)
w_, p_ = torch.from_numpy(c["waters"]).to(dev), torch.from_numpy(c["protein"]).to(dev)
gid = torch.from_numpy(c["groupId"]).to(dev)
cell = torch.from_numpy(c["cellDims"]).to(dev)
pairs = 20000.0 * 1231 * nF
for nb in (30, 60, 100, 149, 200, 230):
    row = []
    for filt, frac in (("tc", "1.0"), ("fold2", None)):
        _capi.set_option("NDA_FILTER", filt)
        if frac:
            _capi.set_option("NDA_TC_MAX_FRACTION", frac)
        counts = torch.zeros((76, nb), dtype=torch.int64, device=dev)
        for _ in range(2):
            counts.zero_()
            _capi.check(lib.nda_dev_radial_nbr_counts(w_.data_ptr(), w_.shape[0], p_.data_ptr(), p_.shape[0], gid.data_ptr(), 76,
                                                      counts.data_ptr(), nb, cell.data_ptr(), nF, 0, nF, 0, 0.1, st))
            ms = ctypes.c_double(0)
            _capi.check(lib.nda_last_kernel_ms(ctypes.byref(ms), None))
        row.append((lib.nda_last_filter(), ms.value, int(counts.sum().item())))
    vol = 4.18879 * (nb * 0.1) ** 3 / 85.0 ** 3
    print("range %4.1f A  sphere/cell %.4f  in-range %.4f | tensor(%d) %.3f ms %.2e pairs/s | fp32(%d) %.3f ms %.2e pairs/s | counts equal %s"
          % (nb * 0.1, vol, row[0][2] / pairs, row[0][0], row[0][1], pairs / row[0][1] * 1e3, row[1][0], row[1][1], pairs / row[1][1] * 1e3,
             row[0][2] == row[1][2]))
