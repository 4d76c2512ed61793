#!/usr/bin/env python
"""Tensor-core filter vs packed-FP32 filter for LONG ranges (range / cell edge up to ~0.48):
RDF O-O, 149 bins of 0.1 A = 14.9 A, water boxes of decreasing size (sameSel)."""
import ctypes, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from namdanalyzer_b200 import _capi
lib = _capi.load()
dev = torch.device("cuda", 0)
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for n_side, nF in ((10, 100), (12, 100), (14, 60), (17, 40), (22, 20)):
    n = n_side ** 3
    L = float(np.float32((n / 0.0334) ** (1.0 / 3.0)))
    xyz = torch.empty((n, nF, 3), dtype=torch.float32, device=dev)
    cell = torch.empty((nF, 3), dtype=torch.float32, device=dev)
    _capi.check(lib.nda_dev_synth_water_box(xyz.data_ptr(), cell.data_ptr(), n_side, nF, 0, nF, L, 0.6, 1005, st))
    pairs = 0.5 * n * (n - 1.0) * nF
    row = []
    for filt in ("tc", "fold2"):
        _capi.set_option("NDA_FILTER", filt)
        counts = torch.zeros(149, dtype=torch.int64, device=dev)
        for _ in range(3):
            counts.zero_()
            _capi.check(lib.nda_dev_radial_nbr_counts(xyz.data_ptr(), n, xyz.data_ptr(), n, None, 1, counts.data_ptr(), 149,
                                                      cell.data_ptr(), nF, 0, nF, 1, 0.1, st))
            ms = ctypes.c_double(0)
            _capi.check(lib.nda_last_kernel_ms(ctypes.byref(ms), None))
        row.append((lib.nda_last_filter(), ms.value, int(counts.sum().item())))
    print("n %6d  L %5.1f  range/L %.2f  in-range %.3f | tensor(%d) %.3f ms %.2e pairs/s | fp32(%d) %.3f ms %.2e pairs/s | equal %s"
          % (n, L, 14.9 / L, row[0][2] / pairs, row[0][0], row[0][1], pairs / row[0][1] * 1e3, row[1][0], row[1][1], pairs / row[1][1] * 1e3,
             row[0][2] == row[1][2]))
