#!/usr/bin/env python
"""Short, deterministic run of the three pair kernels for ncu (one launch each after a warm-up).

    python tools/profile_run.py [within|rdf|rdf3|dist|all] [frames] [--lib tools/variants/libnda_b200_<tag>.so]
"""
import ctypes
import sys
import numpy as np
import torch

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from namdanalyzer_b200 import synth, _capi  # noqa: E402
from tools import _lib  # noqa: E402

_lib.use_variant()                     # --lib PATH: an experiment build of tools/build_variants.py

what = sys.argv[1] if len(sys.argv) > 1 else "all"
WITHIN_FRAMES = int(sys.argv[2]) if len(sys.argv) > 2 else 200
lib = _capi.load()
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev)            # explicit: the library treats a NULL stream as "use my own"
torch.cuda.set_stream(stream)
st = ctypes.c_void_p(stream.cuda_stream)


def ms():
    v = ctypes.c_double(0)
    _capi.check(lib.nda_last_kernel_ms(ctypes.byref(v), None))
    return v.value


if what in ("within", "all"):
    nF = WITHIN_FRAMES
    c = synth.config2(n_frames=nF)
    xyz = torch.from_numpy(c["allAtoms"]).to(dev)
    cell = torch.from_numpy(c["cellDims"]).to(dev)
    rs, os_ = torch.from_numpy(c["refSel"]).to(dev), torch.from_numpy(c["outSel"]).to(dev)
    bits = torch.zeros((nF, (os_.numel() + 31) // 32), dtype=torch.int32, device=dev)
    for _ in range(2):
        _capi.check(lib.nda_dev_within(xyz.data_ptr(), xyz.shape[0], nF, rs.data_ptr(), rs.numel(), os_.data_ptr(),
                                       os_.numel(), bits.data_ptr(), None, cell.data_ptr(), 3.0, 0, nF, st))
        t = ms()
    pairs = 1231.0 * 7923.0 * nF
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(40):                                   # whole resident step: pack + prep + pair kernel
        _capi.check(lib.nda_dev_within(xyz.data_ptr(), xyz.shape[0], nF, rs.data_ptr(), rs.numel(), os_.data_ptr(),
                                       os_.numel(), bits.data_ptr(), None, cell.data_ptr(), 3.0, 0, nF, st))
    e1.record(stream)
    torch.cuda.synchronize()
    print("within: %.3f ms, %.3e pairs/s; whole step %.4f ms" % (t, pairs / t * 1e3, e0.elapsed_time(e1) / 40))

if what in ("rdf", "all"):
    n_side, nF = 48, 2
    n = n_side ** 3
    L = float(np.float32((n / 0.0334) ** (1.0 / 3.0)))
    xyz = torch.empty((n, nF, 3), dtype=torch.float32, device=dev)
    cell = torch.empty((nF, 3), dtype=torch.float32, device=dev)
    _capi.check(lib.nda_dev_synth_water_box(xyz.data_ptr(), cell.data_ptr(), n_side, nF, 0, nF, L, 0.6, 1005, st))
    counts = torch.zeros(149, dtype=torch.int64, device=dev)
    for _ in range(2):
        counts.zero_()
        _capi.check(lib.nda_dev_radial_nbr_counts(xyz.data_ptr(), n, xyz.data_ptr(), n, None, 1, counts.data_ptr(), 149,
                                                  cell.data_ptr(), nF, 0, nF, 1, 0.1, st))
        t = ms()
    pairs = 0.5 * n * (n - 1.0) * nF
    st4 = (ctypes.c_double * 4)()
    print("rdf: %.3f ms, %.3e pairs/s, in-range %d" % (t, pairs / t * 1e3, int(counts.sum().item())))

if what in ("rdf3", "all"):
    nF = 60
    c = synth.config3(n_frames=nF)
    w_, p_ = torch.from_numpy(c["waters"]).to(dev), torch.from_numpy(c["protein"]).to(dev)
    gid = torch.from_numpy(c["groupId"]).to(dev)
    cell = torch.from_numpy(c["cellDims"]).to(dev)
    counts = torch.zeros((76, 149), dtype=torch.int64, device=dev)
    for _ in range(2):
        counts.zero_()
        _capi.check(lib.nda_dev_radial_nbr_counts(w_.data_ptr(), w_.shape[0], p_.data_ptr(), p_.shape[0], gid.data_ptr(), 76,
                                                  counts.data_ptr(), 149, cell.data_ptr(), nF, 0, nF, 0, 0.1, st))
        t = ms()
    pairs = 20000.0 * 1231 * nF
    print("rdf3 (grouped, config 3): %.3f ms, %.3e pairs/s, in-range %d" % (t, pairs / t * 1e3, int(counts.sum().item())))

if what in ("dist", "all"):
    nF = 10000                                      # config 4 at full size
    c = synth.config4(n_frames=nF)
    a, b = torch.from_numpy(c["sel1"]).to(dev), torch.from_numpy(c["sel2"]).to(dev)
    cell = torch.from_numpy(c["cellDims"]).to(dev)
    s = torch.zeros((7, 1231), dtype=torch.float64, device=dev)
    for _ in range(2):
        s.zero_()
        _capi.check(lib.nda_dev_distances_sum(a.data_ptr(), 7, b.data_ptr(), 1231, s.data_ptr(), cell.data_ptr(),
                                              nF, 0, nF, 0, st))
        t = ms()
    print("dist: %.3f ms, %.3e pairs/s, %.1f GB/s" % (t, 7 * 1231.0 * nF / t * 1e3, (7 + 1231) * nF * 12 / t / 1e6))
torch.cuda.synchronize()
