"""GPU (B200): the tensor-core filter (csrc/nda_tc.cuh) against the oracle.

The filter only PRE-SELECTS pairs (periodic embedding -> tcgen05.mma -> sign test); every count and mask
bit must still be the reference's, bit for bit.  These tests keep the tensor-core path on tiny selections
the production heuristic would hand to the packed-FP32 filter (NDA_TC_ALLOW_PADDING=1), attack the two thresholds of the within kernel with pairs
placed a few ulps on either side of the cut-off, and check the routing itself.
"""
import os

import numpy as np
import pytest

from namdanalyzer_b200 import _capi, synth
from namdanalyzer_b200.lib import pylibFuncs as plf
from oracle import port

pytestmark = pytest.mark.gpu
TC = 100


@pytest.fixture()
def force_tc():
    # the library reads NDA_* from the environment once; afterwards nda_set_option is the switch
    with _capi.options(NDA_FILTER="tc", NDA_TC_MAX_FRACTION="1.0", NDA_TC_ALLOW_PADDING="1"):
        yield _capi.load()


def rdf_counts(a, b, cell, same, nb, dr, **kw):
    return plf.py_getRadialNbrCounts(a, b, cell, same, nb, dr, **kw).astype(np.int64)


def within_mask(xyz, refSel, outSel, cell, dist):
    out = np.zeros(xyz.shape[:2], np.int32)
    plf.py_getWithin(xyz, refSel, outSel, out, cell, dist)
    return out


def test_forced_tensor_rdf_same_and_two_selections(force_tc):
    rng = np.random.default_rng(301)
    nF = 3
    cell = rng.uniform(58, 70, (nF, 3)).astype(np.float32)              # orthorhombic, per frame
    a = rng.uniform(-90, 160, (1100, nF, 3)).astype(np.float32)          # unwrapped, several cells away
    b = rng.uniform(-90, 160, (777, nF, 3)).astype(np.float32)
    # range 0.35 A: the fp16 slack of the embedding (2.5 E ~ 0.4 A^2) exceeds the range -> packed-FP32 filter
    for nb, dr, tensor in ((149, 0.1, True), (40, 0.33, True), (30, 0.05, True), (7, 0.05, False)):
        assert np.array_equal(rdf_counts(a, a, cell, 1, nb, dr), port.rdf_counts(a, a, cell, 1, nb, dr).astype(np.int64))
        assert (force_tc.nda_last_filter() == TC) == tensor
        assert np.array_equal(rdf_counts(a, b, cell, 0, nb, dr), port.rdf_counts(a, b, cell, 0, nb, dr).astype(np.int64))
        assert np.array_equal(rdf_counts(b, a, cell, 0, nb, dr), port.rdf_counts(b, a, cell, 0, nb, dr).astype(np.int64))
        assert (force_tc.nda_last_filter() == TC) == tensor


@pytest.mark.parametrize("n", [1, 127, 128, 129, 255, 256, 257, 511, 513, 1025])
def test_forced_tensor_block_boundaries(force_tc, n):
    """selection sizes around the 128-atom streamed block and the 256-atom resident chunk"""
    rng = np.random.default_rng(1000 + n)
    cell = np.array([[61.0, 64.0, 59.0], [62.0, 63.0, 60.5]], np.float32)
    a = rng.uniform(0, 64, (n, 2, 3)).astype(np.float32)
    b = rng.uniform(0, 64, (max(1, (3 * n) // 2), 2, 3)).astype(np.float32)
    assert np.array_equal(rdf_counts(a, a, cell, 1, 100, 0.12), port.rdf_counts(a, a, cell, 1, 100, 0.12).astype(np.int64))
    assert np.array_equal(rdf_counts(a, b, cell, 0, 100, 0.12), port.rdf_counts(a, b, cell, 0, 100, 0.12).astype(np.int64))
    allA = np.concatenate([a, b])
    r = np.arange(0, n, dtype=np.int32)
    o = np.arange(n, allA.shape[0], dtype=np.int32)
    assert np.array_equal(within_mask(allA, r, o, cell, 4.0), port.within(allA, r, o, cell, 4.0))
    assert np.array_equal(within_mask(allA, o, r, cell, 4.0), port.within(allA, o, r, cell, 4.0))
    assert force_tc.nda_last_filter() == TC


def test_forced_tensor_grouped_rdf_config3(force_tc):
    c = synth.config3(n_frames=3)
    w, p, gid, cell = c["waters"][:6000], c["protein"], c["groupId"], c["cellDims"]
    w = np.ascontiguousarray(w)
    got = rdf_counts(w, p, cell, 0, 149, 0.1, groupId=gid, nGroups=76)
    assert force_tc.nda_last_filter() == TC
    ref = np.zeros((76, 149), np.int64)
    for g in range(76):
        sel = np.flatnonzero(gid == g)
        if sel.size:
            ref[g] = port.rdf_counts(w, np.ascontiguousarray(p[sel]), cell, 0, 149, 0.1)
    assert np.array_equal(got.reshape(76, 149), ref)


def _shell_geometry(rng, cell, dist, n_ref, per_ref, rel_eps):
    """ref atoms anywhere (several cells away from the origin), out atoms at distance dist * (1 + eps)
    from 'their' ref atom along random directions, one image away at random: pairs a few ulps on either
    side of the cut-off, through the periodic boundary"""
    ref = rng.uniform(-2.0, 3.0, (n_ref, 3)) * cell
    dirs = rng.normal(size=(n_ref, per_ref, 3))
    dirs /= np.linalg.norm(dirs, axis=-1, keepdims=True)
    eps = rng.choice(rel_eps, size=(n_ref, per_ref, 1))
    out = ref[:, None, :] + dirs * dist * (1.0 + eps)
    out += rng.integers(-1, 2, size=out.shape) * cell                   # another periodic image
    return ref.astype(np.float32), out.reshape(-1, 3).astype(np.float32)


@pytest.mark.parametrize("dist", [1.0, 3.0, 4.75, 9.5])
def test_within_pairs_on_both_sides_of_the_cutoff(dist):
    """the SURE threshold (no strict re-evaluation) and the candidate threshold of the within kernel,
    attacked with pairs at dist * (1 +- k * 2^-24 .. 3e-2): the mask must still be the oracle's"""
    lib = _capi.load()
    rng = np.random.default_rng(int(dist * 100))
    cell3 = np.array([61.3, 64.9, 58.2])
    rel = np.concatenate([[0.0], 2.0 ** -np.arange(5, 25), -(2.0 ** -np.arange(5, 25)), [3e-2, -3e-2, 1e-2, -1e-2]])
    n_ref = int(min(300, max(8, 300 * (3.0 / dist) ** 3)))              # sparse enough that "its" ref atom decides
    frames = []
    for f in range(4):
        ref, out = _shell_geometry(rng, cell3, dist, n_ref, 2400 // n_ref, rel)
        frames.append(np.concatenate([ref, out]))
    xyz = np.ascontiguousarray(np.stack(frames, 1))                     # [atom][frame][3]
    cell = np.tile(cell3.astype(np.float32), (4, 1))
    r = np.arange(0, n_ref, dtype=np.int32)
    o = np.arange(n_ref, xyz.shape[0], dtype=np.int32)
    want = port.within(xyz, r, o, cell, dist)
    _capi.set_option("NDA_TC_ALLOW_PADDING", "1")                       # few ref atoms: keep the tensor path anyway
    try:
        got = within_mask(xyz, r, o, cell, dist)
        assert lib.nda_last_filter() == TC
        assert np.array_equal(got, want)
        frac = want[n_ref:].mean()
        assert 0.2 < frac < 0.8                                         # both sides of the cut-off are populated
        _capi.set_option("NDA_TC_NO_SURE", "1")                         # same masks without the shortcut
        try:
            assert np.array_equal(within_mask(xyz, r, o, cell, dist), want)
        finally:
            _capi.set_option("NDA_TC_NO_SURE", None)
    finally:
        _capi.set_option("NDA_TC_ALLOW_PADDING", None)


def test_within_literal_early_out_below_one_angstrom_on_tensor_path():
    """distance < 1: the reference's per-axis early-outs change the answer (SURVEY B-2); the SURE
    shortcut must be off there and the literal result reproduced"""
    lib = _capi.load()
    rng = np.random.default_rng(77)
    cell3 = np.array([30.0, 31.0, 29.5])
    ref, out = _shell_geometry(rng, cell3, 0.8, 400, 6, np.array([-0.5, -0.3, -0.1, -1e-3, 1e-3, 0.1]))
    xyz = np.ascontiguousarray(np.concatenate([ref, out])[:, None, :])
    cell = cell3.astype(np.float32)[None, :]
    r = np.arange(0, 400, dtype=np.int32)
    o = np.arange(400, xyz.shape[0], dtype=np.int32)
    lit = port.within(xyz, r, o, cell, 0.8, literal_early_out=True)
    pure = port.within(xyz, r, o, cell, 0.8, literal_early_out=False)
    assert (lit != pure).any()
    assert np.array_equal(within_mask(xyz, r, o, cell, 0.8), lit)
    assert lib.nda_last_filter() == TC
    _capi.check(lib.nda_set_within_pure_predicate(1))
    try:
        assert np.array_equal(within_mask(xyz, r, o, cell, 0.8), pure)
    finally:
        _capi.check(lib.nda_set_within_pure_predicate(0))


def test_filter_routing():
    """tensor-core filter for sparse ranges in a real cell; packed-FP32 filter for dense ranges, for the
    'no unit cell' sentinel (100000 A) and when NDA_FILTER asks for it; all with identical results"""
    lib = _capi.load()
    rng = np.random.default_rng(9)
    xyz = rng.uniform(0, 60, (4000, 2, 3)).astype(np.float32)
    r = np.arange(0, 1500, dtype=np.int32)
    o = np.arange(1500, 4000, dtype=np.int32)
    cell = np.full((2, 3), 60.0, np.float32)
    want = port.within(xyz, r, o, cell, 3.0)
    assert np.array_equal(within_mask(xyz, r, o, cell, 3.0), want)
    assert lib.nda_last_filter() == TC
    with _capi.options(NDA_FILTER="fold2"):
        assert np.array_equal(within_mask(xyz, r, o, cell, 3.0), want)
        assert 2 <= lib.nda_last_filter() <= 5
    nocell = np.full((2, 3), 100000.0, np.float32)                      # dcdCell.py:85-89
    assert np.array_equal(within_mask(xyz, r, o, nocell, 3.0), port.within(xyz, r, o, nocell, 3.0))
    assert lib.nda_last_filter() != TC
    a = np.ascontiguousarray(xyz[:1500])
    assert np.array_equal(rdf_counts(a, a, cell, 1, 250, 0.1), port.rdf_counts(a, a, cell, 1, 250, 0.1).astype(np.int64))
    assert lib.nda_last_filter() != TC                                  # 25 A in a 60 A cell: beyond ~0.34 c
    assert np.array_equal(rdf_counts(a, a, cell, 1, 149, 0.1), port.rdf_counts(a, a, cell, 1, 149, 0.1).astype(np.int64))
    assert lib.nda_last_filter() == TC                                  # 6.5 % of the pairs in range: still tensor
    with _capi.options(NDA_TC_MAX_FRACTION="0.01"):                     # ... unless a density cap is asked for
        assert np.array_equal(rdf_counts(a, a, cell, 1, 149, 0.1), port.rdf_counts(a, a, cell, 1, 149, 0.1).astype(np.int64))
        assert lib.nda_last_filter() != TC
    assert np.array_equal(rdf_counts(a, a, cell, 1, 30, 0.1), port.rdf_counts(a, a, cell, 1, 30, 0.1).astype(np.int64))
    assert lib.nda_last_filter() == TC                                  # 0.05 %: sparse
    tiny = np.arange(0, 9, dtype=np.int32)                              # 9 ref atoms would be padded to 256 columns
    assert np.array_equal(within_mask(xyz, tiny, o, cell, 3.0), port.within(xyz, tiny, o, cell, 3.0))
    assert lib.nda_last_filter() != TC
    mixed = cell.copy()
    mixed[1] = 100000.0                                                 # one frame without a cell: whole call falls back
    assert np.array_equal(within_mask(xyz, r, o, mixed, 3.0), port.within(xyz, r, o, mixed, 3.0))
    assert lib.nda_last_filter() != TC


def test_device_api_tensor_path_frame_ranges_add_up():
    """resident trajectory, frame-range calls (the multi-GPU sharding unit) on the tensor-core path"""
    import ctypes
    import torch
    lib = _capi.load()
    c = synth.config2(n_frames=24)
    dev = torch.device("cuda", 0)
    xyz = torch.from_numpy(c["allAtoms"]).to(dev)
    cell = torch.from_numpy(c["cellDims"]).to(dev)
    rs, os_ = torch.from_numpy(c["refSel"]).to(dev), torch.from_numpy(c["outSel"]).to(dev)
    words = (os_.numel() + 31) // 32
    whole = torch.zeros((24, words), dtype=torch.int32, device=dev)
    _capi.check(lib.nda_dev_within(xyz.data_ptr(), xyz.shape[0], 24, rs.data_ptr(), rs.numel(), os_.data_ptr(), os_.numel(),
                                   whole.data_ptr(), None, cell.data_ptr(), 3.0, 0, 24, None))
    assert lib.nda_last_filter() == TC
    parts = torch.zeros((24, words), dtype=torch.int32, device=dev)
    for fb, fe in ((0, 7), (7, 8), (8, 24)):
        _capi.check(lib.nda_dev_within(xyz.data_ptr(), xyz.shape[0], 24, rs.data_ptr(), rs.numel(), os_.data_ptr(), os_.numel(),
                                       ctypes.c_void_p(parts.data_ptr() + fb * words * 4), None, cell.data_ptr(), 3.0, fb, fe, None))
    torch.cuda.synchronize()
    assert torch.equal(whole, parts)
    want = port.within(c["allAtoms"], c["refSel"], c["outSel"], c["cellDims"], 3.0)
    bits = whole.cpu().numpy().view(np.uint32)
    got = ((bits[:, :, None] >> np.arange(32, dtype=np.uint32)) & 1).reshape(24, -1)[:, : os_.numel()]
    assert np.array_equal(got.T.astype(np.int32), want[c["outSel"]])


def test_mixed_exact_and_tensor_frames_in_one_call(force_tc):
    """a runaway coordinate (> 2^20 cells from the origin) sends ITS frame to the all-exact path inside the
    tensor-core kernels, the other frames keep the filter: TMA / MMA warps skip those items, the epilogue
    warps evaluate them strictly, and the barrier protocol must stay consistent"""
    rng = np.random.default_rng(4242)
    nF = 7
    cell = np.tile(np.array([[60.0, 62.0, 58.0]], np.float32), (nF, 1))
    xyz = rng.uniform(0, 60, (1900, nF, 3)).astype(np.float32)
    xyz[5, 2, 0] = 3.0e8                                                # frame 2: ref atom far away
    xyz[1500, 5, 1] = -7.0e8                                            # frame 5: out atom far away
    r = np.arange(0, 700, dtype=np.int32)
    o = np.arange(700, 1900, dtype=np.int32)
    assert np.array_equal(within_mask(xyz, r, o, cell, 3.5), port.within(xyz, r, o, cell, 3.5))
    assert force_tc.nda_last_filter() == TC
    a, b = np.ascontiguousarray(xyz[:700]), np.ascontiguousarray(xyz[700:])
    assert np.array_equal(rdf_counts(a, b, cell, 0, 60, 0.1), port.rdf_counts(a, b, cell, 0, 60, 0.1).astype(np.int64))
    assert np.array_equal(rdf_counts(b, b, cell, 1, 60, 0.1), port.rdf_counts(b, b, cell, 1, 60, 0.1).astype(np.int64))
    assert force_tc.nda_last_filter() == TC


def test_many_small_items_and_single_frame(force_tc):
    """more work items than CTAs with one or two steps each, and a one-frame call: item boundaries, the
    STOP stage and odd step counts per stage"""
    rng = np.random.default_rng(99)
    for nF, n1, n2 in ((1, 130, 3000), (40, 129, 700), (3, 385, 257)):
        cell = np.full((nF, 3), 55.0, np.float32)
        a = rng.uniform(0, 55, (n1, nF, 3)).astype(np.float32)
        b = rng.uniform(0, 55, (n2, nF, 3)).astype(np.float32)
        assert np.array_equal(rdf_counts(a, b, cell, 0, 50, 0.1), port.rdf_counts(a, b, cell, 0, 50, 0.1).astype(np.int64))
        assert force_tc.nda_last_filter() == TC
        allA = np.concatenate([a, b])
        r = np.arange(n1, n1 + n2, dtype=np.int32)
        o = np.arange(0, n1, dtype=np.int32)
        assert np.array_equal(within_mask(allA, r, o, cell, 2.5), port.within(allA, r, o, cell, 2.5))
        assert force_tc.nda_last_filter() == TC


def test_long_ranges_up_to_half_the_cell(force_tc):
    """NDA_TC_RANGE_FACTOR lifts the range limit: the dot-product threshold turns negative, padding atoms
    are kept out by their own K slots, and the counts must still be the oracle's up to ~0.48 c"""
    rng = np.random.default_rng(31)
    cell = np.array([[31.0, 33.0, 30.5], [32.0, 31.5, 31.0]], np.float32)
    a = rng.uniform(-40, 70, (700, 2, 3)).astype(np.float32)
    b = rng.uniform(-40, 70, (300, 2, 3)).astype(np.float32)
    _capi.set_option("NDA_TC_RANGE_FACTOR", "3.9")
    try:
        for nb in (100, 149):                                           # 10 A and 14.9 A in a ~31 A cell
            assert np.array_equal(rdf_counts(a, a, cell, 1, nb, 0.1), port.rdf_counts(a, a, cell, 1, nb, 0.1).astype(np.int64))
            assert force_tc.nda_last_filter() == TC
            assert np.array_equal(rdf_counts(a, b, cell, 0, nb, 0.1), port.rdf_counts(a, b, cell, 0, nb, 0.1).astype(np.int64))
        allA = np.concatenate([a, b])
        r = np.arange(700, 1000, dtype=np.int32)
        o = np.arange(0, 700, dtype=np.int32)
        for dist in (9.0, 14.0):
            assert np.array_equal(within_mask(allA, r, o, cell, dist), port.within(allA, r, o, cell, dist))
            assert force_tc.nda_last_filter() == TC
    finally:
        _capi.set_option("NDA_TC_RANGE_FACTOR", None)
