"""GPU: SURVEY row f-3 -- setWaterDistPBC and waterOrientAtSurface through the reference-shaped
operators, bit-for-bit against the C restatement (which tests/test_oracle.py pins to the compiled
reference sources)."""
import numpy as np
import pytest

from namdanalyzer_b200 import synth
from namdanalyzer_b200.lib import pylibFuncs as plf
from oracle import port

pytestmark = pytest.mark.gpu


def surface_case(seed, nW, nP, nF, L=24.0, spread=(-10, 34)):
    rng = np.random.default_rng(seed)
    cell = (L * (1 + 0.01 * rng.standard_normal((nF, 3)))).astype(np.float32)
    prot = rng.uniform(L / 4, 3 * L / 4, (nP, nF, 3)).astype(np.float32)
    water = rng.uniform(spread[0], spread[1], (3 * nW, nF, 3)).astype(np.float32)
    water[1::3] = water[0::3] + rng.normal(0, 0.6, (nW, nF, 3)).astype(np.float32)
    water[2::3] = water[0::3] + rng.normal(0, 0.6, (nW, nF, 3)).astype(np.float32)
    return water, prot, cell


@pytest.mark.parametrize("nW,nP,nF", [(1, 1, 1), (60, 200, 3), (130, 1025, 2), (500, 2049, 1), (129, 33, 40)])
def test_set_water_dist_pbc(nW, nP, nF):
    water, prot, cell = surface_case(100 + nW, nW, nP, nF)
    want = port.set_water_dist_pbc(water, prot, cell, 3, reset_per_water=True)
    got = water.copy()
    plf.py_setWaterDistPBC(got, prot, cell, 3)
    assert np.array_equal(got, want)
    assert not np.array_equal(got, water) or nW == 1
    # single-atom "molecules" (nbrWAtoms = 1)
    w1 = np.ascontiguousarray(water[0::3])
    want1 = port.set_water_dist_pbc(w1, prot, cell, 1, reset_per_water=True)
    plf.py_setWaterDistPBC(w1, prot, cell, 1)
    assert np.array_equal(w1, want1)


@pytest.mark.parametrize("maxN", [1, 2, 5, 16])
@pytest.mark.parametrize("minR,maxR", [(0.0, 5.0), (2.0, 4.0)])
def test_water_orient_at_surface(maxN, minR, maxR):
    water, prot, cell = surface_case(200 + maxN, 300, 1100, 2)
    wO = np.ascontiguousarray(water[0::3])
    vec = np.ascontiguousarray(((wO - water[1::3]) + (wO - water[2::3])) / np.float32(2))
    with np.errstate(all="ignore"):
        want = port.water_orient_at_surface(wO, vec, prot, cell, minR, maxR, maxN)
    got = wO.copy()
    plf.py_waterOrientAtSurface(got, vec, prot, np.zeros((1, 1), np.float32), cell, minR, maxR, maxN)
    assert np.array_equal(got, want, equal_nan=True)
    assert 0 < want[..., 1].sum() < want[..., 1].size


def test_water_orient_real_geometry_no_unit_cell():
    """the reference's own test trajectory (ubiquitin in a water sphere, cell = 1e5)"""
    fx = synth.fixture()
    xyz, cell = fx["all_xyz"], fx["all_cell"]
    wO = np.ascontiguousarray(xyz[fx["water_o_idx"]])
    h1 = np.ascontiguousarray(xyz[fx["water_o_idx"] + 1])
    h2 = np.ascontiguousarray(xyz[fx["water_o_idx"] + 2])
    vec = np.ascontiguousarray(((wO - h1) + (wO - h2)) / np.float32(2))
    prot = np.ascontiguousarray(xyz[fx["protein_idx"]])
    with np.errstate(all="ignore"):
        want = port.water_orient_at_surface(wO, vec, prot, cell, 0.0, 5.0, 5)
    got = wO.copy()
    plf.py_waterOrientAtSurface(got, vec, prot, np.zeros((1, 1), np.float32), cell, 0.0, 5.0, 5)
    assert np.array_equal(got, want, equal_nan=True)
    moved = wO.copy()
    plf.py_setWaterDistPBC(moved, prot, cell, 1)
    assert np.array_equal(moved, wO)                        # no periodic box: nothing moves


def test_surface_argument_errors():
    w = np.zeros((6, 2, 3), np.float32)
    p = np.zeros((4, 2, 3), np.float32)
    cell = np.ones((2, 3), np.float32)
    with pytest.raises(ValueError):
        plf.py_setWaterDistPBC(w, p, cell, 4)               # 6 atoms are not a multiple of 4
    with pytest.raises(RuntimeError, match="maxN"):
        plf.py_waterOrientAtSurface(w, w.copy(), p, np.zeros((1, 1), np.float32), cell, 0.0, 5.0, 17)
    with pytest.raises(ValueError):
        plf.py_waterOrientAtSurface(w, w[:3].copy(), p, np.zeros((1, 1), np.float32), cell, 0.0, 5.0, 3)


# ---------------------------------------------------------------- SURVEY row f-4 (H-bond autocorrelation)

@pytest.mark.parametrize("continuous", [0, 1])
@pytest.mark.parametrize("n,nF", [(400, 30), (1500, 70), (37, 33)])
def test_hb_corr(n, nF, continuous):
    c = synth.hbond_case(n_water=n, n_frames=nF, seed=1006 + n)
    args = (c["acceptors"], c["donors"], c["hydrogens"], c["cellDims"], c["maxR"], c["minAngle"], continuous)
    want = port.hb_corr(*args)
    got = plf.py_getHBCounts(*args)
    assert np.array_equal(got, want)
    # the reference-shaped operator accumulates into out (getHydrogenBonds.cpp:73)
    out = np.full(nF, 2.0, np.float32)
    plf.py_getHBCorr(c["acceptors"], nF, c["donors"], c["hydrogens"], c["cellDims"], out, 100, 1, 25,
                     c["maxR"], c["minAngle"], continuous)
    assert np.array_equal(out, (want + 2).astype(np.float32))
    assert want[0] > 0 or n < 100


def test_hb_corr_different_selections_and_dense_list():
    """acceptors != donors, and a maxR so large that the frame-0 pair list overflows its first guess"""
    c = synth.hbond_case(n_water=300, n_frames=12, seed=99)
    acc = np.ascontiguousarray(c["acceptors"][:170])
    don = np.ascontiguousarray(c["donors"][40:])
    hyd = np.ascontiguousarray(c["hydrogens"][40:])
    for maxR, ang in ((2.8, 130.0), (60.0, 1.0)):
        for cont in (0, 1):
            want = port.hb_corr(acc, don, hyd, c["cellDims"], maxR, ang, cont)
            got = plf.py_getHBCounts(acc, don, hyd, c["cellDims"], maxR, ang, cont)
            assert np.array_equal(got, want), (maxR, cont)
