"""CPU: the oracle restatements (oracle/port = C, oracle/np_oracle = numpy) against the golden
vectors produced by the reference's own sources (tests/golden/make_golden.py), and -- where
oracle/_ref is present -- against the compiled reference directly.  Bit-exact throughout."""
import numpy as np
import pytest

from conftest import crc, unpack_mask
from namdanalyzer_b200 import synth
from oracle import port, np_oracle, ref

NB = 149


def test_nbins_matches_reference_callers():
    # NAMDAnalyzer/dataAnalysis/RadialDensity.py:206: np.arange(dr, maxR, dr)
    assert synth.n_bins(0.1, 15.0) == NB


def test_generators_are_stable(golden):
    c = synth.config1()
    assert crc(c["sel1"], c["cellDims"]) == golden["c1_crc"]
    c = synth.config2(n_frames=20)
    assert crc(c["allAtoms"], c["cellDims"]) == golden["c2_crc"]
    c = synth.config5(n_side=27, n_frames=2)
    assert crc(c["sel1"], c["cellDims"]) == golden["c5_crc"]


def test_port_rdf_config1_full(golden):
    c = synth.config1()
    got = port.rdf_counts(c["sel1"], c["sel1"], c["cellDims"], 1, NB, 0.1, 15.0)
    assert np.array_equal(got.astype(np.int64), golden["c1_counts"])
    # unique pairs only (getRadialNbrDensity.cpp:27)
    assert got.sum() <= 1000 * 999 // 2 * 100


def test_port_and_numpy_rdf_two_selections(golden):
    c = synth.config1()
    a, b, cell = c["sel1"][:300, :10], c["sel1"][300:, :10], c["cellDims"][:10]
    assert np.array_equal(port.rdf_counts(a, b, cell, 0, NB, 0.1).astype(np.int64), golden["c1ab_counts"])
    assert np.array_equal(np_oracle.rdf_counts(a, b, cell, 0, NB, 0.1), golden["c1ab_counts"])


def test_port_and_numpy_rdf_odd_bin_width(golden):
    c = synth.config1()
    a, cell = c["sel1"][:, :5], c["cellDims"][:5]
    assert np.array_equal(port.rdf_counts(a, a, cell, 1, 17, 0.37).astype(np.int64), golden["c1_dr037_counts"])
    assert np.array_equal(np_oracle.rdf_counts(a, a, cell, 1, 17, 0.37), golden["c1_dr037_counts"])


@pytest.mark.parametrize("tag,dist", [("d08", 0.8), ("d35", 3.5)])
def test_port_and_numpy_within_small_and_normal_cutoff(golden, tag, dist):
    c = synth.config1()
    xyz, cell = c["sel1"][:, :5], c["cellDims"][:5]
    refSel = np.arange(0, 500, dtype=np.int32)
    outSel = np.arange(0, 1000, dtype=np.int32)
    want = unpack_mask(golden["c1_within_%s" % tag], (1000, 5))
    assert np.array_equal(port.within(xyz, refSel, outSel, cell, dist), want)
    assert np.array_equal(np_oracle.within(xyz, refSel, outSel, cell, dist), want)
    # every ref atom that is also in outSel is within 0 of itself (getWithin.cpp:81-84)
    assert want[refSel].all()


def test_within_early_out_quirk_is_visible_below_one_angstrom():
    """getWithin.cpp:64-71 compares a signed component with distance^2: for distance < 1 a hit
    whose positive component lies in (distance^2, distance] is rejected (SURVEY B-2)."""
    xyz = np.zeros((2, 1, 3), np.float32)
    xyz[1, 0, 0] = 0.7                      # atom - ref = +0.7 > 0.8^2 = 0.64, but r = 0.7 <= 0.8
    cell = np.full((1, 3), 50.0, np.float32)
    r, o = np.array([0], np.int32), np.array([1], np.int32)
    assert port.within(xyz, r, o, cell, 0.8, literal_early_out=True)[1, 0] == 0
    assert port.within(xyz, r, o, cell, 0.8, literal_early_out=False)[1, 0] == 1
    assert np_oracle.within(xyz, r, o, cell, 0.8, True)[1, 0] == 0
    xyz[1, 0, 0] = -0.7                     # negative component: never rejected
    assert port.within(xyz, r, o, cell, 0.8, literal_early_out=True)[1, 0] == 1


def test_port_within_config2(golden):
    c = synth.config2(n_frames=20)
    want = unpack_mask(golden["c2_mask_bits"], (25000, 20))
    got = port.within(c["allAtoms"], c["refSel"], c["outSel"], c["cellDims"], c["distance"])
    assert np.array_equal(got, want)
    assert np.array_equal(got.sum(0), golden["c2_hits_per_frame"])
    # SURVEY B-1 (OpenMP semantics): ref atoms outside outSel are never flagged
    assert got[c["refSel"]].sum() == 0


def test_port_grouped_rdf_config3(golden):
    c = synth.config3(n_frames=3)
    got = port.rdf_counts(c["waters"], c["protein"], c["cellDims"], 0, NB, 0.1, 15.0,
                          groupId=c["groupId"], nGroups=c["nGroups"])
    assert np.array_equal(got.astype(np.int64), golden["c3_group_counts"])


def test_port_distances_config4(golden):
    c = synth.config4(n_frames=50)
    s, last = port.distances_sum(c["sel1"], c["sel2"], c["cellDims"], 0, want_last=True)
    assert np.array_equal(s / 50, golden["c4_mean64"])
    assert np.array_equal(last, golden["c4_last32"])
    s = port.distances_sum(c["sel1"], c["sel1"], c["cellDims"], 1)
    assert np.array_equal(s / 50, golden["c4_same_mean64"])
    assert np.all(np.tril(s) == 0)          # getDistances.cpp:44: only j > i is visited


def test_numpy_distances_config4(golden):
    c = synth.config4(n_frames=50)
    d = np_oracle.distances_per_frame(c["sel1"], c["sel2"], c["cellDims"], 0)
    assert np.array_equal(d.astype(np.float64).mean(0), golden["c4_mean64"])


def test_port_rdf_config5_reduced(golden):
    c = synth.config5(n_side=27, n_frames=2)
    got = port.rdf_counts(c["sel1"], c["sel1"], c["cellDims"], 1, NB, 0.1, 15.0)
    assert np.array_equal(got.astype(np.int64), golden["c5_counts"])


def test_port_real_geometry_no_unit_cell(golden):
    fx = synth.fixture()
    xyz, cell = fx["all_xyz"], fx["all_cell"]
    want = unpack_mask(golden["ubq_mask_bits"], (6682, 3))
    assert np.array_equal(port.within(xyz, fx["protein_idx"], fx["water_o_idx"], cell, 3.0), want)
    wo = np.ascontiguousarray(xyz[fx["water_o_idx"]])
    assert np.array_equal(port.rdf_counts(wo, wo, cell, 1, NB, 0.1).astype(np.int64), golden["ubq_oo_counts"])
    pr = np.ascontiguousarray(xyz[fx["protein_idx"]])
    r53 = np.ascontiguousarray(pr[fx["protein_resid"] == 53])
    assert r53.shape[0] == 7
    assert np.array_equal(port.distances_sum(r53, pr, cell, 0) / 3, golden["ubq_dist_mean64"])


@pytest.mark.skipif(not ref.available("strict"), reason="oracle/_ref not built")
def test_port_equals_compiled_reference_on_fresh_inputs():
    """beyond the golden files: random shapes, cells, unwrapped coordinates"""
    R = ref.RefLib("strict")
    rng = np.random.default_rng(7)
    for trial in range(4):
        n1, n2, nF = rng.integers(5, 200), rng.integers(5, 300), rng.integers(1, 4)
        cell = rng.uniform(18, 40, (nF, 3)).astype(np.float32)
        a = rng.uniform(-60, 90, (n1, nF, 3)).astype(np.float32)
        b = rng.uniform(-60, 90, (n2, nF, 3)).astype(np.float32)
        dr = float(rng.uniform(0.05, 0.5))
        nb = int(rng.integers(3, 200))
        assert np.array_equal(port.rdf_counts(a, b, cell, 0, nb, dr).astype(np.int64),
                              R.rdf_counts(a, b, cell, 0, nb, dr))
        assert np.array_equal(port.rdf_counts(a, a, cell, 1, nb, dr).astype(np.int64),
                              R.rdf_counts(a, a, cell, 1, nb, dr))
        d = R.distances_per_frame(a, b, cell, 0)
        assert np.array_equal(port.distances_sum(a, b, cell, 0), d.astype(np.float64).sum(0))
        allA = np.concatenate([a, b])
        refSel = np.arange(n1, dtype=np.int32)
        outSel = rng.permutation(n1 + n2)[: (n1 + n2) // 2].astype(np.int32)
        dist = float(rng.uniform(0.5, 8.0))
        assert np.array_equal(port.within(allA, refSel, outSel, cell, dist),
                              R.within(allA, refSel, outSel, cell, dist))


@pytest.mark.skipif(not (ref.available("strict") and ref.available("fast")), reason="oracle/_ref not built")
def test_shipped_flags_build_differs_only_at_bin_edges():
    """the -ffast-math build moves a handful of pairs across bin edges;
    totals stay equal and every difference is a +1/-1 between neighbouring bins."""
    c = synth.config1(n_frames=3)
    s = ref.RefLib("strict").rdf_counts(c["sel1"], c["sel1"], c["cellDims"], 1, NB, 0.1)
    f = ref.RefLib("fast").rdf_counts(c["sel1"], c["sel1"], c["cellDims"], 1, NB, 0.1)
    diff = f - s
    assert np.abs(diff).sum() <= 1e-4 * s.sum()
    assert abs(int(diff.sum())) <= 2          # only the outermost edge can leak a pair


# ---------------------------------------------------------------- SURVEY row f-3 (water at the surface)

def _surface_case(seed, nW=60, nP=200, nF=3, L=24.0):
    rng = np.random.default_rng(seed)
    cell = (L * (1 + 0.01 * rng.standard_normal((nF, 3)))).astype(np.float32)
    prot = rng.uniform(6, 18, (nP, nF, 3)).astype(np.float32)
    water = rng.uniform(-10, 34, (3 * nW, nF, 3)).astype(np.float32)       # 3 atoms per molecule, unwrapped
    water[1::3] = water[0::3] + rng.normal(0, 0.6, (nW, nF, 3)).astype(np.float32)
    water[2::3] = water[0::3] + rng.normal(0, 0.6, (nW, nF, 3)).astype(np.float32)
    return water, prot, cell


@pytest.mark.skipif(not ref.available("strict"), reason="oracle/_ref not built")
def test_port_set_water_dist_pbc_equals_reference():
    R = ref.RefLib("strict")
    water, prot, cell = _surface_case(31)
    # literal text: one call, running minimum carried across molecules and frames
    assert np.array_equal(port.set_water_dist_pbc(water, prot, cell, 3, reset_per_water=False),
                          R.set_water_dist_pbc_literal(water, prot, cell, 3))
    # per-molecule semantics = the reference called one molecule and one frame at a time
    per = port.set_water_dist_pbc(water, prot, cell, 3, reset_per_water=True)
    assert np.array_equal(per, R.set_water_dist_pbc_per_molecule(water, prot, cell, 3))
    # the carry-over is real: the two disagree on this input
    assert not np.array_equal(per, R.set_water_dist_pbc_literal(water, prot, cell, 3))
    # every molecule ends up within half a box (per axis) of its nearest protein atom: the shift
    # is a whole number of cell vectors
    shift = (water - per)[0::3]
    assert np.allclose(shift / cell[None], np.round(shift / cell[None]), atol=1e-4)


@pytest.mark.skipif(not ref.available("strict"), reason="oracle/_ref not built")
@pytest.mark.parametrize("maxN", [1, 3, 5])
def test_port_water_orient_equals_reference(maxN):
    R = ref.RefLib("strict")
    water, prot, cell = _surface_case(32 + maxN, nW=150)
    wO = np.ascontiguousarray(water[0::3])
    vec = np.ascontiguousarray((wO - water[1::3]) + (wO - water[2::3])) / np.float32(2)
    with np.errstate(all="ignore"):
        got = port.water_orient_at_surface(wO, vec, prot, cell, 0.0, 5.0, maxN)
        want = R.water_orient_at_surface(wO, vec, prot, cell, 0.0, 5.0, maxN)
    assert np.array_equal(got, want, equal_nan=True)
    assert 0 < want[..., 1].sum() < want[..., 1].size          # some waters at the surface, some not


# ---------------------------------------------------------------- SURVEY row f-4 (H-bond autocorrelation)

@pytest.mark.skipif(not ref.available("strict"), reason="oracle/_ref not built")
@pytest.mark.parametrize("continuous", [0, 1])
def test_port_hb_corr_equals_reference(continuous):
    R = ref.RefLib("strict")
    c = synth.hbond_case()
    got = port.hb_corr(c["acceptors"], c["donors"], c["hydrogens"], c["cellDims"], c["maxR"], c["minAngle"], continuous)
    want = R.hb_corr(c["acceptors"], c["donors"], c["hydrogens"], c["cellDims"], c["maxR"], c["minAngle"], continuous)
    assert np.array_equal(got.astype(np.float32), want)
    assert want[0] > 20 and want[-1] < want[0]               # bonds exist at t = 0 and decay
    if continuous:
        assert np.all(np.diff(want) <= 0)                    # an unbroken count can only go down


def test_port_against_full_size_config2_golden():
    """tests/golden/golden_ref_strict_full.npz (reference's own code on all 1000 frames of config 2): the C port
    on every 10th frame -- hit counts and the CRC of the frame's bitmask words"""
    import os
    import zlib
    g = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_ref_strict_full.npz")))
    c = synth.config2(n_frames=1000)
    assert crc(c["allAtoms"], c["cellDims"]) == g["c2full_crc"]
    frames = np.arange(0, 1000, 10)
    xyz = np.ascontiguousarray(c["allAtoms"][:, frames])
    m = port.within(xyz, c["refSel"], c["outSel"], c["cellDims"][frames], c["distance"])
    assert np.array_equal(m.sum(0), g["c2full_hits_per_frame"][frames])
    flags = np.ascontiguousarray(m[c["outSel"]].T).astype(np.uint8)
    words = (c["outSel"].size + 31) // 32
    pad = np.zeros((frames.size, words * 32), np.uint8)
    pad[:, :c["outSel"].size] = flags
    bits = np.packbits(pad, axis=1, bitorder="little")
    got = np.array([zlib.crc32(bits[i].tobytes()) for i in range(frames.size)], np.uint32)
    assert np.array_equal(got, g["c2full_bits_frame_crc"][frames])


def test_port_against_config3_200_frames_and_config4_full_golden():
    """the C port against the reference's own output for config 3 (76 groups, 200 frames, ONE grouped pass against the
    reference's one call per residue) and config 4 at its stated 10 000 frames"""
    import os
    g = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_ref_strict_full.npz")))
    c = synth.config3(n_frames=200)
    assert crc(c["waters"], c["protein"], c["cellDims"]) == g["c3big_crc"]
    got = port.rdf_counts(c["waters"], c["protein"], c["cellDims"], 0, NB, 0.1, groupId=c["groupId"], nGroups=c["nGroups"])
    assert np.array_equal(got.astype(np.int64), g["c3big_group_counts"])
    c = synth.config4(n_frames=10000)
    assert crc(c["sel1"], c["sel2"], c["cellDims"]) == g["c4full_crc"]
    s = port.distances_sum(c["sel1"], c["sel2"], c["cellDims"], 0) / 10000.0
    want = g["c4full_mean64"]
    assert np.max(np.abs(s - want) / np.maximum(want, 1e-30)) <= 1e-12
