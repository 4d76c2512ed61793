"""Round-2 parity gates (all through the C ABI, on the GPU):

 * the hardware property the tensor-core filter rests on, pinned by the suite itself: one production
   tcgen05.mma with fp16 accumulators returns the correctly rounded EXACT sum, and the packed read-back
   maps column 2n / 2n + 1 to the low / high half of register n (nda_debug_tc_probe);
 * BASELINE configs 2 and 5 at FULL size against outputs of the reference's own code
   (tests/golden/golden_ref_strict_full.npz, written by tests/golden/make_golden.py --full);
 * grouped histograms that do not fit in shared memory (proteins above ~230 residues);
 * the frames of one host call sharded over several devices in ONE process (nda_set_devices) with one
   ncclAllReduce -- run when the box has at least two GPUs;
 * nda_set_option / nda_shutdown / the reference's accumulate-into-out convention / compat switch B-1;
 * two caller streams sharing the device API's scratch buffers.
"""
import ctypes
import os
import zlib

import numpy as np
import pytest

from namdanalyzer_b200 import synth, _capi
from namdanalyzer_b200.lib import pylibFuncs as plf
from oracle import port

pytestmark = pytest.mark.gpu
NB = 149
HERE = os.path.dirname(os.path.abspath(__file__))


# ------------------------------------------------------------------------------------------------ tcgen05 probe
def _probe(A16, B16):
    lib = _capi.load()
    A = np.ascontiguousarray(A16.view(np.uint16))
    B = np.ascontiguousarray(B16.view(np.uint16))
    out = np.zeros((128, 128), np.uint32)
    _capi.check(lib.nda_debug_tc_probe(A.ctypes.data_as(ctypes.POINTER(ctypes.c_uint16)),
                                       B.ctypes.data_as(ctypes.POINTER(ctypes.c_uint16)),
                                       out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32))))
    return out


def _decode(raw):
    """raw registers [128][4 warps x 32] -> D float16 [128][256] with the DOCUMENTED map: register n of the
    warp that owns columns 64 cq .. holds column 64 cq + 2n in its low half and 64 cq + 2n + 1 in its high half"""
    D = np.zeros((128, 256), np.float16)
    r = raw.reshape(128, 4, 32)
    lo = (r & 0xffff).astype(np.uint16).view(np.float16)
    hi = (r >> 16).astype(np.uint16).view(np.float16)
    for cq in range(4):
        D[:, 64 * cq + 0:64 * cq + 64:2] = lo[:, cq, :]
        D[:, 64 * cq + 1:64 * cq + 64:2] = hi[:, cq, :]
    return D


def test_tc_probe_register_map():
    """columns: D[i][j] = j + 1; rows: D[i][j] = i + 1 -- any other lane / column / half assignment of
    tcgen05.ld.32x32b.x32.pack::16b (or of the K-major operand layout) shows up as a wrong integer"""
    A = np.zeros((128, 16), np.float16)
    B = np.zeros((256, 16), np.float16)
    A[:, 3] = 1.0
    B[:, 3] = np.arange(1, 257)
    D = _decode(_probe(A, B))
    assert np.array_equal(D, np.broadcast_to(np.arange(1, 257, dtype=np.float16), (128, 256)))
    A[:, 3] = 0
    B[:, 3] = 0
    A[:, 12] = np.arange(1, 129)                            # second K chunk
    B[:, 12] = 1.0
    D = _decode(_probe(A, B))
    assert np.array_equal(D, np.broadcast_to(np.arange(1, 129, dtype=np.float16)[:, None], (128, 256)))


def _embed64(x, c):
    rho = c / (2 * np.pi)
    th = 2 * np.pi * x / c
    return np.stack([rho * np.cos(th), rho * np.sin(th)], -1).reshape(x.shape[0], 6)    # x: [n][3]


def test_tc_fp16_accumulator_is_the_rounded_exact_sum():
    """>= 1e6 pairs placed around the cut-off (D near zero), operands in the PRODUCTION K-slot layout incl. the
    (1, 1) x (-t_hi, -t_lo) threshold slots and the (isPad, -6e4) x (-6e4, isPad) padding slots:
    |D - exact| <= one fp16 rounding of the exact float64 sum + 2^-19 P (the accumulator is summed in at least
    fp32: 16 terms of size <= P), and sign(D) == sign(exact) whenever |exact| exceeds that.  A device that
    accumulated in fp16 internally would be off by ~2^-11 P (= the filter's whole slack E; the threshold keeps
    0.25 E in reserve) and fails this by two orders of magnitude."""
    rng = np.random.default_rng(20)
    total = near = 0
    worst = 0.0
    for rep in range(44):
        c = np.array([rng.uniform(25, 90), rng.uniform(25, 90), rng.uniform(25, 90)])
        rho2 = (c / (2 * np.pi)) ** 2
        P = rho2.sum()
        shell = rep % 5 != 4
        # (shell probes: cut <= 0.06 c, where the chord sum S is within ~1 % of r^2 and D really is near zero at r = cut)
        cut = rng.uniform(2.5, max(2.6, 0.06 * c.min())) if shell else rng.uniform(2.5, 0.2 * c.min())
        thr = cut * cut
        E = P * (2.0 ** -11 * 1.01 + 2.0 ** -18)
        thrDot = np.float32(P - 0.5 * thr - 1.25 * E)
        if shell:
            # ALL 128 x 256 pairs at the cut-off: the A atoms on a sphere of radius ~cut around a centre somewhere in
            # (or many cells away from) the box, the B atoms in a small ball around that centre -> every D is near zero
            centre = rng.uniform(-2 * c, 3 * c, 3)
            u = rng.standard_normal((128, 3))
            u /= np.linalg.norm(u, axis=1)[:, None]
            xa = centre + u * cut * rng.uniform(0.995, 1.005, (128, 1))
            xb = centre + rng.uniform(-0.005, 0.005, (256, 3)) * cut
        else:
            # a random box (D from -2 P to +thr / 2) with every B atom at ~cut from some A atom, and a few padding atoms
            xa = rng.uniform(-2 * c, 3 * c, (128, 3))
            u = rng.standard_normal((256, 3))
            u /= np.linalg.norm(u, axis=1)[:, None]
            xb = xa[rng.integers(0, 128, 256)] + u * cut * rng.uniform(0.98, 1.02, (256, 1))
        ea, eb = _embed64(xa, c), _embed64(xb, c)
        A = np.zeros((128, 16), np.float16)
        B = np.zeros((256, 16), np.float16)
        A[:, :6] = ea.astype(np.float16)
        A[:, 8:14] = (ea - A[:, :6].astype(np.float64)).astype(np.float16)          # lo = fp16(phi - hi)
        A[:, 6] = A[:, 7] = 1.0
        A[:, 14] = 0.0
        A[:, 15] = -60000.0
        B[:, :6] = eb.astype(np.float16)
        B[:, 8:14] = B[:, :6]
        th = np.float16(thrDot)
        if np.float32(th) > thrDot:                                                  # round DOWN like __float2half_rd
            th = np.nextafter(th, np.float16(-np.inf))
        tl = np.float16(np.float32(thrDot) - np.float32(th))
        B[:, 6], B[:, 7] = -th, -tl
        B[:, 14] = -60000.0
        B[:, 15] = 0.0
        if rep % 5 == 4:                                                             # a few padding atoms on both sides
            A[120:, :6] = 0; A[120:, 8:14] = 0; A[120:, 14] = 1.0
            B[250:, :6] = 0; B[250:, 8:14] = 0; B[250:, 15] = 1.0
        D = _decode(_probe(A, B)).astype(np.float64)
        exact = A.astype(np.float64) @ B.astype(np.float64).T                        # exact: 22-bit products, < 53-bit span
        real = np.ones((128, 256), bool)
        if rep % 5 == 4:
            real[120:, :] = False
            real[:, 250:] = False
            assert np.all(D[~real] <= -5.9e4)                                       # padding can never be a candidate
        ex, d = exact[real], D[real]
        floor = P * 2.0 ** -19 + 2.0 ** -14                                          # fp32 accumulation; fp16 subnormals may flush
        tol = np.abs(ex) * 2.0 ** -11 * 1.01 + floor
        assert np.all(np.abs(d - ex) <= tol), float(np.max(np.abs(d - ex) / tol))
        clear = np.abs(ex) > floor
        assert np.array_equal(np.signbit(d[clear]), np.signbit(ex[clear]))
        assert floor < 0.25 * E / 30                                                 # ... far inside the threshold's reserve
        small = np.abs(ex) < P * 2.0 ** -8                                           # decisions near the threshold
        worst = max(worst, float(np.max((np.abs(d - ex) / P)[small])) if np.any(small) else 0.0)
        total += ex.size
        near += int(np.sum(small))
    assert total >= 1_000_000
    assert near >= 1_000_000                     # a million decisions within P / 256 of the threshold
    assert worst <= 2.0 ** -16                   # near the decision the accumulator is within 1.6e-5 P of the exact sum (reserve 0.25 E = 1.2e-4 P)
    print("tc probe: %d pairs, %d near the threshold, worst |D - exact| / P there = %.3g" % (total, near, worst))


# ------------------------------------------------------------------------------------------------ full-size goldens
@pytest.fixture(scope="module")
def golden_full():
    p = os.path.join(HERE, "golden", "golden_ref_strict_full.npz")
    return dict(np.load(p))


def _crc(*arrays):
    c = 0
    for a in arrays:
        c = zlib.crc32(np.ascontiguousarray(a).tobytes(), c)
    return c


def test_config2_all_1000_frames_against_the_reference(golden_full):
    """BASELINE configs[1] at its stated size: every one of the 1000 x 7923 mask bits equals the reference's"""
    c = synth.config2(n_frames=1000)
    assert _crc(c["allAtoms"], c["cellDims"]) == int(golden_full["c2full_crc"])
    bits = plf.py_getWithinBits(c["allAtoms"], c["refSel"], c["outSel"], c["cellDims"], c["distance"])
    assert _capi.load().nda_last_filter() == 100
    hits = np.unpackbits(bits.view(np.uint8), axis=1, bitorder="little").sum(1)
    assert np.array_equal(hits, golden_full["c2full_hits_per_frame"])
    frame_crc = np.array([zlib.crc32(bits[f].tobytes()) for f in range(1000)], np.uint32)
    assert np.array_equal(frame_crc, golden_full["c2full_bits_frame_crc"])
    assert _crc(bits) == int(golden_full["c2full_bits_crc"])
    # and through the reference-shaped operator (int32 matrix)
    out = np.zeros(c["allAtoms"].shape[:2], np.int32)
    plf.py_getWithin(c["allAtoms"], c["refSel"], c["outSel"], out, c["cellDims"], c["distance"])
    assert np.array_equal(out.sum(0), golden_full["c2full_hits_per_frame"])
    with _capi.options(NDA_FILTER="fold2"):                 # the packed-FP32 filter: same bits
        bits2 = plf.py_getWithinBits(c["allAtoms"], c["refSel"], c["outSel"], c["cellDims"], c["distance"])
    assert np.array_equal(bits, bits2)


def test_config5_328509_atoms_against_the_reference(golden_full):
    """BASELINE configs[4] shape, one frame of 69^3 oxygens (5.4e10 pairs): histogram equal bin by bin"""
    c = synth.config5(n_side=69, n_frames=1)
    assert _crc(c["sel1"], c["cellDims"]) == int(golden_full["c5full_crc"])
    got = plf.py_getRadialNbrCounts(c["sel1"], c["sel1"], c["cellDims"], 1, NB, 0.1)
    assert _capi.load().nda_last_filter() == 100
    assert np.array_equal(got.astype(np.int64), golden_full["c5full_counts"].astype(np.int64))


# ------------------------------------------------------------------------------------------------ many groups
@pytest.mark.parametrize("filt", ["tc", "fold2"])
def test_grouped_histogram_larger_than_shared_memory(filt):
    """ResidueWiseWaterDensity on a protein of 700 residues: 700 x 149 = 104 300 cells do not fit in shared
    memory (max ~35k / ~51k); the library runs the pair kernel per batch of groups"""
    rng = np.random.default_rng(12)
    nG, nF = 700, 2
    L = 40.0
    prot = rng.uniform(0, L, (2 * nG, nF, 3)).astype(np.float32)
    wat = rng.uniform(0, L, (1500, nF, 3)).astype(np.float32)
    gid = np.repeat(np.arange(nG, dtype=np.int32), 2)
    rng.shuffle(gid)
    cell = np.full((nF, 3), L, np.float32)
    want = port.rdf_counts(wat, prot, cell, 0, NB, 0.08, groupId=gid, nGroups=nG).astype(np.int64)
    with _capi.options(NDA_FILTER=filt, NDA_TC_ALLOW_PADDING="1"):
        got = plf.py_getRadialNbrCounts(wat, prot, cell, 0, NB, 0.08, groupId=gid, nGroups=nG).astype(np.int64)
        assert (_capi.load().nda_last_filter() == 100) == (filt == "tc")
    assert got.shape == (nG, NB)
    assert np.array_equal(got, want)
    assert want.sum() > 10000


# ------------------------------------------------------------------------------------------------ options, conventions
def test_reference_accumulates_into_out_and_options_round_trip():
    lib = _capi.load()
    c = synth.config1(n_frames=3)
    nb = synth.n_bins(c["dr"], c["maxR"])
    want = (port.rdf_counts(c["sel1"], c["sel1"], c["cellDims"], 1, nb, c["dr"]).astype(np.float64) / 3).astype(np.float32)
    out = np.zeros(nb, np.float32)
    plf.py_getRadialNbrDensity(c["sel1"], c["sel1"], out, c["cellDims"], 1, c["maxR"], c["dr"])
    assert np.array_equal(out, want)
    plf.py_getRadialNbrDensity(c["sel1"], c["sel1"], out, c["cellDims"], 1, c["maxR"], c["dr"])
    assert np.array_equal(out, want + want)                 # getRadialNbrDensity.cpp:40-41 adds to what is there
    with pytest.raises(RuntimeError, match="unknown option"):
        _capi.set_option("NDA_NO_SUCH_SWITCH", "1")
    _capi.check(lib.nda_shutdown())                         # everything released ...
    out[:] = 0
    plf.py_getRadialNbrDensity(c["sel1"], c["sel1"], out, c["cellDims"], 1, c["maxR"], c["dr"])
    assert np.array_equal(out, want)                        # ... and rebuilt on demand


def test_compat_cuda_marks_reference_atoms():
    """SURVEY B-1: the legacy CUDA build flags every refSel atom in every frame; opt-in switch"""
    c = synth.config2(n_frames=3, n_water=900)
    want = port.within(c["allAtoms"], c["refSel"], c["outSel"], c["cellDims"], 3.0)
    out = np.zeros_like(want)
    with _capi.options(NDA_COMPAT_CUDA="1"):
        plf.py_getWithin(c["allAtoms"], c["refSel"], c["outSel"], out, c["cellDims"], 3.0)
    legacy = want.copy()
    legacy[c["refSel"], :] = 1
    assert np.array_equal(out, legacy)
    out[:] = 0
    plf.py_getWithin(c["allAtoms"], c["refSel"], c["outSel"], out, c["cellDims"], 3.0)
    assert np.array_equal(out, want)


def test_two_caller_streams_share_the_scratch_buffers():
    """nda_dev_* calls on DIFFERENT streams use the same lane scratch; the library orders them"""
    import torch
    lib = _capi.load()
    _capi.check(lib.nda_set_device(0))
    dev = torch.device("cuda", 0)
    cs = [synth.config2(n_frames=40, n_water=2500, seed=77 + k) for k in range(2)]
    want = [np.packbits(port.within(c["allAtoms"], c["refSel"], c["outSel"], c["cellDims"], 3.0)[c["outSel"]].T.astype(np.uint8),
                        axis=1, bitorder="little") for c in cs]
    streams = [torch.cuda.Stream(device=dev) for _ in range(2)]
    dv = []
    for c in cs:
        words = (c["outSel"].size + 31) // 32
        dv.append(dict(xyz=torch.from_numpy(c["allAtoms"]).to(dev), cell=torch.from_numpy(c["cellDims"]).to(dev),
                       r=torch.from_numpy(c["refSel"]).to(dev), o=torch.from_numpy(c["outSel"]).to(dev),
                       bits=torch.zeros((40, words), dtype=torch.int32, device=dev)))
    torch.cuda.synchronize()
    for rep in range(6):
        for k in (0, 1):
            d, c = dv[k], cs[k]
            d["bits"].zero_()
        torch.cuda.synchronize()
        for k in (0, 1):
            d, c = dv[k], cs[k]
            _capi.check(lib.nda_dev_within(d["xyz"].data_ptr(), c["allAtoms"].shape[0], 40, d["r"].data_ptr(), c["refSel"].size,
                                           d["o"].data_ptr(), c["outSel"].size, d["bits"].data_ptr(), None,
                                           d["cell"].data_ptr(), 3.0, 0, 40, ctypes.c_void_p(streams[k].cuda_stream)))
        torch.cuda.synchronize()
        for k in (0, 1):
            got = dv[k]["bits"].cpu().numpy().view(np.uint8)
            nbytes = want[k].shape[1]
            assert np.array_equal(got[:, :nbytes], want[k]), (rep, k)


# ------------------------------------------------------------------------------------------------ several devices, one process
def _n_gpus():
    try:
        return int(_capi.load().nda_device_count())
    except Exception:
        return 0


@pytest.mark.skipif(_n_gpus() < 2, reason="needs at least two GPUs in the box")
@pytest.mark.parametrize("collective", ["nccl", "host"])
def test_frames_sharded_over_devices_in_one_process(collective):
    """nda_set_devices: py_getWithin / py_getRadialNbrCounts / py_getDistances on G devices == oracle; the
    histogram and the distance sums are merged by ONE ncclAllReduce (nda_last_collective() == 1)"""
    lib = _capi.load()
    G = min(_n_gpus(), 4)
    try:
        _capi.set_devices(list(range(G)))
        with _capi.options(NDA_COLLECTIVE=collective):
            c = synth.config2(n_frames=37, n_water=2000)
            out = np.zeros(c["allAtoms"].shape[:2], np.int32)
            plf.py_getWithin(c["allAtoms"], c["refSel"], c["outSel"], out, c["cellDims"], 3.0)
            assert np.array_equal(out, port.within(c["allAtoms"], c["refSel"], c["outSel"], c["cellDims"], 3.0))
            assert lib.nda_last_collective() == 0
            bits = plf.py_getWithinBits(c["allAtoms"], c["refSel"], c["outSel"], c["cellDims"], 3.0, 5, 30)
            one = np.unpackbits(bits.view(np.uint8), axis=1, bitorder="little")[:, :c["outSel"].size]
            assert np.array_equal(one.T, out[c["outSel"], 5:30])
            c1 = synth.config1(n_frames=11)
            nb = synth.n_bins(c1["dr"], c1["maxR"])
            got = plf.py_getRadialNbrCounts(c1["sel1"], c1["sel1"], c1["cellDims"], 1, nb, c1["dr"])
            assert np.array_equal(got, port.rdf_counts(c1["sel1"], c1["sel1"], c1["cellDims"], 1, nb, c1["dr"]))
            assert lib.nda_last_collective() == (1 if collective == "nccl" else 2)
            c3 = synth.config3(n_frames=7)
            got = plf.py_getRadialNbrCounts(c3["waters"], c3["protein"], c3["cellDims"], 0, NB, 0.1,
                                            groupId=c3["groupId"], nGroups=c3["nGroups"])
            want = port.rdf_counts(c3["waters"], c3["protein"], c3["cellDims"], 0, NB, 0.1, groupId=c3["groupId"], nGroups=c3["nGroups"])
            assert np.array_equal(got, want)
            c4 = synth.config4(n_frames=203)
            d = plf.py_getDistancesSum(c4["sel1"], c4["sel2"], c4["cellDims"], 0)
            dw = port.distances_sum(c4["sel1"], c4["sel2"], c4["cellDims"], 0)
            assert np.max(np.abs(d - dw) / np.maximum(dw, 1e-30)) < 1e-12
            assert lib.nda_last_collective() == (1 if collective == "nccl" else 2)
            ds = plf.py_getDistancesSum(c4["sel1"], c4["sel1"], c4["cellDims"], 1)
            dsw = port.distances_sum(c4["sel1"], c4["sel1"], c4["cellDims"], 1)
            iu = np.triu_indices(7, 1)
            assert np.max(np.abs(ds[iu] - dsw[iu]) / dsw[iu]) < 1e-12 and np.array_equal(ds, ds.T)
            # fewer frames than devices
            one = np.ascontiguousarray(c1["sel1"][:, :1])
            got = plf.py_getRadialNbrCounts(one, one, c1["cellDims"][:1], 1, nb, c1["dr"])
            assert np.array_equal(got, port.rdf_counts(one, one, c1["cellDims"][:1], 1, nb, c1["dr"]))
            # page-locked input (namdanalyzer_b200.pin): every device DMAs its own frame columns straight from the array
            import namdanalyzer_b200 as nda
            xyz = nda.pin(c["allAtoms"].copy())
            try:
                out2 = np.zeros_like(out)
                plf.py_getWithin(xyz, c["refSel"], c["outSel"], out2, c["cellDims"], 3.0)
                assert np.array_equal(out2, out)
            finally:
                nda.unpin(xyz)
    finally:
        _capi.check(lib.nda_set_device(0))


# ------------------------------------------------------------------------------------------------ stress / determinism
def _run_tool(name, env):
    import subprocess
    import sys
    root = os.path.dirname(HERE)
    e = dict(os.environ, **env)
    r = subprocess.run([sys.executable, os.path.join(root, "tools", name)], capture_output=True, text=True, timeout=600, env=e, cwd=root)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return r.stdout


def test_stress_random_shapes_tensor_vs_fp32_filter():
    """tools/stress_tc.py: random shapes / frame counts / ranges, tensor-core filter against the packed-FP32 filter
    (within masks, two-selection and same-selection histograms must be identical); shakes out hand-over races"""
    out = _run_tool("stress_tc.py", {"ITERS": "120", "SEED": "5"})
    assert "stress ok" in out


def test_repeated_calls_are_bit_identical():
    """tools/repeat_check.py: the same 300-frame config-2 call 60 times -- every bitmask identical to the first"""
    out = _run_tool("repeat_check.py", {"NF": "300", "REPS": "60"})
    assert "calls with differences: 0" in out, out[-500:]


def test_config3_200_frames_and_config4_10000_frames_against_the_reference(golden_full):
    """configs[2] (residue-wise water RDF, one grouped launch against 76 reference calls per frame set) on 200 frames, and
    configs[3] (getDistances 7 x 1231) at its stated 10 000 frames, against the reference's own code"""
    c = synth.config3(n_frames=200)
    assert _crc(c["waters"], c["protein"], c["cellDims"]) == int(golden_full["c3big_crc"])
    got = plf.py_getRadialNbrCounts(c["waters"], c["protein"], c["cellDims"], 0, NB, 0.1, groupId=c["groupId"], nGroups=c["nGroups"])
    assert _capi.load().nda_last_filter() == 100
    assert np.array_equal(got.astype(np.int64), golden_full["c3big_group_counts"])
    c = synth.config4(n_frames=10000)
    assert _crc(c["sel1"], c["sel2"], c["cellDims"]) == int(golden_full["c4full_crc"])
    out = np.zeros((7, 1231), np.float32)
    plf.py_getDistances(c["sel1"], c["sel2"], out, c["cellDims"], 0)
    want = golden_full["c4full_mean64"]
    assert np.max(np.abs(out.astype(np.float64) - want) / np.maximum(want, 1e-30)) <= 1e-6      # north_star tolerance, float32 result
    s = plf.py_getDistancesSum(c["sel1"], c["sel2"], c["cellDims"], 0) / 10000.0
    nz = want > 0
    assert np.max(np.abs(s[nz] - want[nz]) / want[nz]) <= 1e-12                                  # the float64 sums themselves
    assert np.all(s[~nz] == 0.0)                                                                 # an atom against itself


def test_dense_ranges_run_all_exact_with_identical_counts():
    """Round 2: RDF frames whose cut-off sphere exceeds a quarter of the cell skip the filter (NDA_DENSE_FRACTION); small
    ungrouped calls use the 256-atom-tile kernel.  Same counts either way, also when dense and sparse frames mix in a call"""
    lib = _capi.load()
    c = synth.config1(n_frames=6)
    nb = synth.n_bins(c["dr"], c["maxR"])
    want = port.rdf_counts(c["sel1"], c["sel1"], c["cellDims"], 1, nb, c["dr"])
    st = (ctypes.c_double * 4)()
    got = plf.py_getRadialNbrCounts(c["sel1"], c["sel1"], c["cellDims"], 1, nb, c["dr"])
    _capi.check(lib.nda_last_stats(st))
    assert np.array_equal(got, want) and int(st[2]) == 6            # all six frames all-exact
    with _capi.options(NDA_DENSE_FRACTION="0"):
        got = plf.py_getRadialNbrCounts(c["sel1"], c["sel1"], c["cellDims"], 1, nb, c["dr"])
        _capi.check(lib.nda_last_stats(st))
        assert np.array_equal(got, want) and int(st[2]) == 0        # the packed-FP32 filter + candidate ring
    # frames 0-2 in the 31 A cell (dense for 15 A), frames 3-5 the same atoms in a 120 A cell (sparse): one call
    cell = c["cellDims"].copy()
    cell[3:] = 120.0
    a, b = np.ascontiguousarray(c["sel1"][:600]), np.ascontiguousarray(c["sel1"][600:])
    with _capi.options(NDA_FILTER="fold2"):
        got = plf.py_getRadialNbrCounts(a, b, cell, 0, nb, c["dr"])
        _capi.check(lib.nda_last_stats(st))
    assert np.array_equal(got, port.rdf_counts(a, b, cell, 0, nb, c["dr"])) and int(st[2]) == 3
