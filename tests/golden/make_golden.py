#!/usr/bin/env python
"""Generate the committed golden vectors (tests/golden/*.npz).

Run in the BUILD container only (needs /root/reference and oracle/_ref, i.e.
`make -C oracle ref`):

    python tests/golden/make_golden.py            # golden_ref_strict.npz, ubq_ws_fixture.npz
    python tests/golden/make_golden.py --full     # golden_ref_strict_full.npz (configs 2-5 at or near full size, ~15 min)

What it writes
  ubq_ws_fixture.npz   inputs taken from the reference's own test trajectory
                       package/tests/test_data/ubq_ws.{psf,dcd} (6682 atoms, 10 frames,
                       no unit cell): protein frame-0 coordinates + residue ids (used by the
                       synthetic configs 2-4) and 3 full frames (real-geometry parity case).
  golden_*.npz         OUTPUTS of the reference's own OpenMP sources (oracle/_ref/
                       libref_strict.so, strict IEEE flags, driven one frame per call) on the
                       seeded inputs of namdanalyzer_b200/synth.py, plus a CRC of the inputs
                       so a drifting generator is caught.

The reference's test-suite has no known-answer value for this path; these
files are the pins: oracle/port and oracle/np_oracle are checked against them on CPU, the
CUDA path on the GPU.
"""
import os
import sys
import struct
import zlib
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF_DATA = "/root/reference/package/tests/test_data"
PROT = ["GLY", "ALA", "VAL", "LEU", "ILE", "MET", "PHE", "TRP", "PRO", "SER", "THR", "CYS", "TYR",
        "ASN", "GLN", "ASP", "GLU", "LYS", "ARG", "HIS", "HSE", "HSP", "HSD"]


def crc(*arrays):
    c = 0
    for a in arrays:
        c = zlib.crc32(np.ascontiguousarray(a).tobytes(), c)
    return np.uint32(c)


def read_psf(path):
    resid, resname, name = [], [], []
    with open(path) as fh:
        lines = fh.readlines()
    for i, ln in enumerate(lines):
        if "!NATOM" in ln:
            n = int(ln.split()[0])
            for l in lines[i + 1:i + 1 + n]:
                t = l.split()
                resid.append(int(t[2]))
                resname.append(t[3])
                name.append(t[4])
            break
    return np.array(resid), np.array(resname), np.array(name)


def read_dcd(path):
    """-> xyz float32 [nAtoms][nFrames][3] (the layout of NAMDAnalyzer's dcdData), has_cell"""
    b = open(path, "rb").read()
    assert struct.unpack("<i", b[:4])[0] == 84 and b[4:8] == b"CORD"
    icntrl = struct.unpack("<20i", b[8:88])
    nF, has_cell = icntrl[0], icntrl[10]
    p = 92
    tl = struct.unpack("<i", b[p:p + 4])[0]
    p += 8 + tl
    assert struct.unpack("<i", b[p:p + 4])[0] == 4
    nA = struct.unpack("<i", b[p + 4:p + 8])[0]
    p += 12
    xyz = np.empty((nA, nF, 3), np.float32)
    for f in range(nF):
        if has_cell:
            p += 8 + struct.unpack("<i", b[p:p + 4])[0]
        for k in range(3):
            ln = struct.unpack("<i", b[p:p + 4])[0]
            assert ln == 4 * nA
            xyz[:, f, k] = np.frombuffer(b, "<f4", nA, p + 4)
            p += 8 + ln
    return xyz, bool(has_cell)


def make_fixture():
    resid, resname, name = read_psf(os.path.join(REF_DATA, "ubq_ws.psf"))
    xyz, has_cell = read_dcd(os.path.join(REF_DATA, "ubq_ws.dcd"))
    assert xyz.shape == (6682, 10, 3) and not has_cell
    is_prot = np.isin(resname, PROT)
    assert is_prot.sum() == 1231
    water_o = np.flatnonzero(name == "OH2").astype(np.int32)
    assert water_o.size == 1817          # the one count the reference's tests assert
    pres = resid[is_prot]
    uniq = np.unique(pres)
    assert uniq.size == 76
    np.savez_compressed(
        os.path.join(HERE, "ubq_ws_fixture.npz"),
        protein_xyz=xyz[is_prot, 0, :].astype(np.float32),
        protein_resid=pres.astype(np.int32),
        protein_resindex=np.searchsorted(uniq, pres).astype(np.int32),
        all_xyz=np.ascontiguousarray(xyz[:, [0, 4, 9], :]),
        protein_idx=np.flatnonzero(is_prot).astype(np.int32),
        water_o_idx=water_o,
        # NAMDAnalyzer/dataParsers/dcdCell.py:85-89: no unit cell -> every edge 100000.0
        all_cell=np.full((3, 3), 100000.0, np.float32))


def main():
    make_fixture()
    from namdanalyzer_b200 import synth
    from oracle.ref import RefLib
    ref = RefLib("strict")
    assert ref.parallel_backend() == 1
    nb = synth.n_bins(0.1, 15.0)
    assert nb == 149
    out = {}

    # config 1, full size: RDF O-O sameSel
    c = synth.config1()
    out["c1_crc"] = crc(c["sel1"], c["cellDims"])
    out["c1_counts"] = ref.rdf_counts(c["sel1"], c["sel1"], c["cellDims"], 1, nb, 0.1, 15.0)
    # config 1 inputs, two disjoint selections, sameSel = 0, first 10 frames
    a, b = c["sel1"][:300, :10], c["sel1"][300:, :10]
    out["c1ab_counts"] = ref.rdf_counts(a, b, c["cellDims"][:10], 0, nb, 0.1, 15.0)
    # odd bin width / few bins
    out["c1_dr037_counts"] = ref.rdf_counts(c["sel1"][:, :5], c["sel1"][:, :5], c["cellDims"][:5], 1, 17, 0.37, 0.0)
    # within on config-1 atoms, distance < 1 (per-axis early-out of getWithin.cpp:64-71 bites)
    # and a normal cutoff
    refSel = np.arange(0, 500, dtype=np.int32)
    outSel = np.arange(0, 1000, dtype=np.int32)
    for tag, dist in (("d08", 0.8), ("d35", 3.5)):
        m = ref.within(c["sel1"][:, :5], refSel, outSel, c["cellDims"][:5], dist)
        out["c1_within_%s" % tag] = np.packbits(m.astype(np.uint8), axis=None)

    # config 2, first 20 frames
    c = synth.config2(n_frames=20)
    out["c2_crc"] = crc(c["allAtoms"], c["cellDims"])
    m = ref.within(c["allAtoms"], c["refSel"], c["outSel"], c["cellDims"], c["distance"])
    assert m.shape == (25000, 20)
    out["c2_mask_bits"] = np.packbits(m.astype(np.uint8), axis=None)
    out["c2_hits_per_frame"] = m.sum(0).astype(np.int32)

    # config 3: 3 frames, all 76 residues, one reference call per residue
    c = synth.config3(n_frames=3)
    out["c3_crc"] = crc(c["waters"], c["protein"], c["cellDims"])
    g = np.zeros((c["nGroups"], nb), np.int64)
    for r in range(c["nGroups"]):
        sel = np.ascontiguousarray(c["protein"][c["groupId"] == r])
        g[r] = ref.rdf_counts(c["waters"], sel, c["cellDims"], 0, nb, 0.1, 15.0)
    out["c3_group_counts"] = g

    # config 4: 50 frames; per-frame float32 distances, mean formed in float64
    c = synth.config4(n_frames=50)
    out["c4_crc"] = crc(c["sel1"], c["sel2"], c["cellDims"])
    d = ref.distances_per_frame(c["sel1"], c["sel2"], c["cellDims"], 0)
    out["c4_mean64"] = d.astype(np.float64).mean(0)
    out["c4_last32"] = d[-1]
    # sameSel distances on the 7-atom selection
    d = ref.distances_per_frame(c["sel1"], c["sel1"], c["cellDims"], 1)
    out["c4_same_mean64"] = d.astype(np.float64).mean(0)

    # config 5 reduced: 27^3 atoms, 2 frames
    c = synth.config5(n_side=27, n_frames=2)
    out["c5_crc"] = crc(c["sel1"], c["cellDims"])
    out["c5_counts"] = ref.rdf_counts(c["sel1"], c["sel1"], c["cellDims"], 1, nb, 0.1, 15.0)

    # real geometry, no unit cell (cell = 1e5 -> the wrap is a no-op)
    fx = synth.fixture()
    xyz, cell = fx["all_xyz"], fx["all_cell"]
    m = ref.within(xyz, fx["protein_idx"], fx["water_o_idx"], cell, 3.0)
    out["ubq_mask_bits"] = np.packbits(m.astype(np.uint8), axis=None)
    wo = np.ascontiguousarray(xyz[fx["water_o_idx"]])
    out["ubq_oo_counts"] = ref.rdf_counts(wo, wo, cell, 1, nb, 0.1, 15.0)
    pr = np.ascontiguousarray(xyz[fx["protein_idx"]])
    r53 = np.ascontiguousarray(pr[fx["protein_resid"] == 53])
    d = ref.distances_per_frame(r53, pr, cell, 0)
    out["ubq_dist_mean64"] = d.astype(np.float64).mean(0)

    np.savez_compressed(os.path.join(HERE, "golden_ref_strict.npz"), **out)
    for k, v in out.items():
        print("%-22s %-12s %s" % (k, v.dtype, v.shape))


def bitmask_words(m, outSel):
    """int32 mask [nAtoms][nF] -> the library's bitmask uint32 [nF][ceil(nOut/32)]: bit (a & 31) of word a >> 5 = outSel[a]"""
    flags = np.ascontiguousarray(m[outSel].T).astype(np.uint8)             # [nF][nOut]
    words = (outSel.size + 31) // 32
    pad = np.zeros((flags.shape[0], words * 32), np.uint8)
    pad[:, :outSel.size] = flags
    return np.packbits(pad, axis=1, bitorder="little").view("<u4")


def main_full():
    """golden_ref_strict_full.npz: the BASELINE workloads at (or near) full size --
    config 2 at all 1000 frames (hits per frame + CRC of the bitmask), config 5 at 69^3 = 328 509 atoms, one frame
    (5.4e10 pairs: about a minute of the reference's OpenMP code here), config 3 at 200 frames x 76 residues and
    config 4 at all 10 000 frames."""
    from namdanalyzer_b200 import synth
    from oracle.ref import RefLib
    ref = RefLib("strict")
    nb = synth.n_bins(0.1, 15.0)
    out = {}
    c = synth.config2(n_frames=1000)
    out["c2full_crc"] = crc(c["allAtoms"], c["cellDims"])
    m = ref.within(c["allAtoms"], c["refSel"], c["outSel"], c["cellDims"], c["distance"])
    assert m.shape == (25000, 1000)
    out["c2full_hits_per_frame"] = m.sum(0).astype(np.int32)
    bits = bitmask_words(m, c["outSel"])
    out["c2full_bits_crc"] = crc(bits)
    out["c2full_bits_frame_crc"] = np.array([zlib.crc32(bits[f].tobytes()) for f in range(1000)], np.uint32)
    del m, c
    c = synth.config5(n_side=69, n_frames=1)
    out["c5full_crc"] = crc(c["sel1"], c["cellDims"])
    out["c5full_counts"] = ref.rdf_counts(c["sel1"], c["sel1"], c["cellDims"], 1, nb, 0.1, 15.0)
    # config 3, 200 frames, all 76 residues, one reference call per residue (4.9e9 pairs)
    c = synth.config3(n_frames=200)
    out["c3big_crc"] = crc(c["waters"], c["protein"], c["cellDims"])
    g = np.zeros((c["nGroups"], nb), np.int64)
    for r in range(c["nGroups"]):
        sel = np.ascontiguousarray(c["protein"][c["groupId"] == r])
        g[r] = ref.rdf_counts(c["waters"], sel, c["cellDims"], 0, nb, 0.1, 15.0)
    out["c3big_group_counts"] = g
    # config 4 at its stated size: 10 000 frames; per-frame float32 distances, mean formed in float64
    c = synth.config4(n_frames=10000)
    out["c4full_crc"] = crc(c["sel1"], c["sel2"], c["cellDims"])
    acc = np.zeros((7, 1231), np.float64)
    for f0 in range(0, 10000, 500):
        d = ref.distances_per_frame(np.ascontiguousarray(c["sel1"][:, f0:f0 + 500]), np.ascontiguousarray(c["sel2"][:, f0:f0 + 500]),
                                    c["cellDims"][f0:f0 + 500], 0)
        acc += d.astype(np.float64).sum(0)
    out["c4full_mean64"] = acc / 10000.0
    np.savez_compressed(os.path.join(HERE, "golden_ref_strict_full.npz"), **out)
    for k, v in out.items():
        print("%-24s %-12s %s" % (k, v.dtype, v.shape))


if __name__ == "__main__":
    if "--full" in sys.argv:
        main_full()
    else:
        main()
