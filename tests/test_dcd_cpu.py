"""CPU: SURVEY row f-2 -- the DCD ingest drop-ins (py_getDCDCoor / py_getDCDCell) against an
independent numpy parse of synthetic CHARMM/NAMD DCD files (both byte orders, with and without unit
cell) and, in the build container, against the reference's own test trajectory.  The reference's
reader itself cannot be compiled on Linux (lib/common/src/dcdReader.cpp:70,78), so the oracle here
is the file format as the reference's Python side reads it (dataParsers/dcdReader.py:190-236)."""
import os
import struct
import numpy as np
import pytest

from namdanalyzer_b200.lib import pylibFuncs as plf


def write_dcd(path, xyz, cell=None, order="<"):
    """xyz float32 [nAtoms][nFrames][3]; cell float64 [nFrames][6] or None"""
    nA, nF = xyz.shape[:2]
    title = b"synthetic".ljust(80)
    with open(path, "wb") as f:
        icntrl = [nF, 0, 1, nF, 0, 0, 0, 0, 0]
        hdr = struct.pack(order + "i4s9if10i", 84, b"CORD", *icntrl, 1.0, 1 if cell is not None else 0, *([0] * 8), 24)
        f.write(hdr + struct.pack(order + "i", 84))
        f.write(struct.pack(order + "ii", 84, 1) + title + struct.pack(order + "i", 84))
        f.write(struct.pack(order + "iii", 4, nA, 4))
        for fr in range(nF):
            if cell is not None:
                f.write(struct.pack(order + "i", 48) + cell[fr].astype(order + "f8").tobytes() + struct.pack(order + "i", 48))
            for k in range(3):
                f.write(struct.pack(order + "i", 4 * nA) + xyz[:, fr, k].astype(order + "f4").tobytes()
                        + struct.pack(order + "i", 4 * nA))
    title_size = 84
    rec = 12 * nA + (80 if cell is not None else 24)
    return (np.arange(nF, dtype=np.int64) * rec + 112 + title_size)     # dcdReader.py:221-230


@pytest.mark.parametrize("order", ["<", ">"])
@pytest.mark.parametrize("with_cell", [False, True])
def test_synthetic_dcd_roundtrip(tmp_path, order, with_cell):
    rng = np.random.default_rng(5)
    nA, nF = 257, 9
    xyz = rng.normal(0, 30, (nA, nF, 3)).astype(np.float32)
    cell = rng.uniform(20, 90, (nF, 6)) if with_cell else None
    path = str(tmp_path / "t.dcd")
    startPos = write_dcd(path, xyz, cell, order)
    # header sanity the way the reference parses it
    with open(path, "rb") as f:
        rec = struct.unpack(order + "i4c9if11i", f.read(92))
    assert rec[0] == 84 and rec[5] == nF and bool(rec[15]) == with_cell
    frames = np.array([8, 0, 3, 3, 5], np.int32)
    atoms = rng.permutation(nA)[:100].astype(np.int32)
    for dims in (np.array([0, 1, 2], np.int32), np.array([2, 0], np.int32), np.array([1], np.int32)):
        out = np.zeros((atoms.size, frames.size, dims.size), np.float32)
        rc = plf.py_getDCDCoor(bytearray(path, "utf-8"), frames, nA, atoms, dims, with_cell, startPos, out, ord(order))
        assert rc == 0
        assert np.array_equal(out, xyz[np.ix_(atoms, frames, dims)])
    if with_cell:
        c = np.zeros((frames.size, 6), np.float64)
        assert plf.py_getDCDCell(bytearray(path, "utf-8"), frames, startPos, c, ord(order)) == 0
        assert np.array_equal(c, cell[frames])


def test_dcd_error_codes(tmp_path):
    xyz = np.zeros((4, 2, 3), np.float32)
    path = str(tmp_path / "e.dcd")
    startPos = write_dcd(path, xyz)
    out = np.zeros((1, 1, 3), np.float32)
    a, d = np.array([0], np.int32), np.arange(3, dtype=np.int32)
    assert plf.py_getDCDCoor(b"/nonexistent/file.dcd", np.array([0], np.int32), 4, a, d, False, startPos, out, "<") == -1
    assert plf.py_getDCDCoor(path.encode(), np.array([5], np.int32), 4, a, d, False, startPos, out, "<") == -2
    assert plf.py_getDCDCoor(path.encode(), np.array([0], np.int32), 4, np.array([9], np.int32), d, False, startPos, out, "<") == -2
    far = np.array([10 ** 9], np.int64)
    assert plf.py_getDCDCoor(path.encode(), np.array([0], np.int32), 4, a, d, False, far, out, "<") == -3


REF_DCD = "/root/reference/package/tests/test_data/ubq_ws.dcd"


@pytest.mark.skipif(not os.path.exists(REF_DCD), reason="reference test data only exists in the build container")
def test_reference_test_trajectory():
    """the reference's own fixture: 6682 atoms, 10 frames, no unit cell (tests/dataParsers/test_dcdReader.py:8);
    frames 0, 4, 9 are the ones committed in tests/golden/ubq_ws_fixture.npz by an independent parser"""
    from namdanalyzer_b200 import synth
    fx = synth.fixture()
    with open(REF_DCD, "rb") as f:
        rec = struct.unpack("<i4c9if11i", f.read(92))
        title = struct.unpack("<i", f.read(4))[0]
        f.read(title + 4)
        nA = struct.unpack("<iii", f.read(12))[1]
    nF, has_cell = rec[5], bool(rec[15])
    assert (nA, nF, has_cell) == (6682, 10, False)
    startPos = np.arange(nF, dtype=np.int64) * (12 * nA + 24) + 112 + title
    frames = np.array([0, 4, 9], np.int32)
    out = np.zeros((nA, 3, 3), np.float32)
    rc = plf.py_getDCDCoor(bytearray(REF_DCD, "utf-8"), frames, nA, np.arange(nA, dtype=np.int32),
                           np.arange(3, dtype=np.int32), has_cell, startPos, out, ord("<"))
    assert rc == 0 and np.array_equal(out, fx["all_xyz"])


def test_dcdfile_accessor(tmp_path):
    from namdanalyzer_b200.dataParsers.dcdReader import DCDFile
    rng = np.random.default_rng(6)
    xyz = rng.normal(0, 20, (50, 7, 3)).astype(np.float32)
    cell = np.zeros((7, 6))
    cell[:, [0, 2, 5]] = rng.uniform(30, 40, (7, 3))
    p = str(tmp_path / "a.dcd")
    write_dcd(p, xyz, cell, ">")
    t = DCDFile(p)
    assert (t.nbrAtoms, t.nbrFrames, t.cell, t.byteorder) == (50, 7, True, ">")
    assert np.array_equal(t[np.array([3, 1, 49]), np.array([6, 0])], xyz[np.ix_([3, 1, 49], [6, 0])])
    assert np.array_equal(t[:, 2:5], xyz[:, 2:5])
    assert np.array_equal(t.cellDims[np.arange(7)], cell[:, [0, 2, 5]].astype(np.float32))
    p2 = str(tmp_path / "b.dcd")
    write_dcd(p2, xyz, None, "<")
    t2 = DCDFile(p2)
    assert not t2.cell and np.all(t2.cellDims[0:3] == 100000.0)       # dcdCell.py:85-89
    with pytest.raises(IndexError):
        t2[np.array([0]), np.array([7])]
