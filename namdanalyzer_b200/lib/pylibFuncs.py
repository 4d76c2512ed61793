"""Drop-in for ``NAMDAnalyzer.lib.pylibFuncs`` restricted to the minimum-image pair path.

Same callables, positional signatures and in-place output convention as the reference's Cython
module (package/NAMDAnalyzer/lib/libFunc.pyx):

    py_getWithin(allAtoms, refSel, outSel, out, cellDims, distance)            libFunc.pyx:176-186
    py_getDistances(sel1, sel2, out, cellDims, sameSel)                        libFunc.pyx:118-127
    py_getRadialNbrDensity(sel1, sel2, out, cellDims, sameSel, maxR, dr)       libFunc.pyx:131-140
    py_getParallelBackend()                                                    libFunc.pyx:356-358

Argument validation reproduces Cython's typed-buffer errors (TypeError for None / non-arrays,
ValueError for dtype, ndim and contiguity), with one deliberate relaxation: ``cellDims`` may be
non-C-contiguous, because ``dcdCell.__getitem__`` returns an F-ordered array for several frames
(dataParsers/dcdCell.py:114) and the stock callers pass it on unchanged
(dataParsers/dcdParser.py:173, dataAnalysis/RadialDensity.py:82,217).
Shapes are cross-checked (the reference trusts them).  CUDA failures raise RuntimeError.

The work is done by hand-written sm_100a CUDA kernels behind the C ABI of
``include/nda_b200.h``; nothing here computes on the CPU.
"""
import ctypes
import numpy as np

from .. import _capi
from .._capi import c_f32p, c_f64p, c_i32p, c_u32p, c_u64p

__all__ = ["py_getWithin", "py_getDistances", "py_getRadialNbrDensity", "py_getParallelBackend",
           "py_getRadialNbrDensityGrouped", "py_getRadialNbrCounts", "py_getDistancesSum",
           "py_getWithinBits", "py_getWithinLists", "py_setWaterDistPBC", "py_waterOrientAtSurface",
           "py_getHBCorr", "py_getHBCounts", "py_getDCDCoor", "py_getDCDCell"]

_CNAME = {np.dtype(np.float32): "float", np.dtype(np.float64): "double", np.dtype(np.int32): "int",
          np.dtype(np.int64): "long", np.dtype(np.uint64): "unsigned long", np.dtype(np.uint32): "unsigned int"}


def _buf(x, name, dtype, ndim, need_c=True):
    """Cython ``np.ndarray[dtype, ndim=n, mode='c'] name not None`` semantics."""
    if x is None:
        raise TypeError("Argument '%s' must not be None" % name)
    if not isinstance(x, np.ndarray):
        raise TypeError("Argument '%s' has incorrect type (expected numpy.ndarray, got %s)"
                        % (name, type(x).__name__))
    want = np.dtype(dtype)
    if x.dtype != want:
        raise ValueError("Buffer dtype mismatch, expected '%s' but got '%s'"
                         % (_CNAME.get(want, str(want)), _CNAME.get(x.dtype, str(x.dtype))))
    if x.ndim != ndim:
        raise ValueError("Buffer has wrong number of dimensions (expected %d, got %d)" % (ndim, x.ndim))
    if need_c and not x.flags.c_contiguous:
        raise ValueError("ndarray is not C-contiguous")
    return x


def _cell(cellDims, nF):
    c = _buf(cellDims, "cellDims", np.float32, 2, need_c=False)
    if c.shape[1] != 3:
        raise ValueError("cellDims must have shape (nFrames, 3), got %r" % (c.shape,))
    if c.shape[0] != nF:
        raise ValueError("cellDims has %d frames but the coordinates have %d" % (c.shape[0], nF))
    return np.ascontiguousarray(c)


def _coords(x, name):
    x = _buf(x, name, np.float32, 3)
    if x.shape[2] != 3:
        raise ValueError("%s must have shape (nAtoms, nFrames, 3), got %r" % (name, x.shape))
    return x


def _p(a, t):
    return a.ctypes.data_as(t)


def py_getParallelBackend():
    """2, the reference's value for its CUDA build (lib/cuda/src/getParallelBackend.cpp:4-7), so
    ``NAMDDCD.getWithin`` never takes the kd-tree branch (dataParsers/dcdParser.py:181)."""
    return int(_capi.load().nda_get_parallel_backend())


def py_getWithin(allAtoms, refSel, outSel, out, cellDims, distance):
    """out[a, f] = 1 for every outSel atom a within ``distance`` of any refSel atom in frame f
    (lib/openmp/src/getWithin.cpp:9-90).  ``out`` int32 (nAtoms, nFrames) is only ever set, never
    cleared, exactly like the reference."""
    allAtoms = _coords(allAtoms, "allAtoms")
    refSel = _buf(refSel, "refSel", np.int32, 1)
    outSel = _buf(outSel, "outSel", np.int32, 1)
    out = _buf(out, "out", np.int32, 2)
    nA, nF = allAtoms.shape[:2]
    if out.shape != (nA, nF):
        raise ValueError("out must have shape %r, got %r" % ((nA, nF), out.shape))
    cell = _cell(cellDims, nF)
    L = _capi.load()
    _capi.check(L.nda_get_within(_p(allAtoms, c_f32p), nA, nF, _p(refSel, c_i32p), refSel.size,
                                 _p(outSel, c_i32p), outSel.size, _p(out, c_i32p), _p(cell, c_f32p),
                                 float(distance)))


def py_getDistances(sel1, sel2, out, cellDims, sameSel):
    """out[i, j] = mean over frames of the minimum-image distance between sel1[i] and sel2[j]
    (lib/openmp/src/getDistances.cpp:9-63; mean as documented at dataParsers/dcdParser.py:103-105).
    ``sameSel`` fills the matrix symmetrically with a zero diagonal."""
    sel1 = _coords(sel1, "sel1")
    sel2 = _coords(sel2, "sel2")
    out = _buf(out, "out", np.float32, 2)
    nF = sel1.shape[1]
    if sel2.shape[1] != nF:
        raise ValueError("sel1 and sel2 have different frame counts (%d, %d)" % (nF, sel2.shape[1]))
    if out.shape != (sel1.shape[0], sel2.shape[0]):
        raise ValueError("out must have shape %r, got %r" % ((sel1.shape[0], sel2.shape[0]), out.shape))
    cell = _cell(cellDims, nF)
    L = _capi.load()
    _capi.check(L.nda_get_distances(_p(sel1, c_f32p), sel1.shape[0], _p(sel2, c_f32p), sel2.shape[0],
                                    _p(out, c_f32p), _p(cell, c_f32p), nF, int(sameSel)))


def py_getRadialNbrDensity(sel1, sel2, out, cellDims, sameSel, maxR, dr):
    """out[k] = (number of (sel1, sel2, frame) pairs with (int)(dist/dr) == k) / nFrames
    (lib/openmp/src/getRadialNbrDensity.cpp:8-52); ``maxR`` is ignored as in the reference."""
    sel1 = _coords(sel1, "sel1")
    sel2 = _coords(sel2, "sel2")
    out = _buf(out, "out", np.float32, 1)
    nF = sel1.shape[1]
    if sel2.shape[1] != nF:
        raise ValueError("sel1 and sel2 have different frame counts (%d, %d)" % (nF, sel2.shape[1]))
    cell = _cell(cellDims, nF)
    L = _capi.load()
    _capi.check(L.nda_get_radial_nbr_density(_p(sel1, c_f32p), sel1.shape[0], _p(sel2, c_f32p), sel2.shape[0],
                                             _p(out, c_f32p), out.size, _p(cell, c_f32p), nF, int(sameSel),
                                             float(maxR), float(dr)))


# ---- extensions (not in the reference module) ---------------------------------------------------

def py_getRadialNbrCounts(sel1, sel2, cellDims, sameSel, outSize, dr, groupId=None, nGroups=1,
                          frameBegin=0, frameEnd=None):
    """Exact integer pair counts: uint64 (outSize,) or (nGroups, outSize) when ``groupId`` (int32 per
    sel2 atom) is given.  ``frameBegin:frameEnd`` restricts the frames (multi-GPU shard)."""
    sel1 = _coords(sel1, "sel1")
    sel2 = _coords(sel2, "sel2")
    nF = sel1.shape[1]
    if sel2.shape[1] != nF:
        raise ValueError("sel1 and sel2 have different frame counts (%d, %d)" % (nF, sel2.shape[1]))
    cell = _cell(cellDims, nF)
    frameEnd = nF if frameEnd is None else int(frameEnd)
    if groupId is None:
        gp, nGroups = None, 1
    else:
        groupId = _buf(groupId, "groupId", np.int32, 1)
        if groupId.size != sel2.shape[0]:
            raise ValueError("groupId must have one entry per sel2 atom")
        gp = _p(groupId, c_i32p)
    counts = np.zeros((nGroups, int(outSize)), np.uint64)
    L = _capi.load()
    _capi.check(L.nda_get_radial_nbr_counts(_p(sel1, c_f32p), sel1.shape[0], _p(sel2, c_f32p), sel2.shape[0],
                                            gp, int(nGroups), _p(counts, c_u64p), None, int(outSize),
                                            _p(cell, c_f32p), nF, int(frameBegin), frameEnd, int(sameSel),
                                            float(dr)))
    return counts if groupId is not None else counts[0]


def py_getRadialNbrDensityGrouped(sel1, sel2, groupId, out, cellDims, sameSel, maxR, dr):
    """One launch for what dataAnalysis/RadialDensity.py:84-103 does with one reference call per
    residue: out float32 (nGroups, outSize), row g = histogram over the sel2 atoms with groupId g."""
    sel1 = _coords(sel1, "sel1")
    sel2 = _coords(sel2, "sel2")
    groupId = _buf(groupId, "groupId", np.int32, 1)
    out = _buf(out, "out", np.float32, 2)
    nF = sel1.shape[1]
    if sel2.shape[1] != nF:
        raise ValueError("sel1 and sel2 have different frame counts (%d, %d)" % (nF, sel2.shape[1]))
    if groupId.size != sel2.shape[0]:
        raise ValueError("groupId must have one entry per sel2 atom")
    cell = _cell(cellDims, nF)
    L = _capi.load()
    _capi.check(L.nda_get_radial_nbr_counts(_p(sel1, c_f32p), sel1.shape[0], _p(sel2, c_f32p), sel2.shape[0],
                                            _p(groupId, c_i32p), out.shape[0], None, _p(out, c_f32p), out.shape[1],
                                            _p(cell, c_f32p), nF, 0, nF, int(sameSel), float(dr)))


def py_getDistancesSum(sel1, sel2, cellDims, sameSel, frameBegin=0, frameEnd=None):
    """float64 (n1, n2) sum over frames of the per-frame float32 distance (multi-GPU partial)."""
    sel1 = _coords(sel1, "sel1")
    sel2 = _coords(sel2, "sel2")
    nF = sel1.shape[1]
    if sel2.shape[1] != nF:
        raise ValueError("sel1 and sel2 have different frame counts (%d, %d)" % (nF, sel2.shape[1]))
    cell = _cell(cellDims, nF)
    frameEnd = nF if frameEnd is None else int(frameEnd)
    s = np.zeros((sel1.shape[0], sel2.shape[0]), np.float64)
    L = _capi.load()
    _capi.check(L.nda_get_distances_sum(_p(sel1, c_f32p), sel1.shape[0], _p(sel2, c_f32p), sel2.shape[0],
                                        _p(s, c_f64p), _p(cell, c_f32p), nF, int(frameBegin), frameEnd, int(sameSel)))
    return s


def py_getWithinBits(allAtoms, refSel, outSel, cellDims, distance, frameBegin=0, frameEnd=None):
    """The raw cut-off bitmask: uint32 (frameEnd-frameBegin, ceil(len(outSel)/32))."""
    allAtoms = _coords(allAtoms, "allAtoms")
    refSel = _buf(refSel, "refSel", np.int32, 1)
    outSel = _buf(outSel, "outSel", np.int32, 1)
    nA, nF = allAtoms.shape[:2]
    cell = _cell(cellDims, nF)
    frameEnd = nF if frameEnd is None else int(frameEnd)
    bits = np.zeros((max(frameEnd - frameBegin, 0), (outSel.size + 31) // 32), np.uint32)
    L = _capi.load()
    _capi.check(L.nda_get_within_bits(_p(allAtoms, c_f32p), nA, nF, _p(refSel, c_i32p), refSel.size,
                                      _p(outSel, c_i32p), outSel.size, _p(bits, c_u32p), None,
                                      _p(cell, c_f32p), float(distance), int(frameBegin), frameEnd))
    return bits


def py_getWithinLists(allAtoms, refSel, outSel, cellDims, distance, frameBegin=0, frameEnd=None):
    """Per-frame lists of the outSel atoms within ``distance`` of refSel, in CSR form:
    ``(offsets int64 (nFrames+1,), atoms int32 (total,))`` with frame f's atoms at
    ``atoms[offsets[f]:offsets[f+1]]`` in outSel order.  This is what the reference's multi-frame
    selection loop builds one frame at a time with np.argwhere on the int32 matrix
    (NAMDAnalyzer/selection/selParser.py:117-126); here it is decoded from the bitmask in one
    vectorised pass."""
    outSel = _buf(outSel, "outSel", np.int32, 1)
    bits = py_getWithinBits(allAtoms, refSel, outSel, cellDims, distance, frameBegin, frameEnd)
    nFc = bits.shape[0]
    if nFc == 0 or outSel.size == 0:
        return np.zeros(nFc + 1, np.int64), np.zeros(0, np.int32)
    flags = np.unpackbits(bits.view(np.uint8), axis=1, bitorder="little")[:, :outSel.size]
    f_idx, a_idx = np.nonzero(flags)                       # row-major: sorted by frame, then outSel order
    offsets = np.zeros(nFc + 1, np.int64)
    np.cumsum(np.bincount(f_idx, minlength=nFc), out=offsets[1:])
    return offsets, outSel[a_idx]


# ---- SURVEY row f-3 ---------------------------------------------------------------------------------

def py_setWaterDistPBC(water, prot, cellDims, nbrWAtoms):
    """Shifts every water molecule IN PLACE by the periodic image vector of its nearest protein atom
    (NAMDAnalyzer/lib/libFunc.pyx:206-213, lib/openmp/src/setWaterDistPBC.cpp).  ``water`` float32
    (nMolecules*nbrWAtoms, nFrames, 3), atoms of a molecule adjacent, first atom = reference point."""
    water = _coords(water, "water")
    prot = _coords(prot, "prot")
    nF = water.shape[1]
    if prot.shape[1] != nF:
        raise ValueError("water and prot have different frame counts (%d, %d)" % (nF, prot.shape[1]))
    nbrWAtoms = int(nbrWAtoms)
    if nbrWAtoms < 1 or water.shape[0] % nbrWAtoms:
        raise ValueError("water.shape[0] must be a multiple of nbrWAtoms")
    if not water.flags.writeable:
        raise ValueError("water is modified in place and must be writeable")
    cell = _cell(cellDims, nF)
    L = _capi.load()
    _capi.check(L.nda_set_water_dist_pbc(_p(water, c_f32p), water.shape[0] // nbrWAtoms, _p(prot, c_f32p),
                                         prot.shape[0], _p(cell, c_f32p), nF, nbrWAtoms))


def py_waterOrientAtSurface(waterO, watVec, prot, out, cellDims, minR, maxR, maxN):
    """Overwrites ``waterO`` (nWater, nFrames, 3) IN PLACE with (cos angle to the surface normal, 0/1
    flag, nearest distance) -- NAMDAnalyzer/lib/libFunc.pyx:189-203 (``out`` is unused there too),
    lib/openmp/src/waterOrientAtSurface.cpp."""
    waterO = _coords(waterO, "waterO")
    watVec = _coords(watVec, "watVec")
    prot = _coords(prot, "prot")
    _buf(out, "out", np.float32, 2)
    nF = waterO.shape[1]
    if watVec.shape != waterO.shape:
        raise ValueError("watVec must have the shape of waterO")
    if prot.shape[1] != nF:
        raise ValueError("waterO and prot have different frame counts (%d, %d)" % (nF, prot.shape[1]))
    if not waterO.flags.writeable:
        raise ValueError("waterO is modified in place and must be writeable")
    cell = _cell(cellDims, nF)
    L = _capi.load()
    _capi.check(L.nda_water_orient_at_surface(_p(waterO, c_f32p), waterO.shape[0], _p(watVec, c_f32p),
                                              _p(prot, c_f32p), prot.shape[0], _p(out, c_f32p), _p(cell, c_f32p),
                                              nF, float(minR), float(maxR), int(maxN)))


# ---- SURVEY row f-4 ---------------------------------------------------------------------------------

def py_getHBCorr(acceptors, nbrFrames, donors, hydrogens, cellDims, out, maxTime, step, nbrTimeOri, maxR, minAngle,
                 continuous):
    """Hydrogen-bond autocorrelation, ACCUMULATED into ``out`` float32 (nbrFrames,) like the reference
    (NAMDAnalyzer/lib/libFunc.pyx:144-155, lib/openmp/src/getHydrogenBonds.cpp:9-91)."""
    acceptors = _coords(acceptors, "acceptors")
    donors = _coords(donors, "donors")
    hydrogens = _coords(hydrogens, "hydrogens")
    out = _buf(out, "out", np.float32, 1)
    nF = int(nbrFrames)
    for name, x in (("acceptors", acceptors), ("donors", donors), ("hydrogens", hydrogens)):
        if x.shape[1] != nF:
            raise ValueError("%s has %d frames, nbrFrames is %d" % (name, x.shape[1], nF))
    if out.size != nF:
        raise ValueError("out must have nbrFrames entries")
    cell = _cell(cellDims, nF)
    L = _capi.load()
    _capi.check(L.nda_get_hb_corr(_p(acceptors, c_f32p), acceptors.shape[0], nF, _p(donors, c_f32p), donors.shape[0],
                                  _p(hydrogens, c_f32p), hydrogens.shape[0], _p(cell, c_f32p), _p(out, c_f32p),
                                  int(maxTime), int(step), int(nbrTimeOri), float(maxR), float(minAngle), int(continuous)))


def py_getHBCounts(acceptors, donors, hydrogens, cellDims, maxR, minAngle, continuous):
    """the exact integer counts behind py_getHBCorr: uint64 (nFrames,)"""
    acceptors = _coords(acceptors, "acceptors")
    donors = _coords(donors, "donors")
    hydrogens = _coords(hydrogens, "hydrogens")
    nF = acceptors.shape[1]
    cell = _cell(cellDims, nF)
    counts = np.zeros(nF, np.uint64)
    L = _capi.load()
    _capi.check(L.nda_get_hb_counts(_p(acceptors, c_f32p), acceptors.shape[0], nF, _p(donors, c_f32p), donors.shape[0],
                                    _p(hydrogens, c_f32p), hydrogens.shape[0], _p(cell, c_f32p), _p(counts, c_u64p),
                                    float(maxR), float(minAngle), int(continuous)))
    return counts


# ---- DCD ingest ---------------------------------------------------------------------

def _byteorder_char(byteorder):
    if isinstance(byteorder, int):
        byteorder = chr(byteorder)
    if isinstance(byteorder, (bytes, bytearray)):
        byteorder = byteorder.decode()
    return byteorder.encode()[:1] or b"<"


def py_getDCDCoor(fileName, frames, nbrAtoms, selAtoms, dims, cell, startPos, outArr, byteorder):
    """Reads DCD coordinate records into ``outArr`` float32 (len(selAtoms), len(frames), len(dims));
    returns the reference's error code (NAMDAnalyzer/lib/libFunc.pyx:61-75,
    lib/common/src/dcdReader.cpp:28-98 -- which does not compile on Linux)."""
    frames = _buf(frames, "frames", np.int32, 1)
    selAtoms = _buf(selAtoms, "selAtoms", np.int32, 1)
    dims = _buf(dims, "dims", np.int32, 1)
    startPos = _buf(startPos, "startPos", np.int64, 1)
    outArr = _buf(outArr, "outArr", np.float32, 3)
    if outArr.shape != (selAtoms.size, frames.size, dims.size):
        raise ValueError("outArr must have shape (len(selAtoms), len(frames), len(dims))")
    if frames.size and (frames.min() < 0 or frames.max() >= startPos.size):
        return -2
    name = bytes(fileName) if isinstance(fileName, (bytes, bytearray)) else str(fileName).encode()
    L = _capi.load()
    return int(L.nda_dcd_get_coor(name, _p(frames, c_i32p), frames.size, int(nbrAtoms), _p(selAtoms, c_i32p),
                                  selAtoms.size, _p(dims, c_i32p), dims.size, int(bool(cell)),
                                  startPos.ctypes.data_as(ctypes.POINTER(ctypes.c_longlong)), _p(outArr, c_f32p),
                                  _byteorder_char(byteorder)))


def py_getDCDCell(fileName, frames, startPos, outArr, byteorder):
    """Reads the six unit-cell doubles of each frame into ``outArr`` float64 (len(frames), 6)
    (NAMDAnalyzer/lib/libFunc.pyx:82-95, lib/common/src/dcdCellReader.cpp:26-70)."""
    frames = _buf(frames, "frames", np.int32, 1)
    startPos = _buf(startPos, "startPos", np.int64, 1)
    outArr = _buf(outArr, "outArr", np.float64, 2)
    if outArr.shape != (frames.size, 6):
        raise ValueError("outArr must have shape (len(frames), 6)")
    if frames.size and (frames.min() < 0 or frames.max() >= startPos.size):
        return -2
    name = bytes(fileName) if isinstance(fileName, (bytes, bytearray)) else str(fileName).encode()
    L = _capi.load()
    return int(L.nda_dcd_get_cell(name, _p(frames, c_i32p), frames.size,
                                  startPos.ctypes.data_as(ctypes.POINTER(ctypes.c_longlong)), _p(outArr, c_f64p),
                                  _byteorder_char(byteorder)))
