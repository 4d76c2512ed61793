"""oracle.port -- TEST INFRASTRUCTURE.  ctypes loader for ``minimage_port.c`` (our
plain-C, strict-IEEE restatement of lib/openmp/src/{getWithin,getDistances,
getRadialNbrDensity}.cpp).  Built by ``make -C oracle port`` /
``__graft_entry__.build()`` into ``oracle/_build/libminimage_port.so``."""
import ctypes
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libminimage_port.so")
_f32p = ctypes.POINTER(ctypes.c_float)
_f64p = ctypes.POINTER(ctypes.c_double)
_i32p = ctypes.POINTER(ctypes.c_int)
_u64p = ctypes.POINTER(ctypes.c_uint64)
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "minimage_port.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "port"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        L.port_get_within.argtypes = [_f32p, ctypes.c_int, ctypes.c_int, _i32p, ctypes.c_int, _i32p,
                                      ctypes.c_int, _i32p, _f32p, ctypes.c_float, ctypes.c_int]
        L.port_get_within.restype = None
        L.port_get_distances.argtypes = [_f32p, ctypes.c_int, _f32p, ctypes.c_int, _f64p, _f32p, _f32p,
                                         ctypes.c_int, ctypes.c_int]
        L.port_get_distances.restype = None
        L.port_get_rdf_counts.argtypes = [_f32p, ctypes.c_int, _f32p, ctypes.c_int, _u64p, ctypes.c_int,
                                          _i32p, ctypes.c_int, _f32p, ctypes.c_int, ctypes.c_int,
                                          ctypes.c_float, ctypes.c_float]
        L.port_get_rdf_counts.restype = None
        L.port_set_water_dist_pbc.argtypes = [_f32p, ctypes.c_int, _f32p, ctypes.c_int, _f32p, ctypes.c_int,
                                              ctypes.c_int, ctypes.c_int]
        L.port_set_water_dist_pbc.restype = None
        L.port_water_orient_at_surface.argtypes = [_f32p, ctypes.c_int, _f32p, _f32p, ctypes.c_int, _f32p,
                                                   ctypes.c_int, ctypes.c_float, ctypes.c_float, ctypes.c_int]
        L.port_water_orient_at_surface.restype = None
        L.port_get_hb_corr.argtypes = [_f32p, ctypes.c_int, ctypes.c_int, _f32p, ctypes.c_int, _f32p, ctypes.c_int,
                                       _f32p, _u64p, ctypes.c_float, ctypes.c_float, ctypes.c_int]
        L.port_get_hb_corr.restype = None
        L.port_pair_r2.argtypes = [_f32p, _f32p, _f32p]
        L.port_pair_r2.restype = ctypes.c_float
        L.port_num_threads.restype = ctypes.c_int
        _lib = L
    return _lib


def num_threads():
    return lib().port_num_threads()


def within(allAtoms, refSel, outSel, cellDims, distance, literal_early_out=True, out=None):
    """-> int32 [nAtoms][nFrames]; any number of frames."""
    allAtoms = np.ascontiguousarray(allAtoms, np.float32)
    cellDims = np.ascontiguousarray(cellDims, np.float32)
    refSel = np.ascontiguousarray(refSel, np.int32)
    outSel = np.ascontiguousarray(outSel, np.int32)
    nA, nF = allAtoms.shape[:2]
    if out is None:
        out = np.zeros((nA, nF), np.int32)
    lib().port_get_within(allAtoms.ctypes.data_as(_f32p), nA, nF,
                          refSel.ctypes.data_as(_i32p), refSel.size,
                          outSel.ctypes.data_as(_i32p), outSel.size,
                          out.ctypes.data_as(_i32p), cellDims.ctypes.data_as(_f32p),
                          ctypes.c_float(distance), int(bool(literal_early_out)))
    return out


def distances_sum(sel1, sel2, cellDims, sameSel, want_last=False):
    """-> float64 [n1][n2] sum over frames of the per-frame float32 distance
    (+ the last frame's float32 distances if want_last)."""
    sel1 = np.ascontiguousarray(sel1, np.float32)
    sel2 = np.ascontiguousarray(sel2, np.float32)
    cellDims = np.ascontiguousarray(cellDims, np.float32)
    n1, nF = sel1.shape[:2]
    n2 = sel2.shape[0]
    s = np.zeros((n1, n2), np.float64)
    last = np.zeros((n1, n2), np.float32) if want_last else None
    lib().port_get_distances(sel1.ctypes.data_as(_f32p), n1, sel2.ctypes.data_as(_f32p), n2,
                             s.ctypes.data_as(_f64p),
                             last.ctypes.data_as(_f32p) if want_last else None,
                             cellDims.ctypes.data_as(_f32p), nF, int(sameSel))
    return (s, last) if want_last else s


def rdf_counts(sel1, sel2, cellDims, sameSel, outSize, dr, maxR=0.0, groupId=None, nGroups=1):
    """-> uint64 [outSize] (or [nGroups][outSize]) exact pair counts over all frames."""
    sel1 = np.ascontiguousarray(sel1, np.float32)
    sel2 = np.ascontiguousarray(sel2, np.float32)
    cellDims = np.ascontiguousarray(cellDims, np.float32)
    n1, nF = sel1.shape[:2]
    n2 = sel2.shape[0]
    if groupId is None:
        counts = np.zeros(outSize, np.uint64)
        gp = None
        nGroups = 1
    else:
        groupId = np.ascontiguousarray(groupId, np.int32)
        assert groupId.size == n2 and groupId.min() >= 0 and groupId.max() < nGroups
        counts = np.zeros((nGroups, outSize), np.uint64)
        gp = groupId.ctypes.data_as(_i32p)
    lib().port_get_rdf_counts(sel1.ctypes.data_as(_f32p), n1, sel2.ctypes.data_as(_f32p), n2,
                              counts.ctypes.data_as(_u64p), outSize, gp, nGroups,
                              cellDims.ctypes.data_as(_f32p), nF, int(sameSel),
                              ctypes.c_float(maxR), ctypes.c_float(dr))
    return counts


def pair_r2(a, b, cell):
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    cell = np.ascontiguousarray(cell, np.float32)
    return float(lib().port_pair_r2(a.ctypes.data_as(_f32p), b.ctypes.data_as(_f32p),
                                    cell.ctypes.data_as(_f32p)))


def set_water_dist_pbc(water, prot, cellDims, nbrWAtoms, reset_per_water=True):
    """-> shifted copy of water [nMol*nbrWAtoms][nF][3] (setWaterDistPBC.cpp:9-77)"""
    w = np.array(water, np.float32, order="C")
    prot = np.ascontiguousarray(prot, np.float32)
    cell = np.ascontiguousarray(cellDims, np.float32)
    lib().port_set_water_dist_pbc(w.ctypes.data_as(_f32p), w.shape[0] // nbrWAtoms, prot.ctypes.data_as(_f32p),
                                  prot.shape[0], cell.ctypes.data_as(_f32p), w.shape[1], int(nbrWAtoms),
                                  int(bool(reset_per_water)))
    return w


def water_orient_at_surface(waterO, watVec, prot, cellDims, minR, maxR, maxN):
    """-> overwritten copy of waterO (waterOrientAtSurface.cpp:10-112)"""
    w = np.array(waterO, np.float32, order="C")
    v = np.ascontiguousarray(watVec, np.float32)
    prot = np.ascontiguousarray(prot, np.float32)
    cell = np.ascontiguousarray(cellDims, np.float32)
    lib().port_water_orient_at_surface(w.ctypes.data_as(_f32p), w.shape[0], v.ctypes.data_as(_f32p),
                                       prot.ctypes.data_as(_f32p), prot.shape[0], cell.ctypes.data_as(_f32p),
                                       w.shape[1], ctypes.c_float(minR), ctypes.c_float(maxR), int(maxN))
    return w


def hb_corr(acceptors, donors, hydrogens, cellDims, maxR, minAngle, continuous):
    """-> uint64 [nFrames] H-bond autocorrelation counts (getHydrogenBonds.cpp:9-91)"""
    a = np.ascontiguousarray(acceptors, np.float32)
    d = np.ascontiguousarray(donors, np.float32)
    h = np.ascontiguousarray(hydrogens, np.float32)
    cell = np.ascontiguousarray(cellDims, np.float32)
    nF = a.shape[1]
    counts = np.zeros(nF, np.uint64)
    lib().port_get_hb_corr(a.ctypes.data_as(_f32p), a.shape[0], nF, d.ctypes.data_as(_f32p), d.shape[0],
                           h.ctypes.data_as(_f32p), h.shape[0], cell.ctypes.data_as(_f32p),
                           counts.ctypes.data_as(_u64p), ctypes.c_float(maxR), ctypes.c_float(minAngle),
                           int(continuous))
    return counts
