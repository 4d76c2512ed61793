"""oracle.ref -- TEST INFRASTRUCTURE.  ctypes driver for the reference's own OpenMP
sources compiled by ``oracle/Makefile`` into ``oracle/_ref/libref_{strict,fast}.so``.

The reference declares its functions without ``extern "C"``
(package/NAMDAnalyzer/lib/libFunc.h:37-48,57), so they are bound by their mangled
names.  Every call is made ONE FRAME AT A TIME:

* multi-frame ``getWithin`` reads out of bounds and segfaults
  (lib/openmp/src/getWithin.cpp:55);
* multi-frame ``getDistances`` assigns instead of accumulating
  (lib/openmp/src/getDistances.cpp:57);
* with one frame ``getRadialNbrDensity`` adds exactly 1.0 per pair
  (lib/openmp/src/getRadialNbrDensity.cpp:44), so float32 bins are exact integers
  (below 2**24 per bin per frame).
"""
import ctypes
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_f32p = ctypes.POINTER(ctypes.c_float)
_i32p = ctypes.POINTER(ctypes.c_int)

_SYMS = {
    "getWithin": ("_Z9getWithinPfiiPiiS0_iS0_S_f",
                  [_f32p, ctypes.c_int, ctypes.c_int, _i32p, ctypes.c_int, _i32p, ctypes.c_int,
                   _i32p, _f32p, ctypes.c_float]),
    "getDistances": ("_Z12getDistancesPfiS_iS_S_ii",
                     [_f32p, ctypes.c_int, _f32p, ctypes.c_int, _f32p, _f32p, ctypes.c_int, ctypes.c_int]),
    "getRadialNbrDensity": ("_Z19getRadialNbrDensityPfiS_iS_iS_iiff",
                            [_f32p, ctypes.c_int, _f32p, ctypes.c_int, _f32p, ctypes.c_int, _f32p,
                             ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_float]),
    "getParallelBackend": ("_Z18getParallelBackendv", []),
    # SURVEY row f-4 (lib/openmp/src/getHydrogenBonds.cpp)
    "getHBCorr": ("_Z9getHBCorrPfiiS_iS_iS_S_iiiffi",
                  [_f32p, ctypes.c_int, ctypes.c_int, _f32p, ctypes.c_int, _f32p, ctypes.c_int, _f32p, _f32p,
                   ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_float, ctypes.c_int]),
    # SURVEY row f-3 (lib/openmp/src/setWaterDistPBC.cpp, waterOrientAtSurface.cpp)
    "setWaterDistPBC": ("_Z15setWaterDistPBCPfiS_iS_ii",
                        [_f32p, ctypes.c_int, _f32p, ctypes.c_int, _f32p, ctypes.c_int, ctypes.c_int]),
    "waterOrientAtSurface": ("_Z20waterOrientAtSurfacePfiS_S_iS_S_iffi",
                             [_f32p, ctypes.c_int, _f32p, _f32p, ctypes.c_int, _f32p, _f32p, ctypes.c_int,
                              ctypes.c_float, ctypes.c_float, ctypes.c_int]),
}


def lib_path(kind="strict"):
    return os.path.join(_HERE, "_ref", "libref_%s.so" % kind)


def available(kind="strict"):
    return os.path.exists(lib_path(kind))


class RefLib:
    """One compiled build of the reference ('strict' = parity, 'fast' = shipped flags,
    'cuda' = the reference's legacy CUDA build recompiled for sm_100a, bench comparison only)."""

    def __init__(self, kind="strict"):
        self.kind = kind
        self.lib = ctypes.CDLL(lib_path(kind))
        self.fn = {}
        for name, (sym, argtypes) in _SYMS.items():
            try:
                f = getattr(self.lib, sym)
            except AttributeError:           # e.g. the legacy CUDA build holds the three hot-path functions only
                continue
            f.argtypes = argtypes
            f.restype = ctypes.c_int if name == "getParallelBackend" else None
            self.fn[name] = f

    @staticmethod
    def _quiet():
        """The reference printf()s a progress line per frame; silence fd 1 around calls."""
        class _Q:
            def __enter__(s):
                import sys
                sys.stdout.flush()
                s.saved = os.dup(1)
                s.null = os.open(os.devnull, os.O_WRONLY)
                os.dup2(s.null, 1)
            def __exit__(s, *a):
                # the reference leaves text in the C stdio buffer; flush it into /dev/null
                ctypes.CDLL(None).fflush(None)
                os.dup2(s.saved, 1)
                os.close(s.saved)
                os.close(s.null)
        return _Q()

    def parallel_backend(self):
        return self.fn["getParallelBackend"]()

    # -- per-frame drivers ---------------------------------------------------------
    def within(self, allAtoms, refSel, outSel, cellDims, distance):
        """-> int32 [nAtoms][nFrames] mask, per-frame calls of the reference getWithin."""
        allAtoms = np.ascontiguousarray(allAtoms, np.float32)
        cellDims = np.ascontiguousarray(cellDims, np.float32)
        refSel = np.ascontiguousarray(refSel, np.int32)
        outSel = np.ascontiguousarray(outSel, np.int32)
        nA, nF = allAtoms.shape[:2]
        out = np.zeros((nA, nF), np.int32)
        with self._quiet():
            for f in range(nF):
                xyz = np.ascontiguousarray(allAtoms[:, f:f + 1, :])
                cell = np.ascontiguousarray(cellDims[f:f + 1])
                o = np.zeros((nA, 1), np.int32)
                self.fn["getWithin"](xyz.ctypes.data_as(_f32p), nA, 1,
                                     refSel.ctypes.data_as(_i32p), refSel.size,
                                     outSel.ctypes.data_as(_i32p), outSel.size,
                                     o.ctypes.data_as(_i32p), cell.ctypes.data_as(_f32p),
                                     ctypes.c_float(distance))
                out[:, f] = o[:, 0]
        return out

    def distances_per_frame(self, sel1, sel2, cellDims, sameSel):
        """-> float32 [nFrames][n1][n2], the reference getDistances called frame by frame
        (entries the reference does not visit stay 0)."""
        sel1 = np.ascontiguousarray(sel1, np.float32)
        sel2 = np.ascontiguousarray(sel2, np.float32)
        cellDims = np.ascontiguousarray(cellDims, np.float32)
        n1, nF = sel1.shape[:2]
        n2 = sel2.shape[0]
        res = np.zeros((nF, n1, n2), np.float32)
        with self._quiet():
            for f in range(nF):
                a = np.ascontiguousarray(sel1[:, f:f + 1, :])
                b = np.ascontiguousarray(sel2[:, f:f + 1, :])
                cell = np.ascontiguousarray(cellDims[f:f + 1])
                o = np.zeros((n1, n2), np.float32)
                self.fn["getDistances"](a.ctypes.data_as(_f32p), n1, b.ctypes.data_as(_f32p), n2,
                                        o.ctypes.data_as(_f32p), cell.ctypes.data_as(_f32p), 1, int(sameSel))
                res[f] = o
        return res

    def rdf_counts(self, sel1, sel2, cellDims, sameSel, outSize, dr, maxR=0.0, frames=None):
        """-> int64 [outSize] exact pair counts, the reference getRadialNbrDensity per frame."""
        same_obj = sel2 is sel1
        sel1 = np.ascontiguousarray(sel1, np.float32)
        sel2 = sel1 if same_obj else np.ascontiguousarray(sel2, np.float32)
        cellDims = np.ascontiguousarray(cellDims, np.float32)
        n1, nF = sel1.shape[:2]
        n2 = sel2.shape[0]
        tot = np.zeros(outSize, np.int64)
        with self._quiet():
            for f in (range(nF) if frames is None else frames):
                a = np.ascontiguousarray(sel1[:, f:f + 1, :])
                b = a if same_obj else np.ascontiguousarray(sel2[:, f:f + 1, :])
                cell = np.ascontiguousarray(cellDims[f:f + 1])
                o = np.zeros(outSize, np.float32)
                self.fn["getRadialNbrDensity"](a.ctypes.data_as(_f32p), n1, b.ctypes.data_as(_f32p), n2,
                                               o.ctypes.data_as(_f32p), outSize, cell.ctypes.data_as(_f32p),
                                               1, int(sameSel), ctypes.c_float(maxR), ctypes.c_float(dr))
                assert o.max(initial=0.0) < 2 ** 24, "per-frame bin count left float32's exact range"
                tot += o.astype(np.int64)
        return tot

    # -- row f-3 -----------------------------------------------------------------------
    def set_water_dist_pbc_literal(self, water, prot, cellDims, nbrWAtoms):
        """the reference call as its callers make it (one call, all molecules and frames): shifts a
        COPY of water and returns it.  closestDist is a running minimum over the whole call."""
        w = np.array(water, np.float32, order="C")
        prot = np.ascontiguousarray(prot, np.float32)
        cell = np.ascontiguousarray(cellDims, np.float32)
        with self._quiet():
            self.fn["setWaterDistPBC"](w.ctypes.data_as(_f32p), w.shape[0] // nbrWAtoms, prot.ctypes.data_as(_f32p),
                                       prot.shape[0], cell.ctypes.data_as(_f32p), w.shape[1], int(nbrWAtoms))
        return w

    def set_water_dist_pbc_per_molecule(self, water, prot, cellDims, nbrWAtoms):
        """one reference call per (molecule, frame): the per-molecule nearest-atom shift (what the
        legacy CUDA build and the docstring mean), free of the running-minimum carry-over"""
        water = np.ascontiguousarray(water, np.float32)
        prot = np.ascontiguousarray(prot, np.float32)
        cell = np.ascontiguousarray(cellDims, np.float32)
        out = water.copy()
        nW, nF = water.shape[0] // nbrWAtoms, water.shape[1]
        with self._quiet():
            for f in range(nF):
                p1 = np.ascontiguousarray(prot[:, f:f + 1])
                c1 = np.ascontiguousarray(cell[f:f + 1])
                for m in range(nW):
                    w1 = np.ascontiguousarray(water[m * nbrWAtoms:(m + 1) * nbrWAtoms, f:f + 1])
                    self.fn["setWaterDistPBC"](w1.ctypes.data_as(_f32p), 1, p1.ctypes.data_as(_f32p), p1.shape[0],
                                               c1.ctypes.data_as(_f32p), 1, int(nbrWAtoms))
                    out[m * nbrWAtoms:(m + 1) * nbrWAtoms, f] = w1[:, 0]
        return out

    def water_orient_at_surface(self, waterO, watVec, prot, cellDims, minR, maxR, maxN):
        """-> the overwritten copy of waterO: [..,0] cosine, [..,1] flag, [..,2] nearest distance"""
        w = np.array(waterO, np.float32, order="C")
        v = np.ascontiguousarray(watVec, np.float32)
        prot = np.ascontiguousarray(prot, np.float32)
        cell = np.ascontiguousarray(cellDims, np.float32)
        dummy = np.zeros((1, 1), np.float32)
        with self._quiet():
            self.fn["waterOrientAtSurface"](w.ctypes.data_as(_f32p), w.shape[0], v.ctypes.data_as(_f32p),
                                            prot.ctypes.data_as(_f32p), prot.shape[0], dummy.ctypes.data_as(_f32p),
                                            cell.ctypes.data_as(_f32p), w.shape[1], ctypes.c_float(minR),
                                            ctypes.c_float(maxR), int(maxN))
        return w

    # -- row f-4 -----------------------------------------------------------------------
    def hb_corr(self, acceptors, donors, hydrogens, cellDims, maxR, minAngle, continuous,
                maxTime=100, step=1, nbrTimeOri=25):
        """-> float32 [nFrames] as the reference accumulates it (exact integers below 2**24)"""
        a = np.ascontiguousarray(acceptors, np.float32)
        d = np.ascontiguousarray(donors, np.float32)
        h = np.ascontiguousarray(hydrogens, np.float32)
        cell = np.ascontiguousarray(cellDims, np.float32)
        nF = a.shape[1]
        out = np.zeros(nF, np.float32)
        with self._quiet():
            self.fn["getHBCorr"](a.ctypes.data_as(_f32p), a.shape[0], nF, d.ctypes.data_as(_f32p), d.shape[0],
                                 h.ctypes.data_as(_f32p), h.shape[0], cell.ctypes.data_as(_f32p),
                                 out.ctypes.data_as(_f32p), int(maxTime), int(step), int(nbrTimeOri),
                                 ctypes.c_float(maxR), ctypes.c_float(minAngle), int(continuous))
        return out

    # -- raw multi-frame calls, as the reference's callers make them (for timing) ---
    def raw_rdf(self, sel1, sel2, out, cellDims, sameSel, maxR, dr):
        with self._quiet():
            self.fn["getRadialNbrDensity"](sel1.ctypes.data_as(_f32p), sel1.shape[0],
                                           sel2.ctypes.data_as(_f32p), sel2.shape[0],
                                           out.ctypes.data_as(_f32p), out.size,
                                           cellDims.ctypes.data_as(_f32p), cellDims.shape[0],
                                           int(sameSel), ctypes.c_float(maxR), ctypes.c_float(dr))

    def raw_distances(self, sel1, sel2, out, cellDims, sameSel):
        with self._quiet():
            self.fn["getDistances"](sel1.ctypes.data_as(_f32p), sel1.shape[0],
                                    sel2.ctypes.data_as(_f32p), sel2.shape[0],
                                    out.ctypes.data_as(_f32p), cellDims.ctypes.data_as(_f32p),
                                    cellDims.shape[0], int(sameSel))

    def raw_within_frame(self, xyz1, refSel, outSel, out1, cell1, distance):
        """one-frame call on prepared (nAtoms,1,3)/(1,3)/(nAtoms,1) arrays (timing loop)."""
        self.fn["getWithin"](xyz1.ctypes.data_as(_f32p), xyz1.shape[0], 1,
                             refSel.ctypes.data_as(_i32p), refSel.size,
                             outSel.ctypes.data_as(_i32p), outSel.size,
                             out1.ctypes.data_as(_i32p), cell1.ctypes.data_as(_f32p),
                             ctypes.c_float(distance))
