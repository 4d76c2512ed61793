"""Host glue of ``NAMDDCD.getDistances`` / ``NAMDDCD.getWithin``
(package/NAMDAnalyzer/dataParsers/dcdParser.py:96-194) over the CUDA operators.

``PairOps`` is a mixin: the host class supplies ``selection(text)``, ``dcdData[atoms, frames]``
and ``cellDims[frames]`` exactly as NAMDAnalyzer's ``Dataset`` does.  Same arguments, same return
values; the kd-tree branch (dcdParser.py:185-188) does not exist because
``py_getParallelBackend() == 2``.
"""
import numpy as np

from ..lib.pylibFuncs import py_getWithin, py_getDistances, py_getParallelBackend


class PairOps:
    parallelBackend = 2

    def _init_pair_ops(self):
        self.parallelBackend = py_getParallelBackend()          # dcdParser.py:62-65

    def getDistances(self, sel1, sel2=None, frame=0):
        """Pair-wise distances between sel1 (rows) and sel2 (columns); several frames -> the mean
        over the frames (dcdParser.py:96-145)."""
        sameSel = 0
        if isinstance(sel1, str) and isinstance(sel2, str):
            if set(sel1.split(' ')) == set(sel2.split(' ')):    # dcdParser.py:112-115
                sameSel = 1
        if isinstance(sel1, str):
            sel1 = self.selection(sel1)
        if isinstance(sel2, str):
            sel2 = sel1 if sameSel else self.selection(sel2)
        if sel2 is None:                                        # dcdParser.py:133-135
            sel2 = sel1
            sameSel = 1
        i1 = np.asarray(sel1.getIndices() if hasattr(sel1, "getIndices") else sel1)
        i2 = np.asarray(sel2.getIndices() if hasattr(sel2, "getIndices") else sel2)
        out = np.zeros((i1.shape[0], i2.shape[0]), dtype='float32')
        cellDims = np.ascontiguousarray(self.cellDims[frame], dtype='float32')
        c1 = np.ascontiguousarray(self.dcdData[i1, frame], dtype='float32')
        c2 = c1 if sameSel else np.ascontiguousarray(self.dcdData[i2, frame], dtype='float32')
        py_getDistances(c1, c2, out, cellDims, sameSel)
        return out

    def getWithin(self, distance, refSel, outSel='all', frame=0):
        """Atoms of outSel within ``distance`` of refSel: index vector for one frame, 0/1 matrix
        (nAtoms, nFrames) for several (dcdParser.py:147-194)."""
        if isinstance(refSel, str):
            refSel = self.selection(refSel)
        if isinstance(outSel, str):
            outSel = self.selection(outSel)
        refIdx = np.asarray(refSel.getIndices() if hasattr(refSel, "getIndices") else refSel)
        outIdx = np.asarray(outSel.getIndices() if hasattr(outSel, "getIndices") else outSel)
        cellDims = self.cellDims[frame]
        allAtoms = np.ascontiguousarray(self.dcdData[:, frame], dtype='float32')
        keepIdx = np.zeros((allAtoms.shape[0], allAtoms.shape[1]), dtype='int32')
        py_getWithin(allAtoms, refIdx.astype('int32'), outIdx.astype('int32'), keepIdx, cellDims, distance)
        if keepIdx.shape[1] == 1:                               # dcdParser.py:190-192
            keepIdx = keepIdx.flatten()
            keepIdx = np.argwhere(keepIdx)[:, 0]
        return keepIdx
