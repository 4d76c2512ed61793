"""``RadialNumberDensity`` and ``ResidueWiseWaterDensity`` with the constructor arguments,
attributes (``radii``, ``density``, ``residues``) and normalisation of
package/NAMDAnalyzer/dataAnalysis/RadialDensity.py, over the CUDA operators.

``data`` is anything with NAMDAnalyzer's ``Dataset`` surface: ``selection(text)`` returning an
object with ``getIndices()`` / ``coordinates(frames)`` / ``getUniqueResidues()`` / ``getResidues()``,
``data[indices, frames]``, ``data.cellDims[frames]`` and ``data.nbrFrames``
(``namdanalyzer_b200.Dataset.ArrayDataset`` is a small in-memory one).

Differences from the reference, all on purpose:
 * ``ResidueWiseWaterDensity.compDensity`` issues ONE grouped launch
   (``py_getRadialNbrDensityGrouped``) instead of one call per residue, and reads the protein
   coordinates once (RadialDensity.py:84-103 re-reads the DCD per residue);
 * matplotlib is imported only inside ``plotDensity``;
 * the stale ``' '.join(frames)`` of RadialDensity.py:179-180 (SURVEY B-10) is not emulated.
"""
import re
import numpy as np

from ..lib.pylibFuncs import py_getRadialNbrDensity, py_getRadialNbrDensityGrouped


def _indices(sel):
    return np.asarray(sel.getIndices() if hasattr(sel, "getIndices") else sel)


class ResidueWiseWaterDensity:
    """Radial number density of water oxygens around each residue of ``sel``
    (RadialDensity.py:29-134)."""

    def __init__(self, data, sel, maxR=15, dr=0.1, frames=None):
        self.data = data
        self.sel = sel
        self._sel = self.data.selection(sel) if isinstance(sel, str) else sel
        self.maxR = maxR
        self.dr = dr
        self.radii = np.arange(self.dr, self.maxR, self.dr)             # RadialDensity.py:57
        if frames is None:
            self.frames = np.arange(0, self.data.nbrFrames, 1)
        else:
            self.frames = frames
        self.residues = self._sel.getUniqueResidues()                   # RadialDensity.py:66
        self.density = np.zeros((self.residues.size, self.radii.size))

    def compDensity(self):
        waters = np.ascontiguousarray(self.data.selection('name OH2').coordinates(self.frames), dtype='float32')
        cellDims = self.data.cellDims[self.frames]
        prot = np.ascontiguousarray(self._sel.coordinates(self.frames), dtype='float32')
        # histogram row of each protein atom = position of its residue in self.residues
        resOfAtom = np.asarray(self._sel.getResidues())
        order = {r: k for k, r in enumerate(self.residues.tolist())}
        groupId = np.array([order[r] for r in resOfAtom.tolist()], dtype=np.int32)
        dens = np.zeros((self.residues.size, self.radii.size), dtype='float32')
        py_getRadialNbrDensityGrouped(waters, prot, groupId, dens, cellDims, 0, self.maxR, self.dr)
        dens[:, 0] -= dens[:, 0]                                        # RadialDensity.py:100
        dens = dens / (4 * np.pi * self.radii ** 2 * self.dr)           # RadialDensity.py:101
        self.density[:] = dens.astype('float32')

    def plotDensity(self, medianFiltSize=[13, 3]):
        import matplotlib.pyplot as plt
        from matplotlib.cm import get_cmap
        from scipy.signal import medfilt2d
        fig, ax = plt.subplots(subplot_kw={'projection': '3d'})
        X, Y = np.meshgrid(self.radii, self.residues.astype(int))
        ax.plot_surface(X, Y, medfilt2d(self.density, medianFiltSize), cmap=get_cmap('jet'))
        ax.set_xlabel('Radius [$\\AA$]')
        ax.set_ylabel('Residue')
        ax.set_zlabel('Density')
        fig.show()


class RadialNumberDensity:
    """Radial number density of sel2 atoms around sel1 atoms (RadialDensity.py:137-239)."""

    def __init__(self, data, sel1, sel2=None, dr=0.1, maxR=15, frames=None):
        self.data = data
        self.sameSel = 0
        if frames is None:
            self.frames = np.arange(0, self.data.nbrFrames, 1)
        else:
            self.frames = frames
        if isinstance(sel1, str) and isinstance(sel2, str):
            if set(sel1.split(' ')) == set(sel2.split(' ')):            # RadialDensity.py:173-175
                self.sameSel = 1
        if isinstance(sel1, str):
            self.sel1 = self.data.selection(sel1)
        else:
            self.sel1 = sel1
        if isinstance(sel2, str):
            self.sel2 = self.sel1 if self.sameSel else self.data.selection(sel2)
        elif sel2 is not None:
            self.sel2 = sel2
        if sel2 is None:                                                # RadialDensity.py:193-195
            self.sel2 = self.sel1
            self.sameSel = 1
        self.dr = dr
        self.maxR = maxR

    def compDensity(self):
        radii = np.arange(self.dr, self.maxR, self.dr)                  # RadialDensity.py:206
        density = np.zeros(radii.size, dtype='float32')
        sel1 = np.ascontiguousarray(self.data[_indices(self.sel1), self.frames], dtype='float32')
        sel2 = sel1 if self.sameSel else \
            np.ascontiguousarray(self.data[_indices(self.sel2), self.frames], dtype='float32')
        cellDims = self.data.cellDims[self.frames]
        py_getRadialNbrDensity(sel1, sel2, density, cellDims, self.sameSel, self.maxR, self.dr)
        density[0] -= density[0]                                        # RadialDensity.py:222
        density /= (4 * np.pi * radii ** 2 * self.dr)                   # RadialDensity.py:223
        self.radii = radii
        self.density = density

    def plotDensity(self):
        import matplotlib.pyplot as plt
        plt.plot(self.radii, self.density)
        plt.xlabel('Radius [$\\AA$]')
        plt.ylabel('Density')
        plt.show()
