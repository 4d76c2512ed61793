"""A small in-memory stand-in for NAMDAnalyzer's ``Dataset`` (package/NAMDAnalyzer/Dataset.py)
exposing exactly the surface the pair path touches: ``selection()``, ``dcdData[atoms, frames]`` /
``data[atoms, frames]``, ``cellDims[frames]``, ``nbrFrames``, ``getWithin``, ``getDistances``.

It is host plumbing for tests, examples and bench.py -- not a re-implementation of the
reference's PSF/DCD parsers or selection grammar (out of scope).  The
selection strings understood are the ones the BASELINE configs use:

    'all' | 'protein' | 'water' | 'name A B' | 'resid 5 9 12:20' | 'index 0 1 2' |
    terms joined by 'and' | '<sel> within <d> of <sel> [frame <i|a:b:c>]'

'within' follows NAMDAnalyzer/selection/selParser.py:279-354: the text before 'within' is the
out-selection, the text after ' of ' the reference selection.
"""
import numpy as np

from .dataParsers.dcdParser import PairOps

PROTEIN_RESNAMES = ["GLY", "ALA", "VAL", "LEU", "ILE", "MET", "PHE", "TRP", "PRO", "SER", "THR", "CYS",
                    "TYR", "ASN", "GLN", "ASP", "GLU", "LYS", "ARG", "HIS", "HSE", "HSP", "HSD"]
WATER_RESNAMES = ["TIP3", "TIP4", "HOH", "WAT", "SOL"]


class _Coords:
    """``dcdData``: indexing [atoms, frames] -> float32 (nAtoms, nFrames, 3) C-contiguous, like
    NAMDAnalyzer/dataParsers/dcdReader.py:55-175."""

    def __init__(self, xyz):
        self.xyz = xyz

    def __getitem__(self, key):
        atoms, frames = (key if isinstance(key, tuple) else (key, slice(None)))
        a = np.arange(self.xyz.shape[0])[atoms] if not isinstance(atoms, np.ndarray) else atoms
        f = np.arange(self.xyz.shape[1])[frames] if not isinstance(frames, np.ndarray) else frames
        a, f = np.atleast_1d(a), np.atleast_1d(f)
        return np.ascontiguousarray(self.xyz[np.ix_(a, f)], dtype=np.float32)


class _Cell:
    """``cellDims``: indexing [frames] -> float32 (nFrames, 3); F-ordered for several frames, as
    NAMDAnalyzer/dataParsers/dcdCell.py:113-114 returns it (SURVEY B-5)."""

    def __init__(self, cell):
        self.cell = cell

    def __getitem__(self, frames):
        f = np.atleast_1d(np.arange(self.cell.shape[0])[frames] if not isinstance(frames, np.ndarray) else frames)
        return np.asfortranarray(self.cell[f].astype(np.float32))


class Selection:
    """The slice of NAMDAnalyzer/selection/selText.py the pair path uses."""

    def __init__(self, dataset, indices, frames=0):
        self.dataset = dataset
        self._indices = indices
        self.frames = frames
        self.size = len(indices)
        self.shape = (self.size,)

    def getIndices(self):
        return self._indices

    def __len__(self):
        return self.size

    def __getitem__(self, i):
        return self._indices[i]

    def getResidues(self):
        return self.dataset.resid[self._indices].astype(str)

    def getUniqueResidues(self):
        return np.unique(self.getResidues())             # string-sorted, like selText.py:254-262

    def coordinates(self, frames=None):
        if frames is None:
            frames = slice(None)
        return self.dataset.dcdData[self._indices, frames]


class ArrayDataset(PairOps):
    def __init__(self, xyz, cellDims, names=None, resnames=None, resids=None):
        xyz = np.ascontiguousarray(xyz, dtype=np.float32)
        assert xyz.ndim == 3 and xyz.shape[2] == 3
        self.nbrAtoms, self.nbrFrames = xyz.shape[:2]
        self.dcdData = _Coords(xyz)
        self.cellDims = _Cell(np.asarray(cellDims, dtype=np.float32).reshape(self.nbrFrames, 3))
        n = self.nbrAtoms
        self.name = np.asarray(names if names is not None else ["X"] * n).astype(str)
        self.resname = np.asarray(resnames if resnames is not None else ["UNK"] * n).astype(str)
        self.resid = np.asarray(resids if resids is not None else np.zeros(n, int)).astype(int)
        self._init_pair_ops()

    def __getitem__(self, key):                          # Dataset IS the DCD reader (dcdReader.py:42)
        return self.dcdData[key]

    # -- selection mini-language (only what the BASELINE configs need) ----------------------------
    def _term(self, text):
        t = text.split()
        n = self.nbrAtoms
        if not t or t == ["all"]:
            return np.ones(n, bool)
        if t == ["protein"]:
            return np.isin(self.resname, PROTEIN_RESNAMES)
        if t == ["water"]:
            return np.isin(self.resname, WATER_RESNAMES)
        if t[0] == "name":
            return np.isin(self.name, t[1:])
        if t[0] == "resname":
            return np.isin(self.resname, t[1:])
        if t[0] in ("resid", "index"):
            vals = []
            for tok in t[1:]:
                if ":" in tok:
                    a, b = tok.split(":")[:2]
                    vals.extend(range(int(a), int(b)))
                else:
                    vals.append(int(tok))
            if t[0] == "resid":
                return np.isin(self.resid, vals)
            m = np.zeros(n, bool)
            m[np.asarray(vals, int)] = True
            return m
        raise ValueError("selection term not understood: %r" % text)

    def _static(self, text):
        m = np.ones(self.nbrAtoms, bool)
        for term in text.split(" and "):
            m &= self._term(term.strip())
        return m

    @staticmethod
    def _frames(tok, nF):
        if ":" in tok:
            p = [int(x) if x else None for x in tok.split(":")]
            return np.arange(nF)[slice(*p)]
        return int(tok)

    def selection(self, selT="all"):
        if not isinstance(selT, str):
            return Selection(self, np.asarray(selT).astype(int).ravel())
        text, frame = selT, 0
        if " frame " in text:                             # selParser.py:91-109
            text, ftok = text.split(" frame ")
            frame = self._frames(ftok.strip(), self.nbrFrames)
        if " within " not in text and not text.startswith("within "):
            return Selection(self, np.flatnonzero(self._static(text)), frame)
        # '<out terms> within <d> of <ref>'   (selParser.py:279-354)
        before, after = text.split("within", 1)
        dist_txt, ref_txt = after.split(" of ", 1)
        out_terms = [s.strip() for s in before.split(" and ") if s.strip()]
        out_mask = self._static(out_terms[-1]) if out_terms else np.ones(self.nbrAtoms, bool)
        rest_mask = self._static(" and ".join(out_terms[:-1])) if len(out_terms) > 1 else np.ones(self.nbrAtoms, bool)
        ref_idx = np.flatnonzero(self._static(ref_txt.strip()))
        out_idx = np.flatnonzero(out_mask & rest_mask).astype(np.int32)   # ascending -> lists come out sorted
        if np.ndim(frame) == 0:                           # one frame: index vector
            keep = self.getWithin(float(dist_txt), ref_idx, out_idx, frame)
            return Selection(self, keep, frame)
        # several frames: one index array per frame (selParser.py:117-126), decoded from the bitmask
        # in one pass instead of nFrames x (argwhere + intersect1d)
        from .lib.pylibFuncs import py_getWithinLists
        frames = np.asarray(frame)
        xyz = np.ascontiguousarray(self.dcdData[:, frames], dtype=np.float32)
        off, atoms = py_getWithinLists(xyz, ref_idx.astype(np.int32), out_idx, self.cellDims[frames], float(dist_txt))
        per_frame = [atoms[off[k]:off[k + 1]] for k in range(frames.size)]
        return Selection(self, per_frame, frame)
