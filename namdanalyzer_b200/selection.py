"""SURVEY.md section 8 row f-1 for REAL NAMDAnalyzer users: the multi-frame ``within`` post-processing.

The reference turns the (nAtoms, nFrames) int32 mask of ``Dataset.getWithin`` into one index array per
frame with a Python loop of ``nFrames x (np.argwhere on a strided column + text parse + np.intersect1d)``
(NAMDAnalyzer/selection/selParser.py:117-126; consumer: dataAnalysis/ResidenceTime.py:68-78).  With the
mask itself coming off the GPU in milliseconds that loop becomes the whole cost of
``d.selection('name OH2 and within 3 of protein frame 0:1000')``.

``patch_selparser(SelParser)`` replaces ``SelParser.process`` with a version that

  * parses the ``frame`` keyword to the same ``slice`` / index array,
  * calls the class's own ``_getWithin`` (which reaches ``py_getWithin`` through ``install()``),
  * and, for a multi-frame mask whose remaining text has no ``same resid as`` / ``bound to`` clause
    (those depend on the frame's atoms and keep the reference's per-frame route), evaluates the remaining
    text ONCE and decodes all frames in one vectorised pass: rows of the mask restricted to that
    selection, ``np.nonzero`` of the transpose (frame-major order), ``np.split`` at the per-frame counts.

The resulting ``selection`` (list of sorted index arrays, one per frame) is element-wise identical to the
reference's; ``tests/test_selection_patch_cpu.py`` checks that against the reference's own class.
"""
import re

import numpy as np

_FRAME_RE = re.compile(r"frame\s+(.*)$")
_FRAME_STRIP_RE = re.compile(r" frame [0-9]+(:[0-9]+)*(\s[0-9]+)*")


def _parse_frames(text):
    """'a:b' / 'a:b:c' -> slice, 'i j k' -> int array (selParser.py:89-102 semantics)."""
    parts = text.strip()
    if re.fullmatch(r"[0-9]+(:[0-9]+){1,2}", parts):
        return slice(*[int(p) for p in parts.split(":")])
    return np.array(parts.split()).astype(int)


def decode_frames(mask, rows=None):
    """(nAtoms, nFrames) 0/1 mask -> list of nFrames sorted index arrays; ``rows`` (sorted, unique) restricts
    the atoms considered.  One pass over the mask instead of nFrames argwhere + intersect1d calls."""
    mask = np.asarray(mask)
    n_atoms, n_frames = mask.shape
    if rows is None:
        rows = np.arange(n_atoms)
    rows = np.asarray(rows)
    if rows.size == 0:
        return [rows[:0].copy() for _ in range(n_frames)]
    sub = mask[rows.astype(np.intp)]                        # (nRows, nFrames)
    f_idx, r_idx = np.nonzero(sub.T)                        # frame-major, rows ascending inside a frame
    counts = np.bincount(f_idx, minlength=n_frames)
    return np.split(rows[r_idx], np.cumsum(counts)[:-1])


def _fast_process(self):
    if not isinstance(self.selT, str):
        print("Selection text should be a string instance, the given argument cannot be parsed.")
        return
    m = _FRAME_RE.search(self.selT)
    if m is not None:
        self.frame = _parse_frames(m.group(1))
        self.selT = _FRAME_STRIP_RE.sub("", self.selT)
    else:
        self.frame = 0

    if "within" in self.selT:
        self.withinList = self._getWithin(self.selT)

    wl = self.withinList
    if wl is None:
        self.selection = self._parseText()
    elif wl.ndim <= 1:
        self.selection = self._parseText(wl)
    elif "same resid as" in self.selT or "bound to" in self.selT:
        # these clauses consume the frame's own atom list: one parse per frame, as the reference does
        text = self.selT
        self.selection = []
        for f in range(wl.shape[1]):
            self.selection.append(self._parseText(np.flatnonzero(wl[:, f])))
            self.selT = text
    else:
        rows = self._getSelection(self.selT) if self.selT != "" else None
        if rows is not None:
            rows = np.unique(np.asarray(rows))
        self.selection = decode_frames(wl, rows)


def patch_selparser(cls):
    """Replace ``cls.process`` (NAMDAnalyzer.selection.selParser.SelParser); idempotent; returns cls."""
    if getattr(cls, "_nda_b200_patched", False):
        return cls
    cls._nda_b200_reference_process = cls.process
    cls.process = _fast_process
    cls._nda_b200_patched = True
    return cls
