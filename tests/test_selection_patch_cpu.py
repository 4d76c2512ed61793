"""Row f-1 for real NAMDAnalyzer callers: ``namdanalyzer_b200.selection.patch_selparser`` against the
reference's OWN SelParser (loaded from its source file; skipped where /root/reference does not exist)."""
import importlib.util
import os

import numpy as np
import pytest

from namdanalyzer_b200.selection import decode_frames, patch_selparser

REF = "/root/reference/package/NAMDAnalyzer/selection/selParser.py"


class FakeData:
    """The slice of ``Dataset`` that SelParser touches: keyword selections and getWithin."""

    def __init__(self, n_atoms, n_frames, seed):
        rng = np.random.default_rng(seed)
        self.n_atoms, self.n_frames = n_atoms, n_frames
        self.protein = np.arange(0, n_atoms // 5)
        self.water = np.arange(n_atoms // 5, n_atoms)
        self.oh2 = self.water[::3]
        self.mask = (rng.random((n_atoms, n_frames)) < 0.2).astype(np.int32)
        self.calls = []

    def getSelection(self, selT=None, invert=False, **kw):
        if selT == "protein":
            sel = self.protein
        elif selT == "water":
            sel = self.water
        elif kw.get("atom"):
            sel = self.oh2
        else:
            sel = np.arange(self.n_atoms)
        return np.setdiff1d(np.arange(self.n_atoms), sel) if invert else sel

    def getWithin(self, distance, refSel, outSel="all", frame=0):
        self.calls.append((distance, frame))
        if isinstance(frame, slice):
            m = self.mask[:, frame]
        elif np.ndim(frame) == 0:
            m = self.mask[:, int(frame):int(frame) + 1]
        else:
            m = self.mask[:, np.asarray(frame)]
        if m.shape[1] == 1:
            return np.argwhere(m.flatten())[:, 0]
        return m


def _load_reference_class():
    spec = importlib.util.spec_from_file_location("ref_selParser", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.SelParser


def test_decode_frames_matches_argwhere_loop():
    rng = np.random.default_rng(4)
    mask = (rng.random((300, 17)) < 0.3).astype(np.int32)
    rows = np.unique(rng.integers(0, 300, 120))
    got = decode_frames(mask, rows)
    for f in range(17):
        want = np.intersect1d(rows, np.argwhere(mask[:, f])[:, 0])
        assert np.array_equal(got[f], want)
    got = decode_frames(mask)
    for f in range(17):
        assert np.array_equal(got[f], np.argwhere(mask[:, f])[:, 0])
    assert [a.size for a in decode_frames(np.zeros((5, 3), np.int32))] == [0, 0, 0]


@pytest.mark.skipif(not os.path.exists(REF), reason="reference sources not present on this machine")
@pytest.mark.parametrize("text", ["name OH2 and within 3 of protein frame 0:12",
                                  "within 3 of protein frame 2:14:3",
                                  "water and within 4.5 of protein frame 1 5 7",
                                  "name OH2 and within 3 of protein frame 4",
                                  "name OH2 and within 3 of protein",
                                  "protein"])
def test_patched_selparser_equals_reference(text):
    Ref = _load_reference_class()
    Fast = patch_selparser(type("FastSelParser", (_load_reference_class(),), {}))
    d1, d2 = FakeData(400, 16, 7), FakeData(400, 16, 7)
    a, b = Ref(d1, text), Fast(d2, text)
    def norm(calls):                                        # same getWithin request (distance, frames)
        out = []
        for dist, fr in calls:
            if isinstance(fr, slice):
                fr = ("slice", int(fr.start), int(fr.stop), None if fr.step is None else int(fr.step))
            elif isinstance(fr, np.ndarray):
                fr = tuple(int(v) for v in fr)
            out.append((float(dist), fr))
        return out
    assert norm(d1.calls) == norm(d2.calls)
    if isinstance(a.selection, list):
        assert isinstance(b.selection, list) and len(a.selection) == len(b.selection)
        for x, y in zip(a.selection, b.selection):
            assert np.array_equal(x, y)
    else:
        assert np.array_equal(a.selection, b.selection)
