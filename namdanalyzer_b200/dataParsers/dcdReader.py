"""Minimal DCD accessor over the ingest drop-ins: what
NAMDAnalyzer/dataParsers/dcdReader.py (``dcdData[atoms, frames]``) and dcdCell.py
(``cellDims[frames]``) hand to the pair path, for ONE trajectory file.

    traj = DCDFile("run.dcd")
    xyz  = traj[atomIdx, frames]            # float32 (nAtoms, nFrames, 3) C-contiguous
    cell = traj.cellDims[frames]            # float32 (nFrames, 3); 100000.0 when the file has no cell
    T    = traj.resident(atomIdx, frames)   # namdanalyzer_b200.resident.ResidentTrajectory in HBM

Not a re-implementation of the reference's reader classes (multi-file append, PSF coupling, time axis
are out of scope); only the header fields and offsets the records need
(dataParsers/dcdReader.py:190-236).
"""
import os
from struct import unpack
import numpy as np

from ..lib.pylibFuncs import py_getDCDCoor, py_getDCDCell


class _CellAccessor:
    def __init__(self, traj):
        self.traj = traj

    def __getitem__(self, frames):
        t = self.traj
        f = np.atleast_1d(np.arange(t.nbrFrames)[frames] if not isinstance(frames, np.ndarray) else frames).astype(np.int32)
        if not t.cell:                                   # dcdCell.py:85-89
            return np.full((f.size, 3), 100000.0, np.float32)
        out = np.zeros((f.size, 6), np.float64)
        rc = py_getDCDCell(bytearray(t.path, "utf-8"), np.ascontiguousarray(f), t.startPos, out, ord(t.byteorder))
        if rc:
            raise IOError("DCD cell read failed with code %d" % rc)
        return np.ascontiguousarray(out[:, [0, 2, 5]], dtype=np.float32)   # dcdCell.py:113-114


class DCDFile:
    def __init__(self, path):
        self.path = os.path.abspath(path)
        self.byteorder = "<"
        with open(self.path, "rb") as f:
            data = f.read(92)
            rec = unpack("%si4c9if11i" % self.byteorder, data)
            if int(rec[0]) != 84:                        # dcdReader.py:199-201
                self.byteorder = ">"
                rec = unpack("%si4c9if11i" % self.byteorder, data)
            if int(rec[0]) != 84:
                raise IOError("%s is not a DCD file" % path)
            self.nbrFrames = int(rec[5])
            self.cell = bool(rec[15])
            title = unpack("%si" % self.byteorder, f.read(4))[0]
            f.read(title + 4)
            self.nbrAtoms = unpack("%siii" % self.byteorder, f.read(12))[1]
        rec_size = 12 * self.nbrAtoms + (80 if self.cell else 24)          # dcdReader.py:221-230
        self.startPos = np.arange(self.nbrFrames, dtype=np.int64) * rec_size + 112 + title
        # a truncated file holds fewer frames than its header claims
        size = os.path.getsize(self.path)
        self.nbrFrames = int(min(self.nbrFrames, max(0, (size - 112 - title) // rec_size)))
        self.startPos = self.startPos[:self.nbrFrames]
        self.cellDims = _CellAccessor(self)

    def __getitem__(self, key):
        atoms, frames = key if isinstance(key, tuple) else (key, slice(None))
        a = np.atleast_1d(np.arange(self.nbrAtoms)[atoms] if not isinstance(atoms, np.ndarray) else atoms).astype(np.int32)
        f = np.atleast_1d(np.arange(self.nbrFrames)[frames] if not isinstance(frames, np.ndarray) else frames).astype(np.int32)
        out = np.zeros((a.size, f.size, 3), np.float32)
        rc = py_getDCDCoor(bytearray(self.path, "utf-8"), np.ascontiguousarray(f), self.nbrAtoms,
                           np.ascontiguousarray(a), np.arange(3, dtype=np.int32), self.cell, self.startPos, out,
                           ord(self.byteorder))
        if rc == -1:
            raise IOError("Error while opening the file.")
        if rc == -2:
            raise IndexError("Out of range index.")
        if rc:
            raise IOError("Error while reading the file.")
        return out

    def resident(self, atoms=slice(None), frames=slice(None), device=0):
        """file -> HBM: the selected atoms / frames as a ResidentTrajectory"""
        from ..resident import ResidentTrajectory
        return ResidentTrajectory(self[atoms, frames], self.cellDims[frames], device=device)
