"""A trajectory kept resident in HBM and queried many times.

The reference re-reads the DCD file for every selection (NAMDAnalyzer/dataParsers/dcdReader.py:55-175)
and its CUDA build re-uploads the whole selection on every call
(lib/cuda/src/getRadialNbrDensity.cu:64-76); ``ResidueWiseWaterDensity`` therefore moves the water
coordinates 76 times.  On a B200 a whole trajectory fits in HBM (25 000 atoms x 10 000 frames =
3 GB of 180 GB), so ``ResidentTrajectory`` uploads ``(nAtoms, nFrames, 3)`` once and every query --
given as atom INDEX arrays, like ``Dataset.getWithin`` / ``getDistances`` -- runs the CUDA path on
the resident copy through the ``nda_dev_*`` entry points.  Results are the same arrays the
reference-shaped operators return (tests/test_gpu_host_api.py compares them).

torch is used for device memory and the stream only.
"""
import ctypes
import numpy as np

from . import _capi


def _i32(a):
    return np.ascontiguousarray(np.asarray(a).ravel(), dtype=np.int32)


class ResidentTrajectory:
    def __init__(self, xyz, cellDims, device=0):
        import torch
        self._torch = torch
        xyz = np.ascontiguousarray(xyz, dtype=np.float32)
        if xyz.ndim != 3 or xyz.shape[2] != 3:
            raise ValueError("xyz must have shape (nAtoms, nFrames, 3)")
        self.nAtoms, self.nFrames = xyz.shape[:2]
        cell = np.ascontiguousarray(cellDims, dtype=np.float32).reshape(self.nFrames, 3)
        self.lib = _capi.load()
        _capi.check(self.lib.nda_set_device(int(device)))
        self.dev = torch.device("cuda", int(device))
        self.stream = torch.cuda.Stream(device=self.dev)
        with torch.cuda.stream(self.stream):
            self.xyz = torch.from_numpy(xyz).to(self.dev, non_blocking=False)
            self.cell = torch.from_numpy(cell).to(self.dev)
        self.stream.synchronize()
        self._st = ctypes.c_void_p(self.stream.cuda_stream)

    def _frames(self, frames):
        # the library's device is per calling thread: a query from another thread, or after a second
        # ResidentTrajectory was created on another device, must re-select this trajectory's device
        _capi.check(self.lib.nda_set_device(self.dev.index))
        if frames is None:
            return 0, self.nFrames
        fb, fe = frames
        if not (0 <= fb <= fe <= self.nFrames):
            raise ValueError("frame range outside the trajectory")
        return int(fb), int(fe)

    def _dev_idx(self, idx):
        t = self._torch.from_numpy(_i32(idx)).to(self.dev)
        if t.numel() and (int(t.min()) < 0 or int(t.max()) >= self.nAtoms):
            raise ValueError("atom index outside [0, %d)" % self.nAtoms)
        return t

    # ---- within ---------------------------------------------------------------------------------
    def within_bits(self, refSel, outSel, distance, frames=None):
        """uint32 (nFramesSel, ceil(len(outSel)/32)) bitmask (py_getWithinBits on the resident copy)"""
        torch = self._torch
        fb, fe = self._frames(frames)
        r, o = self._dev_idx(refSel), self._dev_idx(outSel)
        words = (o.numel() + 31) // 32
        with torch.cuda.stream(self.stream):
            bits = torch.zeros((fe - fb, words), dtype=torch.int32, device=self.dev)
            if o.numel() and fe > fb:
                _capi.check(self.lib.nda_dev_within(self.xyz.data_ptr(), self.nAtoms, self.nFrames,
                                                    r.data_ptr() if r.numel() else None, r.numel(),
                                                    o.data_ptr(), o.numel(), bits.data_ptr(), None,
                                                    self.cell.data_ptr(), float(distance), fb, fe, self._st))
            host = bits.cpu()
        self.stream.synchronize()
        return host.numpy().view(np.uint32)

    def within(self, refSel, outSel, distance, frames=None):
        """int32 (nAtoms, nFramesSel) 0/1 matrix, as NAMDDCD.getWithin builds it (dcdParser.py:178-184)"""
        outSel = _i32(outSel)
        bits = self.within_bits(refSel, outSel, distance, frames)
        flags = np.unpackbits(bits.view(np.uint8), axis=1, bitorder="little")[:, :outSel.size]
        out = np.zeros((self.nAtoms, bits.shape[0]), np.int32)
        f_idx, a_idx = np.nonzero(flags)
        out[outSel[a_idx], f_idx] = 1
        return out

    # ---- radial histogram -----------------------------------------------------------------------
    def radial_nbr_counts(self, sel1, sel2, outSize, dr, sameSel=0, groupId=None, nGroups=1, frames=None):
        """exact pair counts int64 (outSize,) or (nGroups, outSize); selections are atom indices"""
        torch = self._torch
        fb, fe = self._frames(frames)
        i1 = self._dev_idx(sel1)
        i2 = i1 if sameSel else self._dev_idx(sel2)
        G = 1 if groupId is None else int(nGroups)
        with torch.cuda.stream(self.stream):
            counts = torch.zeros((G, int(outSize)), dtype=torch.int64, device=self.dev)
            gid = None
            if groupId is not None:
                g = _i32(groupId)
                if g.size != i2.numel() or (g.size and (g.min() < 0 or g.max() >= G)):
                    raise ValueError("groupId must have one entry in [0, nGroups) per sel2 atom")
                gid = torch.from_numpy(g).to(self.dev)
            if i1.numel() and i2.numel() and fe > fb and outSize > 0:
                _capi.check(self.lib.nda_dev_radial_nbr_counts_sel(
                    self.xyz.data_ptr(), self.nAtoms, i1.data_ptr(), i1.numel(), i2.data_ptr(), i2.numel(),
                    gid.data_ptr() if gid is not None else None, G, counts.data_ptr(), int(outSize),
                    self.cell.data_ptr(), self.nFrames, fb, fe, int(bool(sameSel)), float(dr), self._st))
            host = counts.cpu()
        self.stream.synchronize()
        h = host.numpy()
        return h if groupId is not None else h[0]

    def radial_nbr_density(self, sel1, sel2, out, sameSel, maxR, dr, frames=None):
        """py_getRadialNbrDensity semantics on the resident copy: out float32 = count / nFramesSel"""
        fb, fe = self._frames(frames)
        c = self.radial_nbr_counts(sel1, sel2, out.size, dr, sameSel=sameSel, frames=(fb, fe))
        out[...] = (c.astype(np.float64) / max(fe - fb, 1)).astype(np.float32).reshape(out.shape)

    # ---- distances --------------------------------------------------------------------------------
    def distances(self, sel1, sel2=None, frames=None):
        """float32 (n1, n2) mean minimum-image distance over the frames (NAMDDCD.getDistances)"""
        torch = self._torch
        fb, fe = self._frames(frames)
        same = sel2 is None
        i1 = self._dev_idx(sel1)
        i2 = i1 if same else self._dev_idx(sel2)
        with torch.cuda.stream(self.stream):
            s = torch.zeros((i1.numel(), i2.numel()), dtype=torch.float64, device=self.dev)
            if i1.numel() and i2.numel() and fe > fb:
                _capi.check(self.lib.nda_dev_distances_sum_sel(
                    self.xyz.data_ptr(), self.nAtoms, i1.data_ptr(), i1.numel(), i2.data_ptr(), i2.numel(),
                    s.data_ptr(), self.cell.data_ptr(), self.nFrames, fb, fe, int(same), self._st))
            if same:
                s = s + s.T                       # the kernel accumulates j > i only
            host = s.cpu()
        self.stream.synchronize()
        return (host.numpy() / max(fe - fb, 1)).astype(np.float32)
