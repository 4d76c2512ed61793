"""Frame sharding across the GPUs of one box (one process per GPU, ``torch.distributed``).

Every reference loop has the frame as its outermost index (lib/openmp/src/getWithin.cpp:31,
getDistances.cpp:26, getRadialNbrDensity.cpp:12), so rank g of G takes the contiguous range
[floor(g*nF/G), floor((g+1)*nF/G)) and

 * RDF        : one all-reduce (SUM) of the integer histogram [nGroups][outSize] (1.2-90 KB; NCCL
                over NVLink when the group's backend is nccl), then counts / nFrames;
 * distances  : one all-reduce (SUM) of the float64 partial sums [n1][n2];
 * within     : NO collective -- ranks own disjoint frame columns of the mask.

The per-shard compute is the CUDA path (``lib.pylibFuncs``).  ``compute`` arguments exist so the
CPU (gloo, world_size 2) tests can exercise the sharding and the collective with a stand-in; the
product never passes one.
"""
import numpy as np

from .lib import pylibFuncs as plf


def frame_range(nFrames, rank, world):
    """Contiguous, balanced (sizes differ by at most 1) and exhaustive."""
    return (rank * nFrames) // world, ((rank + 1) * nFrames) // world


def _dist():
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized():
        return None
    return dist


def _all_reduce_sum(arr, group=None):
    """In-place SUM all-reduce of a numpy array through torch.distributed (on the GPU for nccl)."""
    dist = _dist()
    if dist is None or dist.get_world_size(group) == 1:
        return arr
    import torch
    t = torch.from_numpy(arr)
    if dist.get_backend(group) == "nccl":
        t = t.cuda()
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    arr[...] = t.cpu().numpy()
    return arr


def _rank_world(group=None):
    dist = _dist()
    if dist is None:
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


def radial_nbr_counts(sel1, sel2, cellDims, sameSel, outSize, dr, groupId=None, nGroups=1,
                      group=None, compute=None):
    """Exact pair counts over ALL frames, int64 (outSize,) or (nGroups, outSize), on every rank."""
    rank, world = _rank_world(group)
    nF = sel1.shape[1]
    fb, fe = frame_range(nF, rank, world)
    fn = compute or plf.py_getRadialNbrCounts
    local = np.asarray(fn(sel1, sel2, cellDims, sameSel, outSize, dr, groupId=groupId, nGroups=nGroups,
                          frameBegin=fb, frameEnd=fe)).astype(np.int64)
    return _all_reduce_sum(np.ascontiguousarray(local), group)


def radial_nbr_density(sel1, sel2, out, cellDims, sameSel, maxR, dr, group=None, compute=None):
    """``py_getRadialNbrDensity`` with the frames sharded over the ranks; ``out`` is filled on every
    rank with count / nFrames (float32)."""
    counts = radial_nbr_counts(sel1, sel2, cellDims, sameSel, out.size, dr, group=group, compute=compute)
    out[...] = (counts.astype(np.float64) / sel1.shape[1]).astype(np.float32).reshape(out.shape)


def distances(sel1, sel2, out, cellDims, sameSel, group=None, compute=None):
    """``py_getDistances`` with the frames sharded: float64 partial sums, one all-reduce, then the
    mean as float32 (the 1e-6 gate survives because the re-association happens in float64)."""
    rank, world = _rank_world(group)
    nF = sel1.shape[1]
    fb, fe = frame_range(nF, rank, world)
    fn = compute or plf.py_getDistancesSum
    s = np.ascontiguousarray(fn(sel1, sel2, cellDims, sameSel, fb, fe), dtype=np.float64)
    _all_reduce_sum(s, group)
    out[...] = (s / nF).astype(np.float32)


def within_bits(allAtoms, refSel, outSel, cellDims, distance, group=None, compute=None, gather=False):
    """This rank's rows of the bitmask: (frame_begin, frame_end, uint32 (fe-fb, words)).  With
    ``gather`` every rank receives the full (nFrames, words) mask (all_gather of equal-size
    padded blocks) -- a convenience for the caller, the computation itself needs no exchange."""
    rank, world = _rank_world(group)
    nF = allAtoms.shape[1]
    fb, fe = frame_range(nF, rank, world)
    fn = compute or plf.py_getWithinBits
    bits = np.ascontiguousarray(fn(allAtoms, refSel, outSel, cellDims, distance, fb, fe), dtype=np.uint32)
    if not gather or world == 1:
        return fb, fe, bits
    import torch
    dist = _dist()
    words = bits.shape[1]
    maxRows = (nF + world - 1) // world
    pad = np.zeros((maxRows, words), np.int32)
    pad[: fe - fb] = bits.view(np.int32)
    t = torch.from_numpy(pad)
    if dist.get_backend(group) == "nccl":
        t = t.cuda()
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t, group=group)
    full = np.zeros((nF, words), np.uint32)
    for r, p in enumerate(parts):
        b, e = frame_range(nF, r, world)
        full[b:e] = p.cpu().numpy().view(np.uint32)[: e - b]
    return 0, nF, full
