"""oracle.np_oracle -- TEST INFRASTRUCTURE.  numpy float32 restatement of the
per-pair arithmetic of the reference's OpenMP loops:

    d = a - b ; q = d / c ; n = roundf(q) ; w = d - c*n        (each op rounded to f32)
    r2 = (wx*wx + wy*wy) + wz*wz

    within : hit  <=> r2 <= distance*distance        lib/openmp/src/getWithin.cpp:29,63-77
    rdf    : idx = (int)(sqrtf(r2) / dr)              lib/openmp/src/getRadialNbrDensity.cpp:29-45
    dist   : sqrtf(r2)                                lib/openmp/src/getDistances.cpp:47-57

numpy evaluates every float32 ufunc with one IEEE rounding, so this is the strict
(no FMA, no fast-math) reading; tests/test_oracle.py checks it bit-for-bit against
the compiled reference.  O(n1*n2) memory per frame: small cases only.
"""
import numpy as np

F = np.float32


def roundf(q):
    """C roundf (half away from zero) on a float32 array; np.round is half-to-even."""
    t = np.trunc(q)
    f = q - t                      # exact
    return (t + np.where(np.abs(f) >= F(0.5), np.copysign(F(1), q), F(0))).astype(F)


def wrapped_delta(a, b, cell):
    """a [n1,3], b [n2,3], cell [3] -> w [n1,n2,3] float32 (a - b, wrapped)."""
    a = np.asarray(a, F)
    b = np.asarray(b, F)
    c = np.asarray(cell, F)
    with np.errstate(all="ignore"):
        d = a[:, None, :] - b[None, :, :]
        q = d / c
        n = roundf(q)
        p = c * n
        return (d - p).astype(F)


def r2_matrix(a, b, cell):
    w = wrapped_delta(a, b, cell)
    with np.errstate(all="ignore"):
        xx = w[..., 0] * w[..., 0]
        yy = w[..., 1] * w[..., 1]
        zz = w[..., 2] * w[..., 2]
        return ((xx + yy) + zz).astype(F)


def within(allAtoms, refSel, outSel, cellDims, distance, literal_early_out=True):
    """-> int32 [nAtoms][nFrames] (getWithin.cpp:9-90; atom - ref sign convention)."""
    allAtoms = np.asarray(allAtoms, F)
    nA, nF = allAtoms.shape[:2]
    out = np.zeros((nA, nF), np.int32)
    d2 = F(distance) * F(distance)
    refSel = np.asarray(refSel)
    outSel = np.asarray(outSel)
    for f in range(nF):
        w = wrapped_delta(allAtoms[outSel, f], allAtoms[refSel, f], cellDims[f])
        with np.errstate(all="ignore"):
            r2 = ((w[..., 0] * w[..., 0] + w[..., 1] * w[..., 1]) + w[..., 2] * w[..., 2]).astype(F)
            hit = r2 <= d2
            if literal_early_out:      # getWithin.cpp:64-71: signed component vs distance^2
                hit &= ~((w[..., 0] > d2) | (w[..., 1] > d2) | (w[..., 2] > d2))
        out[outSel[hit.any(axis=1)], f] = 1
    return out


def distances_per_frame(sel1, sel2, cellDims, sameSel):
    """-> float32 [nF][n1][n2] (getDistances.cpp; sel2 - sel1; j > i only when sameSel)."""
    sel1 = np.asarray(sel1, F)
    sel2 = np.asarray(sel2, F)
    n1, nF = sel1.shape[:2]
    n2 = sel2.shape[0]
    res = np.zeros((nF, n1, n2), F)
    for f in range(nF):
        # r2 is even in d, but keep the reference's operand order for bit-exact q
        r2 = r2_matrix(sel2[:, f], sel1[:, f], cellDims[f]).T
        with np.errstate(all="ignore"):
            d = np.sqrt(r2).astype(F)
        if sameSel:
            d = np.triu(d, 1)
        res[f] = d
    return res


def rdf_counts(sel1, sel2, cellDims, sameSel, outSize, dr, groupId=None, nGroups=1):
    """-> int64 [outSize] or [nGroups][outSize] (getRadialNbrDensity.cpp:8-52; sel1 - sel2)."""
    sel1 = np.asarray(sel1, F)
    sel2 = np.asarray(sel2, F)
    n1, nF = sel1.shape[:2]
    n2 = sel2.shape[0]
    dr = F(dr)
    G = 1 if groupId is None else nGroups
    counts = np.zeros((G, outSize), np.int64)
    gid = np.zeros(n2, np.int64) if groupId is None else np.asarray(groupId, np.int64)
    for f in range(nF):
        r2 = r2_matrix(sel1[:, f], sel2[:, f], cellDims[f])
        with np.errstate(all="ignore"):
            t = (np.sqrt(r2).astype(F) / dr).astype(F)
        ok = np.isfinite(t) & (t > F(-1)) & (t < F(2147483648.0))
        if sameSel:
            ok &= np.triu(np.ones((n1, n2), bool), 1)
        idx = np.where(ok, t, F(-1)).astype(np.int64)      # trunc toward zero, like (int)
        ok &= (idx >= 0) & (idx < outSize)
        g = np.broadcast_to(gid[None, :], (n1, n2))[ok]
        np.add.at(counts, (g, idx[ok]), 1)
    return counts[0] if groupId is None else counts
