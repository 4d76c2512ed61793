"""Synthetic inputs of the five BASELINE.json configs.

Every generator returns plain numpy arrays in the reference's own layouts
(coordinates ``float32 (nAtoms, nFrames, 3)`` C-contiguous as produced by
NAMDAnalyzer/dataParsers/dcdReader.py:127,175; cells ``float32 (nFrames, 3)`` as in
NAMDAnalyzer/dataParsers/dcdCell.py:113-114).  Seeds 1001..1005 are fixed so the
tests, the golden-vector script and bench.py see the same numbers.

The "protein" of configs 2-4 is frame 0 of the reference's own test trajectory
(package/tests/test_data/ubq_ws.{psf,dcd}: ubiquitin, 1231 atoms, 76 residues),
committed as tests/golden/ubq_ws_fixture.npz by tests/golden/make_golden.py.
"""
import os
import numpy as np

F = np.float32
_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIXTURE = os.path.join(_ROOT, "tests", "golden", "ubq_ws_fixture.npz")
_fix = None


def fixture():
    """ubq_ws frame-0 protein coordinates [1231,3] f32 and residue index [1231] (0..75)."""
    global _fix
    if _fix is None:
        z = np.load(FIXTURE)
        _fix = {k: z[k] for k in z.files}
    return _fix


def _cells(rng, L, n_frames, rel=1e-3):
    """per-frame cubic cell L*(1 + rel*N(0,1)) so that the cell really varies per frame"""
    Lf = (L * (1.0 + rel * rng.standard_normal(n_frames))).astype(F)
    return np.ascontiguousarray(np.repeat(Lf[:, None], 3, axis=1))


def config1(n_frames=100, n_side=10, L=31.04, sigma=0.6, seed=1001):
    """RadialNumberDensity O-O on a TIP3P-like box: n_side^3 O on a jittered simple-cubic
    lattice, NOT wrapped back into the box (exercises |q| > 0.5), dr=0.1, maxR=15, sameSel."""
    rng = np.random.default_rng(seed)
    g = (np.arange(n_side) + 0.5) * (L / n_side)
    lat = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)
    xyz = lat[:, None, :] + sigma * rng.standard_normal((lat.shape[0], n_frames, 3))
    return dict(sel1=np.ascontiguousarray(xyz, F), cellDims=_cells(rng, L, n_frames),
                sameSel=1, dr=0.1, maxR=15.0)


def config2(n_frames=1000, n_water=7923, L=63.0, sigma=0.3, seed=1002):
    """selection 'name OH2 and within 3 of protein': 1231 protein atoms + n_water x 3 water
    atoms (O first) uniform in the box; N = 25000 for the default sizes."""
    rng = np.random.default_rng(seed)
    prot = fixture()["protein_xyz"].astype(np.float64)
    prot = prot - prot.mean(0) + L / 2
    nP = prot.shape[0]
    N = nP + 3 * n_water
    xyz = np.empty((N, n_frames, 3), F)
    xyz[:nP] = prot[:, None, :] + sigma * rng.standard_normal((nP, n_frames, 3))
    xyz[nP:] = rng.uniform(0.0, L, (3 * n_water, n_frames, 3))
    return dict(allAtoms=xyz, cellDims=_cells(rng, L, n_frames),
                refSel=np.arange(nP, dtype=np.int32),
                outSel=(nP + 3 * np.arange(n_water)).astype(np.int32), distance=3.0)


def config3(n_frames=5000, n_water=20000, L=85.0, sigma=0.3, seed=1003):
    """ResidueWiseWaterDensity('protein'): sel1 = water O, sel2 = protein atoms with a residue
    index per atom (76 groups), dr=0.1, maxR=15, sameSel=0."""
    rng = np.random.default_rng(seed)
    fx = fixture()
    prot = fx["protein_xyz"].astype(np.float64)
    prot = prot - prot.mean(0) + L / 2
    nP = prot.shape[0]
    p = (prot[:, None, :] + sigma * rng.standard_normal((nP, n_frames, 3))).astype(F)
    w = rng.uniform(0.0, L, (n_water, n_frames, 3)).astype(F)
    return dict(waters=np.ascontiguousarray(w), protein=np.ascontiguousarray(p),
                groupId=fx["protein_resindex"].astype(np.int32),
                nGroups=int(fx["protein_resindex"].max()) + 1,
                cellDims=_cells(rng, L, n_frames), dr=0.1, maxR=15.0)


def config4(n_frames=10000, L=63.0, sigma=0.5, seed=1004):
    """getDistances('protein and resid 53', 'protein'): 7 x 1231 atoms."""
    rng = np.random.default_rng(seed)
    fx = fixture()
    prot = fx["protein_xyz"].astype(np.float64)
    prot = prot - prot.mean(0) + L / 2
    nP = prot.shape[0]
    p = (prot[:, None, :] + sigma * rng.standard_normal((nP, n_frames, 3))).astype(F)
    idx = np.flatnonzero(fx["protein_resid"] == 53)
    return dict(sel1=np.ascontiguousarray(p[idx]), sel2=np.ascontiguousarray(p),
                sel1_idx=idx.astype(np.int32), cellDims=_cells(rng, L, n_frames), sameSel=0)


# ---- config 5: counter-based generator, reproduced bit-for-bit by the CUDA library ----------
# (csrc/nda_synth.cuh uses the same integer hash and the same individually rounded f32 ops)

def _mix32(x):
    """lowbias32 integer hash on uint32 arrays (wraps mod 2^32)."""
    x = x.astype(np.uint32)
    x ^= x >> np.uint32(16)
    x = (x * np.uint32(0x7FEB352D)).astype(np.uint32)
    x ^= x >> np.uint32(15)
    x = (x * np.uint32(0x846CA68B)).astype(np.uint32)
    x ^= x >> np.uint32(16)
    return x


def water_box_frame(n_side, frame, L=None, seed=1005, sigma=0.6, rho=0.0334):
    """One frame of an n_side^3 O lattice + triangular-ish jitter -> ([n,3] f32, cell [3] f32).
    jitter_k = ((h1>>10)+(h2>>10)+(h3>>10)+(h4>>10)) * 2^-22 - 2  (Irwin-Hall, var 1/3), times
    sigma*sqrt(3); every float op is a single IEEE f32 operation in a fixed order."""
    n = n_side ** 3
    if L is None:
        L = (n / rho) ** (1.0 / 3.0)
    L = F(L)
    a = np.arange(n, dtype=np.uint32)
    ix = (a // np.uint32(n_side * n_side)).astype(np.uint32)
    iy = ((a // np.uint32(n_side)) % np.uint32(n_side)).astype(np.uint32)
    iz = (a % np.uint32(n_side)).astype(np.uint32)
    h = F(L / F(n_side))
    amp = F(np.float64(F(sigma)) * np.sqrt(3.0))        # float32 sigma, as the C ABI receives it
    out = np.empty((n, 3), F)
    with np.errstate(over="ignore"):
        key = _mix32(np.uint32(seed) ^ _mix32(np.full(1, frame, np.uint32) + np.uint32(0x9E3779B9)))
        for k, idx in enumerate((ix, iy, iz)):
            s = np.zeros(n, np.uint32)
            for j in range(4):
                ctr = (a * np.uint32(12) + np.uint32(4 * k + j)).astype(np.uint32)
                s += _mix32(ctr ^ key) >> np.uint32(10)
            u = (s.astype(F) * F(2.0 ** -22)).astype(F) - F(2.0)
            base = ((idx.astype(F) + F(0.5)) * h).astype(F)
            out[:, k] = base + (u * amp).astype(F)
    cell = np.full(3, L, F)
    return out, cell


def config5(n_side=69, n_frames=4, seed=1005, sigma=0.6, rho=0.0334):
    """1M-atom water box RDF O-O (n_side=69 -> 328 509 O, the cube nearest 333 333) at
    reduced or full frame count; host-array form of what the CUDA library generates per
    frame on the device."""
    frames = [water_box_frame(n_side, f, None, seed, sigma, rho) for f in range(n_frames)]
    xyz = np.ascontiguousarray(np.stack([f[0] for f in frames], axis=1))
    cell = np.ascontiguousarray(np.stack([f[1] for f in frames], axis=0))
    return dict(sel1=xyz, cellDims=cell, sameSel=1, dr=0.1, maxR=15.0)


def n_bins(dr, maxR):
    """outSize exactly as the reference's callers size it
    (NAMDAnalyzer/dataAnalysis/RadialDensity.py:206-208): len(np.arange(dr, maxR, dr))."""
    return int(np.arange(dr, maxR, dr).size)


def hbond_case(n_water=400, n_frames=30, L=22.0, seed=1006, persist=0.9):
    """Water-like triples for the H-bond autocorrelation: oxygens on a jittered
    lattice that move slowly from frame to frame (so bonds persist for a while), one hydrogen per
    oxygen pointing at a neighbour.  acceptors = donors = oxygens, hydrogens[i] bound to donors[i]."""
    rng = np.random.default_rng(seed)
    side = int(np.ceil(n_water ** (1.0 / 3.0)))
    g = (np.arange(side) + 0.5) * (L / side)
    lat = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)[:n_water]
    o = np.empty((n_water, n_frames, 3))
    o[:, 0] = lat + 0.3 * rng.standard_normal((n_water, 3))
    for f in range(1, n_frames):
        o[:, f] = o[:, f - 1] + 0.12 * rng.standard_normal((n_water, 3))
    u = rng.standard_normal((n_water, n_frames, 3))
    for f in range(1, n_frames):                                  # slowly rotating O-H directions
        u[:, f] = persist * u[:, f - 1] + (1 - persist) * u[:, f]
    u /= np.linalg.norm(u, axis=2, keepdims=True)
    h = o + 0.96 * u
    cell = _cells(rng, L, n_frames)
    return dict(acceptors=np.ascontiguousarray(o, F), donors=np.ascontiguousarray(o, F),
                hydrogens=np.ascontiguousarray(h, F), cellDims=cell, maxR=2.8, minAngle=130.0)
