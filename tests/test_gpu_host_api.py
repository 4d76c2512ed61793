"""GPU: the host-side mirror of the reference interface (Dataset.selection('... within ...'),
Dataset.getWithin / getDistances, RadialNumberDensity, ResidueWiseWaterDensity) on the reference's
own test geometry (ubq_ws, no unit cell) and on periodic synthetic boxes, checked against the
oracle assembled the way the reference's Python glue assembles it."""
import numpy as np
import pytest

from namdanalyzer_b200 import synth, _capi
from namdanalyzer_b200.Dataset import ArrayDataset
from namdanalyzer_b200.dataAnalysis.RadialDensity import RadialNumberDensity, ResidueWiseWaterDensity
from oracle import port

pytestmark = pytest.mark.gpu


def ubq_dataset():
    fx = synth.fixture()
    n = fx["all_xyz"].shape[0]
    names = np.array(["X"] * n, dtype=object)
    names[fx["water_o_idx"]] = "OH2"
    resn = np.array(["TIP3"] * n, dtype=object)
    resn[fx["protein_idx"]] = "ALA"
    resid = np.zeros(n, int)
    resid[fx["protein_idx"]] = fx["protein_resid"]
    resid[~np.isin(np.arange(n), fx["protein_idx"])] = 1000
    return ArrayDataset(fx["all_xyz"], fx["all_cell"], names, resn, resid), fx


def test_selection_within_single_and_multi_frame():
    d, fx = ubq_dataset()
    assert d.selection("name OH2").size == 1817                      # the reference's own asserted count
    assert d.parallelBackend == 2
    want = port.within(fx["all_xyz"], fx["protein_idx"], fx["water_o_idx"], fx["all_cell"], 3.0)
    s0 = d.selection("name OH2 and within 3 of protein")             # default frame 0
    assert np.array_equal(s0.getIndices(), np.flatnonzero(want[:, 0]))
    s2 = d.selection("name OH2 and within 3 of protein frame 2")
    assert np.array_equal(s2.getIndices(), np.flatnonzero(want[:, 2]))
    sm = d.selection("name OH2 and within 3 of protein frame 0:3")   # list of per-frame index arrays
    assert len(sm.getIndices()) == 3
    for k in range(3):
        assert np.array_equal(sm.getIndices()[k], np.flatnonzero(want[:, k]))
    # Dataset.getWithin: index vector for one frame, 0/1 matrix for several (dcdParser.py:190-194)
    k1 = d.getWithin(3.0, "protein", "name OH2", frame=1)
    assert np.array_equal(k1, np.flatnonzero(want[:, 1]))
    km = d.getWithin(3.0, "protein", "name OH2", frame=np.arange(3))
    assert km.shape == (6682, 3) and np.array_equal(km, want)


def test_get_distances_host_method():
    d, fx = ubq_dataset()
    out = d.getDistances("protein and resid 53", "protein", frame=np.arange(3))
    pr = np.ascontiguousarray(fx["all_xyz"][fx["protein_idx"]])
    r53 = np.ascontiguousarray(pr[fx["protein_resid"] == 53])
    w = port.distances_sum(r53, pr, fx["all_cell"], 0) / 3
    assert out.shape == (7, 1231) and out.dtype == np.float32
    assert np.max(np.abs(out - w) / np.maximum(w, 1e-30)) <= 1e-6
    same = d.getDistances("protein and resid 53", frame=0)            # sel2 None -> sameSel
    assert same.shape == (7, 7) and np.array_equal(same, same.T) and np.all(np.diag(same) == 0)


def test_radial_number_density_class():
    c = synth.config1(n_frames=6)
    d = ArrayDataset(c["sel1"], c["cellDims"], names=["OH2"] * 1000, resnames=["TIP3"] * 1000)
    rdf = RadialNumberDensity(d, "name OH2", "name OH2", dr=0.1, maxR=15)
    assert rdf.sameSel == 1
    rdf.compDensity()
    counts = port.rdf_counts(c["sel1"], c["sel1"], c["cellDims"], 1, 149, 0.1).astype(np.float64)
    dens = (counts / 6).astype(np.float32)
    dens[0] -= dens[0]                                               # RadialDensity.py:222-223
    radii = np.arange(0.1, 15, 0.1)
    dens /= (4 * np.pi * radii ** 2 * 0.1)
    assert rdf.radii.size == 149 and np.array_equal(rdf.density, dens)


def test_residue_wise_water_density_class():
    c = synth.config3(n_frames=2, n_water=3000)
    nP, nW = c["protein"].shape[0], c["waters"].shape[0]
    fx = synth.fixture()
    xyz = np.concatenate([c["protein"], c["waters"]])
    d = ArrayDataset(xyz, c["cellDims"], names=["CA"] * nP + ["OH2"] * nW,
                     resnames=["ALA"] * nP + ["TIP3"] * nW,
                     resids=np.concatenate([fx["protein_resid"], np.full(nW, 5000)]))
    rw = ResidueWiseWaterDensity(d, "protein", maxR=15, dr=0.1)
    rw.compDensity()
    assert rw.density.shape == (76, 149)
    radii = np.arange(0.1, 15, 0.1)
    # the reference's loop (RadialDensity.py:84-103): one call per residue, string-sorted residues
    for k in (0, 17, 75):
        res = int(rw.residues[k])
        sel = np.ascontiguousarray(c["protein"][fx["protein_resid"] == res])
        cnt = port.rdf_counts(c["waters"], sel, c["cellDims"], 0, 149, 0.1).astype(np.float64)
        dens = (cnt / 2).astype(np.float32)
        dens[0] -= dens[0]
        dens = (dens / (4 * np.pi * radii ** 2 * 0.1)).astype(np.float32)
        assert np.array_equal(rw.density[k].astype(np.float32), dens)


def test_nccl_frame_sharding_if_two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import os, socket, sys
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port_no = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port_no, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    c = synth.config1(n_frames=9)
    want = port.rdf_counts(c["sel1"], c["sel1"], c["cellDims"], 1, 149, 0.1).astype(np.int64)
    assert np.array_equal(res, want)


def _nccl_worker(rank, world, port_no, q):
    import os, sys
    from conftest import ROOT
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from namdanalyzer_b200 import synth, distributed as nd, _capi
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    torch.cuda.set_device(rank)
    _capi.check(_capi.load().nda_set_device(rank))
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        c = synth.config1(n_frames=9)
        counts = nd.radial_nbr_counts(c["sel1"], c["sel1"], c["cellDims"], 1, 149, 0.1)
        if rank == 0:
            q.put(counts)
    finally:
        dist.destroy_process_group()


def test_within_lists_csr_matches_mask():
    from namdanalyzer_b200.lib import pylibFuncs as plf
    c = synth.config2(n_frames=7, n_water=1200)
    off, atoms = plf.py_getWithinLists(c["allAtoms"], c["refSel"], c["outSel"], c["cellDims"], 3.5)
    want = port.within(c["allAtoms"], c["refSel"], c["outSel"], c["cellDims"], 3.5)
    assert off.shape == (8,) and off[0] == 0 and off[-1] == atoms.size == want.sum()
    for f in range(7):
        assert np.array_equal(atoms[off[f]:off[f + 1]], np.flatnonzero(want[:, f]))


def test_resident_trajectory_matches_host_operators():
    from namdanalyzer_b200.resident import ResidentTrajectory
    from namdanalyzer_b200.lib import pylibFuncs as plf
    c = synth.config2(n_frames=9, n_water=1500)
    xyz, cell, r, o = c["allAtoms"], c["cellDims"], c["refSel"], c["outSel"]
    T = ResidentTrajectory(xyz, cell)
    want = port.within(xyz, r, o, cell, 3.5)
    assert np.array_equal(T.within(r, o, 3.5), want)
    assert np.array_equal(T.within(r, o, 3.5, frames=(2, 7)), want[:, 2:7])
    # RDF between two index selections, grouped and plain, and a frame range
    fx = synth.fixture()
    gid = fx["protein_resindex"].astype(np.int32)
    wO = o[:800]
    cnt = T.radial_nbr_counts(wO, r, 149, 0.1, groupId=gid, nGroups=76)
    w = port.rdf_counts(np.ascontiguousarray(xyz[wO]), np.ascontiguousarray(xyz[r]), cell, 0, 149, 0.1,
                        groupId=gid, nGroups=76).astype(np.int64)
    assert np.array_equal(cnt, w)
    oo = T.radial_nbr_counts(wO, None, 149, 0.1, sameSel=1, frames=(1, 4))
    a = np.ascontiguousarray(xyz[wO][:, 1:4])
    assert np.array_equal(oo, port.rdf_counts(a, a, cell[1:4], 1, 149, 0.1).astype(np.int64))
    dens = np.zeros(149, np.float32)
    T.radial_nbr_density(wO, wO, dens, 1, 15.0, 0.1)
    ref = np.zeros(149, np.float32)
    aa = np.ascontiguousarray(xyz[wO])
    plf.py_getRadialNbrDensity(aa, aa, ref, cell, 1, 15.0, 0.1)
    assert np.array_equal(dens, ref)
    # distances between index selections
    d = T.distances(r[:12], r)
    w = port.distances_sum(np.ascontiguousarray(xyz[r[:12]]), np.ascontiguousarray(xyz[r]), cell, 0) / 9
    assert np.max(np.abs(d - w) / np.maximum(w, 1e-30)) <= 1e-6
    ds = T.distances(r[:12])
    assert np.array_equal(ds, ds.T) and np.all(np.diag(ds) == 0)
    assert np.max(np.abs(ds - d[:, :12]) / np.maximum(d[:, :12], 1e-30)) <= 1e-6


def test_dcd_file_to_resident_trajectory(tmp_path):
    """row f-2 end to end: DCD file -> threaded reader -> HBM -> within, against the oracle"""
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from test_dcd_cpu import write_dcd
    from namdanalyzer_b200.dataParsers.dcdReader import DCDFile
    c = synth.config2(n_frames=4, n_water=500)
    cell6 = np.zeros((4, 6))
    cell6[:, [0, 2, 5]] = c["cellDims"]
    p = str(tmp_path / "traj.dcd")
    write_dcd(p, c["allAtoms"], cell6, "<")
    T = DCDFile(p).resident()
    want = port.within(c["allAtoms"], c["refSel"], c["outSel"], c["cellDims"], 3.0)
    assert np.array_equal(T.within(c["refSel"], c["outSel"], 3.0), want)


@pytest.mark.parametrize("pinned", [False, True])
def test_upload_pipeline_chunk_boundaries(pinned):
    """The host entry points cut the frames into chunks (tapered tail); force many
    small chunks on a small system so that every boundary, both staging buffers, the pinned (strided
    DMA) and the pageable (gathered) upload and the per-chunk mask scatter are exercised, and compare
    every frame with the oracle.  The second call with the same selections reuses the cached row plan,
    the third one must not."""
    import torch
    from namdanalyzer_b200.lib import pylibFuncs as plf
    c = synth.config2(n_frames=150, n_water=700)
    xyz, cell, r, o = c["allAtoms"], c["cellDims"], c["refSel"], c["outSel"]
    if pinned:
        xyz = torch.from_numpy(xyz).pin_memory().numpy()
        cell = torch.from_numpy(cell).pin_memory().numpy()
    want = port.within(c["allAtoms"], r, o, c["cellDims"], 3.0)
    for chunks in ("7", "3", "1"):
        _capi.set_option("NDA_WITHIN_CHUNKS", chunks)
        for rep in range(2):
            out = np.zeros(xyz.shape[:2], np.int32)
            plf.py_getWithin(xyz, r, o, out, cell, 3.0)
            assert np.array_equal(out, want), (chunks, rep)
    o2 = np.ascontiguousarray(o[::2])                          # another selection: the cached plan must go
    out = np.zeros(xyz.shape[:2], np.int32)
    plf.py_getWithin(xyz, r, o2, out, cell, 3.0)
    assert np.array_equal(out, port.within(c["allAtoms"], r, o2, c["cellDims"], 3.0))
    # RDF and distances share the uploader
    _capi.set_option("NDA_WITHIN_CHUNKS", None)
    _capi.set_option("NDA_UPLOAD_CHUNKS", "5")
    w = np.ascontiguousarray(xyz[o[:300]])
    pr = np.ascontiguousarray(xyz[r[:200]])
    got = plf.py_getRadialNbrCounts(w, pr, cell, 0, 40, 0.1)
    assert np.array_equal(got, port.rdf_counts(w, pr, c["cellDims"], 0, 40, 0.1))
    d = plf.py_getDistancesSum(w[:5], pr, cell, 0)
    dw = port.distances_sum(w[:5], pr, c["cellDims"], 0)
    _capi.set_option("NDA_UPLOAD_CHUNKS", None)
    assert np.max(np.abs(d - dw) / np.maximum(dw, 1e-30)) < 1e-12
