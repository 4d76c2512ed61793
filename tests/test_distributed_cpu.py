"""CPU, world_size 2, gloo: the frame sharding and the collectives of namdanalyzer_b200.distributed.
There is no GPU here, so the per-shard compute is a stand-in supplied through the ``compute``
hook (the oracle's C restatement -- test infrastructure acting as the checker's stand-in); what
is under test is the host logic: frame ranges, the all-reduce of histograms / float64 sums, the
gather of mask rows."""
import os
import sys
import numpy as np
import pytest
import torch.multiprocessing as mp

from conftest import ROOT
from namdanalyzer_b200 import distributed as nd


def test_frame_ranges_partition_the_frames():
    for nF in (1, 2, 7, 100, 1000, 10001):
        for world in (1, 2, 3, 4, 8):
            rs = [nd.frame_range(nF, r, world) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == nF
            assert all(rs[i][1] == rs[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in rs]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port_no, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from namdanalyzer_b200 import synth, distributed as nd
    from oracle import port
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        def rdf(sel1, sel2, cell, same, nb, dr, groupId=None, nGroups=1, frameBegin=0, frameEnd=None):
            return port.rdf_counts(sel1[:, frameBegin:frameEnd], sel2[:, frameBegin:frameEnd],
                                   cell[frameBegin:frameEnd], same, nb, dr, groupId=groupId, nGroups=nGroups)

        def dsum(sel1, sel2, cell, same, fb, fe):
            return port.distances_sum(sel1[:, fb:fe], sel2[:, fb:fe], cell[fb:fe], same)

        def wbits(xyz, r, o, cell, d, fb, fe):
            m = port.within(xyz[:, fb:fe], r, o, cell[fb:fe], d)[o]          # (nOut, nFc)
            pad = (-m.shape[0]) % 32
            m = np.concatenate([m, np.zeros((pad, m.shape[1]), m.dtype)]).T   # (nFc, nOutPad)
            return np.ascontiguousarray(np.packbits(np.ascontiguousarray(m.astype(np.uint8)), axis=1, bitorder="little")).view(np.uint32)

        c = synth.config1(n_frames=7)
        counts = nd.radial_nbr_counts(c["sel1"], c["sel1"], c["cellDims"], 1, 149, 0.1, compute=rdf)
        out = np.zeros(149, np.float32)
        nd.radial_nbr_density(c["sel1"], c["sel1"], out, c["cellDims"], 1, 15.0, 0.1, compute=rdf)
        c4 = synth.config4(n_frames=9)
        d = np.zeros((7, 1231), np.float32)
        nd.distances(c4["sel1"], c4["sel2"], d, c4["cellDims"], 0, compute=dsum)
        c2 = synth.config2(n_frames=5, n_water=400)
        fb, fe, bits = nd.within_bits(c2["allAtoms"], c2["refSel"], c2["outSel"], c2["cellDims"], 3.0,
                                      compute=wbits, gather=True)
        if rank == 0:
            q.put(dict(counts=counts, out=out, d=d, bits=bits, fbfe=(fb, fe)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_world_size_2_gloo_sharding_and_collectives():
    from namdanalyzer_b200 import synth
    from oracle import port
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port_no = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port_no, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    c = synth.config1(n_frames=7)
    want = port.rdf_counts(c["sel1"], c["sel1"], c["cellDims"], 1, 149, 0.1).astype(np.int64)
    assert np.array_equal(res["counts"], want)
    assert np.array_equal(res["out"], (want / 7.0).astype(np.float32))
    c4 = synth.config4(n_frames=9)
    w = port.distances_sum(c4["sel1"], c4["sel2"], c4["cellDims"], 0) / 9
    assert np.max(np.abs(res["d"] - w) / np.maximum(w, 1e-30)) <= 1e-6
    c2 = synth.config2(n_frames=5, n_water=400)
    m = port.within(c2["allAtoms"], c2["refSel"], c2["outSel"], c2["cellDims"], 3.0)[c2["outSel"]]
    got = np.unpackbits(res["bits"].view(np.uint8), axis=1, bitorder="little")[:, :m.shape[0]]
    assert res["fbfe"] == (0, 5) and np.array_equal(got.T.astype(np.int32), m)


def test_install_routes_the_reference_import_path():
    import namdanalyzer_b200
    mod = namdanalyzer_b200.install()
    from NAMDAnalyzer.lib.pylibFuncs import py_getWithin, py_getParallelBackend   # the reference's import line
    from namdanalyzer_b200.lib import pylibFuncs as ours
    assert py_getWithin is ours.py_getWithin and mod.py_getRadialNbrDensity is ours.py_getRadialNbrDensity
    assert py_getParallelBackend() == 2
