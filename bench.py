#!/usr/bin/env python
"""bench.py -- the driver's measurement contract for the minimum-image pair path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Headline workload (config.workload): BASELINE.json configs[1] -- selection 'name OH2 and within 3 of
protein' on a ubiquitin-in-water box, 25 000 atoms, 1231 reference atoms x 7923 water O, 1000 frames
(namdanalyzer_b200.synth.config2, seed 1002).  One STEP = one pass of the within path over that
1000-frame batch = 9.75e9 minimum-image pair evaluations.  With N GPUs every rank owns its own batch
(frames shard with no data-path collective): scaling "weak".

  value      pair evaluations / s over EXACTLY --steps steps, inputs resident in HBM (reference layout),
             CUDA events on the launching stream: pack + prep + within kernel + mask expansion.
  sustained  the same loop run for >= 2 s, with clocks / power sampled while it runs.
  e2e        the same metric through the reference-shaped operator py_getWithin with plain (PAGEABLE) numpy
             arrays, what a drop-in caller passes: host->device copies of coordinates / selections / cells and
             the device->host read of the bitmask are inside the timed region.  e2e.pinned: the same call on
             page-locked arrays.
  roofline   the dominant kernel (within_tc_kernel) against the tensor roofline; fp32_filter_arm and
             rdf.fp32_filter_arm: the packed-FP32 kernels against the FP32 peak measured live.
  rdf        co-headline (BASELINE's metric is "RDF + within"): RDF O-O on a config-5 shaped water box --
             resident value, e2e from host arrays through py_getRadialNbrDensity, its own cpu_baseline.
  configs    all five BASELINE shapes: end to end from plain numpy arrays and pair-kernel time.
  cpu_baseline  the reference's own OpenMP getWithin (oracle/_ref/libref_fast.so, shipped flags), one frame
             per call, all host threads, on a bounded sample of the same frames.
  N > 1      strong-scaling legs (ONE fixed trajectory split over the ranks), each with a parity record against
             rank 0 computing alone; the drop-in call sharding over all GPUs inside ONE process
             (nda_set_devices); the concurrent pinned-H2D ceiling of the box.

`--impl reference` times only the reference CPU implementation on bounded samples of the workload.
"""
import argparse
import ctypes
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_WITHIN = 21.0      # wrap + r2 = 20, + compare
FLOP_RDF = 22.0         # + sqrtf, /dr
FP32_NOMINAL_TFLOPS = 2 * 128 * 148 * 1.965e9 / 1e12     # 74.5
N_SIDE5 = 69            # config-5 shape: 69^3 = 328 509 oxygens at rho = 0.0334 A^-3 (BASELINE: ~333 333)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=1000, help="frames per step (BASELINE: 1000)")
    ap.add_argument("--no-extras", action="store_true", help="headline + e2e only (A/B runs)")
    ap.add_argument("--sustain-s", type=float, default=2.2, help="length of the sustained loop in seconds")
    ap.add_argument("--rdf-full-frames", type=int, default=10000,
                    help="frames of the full-length config-5 RDF pass (BASELINE: 10000; ~33 s on one B200; 0 skips it)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append((time.perf_counter(), ln.strip()))

    def stop(self, t_begin=None, t_end=None):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, ln in self.lines:
            if t_begin is not None and not (t_begin <= ts <= t_end + 0.06):
                continue
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0])); mx.append(float(p[1])); pw.append(float(p[2]))
            except ValueError:
                continue
            for n, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "power_w_median": float(np.median(pw)) if pw else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------- reference arm
def cpu_within_sample(cfg, frames, threads, repeats=1):
    """Reference OpenMP getWithin (shipped flags) one frame per call on `frames`; -> (pairs/s, s)."""
    from oracle import ref
    os.environ["OMP_NUM_THREADS"] = str(threads)
    R = ref.RefLib("fast")
    xyz, cell, rs, os_ = cfg["allAtoms"], cfg["cellDims"], cfg["refSel"], cfg["outSel"]
    prepared = []
    for f in frames:
        prepared.append((np.ascontiguousarray(xyz[:, f:f + 1, :]), np.ascontiguousarray(cell[f:f + 1]),
                         np.zeros((xyz.shape[0], 1), np.int32)))
    x0, c0, o0 = prepared[0]
    R.raw_within_frame(x0, rs, os_, o0.copy(), c0, cfg["distance"])      # thread-pool start-up, untimed
    t0 = time.perf_counter()
    for _ in range(repeats):
        for x, c, o in prepared:
            o[:] = 0
            R.raw_within_frame(x, rs, os_, o, c, cfg["distance"])
    dt = time.perf_counter() - t0
    pairs = float(rs.size) * float(os_.size) * len(frames) * repeats
    return pairs / dt, dt


def cpu_rdf_sample(threads, budget_s=8.0):
    """Reference OpenMP getRadialNbrDensity (shipped flags), one frame per call, on a 32^3-atom box of the
    config-5 density (the full 69^3 frame takes the CPU minutes); -> dict."""
    from namdanalyzer_b200 import synth
    from oracle import ref
    os.environ["OMP_NUM_THREADS"] = str(threads)
    R = ref.RefLib("fast")
    c = synth.config5(n_side=32, n_frames=2)
    n = c["sel1"].shape[0]
    f32p = ctypes.POINTER(ctypes.c_float)
    frames = [(np.ascontiguousarray(c["sel1"][:, f:f + 1]), np.ascontiguousarray(c["cellDims"][f:f + 1])) for f in range(2)]
    d = np.zeros(149, np.float32)

    def one(x, cell):
        with R._quiet():
            R.fn["getRadialNbrDensity"](x.ctypes.data_as(f32p), n, x.ctypes.data_as(f32p), n, d.ctypes.data_as(f32p), 149,
                                        cell.ctypes.data_as(f32p), 1, 1, ctypes.c_float(15.0), ctypes.c_float(0.1))
    one(*frames[0])
    calls, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < budget_s and calls < 200:
        one(*frames[calls % 2])
        calls += 1
    dt = time.perf_counter() - t0
    pairs = 0.5 * n * (n - 1.0) * calls
    return {"value": pairs / dt, "unit": "pairs/s", "cores": threads, "kind": "reference",
            "sample": "reference OpenMP getRadialNbrDensity (oracle/_ref/libref_fast.so, shipped flags), one frame per call, "
                      "%d calls on a 32^3 = %d atom box of the config-5 density (%.1f s)" % (calls, n, dt)}


def legacy_gpu_sample():
    """The reference's OWN CUDA build (lib/cuda/src/*.cu, unmodified, compiled for sm_100a into
    oracle/_ref/libref_cuda.so) on bounded samples: the recompiled legacy kernels this repo is meant
    to beat.  End to end like the reference runs them (cudaMalloc + blocking copies + one launch and
    device sync per frame + cudaFree).  Not a parity oracle: the two reference builds disagree."""
    from namdanalyzer_b200 import synth
    from oracle import ref
    if not ref.available("cuda"):
        return {"unavailable": "oracle/_ref/libref_cuda.so not built"}
    R = ref.RefLib("cuda")
    out = {}
    nF = 100
    c = synth.config2(n_frames=nF)
    o = np.zeros(c["allAtoms"].shape[:2], np.int32)
    f32p, i32p = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int)

    def within():
        with R._quiet():
            R.fn["getWithin"](c["allAtoms"].ctypes.data_as(f32p), c["allAtoms"].shape[0], nF,
                              c["refSel"].ctypes.data_as(i32p), c["refSel"].size,
                              c["outSel"].ctypes.data_as(i32p), c["outSel"].size,
                              o.ctypes.data_as(i32p), c["cellDims"].ctypes.data_as(f32p), ctypes.c_float(3.0))
    within()
    t0 = time.perf_counter()
    within()
    dt = time.perf_counter() - t0
    out["within_config2"] = {"frames": nF, "pairs_per_s": 1231.0 * 7923 * nF / dt, "ms": dt * 1e3}
    c5 = synth.config5(n_side=32, n_frames=2)                 # 32 768 atoms: the legacy kernel is O(n^2) global atomics
    d = np.zeros(149, np.float32)
    n = c5["sel1"].shape[0]

    def rdf():
        with R._quiet():
            R.fn["getRadialNbrDensity"](c5["sel1"].ctypes.data_as(f32p), n, c5["sel1"].ctypes.data_as(f32p), n,
                                        d.ctypes.data_as(f32p), 149, c5["cellDims"].ctypes.data_as(f32p), 2, 1,
                                        ctypes.c_float(15.0), ctypes.c_float(0.1))
    rdf()
    t0 = time.perf_counter()
    rdf()
    dt = time.perf_counter() - t0
    # the legacy kernel ignores sameSel and evaluates the full n x n matrix (SURVEY B-6)
    out["rdf_32768_atoms"] = {"frames": 2, "pairs_evaluated_per_s": float(n) * n * 2 / dt,
                              "unique_pairs_per_s": 0.5 * n * (n - 1.0) * 2 / dt, "ms": dt * 1e3}
    out["note"] = ("reference CUDA build recompiled for sm_100a, end to end from host arrays "
                   "(its only mode); semantics differ from the OpenMP oracle (SURVEY Appendix B)")
    return out


def run_reference(args, rank, world):
    if rank != 0:
        return
    from namdanalyzer_b200 import synth
    threads = os.cpu_count() or 1
    sample_frames = 24
    cfg = synth.config2(n_frames=sample_frames)
    for _ in range(args.warmup):
        cpu_within_sample(cfg, range(0, 4), threads)
    times = []
    for s in range(args.steps):
        r, dt = cpu_within_sample(cfg, range(sample_frames), threads)
        times.append(dt)
        if sum(times) > 150.0:                              # bounded: the whole run ends within a few minutes
            break
    steps_done = len(times)
    pairs_step = float(cfg["refSel"].size) * cfg["outSel"].size * sample_frames
    total = sum(times)
    value = pairs_step * steps_done / total
    line = {
        "impl": "reference", "metric": "min-image pair evals/sec (within, config 2)", "value": value,
        "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / steps_done, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": threads, "kind": "reference",
                         "sample": "reference OpenMP getWithin (oracle/_ref/libref_fast.so, -O2 -fopenmp -ffast-math), "
                                   "one frame per call, each step = %d of the workload's 1000 frames x 1231 x 7923 pairs; "
                                   "%d steps run" % (sample_frames, steps_done)},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def workload_config(args):
    return {"workload": "BASELINE configs[1]: 'name OH2 and within 3 of protein', 25000 atoms "
                        "(1231 ref x 7923 out), %d frames per GPU per step, synthetic seed 1002" % args.frames,
            "pairs_per_step_per_gpu": 1231.0 * 7923.0 * args.frames,
            "l2_policy": "inputs (300 MB raw coordinates + ~450 MB packed coordinates and fp16 embeddings per step) exceed the 126 MB L2; no explicit flush",
            "frame_sharding": "each rank owns its own frame batch; no data-path collective"}


def bind_to_gpu_numa_node(local_rank):
    """One process per GPU: run (and so allocate pinned host buffers) on the CPU socket the GPU hangs off.
    Best effort: returns a description or None."""
    try:
        out = subprocess.run(["nvidia-smi", "-i", str(local_rank), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=20).stdout.strip().lower()
        if not out:
            return None
        dom, rest = out.split(":", 1)
        bus = "%s:%s" % (dom[-4:], rest)                       # 00000000:1b:00.0 -> 0000:1b:00.0
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return {"gpu_bus": bus, "numa_node": node, "cpus": len(cpus)}
    except Exception:
        return None


# ------------------------------------------------------------------------------------- helpers of our arm
class Env:
    """what every leg needs: torch, the library, this rank's device and stream, barriers"""
    pass


def gen_frames(E, xyz, cell, stride, gfb, gfe, Lbox):
    """frames [gfb, gfe) of the config-5 generator into a LOCAL array whose column 0 is frame gfb (row stride `stride` frames)"""
    if gfe > gfb:
        E.capi.check(E.lib.nda_dev_synth_water_box_slice(xyz.data_ptr(), cell.data_ptr(), N_SIDE5, stride, gfb, gfe - gfb,
                                                         Lbox, 0.6, 1005, E.st))


def kernel_ms(E):
    ms = ctypes.c_double(0.0)
    nl = ctypes.c_int(0)
    E.capi.check(E.lib.nda_last_kernel_ms(ctypes.byref(ms), ctypes.byref(nl)))
    return ms.value, nl.value


def max_over_ranks(E, x):
    t = E.torch.tensor([float(x)], dtype=E.torch.float64, device=E.dev)
    if E.world > 1:
        E.dist.all_reduce(t, op=E.dist.ReduceOp.MAX)
    return float(t.item())


def timed_calls(fn, n, sync):
    """wall clock of n calls of a host entry point (each returns with its result on the host)"""
    sync()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    sync()
    return (time.perf_counter() - t0) / n


def tensor_peak():
    tpeak, tsus, tsrc = 1590.0, None, "fallback of B200_PROFILING.md"
    mp = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(mp):
        try:
            j = json.load(open(mp))
            tpeak, tsus = float(j["bf16_tflops"]), j.get("bf16_tflops_sustained")
            tsrc = "MEASURED_PEAKS.json bf16_tflops (burst figure; the kernel is timed alone)"
        except Exception:
            pass
    return tpeak, tsus, tsrc


def hbm_peak():
    mp = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(mp))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    except Exception:
        return 6550.0, "fallback of B200_PROFILING.md"


def committed_traffic(kname):
    tp = os.path.join(ROOT, "profiles", "traffic_latest.json")
    if os.path.exists(tp):
        try:
            j = json.load(open(tp))
            return j.get(kname + "_dram_bytes_per_launch"), j.get("source", "profiles/traffic_latest.json")
        except Exception:
            pass
    return None, None


# ------------------------------------------------------------------------------------- RDF co-headline
def rdf_leg(E, peak_fp32, n_frames=64, e2e_frames=12):
    """RDF O-O, config-5 shape: 69^3 oxygens, frames generated on the device (bit-identical to synth.config5)."""
    torch, lib, capi, plf = E.torch, E.lib, E.capi, E.plf
    n = N_SIDE5 ** 3
    L = float(np.float32((n / 0.0334) ** (1.0 / 3.0)))
    xyz = torch.empty((n, n_frames, 3), dtype=torch.float32, device=E.dev)
    cell = torch.empty((n_frames, 3), dtype=torch.float32, device=E.dev)
    capi.check(lib.nda_dev_synth_water_box(xyz.data_ptr(), cell.data_ptr(), N_SIDE5, n_frames, 0, n_frames, L, 0.6, 1005, E.st))
    counts = torch.zeros(149, dtype=torch.int64, device=E.dev)
    per_frame = 0.5 * n * (n - 1.0)

    def go(fb, fe):
        counts.zero_()
        capi.check(lib.nda_dev_radial_nbr_counts(xyz.data_ptr(), n, xyz.data_ptr(), n, None, 1, counts.data_ptr(), 149,
                                                 cell.data_ptr(), n_frames, fb, fe, 1, 0.1, E.st))
    go(0, 4)
    torch.cuda.synchronize()
    reps = 3
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(E.stream)
    ks = []
    for _ in range(reps):
        go(0, n_frames)
    e1.record(E.stream)
    torch.cuda.synchronize()
    step_ms = e0.elapsed_time(e1) / reps
    go(0, n_frames)
    k_ms, k_n = kernel_ms(E)
    total_resident = int(counts.sum().item())
    filt = int(lib.nda_last_filter())
    pairs = per_frame * n_frames
    tpeak, tsus, tsrc = tensor_peak()
    out = {"metric": "min-image pair evals/sec (RDF O-O, config-5 shape)",
           "workload": "BASELINE configs[4] shape: %d O atoms (69^3 at rho = 0.0334 A^-3; BASELINE names ~333 333), %d frames "
                       "generated on the device, 149 bins of 0.1 A, sameSel; one step = one pass over all frames" % (n, n_frames),
           "value": pairs / (step_ms * 1e-3), "unit": "pairs/s", "ms_per_step": step_ms, "pairs_per_step": pairs,
           "steps": reps, "filter": "tensor-core" if filt == 100 else "fp32 %d" % filt,
           "roofline": {"bound": "tensor", "kernel": "rdf_tc_kernel<false>" if filt == 100 else "rdf_kernel",
                        "kernel_ms": k_ms, "launches": k_n, "kernel_pairs_per_s": pairs / (k_ms * 1e-3),
                        "achieved": pairs * 32.0 / (k_ms * 1e-3) / 1e12, "peak": tpeak, "unit": "TFLOP/s",
                        "frac": pairs * 32.0 / (k_ms * 1e-3) / 1e12 / tpeak, "flop_per_pair": 32.0, "peak_source": tsrc,
                        "traffic": committed_traffic("rdf_tc_kernel")[0],
                        "traffic_source": committed_traffic("rdf_tc_kernel")[1]},
           "counts_total": total_resident}
    # ---- the packed-FP32 kernel on the same box: the north-star ">= 50 % of FP32 peak for the RDF kernel" figure
    with capi.options(NDA_FILTER="fold2a1"):
        nf32 = 4
        go(0, nf32)
        ks = []
        for _ in range(3):
            go(0, nf32)
            ks.append(kernel_ms(E)[0])
        fk = float(np.mean(ks))
        f32_total = int(counts.sum().item())
    go(0, nf32)
    tc4_total = int(counts.sum().item())
    ach = per_frame * nf32 * FLOP_RDF / (fk * 1e-3) / 1e12
    out["fp32_filter_arm"] = {
        "filter": "fold2a1 (packed FP32: FFMA2 / FADD2)", "kernel": "rdf_kernel<false, fold2a1>", "frames": nf32, "kernel_ms": fk,
        "kernel_pairs_per_s": per_frame * nf32 / (fk * 1e-3),
        "roofline": {"bound": "fp32", "achieved": ach, "peak": peak_fp32, "unit": "TFLOP/s", "frac": ach / peak_fp32 if peak_fp32 else None,
                     "frac_of_nominal": ach / FP32_NOMINAL_TFLOPS, "flop_per_pair": FLOP_RDF,
                     "peak_source": "FFMA loop measured live in this run by nda_measure_fp32_peak (not MEASURED_PEAKS.json, which has no FP32 figure)"},
        "counts_equal_tensor_filter": f32_total == tc4_total}
    # ---- end to end: host arrays through the reference-shaped operator
    h_xyz = xyz[:, :e2e_frames].contiguous().cpu().numpy()      # plain (pageable) numpy, as a caller holds it
    h_cell = cell[:e2e_frames].cpu().numpy()
    h_out = np.zeros(149, np.float32)

    def call():
        h_out[:] = 0
        plf.py_getRadialNbrDensity(h_xyz, h_xyz, h_out, h_cell, 1, 15.0, 0.1)
    call()
    dt = timed_calls(call, 2, torch.cuda.synchronize)
    go(0, e2e_frames)
    want = (counts.cpu().numpy().astype(np.float64) / e2e_frames).astype(np.float32)
    out["e2e"] = {"value": per_frame * e2e_frames / dt, "unit": "pairs/s", "ms_per_call": dt * 1e3, "frames": e2e_frames,
                  "h2d_bytes_per_step": int(h_xyz.nbytes + h_cell.nbytes), "d2h_bytes_per_step": 149 * 8,
                  "api": "namdanalyzer_b200.lib.pylibFuncs.py_getRadialNbrDensity (pageable numpy arrays)",
                  "equals_resident_result": bool(np.array_equal(h_out, want))}
    hist64 = None
    go(0, n_frames)
    hist64 = counts.cpu().numpy().copy()
    del xyz, h_xyz
    # ---- BASELINE configs[4] at its stated length: 10 000 frames, the whole trajectory resident in HBM
    nFull = E.args.rdf_full_frames
    if nFull > 0:
        torch.cuda.empty_cache()
        xyz = torch.empty((n, nFull, 3), dtype=torch.float32, device=E.dev)
        cell = torch.empty((nFull, 3), dtype=torch.float32, device=E.dev)
        capi.check(lib.nda_dev_synth_water_box(xyz.data_ptr(), cell.data_ptr(), N_SIDE5, nFull, 0, nFull, L, 0.6, 1005, E.st))

        def gof(fb, fe):
            counts.zero_()
            capi.check(lib.nda_dev_radial_nbr_counts(xyz.data_ptr(), n, xyz.data_ptr(), n, None, 1, counts.data_ptr(), 149,
                                                     cell.data_ptr(), nFull, fb, fe, 1, 0.1, E.st))
        gof(0, min(n_frames, nFull))
        same = bool(np.array_equal(counts.cpu().numpy(), hist64)) if nFull >= n_frames else None
        torch.cuda.synchronize()
        e0.record(E.stream)
        gof(0, nFull)
        e1.record(E.stream)
        torch.cuda.synchronize()
        sec = e0.elapsed_time(e1) * 1e-3
        kf_ms, kf_n = kernel_ms(E)
        out["full_size"] = {"workload": "BASELINE configs[4] at its stated length: %d O atoms x %d frames, %.1f GB of coordinates generated and "
                                        "resident in HBM, ONE call" % (n, nFull, n * nFull * 12 / 1e9),
                            "pairs": per_frame * nFull, "seconds": sec, "value": per_frame * nFull / sec, "unit": "pairs/s",
                            "pair_kernel_seconds": kf_ms * 1e-3, "pair_kernel_launches": kf_n,
                            "counts_total": int(counts.sum().item()),
                            "first_%d_frames_equal_the_%d_frame_run" % (n_frames, n_frames): same}
        del xyz, cell
        torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------------- all five BASELINE shapes
def configs_leg(E, headline):
    """End to end from plain numpy arrays through the reference-shaped operators + pair-kernel time."""
    torch, lib, capi, plf, synth = E.torch, E.lib, E.capi, E.plf, E.synth
    sync = torch.cuda.synchronize
    rows = []

    def exact_frames():
        st4 = (ctypes.c_double * 4)()
        capi.check(lib.nda_last_stats(st4))
        return int(st4[2])

    def row(i, workload, pairs, fn, calls, h2d, extra=None):
        fn()
        dt = timed_calls(fn, calls, sync)
        fn()
        k_ms, k_n = kernel_ms(E)
        filt = int(lib.nda_last_filter())
        r = {"config": i, "workload": workload, "pairs": pairs, "e2e_ms": dt * 1e3, "e2e_pairs_per_s": pairs / dt,
             "kernel_ms": k_ms, "kernel_launches": k_n, "kernel_pairs_per_s": pairs / (k_ms * 1e-3) if k_ms > 0 else None,
             "h2d_bytes": int(h2d), "inputs": "pageable numpy"}
        if extra:
            r.update(extra(filt, k_ms))
        rows.append(r)

    # 1: RadialNumberDensity O-O, 1000 O atoms (3k-atom TIP3P box), 100 frames, dr 0.1, maxR 15
    c = synth.config1()
    nb = synth.n_bins(c["dr"], c["maxR"])
    o1 = np.zeros(nb, np.float32)

    def f1():
        o1[:] = 0
        plf.py_getRadialNbrDensity(c["sel1"], c["sel1"], o1, c["cellDims"], 1, c["maxR"], c["dr"])
    row(1, "RDF O-O, 1000 O atoms x 100 frames, dr 0.1 maxR 15 (15 A in a 31 A cell, 46 % of the pairs in range)",
        0.5 * 1000 * 999 * 100, f1, 5, c["sel1"].nbytes + c["cellDims"].nbytes,
        lambda filt, k: {"filter": ("tensor-core" if filt == 100 else "fp32 %d" % filt) +
                                   (", all %d frames all-exact (dense range: every pair through the strict arithmetic)" % exact_frames()
                                    if exact_frames() else "")})
    rows.append(headline)
    # 3: ResidueWiseWaterDensity('protein', dr 0.1, maxR 15): 76 residues, 20 000 water O, 5000 frames -- ONE grouped call
    c3 = synth.config3(n_frames=5000)
    o3 = np.zeros((c3["nGroups"], nb), np.float32)

    def f3():
        plf.py_getRadialNbrDensityGrouped(c3["waters"], c3["protein"], c3["groupId"], o3, c3["cellDims"], 0, 15.0, 0.1)
    row(3, "residue-wise water RDF: 20000 water O x 1231 protein atoms (76 residues) x 5000 frames, one grouped call",
        20000.0 * 1231 * 5000, f3, 1, c3["waters"].nbytes + c3["protein"].nbytes + c3["cellDims"].nbytes,
        lambda filt, k: {"filter": "tensor-core" if filt == 100 else "fp32 %d" % filt})
    del c3
    # 4: getDistances 'protein and resid 53' (7 atoms) vs 'protein' (1231), 10 000 frames
    c4 = synth.config4(n_frames=10000)
    o4 = np.zeros((7, 1231), np.float32)

    def f4():
        plf.py_getDistances(c4["sel1"], c4["sel2"], o4, c4["cellDims"], 0)
    bytes4 = c4["sel1"].nbytes + c4["sel2"].nbytes
    hbm, hsrc = hbm_peak()
    row(4, "getDistances 7 x 1231 atoms x 10000 frames (strict float32 distance per pair and frame, float64 mean)",
        7.0 * 1231 * 10000, f4, 3, bytes4 + c4["cellDims"].nbytes,
        lambda filt, k: {"roofline": {"bound": "hbm", "kernel": "distances_kernel", "achieved": bytes4 / (k * 1e-3) / 1e9, "peak": hbm,
                                      "unit": "GB/s", "frac": bytes4 / (k * 1e-3) / 1e9 / hbm, "peak_source": hsrc,
                                      "algorithmic_bytes": bytes4,
                                      "note": "12 B per atom-frame feed 7 strict evaluations (~40 instructions each): the kernel is "
                                              "issue-bound at n1 = 7, see DESIGN.md 5.4; kernel_ms sums the launches of the upload chunks"}})
    # the same kernel on a resident trajectory (one launch, no upload)
    d1, d2 = torch.from_numpy(c4["sel1"]).to(E.dev), torch.from_numpy(c4["sel2"]).to(E.dev)
    dc = torch.from_numpy(c4["cellDims"]).to(E.dev)
    s = torch.zeros((7, 1231), dtype=torch.float64, device=E.dev)
    ks = []
    for _ in range(4):
        s.zero_()
        capi.check(lib.nda_dev_distances_sum(d1.data_ptr(), 7, d2.data_ptr(), 1231, s.data_ptr(), dc.data_ptr(), 10000, 0, 10000, 0, E.st))
        ks.append(kernel_ms(E)[0])
    kres = float(np.mean(ks[1:]))
    rows[-1]["resident_kernel"] = {"kernel_ms": kres, "pairs_per_s": 7.0 * 1231 * 10000 / (kres * 1e-3),
                                   "achieved_gbs": bytes4 / (kres * 1e-3) / 1e9, "frac_of_hbm": bytes4 / (kres * 1e-3) / 1e9 / hbm}
    del c4, d1, d2
    return rows


# ------------------------------------------------------------------------------------- N > 1 legs
def multi_gpu_legs(E, cfg):
    """Strong scaling: ONE fixed trajectory, frames split over the ranks.  Every leg carries a parity record
    against rank 0 computing the whole thing alone."""
    torch, dist, lib, capi, plf, synth = E.torch, E.dist, E.lib, E.capi, E.plf, E.synth
    from namdanalyzer_b200 import distributed as nd
    from namdanalyzer_b200.distributed import frame_range
    out = {}
    rank, world = E.rank, E.world
    gloo = dist.new_group(backend="gloo")                   # host-side barriers that keep the GPUs idle

    def hbar():
        torch.cuda.synchronize()
        dist.barrier(group=gloo)

    # ---- within, config 2, the SAME 1000 frames on every rank (seed 1002), each rank takes its frame columns
    c2 = synth.config2(n_frames=1000, seed=1002)
    rs, os_ = c2["refSel"], c2["outSel"]
    pairs2 = float(rs.size) * os_.size * 1000
    fb, fe = frame_range(1000, rank, world)
    words = (os_.size + 31) // 32
    d_xyz = torch.from_numpy(np.ascontiguousarray(c2["allAtoms"][:, fb:fe])).to(E.dev)
    d_cell = torch.from_numpy(np.ascontiguousarray(c2["cellDims"][fb:fe])).to(E.dev)
    d_rs, d_os = torch.from_numpy(rs).to(E.dev), torch.from_numpy(os_).to(E.dev)
    d_bits = torch.zeros((fe - fb, words), dtype=torch.int32, device=E.dev)

    def res():
        capi.check(lib.nda_dev_within(d_xyz.data_ptr(), c2["allAtoms"].shape[0], fe - fb, d_rs.data_ptr(), rs.size, d_os.data_ptr(),
                                      os_.size, d_bits.data_ptr(), None, d_cell.data_ptr(), 3.0, 0, fe - fb, E.st))
    for _ in range(3):
        res()
    hbar()
    reps = 50
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(E.stream)
    for _ in range(reps):
        res()
    e1.record(E.stream)
    torch.cuda.synchronize()
    ms = max_over_ranks(E, e0.elapsed_time(e1) / reps)
    # end to end: every rank holds the whole host array (as one process would) and uploads only its frame columns
    for _ in range(2):
        nd.within_bits(c2["allAtoms"], rs, os_, c2["cellDims"], 3.0)
    hbar()
    t0 = time.perf_counter()
    n_e2e = 8
    for _ in range(n_e2e):
        nd.within_bits(c2["allAtoms"], rs, os_, c2["cellDims"], 3.0)
    dt = max_over_ranks(E, (time.perf_counter() - t0) / n_e2e)
    _, _, full = nd.within_bits(c2["allAtoms"], rs, os_, c2["cellDims"], 3.0, gather=True)
    parity = None
    if rank == 0:
        alone = plf.py_getWithinBits(c2["allAtoms"], rs, os_, c2["cellDims"], 3.0)
        parity = bool(np.array_equal(alone, full))
    out["within_config2"] = {"workload": "config 2, ONE 1000-frame trajectory split over %d ranks (strong scaling), no collective" % world,
                             "resident_ms": ms, "resident_pairs_per_s": pairs2 / (ms * 1e-3),
                             "e2e_ms": dt * 1e3, "e2e_pairs_per_s": pairs2 / dt,
                             "parity": parity, "parity_how": "all-gathered bitmask (NCCL) == rank 0 computing all 1000 frames alone"}
    del d_xyz, c2
    hbar()

    # ---- distances, config 4, ONE 10 000-frame trajectory: float64 partial sums + one NCCL all-reduce
    c4 = synth.config4(n_frames=10000)
    o4 = np.zeros((7, 1231), np.float32)
    for _ in range(2):
        nd.distances(c4["sel1"], c4["sel2"], o4, c4["cellDims"], 0)
    hbar()
    t0 = time.perf_counter()
    for _ in range(5):
        nd.distances(c4["sel1"], c4["sel2"], o4, c4["cellDims"], 0)
    dt = max_over_ranks(E, (time.perf_counter() - t0) / 5)
    parity = None
    if rank == 0:
        alone = plf.py_getDistancesSum(c4["sel1"], c4["sel2"], c4["cellDims"], 0) / 10000.0
        parity = bool(np.max(np.abs(o4.astype(np.float64) - alone) / np.maximum(alone, 1e-30)) < 1e-6)
        fb4, fe4 = frame_range(10000, 0, world)
    out["distances_config4"] = {"workload": "config 4, 7 x 1231 x 10000 frames split over %d ranks, one NCCL all-reduce of 8617 float64" % world,
                                "e2e_ms": dt * 1e3, "e2e_pairs_per_s": 7.0 * 1231 * 10000 / dt,
                                "parity": parity, "parity_how": "mean matrix == rank 0 alone within 1e-6 relative (float64 sums re-associated)"}
    del c4
    hbar()

    # ---- RDF O-O, config 1 (1000 O x 100 frames, FP32 filter) and residue-wise water RDF, config 3 at 5000 frames:
    #      host arrays, frames sharded, one NCCL all-reduce of the integer histogram each
    c1 = synth.config1()
    nb = synth.n_bins(c1["dr"], c1["maxR"])
    for _ in range(2):
        h1 = nd.radial_nbr_counts(c1["sel1"], c1["sel1"], c1["cellDims"], 1, nb, c1["dr"])
    hbar()
    t0 = time.perf_counter()
    for _ in range(5):
        h1 = nd.radial_nbr_counts(c1["sel1"], c1["sel1"], c1["cellDims"], 1, nb, c1["dr"])
    dt = max_over_ranks(E, (time.perf_counter() - t0) / 5)
    parity = None
    if rank == 0:
        parity = bool(np.array_equal(h1, plf.py_getRadialNbrCounts(c1["sel1"], c1["sel1"], c1["cellDims"], 1, nb, c1["dr"]).astype(np.int64)))
    out["rdf_config1"] = {"workload": "config 1, RDF O-O 1000 atoms x 100 frames split over %d ranks, one NCCL all-reduce of 149 x i64" % world,
                          "e2e_ms": dt * 1e3, "e2e_pairs_per_s": 0.5 * 1000 * 999 * 100 / dt,
                          "parity": parity, "parity_how": "all-reduced histogram bit-equal to rank 0 alone"}
    hbar()
    c3 = synth.config3(n_frames=5000)
    g3 = nd.radial_nbr_counts(c3["waters"], c3["protein"], c3["cellDims"], 0, nb, 0.1, groupId=c3["groupId"], nGroups=c3["nGroups"])
    hbar()
    t0 = time.perf_counter()
    for _ in range(2):
        g3 = nd.radial_nbr_counts(c3["waters"], c3["protein"], c3["cellDims"], 0, nb, 0.1, groupId=c3["groupId"], nGroups=c3["nGroups"])
    dt = max_over_ranks(E, (time.perf_counter() - t0) / 2)
    parity = None
    if rank == 0:
        alone = plf.py_getRadialNbrCounts(c3["waters"], c3["protein"], c3["cellDims"], 0, nb, 0.1, groupId=c3["groupId"], nGroups=c3["nGroups"])
        parity = bool(np.array_equal(g3, alone.astype(np.int64)))
    out["rdf_config3"] = {"workload": "config 3, residue-wise water RDF 20000 x 1231 x 5000 frames (76 groups) split over %d ranks, "
                                      "one NCCL all-reduce of 76 x 149 x i64" % world,
                          "e2e_ms": dt * 1e3, "e2e_pairs_per_s": 20000.0 * 1231 * 5000 / dt,
                          "parity": parity, "parity_how": "all-reduced grouped histogram bit-equal to rank 0 alone"}
    del c3
    hbar()

    # ---- RDF, config-5 shape, frames generated on the devices: u64 histogram + one NCCL all-reduce.
    #      (i) ONE 64-frame trajectory with a parity pass on rank 0; (ii) config 5 at its stated length, 10 000 frames.
    n = N_SIDE5 ** 3
    Lbox = float(np.float32((n / 0.0334) ** (1.0 / 3.0)))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    counts = torch.zeros(149, dtype=torch.int64, device=E.dev)

    def sharded_pass(nF, reps, parity_pass):
        fb, fe = frame_range(nF, rank, world)
        own = fe - fb
        xyz = torch.empty((n, max(own, 1), 3), dtype=torch.float32, device=E.dev)      # only this rank's frames live here
        cell = torch.empty((max(own, 1), 3), dtype=torch.float32, device=E.dev)
        gen_frames(E, xyz, cell, own, fb, fe, Lbox)

        def go():
            counts.zero_()
            capi.check(lib.nda_dev_radial_nbr_counts(xyz.data_ptr(), n, xyz.data_ptr(), n, None, 1, counts.data_ptr(), 149,
                                                     cell.data_ptr(), own, 0, own, 1, 0.1, E.st))
            dist.all_reduce(counts, op=dist.ReduceOp.SUM)   # NCCL, on the current (= launching) stream
        if parity_pass:
            go()                                            # warm-up
        hbar()
        e0.record(E.stream)
        for _ in range(reps):
            go()
        e1.record(E.stream)
        torch.cuda.synchronize()
        ms = max_over_ranks(E, e0.elapsed_time(e1) / reps)
        sharded = counts.cpu().numpy().copy()
        del xyz, cell
        parity = None
        if parity_pass and rank == 0:
            full = torch.empty((n, nF, 3), dtype=torch.float32, device=E.dev)
            fcell = torch.empty((nF, 3), dtype=torch.float32, device=E.dev)
            gen_frames(E, full, fcell, nF, 0, nF, Lbox)
            alone = torch.zeros(149, dtype=torch.int64, device=E.dev)
            capi.check(lib.nda_dev_radial_nbr_counts(full.data_ptr(), n, full.data_ptr(), n, None, 1, alone.data_ptr(), 149,
                                                     fcell.data_ptr(), nF, 0, nF, 1, 0.1, E.st))
            torch.cuda.synchronize()
            parity = bool(np.array_equal(alone.cpu().numpy(), sharded))
            del full, fcell
        hbar()
        return ms, sharded, parity

    ms, sharded, parity = sharded_pass(64, 3, True)
    out["rdf_config5"] = {"workload": "config-5 shape: %d O atoms, ONE 64-frame trajectory split over %d ranks, one NCCL all-reduce of 149 x u64"
                                      % (n, world),
                          "ms_per_pass": ms, "pairs_per_s": 0.5 * n * (n - 1.0) * 64 / (ms * 1e-3), "counts_total": int(sharded.sum()),
                          "parity": parity, "parity_how": "all-reduced histogram bit-equal to rank 0 computing all 64 frames alone"}
    if E.args.rdf_full_frames > 0:
        nFull = E.args.rdf_full_frames
        ms, sharded, _ = sharded_pass(nFull, 1, False)
        out["rdf_config5_full"] = {"workload": "BASELINE configs[4] at its stated length: %d O atoms x %d frames (%.1f GB of coordinates generated in "
                                               "HBM, %d frames per rank), one pass, one NCCL all-reduce" % (n, nFull, n * nFull * 12 / 1e9, nFull // world),
                                   "seconds_per_pass": ms * 1e-3, "pairs_per_s": 0.5 * n * (n - 1.0) * nFull / (ms * 1e-3),
                                   "counts_total": int(sharded.sum()),
                                   "counts_per_frame_consistent_with_64_frame_run": bool(abs(sharded.sum() / nFull - out["rdf_config5"]["counts_total"] / 64.0)
                                                                                         < 2e-3 * sharded.sum() / nFull)}
    hbar()

    # ---- the box's concurrent pinned-H2D ceiling: every rank copies 110 MB of contiguous pinned memory at the same time
    nbytes = 110 * (1 << 20)
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d = torch.empty(nbytes, dtype=torch.uint8, device=E.dev)
    d.copy_(h, non_blocking=True)
    hbar()
    reps = 10
    e0.record(E.stream)
    for _ in range(reps):
        d.copy_(h, non_blocking=True)
    e1.record(E.stream)
    torch.cuda.synchronize()
    ms_all = max_over_ranks(E, e0.elapsed_time(e1) / reps)
    hbar()
    ms_one = None
    if rank == 0:                                           # and alone, the other ranks idle
        e0.record(E.stream)
        for _ in range(reps):
            d.copy_(h, non_blocking=True)
        e1.record(E.stream)
        torch.cuda.synchronize()
        ms_one = e0.elapsed_time(e1) / reps
    hbar()
    out["h2d_ceiling"] = {"bytes_per_rank": nbytes, "concurrent_aggregate_gbs": nbytes * world / (ms_all * 1e-3) / 1e9,
                          "concurrent_per_gpu_gbs": nbytes / (ms_all * 1e-3) / 1e9,
                          "one_rank_alone_gbs": nbytes / (ms_one * 1e-3) / 1e9 if ms_one else None,
                          "how": "one contiguous pinned cudaMemcpyAsync of 110 MB per rank, all ranks at once, max over ranks (CUDA events)"}
    del h, d

    # ---- the drop-in call itself sharding over all GPUs inside ONE process (rank 0; the other ranks wait on the host)
    drop = None
    if rank == 0:
        # the other ranks are parked on a host barrier: this process may use the host threads a single process would
        capi.set_option("NDA_HOST_THREADS", str(max(1, min(16, os.cpu_count() or 1))))
        c2 = synth.config2(n_frames=1000, seed=1002)
        rs, os_ = c2["refSel"], c2["outSel"]
        xyz_np, cell_np = c2["allAtoms"], c2["cellDims"]
        out_np = np.zeros(xyz_np.shape[:2], np.int32)
        plf.py_getWithin(xyz_np, rs, os_, out_np, cell_np, 3.0)
        one_ms = timed_calls(lambda: plf.py_getWithin(xyz_np, rs, os_, out_np, cell_np, 3.0), 5, torch.cuda.synchronize) * 1e3
        bits_one = plf.py_getWithinBits(xyz_np, rs, os_, cell_np, 3.0)
        try:
            capi.set_devices(list(range(world)))
            for _ in range(2):
                plf.py_getWithin(xyz_np, rs, os_, out_np, cell_np, 3.0)
            all_ms = timed_calls(lambda: plf.py_getWithin(xyz_np, rs, os_, out_np, cell_np, 3.0), 8, torch.cuda.synchronize) * 1e3
            bits_all = plf.py_getWithinBits(xyz_np, rs, os_, cell_np, 3.0)
            # the same call on the same array after namdanalyzer_b200.pin(): every device DMAs its own frame columns
            # over its own link, no gather through host staging (which is bounded by the host's memory bandwidth)
            import namdanalyzer_b200 as nda
            nda.pin(xyz_np)
            try:
                for _ in range(2):
                    plf.py_getWithin(xyz_np, rs, os_, out_np, cell_np, 3.0)
                pin_all_ms = timed_calls(lambda: plf.py_getWithin(xyz_np, rs, os_, out_np, cell_np, 3.0), 8, torch.cuda.synchronize) * 1e3
                bits_pin = plf.py_getWithinBits(xyz_np, rs, os_, cell_np, 3.0)
                capi.check(lib.nda_set_device(E.local_rank))
                for _ in range(2):
                    plf.py_getWithin(xyz_np, rs, os_, out_np, cell_np, 3.0)
                pin_one_ms = timed_calls(lambda: plf.py_getWithin(xyz_np, rs, os_, out_np, cell_np, 3.0), 5, torch.cuda.synchronize) * 1e3
                capi.set_devices(list(range(world)))
            finally:
                nda.unpin(xyz_np)
            c1 = synth.config1()
            nb = synth.n_bins(c1["dr"], c1["maxR"])
            h_all = plf.py_getRadialNbrCounts(c1["sel1"], c1["sel1"], c1["cellDims"], 1, nb, c1["dr"])
            coll = int(lib.nda_last_collective())
            c4 = synth.config4(n_frames=10000)
            s_all = plf.py_getDistancesSum(c4["sel1"], c4["sel2"], c4["cellDims"], 0)
            d_ms = timed_calls(lambda: plf.py_getDistancesSum(c4["sel1"], c4["sel2"], c4["cellDims"], 0), 3, torch.cuda.synchronize) * 1e3
        finally:
            capi.check(lib.nda_set_device(E.local_rank))
        h_one = plf.py_getRadialNbrCounts(c1["sel1"], c1["sel1"], c1["cellDims"], 1, nb, c1["dr"])
        s_one = plf.py_getDistancesSum(c4["sel1"], c4["sel2"], c4["cellDims"], 0)
        d1_ms = timed_calls(lambda: plf.py_getDistancesSum(c4["sel1"], c4["sel2"], c4["cellDims"], 0), 3, torch.cuda.synchronize) * 1e3
        capi.set_option("NDA_HOST_THREADS", None)
        drop = {"api": "nda_set_devices(0..%d) + py_getWithin / py_getRadialNbrCounts / py_getDistancesSum, plain numpy, ONE process" % (world - 1),
                "within_config2_e2e_ms": all_ms, "within_config2_e2e_pairs_per_s": pairs2 / (all_ms * 1e-3),
                "within_config2_one_gpu_e2e_ms": one_ms, "speedup_vs_one_gpu": one_ms / all_ms,
                "note": "pageable input: the rows are gathered into pinned staging by host threads, a host-memory-bandwidth "
                        "bound that more GPUs cannot lift; page-locked input (namdanalyzer_b200.pin) is pulled by every device "
                        "over its own link",
                "pinned": {"within_config2_e2e_ms": pin_all_ms, "within_config2_e2e_pairs_per_s": pairs2 / (pin_all_ms * 1e-3),
                           "within_config2_one_gpu_e2e_ms": pin_one_ms, "speedup_vs_one_gpu": pin_one_ms / pin_all_ms,
                           "within_bits_equal_one_gpu": bool(np.array_equal(bits_one, bits_pin))},
                "distances_config4_e2e_ms": d_ms, "distances_config4_one_gpu_e2e_ms": d1_ms,
                "collective": {0: "none", 1: "ncclAllReduce (ncclCommInitAll)", 2: "host sum"}.get(coll, str(coll)),
                "parity": {"within_bits_equal_one_gpu": bool(np.array_equal(bits_one, bits_all)),
                           "rdf_counts_equal_one_gpu": bool(np.array_equal(h_one, h_all)),
                           "distance_sums_rel_diff": float(np.max(np.abs(s_all - s_one) / np.maximum(s_one, 1e-30)))}}
    hbar()
    out["dropin_multi_gpu"] = drop
    return out


# ------------------------------------------------------------------------------------- our arm
def run_ours(args, rank, world, local_rank):
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else None     # before any host buffer exists
    import torch
    import torch.distributed as dist
    from namdanalyzer_b200 import synth, _capi
    from namdanalyzer_b200.lib import pylibFuncs as plf

    torch.cuda.set_device(local_rank)
    lib = _capi.load()
    _capi.check(lib.nda_set_device(local_rank))
    dev = torch.device("cuda", local_rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    nF = args.frames
    cfg = synth.config2(n_frames=nF, seed=1002 + 7919 * rank)
    nA = cfg["allAtoms"].shape[0]
    rs, os_ = cfg["refSel"], cfg["outSel"]
    pairs_step = float(rs.size) * float(os_.size) * nF
    words = (os_.size + 31) // 32

    xyz_np, cell_np = cfg["allAtoms"], cfg["cellDims"]      # plain (pageable) numpy: what a drop-in caller passes
    out_np = np.zeros((nA, nF), np.int32)
    d_xyz = torch.from_numpy(xyz_np).to(dev)
    d_cell = torch.from_numpy(cell_np).to(dev)
    d_rs = torch.from_numpy(rs).to(dev)
    d_os = torch.from_numpy(os_).to(dev)
    d_bits = torch.zeros((nF, words), dtype=torch.int32, device=dev)
    d_out = torch.zeros((nA, nF), dtype=torch.int32, device=dev)
    # an explicit (non-default) stream: the library treats a NULL stream as "use my own", and the
    # CUDA events below must sit on the stream the kernels are launched on
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    st = ctypes.c_void_p(stream.cuda_stream)
    assert stream.cuda_stream != 0
    torch.cuda.synchronize()                 # inputs above were staged on the default stream

    E = Env()
    E.torch, E.dist, E.lib, E.capi, E.plf, E.synth = torch, dist, lib, _capi, plf, synth
    E.dev, E.stream, E.st, E.rank, E.world, E.local_rank, E.args = dev, stream, st, rank, world, local_rank, args

    def step_resident():
        _capi.check(lib.nda_dev_within(d_xyz.data_ptr(), nA, nF, d_rs.data_ptr(), rs.size, d_os.data_ptr(), os_.size,
                                       d_bits.data_ptr(), d_out.data_ptr(), d_cell.data_ptr(), float(cfg["distance"]),
                                       0, nF, st))

    warm = max(args.warmup, 3)
    for _ in range(warm):
        step_resident()
    barrier()

    # ---- timed: resident inputs, EXACTLY --steps steps; then the same loop for >= sustain-s seconds
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.15)
    launches0 = int(lib.nda_launch_count())
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_begin = time.perf_counter()
    ev0.record(stream)
    for _ in range(args.steps):
        step_resident()                      # back to back, no host synchronisation inside the timed region
    ev1.record(stream)
    barrier()
    launches = int(lib.nda_launch_count()) - launches0
    ms_total = max_over_ranks(E, ev0.elapsed_time(ev1))
    value = pairs_step * world * args.steps / (ms_total * 1e-3)
    ms_step = ms_total / args.steps
    n_sus = max(args.steps, int(math.ceil(args.sustain_s * 1e3 / ms_step)))
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record(stream)
    for _ in range(n_sus):
        step_resident()
    s1.record(stream)
    barrier()
    t_end = time.perf_counter()
    sus_ms = max_over_ranks(E, s0.elapsed_time(s1))
    clocks = sampler.stop(t_begin, t_end) if rank == 0 else None
    sustained = {"value": pairs_step * world * n_sus / (sus_ms * 1e-3), "unit": "pairs/s", "steps": n_sus,
                 "seconds": sus_ms * 1e-3, "ms_per_step": sus_ms / n_sus,
                 "note": "same loop as `value`, run for >= %.1f s right after it; `clocks` covers both" % args.sustain_s}

    # the dominant kernel's own duration (CUDA events the library records around it on the launching
    # stream), measured live on the same inputs right after the timed region
    kern_ms = []
    for _ in range(5):
        step_resident()
        kern_ms.append(kernel_ms(E)[0])
    hits_dev = int(d_out.sum().item())

    # ---- timed: end to end through the reference-shaped operator, host buffers (pageable first, pinned second)
    def e2e_call(x, c, o):
        plf.py_getWithin(x, rs, os_, o, c, cfg["distance"])
    for _ in range(2):
        e2e_call(xyz_np, cell_np, out_np)
    barrier()
    e2e_steps = 10
    dt = max_over_ranks(E, timed_calls(lambda: e2e_call(xyz_np, cell_np, out_np), e2e_steps, torch.cuda.synchronize))
    e2e_value = pairs_step * world / dt
    hits_e2e = int(out_np.sum())
    h_xyz = torch.from_numpy(xyz_np).pin_memory()
    h_cell = torch.from_numpy(cell_np).pin_memory()
    h_out = torch.zeros((nA, nF), dtype=torch.int32).pin_memory()
    px, pc, po = h_xyz.numpy(), h_cell.numpy(), h_out.numpy()
    for _ in range(2):
        e2e_call(px, pc, po)
    barrier()
    dtp = max_over_ranks(E, timed_calls(lambda: e2e_call(px, pc, po), e2e_steps, torch.cuda.synchronize))
    hits_pinned = int(po.sum())
    del h_xyz, h_out, px, po

    multi = None
    if world > 1 and not args.no_extras:
        multi = multi_gpu_legs(E, cfg)

    if rank != 0:
        return

    # ---- parity inside the run: a sample of frames against the oracle (the checker, never the thing measured)
    from oracle import port
    sample = [0, nF // 3, nF // 2, nF - 1]
    want = port.within(np.ascontiguousarray(xyz_np[:, sample]), rs, os_, np.ascontiguousarray(cell_np[sample]), cfg["distance"])
    got = d_out[:, sample].cpu().numpy()
    diff = np.argwhere(want != got)
    edge_pairs = {"count": int(diff.shape[0]), "frames_checked": sample,
                  "list": [[int(a), int(sample[f])] for a, f in diff[:32]],
                  "how": "mask of %d sampled frames x 25000 atoms against oracle/ (strict IEEE restatement of getWithin.cpp); "
                         "north_star allows pairs within 1 ulp of the cut-off to differ and asks for them to be listed" % len(sample)}

    # ---- roofline of the dominant kernel
    peak0, peak1 = ctypes.c_double(0), ctypes.c_double(0)
    _capi.check(lib.nda_measure_fp32_peak(0, ctypes.byref(peak0)))
    _capi.check(lib.nda_measure_fp32_peak(1, ctypes.byref(peak1)))
    peak = max(peak0.value, peak1.value)
    k_ms = float(np.mean(kern_ms))
    achieved = pairs_step * FLOP_WITHIN / (k_ms * 1e-3) / 1e12
    filt = int(lib.nda_last_filter())
    tensor_path = filt == 100
    kname = "within_tc_kernel" if tensor_path else "within_kernel"
    traffic, traffic_src = committed_traffic(kname)
    fp32_src = ("FFMA loop measured live in this run by the library's own nda_measure_fp32_peak (scalar %.1f, f32x2 %.1f TFLOP/s); "
                "MEASURED_PEAKS.json has no FP32 figure; nominal 2*128*148*1.965 GHz = %.1f" % (peak0.value, peak1.value, FP32_NOMINAL_TFLOPS))
    fp32_view = {"achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if peak else None,
                 "frac_of_nominal": achieved / FP32_NOMINAL_TFLOPS, "flop_per_pair": FLOP_WITHIN, "peak_source": fp32_src,
                 "note": "algorithmic flops of the REFERENCE arithmetic (SURVEY 8d: 21 per within pair) over the FP32 "
                         "peak; above 1 when the filter runs on the tensor cores -- not a roofline, a conversion"}
    common = {"kernel": kname, "pairs_per_launch": pairs_step, "kernel_ms": k_ms,
              "kernel_share_of_step": k_ms / ms_step if world == 1 else None,
              "kernel_pairs_per_s": pairs_step / (k_ms * 1e-3), "traffic": traffic,
              "traffic_source": (traffic_src or "none") + " (a committed ncu --set full capture, per launch; not measured in this run)"}
    if tensor_path:
        tpeak, tsus, tsrc = tensor_peak()
        t_ach = pairs_step * 32.0 / (k_ms * 1e-3) / 1e12
        sm_count = torch.cuda.get_device_properties(dev).multi_processor_count
        roofline = dict(common, bound="tensor", achieved=t_ach, peak=tpeak, unit="TFLOP/s", frac=t_ach / tpeak,
                        frac_of_sustained_peak=(t_ach / tsus) if tsus else None,
                        flop_per_pair=32.0, peak_source=tsrc,
                        note="one tcgen05.mma M=128 N=256 K=16 per 32768 pairs (12 of the 16 K slots carry the embedding, 2 the "
                             "threshold); every accumulator is read back exactly once, so the kernel is bound by TMEM capacity / "
                             "hand-over latency (65536 cells per SM over a ~950-clock round trip) and the epilogue's ALU work, "
                             "not by MMA rate: see DESIGN.md 5.7",
                        tmem_cells_per_clk_per_sm={
                            "achieved": pairs_step / (k_ms * 1e-3) / (sm_count * clocks["sm_mhz"] * 1e6)
                            if (clocks and clocks.get("sm_mhz")) else None,
                            "mma_write_peak": 256.0, "tmem_read_peak_measured": 230.0,
                            "how": "tools/tmem_ld_rate.cu (16 warps, tcgen05.ld 32x32b.x32.pack::16b) on B200"},
                        fp32_equivalent=fp32_view)
    else:
        roofline = dict(common, bound="fp32", **{k: fp32_view[k] for k in ("achieved", "peak", "unit", "frac", "frac_of_nominal",
                                                                           "flop_per_pair", "peak_source")})

    # ---- the same resident step with the packed-FP32 filter (the north-star design), for comparison
    fp32_arm = None
    if tensor_path and not args.no_extras:
        with _capi.options(NDA_FILTER="fold2a1"):
            for _ in range(2):
                step_resident()
            ks = []
            for _ in range(3):
                step_resident()
                ks.append(kernel_ms(E)[0])
            fk = float(np.mean(ks))
            fp32_arm = {"filter": "fold2a1 (packed FP32, FFMA2/FADD2)", "kernel": "within_kernel", "kernel_ms": fk,
                        "kernel_pairs_per_s": pairs_step / (fk * 1e-3),
                        "roofline": {"bound": "fp32", "achieved": pairs_step * FLOP_WITHIN / (fk * 1e-3) / 1e12, "peak": peak, "unit": "TFLOP/s",
                                     "frac": pairs_step * FLOP_WITHIN / (fk * 1e-3) / 1e12 / peak if peak else None,
                                     "flop_per_pair": FLOP_WITHIN, "peak_source": fp32_src},
                        "hits": int(d_out.sum().item())}
        step_resident()

    h2d = int(np.union1d(rs, os_).size * nF * 12 + cell_np.nbytes + rs.nbytes + os_.nbytes)
    line = {
        "metric": "min-image pair evals/sec (within, config 2)", "value": value, "unit": "pairs/s",
        "n_gpus": world, "steps": args.steps, "warmup": warm, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args),
        "sustained": sustained,
        "e2e": {"value": e2e_value, "unit": "pairs/s", "ms_per_call": dt * 1e3,
                # only the rows of refSel | outSel atoms travel (gathered by host threads into pinned staging, one DMA per chunk)
                "h2d_bytes_per_step": h2d, "host_array_bytes": int(xyz_np.nbytes),
                "d2h_bytes_per_step": int(nF * words * 4), "steps": e2e_steps,
                "h2d_gbs": h2d * world / dt / 1e9,
                "api": "namdanalyzer_b200.lib.pylibFuncs.py_getWithin(allAtoms, refSel, outSel, out, cellDims, 3.0) on plain "
                       "(pageable) numpy arrays -- what the reference's callers pass",
                "pinned": {"value": pairs_step * world / dtp, "ms_per_call": dtp * 1e3, "h2d_gbs": h2d * world / dtp / 1e9,
                           "api": "the same call on page-locked (cudaHostAlloc) arrays: strided 2-D DMA, no host gather"}},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": roofline,
        "check": {"hits_resident": hits_dev, "hits_e2e": hits_e2e, "hits_e2e_pinned": hits_pinned,
                  "oracle_frames": len(sample), "oracle_mismatches": int(diff.shape[0])},
        "edge_pairs": edge_pairs,
        "numa_binding": numa,
        "filter": "tensor-core (tcgen05) periodic-embedding filter + strict FP32 re-evaluation" if tensor_path
                  else "packed-FP32 filter %d + strict FP32 re-evaluation" % filt,
    }
    if fp32_arm is not None:
        line["fp32_filter_arm"] = fp32_arm
    if multi is not None:
        line["strong_scaling"] = {k: v for k, v in multi.items() if k not in ("h2d_ceiling", "dropin_multi_gpu")}
        line["h2d_ceiling"] = multi["h2d_ceiling"]
        line["dropin_multi_gpu"] = multi["dropin_multi_gpu"]
        line["e2e"]["h2d_fraction_of_ceiling"] = (h2d * world / dt / 1e9) / multi["h2d_ceiling"]["concurrent_aggregate_gbs"]
        line["e2e"]["pinned"]["h2d_fraction_of_ceiling"] = (h2d * world / dtp / 1e9) / multi["h2d_ceiling"]["concurrent_aggregate_gbs"]
    if not args.no_extras:
        del d_xyz, d_out
        torch.cuda.empty_cache()
        if world == 1:
            line["rdf"] = rdf_leg(E, peak)
            headline_row = {"config": 2, "workload": line["config"]["workload"], "pairs": pairs_step, "e2e_ms": dt * 1e3,
                            "e2e_pairs_per_s": pairs_step / dt, "kernel_ms": k_ms, "kernel_launches": 1,
                            "kernel_pairs_per_s": pairs_step / (k_ms * 1e-3), "h2d_bytes": h2d, "inputs": "pageable numpy",
                            "filter": "tensor-core" if tensor_path else "fp32 %d" % filt,
                            "note": "kernel_ms = one resident launch over 1000 frames; the e2e call runs ~6 chunk launches"}
            rows = configs_leg(E, headline_row)
            rows.append({"config": 5, "workload": line["rdf"]["workload"], "pairs": line["rdf"]["pairs_per_step"],
                         "resident_ms": line["rdf"]["ms_per_step"], "resident_pairs_per_s": line["rdf"]["value"],
                         "kernel_ms": line["rdf"]["roofline"]["kernel_ms"], "kernel_pairs_per_s": line["rdf"]["roofline"]["kernel_pairs_per_s"],
                         "e2e_pairs_per_s": line["rdf"]["e2e"]["value"], "e2e_frames": line["rdf"]["e2e"]["frames"],
                         "inputs": "frames generated on the device (a 10 000-frame host array would be 39 GB); e2e on %d frames from pageable numpy"
                                   % line["rdf"]["e2e"]["frames"], "filter": line["rdf"]["filter"],
                         "full_size": line["rdf"].get("full_size")})
            line["configs"] = rows
        threads = os.cpu_count() or 1
        if world == 1:
            line["rdf"]["cpu_baseline"] = cpu_rdf_sample(threads)
        smp = 24
        scfg = synth.config2(n_frames=smp)
        cpu_within_sample(scfg, range(0, 4), threads)
        reps = 1
        rate, cdt = cpu_within_sample(scfg, range(smp), threads, reps)
        while cdt < 8.0 and reps < 64:                       # bounded sample: ~10-30 s of CPU work
            reps *= 2
            rate, cdt = cpu_within_sample(scfg, range(smp), threads, reps)
        if world == 1:
            try:
                line["legacy_gpu"] = legacy_gpu_sample()
            except Exception as e:                           # never let the optional arm break the bench line
                line["legacy_gpu"] = {"unavailable": repr(e)[:200]}
        line["cpu_baseline"] = {"value": rate, "unit": "pairs/s", "cores": threads, "kind": "reference",
                                "sample": "reference OpenMP getWithin (oracle/_ref/libref_fast.so, shipped flags), "
                                          "one frame per call, %d frames x %d repeats of config 2 (%.1f s)"
                                          % (smp, reps, cdt)}
    emit(line)


_REAL_STDOUT = None


def emit(line):
    """the ONE JSON line goes to the real stdout; everything else (NCCL banners, library chatter) was
    redirected to stderr by main()"""
    data = (json.dumps(line) + "\n").encode()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    args = parse()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
