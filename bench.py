#!/usr/bin/env python
"""bench.py -- the driver's measurement contract for the minimum-image pair path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config.workload): BASELINE.json configs[1] -- selection 'name OH2 and within 3 of
protein' on a ubiquitin-in-water box, N = 25 000 atoms, 1231 reference atoms x 7923 water O,
1000 frames (namdanalyzer_b200.synth.config2, seed 1002).  One STEP = one pass of the within path
over that 1000-frame batch = 9.75e9 minimum-image pair evaluations.  With N GPUs every rank owns
its own 1000-frame batch (frames shard with no data-path collective): scaling "weak".

  value  pair evaluations / s, inputs resident in HBM (reference layout), timed with CUDA events on
         the launching stream: pack + prep + within kernel + mask expansion.
  e2e    the same metric through the reference-shaped operator py_getWithin with HOST buffers:
         host->device copies of coordinates / selections / cells and the device->host read of the
         bitmask are inside the timed region.
  roofline      the within pair kernel against the FP32 pipe (21 algorithmic flop / pair,
                SURVEY.md section 8d); peak = FFMA rate measured live on this GPU.
  cpu_baseline  the reference's own OpenMP getWithin (oracle/_ref/libref_fast.so, shipped flags),
                one frame per call, all host threads, on a bounded sample of the same frames.
  rdf    (extra) the RDF kernel on a config-5 shaped water box generated on the device.

`--impl reference` times only the reference CPU implementation on bounded samples of the workload.
"""
import argparse
import ctypes
import json
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


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=1000, help="frames per step (BASELINE: 1000)")
    ap.add_argument("--no-extras", action="store_true", help="skip the RDF extra and the CPU baseline")
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
            self.lines.append(ln.strip())

    def stop(self):
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
        for ln in self.lines:
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
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


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
    rates, times = [], []
    for s in range(args.steps):
        r, dt = cpu_within_sample(cfg, range(sample_frames), threads)
        rates.append(r); times.append(dt)
    pairs_step = float(cfg["refSel"].size) * cfg["outSel"].size * sample_frames
    total = sum(times)
    value = pairs_step * args.steps / total
    line = {
        "impl": "reference", "metric": "min-image pair evals/sec (within, config 2)", "value": value,
        "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, sample="%d of the 1000 frames per step" % sample_frames),
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": threads, "kind": "reference",
                         "sample": "reference OpenMP getWithin (oracle/_ref/libref_fast.so, -O2 -fopenmp -ffast-math), "
                                   "one frame per call, %d frames x 1231 x 7923 pairs per step" % sample_frames},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def workload_config(args, sample=None):
    c = {"workload": "BASELINE configs[1]: 'name OH2 and within 3 of protein', 25000 atoms "
                     "(1231 ref x 7923 out), %d frames per GPU per step, synthetic seed 1002" % args.frames,
         "pairs_per_step_per_gpu": 1231.0 * 7923.0 * args.frames,
         "l2_policy": "inputs (300 MB raw coordinates + ~450 MB packed coordinates and fp16 embeddings per step) exceed the 126 MB L2; no explicit flush",
         "frame_sharding": "each rank owns its own frame batch; no data-path collective"}
    if sample:
        c["sample"] = sample
    return c


def bind_to_gpu_numa_node(local_rank):
    """One process per GPU: run (and so allocate pinned host buffers) on the CPU socket the GPU hangs off.
    With 8 ranks pulling 8 x 47 GB/s out of host memory, buffers on the far socket cost the end-to-end
    number more than anything on the GPU side.  Best effort: returns a description or None."""
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

    # ---- pinned host copies (e2e) and resident device copies (value)
    h_xyz = torch.from_numpy(cfg["allAtoms"]).pin_memory()
    h_cell = torch.from_numpy(cfg["cellDims"]).pin_memory()
    h_out = torch.zeros((nA, nF), dtype=torch.int32).pin_memory()
    d_xyz = h_xyz.to(dev, non_blocking=True)
    d_cell = h_cell.to(dev)
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

    def step_resident():
        _capi.check(lib.nda_dev_within(d_xyz.data_ptr(), nA, nF, d_rs.data_ptr(), rs.size, d_os.data_ptr(), os_.size,
                                       d_bits.data_ptr(), d_out.data_ptr(), d_cell.data_ptr(), float(cfg["distance"]),
                                       0, nF, st))

    for _ in range(max(args.warmup, 3)):
        step_resident()
    barrier()

    # ---- timed: resident inputs
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    kern_ms, launches0 = [], int(lib.nda_launch_count())
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    for _ in range(args.steps):
        step_resident()                      # back to back, no host synchronisation inside the timed region
    ev1.record(stream)
    barrier()
    launches = int(lib.nda_launch_count()) - launches0
    ms_total = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    # the dominant kernel's own duration (CUDA events the library records around it on the launching
    # stream), measured live on the same inputs right after the timed region
    for _ in range(min(args.steps, 5)):
        step_resident()
        ms = ctypes.c_double(0.0)
        nl = ctypes.c_int(0)
        _capi.check(lib.nda_last_kernel_ms(ctypes.byref(ms), ctypes.byref(nl)))   # waits for this step's kernel
        kern_ms.append(ms.value)

    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = pairs_step * world * args.steps / (ms_total * 1e-3)

    # ---- timed: end to end through the reference-shaped operator, host buffers
    xyz_np, cell_np, out_np = h_xyz.numpy(), h_cell.numpy(), h_out.numpy()
    for _ in range(2):
        plf.py_getWithin(xyz_np, rs, os_, out_np, cell_np, cfg["distance"])
    barrier()
    e2e_steps = max(3, min(args.steps, 10))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        plf.py_getWithin(xyz_np, rs, os_, out_np, cell_np, cfg["distance"])
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = pairs_step * world * e2e_steps / float(t.item())
    hits_e2e = int(out_np.sum())
    hits_dev = int(d_out.sum().item())

    rdf_sharded = None
    if world > 1 and not args.no_extras:
        rdf_sharded = rdf_sharded_extra(lib, _capi, torch, dist, dev, st, rank, world)

    if rank != 0:
        return

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
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic_latest.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get(kname + "_dram_bytes_per_launch")
        except Exception:
            traffic = None
    fp32_view = {"achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if peak else None,
                 "frac_of_nominal": achieved / FP32_NOMINAL_TFLOPS, "flop_per_pair": FLOP_WITHIN,
                 "peak_source": "FFMA loop measured live on this GPU (scalar %.1f, f32x2 %.1f TFLOP/s); "
                                "MEASURED_PEAKS.json has no FP32 figure; nominal 2*128*148*1.965 GHz = %.1f"
                                % (peak0.value, peak1.value, FP32_NOMINAL_TFLOPS),
                 "note": "algorithmic flops of the REFERENCE arithmetic (SURVEY 8d: 21 per within pair) over the FP32 "
                         "peak; above 1 when the filter runs on the tensor cores"}
    common = {"kernel": kname, "pairs_per_launch": pairs_step, "kernel_ms": k_ms,
              "kernel_share_of_step": k_ms * args.steps / ms_total if world == 1 else None,
              "kernel_pairs_per_s": pairs_step / (k_ms * 1e-3), "traffic": traffic}
    if tensor_path:
        # executed tensor flops: one tcgen05.mma (kind::f16, K = 16) row x column = 2 * 16 flop per pair
        tpeak, tsrc = 1590.0, "fallback of B200_PROFILING.md"
        mp = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(mp):
            try:
                tpeak, tsrc = float(json.load(open(mp))["bf16_tflops"]), "MEASURED_PEAKS.json bf16_tflops (burst; kernel timed alone)"
            except Exception:
                pass
        t_ach = pairs_step * 32.0 / (k_ms * 1e-3) / 1e12
        roofline = dict(common, bound="tensor", achieved=t_ach, peak=tpeak, unit="TFLOP/s", frac=t_ach / tpeak,
                        flop_per_pair=32.0, peak_source=tsrc,
                        note="one tcgen05.mma M=128 N=256 K=16 per 32768 pairs (12 of the 16 K slots carry the embedding, 2 the "
                             "threshold); every accumulator is read back exactly once, so the kernel is bound by the TMEM "
                             "hand-over (2 buffers of 256 columns) and the epilogue's ALU work, not by MMA rate: see DESIGN.md 5.7",
                        tmem_cells_per_clk_per_sm={
                            "achieved": pairs_step / (k_ms * 1e-3) / (torch.cuda.get_device_properties(dev).multi_processor_count
                                                                      * clocks["sm_mhz"] * 1e6)
                            if (clocks and clocks.get("sm_mhz")) else None,
                            "mma_write_peak": 256.0, "tmem_read_peak_measured": 230.0,
                            "how": "tools/tmem_ld_rate.cu (16 warps, tcgen05.ld 32x32b.x32.pack::16b) on B200"},
                        fp32_equivalent=fp32_view)
    else:
        roofline = dict(common, bound="fp32", **{k: fp32_view[k] for k in ("achieved", "peak", "unit", "frac", "frac_of_nominal",
                                                                           "flop_per_pair", "peak_source")})

    # ---- the same resident step with the packed-FP32 filter (the round-1 kernel), for comparison
    fp32_arm = None
    if tensor_path and not args.no_extras and world == 1:
        os.environ["NDA_FILTER"] = "fold2a1"
        try:
            for _ in range(2):
                step_resident()
            ks = []
            for _ in range(3):
                step_resident()
                ms = ctypes.c_double(0.0)
                _capi.check(lib.nda_last_kernel_ms(ctypes.byref(ms), None))
                ks.append(ms.value)
            fk = float(np.mean(ks))
            fp32_arm = {"filter": "fold2a1 (packed FP32, FFMA2/FADD2)", "kernel": "within_kernel", "kernel_ms": fk,
                        "kernel_pairs_per_s": pairs_step / (fk * 1e-3),
                        "frac_of_fp32_peak": pairs_step * FLOP_WITHIN / (fk * 1e-3) / 1e12 / peak if peak else None,
                        "hits": int(d_out.sum().item())}
        finally:
            del os.environ["NDA_FILTER"]

    line = {
        "metric": "min-image pair evals/sec (within, config 2)", "value": value, "unit": "pairs/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args),
        "e2e": {"value": e2e_value, "unit": "pairs/s",
                # only the rows of refSel | outSel atoms travel (strided 2-D copies of the host array)
                "h2d_bytes_per_step": int(np.union1d(rs, os_).size * nF * 12 + cell_np.nbytes + rs.nbytes + os_.nbytes),
                "host_array_bytes": int(xyz_np.nbytes),
                "d2h_bytes_per_step": int(nF * words * 4), "steps": e2e_steps,
                "api": "namdanalyzer_b200.lib.pylibFuncs.py_getWithin (pinned host arrays)"},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": roofline,
        "check": {"hits_resident": hits_dev, "hits_e2e": hits_e2e},
        "numa_binding": numa,
        "filter": "tensor-core (tcgen05) periodic-embedding filter + strict FP32 re-evaluation" if tensor_path
                  else "packed-FP32 filter %d + strict FP32 re-evaluation" % filt,
    }
    if fp32_arm is not None:
        line["fp32_filter_arm"] = fp32_arm

    if rdf_sharded is not None:
        line["rdf_sharded"] = rdf_sharded
    if not args.no_extras and world == 1:
        line["rdf"] = rdf_extra(lib, _capi, torch, dev, st, peak)
        threads = os.cpu_count() or 1
        sample = 24
        scfg = synth.config2(n_frames=sample)
        cpu_within_sample(scfg, range(0, 4), threads)
        reps = 1
        rate, dt = cpu_within_sample(scfg, range(sample), threads, reps)
        while dt < 8.0 and reps < 64:                       # bounded sample: ~10-30 s of CPU work
            reps *= 2
            rate, dt = cpu_within_sample(scfg, range(sample), threads, reps)
        try:
            line["legacy_gpu"] = legacy_gpu_sample()
        except Exception as e:                               # never let the optional arm break the bench line
            line["legacy_gpu"] = {"unavailable": repr(e)[:200]}
        line["cpu_baseline"] = {"value": rate, "unit": "pairs/s", "cores": threads, "kind": "reference",
                                "sample": "reference OpenMP getWithin (oracle/_ref/libref_fast.so, shipped flags), "
                                          "one frame per call, %d frames x %d repeats of config 2 (%.1f s)"
                                          % (sample, reps, dt)}
    emit(line)


def rdf_extra(lib, _capi, torch, dev, st, peak):
    """RDF O-O kernel on a config-5 shaped box generated on the device (n_side=69 -> 328 509 O,
    rho = 0.0334 A^-3), a few frames per launch; counts checked for conservation only."""
    n_side, nF = 69, 4
    n = n_side ** 3
    L = float(np.float32((n / 0.0334) ** (1.0 / 3.0)))
    xyz = torch.empty((n, nF, 3), dtype=torch.float32, device=dev)
    cell = torch.empty((nF, 3), dtype=torch.float32, device=dev)
    _capi.check(lib.nda_dev_synth_water_box(xyz.data_ptr(), cell.data_ptr(), n_side, nF, 0, nF, L, 0.6, 1005, st))
    counts = torch.zeros(149, dtype=torch.int64, device=dev)
    pairs = 0.5 * n * (n - 1.0) * nF

    def go():
        counts.zero_()
        _capi.check(lib.nda_dev_radial_nbr_counts(xyz.data_ptr(), n, xyz.data_ptr(), n, None, 1, counts.data_ptr(), 149,
                                                  cell.data_ptr(), nF, 0, nF, 1, 0.1, st))
    go()
    torch.cuda.synchronize()
    ks = []
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        go()
        ms = ctypes.c_double(0.0)
        _capi.check(lib.nda_last_kernel_ms(ctypes.byref(ms), None))
        ks.append(ms.value)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / reps
    k_ms = float(np.mean(ks))
    ach = pairs * FLOP_RDF / (k_ms * 1e-3) / 1e12
    return {"workload": "config-5 shape: %d O atoms, %d frames, 149 bins, sameSel (device-generated)" % (n, nF),
            "pairs_per_launch": pairs, "kernel_ms": k_ms, "kernel_pairs_per_s": pairs / (k_ms * 1e-3),
            "step_pairs_per_s": pairs / wall,
            "roofline": {"bound": "fp32", "kernel": "rdf_tc_kernel" if int(lib.nda_last_filter()) == 100 else "rdf_kernel",
                         "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                         "frac": ach / peak if peak else None, "frac_of_nominal": ach / FP32_NOMINAL_TFLOPS,
                         "flop_per_pair": FLOP_RDF,
                         "note": "algorithmic flops of the reference arithmetic over the FP32 peak; above 1 when the "
                                 "filter runs on the tensor cores"},
            "counts_total": int(counts.sum().item())}


def rdf_sharded_extra(lib, _capi, torch, dist, dev, st, rank, world):
    """Frame-sharded RDF with ONE NCCL all-reduce of the u64 histogram: a config-5 shaped trajectory of
    2*world frames, rank r generates and processes frames [2r, 2r+2) (strong scaling of the frame
    axis); device time = max over ranks of (kernel + all-reduce)."""
    from namdanalyzer_b200.distributed import frame_range
    n_side = 69
    n = n_side ** 3
    nF = 2 * world
    fb, fe = frame_range(nF, rank, world)
    L = float(np.float32((n / 0.0334) ** (1.0 / 3.0)))
    xyz = torch.empty((n, nF, 3), dtype=torch.float32, device=dev)
    cell = torch.empty((nF, 3), dtype=torch.float32, device=dev)
    _capi.check(lib.nda_dev_synth_water_box(xyz.data_ptr(), cell.data_ptr(), n_side, nF, fb, fe, L, 0.6, 1005, st))
    counts = torch.zeros(149, dtype=torch.int64, device=dev)

    def go():
        counts.zero_()
        _capi.check(lib.nda_dev_radial_nbr_counts(xyz.data_ptr(), n, xyz.data_ptr(), n, None, 1, counts.data_ptr(), 149,
                                                  cell.data_ptr(), nF, fb, fe, 1, 0.1, st))
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    go()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3
    e0.record()
    for _ in range(reps):
        go()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    pairs = 0.5 * n * (n - 1.0) * nF
    return {"workload": "config-5 shape: %d O atoms, %d frames sharded over %d ranks, one NCCL all-reduce of 149 x u64"
                        % (n, nF, world),
            "ms_per_pass": float(t.item()), "pairs_per_s": pairs / (float(t.item()) * 1e-3),
            "counts_total": int(counts.sum().item())}


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
