#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into JSON + markdown for profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_prof
"""
import collections
import csv
import io
import json
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__occupancy_limit_registers", "occ limit regs (CTAs)"),
    ("launch__occupancy_limit_shared_mem", "occ limit smem (CTAs)"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe (inst) %"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe (cycles) %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
    ("smsp__cycles_active.avg", "SMSP active cycles"),
    ("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor pipe active %"),
    ("sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active", "tcgen05 issue pipe %"),
    ("sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active", "TMA issue pipe %"),
    ("lts__t_bytes.sum", "L2 bytes"),
]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    return hdr, units, rows[2:]


def stalls(rep, kernel_regex):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kernel_regex,
                          "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    if len(rows) < 3:
        return {}
    hdr = rows[1]
    cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot = collections.Counter()
    seen = set()
    for r in rows[2:]:
        if len(r) < len(hdr) or r[0] in seen:
            continue
        seen.add(r[0])
        for i, h in cols:
            try:
                tot[h] += int(r[i])
            except ValueError:
                pass
    s = sum(tot.values()) or 1
    return {k: round(100.0 * v / s, 1) for k, v in tot.most_common(8)}


def main():
    rep, outbase = sys.argv[1], sys.argv[2]
    hdr, units, rows = raw(rep)
    idx = {h: i for i, h in enumerate(hdr)}
    kernels = collections.OrderedDict()
    for r in rows:
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "")
        kernels[name] = r                   # keep the last (warm) launch of each kernel
    res = {}
    md = ["| metric | " + " | ".join(kernels) + " |", "|---|" + "---|" * len(kernels)]
    for key, label in METRICS:
        if key not in idx:
            continue
        vals = []
        for name, r in kernels.items():
            v = r[idx[key]]
            res.setdefault(name, {})[key] = {"value": v, "unit": units[idx[key]]}
            vals.append("%s %s" % (v, units[idx[key]]))
        md.append("| %s (`%s`) | %s |" % (label, key, " | ".join(vals)))
    for name in kernels:
        st = stalls(rep, name.split("<")[0])
        res[name]["warp_stall_sampling_pct"] = st
        md.append("")
        md.append("Warp-stall sampling, %s: " % name + ", ".join("%s %.1f%%" % kv for kv in st.items()))
    json.dump(res, open(outbase + ".json", "w"), indent=1)
    open(outbase + ".md", "w").write("\n".join(md) + "\n")
    print("\n".join(md))


if __name__ == "__main__":
    main()
