#!/usr/bin/env python
"""Per-instruction view of an .ncu-rep source page: hottest instructions and (optionally) the instructions
executed by few warps (a control / producer warp).

    python tools/ncu_src.py rep.ncu-rep [kernel-regex] [--rare N]   # --rare: only instructions executed < N times
"""
import csv, io, subprocess, sys
rep = sys.argv[1]
kre = sys.argv[2] if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else None
rare = int(sys.argv[sys.argv.index("--rare") + 1]) if "--rare" in sys.argv else None
cmd = ["ncu", "-i", rep, "--page", "source", "--csv"]
if kre:
    cmd += ["--kernel-name", "regex:" + kre, "--launch-count", "1"]
out = subprocess.run(cmd, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    try:
        st = {h: int(r[ix[h]]) for h in stall_cols}
        data.append((int(r[ix["Address"]], 16), int(r[ix["# Samples"]]), int(r[ix["Instructions Executed"]]), r[ix["Source"]].strip(), st))
    except ValueError:
        pass
tot = sum(d[1] for d in data) or 1
base = data[0][0]
print("total samples", tot, "instructions", len(data))
if rare is None:
    for a, s, n, src, st in sorted(data, key=lambda d: -d[1])[:45]:
        top = max(st.items(), key=lambda kv: kv[1])
        print("%5.1f%%  +%05x  x%-9d %-18s %s" % (100.0 * s / tot, a - base, n, top[0] if top[1] else "", src[:100]))
else:
    for a, s, n, src, st in data:
        if n < rare and n > 0:
            top = max(st.items(), key=lambda kv: kv[1])
            print("%5d  +%05x  x%-8d %-18s %s" % (s, a - base, n, top[0] if top[1] else "", src[:110]))
