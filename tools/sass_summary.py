#!/usr/bin/env python
"""Per-kernel counts of the Blackwell-specific SASS mnemonics in libnda_b200.so -> profiles/sass_summary.md.

    python tools/sass_summary.py            (needs cuobjdump and c++filt; no GPU)
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "namdanalyzer_b200", "lib", "libnda_b200.so")
KEYS = ["UTCHMMA", "LDTM", "UTCBAR", "UBLKCP", "SYNCS", "FFMA2", "FADD2", "FMUL2", "HMNMX2", "VHMNMX", "LDGSTS",
        "REDG", "ATOMS", "ATOMG", "DADD", "MUFU.RSQ"]


def demangle(n):
    try:
        return re.sub(r"\(.*", "", subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip())
    except OSError:
        return n


def main():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    rows, tot = [], collections.Counter()
    for f in re.split(r"\n\s*Function : ", txt)[1:]:
        name = f.split("\n", 1)[0].strip()
        c = {k: len(re.findall(r"\b" + re.escape(k), f)) for k in KEYS}
        n = len(re.findall(r"^\s+/\*[0-9a-f]{4}\*/", f, flags=re.M))
        rows.append((name, n, c))
        for k, v in c.items():
            tot[k] += v
    out = ["# SASS evidence, libnda_b200.so", "",
           "`cuobjdump -sass namdanalyzer_b200/lib/libnda_b200.so`, counted per kernel by `tools/sass_summary.py` (this file is its output;",
           "kernels of fewer than 40 instructions are left out).  Blackwell-only mnemonics: `UTCHMMA` = tcgen05.mma (kind::f16), `LDTM` =",
           "tcgen05.ld, `UTCBAR` = tcgen05.commit, `UBLKCP` = cp.async.bulk (TMA), `SYNCS` = mbarrier operations, `FFMA2 / FADD2 / FMUL2` =",
           "packed f32x2, `HMNMX2 / VHMNMX` = packed fp16 maximum, `LDGSTS` = cp.async.", "",
           "| kernel | instructions | " + " | ".join(KEYS) + " |", "|---|---|" + "---|" * len(KEYS)]
    for name, n, c in sorted(rows, key=lambda r: -r[1]):
        if n >= 40:
            out.append("| `%s` | %d | %s |" % (demangle(name)[:72], n, " | ".join(str(c[k]) if c[k] else "" for k in KEYS)))
    out.append("| **whole library** | | " + " | ".join("**%d**" % tot[k] for k in KEYS) + " |")
    open(os.path.join(ROOT, "profiles", "sass_summary.md"), "w").write("\n".join(out) + "\n")
    print("\n".join(out[-3:]))


if __name__ == "__main__":
    main()
