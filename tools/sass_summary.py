#!/usr/bin/env python
"""Counts the Blackwell-specific SASS mnemonics of every kernel in libaz_b200.so (cuobjdump -sass) and writes a table:
UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA (cp.async.bulk.tensor), UTCBAR = tcgen05.commit,
HMMA = mma.sync (legacy warp-level tensor path), plus registers per thread from the ELF section flags.

    python tools/sass_summary.py [out.md]
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "alphazero-chess_b200", "libaz_b200.so")
PATTERNS = collections.OrderedDict([
    ("UTC*MMA (tcgen05.mma)", re.compile(r"\bUTC[A-Z]*MMA")), ("UTCBAR (tcgen05.commit)", re.compile(r"\bUTCBAR")),
    ("LDTM (tcgen05.ld)", re.compile(r"\bLDTM")), ("UTMALDG (TMA load)", re.compile(r"\bUTMALDG")),
    ("SYNCS (mbarrier)", re.compile(r"\bSYNCS")), ("HMMA (mma.sync)", re.compile(r"\bHMMA")),
    ("LDG", re.compile(r"\bLDG")), ("STG", re.compile(r"\bSTG")), ("ATOM/RED", re.compile(r"\b(ATOM|ATOMG|RED)\b")),
    ("SHFL", re.compile(r"\bSHFL")), ("STL/LDL (spills)", re.compile(r"\b(STL|LDL)\b")),
])


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return [o.split("(")[0].replace("azb::", "") for o in out]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts, sizes, order = {}, {}, []
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            sizes[cur] = 0
            order.append(cur)
            continue
        if cur is None or "/*" not in line:
            continue
        body = line.split("*/", 1)[-1]
        if re.search(r"\b[A-Z][A-Z0-9_.]+\b", body):
            sizes[cur] += 1
        for k, pat in PATTERNS.items():
            if pat.search(body):
                counts[cur][k] += 1
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    regs = {}
    fn = None
    for line in res.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            fn = m.group(1)
        m = re.search(r"REG:(\d+)", line)
        if m and fn:
            regs[fn] = int(m.group(1))
    names = demangle(order)
    lines = ["# SASS summary of libaz_b200.so (cuobjdump -sass; sm_100a)", "",
             "Instruction counts per kernel of the mnemonics that prove the Blackwell paths (B200_PROFILING.md): `UTC*MMA` = tcgen05.mma,",
             "`UTCBAR` = tcgen05.commit, `LDTM` = tcgen05.ld, `UTMALDG` = TMA tensor load, `SYNCS` = mbarrier, `HMMA` = mma.sync.", "",
             "| kernel | SASS lines | regs | " + " | ".join(PATTERNS) + " |", "|---|---:|---:|" + "---:|" * len(PATTERNS)]
    tot = collections.Counter()
    for f, n in zip(order, names):
        c = counts[f]
        tot.update(c)
        lines.append(f"| `{n[:64]}` | {sizes[f]} | {regs.get(f, '')} | " + " | ".join(str(c[k]) if c[k] else "" for k in PATTERNS) + " |")
    lines.append(f"| **total ({len(order)} kernels)** | {sum(sizes.values())} | | " + " | ".join(str(tot[k]) for k in PATTERNS) + " |")
    text = "\n".join(lines) + "\n"
    if len(sys.argv) > 1:
        with open(sys.argv[1], "w") as f:
            f.write(text)
    else:
        print(text)


if __name__ == "__main__":
    main()
