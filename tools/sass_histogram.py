#!/usr/bin/env python3
"""SASS opcode histogram of every kernel in libqtet_b200.so (or the objects given): proves which pipes the code is written for
(DFMA / DMMA / LDGSTS ... and that there is no UTCMMA / TMA: tcgen05 has no FP64 kind).  Markdown on stdout.

    python tools/sass_histogram.py [lib_or_object ...] > profiles/r2_sass_histogram.md
"""
import os, re, subprocess, sys
from collections import Counter

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
targets = sys.argv[1:] or [os.path.join(ROOT, "quantum-targeted-energy-transfer_b200", "lib", "libqtet_b200.so")]
print("# SASS opcode histogram (sm_100a cubins)\n")
for t in targets:
    out = subprocess.run(["cuobjdump", "-sass", t], capture_output=True, text=True).stdout
    kern, hist = None, {}
    for line in out.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            kern = re.sub(r"qtet::\(anonymous namespace\)::|\(anonymous namespace\)::|qtet::splitk::|qtet::", "", kern)
            hist[kern] = Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and kern:
            hist[kern][m.group(1)] += 1
    print(f"## {os.path.basename(t)}\n")
    print("| kernel | instr | DFMA | DMUL | DADD | DMMA | DSETP | MUFU | LDS | STS | LDG | STG | LDGSTS | SHFL | BAR | UTCMMA/UTMA |")
    print("|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|")
    for k, c in sorted(hist.items(), key=lambda x: -sum(x[1].values())):
        tot = sum(c.values())
        if tot < 50:
            continue
        tc = sum(v for op, v in c.items() if op.startswith("UTC") or op.startswith("UTMA") or op.startswith("UBLKCP"))
        print(f"| `{k[:90]}` | {tot} | {c['DFMA']} | {c['DMUL']} | {c['DADD']} | {c['DMMA']} | {c['DSETP']} | {c['MUFU']} | {c['LDS']} | {c['STS']} | "
              f"{c['LDG']} | {c['STG']} | {c['LDGSTS']} | {c['SHFL']} | {c['BAR']} | {tc} |")
    print()
