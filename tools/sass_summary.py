#!/usr/bin/env python3
"""Opcode summary of every kernel in libygz_b200.so from `cuobjdump -sass` (no GPU needed): instruction count and the
mnemonics that prove what a kernel is built from (TMA, mbarrier, cluster barriers, distributed shared memory, FP64, POPC,
shuffles, tensor-core opcodes).  usage: tools/sass_summary.py [lib] > profiles/r2_sass_opcodes.txt"""
import collections
import re
import subprocess
import sys
from pathlib import Path

lib = sys.argv[1] if len(sys.argv) > 1 else str(Path(__file__).resolve().parent.parent / "ygz_slam_b200" / "libygz_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
WATCH = ["UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "UCGABAR", "MAPA", "LD.E", "ST.E", "LDS", "STS", "DFMA", "DMUL", "DADD", "MUFU.RCP64H", "MUFU.RSQ64H",
         "MUFU", "FFMA", "IMAD", "POPC", "SHFL", "CREDUX", "VABSDIFF4", "VIMNMX", "BAR.SYNC", "LDG", "STG", "LDL", "STL",
         "ATOM", "RED", "HMMA", "DMMA", "UTCHMMA", "UTCQMMA", "LDTM", "STTM"]
kern, counts, total = None, None, None
rows = []
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        if kern:
            rows.append((kern, total, counts))
        kern, counts, total = m.group(1), collections.Counter(), 0
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and kern:
        op = m.group(1)
        total += 1
        for w in WATCH:
            if op.startswith(w):
                counts[w] += 1
if kern:
    rows.append((kern, total, counts))
dem = subprocess.run(["c++filt"] + [r[0] for r in rows], capture_output=True, text=True).stdout.splitlines()
print(f"# {lib}: SASS opcode summary (sm_100a), {len(rows)} kernels / device functions")
print("# tensor-core opcodes (HMMA / DMMA / UTC*MMA / LDTM / STTM) are listed when present: none of these integer / FP64 latency-bound kernels uses them")
for (k, tot, c), name in sorted(zip(rows, dem), key=lambda t: -t[0][1]):
    short = re.sub(r"\(.*", "", name.replace("(anonymous namespace)::", "")).split("::")[-1]
    print(f"{short:34s} {tot:6d} instr  " + "  ".join(f"{w}={c[w]}" for w in WATCH if c[w]))
