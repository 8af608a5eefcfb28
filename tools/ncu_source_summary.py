#!/usr/bin/env python
"""Aggregates `ncu -i X.ncu-rep --page source --csv` per SASS opcode: instructions, shared-memory wavefronts,
stall samples. Usage: ncu -i rep --page source --csv | python tools/ncu_source_summary.py [--top N]"""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(sys.stdin))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
col = {n: i for i, n in enumerate(hdr)}
agg = defaultdict(lambda: [0, 0, 0, 0, 0])
stall_cols = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
stalls = defaultdict(int)
tot_samples = 0
lines = []
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    src = r[col["Source"]].strip()
    op = src.split()[0] if src else "?"
    if op.startswith("@"):
        op = src.split()[1]
    op = op.rstrip(";")
    def f(n):
        try:
            return int(float(r[col[n]] or 0)) if n in col else 0
        except ValueError:      # a second header row / a text cell ("-")
            return 0
    a = agg[op]
    a[0] += f("Instructions Executed")
    a[1] += f("L1 Wavefronts Shared")
    a[2] += f("L1 Wavefronts Shared Ideal")
    a[3] += f("# Samples")
    a[4] += f("L2 Theoretical Sectors Global") + f("L2 Theoretical Sectors Local")
    tot_samples += f("# Samples")
    for s in stall_cols:
        stalls[s] += f(s)
    lines.append((f("# Samples"), r[col["Address"]], src, {s: f(s) for s in stall_cols if f(s)}))
print(f"{'opcode':28s} {'inst':>12s} {'smem wavefr':>12s} {'ideal':>12s} {'samples':>8s} {'L2 sectors':>12s}")
for op, a in sorted(agg.items(), key=lambda kv: -kv[1][1] - kv[1][0] / 1000)[:40]:
    print(f"{op:28s} {a[0]:12d} {a[1]:12d} {a[2]:12d} {a[3]:8d} {a[4]:12d}")
print("total samples", tot_samples)
for s, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:10]:
    print(f"  {s:28s} {v:8d} {100.0 * v / max(tot_samples, 1):5.1f}%")
top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 0
for n, addr, src, st in sorted(lines, key=lambda x: -x[0])[:top]:
    print(n, addr, src[:70], st)
