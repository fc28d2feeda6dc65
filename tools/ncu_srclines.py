#!/usr/bin/env python3
"""Per-source-line instruction and stall-sample shares from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`.
usage: tools/ncu_srclines.py <csv> [top N]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
h = rows[hi]
iL, iS, iI, iSm, iT = 0, 1, h.index("Instructions Executed"), h.index("# Samples"), h.index("Thread Instructions Executed")
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= iT or not r[iL].strip().isdigit(): continue
    k = int(r[iL])
    a = agg.setdefault(k, [r[iS], 0, 0, 0])
    try: a[1] += int(r[iI] or 0); a[2] += int(r[iSm] or 0); a[3] += int(r[iT] or 0)
    except ValueError: pass
ti = sum(a[1] for a in agg.values()) or 1; ts = sum(a[2] for a in agg.values()) or 1
print(f"total warp instructions {ti}, samples {ts}")
print("| line | inst % | samples % | thr/inst | source |\n|---:|---:|---:|---:|---|")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][2])[:top]:
    print(f"| {k} | {100*a[1]/ti:.1f} | {100*a[2]/ts:.1f} | {a[3]/max(1,a[1]):.1f} | `{a[0].strip()[:110]}` |")
