#!/usr/bin/env python
"""Summarise `ncu --page source --csv` output: top instructions by stall samples with their dominant stall reasons.
usage: ncu -i X.ncu-rep --page source --csv > src.csv ; python tools/ncu_src.py src.csv [top_n] [window]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
data = []
for r in rows[hdr_i + 1:]:
    if r and r[0] in ("Address", "Kernel Name"):      # the next section / kernel of a multi-kernel export
        break
    if len(r) >= len(hdr):
        data.append(r)
ci = {n: i for i, n in enumerate(hdr)}
samp = ci["# Samples"]
stall_cols = [(n, i) for n, i in ci.items() if n.startswith("stall_") and "Not Issued" not in n]
total = sum(int(r[samp] or 0) for r in data)
print("kernel:", rows[0][1] if rows[0] else "?", " instructions:", len(data), " samples:", total)
agg = {}
for r in data:
    for n, i in stall_cols:
        agg[n] = agg.get(n, 0) + int(r[i] or 0)
print("stall totals:", ", ".join("%s %.1f%%" % (n[6:], 100.0 * v / max(total, 1)) for n, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v))
order = sorted(range(len(data)), key=lambda k: -int(data[k][samp] or 0))[:top_n]
for k in sorted(order):
    r = data[k]
    st = sorted(((int(r[i] or 0), n[6:]) for n, i in stall_cols), reverse=True)[:3]
    print("%5d %6s %5.1f%%  %-60s %s" % (k, r[samp], 100.0 * int(r[samp] or 0) / max(total, 1), r[ci["Source"]].strip()[:60],
                                          " ".join("%s:%d" % (n, v) for v, n in st if v)))
