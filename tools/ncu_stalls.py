#!/usr/bin/env python
"""DEVELOPMENT: per-instruction warp-stall samples of the first kernel of an ncu report (needs --import-source on).
    python tools/ncu_stalls.py report.ncu-rep [top N]
Prints the instructions with the most samples, their dominant stall reasons, and totals per reason."""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]
ia, isrc, iall, iexe = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
reasons = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = collections.Counter()
items = []
base = None
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    addr = int(r[ia], 16)
    base = addr if base is None else base
    n = int(r[iall] or 0)
    per = {h: int(r[i] or 0) for i, h in reasons if int(r[i] or 0)}
    for h, v in per.items():
        tot[h] += v
    items.append((addr - base, r[isrc].strip(), n, int(r[iexe] or 0), per))
total = sum(n for _, _, n, _, _ in items)
print("total samples", total)
for h, v in tot.most_common():
    print("  %-22s %7d  %.1f%%" % (h, v, 100.0 * v / total))
print()
for off, src, n, exe, per in sorted(items, key=lambda t: -t[2])[:top]:
    why = " ".join("%s:%d" % (h[6:], v) for h, v in sorted(per.items(), key=lambda kv: -kv[1])[:3])
    print("%05x %6d exe=%-7d %-70s %s" % (off, n, exe, src[:70], why))
