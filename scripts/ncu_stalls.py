#!/usr/bin/env python3
"""Stall-reason totals of the first kernel in an .ncu-rep (SASS page), and the instructions that collect the most stall samples.
usage: ncu_stalls.py <report.ncu-rep> [top-n]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
hdr, blocks = None, 0
reasons = collections.Counter()
per_ins = []
for r in csv.reader(io.StringIO(src)):
    if r and r[0] == "Kernel Name":
        blocks += 1
        if blocks > 1:
            break
        continue
    if r and r[0] == "Address":
        hdr = r
        cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if hdr and len(r) > 10:
        tot = 0
        for i, h in cols:
            v = int(r[i] or 0)
            reasons[h] += v
            tot += v
        per_ins.append((tot, r[0][-5:], r[1].strip(), {h: int(r[i] or 0) for i, h in cols if int(r[i] or 0)}, int(r[hdr.index("Instructions Executed")])))
total = sum(reasons.values()) or 1
print("stall samples by reason:", ", ".join(f"{k[6:]} {100.0 * v / total:.1f}%" for k, v in reasons.most_common(12)))
print("instructions with the most samples:")
for tot, addr, ins, rs, n in sorted(per_ins, key=lambda t: -t[0])[:topn]:
    top = ", ".join(f"{k[6:]} {v}" for k, v in sorted(rs.items(), key=lambda kv: -kv[1])[:3])
    print(f"  {100.0 * tot / total:5.2f}%  {addr}  {ins[:70]:70s}  x{n}  [{top}]")
