#!/usr/bin/env python3
"""Per-source-line executed warp instructions of one kernel in an .ncu-rep (needs -lineinfo and --import-source on).
usage: ncu_lines.py <report.ncu-rep> [kernel-substring] [pixels_per_launch] [top-n]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else ""
px = int(sys.argv[3]) if len(sys.argv) > 3 else 3840 * 2160
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 60
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
cur_file, cur_fn, hdr, active, done = None, None, None, False, False
lines = {}
stalls = {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        cur_fn = r[1]
        active = (want in cur_fn)
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if not active or hdr is None or len(r) < 8:
        continue
    if r[0] != "":  # a source line row (aggregated over its SASS)
        try:
            n = int(r[hdr.index("Instructions Executed")])
            s = int(r[hdr.index("# Samples")])
        except ValueError:
            continue
        key = (cur_file, int(r[0]), r[1].strip()[:110])
        lines[key] = lines.get(key, 0) + n
        stalls[key] = stalls.get(key, 0) + s
tot = sum(lines.values())
tot_s = sum(stalls.values()) or 1
print(f"kernel filter '{want}': {tot / (px / 32):.1f} warp instructions per pixel-thread in total")
order = (lambda kv: -stalls[kv[0]]) if len(sys.argv) > 5 and sys.argv[5] == "stalls" else (lambda kv: -kv[1])
for key, n in sorted(lines.items(), key=order)[:topn]:
    print(f"{n / (px / 32):7.1f}  {100.0 * stalls[key] / tot_s:5.1f}%  {key[0]}:{key[1]:<4d} {key[2]}")
