#!/usr/bin/env python3
"""Summarises an .ncu-rep (read here, no GPU needed) into a small text file for profiles/.
usage: ncu_summary.py <report.ncu-rep> <out.txt> [pixels_per_launch]"""
import collections
import csv
import io
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
px = int(sys.argv[3]) if len(sys.argv) > 3 else 3840 * 2160
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__grid_size", "launch__block_size",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__cycles_elapsed.avg", "sm__cycles_active.avg"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
lines = [f"source report: {rep}", f"pixels per launch assumed: {px}", ""]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    lines.append(f"== {name}  grid {r[hdr.index('Grid Size')] if 'Grid Size' in hdr else ''} block {r[hdr.index('Block Size')] if 'Block Size' in hdr else ''}")
    vals = {}
    for k in KEYS:
        if k in hdr:
            vals[k] = r[hdr.index(k)]
            lines.append(f"  {k:70s} {r[hdr.index(k)]} {units[hdr.index(k)]}")
    try:
        inst = float(vals["smsp__inst_executed.sum"])
        lines.append(f"  -> warp instructions per pixel-thread: {inst / (px / 32):.1f}")
        rd = float(vals["dram__bytes_read.sum"]); wr = float(vals["dram__bytes_write.sum"])
        ur, uw = units[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_write.sum")]
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        tot = rd * scale[ur] + wr * scale[uw]
        lines.append(f"  -> DRAM traffic per launch: {tot / 1e6:.1f} MB = {tot / px:.1f} B/px")
    except Exception as e:
        lines.append(f"  (derived values unavailable: {e})")
    lines.append("")
# executed-instruction mix of the first kernel (SASS page)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
agg, tot, hdr2, blocks = collections.Counter(), 0, None, 0
for r in csv.reader(io.StringIO(src)):
    if r and r[0] == "Kernel Name":
        blocks += 1
        if blocks > 1:
            break
        continue
    if r and r[0] == "Address":
        hdr2 = r
        continue
    if hdr2 and len(r) > 5:
        toks = r[1].split()
        o = toks[1] if toks[0].startswith("@") else toks[0]
        n = int(r[hdr2.index("Instructions Executed")])
        agg[o.split(".")[0]] += n
        tot += n
if tot:
    lines.append(f"executed SASS mix of the first launch (warp instructions per pixel-thread, total {tot / (px / 32):.1f}):")
    lines.append("  " + "  ".join(f"{k} {v / (px / 32):.1f}" for k, v in agg.most_common(28)))
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
