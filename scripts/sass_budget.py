#!/usr/bin/env python3
"""Static instruction budget of the strip kernel (no GPU needed): disassembles libtaa_b200.so and reports, per kernel variant, the
number of SASS instructions in each loop that follows the last block-wide barrier (= the strip loops of phase 2; the first one is the
uniform-motion loop, executed once per pixel). The kernel is issue-bound: this is the number to watch.
usage: sass_budget.py [path/to/libtaa_b200.so] [substring of the mangled kernel name]"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def strip_loops(lib=None, want="taa_resolve_strip_kernel"):
    lib = lib or os.path.join(ROOT, "taa_star_b200", "libtaa_b200.so")
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    out, name, ins = {}, None, []

    def flush():
        if name and want in name and ins:
            last_bar = max((i for i, (_, t) in enumerate(ins) if "BAR." in t), default=-1)
            loops = []
            for i, (addr, t) in enumerate(ins):
                m = re.search(r"\bBRA\b.*?(0x[0-9a-f]+)", t)
                if m and i > last_bar:
                    tgt = int(m.group(1), 16)
                    if tgt < addr:  # backward branch: a loop of (addr - tgt) / 16 + 1 instructions
                        j = next((k for k, (a, _) in enumerate(ins) if a == tgt), None)
                        if j is not None and j > last_bar:
                            loops.append(i - j + 1)
            out[name] = {"instructions": len(ins), "loops_after_last_barrier": loops}

    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            flush()
            name, ins = m.group(1), []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m and name:
            ins.append((int(m.group(1), 16), m.group(2)))
    flush()
    return out


def variant(name):
    m = re.search(r"strip_kernelILb(\d)ELb(\d)ELb(\d)ELi(\d)ELi(\d)E", name)
    return tuple(int(x) for x in m.groups()) if m else None


if __name__ == "__main__":
    res = strip_loops(sys.argv[1] if len(sys.argv) > 1 else None, sys.argv[2] if len(sys.argv) > 2 else "taa_resolve_strip_kernel")
    print("REJ ALPHA DIAG MINB UNR : SASS instructions, loops after the last barrier (instructions each)")
    for n, r in sorted(res.items(), key=lambda kv: variant(kv[0]) or ()):
        v = variant(n)
        if v:
            print(f"{v[0]:3d} {v[1]:5d} {v[2]:4d} {v[3]:4d} {v[4]:3d} : {r['instructions']:5d}  {r['loops_after_last_barrier']}")
