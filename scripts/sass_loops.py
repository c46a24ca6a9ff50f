#!/usr/bin/env python3
"""Static look at the streaming kernel's loops (no GPU needed): for a kernel variant, the SASS instruction count and opcode histogram
of the uniform-motion loop (four pixel rows of two columns per lane) and of the general loop (two rows).
usage: sass_loops.py <object or .so> <REJ><ALPHA><DIAG> (e.g. 000)"""
import collections
import re
import subprocess
import sys


def loops(path, variant, kernel="taa_resolve_stream_kernel"):
    txt = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    want = kernel + "ILb%sELb%sELb%sE" % tuple(variant)
    for f in re.split(r"\n\s*Function : ", txt)[1:]:
        if want not in f.split("\n")[0]:
            continue
        ins = [(int(a, 16), t) for a, t in re.findall(r"/\*([0-9a-f]{4,})\*/\s+(.*?);", f)]
        idx = {a: k for k, (a, _) in enumerate(ins)}
        # the out-of-line stubs of divergent collectives (BRA.DIV targets) sit behind the last EXIT and branch back: not loops
        last_exit = max(k for k, (_, t) in enumerate(ins) if "EXIT" in t)
        cand = []
        for i, (a, t) in enumerate(ins[:last_exit]):
            m = re.search(r"\bBRA\b.*?(0x[0-9a-f]+)", t)
            if m and int(m.group(1), 16) < a and int(m.group(1), 16) in idx:
                j = idx[int(m.group(1), 16)]
                if i - j + 1 >= 200:
                    cand.append((j, i))
        # keep innermost loops only
        inner = [c for c in cand if not any(o != c and c[0] <= o[0] and o[1] <= c[1] for o in cand)]
        res = []
        for j, i in inner:
            c = collections.Counter()
            for _, t in ins[j:i + 1]:
                parts = t.split()
                op = parts[1] if parts[0].startswith("@") else parts[0]
                c[op.split(".")[0]] += 1
            res.append((i - j + 1, c, ins[j:i + 1]))
        return len(ins), res
    return 0, []


if __name__ == "__main__":
    total, res = loops(sys.argv[1], sys.argv[2])
    print("instructions in the kernel:", total)
    for n, c, body in res:
        print(f"loop of {n} instructions:", ", ".join(f"{k} {v}" for k, v in sorted(c.items(), key=lambda kv: -kv[1])))
        if len(sys.argv) > 3:
            open(sys.argv[3] + f".{n}.sass", "w").write("\n".join(f"{a:05x} {t}" for a, t in body))
