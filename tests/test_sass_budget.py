"""CPU: the static instruction budget of the strip kernel's per-pixel loops (scripts/sass_budget.py). The kernel is issue-bound, so the
length of the uniform-motion strip loop IS the performance model; this guards it against silent regressions."""
import os
import shutil
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))


@pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not available")
def test_uniform_motion_strip_loop_stays_within_its_instruction_budget(taalib):
    import sass_budget
    res = {sass_budget.variant(n): r for n, r in sass_budget.strip_loops().items() if sass_budget.variant(n)}
    # (REJ, ALPHA, DIAG, CTAs/SM, unroll): the default variants; the first loop after the last barrier is the uniform-motion loop
    budget = {(0, 0, 0, 3, 1): 200, (0, 0, 1, 3, 1): 230, (1, 1, 1, 2, 1): 420}
    for v, limit in budget.items():
        assert v in res, f"variant {v} is not in the library"
        loops = res[v]["loops_after_last_barrier"]
        assert loops and loops[0] <= limit, f"variant {v}: uniform-motion loop has {loops[0] if loops else None} SASS instructions (budget {limit})"
        assert loops[0] >= 100, f"variant {v}: loop detection is off ({loops})"


@pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not available")
def test_default_streaming_kernel_uses_tensor_map_tma_and_stays_within_its_budget(taalib):
    """The DEFAULT resolve kernel (taa_resolve_stream_kernel<REJ=0,ALPHA=0,DIAG=0,FX=0,MINB=6,EPI=0,PEER=0>) stages its rows with 2-D tensor-map TMA
    (SASS UTMALDG.2D, completion on mbarriers: SYNCS.*), has no block-wide barrier, does not spill, and its uniform-motion loop (four pixel rows
    of two columns per lane) stays within its instruction budget (it is issue-bound: the loop length is the performance model)."""
    import re
    import subprocess
    lib = os.path.join(ROOT, "taa_star_b200", "libtaa_b200.so")
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    want = "taa_resolve_stream_kernelILb0ELb0ELb0ELi0ELi6ELi0ELb0E"
    body = [f for f in re.split(r"\n\s*Function : ", txt)[1:] if want in f.split("\n")[0]]
    assert len(body) == 1, f"{len(body)} functions match {want}"
    ins = [(int(a, 16), t) for a, t in re.findall(r"/\*([0-9a-f]{4,})\*/\s+(.*?);", body[0])]
    ops = [t.split()[1] if t.startswith("@") else t.split()[0] for _, t in ins]
    assert sum(o.startswith("UTMALDG.2D") for o in ops) >= 4, "no tensor-map TMA loads in the default kernel"
    assert any(o.startswith("SYNCS.PHASECHK") for o in ops) and any(o.startswith("SYNCS.ARRIVE") for o in ops), "no mbarrier wait / arrive"
    assert not any(o.startswith("BAR.") for o in ops), "a block-wide barrier in the streaming kernel"
    assert not any(o.startswith(("STL", "LDL")) for o in ops), "the default variant spills"
    idx = {a: k for k, (a, _) in enumerate(ins)}
    loops = []
    for i, (a, t) in enumerate(ins):
        m = re.search(r"\bBRA\b.*?(0x[0-9a-f]+)", t)
        if m and int(m.group(1), 16) < a and int(m.group(1), 16) in idx and i - idx[int(m.group(1), 16)] + 1 >= 400:
            loops.append(i - idx[int(m.group(1), 16)] + 1)
    assert loops, "loop detection is off"
    # the first long loop is the uniform-motion loop: 8 pixels per lane and turn; 1558 instructions at the end of round 2 (194.75 per pixel)
    assert loops[0] <= 1650, f"uniform-motion loop of the streaming kernel: {loops[0]} SASS instructions per 8 pixels (budget 1650)"
