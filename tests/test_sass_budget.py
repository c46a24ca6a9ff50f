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
