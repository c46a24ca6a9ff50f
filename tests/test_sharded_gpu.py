"""Row-band sharding on real GPUs: every rank's band (and the halo rows its neighbours stored into it) equals the whole-frame resolve bit for
bit, frame after frame — config 2 with peer stores from the kernel (taa_band_peers), configs 2 / 3 over the NCCL halo exchange, and the
replicated-history variant. Spawns one process per GPU (torchrun); skipped where fewer GPUs are visible."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_equals_whole_frame(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"{world} GPUs needed, {torch.cuda.device_count()} visible")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29540 + world), os.path.join(ROOT, "scripts", "sharded_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    tail = (r.stdout + r.stderr)[-3000:]
    assert r.returncode == 0 and "sharded_check: OK" in r.stdout, tail
