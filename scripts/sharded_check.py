#!/usr/bin/env python
"""torchrun, one rank per GPU: the row-band sharded resolve (taa_star_b200/sharded.py; peer stores and NCCL halo exchange) against the whole-frame
resolve of the same sequence computed redundantly on every rank. Bit-identical history and result rows are required, frame after frame.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 scripts/sharded_check.py"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from taa_star_b200 import abi, configs, host  # noqa: E402
from taa_star_b200.sharded import ShardedTaa  # noqa: E402
from taa_star_b200.synth import SyntheticScene  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    W, H, halo = 1920, 1080, 20
    ok = True
    for cfg_id, replicate, exchange in ((2, False, "peer"), (2, False, "nccl"), (3, False, "nccl"), (2, True, "nccl")):
        p = configs.config2_resolve() if cfg_id == 2 else configs.config3_full_chain()
        sh = ShardedTaa(W, H, halo=halo, device=dev, apron=halo if cfg_id == 3 else 2, replicate=replicate, exchange=exchange)
        L = sh.L
        sc = SyntheticScene(W, H, device=dev, with_aux=False, pan_px=(3.25, 7.5), mover_px=(-6.5, 5.25))  # (not whole texels: see _whole_frame_reference in sharded.py)
        whole = host.TaaContext((W, H))
        hist = [torch.zeros(H, W, 4, dtype=torch.float16, device=dev) for _ in range(2)]
        res = torch.zeros(H, W, 4, dtype=torch.float16, device=dev)
        prev_depth = None
        for n in range(12):
            f = sc.frame(n)
            u = configs.uniforms_for(p, f.jitter_ndc, reset_history=(n == 0))
            hd = prev_depth if prev_depth is not None else f.depth
            whole.resolve(u, color=f.color, depth=f.depth, velocity=f.velocity, history_in=hist[n & 1], history_out=hist[1 - (n & 1)], result=res,
                          history_depth=hd if cfg_id == 3 else None)
            a0 = L.iy0 if cfg_id == 2 else max(0, L.y0 - halo)
            a1 = L.iy1 if cfg_id == 2 else min(H, L.y1 + halo)
            sh.step(u, f.color[a0:a1].contiguous(), f.depth[a0:a1].contiguous(), f.velocity[a0:a1].contiguous(), a0,
                    history_depth=hd[a0:a1].contiguous() if cfg_id == 3 else None)
            if exchange == "peer" and n % 3 != 2:
                prev_depth = f.depth
                continue  # (no host synchronisation between the ranks' frames: the flags alone order the halo stores and reads)
            torch.cuda.synchronize()
            dist.barrier()
            assert sh.poll() == abi.TAA_OK, "halo overflow / peer time-out"
            got_hist = sh.hist[sh.parity]  # the buffer just written (parity already flipped)
            ref_hist = hist[1 - (n & 1)]
            lo, hi = (0, H) if replicate else (L.hy0, L.hy1)
            same_h = torch.equal(got_hist.view(torch.int16)[(lo - sh.hist_y0):(hi - sh.hist_y0)], ref_hist.view(torch.int16)[lo:hi])
            same_r = torch.equal(sh.result.view(torch.int16), res.view(torch.int16)[L.y0:L.y1])
            if not (same_h and same_r):
                ok = False
                print(f"rank {rank} cfg {cfg_id} replicate {replicate} exchange {exchange} frame {n}: history rows equal {same_h}, result rows equal {same_r}", flush=True)
                break
            prev_depth = f.depth
        sh.close()
        del sh, whole
    flag = torch.tensor([0 if ok else 1], device=dev)
    dist.all_reduce(flag)
    if rank == 0:
        print("sharded_check:", "OK" if flag.item() == 0 else "FAILED", f"({world} ranks, {W}x{H}, halo {halo}, config 2 with peer stores, configs 2/3 over NCCL, replicated history; 12 frames each)", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 0 else 1)


if __name__ == "__main__":
    main()
