"""CPU, world_size 2 and 3 over gloo: the band layout, the halo exchange and the replicated-history gather of
taa_star_b200/sharded.py. The per-band compute is done by the oracle here (no GPU), which also proves the halo sizing rule:
a band with `halo` rows of history (and of the previous depth buffer, which taa.comp:818 reads at the HISTORY position) and a
2-row apron of the current inputs reproduces the whole-frame result exactly."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, replicate, q, geom=(96, 66, 9, 6, (2.0, 5.5), (-4.0, 3.0))):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle_py
        from taa_star_b200 import configs
        from taa_star_b200.sharded import BandLayout, HaloExchanger, band_of, gather_full_history
        from taa_star_b200.synth import SyntheticScene
        W, H, halo, nframes, pan, mover = geom
        L = BandLayout(H, world, rank, halo)
        assert (L.y0, L.y1) == band_of(H, world, rank)
        sc = SyntheticScene(W, H, pan_px=pan, mover_px=mover, with_aux=False)
        p = configs.config3_full_chain()
        xchg = HaloExchanger(L)
        # every rank also runs the whole frame (the reference result)
        full_hist = np.zeros((H, W, 4), np.float16)
        # band-local state: full-size arrays poisoned outside the rows this rank may touch
        band_hist = torch.full((H, W, 4), float("nan"), dtype=torch.float16)
        band_hist[L.hy0:L.hy1] = 0
        prev_depth = None
        for n in range(nframes):
            f = sc.frame(n)
            u = configs.uniforms_for(p, f.jitter_ndc, reset_history=(n == 0))
            col, dep, vel = f.color.numpy(), f.depth.numpy(), f.velocity.numpy()
            hd = prev_depth if prev_depth is not None else dep
            ref = oracle_py.resolve(u, col, dep, vel, full_hist, history_depth=hd, want=("history_out", "result", "mask"))
            # band: inputs poisoned outside the apron rows
            def poison(a, lo, hi):
                b = a.copy()
                b[:lo] = np.nan if a.dtype != np.uint32 else 0
                b[hi:] = np.nan if a.dtype != np.uint32 else 0
                return b
            got = oracle_py.resolve(u, poison(col, L.iy0, L.iy1), poison(dep, L.iy0, L.iy1), poison(vel, L.iy0, L.iy1), band_hist.numpy(),
                                    history_depth=poison(hd, L.hy0, L.hy1), want=("history_out", "result", "mask"), rows=(L.y0, L.y1))
            for k in ("history_out", "result"):
                a, b = ref[k][L.y0:L.y1].view(np.uint16), got[k][L.y0:L.y1].view(np.uint16)
                assert (a == b).all(), f"rank {rank} frame {n}: {k} differs from the whole-frame result"
            assert (ref["mask"][L.y0:L.y1] == got["mask"][L.y0:L.y1]).all()
            new_hist = torch.full((H, W, 4), float("nan"), dtype=torch.float16)
            new_hist[L.y0:L.y1] = torch.from_numpy(got["history_out"][L.y0:L.y1])
            if replicate:
                full = torch.zeros((H, W, 4), dtype=torch.float16)
                gather_full_history(new_hist[L.y0:L.y1], full, L)
                assert (full.numpy().view(np.uint16) == ref["history_out"].view(np.uint16)).all()
                new_hist = full
            else:
                view = new_hist[L.hy0:L.hy1]
                for w in xchg.exchange(view):
                    w.wait()
                a = new_hist[L.hy0:L.hy1].numpy().view(np.uint16)
                b = ref["history_out"][L.hy0:L.hy1].view(np.uint16)
                assert (a == b).all(), f"rank {rank} frame {n}: halo rows differ after the exchange"
            band_hist = new_hist
            full_hist = ref["history_out"]
            prev_depth = dep
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, "FAIL: " + traceback.format_exc()))
        raise
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,replicate", [(2, False), (3, False), (2, True)])
def test_band_sharding_over_gloo(world, replicate):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500) + world * 7 + (3 if replicate else 0)
    procs = [ctx.Process(target=_worker, args=(r, world, port, replicate, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(msg == "ok" for _, msg in results), results


def test_band_sharding_at_16k_width_over_gloo():
    """BASELINE configs[3] names 15360x8640: the synthetic motion scales with the frame (the same angular motion covers twice the pixels of
    8K), and so does the halo the sharded bench uses (sharded.bench_main: halo * W // 7680 = 40 rows). A strip of that width, two bands, a
    vertical motion of 30 px per frame: the band with 40 halo rows reproduces the whole frame, and the exchange moves the right rows."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500) + 61
    geom = (15360, 96, 40, 3, (12.0, 30.0), (-9.0, 14.0))
    procs = [ctx.Process(target=_worker, args=(r, world, port, False, q, geom)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(msg == "ok" for _, msg in results), results


def test_band_layout_rules():
    sys.path.insert(0, ROOT)
    from taa_star_b200.sharded import BandLayout, band_of
    for H, R in ((4320, 8), (2160, 7), (66, 3)):
        rows = [band_of(H, R, r) for r in range(R)]
        assert rows[0][0] == 0 and rows[-1][1] == H and all(rows[i][1] == rows[i + 1][0] for i in range(R - 1))
    L = BandLayout(4320, 8, 3, 20)
    assert (L.y0, L.y1, L.hy0, L.hy1, L.iy0, L.iy1) == (1620, 2160, 1600, 2180, 1618, 2162)
    assert BandLayout(4320, 8, 0, 20).hy0 == 0 and BandLayout(4320, 8, 7, 20).hy1 == 4320
    with pytest.raises(ValueError):
        BandLayout(64, 8, 0, 20)


def test_balanced_bounds_and_uneven_layouts():
    """Band boundaries that follow a measured cost (sharded.balanced_bounds) and the layout of uneven bands (pure host logic)."""
    from taa_star_b200.sharded import BandLayout, balanced_bounds
    H = 4320
    eq = [r * H // 4 for r in range(5)]
    assert balanced_bounds(eq, [1.0, 1.0, 1.0, 1.0]) == eq                      # equal cost: nothing moves
    b = balanced_bounds(eq, [0.08, 0.11, 0.105, 0.08])
    assert b[0] == 0 and b[-1] == H and all(y % 2 == 0 for y in b) and all(q > p for p, q in zip(b, b[1:]))
    rows = [q - p for p, q in zip(b, b[1:])]
    assert rows[1] < 1080 < rows[0] and rows[2] < 1080 < rows[3]                 # the expensive bands shrink, the cheap ones grow
    # cost implied by the new boundaries under the same piecewise-constant density: equal within a few rows' worth
    dens = [t / 1080 for t in (0.08, 0.11, 0.105, 0.08)]
    def cost(a, c):
        return sum(dens[k] * max(0, min(c, eq[k + 1]) - max(a, eq[k])) for k in range(4))
    costs = [cost(p, q) for p, q in zip(b, b[1:])]
    assert max(costs) - min(costs) < 3 * max(dens)
    assert balanced_bounds([0, 10, 20], [1.0, 0.0], min_rows=2) == [0, 4, 20]   # every band keeps min_rows
    for r in range(4):
        L = BandLayout(H, 4, r, halo=20, apron=2, bounds=b)
        assert (L.y0, L.y1) == (b[r], b[r + 1]) and L.hy0 == max(0, b[r] - 20) and L.hy1 == min(H, b[r + 1] + 20)
    with pytest.raises(ValueError):
        BandLayout(100, 2, 0, halo=20, bounds=[0, 90, 100])                      # a band lower than the halo
