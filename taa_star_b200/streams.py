"""Batches of independent camera streams (BASELINE configs[4]: 64 x 1080p offline-render streams over up to 8 GPUs, no communication).

Every camera stream is its own temporal sequence: its own history ping-pong, its own context (the tuned path keeps a fix-up list per
context). Streams are dealt to ranks round-robin (stream s -> rank s % world); on a rank they are dealt to a few CUDA streams, so that
the kernels of different camera streams overlap: a 1080p frame is only 2.3 waves of CTAs, and its latency-bound fix-up pass would leave
the GPU idle for ~10 us per frame if the frames ran back to back on one CUDA stream.

The reference is single-stream (one window, one `taa<CF>` invokee, main.cpp:4973); this is the batch shape the north star adds.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from . import host


def streams_of_rank(n_streams: int, world: int, rank: int) -> List[int]:
    """Round-robin deal: camera stream s runs on rank s % world."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return list(range(rank, n_streams, world))


class StreamBatch:
    """The camera streams of one rank. resolve(k, ...) enqueues one frame of local stream k; frames of one camera stream are ordered,
    frames of different camera streams are not."""

    def __init__(self, width: int, height: int, n_local: int, n_cuda_streams: int = 8, flags: int = 0, device=None):
        self.W, self.H, self.n = width, height, n_local
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.ctx = [host.TaaContext((width, height), flags=flags) for _ in range(n_local)]
        self.hist = [[torch.zeros(height, width, 4, dtype=torch.float16, device=self.device) for _ in range(2)] for _ in range(n_local)]
        self.result = [torch.zeros(height, width, 4, dtype=torch.float16, device=self.device) for _ in range(n_local)]
        self.parity = [0] * n_local
        self.cuda_streams = [torch.cuda.Stream(device=self.device) for _ in range(max(1, min(n_cuda_streams, max(n_local, 1))))]
        self._prepared = {}

    def cuda_stream_of(self, k: int) -> torch.cuda.Stream:
        return self.cuda_streams[k % len(self.cuda_streams)]

    @property
    def launch_count(self) -> int:
        return sum(c.launch_count for c in self.ctx)

    def resolve(self, k: int, uniforms, color, depth, velocity, history_depth=None, key=None):
        """One frame of local camera stream k on its CUDA stream. `key` (hashable) lets the argument block be reused when the same
        buffers come round again (benchmarks rotate a few frame sets)."""
        par = self.parity[k]
        s = self.cuda_stream_of(k)
        ck = None if key is None else (k, par, key)
        im = self._prepared.get(ck) if ck is not None else None
        if im is None:
            kw = dict(color=color, depth=depth, velocity=velocity, history_in=self.hist[k][par], history_out=self.hist[k][1 - par], result=self.result[k])
            if history_depth is not None:
                kw["history_depth"] = history_depth
            im = self.ctx[k].images(**kw)
            if ck is not None:
                self._prepared[ck] = im
        self.ctx[k].resolve_prepared(im, uniforms, s.cuda_stream)
        self.parity[k] ^= 1

    def fork_from(self, main: torch.cuda.Stream):
        """Every CUDA stream of the batch waits for what `main` holds so far."""
        ev = torch.cuda.Event()
        ev.record(main)
        for s in self.cuda_streams:
            s.wait_event(ev)

    def join_into(self, main: torch.cuda.Stream):
        """`main` waits for everything enqueued on the batch's CUDA streams."""
        for s in self.cuda_streams:
            ev = torch.cuda.Event()
            ev.record(s)
            main.wait_event(ev)

    def close(self):
        for c in self.ctx:
            c.close()
        self.ctx = []


def bench_main(args, ClockSampler, measured_peak, BYTES_PER_PX):
    """bench.py --config 5 [--gpus N]: 64 camera streams of 1920x1080 (config 2 settings), 64 / N per rank, no communication.
    A step = one frame of every camera stream. Launch with torchrun for N > 1."""
    import json
    import os
    import time

    import torch.distributed as dist

    from . import configs
    from .synth import SyntheticScene
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)
    W, H, NSTREAMS, NSETS = 1920, 1080, 64, 2
    mine = streams_of_rank(NSTREAMS, world, rank)
    p = configs.config2_resolve()
    n_cuda = int(os.environ.get("TAA_BATCH_CUDA_STREAMS", "8"))  # measured on one B200, 64 streams: 1 -> 40.9, 4 -> 66.1, 8 -> 68.8 Gpixel/s
    batch = StreamBatch(W, H, len(mine), n_cuda_streams=n_cuda, device=dev)
    # every camera stream has its own scene (seed) and pan, as different cameras would
    frames, unis = [], []
    for s in mine:
        sc = SyntheticScene(W, H, device=dev, with_aux=False, seed=0x7AA57A2 + s, pan_px=(3.0 - 0.125 * (s % 16), 0.5 + 0.25 * (s % 5)))
        fs = [sc.frame(n) for n in range(NSETS)]
        frames.append(fs)
        unis.append([configs.uniforms_for(p, f.jitter_ndc) for f in fs])
    main = torch.cuda.Stream(device=dev)

    def step(i):
        for k in range(len(mine)):
            f = frames[k][i % NSETS]
            u = unis[k][i % NSETS] if i > 0 else configs.uniforms_for(p, f.jitter_ndc, reset_history=True)
            batch.resolve(k, u, f.color, f.depth, f.velocity, key=i % NSETS if i > 0 else None)

    for i in range(args.warmup + 1):
        step(i)
    torch.cuda.synchronize()
    clocks = ClockSampler(local) if rank == 0 else None
    launches0 = batch.launch_count
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if clocks:
        clocks.region(True)
    ev0.record(main)
    batch.fork_from(main)
    for i in range(args.steps):
        step(i + args.warmup + 1)
    batch.join_into(main)
    ev1.record(main)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    if clocks:
        clocks.region(False)
    ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
    launches = torch.tensor([batch.launch_count - launches0], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(launches)
    ms_per_step = float(ms.item()) / args.steps
    px = W * H * NSTREAMS
    mpx_s = px / (ms_per_step * 1e-3) / 1e6
    peak, peak_src = measured_peak()
    achieved = BYTES_PER_PX[2] * px / world / (ms_per_step * 1e-3) / 1e9  # per GPU

    # ---- e2e: every frame's G-buffer from pinned host memory, every result back to the host ----
    e2e_steps = max(2, min(args.steps, 4))
    hcol = [[f.color.cpu().pin_memory() for f in fs] for fs in frames]
    hdep = [[f.depth.cpu().pin_memory() for f in fs] for fs in frames]
    hvel = [[f.velocity.cpu().pin_memory() for f in fs] for fs in frames]
    hres = [torch.empty(H, W, 4, dtype=torch.float16).pin_memory() for _ in mine]
    dbuf = [(torch.empty_like(frames[k][0].color), torch.empty_like(frames[k][0].depth), torch.empty_like(frames[k][0].velocity)) for k in range(len(mine))]

    def e2e_step(i):
        for k in range(len(mine)):
            s = batch.cuda_stream_of(k)
            with torch.cuda.stream(s):
                dbuf[k][0].copy_(hcol[k][i % NSETS], non_blocking=True)
                dbuf[k][1].copy_(hdep[k][i % NSETS], non_blocking=True)
                dbuf[k][2].copy_(hvel[k][i % NSETS], non_blocking=True)
            batch.resolve(k, unis[k][i % NSETS], *dbuf[k])
            with torch.cuda.stream(s):
                hres[k].copy_(batch.result[k], non_blocking=True)

    e2e_step(0)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        e2e_step(i + 1)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dt = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    e2e_mpx = e2e_steps * px / float(dt.item()) / 1e6
    if rank == 0:
        line = {
            "metric": "resolved Mpixels/s", "value": round(mpx_s, 1), "unit": "Mpixels/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(ms_per_step, 5), "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "fps": round(NSTREAMS * 1e3 / ms_per_step, 1),
            "config": {"workload": f"{NSTREAMS} independent {W}x{H} camera streams, BASELINE configs[4] (config 2 settings), {len(mine)} per GPU, no communication",
                       "arithmetic": "tuned kernel alone (no mask bound: nothing for the exact fix-up pass to decide)", "step": "one frame of every camera stream", "cuda_streams_per_gpu": len(batch.cuda_streams),
                       "l2": f"{len(mine)} streams x {NSETS} frame sets rotated per GPU ({len(mine) * NSETS * W * H * 20 / 1e6:.0f} MB of inputs), launches of different streams interleaved"},
            "gpu_launches": int(launches.item()),
            "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4), "traffic": None,
                         "peak_source": peak_src, "bytes_per_px": BYTES_PER_PX[2], "kernel": "taa_resolve_stream_kernel (per GPU)"},
            "cpu_baseline": None,
            "e2e": {"value": round(e2e_mpx, 1), "unit": "Mpixels/s", "h2d_bytes_per_step": px * 20, "d2h_bytes_per_step": px * 8, "steps": e2e_steps,
                    "path": "per camera stream: pinned host G-buffer -> H2D -> resolve -> D2H of the result, on the stream's CUDA stream"},
            "clocks": clocks.result() if clocks else None,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
