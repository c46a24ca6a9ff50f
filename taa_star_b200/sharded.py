"""Row-band sharding of one large frame across GPUs (BASELINE configs[3]).

The reference is single-GPU (source/main.cpp:4973); this is new work. Every output pixel of taa.comp depends on
  (i)  current-frame inputs within a 1-texel apron (3x3 neighbourhood, taa.comp:199-213; 5-tap velocity test, :802-806), and
  (ii) history_in within (motion + filter footprint) of the pixel (taa.comp:421, 441-513),
so rank r of R resolves output rows [r*H/R, (r+1)*H/R) from band-local buffers that carry `halo` extra history rows above and
below. With depth culling the previous frame's depth is read at the history position too (taa.comp:818), so that buffer needs
the same halo: each rank keeps (or renders) depth for its band +- halo rows. After each frame a rank sends the first / last `halo` rows of the band it just wrote to its upper / lower neighbour
(one grouped NCCL send/recv over NVLink). The boundary strips are resolved first so that the exchange overlaps the interior.
If motion is unbounded, `replicate=True` all-gathers the whole history instead (correct for any motion, but a scaling cliff).

One process per GPU; torch.distributed is the plumbing (NCCL on GPUs, gloo in the CPU tests of the exchange logic).
"""
from __future__ import annotations

import json
import os
import time
from dataclasses import dataclass
from typing import Callable, Optional

import torch
import torch.distributed as dist


def band_of(height: int, world: int, rank: int):
    """Rows [y0, y1) owned by `rank`."""
    return rank * height // world, (rank + 1) * height // world


def balanced_bounds(bounds, times, quantum: int = 2, min_rows: int = 1):
    """Row boundaries that equalise the bands' cost, given the time each band took with the boundaries `bounds` (cost taken as uniform
    inside a band): the cumulative cost is cut into equal parts. Boundaries are multiples of `quantum`, every band keeps >= min_rows rows."""
    world = len(times)
    dens = [t / max(1, bounds[r + 1] - bounds[r]) for r, t in enumerate(times)]
    total = sum(times)
    out = [bounds[0]]
    r, acc = 0, 0.0  # acc: cost of the bands before band r
    for k in range(1, world):
        target = total * k / world
        while r < world - 1 and acc + times[r] < target:
            acc += times[r]
            r += 1
        y = bounds[r] + (target - acc) / dens[r] if dens[r] > 0 else bounds[r + 1]
        y = int(round(y / quantum)) * quantum
        y = max(out[-1] + min_rows, min(y, bounds[-1] - (world - k) * min_rows))
        out.append(y)
    out.append(bounds[-1])
    return out


@dataclass
class BandLayout:
    height: int
    world: int
    rank: int
    halo: int        # history rows kept above/below the band
    apron: int = 2   # input rows kept above/below the band (3x3 neighbourhood + velocity taps + sampler bleed)
    bounds: Optional[list] = None  # world + 1 row boundaries (default: equal bands)

    def band(self, r: int):
        return (self.bounds[r], self.bounds[r + 1]) if self.bounds is not None else band_of(self.height, self.world, r)

    def __post_init__(self):
        if self.bounds is not None:
            assert len(self.bounds) == self.world + 1 and self.bounds[0] == 0 and self.bounds[-1] == self.height and all(b > a for a, b in zip(self.bounds, self.bounds[1:]))
        self.y0, self.y1 = self.band(self.rank)
        self.hy0, self.hy1 = max(0, self.y0 - self.halo), min(self.height, self.y1 + self.halo)
        self.iy0, self.iy1 = max(0, self.y0 - self.apron), min(self.height, self.y1 + self.apron)
        smallest = min(b - a for a, b in (self.band(r) for r in range(self.world)))
        if self.world > 1 and smallest < self.halo:
            raise ValueError(f"halo {self.halo} exceeds the smallest band ({smallest} rows): use fewer ranks or replicate=True")

    @property
    def rows(self):
        return self.y1 - self.y0


class HaloExchanger:
    """Moves the freshly written boundary rows of a band-local history buffer into the neighbours' halo rows."""

    def __init__(self, layout: BandLayout, group=None):
        self.L = layout
        self.group = group

    def ops(self, hist: torch.Tensor):
        """hist: (hy1 - hy0, W, 4) band-local history buffer whose row 0 is global row hy0."""
        L = self.L
        ops = []
        h = L.halo
        if L.rank > 0:  # upper neighbour: it needs my first h rows; I need its last h rows
            send = hist[L.y0 - L.hy0: L.y0 - L.hy0 + h]
            recv = hist[L.y0 - h - L.hy0: L.y0 - L.hy0]
            ops += [dist.P2POp(dist.isend, send, L.rank - 1, self.group), dist.P2POp(dist.irecv, recv, L.rank - 1, self.group)]
        if L.rank < L.world - 1:
            send = hist[L.y1 - h - L.hy0: L.y1 - L.hy0]
            recv = hist[L.y1 - L.hy0: L.y1 - L.hy0 + h]
            ops += [dist.P2POp(dist.isend, send, L.rank + 1, self.group), dist.P2POp(dist.irecv, recv, L.rank + 1, self.group)]
        return ops

    def exchange(self, hist: torch.Tensor):
        ops = self.ops(hist)
        if not ops:
            return []
        return dist.batch_isend_irecv(ops)


def gather_full_history(band: torch.Tensor, full: torch.Tensor, layout: BandLayout, group=None):
    """replicate=True: all-gather the bands of history_out into a full-frame buffer on every rank."""
    sizes = [layout.band(r) for r in range(layout.world)]
    if len({b - a for a, b in sizes}) == 1:
        dist.all_gather_into_tensor(full, band.contiguous(), group=group)
    else:
        outs = [full[a:b] for a, b in sizes]
        dist.all_gather(outs, band.contiguous(), group=group)


class ShardedTaa:
    """One rank's share of a row-band sharded resolve. Inputs for the band (plus apron) are produced locally."""

    def __init__(self, width: int, height: int, halo: int = 20, replicate: bool = False, flags: int = 0, device=None, group=None, apron: int = 2,
                 exchange: str = "nccl", bounds=None):
        """exchange = "nccl": the boundary strips are resolved by their own launches and sent / received (works for every settings block);
        "peer": ONE launch per band and frame, the resolve kernel stores the boundary rows into the neighbours' halos itself and the next frame's
        boundary units wait on a flag (taa_band_peers; calls the streaming kernel serves alone, i.e. the config 2 family)."""
        from . import host
        assert exchange in ("nccl", "peer")
        self.W, self.H = width, height
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.L = BandLayout(height, self.world, self.rank, halo, apron, bounds)
        self.replicate = replicate
        self.group = group
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.peer = exchange == "peer" and self.world > 1 and not replicate
        L = self.L
        rows = (height if replicate else L.hy1 - L.hy0)
        self.hist_y0 = 0 if replicate else L.hy0
        if self.peer:
            # every rank must end up in the same mode: if any rank cannot map its neighbours (no peer access between the GPUs, IPC refused),
            # all of them fall back to the NCCL exchange
            ok, why = 1, ""
            try:
                self._peer_setup(rows, flags)
            except Exception as e:  # noqa: BLE001
                ok, why = 0, f"{type(e).__name__}: {e}"
            flag = torch.tensor([ok], device=self.device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
            if int(flag.item()) == 1:
                self.peer_reset()
                return
            self.peer = False
            self.peer_fallback_reason = why or "another rank could not map its neighbours"
        self.hist = [torch.zeros(rows, width, 4, dtype=torch.float16, device=self.device) for _ in range(2)]
        self.result = torch.zeros(L.rows, width, 4, dtype=torch.float16, device=self.device)
        self.band_tmp = torch.zeros(L.rows, width, 4, dtype=torch.float16, device=self.device) if replicate else None
        # three sub-bands: the two boundary strips (exchanged) first, then the interior
        h = min(halo, L.rows // 2) if (self.world > 1 and not replicate) else 0
        strips = []
        if h and L.rank > 0:
            strips.append((L.y0, h))
        if h and L.rank < L.world - 1:
            strips.append((L.y1 - h, h))
        a = L.y0 + (h if (h and L.rank > 0) else 0)
        b = L.y1 - (h if (h and L.rank < L.world - 1) else 0)
        self.boundary = [host.TaaContext((width, height), band=s, flags=flags) for s in strips]
        self.interior = host.TaaContext((width, height), band=(a, b - a), flags=flags) if b > a else None
        self.xchg = HaloExchanger(L, group)
        self.compute = torch.cuda.Stream(device=self.device)
        self.comm = torch.cuda.Stream(device=self.device)
        # the boundary strips gate the exchange: each gets its own high-priority stream and runs beside the interior resolve (a 20-row
        # strip is a quarter of a wave of CTAs; back to back with its fix-up pass it would cost a CTA lifetime of an otherwise idle GPU)
        self.s_boundary = [torch.cuda.Stream(device=self.device, priority=-1) for _ in self.boundary]
        self.ev_boundary = [torch.cuda.Event() for _ in self.boundary]
        self.ev_start = torch.cuda.Event()
        self.ev_comm = torch.cuda.Event()
        self.parity = 0
        self._pending = []
        self._capturing = False
        self._ev_comm_captured = False
        self._comm_in_capture = False

    # ---- exchange = "peer" ---------------------------------------------------------------------------------------------------------
    def _peer_setup(self, rows: int, flags: int):
        """One exportable allocation per rank: [history 0 | history 1 | flag block]; the handles travel once through torch.distributed."""
        import ctypes as C
        from . import abi, host
        L, W = self.L, self.W
        lib = abi.load_library()
        self._lib = lib
        hbytes = rows * W * 8
        handle = (C.c_ubyte * abi.TAA_IPC_HANDLE_BYTES)()
        self._arena = lib.taa_device_alloc(2 * hbytes + 4 * abi.TAA_BAND_FLAG_WORDS)
        local_ok = bool(self._arena) and lib.taa_ipc_export(self._arena, handle) == abi.TAA_OK
        mine = dict(ok=local_ok, handle=bytes(handle), hbytes=hbytes, hy0=L.hy0, band_rows=L.rows)
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine, group=self.group)  # (the only collective in here: every rank reaches it, whatever failed locally)
        if not all(e["ok"] for e in everyone):
            raise RuntimeError("a rank could not allocate or export its band buffers")
        self.hist = [host.tensor_from_ptr(self._arena + k * hbytes, (rows, W, 4), torch.float16) for k in range(2)]
        self._flags_ptr = self._arena + 2 * hbytes
        self._mapped = []
        peers = [None, None]
        for sd, nb in enumerate((self.rank - 1, self.rank + 1)):
            if nb < 0 or nb >= self.world:
                continue
            info = everyone[nb]
            base = C.c_void_p()
            st = lib.taa_ipc_open((C.c_ubyte * abi.TAA_IPC_HANDLE_BYTES).from_buffer_copy(info["handle"]), C.byref(base))
            assert st == abi.TAA_OK, f"taa_ipc_open failed: {lib.taa_last_error_string(None).decode()}"
            self._mapped.append(base.value)
            pb = abi.taa_band_peer()
            pb.history[0], pb.history[1] = base.value, base.value + info["hbytes"]
            pb.row_pitch, pb.y0, pb.band_rows = W * 8, info["hy0"], info["band_rows"]
            pb.flags = base.value + 2 * info["hbytes"]
            peers[sd] = pb
        self._peers = peers
        self.result = torch.zeros(L.rows, W, 4, dtype=torch.float16, device=self.device)
        self.band_tmp = None
        self.boundary, self.s_boundary, self.ev_boundary = [], [], []
        self.interior = host.TaaContext((W, self.H), band=(L.y0, L.rows), flags=flags)
        self.compute = torch.cuda.Stream(device=self.device)
        self.comm = self.compute
        self.ev_comm = torch.cuda.Event()
        self.parity = 0
        self._pending = []
        self._capturing = self._ev_comm_captured = self._comm_in_capture = False

    def peer_reset(self):
        """(Re)starts a frame sequence: flag blocks zeroed, the next step waits for nobody. Collective (a barrier on each side)."""
        import ctypes as C
        torch.cuda.synchronize()
        dist.barrier(group=self.group)  # nobody is still signalling into a block that is about to be zeroed
        up, dn = self._peers
        st = self._lib.taa_band_peers(self.interior._h, C.byref(up) if up is not None else None, C.byref(dn) if dn is not None else None,
                                      self.hist[0].data_ptr(), self.hist[1].data_ptr(), self._flags_ptr, self.L.halo)
        self.interior._check(st, "taa_band_peers")
        self.parity = 0
        dist.barrier(group=self.group)

    def _peer_step(self, uniforms, color, depth, velocity, in_y0, history_depth=None):
        L = self.L
        hin, hout = self.hist[self.parity], self.hist[1 - self.parity]
        kw = dict(color=(color, in_y0), depth=(depth, in_y0), velocity=(velocity, in_y0), history_in=(hin, self.hist_y0),
                  history_out=(hout, self.hist_y0), result=(self.result, L.y0))
        if history_depth is not None:
            kw["history_depth"] = (history_depth, in_y0)
        with torch.cuda.stream(self.compute):
            self.interior.resolve(uniforms, stream=self.compute, **kw)
        self.parity ^= 1

    def close(self):
        if self.peer and getattr(self, "_arena", None):
            torch.cuda.synchronize()
            dist.barrier(group=self.group)
            self.hist = []
            for m in self._mapped:
                self._lib.taa_ipc_close(m)
            self.interior.close()
            self._lib.taa_device_free(self._arena)
            self._arena = None

    @property
    def launch_count(self):
        return sum(c.launch_count for c in self.boundary) + (self.interior.launch_count if self.interior else 0)

    def step(self, uniforms, color, depth, velocity, in_y0: int, history_depth=None):
        """Resolves this rank's band for one frame. color/depth/velocity hold input rows starting at global row in_y0."""
        if not self._capturing:  # inputs are usually produced on the caller's current stream: order them before this step's launches
            self.compute.wait_stream(torch.cuda.current_stream(self.device))
        if self.peer:
            return self._peer_step(uniforms, color, depth, velocity, in_y0, history_depth)
        L = self.L
        hin, hout = self.hist[self.parity], self.hist[1 - self.parity]
        kw = dict(color=(color, in_y0), depth=(depth, in_y0), velocity=(velocity, in_y0), history_in=(hin, self.hist_y0),
                  history_out=(hout, self.hist_y0), result=(self.result, L.y0))
        if history_depth is not None:
            kw["history_depth"] = (history_depth, in_y0)
        exchanging = self.world > 1 and not self.replicate
        # The interior rows [y0 + halo, y1 - halo) never read a halo row (motion bound + filter footprint < halo): the interior resolve of
        # step n+1 depends on this rank's own step-n kernels only, NOT on exchange n. It is handed the band rows of the history alone, so a
        # read beyond them is reported (TAA_E_HALO_OVERFLOW) instead of racing with the exchange. Only the boundary strips wait for the halos.
        kw_int = kw
        if exchanging:
            kw_int = dict(kw, history_in=(hin[L.y0 - self.hist_y0: L.y1 - self.hist_y0], L.y0), history_out=(hout[L.y0 - self.hist_y0: L.y1 - self.hist_y0], L.y0))
        with torch.cuda.stream(self.compute):
            self.ev_start.record(self.compute)  # everything queued on `compute` so far (input copies, the previous step) is ordered before
        for c, sb, ev in zip(self.boundary, self.s_boundary, self.ev_boundary):
            with torch.cuda.stream(sb):
                sb.wait_event(self.ev_start)
                if exchanging:  # halos of `hin` (written by the previous exchange) must have landed
                    if self._capturing:
                        if self._comm_in_capture:
                            sb.wait_event(self.ev_comm)  # (first step of a capture: the previous graph ended with the exchange joined)
                    else:
                        if self._ev_comm_captured:  # an event last recorded inside a capture cannot be waited on outside of it
                            self.ev_comm, self._ev_comm_captured = torch.cuda.Event(), False
                        sb.wait_event(self.ev_comm)
                c.resolve(uniforms, stream=sb, **kw)
                ev.record(sb)
        if self.interior is not None:
            with torch.cuda.stream(self.compute):
                self.interior.resolve(uniforms, stream=self.compute, **kw_int)
        if self.world > 1:
            if self.replicate:
                with torch.cuda.stream(self.compute):
                    band = hout[L.y0:L.y1]
                    gather_full_history(band, hout, L, self.group)
                    self.ev_comm.record(self.compute)
            else:
                with torch.cuda.stream(self.comm):
                    for ev in self.ev_boundary:
                        self.comm.wait_event(ev)
                    for w in self._pending:
                        w.wait()
                    self._pending = self.xchg.exchange(hout)
                    for w in self._pending:
                        w.wait()
                    self._pending = []
                    self.ev_comm.record(self.comm)
                    self._comm_in_capture = self._capturing
        with torch.cuda.stream(self.compute):  # join: what follows on `compute` (next step, result copies) sees the whole band
            for ev in self.ev_boundary:
                self.compute.wait_event(ev)
        self.parity ^= 1

    def capture_step(self, *args, **kw) -> "torch.cuda.CUDAGraph":
        """Captures one step() (all launches on the three resolve streams, the NCCL halo exchange and their dependencies) into a CUDA
        graph on `compute`; the caller replays it on `compute`. A step costs ~0.2-0.3 ms of host time to enqueue, more than the GPU needs
        at 4+ ranks. The graph bakes in this step's arguments (uniforms, buffers, history parity, the contexts' fix-up counter parity):
        capture one graph per distinct step of a cycle of EVEN length and replay them in capture order. Capturing does not run the step
        but does advance the host-side parities, exactly as the replay will find them."""
        return self.capture_steps([(args, kw)])

    def capture_steps(self, steps) -> "torch.cuda.CUDAGraph":
        """Several consecutive steps [(args, kwargs), ...] in ONE graph: no graph-launch gap between them."""
        assert not self._pending
        g = torch.cuda.CUDAGraph()
        self._capturing = True
        try:
            with torch.cuda.graph(g, stream=self.compute, capture_error_mode="thread_local"):
                self._comm_in_capture = False
                for a, k in steps:
                    self.step(*a, **k)
                if self._comm_in_capture:  # every forked stream joins `compute` before the capture ends
                    self.compute.wait_event(self.ev_comm)
        finally:
            self._capturing = False
            self._ev_comm_captured = self._comm_in_capture
            self._comm_in_capture = False
        return g

    def poll(self) -> int:
        st = 0
        self.last_poll_error = ""
        for c in self.boundary + ([self.interior] if self.interior else []):
            s1 = c.poll_status(self.compute)
            if s1 != 0:
                self.last_poll_error = c._lib.taa_last_error_string(c._h).decode()
            st = min(st, s1)
        return st


def _whole_frame_reference(W, H, p, flags, dev, cfg_id, sh, halo, replicate):
    """Runs frames 0 (history reset) and 1 of the synthetic sequence on this GPU alone over the whole W x H frame and through the sharded
    pipeline; returns (ms per whole-frame step on one GPU, this rank's band identical bit for bit). Frame 1 reads the halo rows the
    neighbours wrote in frame 0, so the comparison covers the exchange."""
    from . import abi, configs, host
    from .synth import SyntheticScene
    L = sh.L
    sc = SyntheticScene(W, H, device=dev, with_aux=False)
    f = [sc.frame(0), sc.frame(1)]
    ctx = host.TaaContext((W, H), flags=flags)
    hist = [torch.zeros(H, W, 4, dtype=torch.float16, device=dev) for _ in range(2)]
    res = torch.zeros(H, W, 4, dtype=torch.float16, device=dev)
    u = [configs.uniforms_for(p, f[0].jitter_ndc, reset_history=True), configs.uniforms_for(p, f[1].jitter_ndc)]

    def whole(i, par):
        ctx.resolve(u[i], color=f[i].color, depth=f[i].depth, velocity=f[i].velocity, history_in=hist[par], history_out=hist[1 - par], result=res,
                    history_depth=f[i - 1].depth if cfg_id == 3 else None)
    whole(0, 0)
    whole(1, 1)
    torch.cuda.synchronize()
    want_hist, want_res = hist[0][L.y0:L.y1].clone(), res[L.y0:L.y1].clone()
    # the sharded pipeline on the same two frames (band rows + apron of the same inputs)
    dist.barrier()  # (the ranks reach this point seconds apart; a peer-store step waits for its neighbours on the device)
    for i in range(2):
        sh.step(u[i], f[i].color[L.iy0:L.iy1], f[i].depth[L.iy0:L.iy1], f[i].velocity[L.iy0:L.iy1], L.iy0,
                history_depth=f[i - 1].depth[L.iy0:L.iy1] if cfg_id == 3 else None)
    torch.cuda.synchronize()
    st = sh.poll()
    assert st == 0, f"status {st} in the sharded steps of the whole-frame check: {sh.last_poll_error}"
    got_hist = sh.hist[sh.parity][L.y0 - sh.hist_y0: L.y1 - sh.hist_y0]
    # Exact kernel: bit for bit. Tuned kernels: bit for bit too wherever the motion is not EXACTLY a whole number of texels; where it is (this
    # scene's 3.0 px pan and its mover), the Catmull-Rom footprint start sits on a rounding tie, the uniform-motion rows break the tie once per
    # unit, and the unit boundaries of a band differ from the whole frame's: a few hundred pixels of an 8K frame then differ in the last fp16 bit.
    # The bar is the parity bar of the tuned path (2^-10 per channel); the count of differing pixels is reported.
    dh = (got_hist.view(torch.int16) != want_hist.view(torch.int16)).any(dim=2)
    dr = (sh.result.view(torch.int16) != want_res.view(torch.int16)).any(dim=2)
    ndiff = int(dh.sum()) + int(dr.sum())
    maxd = max(float((got_hist.float() - want_hist.float()).abs().max()), float((sh.result.float() - want_res.float()).abs().max()))
    same = (ndiff == 0) if (flags & 1) else (maxd <= 2.0 ** -10)
    if ndiff:
        rows_h = torch.nonzero(dh.any(dim=1)).flatten()
        print(f"[whole-frame check] rank {L.rank}: {ndiff} texels differ from the whole-frame run, max |d| = {maxd:.3g} (band rows {rows_h[:4].tolist()} ..)", flush=True)
    # time the whole frame on one GPU (history ping-pong over the two frames; 2 x 20 B/px of inputs exceed the L2 from 4K upwards)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for k in range(4):
        whole(k & 1, k & 1)
    torch.cuda.synchronize()
    n = 20
    ev0.record()
    for k in range(n):
        whole(k & 1, k & 1)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / n
    ctx.close()
    del hist, res, f, sc
    torch.cuda.empty_cache()
    # leave the sharded pipeline as a fresh one would be
    for h in sh.hist:
        h.zero_()
    if sh.parity:
        sh.parity = 0
    if sh.peer:
        sh.peer_reset()
    dist.barrier()
    return ms, same, ndiff, maxd


def _band_alone_ms(W, H, p, flags, dev, cfg_id, y0, y1, halo, apron, frames_n: int = 24):
    """Device time per frame of rows [y0, y1) resolved on this GPU alone (no neighbours, no exchange): what the band's content costs."""
    from . import configs, host
    from .synth import SyntheticScene
    hy0, hy1, iy0, iy1 = max(0, y0 - halo), min(H, y1 + halo), max(0, y0 - apron), min(H, y1 + apron)
    sc = SyntheticScene(W, H, device=dev, with_aux=False, rows=(iy0, iy1))
    f = [sc.frame(0), sc.frame(1)]
    ctx = host.TaaContext((W, H), band=(y0, y1 - y0), flags=flags)
    hist = [torch.zeros(hy1 - hy0, W, 4, dtype=torch.float16, device=dev) for _ in range(2)]
    res = torch.zeros(y1 - y0, W, 4, dtype=torch.float16, device=dev)
    u = [configs.uniforms_for(p, x.jitter_ndc) for x in f]
    stream = torch.cuda.Stream(device=dev)
    prep = [ctx.images(color=(f[k].color, iy0), depth=(f[k].depth, iy0), velocity=(f[k].velocity, iy0), history_in=(hist[k], hy0), history_out=(hist[1 - k], hy0),
                       result=(res, y0), history_depth=(f[1 - k].depth, iy0) if cfg_id == 3 else None) for k in range(2)]
    torch.cuda.synchronize()
    for k in range(6):
        ctx.resolve_prepared(prep[k & 1], u[k & 1], stream.cuda_stream)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for k in range(frames_n):
        ctx.resolve_prepared(prep[k & 1], u[k & 1], stream.cuda_stream)
    ev1.record(stream)
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / frames_n
    ctx.close()
    return ms


# ---- bench (N > 1) ------------------------------------------------------------------------------------------------------------------
def bench_main(args, ClockSampler, measured_peak, BYTES_PER_PX):
    from . import abi, configs
    from .synth import SyntheticScene
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)
    W, H = (7680, 4320) if (args.width, args.height) == (3840, 2160) else (args.width, args.height)
    cfg_id = args.config
    p = configs.config2_resolve() if cfg_id == 2 else configs.config3_full_chain()
    flags = abi.TAA_FLAG_EXACT if args.exact else 0
    halo = 20  # |v_y| <= 16 px guaranteed by the generator + 2 filter + 2 guard
    replicate = bool(getattr(args, "replicate", False))
    if W > 7680:
        halo = halo * W // 7680  # (the synthetic motion and the fix-up band scale with the frame: same angular motion, more pixels)
    # config 2 (the streaming kernel alone): boundary rows stored into the neighbours' halos by the kernel, one launch per band and frame;
    # config 3 / the exact kernel (a fix-up or general launch rewrites pixels): boundary strips + NCCL send/recv. TAA_SHARDED_EXCHANGE overrides.
    # (measured, 8K: 2 GPUs 0.187 ms over NCCL against 0.195 with peer stores — the three launches of the NCCL variant overlap consecutive frames;
    # 4 GPUs 0.1215 against 0.1116; 8 GPUs 0.1071 against 0.0702)
    exchange = os.environ.get("TAA_SHARDED_EXCHANGE", "peer" if (cfg_id == 2 and not args.exact and not replicate and world >= 3) else "nccl")
    apron = halo if cfg_id == 3 else 2
    # Optional (TAA_SHARDED_BALANCE=1): band heights that follow the content's cost — every rank times its band alone, the cumulative cost is
    # cut into equal parts, twice. Measured on 4 x B200 (8K, the bench scene): rows [1130, 1002, 1048, 1140], 0.1121 ms against 0.1109 with
    # equal bands: the ranks run in lockstep and the step time is not set by the content of the slowest band, so equal bands stay the default.
    bounds, balance_note = None, "equal bands"
    if world > 1 and not replicate and os.environ.get("TAA_SHARDED_BALANCE", "0") == "1":
        bounds = [r * H // world for r in range(world + 1)]
        for _ in range(2):
            t = torch.tensor([_band_alone_ms(W, H, p, flags, dev, cfg_id, bounds[rank], bounds[rank + 1], halo, apron)], device=dev)
            ts = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(ts, t)
            times = [float(x.item()) for x in ts]
            bounds = balanced_bounds(bounds, times, quantum=2, min_rows=max(3 * halo, 64))
        balance_note = f"band heights follow the measured cost of the content (each band timed alone, two rounds): rows {[b - a for a, b in zip(bounds, bounds[1:])]}"
    sh = ShardedTaa(W, H, halo=halo, replicate=replicate, flags=flags, device=dev, apron=apron, exchange=exchange, bounds=bounds)
    L = sh.L
    NSETS = 4
    # ---- before anything is timed: the same two frames on ONE GPU (every rank does it for itself), (a) as the strong-scaling reference
    # of this very frame size, (b) to check that this rank's band of the sharded result equals the whole-frame result bit for bit ----
    single_ms, same, ndiff, maxd = None, None, None, None
    if not getattr(args, "no_verify", False):
        single_ms, same, ndiff, maxd = _whole_frame_reference(W, H, p, flags, dev, cfg_id, sh, halo, replicate)
    sc = SyntheticScene(W, H, device=dev, with_aux=False, rows=(L.iy0, L.iy1))
    frames = [sc.frame(n) for n in range(NSETS)]
    unis = [configs.uniforms_for(p, f.jitter_ndc) for f in frames]
    u0 = configs.uniforms_for(p, frames[0].jitter_ndc, reset_history=True)

    def step(i, u=None):
        f, fp = frames[i % NSETS], frames[(i - 1) % NSETS]
        sh.step(u or unis[i % NSETS], f.color, f.depth, f.velocity, L.iy0, history_depth=fp.depth if cfg_id == 3 else None)

    step(0, u0)
    if os.environ.get("TAA_SHARDED_DEBUG"):
        torch.cuda.synchronize()
        print(f"[debug] rank {rank} step 0: status {sh.poll()} {sh.last_poll_error}; velocity finite {[bool(torch.isfinite(f.velocity.float()).all()) for f in frames]}", flush=True)
    nwarm = max(args.warmup, 1)
    nwarm += (2 * NSETS - (nwarm + 1) % (2 * NSETS)) % (2 * NSETS)  # the next step index is a multiple of the cycle (frame set x history parity)
    for i in range(1, nwarm + 1):
        step(i)
        if os.environ.get("TAA_SHARDED_DEBUG"):
            torch.cuda.synchronize()
            print(f"[debug] rank {rank} warm-up step {i}: status {sh.poll()} {sh.last_poll_error}", flush=True)
    torch.cuda.synchronize()
    st = sh.poll()
    assert st == abi.TAA_OK, f"status {st} during warm-up: {sh.last_poll_error}"
    graphs, cycle_graph, graph_note = None, None, "eager"
    if os.environ.get("TAA_SHARDED_GRAPHS", "1") != "0":
        # One CUDA graph per step of the cycle (kernels + NCCL exchange, see capture_steps) and the whole cycle as one graph (no launch gap
        # between its steps; used wherever a full cycle fits into the timed steps). Capturing does not communicate, replaying does: every
        # rank captures everything first, the ranks agree that all captures succeeded, and only then is anything replayed.
        def args_of(i):
            f, fp = frames[i % NSETS], frames[(i - 1) % NSETS]
            return (unis[i % NSETS], f.color, f.depth, f.velocity, L.iy0), dict(history_depth=fp.depth if cfg_id == 3 else None)

        parity0, ok, why = sh.parity, 1, ""
        launches_c0 = sh.launch_count  # (the contexts count launches as they record them: a replay repeats what the capture recorded)
        try:
            gs = [sh.capture_steps([args_of(nwarm + 1 + k)]) for k in range(2 * NSETS)]
            launches_per_cycle = sh.launch_count - launches_c0
            cg = sh.capture_steps([args_of(nwarm + 1 + 2 * NSETS + k) for k in range(2 * NSETS)])
        except Exception as e:  # noqa: BLE001 - any capture failure means: run eagerly
            ok, why = 0, f"{type(e).__name__}: {str(e)[:120]}"
            sh.parity = parity0
            sh._pending = []
        flag = torch.tensor([ok], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 1:
            graphs, cycle_graph = gs, cg
            graph_note = ("CUDA graphs (one resolve launch per step): one per cycle of 8 steps, single-step graphs for the remainder" if sh.peer else
                          "CUDA graphs (kernels on 3 streams + the NCCL exchange): one per cycle of 8 steps, single-step graphs for the remainder")
            with torch.cuda.stream(sh.compute):
                for g in graphs:
                    g.replay()  # run them once: the device state follows the host-side parities the captures advanced
                cycle_graph.replay()
            torch.cuda.synchronize()
            assert sh.poll() == abi.TAA_OK
            nwarm += 4 * NSETS

            def step(i, u=None):  # noqa: F811
                with torch.cuda.stream(sh.compute):
                    graphs[(i - (nwarm + 1)) % (2 * NSETS)].replay()
        else:
            graph_note = "eager (graph capture failed on some rank" + (f": {why}" if why else "") + ")"
    clocks = ClockSampler(local) if rank == 0 else None
    launches0 = sh.launch_count
    dist.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if clocks:
        clocks.region(True)
    ev0.record(sh.compute)
    t_host0 = time.perf_counter()
    i = 0
    while i < args.steps:  # exactly args.steps steps
        if graphs is not None and i % (2 * NSETS) == 0 and args.steps - i >= 2 * NSETS:
            with torch.cuda.stream(sh.compute):
                cycle_graph.replay()
            i += 2 * NSETS
        else:
            step(i + nwarm + 1)
            i += 1
    if graphs is None:  # (a captured step ends with the exchange joined into `compute`)
        sh.compute.wait_event(sh.ev_comm)
    ev1.record(sh.compute)
    host_ms_per_step = (time.perf_counter() - t_host0) * 1e3 / args.steps  # time to ENQUEUE a step (a step is host-bound if this ~ ms_per_step)
    torch.cuda.synchronize()
    if graphs is not None:  # complete the cycle, so that the device-side state matches the host-side parities again
        for i in range(args.steps, args.steps + (-args.steps) % (2 * NSETS)):
            step(i + nwarm + 1)
        torch.cuda.synchronize()
    dist.barrier()
    if clocks:
        clocks.region(False)
    ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
    per_rank = [torch.zeros_like(ms) for _ in range(world)]
    dist.all_gather(per_rank, ms)
    per_rank_ms = [round(float(t.item()) / args.steps, 5) for t in per_rank]  # each rank's own device time per step (content differs per band)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    launches = torch.tensor([sh.launch_count - launches0 if graphs is None else launches_per_cycle * args.steps // (2 * NSETS)], device=dev)
    dist.all_reduce(launches)
    # kernel-only time of this rank's interior + boundary launches is not separable from the exchange here; the roofline entry uses
    # the per-step time of the whole sharded step (conservative).
    px = W * H
    ms_per_step = ms / args.steps
    mpx_s = px / (ms_per_step * 1e-3) / 1e6
    peak, peak_src = measured_peak()
    achieved = BYTES_PER_PX[cfg_id] * px / world / (ms_per_step * 1e-3) / 1e9  # per GPU

    # ---- e2e: band inputs from pinned host memory, result band back to the host, every step ----
    hcol = [f.color.cpu().pin_memory() for f in frames]
    hdep = [f.depth.cpu().pin_memory() for f in frames]
    hvel = [f.velocity.cpu().pin_memory() for f in frames]
    hres = torch.empty_like(sh.result, device="cpu").pin_memory()
    dcol, ddep, dvel = torch.empty_like(frames[0].color), torch.empty_like(frames[0].depth), torch.empty_like(frames[0].velocity)
    e2e_steps = max(8, min(args.steps, 32))

    def e2e_step(i):
        k = i % NSETS
        with torch.cuda.stream(sh.compute):
            dcol.copy_(hcol[k], non_blocking=True)
            ddep.copy_(hdep[k], non_blocking=True)
            dvel.copy_(hvel[k], non_blocking=True)
        sh.step(unis[k], dcol, ddep, dvel, L.iy0, history_depth=None if cfg_id == 2 else frames[(i - 1) % NSETS].depth)
        with torch.cuda.stream(sh.compute):
            hres.copy_(sh.result, non_blocking=True)

    for i in range(3):
        e2e_step(i)
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        e2e_step(i + 3)
    torch.cuda.synchronize()
    dist.barrier()
    dt = torch.tensor([time.perf_counter() - t0], device=dev)
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    e2e_mpx = e2e_steps * px / float(dt.item()) / 1e6
    in_rows = L.iy1 - L.iy0
    if same is not None:  # every rank's band must equal the whole-frame result
        ok = torch.tensor([1 if same else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        same = bool(ok.item())
        nd = torch.tensor([float(ndiff), maxd], device=dev)
        nds = nd.clone()
        dist.all_reduce(nds, op=dist.ReduceOp.SUM)
        dist.all_reduce(nd, op=dist.ReduceOp.MAX)
        ndiff, maxd = int(nds[0].item()), float(nd[1].item())
        assert same, "a band of the sharded result differs from the whole-frame result by more than the parity bar"
    lps = float(launches.item()) / max(args.steps, 1) / world  # launches per step and rank
    if rank == 0:
        line = {
            "metric": "resolved Mpixels/s", "value": round(mpx_s, 1), "unit": "Mpixels/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(ms_per_step, 5), "host_enqueue_ms_per_step": round(host_ms_per_step, 5), "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "fps": round(1e3 / ms_per_step, 1),
            "config": {"workload": f"{W}x{H} TAA resolve sharded in {world} row bands, BASELINE configs[3] (config {cfg_id} settings)",
                       "arithmetic": "exact general kernel" if args.exact else ("tuned kernels + exact fix-up pass" if cfg_id == 3 else "tuned kernels alone (no mask bound: nothing for the exact fix-up pass to decide)"),
                       "halo_rows": halo, "bands": balance_note,
                       "exchange": ("all-gather of the history bands (replicated history: correct for unbounded motion)" if replicate else
                                    "peer stores: the resolve kernel writes its first / last halo rows into the neighbours' halos over NVLink and signals a flag; "
                                    "the neighbours' boundary units of the next frame wait on it (taa_band_peers) - one launch per band and frame, no collective" if sh.peer else
                                    "NCCL send/recv of 2 x halo rows of history per neighbour per frame, overlapped with the interior resolve"),
                       "peer_fallback": getattr(sh, "peer_fallback_reason", None),
                       "launch": graph_note,
                       "l2": f"inputs larger than L2: {NSETS} frame sets rotated, history ping-pong"},
            "gpu_launches": int(launches.item()), "gpu_launches_per_step_and_rank": round(lps, 2),
            "per_rank_ms_per_step": per_rank_ms,
            "single_gpu_same_frame_ms": round(single_ms, 5) if single_ms is not None else None,
            "strong_scaling_efficiency_vs_same_frame": round(single_ms / ms_per_step / world, 4) if single_ms is not None else None,
            "sharded_vs_whole_frame": None if same is None else {"within_parity_bar": same, "differing_texels": ndiff, "max_abs_diff": maxd,
                                                                  "note": "bit for bit unless the motion is exactly a whole number of texels (rounding tie of the footprint start, broken per unit)"},
            "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4), "traffic": None,
                         "peak_source": peak_src, "bytes_per_px": BYTES_PER_PX[cfg_id], "kernel": "taa_resolve (per GPU, whole sharded step incl. exchange)"},
            "cpu_baseline": None,
            "e2e": {"value": round(e2e_mpx, 1), "unit": "Mpixels/s", "h2d_bytes_per_step": in_rows * W * 20 * world, "d2h_bytes_per_step": px * 8, "steps": e2e_steps,
                    "path": "per rank: pinned host band (+apron) -> H2D -> band resolve + halo exchange -> D2H of the result band"},
            "clocks": clocks.result() if clocks else None,
        }
        print(json.dumps(line), flush=True)
    # tear-down: graphs that captured NCCL work go first; the process-group destructor has been seen to hang behind them, so it gets
    # a deadline (the line is printed and flushed by now)
    import gc
    import sys
    import threading
    graphs = None
    step = None
    cycle_graph = None
    gc.collect()
    torch.cuda.synchronize()
    dist.barrier()
    sys.stdout.flush()
    t = threading.Thread(target=dist.destroy_process_group, daemon=True)
    t.start()
    t.join(20.0)
    if t.is_alive():
        os._exit(0)
