#!/usr/bin/env python
"""bench.py — resolved Mpixels/s of the TAA resolve on N B200s (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 2|3]

N = 1 : one 3840x2160 stream, BASELINE configs[1] ("4K TAA resolve: YCoCg variance clip, Catmull-Rom history, alpha 0.1").
        A step = one taa_resolve_ex call on one synthetic frame; history ping-pongs, inputs rotate through 4 frame sets
        (4 x 166 MB, larger than the 126 MB L2), so every step streams its inputs from HBM.
N > 1 : one 7680x4320 frame sharded by row bands over N ranks (BASELINE configs[3]); every step each rank resolves its band and
        exchanges a history halo with its neighbours over NCCL (taa_star_b200/sharded.py). Launched with torchrun.
--impl reference : the CPU restatement of the reference shader (oracle/, all host threads) on the same config; rank 0 only.
Prints ONE JSON line (see the contract in the task description).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BYTES_PER_PX = {2: 44, 3: 48}  # colour 8 + depth 4 + velocity 8 + history 8 | history out 8 + result 8 (+ history depth 4)
FALLBACK_HBM_GBS = 6650.0      # /opt/skills/guides/B200_PROFILING.md


def ncu_traffic(cfg_id: int):
    """DRAM bytes per step of the resolve launches, from the committed `ncu --set full` capture (profiles/traffic.json)."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))[f"config{cfg_id}"]
        return float(t["dram_bytes_per_step"]), t["source"]
    except Exception:
        return None, None


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clock and throttle reasons through NVML while the timed regions run."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap", 0x80: "hw_power_brake_slowdown"}

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz, self.ok = [], set(), None, False
        self._stop = threading.Event()
        self._active = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            return
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()

    def _run(self):
        while not self._stop.is_set():
            if self._active.is_set():
                try:
                    self.samples.append(int(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
                    r = int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                    for bit, name in self.REASONS.items():
                        if r & bit:
                            self.reasons.add(name)
                except Exception:
                    pass
            time.sleep(0.002)

    def region(self, on: bool):
        (self._active.set if on else self._active.clear)()

    def result(self):
        self._stop.set()
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": int(statistics.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def cpu_impl(oracle_py):
    """The CPU arm: the reference's own shader text compiled by oracle/ref_build.py when that library exists, else the restatement."""
    if oracle_py.ref_available():
        return "ref", "reference", "oracle/_ref/libtaa_ref.so (the reference's taa.comp compiled through oracle/glsl_shim.h)"
    return "oracle", "port", "oracle/libtaa_oracle.so"


def cpu_oracle_throughput(cfg_id: int, width: int, height: int, budget_s: float = 12.0):
    """Times the CPU restatement (oracle/) on a band of rows of the bench workload. Returns (Mpx/s, cores, sample text)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import oracle_py
    from taa_star_b200 import configs
    from taa_star_b200.synth import SyntheticScene
    p = configs.config2_resolve() if cfg_id == 2 else configs.config3_full_chain()
    sc = SyntheticScene(width, height, with_aux=False)
    f0, f1 = sc.frame(8), sc.frame(9)
    ins = dict(color=f1.color.numpy(), depth=f1.depth.numpy(), velocity=f1.velocity.numpy())
    hist = f0.color.numpy().copy()
    u = configs.uniforms_for(p, f1.jitter_ndc)
    cores = oracle_py.max_threads()
    impl, kind, libname = cpu_impl(oracle_py)
    rows = min(height, 64)
    t0 = time.perf_counter()
    oracle_py.resolve(u, ins["color"], ins["depth"], ins["velocity"], hist, history_depth=f0.depth.numpy(), want=("history_out", "result"), rows=(height // 2, height // 2 + rows), impl=impl)
    per_row = (time.perf_counter() - t0) / rows
    rows = int(max(64, min(height, budget_s / 3 / max(per_row, 1e-9))))
    y0 = (height - rows) // 2
    times = []
    for _ in range(3):
        t0 = time.perf_counter()
        oracle_py.resolve(u, ins["color"], ins["depth"], ins["velocity"], hist, history_depth=f0.depth.numpy(), want=("history_out", "result"), rows=(y0, y0 + rows), impl=impl)
        times.append(time.perf_counter() - t0)
    best = min(times)
    return rows * width / best / 1e6, cores, kind, f"{rows} rows of one {width}x{height} frame (config {cfg_id}), best of 3, {cores} OpenMP threads, {libname}"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg_id = 2 if args.config == 5 else args.config  # (config 5 = 64 camera streams with the config 2 settings: the same per-pixel work)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py
    from taa_star_b200 import configs
    from taa_star_b200.synth import SyntheticScene
    p = configs.config2_resolve() if cfg_id == 2 else configs.config3_full_chain()
    sc = SyntheticScene(3840, 2160, with_aux=False)  # the band sample is taken from a 4K frame in both cases (same per-pixel work)
    f0, f1 = sc.frame(8), sc.frame(9)
    hist = f0.color.numpy().copy()
    u = configs.uniforms_for(p, f1.jitter_ndc)
    cores = oracle_py.max_threads()
    impl, kind, libname = cpu_impl(oracle_py)
    rows = 192  # bounded sample per step
    y0 = (2160 - rows) // 2

    def step():
        oracle_py.resolve(u, f1.color.numpy(), f1.depth.numpy(), f1.velocity.numpy(), hist, history_depth=f0.depth.numpy(),
                          want=("history_out", "result"), rows=(y0, y0 + rows), impl=impl)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    mpx = args.steps * rows * 3840 / dt / 1e6
    sample = f"each step = {rows} rows of a 3840x2160 frame (config {cfg_id}) through {libname} with {cores} OpenMP threads"
    if args.config == 5:
        workload = f"64 independent 1920x1080 camera streams, BASELINE configs[4] (config 2 settings), {64 // max(args.gpus, 1)} per GPU, no communication"
    elif args.gpus > 1:
        workload = f"7680x4320 TAA resolve sharded in {args.gpus} row bands, BASELINE configs[3] (config {cfg_id} settings)"
    else:
        workload = f"3840x2160 TAA resolve, BASELINE configs[{1 if cfg_id == 2 else 2}] (config {cfg_id})"
    line = {"impl": "reference", "metric": "resolved Mpixels/s", "value": round(mpx, 3), "unit": "Mpixels/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(dt / args.steps * 1e3, 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            # the workload string is the one the GPU arm prints for this N (the driver compares them); what the CPU arm executes per step is a
            # bounded sample of that workload's per-pixel work: `sample`
            "config": {"workload": workload, "sample": sample, "sampled_from": "3840x2160 frame with the same settings (the per-pixel work does not depend on the frame size)"},
            "cpu_baseline": {"value": round(mpx, 3), "unit": "Mpixels/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": round(mpx, 3), "unit": "Mpixels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def run_single(args):
    import torch
    from taa_star_b200 import abi, configs, host
    from taa_star_b200.synth import SyntheticScene
    W, H = args.width, args.height
    cfg_id = args.config
    torch.cuda.set_device(0)
    dev = torch.device("cuda:0")
    p = configs.config2_resolve() if cfg_id == 2 else configs.config3_full_chain()
    flags = abi.TAA_FLAG_EXACT if args.exact else 0
    NSETS = 4
    sc = SyntheticScene(W, H, device=dev, with_aux=False)
    frames = [sc.frame(n) for n in range(NSETS)]
    px = W * H
    peak, peak_src = measured_peak()
    stream = torch.cuda.Stream()
    sptr = stream.cuda_stream
    hist = [torch.zeros(H, W, 4, dtype=torch.float16, device=dev) for _ in range(2)]
    result = torch.zeros(H, W, 4, dtype=torch.float16, device=dev)
    final = torch.zeros(H, W, 4, dtype=torch.float16, device=dev)

    def varying_velocity():
        """A smoothly varying velocity field: no strip of the frame has uniform motion, no column shares its history u."""
        yy, xx = torch.meshgrid(torch.arange(H, device=dev, dtype=torch.float32), torch.arange(W, device=dev, dtype=torch.float32), indexing="ij")
        gain = 1.0 + 0.25 * torch.sin(xx * 0.011) * torch.cos(yy * 0.013)
        out = []
        for f in frames:
            v = f.velocity.clone()
            v[..., 0:2] = (f.velocity[..., 0:2].float() * gain[..., None]).half()
            out.append(v)
        return out

    def timed(params, velocities, steps, warmup, chain=None, with_hist_depth=False, bind_result=True):
        """Device time per step of `steps` back-to-back calls (CUDA events on the launching stream) after `warmup` calls, history ping-pong,
        inputs rotating through NSETS frame sets. chain = None: taa_resolve_ex; else taa_frame (the resolve and its follow-on passes)."""
        ctx = host.TaaContext((W, H), flags=flags)
        prepared = []
        for n in range(NSETS):
            f, fprev = frames[n], frames[(n - 1) % NSETS]
            for par in range(2):
                kw = dict(color=f.color, depth=f.depth, velocity=velocities[n], history_in=hist[par], history_out=hist[1 - par],
                          history_depth=fprev.depth if with_hist_depth else None)
                if bind_result:
                    kw["result"] = result
                prepared.append((ctx.images(**kw), configs.uniforms_for(params, f.jitter_ndc)))
        u0 = configs.uniforms_for(params, frames[0].jitter_ndc, reset_history=True)
        fin = ctx.image(final)

        def step(i, u=None):
            im, uu = prepared[(i % NSETS) * 2 + (i % 2)]
            if chain is None:
                ctx.resolve_prepared(im, u or uu, sptr)
            else:
                ctx.frame_prepared(im, u or uu, chain, fin, sptr)

        with torch.cuda.stream(stream):
            step(0, u0)
            for i in range(1, warmup + 1):
                step(i)
        torch.cuda.synchronize()
        l0 = ctx.launch_count
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        for i in range(steps):
            step(i + warmup + 1)
        ev1.record(stream)
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1) / steps
        launches = ctx.launch_count - l0
        fix = ctx.fixup_pixels()
        ctx.close()
        return ms, launches, fix

    def block(ms, launches, bytes_per_px, **extra):
        gbs = bytes_per_px * px / (ms * 1e-3) / 1e9
        d = {"ms_per_step": round(ms, 5), "Mpixels/s": round(px / (ms * 1e-3) / 1e6, 1), "bytes_per_px": bytes_per_px, "frac": round(gbs / peak, 4),
             "gpu_launches_per_step": round(launches, 2)}
        d.update(extra)
        return d

    vel_pan = [f.velocity for f in frames]
    assert args.motion == "pan" or args.kernel_only, "--motion varying is a kernel-only tuning aid; the bench line is measured on the pan of SURVEY 8d"
    vel = vel_pan if args.motion == "pan" else varying_velocity()

    # ---- the headline: `steps` resolves of the configuration the metric is quoted on ----
    clocks = ClockSampler(0)
    clocks.region(True)
    ms_per_step, launches, fixpx = timed(p, vel, args.steps, args.warmup, with_hist_depth=(cfg_id == 3))
    clocks.region(False)
    launches_per_step = launches / args.steps
    mpx_s = px / (ms_per_step * 1e-3) / 1e6
    achieved = BYTES_PER_PX[cfg_id] * px / (ms_per_step * 1e-3) / 1e9

    if args.kernel_only:  # tuning aid: device-timed kernel numbers only
        print(json.dumps({"kernel_only": True, "config": cfg_id, "ms_per_step": round(ms_per_step, 5), "Mpixels/s": round(mpx_s, 1),
                          "frac": round(achieved / peak, 4), "gpu_launches": int(launches), "fixup_pixels": fixpx,
                          "variant": os.environ.get("TAA_TUNED_VARIANT"), "motion": args.motion}), flush=True)
        return

    # ---- beside it (same run, same clocks): what the headline does not show ----
    side_steps = max(20, min(args.steps, 100))
    extra = {}
    if not args.exact and cfg_id == 2:
        # the same kernel on a smoothly varying velocity field (the pan of SURVEY 8d is the kindest input: bit-identical motion everywhere)
        ms_v, l_v, _ = timed(p, varying_velocity(), side_steps, args.warmup)
        extra["varying_motion"] = block(ms_v, l_v / side_steps, 44, note="config 2 on a smoothly varying velocity field: every row takes the general path")
        # BASELINE configs[2]: config 3 settings through taa_frame with CAS 0.5 + post-process (resolve + exact fix-up + [CAS + post] launch);
        # algorithmic bytes 48 B/px (a fused chain adds none)
        p3 = configs.config3_full_chain()
        ch = abi.taa_post_chain()
        ch.sharpener = 2
        ch.sharpen.sharpeningFactor = 0.5
        ch.cas = host.cas_setup(0.5, W, H)
        ch.postprocess = 1
        ch.pp = host.postprocess_default(W, H)
        ms_c, l_c, fix_c = timed(p3, vel_pan, side_steps, args.warmup, chain=ch, with_hist_depth=True, bind_result=False)
        extra["config3_full_chain"] = block(ms_c, l_c / side_steps, 48, fixup_pixels=fix_c,
                                            note="taa_frame: config 3 resolve (strip kernel; the exact pass decides the anti-ghosting predicate) + CAS + post-process in one follow-on launch")
        ms_r, l_r, _ = timed(p3, vel_pan, side_steps, args.warmup, with_hist_depth=True)
        extra["config3_resolve_only"] = block(ms_r, l_r / side_steps, 48)
        # BASELINE configs[0]: the reference's default settings (RGB min / max clamp, bilinear history, velocity for movers only + matrix
        # reprojection) on the exact arithmetic: the specialised exact kernel (bit-identical to the oracle), not a tuned one
        ms_d, l_d, _ = timed(configs.config1_defaults(), vel_pan, side_steps, args.warmup)
        extra["config1_defaults"] = block(ms_d, l_d / side_steps, 44, note="the reference's default settings on taa_resolve_defaults_kernel: exact arithmetic (all outputs "
                                          "bit-identical to the oracle), colour taps shared through a shared-memory tile, default switches folded at compile time")
        # the north-star target: fused TAA resolve + CAS — config 2 settings, CAS 0.5 + identity post-process in the resolve's epilogue
        ms_f, l_f, _ = timed(p, vel_pan, side_steps, args.warmup, chain=ch, bind_result=False)
        extra["fused_resolve_cas"] = block(ms_f, l_f / side_steps, 36, note="taa_frame: config 2 settings + CAS 0.5 + post-process, ONE launch (sharpening in the resolve's epilogue); "
                                           "36 B/px = 28 read + history_out 8 + final 8 (the unsharpened screen result is never written)")
    other = {}
    try:
        other = json.load(open(os.path.join(ROOT, "profiles", "other_configs.json")))
    except Exception:
        pass

    # ---- e2e: host buffers through the invokee (taa<CF>::render path), H2D + D2H inside the timed region ----
    t = host.Taa(3, flags=flags)
    t.set_sizes_for_host_frames((W, H), (W, H))
    for i in range(2):
        C.memmove(C.addressof(t.mParameters[i]), C.addressof(p), C.sizeof(p))
    s = t.settings
    s.jitter.mSampleDistribution = 2
    s.mPostProcessEnabled = 1 if cfg_id == 3 else 0
    s.mSharpener = 2 if cfg_id == 3 else 0
    s.mResetHistoryOnChange = 0
    lib = abi.load_library()

    def pinned(tensor):
        n = tensor.numel() * tensor.element_size()
        ptr = lib.taa_host_alloc(n)
        assert ptr, "taa_host_alloc failed"
        ht = torch.frombuffer((C.c_ubyte * n).from_address(ptr), dtype=torch.uint8).view(tensor.dtype).view(tensor.shape)
        ht.copy_(tensor.cpu())
        return ht
    hsets = [(pinned(f.color), pinned(f.depth), pinned(f.velocity)) for f in frames]
    houts = [pinned(result) for _ in range(3)]
    e2e_steps = max(8, min(args.steps, 48))
    for n in range(6):  # warm-up of the pipeline
        f = frames[n % NSETS]
        c, d, v = hsets[n % NSETS]
        t.frame_host(n, c, d, v, f.view, f.proj, houts[n % 3])
    for n in range(3, 6):
        t.wait(n)
    torch.cuda.synchronize()
    clocks.region(True)
    e2e_launch0 = t.launch_count
    e2e_h2d0 = t.h2d_bytes
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        n = 6 + k
        f = frames[n % NSETS]
        c, d, v = hsets[n % NSETS]
        t.frame_host(n, c, d, v, f.view, f.proj, houts[n % 3])
    for n in range(6 + e2e_steps - 3, 6 + e2e_steps):
        t.wait(n)
    torch.cuda.synchronize()
    e2e_dt = time.perf_counter() - t0
    clocks.region(False)
    e2e_mpx = e2e_steps * px / e2e_dt / 1e6
    h2d = (t.h2d_bytes - e2e_h2d0) // e2e_steps  # counted by the library from the copies it issued (colour + velocity; + depth when a kernel reads it)
    d2h = px * 8
    checksum = float(houts[(6 + e2e_steps - 1) % 3][::97, ::89, :3].float().mean())
    assert 0.05 < checksum < 0.95, f"implausible result mean {checksum}"

    cpu_mpx, cores, cpu_kind, sample = cpu_oracle_throughput(cfg_id, W, H)
    traffic, traffic_src = ncu_traffic(cfg_id) if (not args.exact and (W, H) == (3840, 2160)) else (None, None)
    if args.exact:
        arithmetic, kernel = "exact general kernel", "taa_resolve_generic_kernel"
    else:
        kernel = "taa_resolve_stream_kernel" if cfg_id == 2 else "taa_resolve_strip_kernel"
        arithmetic = ("tuned kernel alone (no mask bound: nothing for the exact fix-up pass to decide)" if launches_per_step < 1.5
                      else "tuned kernel + exact fix-up pass")
    line = {
        "metric": "resolved Mpixels/s", "value": round(mpx_s, 1), "unit": "Mpixels/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(ms_per_step, 5), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "fps": round(1e3 / ms_per_step, 1),
        "config": {"workload": f"{W}x{H} TAA resolve, BASELINE configs[{1 if cfg_id == 2 else 2}] (config {cfg_id})", "arithmetic": arithmetic,
                   "l2": f"inputs larger than L2: {NSETS} frame sets rotated ({NSETS} x {px * 20 / 1e6:.0f} MB), history ping-pong",
                   "outputs": "history_out + result", "motion": "analytic pan (+3.0, +0.5) px/frame + one mover (SURVEY 8d)"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4), "traffic": traffic,
                     "traffic_source": traffic_src, "algorithmic_bytes": BYTES_PER_PX[cfg_id] * px,
                     "peak_source": peak_src, "bytes_per_px": BYTES_PER_PX[cfg_id], "kernel": kernel},
        "cpu_baseline": {"value": round(cpu_mpx, 3), "unit": "Mpixels/s", "cores": cores, "kind": cpu_kind, "sample": sample},
        "e2e": {"value": round(e2e_mpx, 1), "unit": "Mpixels/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                "fps": round(e2e_steps / e2e_dt, 1), "path": "taa_invokee_frame_host: pinned host G-buffer -> H2D -> render() -> D2H of the final image, 3 frames in flight",
                "gpu_launches": int(t.launch_count - e2e_launch0),
                "bound": f"PCIe: {h2d / 1e6:.0f} MB up + {d2h / 1e6:.0f} MB down per frame at ~55-60 GB/s each way; the kernel is ~4 % of the frame time",
                "inputs_uploaded": "colour + velocity" + (" + depth" if h2d >= px * 20 else " (depth stays on the host: no kernel of this configuration reads it)")},
        "clocks": clocks.result(),
    }
    line.update(extra)
    if other:
        line["other_configs"] = other
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 5], help="2: 4K resolve (the bench line); 3: config 3 settings; 5: 64 independent 1080p camera streams (BASELINE configs[4])")
    ap.add_argument("--motion", default="pan", choices=["pan", "varying"], help="varying: perturb the synthetic velocity field (kernel-only tuning aid; the bench line is 'pan', SURVEY 8d)")
    ap.add_argument("--width", type=int, default=3840)
    ap.add_argument("--height", type=int, default=2160)
    ap.add_argument("--kernel-only", action="store_true", help="skip the e2e and CPU legs (tuning aid; not a bench line)")
    ap.add_argument("--exact", action="store_true", help="TAA_FLAG_EXACT: force the exact general kernel")
    ap.add_argument("--replicate", action="store_true", help="--gpus N: all-gather the whole history every frame instead of exchanging halos (unbounded motion)")
    ap.add_argument("--no-verify", action="store_true", help="--gpus N: skip the whole-frame reference run (timing of the same frame on one GPU, bit-for-bit check of the bands)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        # torchrun exports OMP_NUM_THREADS=1 when the user has not set it; the CPU arm is to use every host thread it can get.
        # (Set before anything loads an OpenMP runtime: torch and the oracle libraries are imported inside run_reference.)
        if os.environ.get("OMP_NUM_THREADS") == "1" and "TORCHELASTIC_RUN_ID" in os.environ:
            os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
        if args.steps > 20:
            args.steps = 20  # bounded: each step is ~1 s of CPU work
        return run_reference(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.config == 5 and args.impl != "reference":
        from taa_star_b200 import streams
        return streams.bench_main(args, ClockSampler, measured_peak, BYTES_PER_PX)
    if args.gpus > 1 or world > 1:
        from taa_star_b200 import sharded
        return sharded.bench_main(args, ClockSampler, measured_peak, BYTES_PER_PX)
    return run_single(args)


if __name__ == "__main__":
    main()
