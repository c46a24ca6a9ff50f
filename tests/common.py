"""Shared helpers of the parity tests: run the same frame through the oracle (CPU) and the C-ABI (GPU) and compare."""
from __future__ import annotations

import numpy as np
import torch

from taa_star_b200 import abi, configs
from taa_star_b200.synth import SyntheticScene

TOL_ABS = 2.0 ** -10  # north_star: max per-channel abs error on rgba16f outputs
PSNR_MIN = 60.0      # north_star: after 64 accumulated frames


def np_inputs(f):
    return dict(color=f.color.cpu().numpy(), depth=f.depth.cpu().numpy(), velocity=f.velocity.cpu().numpy(),
                matid=f.matid.cpu().numpy() if f.matid.numel() else None, uvnrm=f.uvnrm.cpu().numpy() if f.uvnrm.numel() else None)


def random_history(h, w, seed, alpha_binary=True):
    rng = np.random.default_rng(seed)
    hist = rng.random((h, w, 4), dtype=np.float32)
    if alpha_binary:
        hist[..., 3] = (rng.random((h, w)) > 0.7).astype(np.float32)
    return hist.astype(np.float16)


def bits16(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint16)


def same_f16(a, b):
    """Bitwise equality of fp16 images, with -0 == +0 and NaN == NaN."""
    ua, ub = bits16(a).copy(), bits16(b).copy()
    ua[ua == 0x8000] = 0
    ub[ub == 0x8000] = 0
    nan_a = (ua & 0x7fff) > 0x7c00
    nan_b = (ub & 0x7fff) > 0x7c00
    return (ua == ub) | (nan_a & nan_b)


def mismatch_report(name, a, b):
    eq = same_f16(a, b)
    n = int((~eq).sum())
    if n == 0:
        return None
    idx = np.argwhere(~eq)[:5]
    fa, fb = a.astype(np.float32), b.astype(np.float32)
    d = np.abs(fa - fb)
    return f"{name}: {n} of {eq.size} fp16 values differ; max |d| = {np.nanmax(d):.3e}; first at {idx.tolist()}: " + \
        ", ".join(f"{fa[tuple(i)]!r} vs {fb[tuple(i)]!r}" for i in idx)


def psnr(a, b, peak=1.0):
    mse = float(np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2))
    return 200.0 if mse == 0 else 10.0 * np.log10(peak * peak / mse)


def to_dev(a, dtype=None):
    if a is None:
        return None
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.view(dtype)
    return t.cuda()


def run_gpu_resolve(ctx, u, ins, hist_in, hist_depth=None, prev_matid=None, prev_segmask=None, out_size=None,
                    want=("history_out", "result", "mask")):
    """One taa_resolve_ex through the C-ABI; returns numpy outputs like oracle_py.resolve."""
    in_h, in_w = ins["depth"].shape
    out_w, out_h = out_size if out_size else (in_w, in_h)
    outs = {}
    for name in want:
        if name in ("history_out", "result", "debug"):
            outs[name] = torch.full((out_h, out_w, 4), float("nan"), dtype=torch.float16, device="cuda")
        else:
            outs[name] = torch.full((out_h, out_w), -1, dtype=torch.int32, device="cuda")
    kw = dict(color=to_dev(ins["color"]), depth=to_dev(ins["depth"]), velocity=to_dev(ins["velocity"]), history_in=to_dev(hist_in),
              history_depth=to_dev(hist_depth), matid=to_dev(ins.get("matid")), uvnrm=to_dev(ins.get("uvnrm")),
              prev_matid=to_dev(prev_matid), prev_segmask=to_dev(prev_segmask))
    kw.update(outs)
    ctx.resolve(u, **kw)
    torch.cuda.synchronize()
    res = {}
    for k, v in outs.items():
        a = v.cpu().numpy()
        res[k] = a.view(np.uint32) if a.dtype == np.int32 else a
    return res
