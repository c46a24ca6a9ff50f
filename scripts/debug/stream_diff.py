"""Where does the default (streaming) resolve differ from the oracle? Prints per-case maps of the bad pixels (tuning / debugging aid)."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_py
from common import np_inputs, random_history, run_gpu_resolve
from taa_star_b200 import abi, configs, host
from taa_star_b200.synth import SyntheticScene


def with_params(base, **kw):
    p = abi.TaaParameters.from_buffer_copy(base)
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def report(name, ref, got):
    for k in ("history_out", "result"):
        d = np.abs(ref[k].astype(np.float32) - got[k].astype(np.float32)).max(axis=-1)
        bad = d > 2.0 ** -10
        print(f"{name} {k}: max |d| = {d.max():.4f}, bad pixels = {int(bad.sum())}")
        if bad.any():
            ys, xs = np.nonzero(bad)
            print(f"   rows {ys.min()}..{ys.max()}, cols {xs.min()}..{xs.max()}")
            rows = np.nonzero(bad.any(axis=1))[0]
            print("   bad rows:", rows.tolist()[:60])
            cols = np.nonzero(bad.any(axis=0))[0]
            print("   bad cols:", cols.tolist()[:80])
    if "mask" in ref:
        bad = ref['mask'] != got['mask']
        print(f"{name} mask: {int(bad.sum())} differ")
        if bad.any():
            ys, xs = np.nonzero(bad)
            for y, x in list(zip(ys, xs))[:4]:
                print(f"   ({x},{y}): ref mask {ref['mask'][y, x]:#x} got {got['mask'][y, x]:#x}; ref hist {ref['history_out'][y, x]} got {got['history_out'][y, x]}")


W, H = 256, 144
sc = SyntheticScene(W, H, pan_px=(5.25, -2.5))
f0, f1 = sc.frame(2), sc.frame(3)
for sw in (dict(mRejectOutside=1), dict(mDepthCulling=1)):
    u = configs.uniforms_for(with_params(configs.config2_resolve(), **sw), f1.jitter_ndc)
    ins, hist = np_inputs(f1), random_history(H, W, 11)
    want = ("history_out", "result", "mask")
    ref = oracle_py.resolve(u, ins["color"], ins["depth"], ins["velocity"], hist, history_depth=f0.depth.numpy(), want=want)
    ctx = host.TaaContext((W, H))
    for rep in range(3):
        got = run_gpu_resolve(ctx, u, ins, hist, hist_depth=f0.depth.numpy(), want=want)
        report(f"{sw} call {rep}", ref, got)
    ctx.close()

# sequence on one context (the partition adapts from frame to frame)
w, h = 192, 108
sc = SyntheticScene(w, h)
p = configs.config2_resolve()
ctx = host.TaaContext((w, h))
hist = np.zeros((h, w, 4), np.float16)
for n in range(12):
    f = sc.frame(n)
    ins = np_inputs(f)
    u = configs.uniforms_for(p, f.jitter_ndc, reset_history=(n == 0))
    ref = oracle_py.resolve(u, ins["color"], ins["depth"], ins["velocity"], hist, history_depth=ins["depth"], want=("history_out", "result"))
    try:
        got = run_gpu_resolve(ctx, u, ins, hist, hist_depth=ins["depth"], want=("history_out", "result"))
    except Exception as e:
        print("frame", n, "FAILED:", str(e)[:200])
        break
    report(f"seq frame {n}", ref, got)
    hist = ref["history_out"]
