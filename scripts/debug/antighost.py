import sys, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, ROOT + "/oracle"); sys.path.insert(0, ROOT + "/tests")
import oracle_py
from common import np_inputs, random_history, run_gpu_resolve
from taa_star_b200 import abi, configs, host
from taa_star_b200.synth import SyntheticScene
W, H = 256, 144
sc = SyntheticScene(W, H, pan_px=(5.25, -2.5))
f0, f1 = sc.frame(2), sc.frame(3)
p = configs.config2_resolve(); p.mDynamicAntiGhosting = 1
u = configs.uniforms_for(p, f1.jitter_ndc)
ins = np_inputs(f1); hist = random_history(H, W, 11)
ref = oracle_py.resolve(u, ins["color"], ins["depth"], ins["velocity"], hist, history_depth=f0.depth.numpy(), want=("history_out", "result", "mask"))
for flags in (0, abi.TAA_FLAG_EXACT):
    ctx = host.TaaContext((W, H), flags=flags)
    got = run_gpu_resolve(ctx, u, ins, hist, hist_depth=f0.depth.numpy())
    d = np.abs(ref["history_out"].astype(np.float32) - got["history_out"].astype(np.float32)).max(axis=2)
    bad = np.argwhere(d > 2**-10)
    mm = np.argwhere(ref["mask"] != got["mask"])
    print("flags", flags, "colour bad", len(bad), "mask mismatches", len(mm), "fix", ctx.fixup_pixels())
    for (y, x) in bad[:12]:
        print(" px", x, y, "ref mask", ref["mask"][y, x], "got", got["mask"][y, x], "ref", ref["history_out"][y, x], "got", got["history_out"][y, x],
              "w around", ins["velocity"][max(0,y-2):y+3, max(0,x-2):x+3, 3].astype(np.float32).ravel().tolist())
