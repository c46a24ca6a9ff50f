"""Per-unit timeline of the streaming kernel (needs a build with EXTRA_NVCC_FLAGS=-DTAA_STREAM_TRACE): which units are slow, when SMs run dry."""
import ctypes as C, os, sys, json
import numpy as np, torch
sys.path.insert(0, os.getcwd())
from taa_star_b200 import abi, configs, host
from taa_star_b200.synth import SyntheticScene
W, H = int(os.environ.get("W", 3840)), int(os.environ.get("H", 2160))
dev = torch.device("cuda:0")
mover = os.environ.get("MOVER", "1") == "1"
sc = SyntheticScene(W, H, device=dev, with_aux=False)
if not mover:
    sc.mhalf = (-1.0, -1.0)
frames = [sc.frame(n) for n in range(4)]
p = configs.config2_resolve()
ctx = host.TaaContext((W, H))
hist = [torch.zeros(H, W, 4, dtype=torch.float16, device=dev) for _ in range(2)]
result = torch.zeros(H, W, 4, dtype=torch.float16, device=dev)
stream = torch.cuda.Stream()
prep = []
for n in range(4):
    for par in range(2):
        f = frames[n]
        prep.append((ctx.images(color=f.color, depth=f.depth, velocity=f.velocity, history_in=hist[par], history_out=hist[1 - par], result=result), configs.uniforms_for(p, f.jitter_ndc)))
def run(n):
    for i in range(n):
        im, u = prep[(i % 4) * 2 + (i % 2)]; ctx.resolve_prepared(im, u, stream.cuda_stream)
run(8); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream); run(100); e1.record(stream); torch.cuda.synchronize()
print("ms per frame", e0.elapsed_time(e1) / 100, "mover", mover, "R", os.environ.get("TAA_STREAM_R"))
lib = abi.load_library()
if hasattr(lib, "taa_debug_stream_trace"):
    buf = np.zeros(4 * 16384, dtype=np.uint64)
    lib.taa_debug_stream_trace.argtypes = [C.c_void_p, C.c_size_t]
    lib.taa_debug_stream_trace(buf.ctypes.data, buf.nbytes)
    t = buf.reshape(-1, 4)
    t = t[t[:, 0] > 0]
    t0, t1 = t[:, 0].astype(np.int64), t[:, 1].astype(np.int64)
    sm = (t[:, 2] & 0xffff).astype(int); gen = ((t[:, 2] >> 16) & 0xffff).astype(int)
    base = t0.min(); t0 -= base; t1 -= base
    dur = t1 - t0
    print("units", len(t), "span us", t1.max() / 1e3, "unit us: median %.1f p90 %.1f max %.1f" % (np.median(dur) / 1e3, np.percentile(dur, 90) / 1e3, dur.max() / 1e3))
    g = gen != 0xffff
    print("units that left the uniform rows:", int(g.sum()), "their median us %.1f max %.1f; uniform units median %.1f max %.1f" % (np.median(dur[g]) / 1e3 if g.any() else 0, dur[g].max() / 1e3 if g.any() else 0, np.median(dur[~g]) / 1e3, dur[~g].max() / 1e3))
    # per SM: first start, last end, busy warps over time
    ends = np.array([t1[sm == s].max() for s in np.unique(sm)]); starts = np.array([t0[sm == s].min() for s in np.unique(sm)])
    print("SMs", len(ends), "last-end us: min %.1f median %.1f max %.1f; first-start us: max %.1f" % (ends.min() / 1e3, np.median(ends) / 1e3, ends.max() / 1e3, starts.max() / 1e3))
    # resident-warp curve
    T = int(t1.max()); grid = np.linspace(0, T, 41)
    occ = [(int(((t0 <= x) & (t1 > x)).sum())) for x in grid]
    print("resident warps over time (40 bins):", occ)
    late = np.argsort(-t1)[:12]
    print("last units to finish (end us, dur us, Y0, strip, general-from-row):", [(round(t1[i] / 1e3, 1), round(dur[i] / 1e3, 1), int(t[i, 3] & 0xffff), int(t[i, 3] >> 16), int(gen[i])) for i in late])
    # start-time distribution of waves
    print("start times us (sorted, every 400th):", [round(x / 1e3, 1) for x in np.sort(t0)[::400]])
