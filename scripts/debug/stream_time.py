"""Per-frame device times of the default resolve on the 4K bench scene, with and without the mover (tuning aid)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from taa_star_b200 import configs, host
from taa_star_b200.synth import SyntheticScene
W, H = 3840, 2160
dev = torch.device("cuda:0")
for mover in (True, False):
    sc = SyntheticScene(W, H, device=dev, with_aux=False)
    if not mover:
        sc.mhalf = (-1.0, -1.0)
    frames = [sc.frame(n) for n in range(4)]
    p = configs.config2_resolve()
    ctx = host.TaaContext((W, H))
    hist = [torch.zeros(H, W, 4, dtype=torch.float16, device=dev) for _ in range(2)]
    result = torch.zeros(H, W, 4, dtype=torch.float16, device=dev)
    stream = torch.cuda.Stream()
    prepared = []
    for n in range(4):
        for par in range(2):
            f = frames[n]
            im = ctx.images(color=f.color, depth=f.depth, velocity=f.velocity, history_in=hist[par], history_out=hist[1 - par], result=result)
            prepared.append((im, configs.uniforms_for(p, f.jitter_ndc)))
    times = []
    with torch.cuda.stream(stream):
        for i in range(40):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            im, u = prepared[(i % 4) * 2 + (i % 2)]
            e0.record(stream)
            ctx.resolve_prepared(im, u, stream.cuda_stream)
            e1.record(stream)
            times.append((e0, e1))
    torch.cuda.synchronize()
    ms = [a.elapsed_time(b) for a, b in times]
    print("mover" if mover else "no mover", " ".join(f"{t * 1000:.0f}" for t in ms))
