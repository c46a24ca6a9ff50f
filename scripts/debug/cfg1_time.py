"""The reference default settings (BASELINE configs[0]) on a 4K frame: device time per frame of the specialised exact kernel (tuning / ncu aid)."""
import os, sys, torch
sys.path.insert(0, os.getcwd())
from taa_star_b200 import abi, configs, host
from taa_star_b200.synth import SyntheticScene
W, H = 3840, 2160
dev = torch.device("cuda:0")
sc = SyntheticScene(W, H, device=dev, with_aux=False)
frames = [sc.frame(n) for n in range(4)]
p = configs.config1_defaults()
ctx = host.TaaContext((W, H))
hist = [torch.zeros(H, W, 4, dtype=torch.float16, device=dev) for _ in range(2)]
res = torch.zeros(H, W, 4, dtype=torch.float16, device=dev)
stream = torch.cuda.Stream()
prep = []
for n in range(4):
    for par in range(2):
        f = frames[n]
        prep.append((ctx.images(color=f.color, depth=f.depth, velocity=f.velocity, history_in=hist[par], history_out=hist[1 - par], result=res), configs.uniforms_for(p, f.jitter_ndc)))
def run(n):
    for i in range(n):
        im, u = prep[(i % 4) * 2 + (i % 2)]; ctx.resolve_prepared(im, u, stream.cuda_stream)
run(8); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream); run(50); e1.record(stream); torch.cuda.synchronize()
print("config 1 ms per frame", e0.elapsed_time(e1) / 50)
