"""taa_frame on a 4K frame with config 2 settings + [sharpen | CAS] + identity post-process: the one-launch fused chain (TAA_FUSED_CHAIN=0: resolve + streaming\nsharpening pass); SHARP=1|2 picks the sharpener, TAA_STREAM_EPI_MINB the register budget of the epilogue variant (tuning / ncu aid)."""
import os, sys, torch
sys.path.insert(0, os.getcwd())
from taa_star_b200 import abi, configs, host
from taa_star_b200.synth import SyntheticScene
W, H = 3840, 2160
dev = torch.device("cuda:0")
sc = SyntheticScene(W, H, device=dev, with_aux=False)
frames = [sc.frame(n) for n in range(4)]
p = configs.config2_resolve()
ctx = host.TaaContext((W, H))
hist = [torch.zeros(H, W, 4, dtype=torch.float16, device=dev) for _ in range(2)]
final = torch.zeros(H, W, 4, dtype=torch.float16, device=dev)
ch = abi.taa_post_chain(); ch.sharpener = int(os.environ.get("SHARP", "2")); ch.sharpen.sharpeningFactor = 0.5; ch.cas = host.cas_setup(0.5, W, H); ch.postprocess = 1; ch.pp = host.postprocess_default(W, H)
stream = torch.cuda.Stream()
prep = []
for n in range(4):
    for par in range(2):
        f = frames[n]
        prep.append((ctx.images(color=f.color, depth=f.depth, velocity=f.velocity, history_in=hist[par], history_out=hist[1 - par]), configs.uniforms_for(p, f.jitter_ndc)))
fin = ctx.image(final)
with torch.cuda.stream(stream):
    for i in range(8):
        im, u = prep[(i % 4) * 2 + (i % 2)]; ctx.frame_prepared(im, u, ch, fin, stream.cuda_stream)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
for i in range(100):
    im, u = prep[(i % 4) * 2 + (i % 2)]; ctx.frame_prepared(im, u, ch, fin, stream.cuda_stream)
e1.record(stream); torch.cuda.synchronize()
print("fused chain ms", e0.elapsed_time(e1) / 100, "minb", os.environ.get("TAA_STREAM_EPI_MINB"), "sharp", os.environ.get("SHARP", "2"))
