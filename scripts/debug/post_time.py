"""The follow-on passes on a 4K frame (for an ncu capture): unfused [CAS + post-process] streaming pass, sharpen, the general post-process with the zoom box,
FXAA prepare + FXAA on a masked frame."""
import os, sys, torch
sys.path.insert(0, os.getcwd())
from taa_star_b200 import abi, configs, host
from taa_star_b200.synth import SyntheticScene
W, H = 3840, 2160
dev = torch.device("cuda:0")
sc = SyntheticScene(W, H, device=dev, with_aux=False)
f = sc.frame(3)
ctx = host.TaaContext((W, H))
src = f.color.clone()
dst = torch.zeros_like(src); dst2 = torch.zeros_like(src)
seg = ((torch.arange(H, device=dev)[:, None] // 16 + torch.arange(W, device=dev)[None, :] // 16) % 3 == 0).to(torch.int32).contiguous()  # a third of the tiles marked for FXAA
cas = host.cas_setup(0.5, W, H)
pp = host.postprocess_default(W, H)
ppz = host.postprocess_default(W, H); ppz.zoom = 1; ppz.showZoomBox = 1
ch = abi.taa_post_chain(); ch.sharpener = 2; ch.cas = cas; ch.postprocess = 1; ch.pp = pp
def once():
    ctx.sharpen_cas(src, dst, cas)
    ctx.sharpen(src, dst, 0.5)
    ctx.post_process(src, None, dst, ppz)
    ctx.fxaa_prepare(src, dst2)
    ctx.fxaa(dst2, seg, dst, host.fxaa_default(W, H))
    ctx.fxaa(src, seg, dst, host.fxaa_default(W, H), fused=True)
for _ in range(3): once()
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
ev[0].record(); 
for _ in range(20): once()
ev[1].record(); torch.cuda.synchronize()
print("six follow-on launches per iteration: %.4f ms per iteration" % (ev[0].elapsed_time(ev[1]) / 20))
