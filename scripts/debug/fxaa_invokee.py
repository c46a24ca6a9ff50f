import ctypes as C, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")]
import numpy as np, torch, oracle_py as oracle
from common import np_inputs
from taa_star_b200 import abi, configs, host
from taa_star_b200.synth import SyntheticScene
W, H, CF = 160, 96, 3
sc = SyntheticScene(W, H, pan_px=(1.5, 0.75))
frames = [sc.frame(n) for n in range(2)]
t = host.Taa(CF, flags=abi.TAA_FLAG_EXACT)
p = configs.config3_full_chain()
p.mRayTraceAugment = 1
p.mRayTraceAugmentFlags = abi.TAA_RTFLAG_OUT | abi.TAA_RTFLAG_DPT | abi.TAA_RTFLAG_LUM | abi.TAA_RTFLAG_MID | abi.TAA_RTFLAG_CNT | abi.TAA_RTFLAG_FXA
p.mRayTraceAugment_WDpt, p.mRayTraceAugment_WLum = 4.0, 1.5
print("hist count default", p.mRayTraceHistoryCount)
for i in range(2):
    C.memmove(C.addressof(t.mParameters[i]), C.addressof(p), C.sizeof(p))
s = t.settings
s.jitter.mSampleDistribution = 2
s.mSharpener, s.mSharpenFactor, s.mPostProcessEnabled = 0, 0.5, 0
slots = [dict(color=torch.empty_like(frames[0].color.cuda()), depth=torch.empty_like(frames[0].depth.cuda()), velocity=torch.empty_like(frames[0].velocity.cuda()),
              matid=torch.empty_like(frames[0].matid.cuda()), uvnrm=torch.empty_like(frames[0].uvnrm.cuda())) for _ in range(CF)]
t.set_source_image_views((W, H), [x["color"] for x in slots], [x["depth"] for x in slots], [x["uvnrm"] for x in slots], [x["velocity"] for x in slots],
                         [x["matid"] for x in slots])
for n, f in enumerate(frames):
    i, last = n % CF, (n + CF - 1) % CF
    for k in slots[i]:
        slots[i][k].copy_(getattr(f, k).cuda())
    t.get_jittered_projection_matrix(f.proj, n)
    t.save_history_proj_matrix(f.proj, n)
    t.update(n, f.view)
    ptr = t.render(n)
    torch.cuda.synchronize()
    got = t.image_by_ptr(ptr).cpu().numpy()
n = 1; f = frames[1]
(jx, jy), npat = host.jitter_offset_for_frame(n, W, H, sample_distribution=2)
pp = abi.TaaParameters.from_buffer_copy(p); pp.mRayTraceHistoryCount = npat
u = configs.uniforms_for(pp, (jx, jy))
for k in range(4): u.mSinTime[k] = 0.0
m = lambda a: (C.c_float * 16)(*a)
abi.load_library().taa_reprojection_matrices(m(f.proj), m(f.view), m(frames[0].proj), m(frames[0].view), u.mInverseViewProjMatrix, u.mHistoryViewProjMatrix)
ins = np_inputs(f)
r = oracle.resolve(u, ins["color"], ins["depth"], ins["velocity"], np.zeros((H, W, 4), np.float16), history_depth=frames[0].depth.numpy(), matid=ins["matid"],
                   prev_matid=frames[0].matid.numpy(), uvnrm=ins["uvnrm"], prev_segmask=np.zeros((H, W), np.uint32), want=("history_out", "result", "segmask"))
U = t.uniforms
ub, vb = bytes(U), bytes(u)
print("uniform bytes differ at", [k for k in range(544) if ub[k] != vb[k]][:20])
seg = t.image(abi.TAA_IMG_SEGMASK, 1, torch.int32).cpu().numpy().view(np.uint32)
print("seg equal:", (seg == r["segmask"]).mean(), np.bincount(seg.ravel() & 3, minlength=4), np.bincount(r["segmask"].ravel() & 3, minlength=4))
res = t.image(abi.TAA_IMG_RESULT, 1).cpu().numpy()
print("result equal:", (res.view(np.uint16) == r["result"].view(np.uint16)).mean())
img = oracle.fxaa(oracle.fxaa_prepare(r["result"]), r["segmask"], host.fxaa_default(W, H), gather4=True)
d = (img.view(np.uint16) != got.view(np.uint16)).any(-1)
print("final differs px", d.sum(), "rows", np.unique(np.argwhere(d)[:, 0])[:10], "cols", np.unique(np.argwhere(d)[:, 1])[:10])
print("got==result?", (got.view(np.uint16) == res.view(np.uint16)).all(-1).mean(), " final ptr is TEMP0:", ptr == t._lib.taa_invokee_image(t._h, abi.TAA_IMG_TEMP0, 1))
