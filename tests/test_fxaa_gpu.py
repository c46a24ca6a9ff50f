"""GPU parity of the FXAA branch of render() (taa.hpp:1061-1107): antialias_fxaa_prepare.comp, antialias_fxaa.comp (FxaaPixelShader,
Fxaa3_11_mod.h:884-1243, preset 12), the fused single launch, and taa_frame with a segmentation mask that marks pixels for FXAA.
Everything is compared bit for bit with the CPU oracle (which tests/test_oracle_vs_ref.py pins to the reference's shader text)."""
import ctypes as C

import numpy as np
import pytest
import torch

from common import mismatch_report, np_inputs, random_history, to_dev
from taa_star_b200 import abi, configs, host
from taa_star_b200.synth import SyntheticScene
from test_oracle_vs_ref import fxaa_test_image

pytestmark = pytest.mark.gpu


def gpu_out(h, w):
    return torch.full((h, w, 4), float("nan"), dtype=torch.float16, device="cuda")


def seg_masks(h, w, seed):
    rng = np.random.default_rng(seed)
    yield np.full((h, w), 1 | (5 << 16), np.uint32)
    yield rng.integers(0, 4, (h, w)).astype(np.uint32) | (rng.integers(0, 9, (h, w)).astype(np.uint32) << 16)
    yield np.zeros((h, w), np.uint32)


@pytest.mark.parametrize("h,w", [(96, 160), (37, 53), (1, 1), (2, 64), (270, 481)])
def test_fxaa_kernels(oracle, h, w):
    ctx = host.TaaContext((w, h))
    src = fxaa_test_image(h, w, 40 + h)
    prep_ref = oracle.fxaa_prepare(src)
    prep = gpu_out(h, w)
    ctx.fxaa_prepare(to_dev(src), prep)
    assert mismatch_report("fxaa_prepare", prep_ref, prep.cpu().numpy()) is None
    pcs = [host.fxaa_default(w, h)]
    pc2 = host.fxaa_default(w, h)
    pc2.fxaaQualitySubpix, pc2.fxaaQualityEdgeThreshold, pc2.fxaaQualityEdgeThresholdMin = 1.0, 0.063, 0.0
    pcs.append(pc2)
    for seg in seg_masks(h, w, h * w):
        for pc in pcs:
            ref = oracle.fxaa(prep_ref, seg, pc, gather4=True)
            two, one = gpu_out(h, w), gpu_out(h, w)
            ctx.fxaa(prep, to_dev(seg.view(np.int32)), two, pc)
            ctx.fxaa(to_dev(src), to_dev(seg.view(np.int32)), one, pc, fused=True)
            torch.cuda.synchronize()
            r = mismatch_report("fxaa (prepared)", ref, two.cpu().numpy())
            assert r is None, r
            r = mismatch_report("fxaa (fused)", ref, one.cpu().numpy())
            assert r is None, r


def test_fxaa_default_push_constants():
    pc = host.fxaa_default(3840, 2160)
    assert pc.fxaaQualityRcpFrame[0] == np.float32(1.0) / np.float32(3840) and pc.fxaaQualityRcpFrame[1] == np.float32(1.0) / np.float32(2160)
    assert (pc.fxaaQualitySubpix, pc.fxaaQualityEdgeThreshold, pc.fxaaQualityEdgeThresholdMin) == (np.float32(0.75), np.float32(0.116), np.float32(0.0833))


@pytest.mark.parametrize("sharpener,post", [(0, 0), (2, 1), (1, 0), (0, 1)])
def test_taa_frame_with_fxaa(oracle, sharpener, post):
    """taa.comp writes the seg-mask (mRayTraceAugment), FXAA consumes it, then the sharpener / post-process follow (taa.hpp:1029-1161)."""
    W, H = 200, 112
    sc = SyntheticScene(W, H, pan_px=(2.5, -1.25))
    f0, f1 = sc.frame(3), sc.frame(4)
    ins, hist = np_inputs(f1), f0.color.numpy().copy()
    p = configs.config3_full_chain()
    p.mRayTraceAugment = 1
    # seg-mask value 1 (FXAA) comes from the off-screen-history test (OUT) or the debug border (FXD: 100 px, i.e. every pixel here)
    p.mRayTraceAugmentFlags = abi.TAA_RTFLAG_DPT | abi.TAA_RTFLAG_LUM | abi.TAA_RTFLAG_MID | abi.TAA_RTFLAG_FXA | (abi.TAA_RTFLAG_OUT if (sharpener or post) else abi.TAA_RTFLAG_FXD)
    p.mRayTraceAugment_WDpt, p.mRayTraceAugment_WLum = 4.0, 1.5
    p.mRayTraceHistoryCount = 8
    u = configs.uniforms_for(p, f1.jitter_ndc)
    r = oracle.resolve(u, ins["color"], ins["depth"], ins["velocity"], hist, history_depth=f0.depth.numpy(), matid=ins["matid"], prev_matid=f0.matid.numpy(),
                       uvnrm=ins["uvnrm"], want=("history_out", "result", "segmask"))
    assert ((r["segmask"] & 3) == 1).any(), "the test scene must mark some pixels for FXAA"
    pc = host.fxaa_default(W, H)
    img = oracle.fxaa(oracle.fxaa_prepare(r["result"]), r["segmask"], pc, gather4=True)
    if sharpener == 1:
        img = oracle.sharpen(img, 0.5)
    elif sharpener == 2:
        cs = host.cas_setup(0.5, W, H)
        img = oracle.cas(img, list(cs.const0), list(cs.const1))
    pp = host.postprocess_default(W, H)
    if post:
        img = oracle.post_process(img, None, pp)
    ch = abi.taa_post_chain()
    ch.sharpener, ch.postprocess, ch.pp, ch.fxaa, ch.fxaa_pc = sharpener, post, pp, 1, pc
    ch.sharpen.sharpeningFactor = 0.5
    ch.cas = host.cas_setup(0.5, W, H)
    ctx = host.TaaContext((W, H), flags=abi.TAA_FLAG_EXACT)
    final, ho = gpu_out(H, W), gpu_out(H, W)
    seg = torch.zeros(H, W, dtype=torch.int32, device="cuda")
    n0 = ctx.launch_count
    ctx.frame(u, ch, final, color=to_dev(ins["color"]), depth=to_dev(ins["depth"]), velocity=to_dev(ins["velocity"]), history_in=to_dev(hist),
              history_depth=to_dev(f0.depth.numpy()), matid=to_dev(ins["matid"]), prev_matid=to_dev(f0.matid.numpy()), uvnrm=to_dev(ins["uvnrm"]),
              history_out=ho, segmask=seg)
    torch.cuda.synchronize()
    assert ctx.launch_count - n0 == 2 + (1 if (sharpener or post) else 0)  # resolve, FXAA (prepare fused in), [sharpener + post-process]
    assert (seg.cpu().numpy().view(np.uint32) == r["segmask"]).all()
    cmp = slice(0, 3) if sharpener == 2 else slice(0, 4)
    rep = mismatch_report("final", img[..., cmp], final.cpu().numpy()[..., cmp])
    assert rep is None, rep
    # without the segmentation mask the chain must refuse, not skip FXAA silently
    with pytest.raises(abi.TaaError):
        ctx.frame(u, ch, final, color=to_dev(ins["color"]), depth=to_dev(ins["depth"]), velocity=to_dev(ins["velocity"]), history_in=to_dev(hist),
                  history_depth=to_dev(f0.depth.numpy()), history_out=ho)


def test_invokee_fxaa_branch(oracle):
    """`taa<CF>::render()` with mRayTraceAugment + TAA_RTFLAG_FXA: seg-mask ring, FXAA, CAS, post-process over a few frames."""
    W, H, CF = 160, 96, 3
    sc = SyntheticScene(W, H, pan_px=(1.5, 0.75))
    frames = [sc.frame(n) for n in range(5)]
    t = host.Taa(CF, flags=abi.TAA_FLAG_EXACT)
    p = configs.config3_full_chain()
    p.mRayTraceAugment = 1
    p.mRayTraceAugmentFlags = abi.TAA_RTFLAG_OUT | abi.TAA_RTFLAG_DPT | abi.TAA_RTFLAG_LUM | abi.TAA_RTFLAG_MID | abi.TAA_RTFLAG_CNT | abi.TAA_RTFLAG_FXA
    p.mRayTraceAugment_WDpt, p.mRayTraceAugment_WLum = 4.0, 1.5
    for i in range(2):
        C.memmove(C.addressof(t.mParameters[i]), C.addressof(p), C.sizeof(p))
    s = t.settings
    s.jitter.mSampleDistribution = 2
    s.mSharpener, s.mSharpenFactor, s.mPostProcessEnabled = 2, 0.5, 1
    slots = [dict(color=torch.empty_like(frames[0].color.cuda()), depth=torch.empty_like(frames[0].depth.cuda()), velocity=torch.empty_like(frames[0].velocity.cuda()),
                  matid=torch.empty_like(frames[0].matid.cuda()), uvnrm=torch.empty_like(frames[0].uvnrm.cuda())) for _ in range(CF)]
    t.set_source_image_views((W, H), [x["color"] for x in slots], [x["depth"] for x in slots], [x["uvnrm"] for x in slots], [x["velocity"] for x in slots],
                             [x["matid"] for x in slots])
    hist = [np.zeros((H, W, 4), np.float16) for _ in range(CF)]
    segs = [np.zeros((H, W), np.uint32) for _ in range(CF)]
    marked = 0
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
        if n == 0:
            want, cmp = f.color.numpy(), slice(0, 4)
        else:
            (jx, jy), npat = host.jitter_offset_for_frame(n, W, H, sample_distribution=2)
            pp = abi.TaaParameters.from_buffer_copy(p)
            pp.mRayTraceHistoryCount = npat  # -1 is replaced by the pattern length (taa.hpp:941)
            u = configs.uniforms_for(pp, (jx, jy))
            for k in range(4):
                u.mSinTime[k] = 0.0
            m = lambda a: (C.c_float * 16)(*a)
            abi.load_library().taa_reprojection_matrices(m(f.proj), m(f.view), m(frames[n - 1].proj), m(frames[n - 1].view),
                                                         u.mInverseViewProjMatrix, u.mHistoryViewProjMatrix)
            ins = np_inputs(f)
            r = oracle.resolve(u, ins["color"], ins["depth"], ins["velocity"], hist[last], history_depth=frames[n - 1].depth.numpy(), matid=ins["matid"],
                               prev_matid=frames[n - 1].matid.numpy(), uvnrm=ins["uvnrm"], prev_segmask=segs[last], want=("history_out", "result", "segmask"))
            hist[i], segs[i] = r["history_out"], r["segmask"]
            marked += int(((r["segmask"] & 3) == 1).sum())
            img = oracle.fxaa(oracle.fxaa_prepare(r["result"]), r["segmask"], host.fxaa_default(W, H), gather4=True)
            cs = host.cas_setup(0.5, W, H)
            want = oracle.post_process(oracle.cas(img, list(cs.const0), list(cs.const1)), None, host.postprocess_default(W, H))
            cmp = slice(0, 3)
        rep = mismatch_report(f"frame {n}", want[..., cmp], got[..., cmp])
        assert rep is None, rep
    assert marked > 0
    t.close()
