"""GPU parity of the follow-on passes (sharpen.comp, sharpen_cas.comp, post_process.comp), of taa_frame (the whole chain render() records,
taa.hpp:1008-1159) and of the invokee (`class taa<CF>`: history ring, first-frame blit, jitter, host-buffer frames) against the CPU oracle."""
import ctypes as C

import numpy as np
import pytest
import torch

from common import TOL_ABS, mismatch_report, np_inputs, random_history, same_f16, to_dev
from taa_star_b200 import abi, configs, host
from taa_star_b200.synth import SyntheticScene

pytestmark = pytest.mark.gpu

W, H = 200, 112


def rand_img(seed, hdr=False):
    rng = np.random.default_rng(seed)
    a = rng.random((H, W, 4), dtype=np.float32)
    if hdr:
        a[..., :3] *= 3.0
    return a.astype(np.float16)


def gpu_out():
    return torch.full((H, W, 4), float("nan"), dtype=torch.float16, device="cuda")


def pp_cases():
    base = host.postprocess_default(W, H)
    out = []
    for c in (dict(), dict(splitX=W // 2), dict(zoom=1), dict(zoom=1, showZoomBox=0), dict(debugL_show=1), dict(splitX=W // 3, debugR_show=1),
              dict(zoom=1, debugL_show=1, splitX=5)):
        pc = abi.TaaPostProcessPush.from_buffer_copy(base)
        for k, v in c.items():
            setattr(pc, k, v)
        pc.debugR_mask[3] = 1.0
        out.append(pc)
    return out


def test_sharpen_cas_post_process_kernels(oracle):
    ctx = host.TaaContext((W, H))
    for seed, hdr in ((1, False), (2, True)):
        src, dbg = rand_img(seed, hdr), rand_img(seed + 10)
        for f in (0.0, 0.5, 2.0):
            dst = gpu_out()
            ctx.sharpen(to_dev(src), dst, f)
            assert mismatch_report(f"sharpen {f}", oracle.sharpen(src, f), dst.cpu().numpy()) is None
        for sharp in (0.0, 0.5, 1.0):
            pc = host.cas_setup(sharp, W, H)
            dst = gpu_out()
            ctx.sharpen_cas(to_dev(src), dst, pc)
            ref = oracle.cas(src, list(pc.const0), list(pc.const1))
            r = mismatch_report(f"cas {sharp}", ref[..., :3], dst.cpu().numpy()[..., :3])  # alpha is undefined in the reference (sharpen_cas.comp:38)
            assert r is None, r
        for pc in pp_cases():
            dst = gpu_out()
            ctx.post_process(to_dev(src), to_dev(dbg), dst, pc)
            assert mismatch_report("post_process", oracle.post_process(src, dbg, pc), dst.cpu().numpy()) is None


def chain_of(sharpener, post, pp=None, factor=0.5):
    ch = abi.taa_post_chain()
    ch.sharpener = sharpener
    ch.sharpen.sharpeningFactor = factor
    ch.cas = host.cas_setup(factor, W, H)
    ch.postprocess = post
    ch.pp = pp if pp is not None else host.postprocess_default(W, H)
    return ch


def oracle_chain(oracle, u, ins, hist, hd, ch, debug=False):
    want = ("history_out", "result") + (("debug",) if debug else ())
    r = oracle.resolve(u, ins["color"], ins["depth"], ins["velocity"], hist, history_depth=hd, want=want)
    img = r["result"]
    if ch.sharpener == 1:
        img = oracle.sharpen(img, ch.sharpen.sharpeningFactor)
    elif ch.sharpener == 2:
        img = oracle.cas(img, list(ch.cas.const0), list(ch.cas.const1))
    if ch.postprocess:
        img = oracle.post_process(img, r.get("debug"), ch.pp)
    return r, img


@pytest.mark.parametrize("flags", [abi.TAA_FLAG_EXACT, 0], ids=["exact", "tuned"])
@pytest.mark.parametrize("sharpener,post", [(0, 0), (0, 1), (1, 0), (1, 1), (2, 0), (2, 1)])
def test_taa_frame_chain(oracle, flags, sharpener, post):
    sc = SyntheticScene(W, H, pan_px=(2.5, -1.25))
    f0, f1 = sc.frame(3), sc.frame(4)
    ins, hist = np_inputs(f1), f0.color.numpy().copy()
    u = configs.uniforms_for(configs.config3_full_chain(), f1.jitter_ndc)
    ctx = host.TaaContext((W, H), flags=flags)
    for pp in ([None] if not post else [None] + pp_cases()[1:4]):
        ch = chain_of(sharpener, post, pp)
        ref, ref_final = oracle_chain(oracle, u, ins, hist, f0.depth.numpy(), ch)
        final, ho = gpu_out(), gpu_out()
        n0 = ctx.launch_count
        ctx.frame(u, ch, final, color=to_dev(ins["color"]), depth=to_dev(ins["depth"]), velocity=to_dev(ins["velocity"]), history_in=to_dev(hist),
                  history_depth=to_dev(f0.depth.numpy()), history_out=ho)
        torch.cuda.synchronize()
        launches = ctx.launch_count - n0
        resolve_launches = 1 if flags else 2
        # without a sharpener an identity post-process is a copy: the tuned path lets the resolve write `final` itself
        identity_copy = (not flags) and (not sharpener) and post and (pp is None)
        assert launches == resolve_launches + (1 if ((sharpener or post) and not identity_copy) else 0), "sharpen/CAS + post-process must be one launch"
        got = final.cpu().numpy()
        ch_cmp = slice(0, 3) if sharpener == 2 else slice(0, 4)
        if flags:
            assert mismatch_report("final", ref_final[..., ch_cmp], got[..., ch_cmp]) is None
            assert mismatch_report("history_out", ref["history_out"], ho.cpu().numpy()) is None
        else:
            # CAS and the unsharp mask amplify a difference of the resolved value: <= (1 + 4*0.5) x for sharpen at factor 0.5, <= ~1.6 x for CAS
            tol = TOL_ABS * (3.0 if sharpener == 1 else 2.0 if sharpener == 2 else 1.0)
            d = np.abs(ref_final[..., ch_cmp].astype(np.float32) - got[..., ch_cmp].astype(np.float32)).max()
            assert d <= tol, f"final: max |d| = {d}"
            assert np.abs(ref["history_out"].astype(np.float32) - ho.cpu().numpy().astype(np.float32)).max() <= TOL_ABS


def test_taa_frame_debug_view(oracle):
    sc = SyntheticScene(W, H)
    f0, f1 = sc.frame(1), sc.frame(2)
    ins, hist = np_inputs(f1), f0.color.numpy().copy()
    p = configs.config2_resolve()
    p.mDebugMode = 3
    p.mDebugToScreenOutput = 1
    u = configs.uniforms_for(p, f1.jitter_ndc)
    pp = host.postprocess_default(W, H)
    pp.debugL_show = 1
    ch = chain_of(2, 1, pp)
    ref, ref_final = oracle_chain(oracle, u, ins, hist, None, ch, debug=True)
    ctx = host.TaaContext((W, H))
    final, ho, dbg = gpu_out(), gpu_out(), gpu_out()
    ctx.frame(u, ch, final, color=to_dev(ins["color"]), depth=to_dev(ins["depth"]), velocity=to_dev(ins["velocity"]), history_in=to_dev(hist),
              history_out=ho, debug=dbg)
    torch.cuda.synchronize()
    assert mismatch_report("final (debug view)", ref_final, final.cpu().numpy()) is None  # a debug image forces the exact general kernel


def _run_invokee_sequence(oracle, host_frames: bool, sharpener: int, nframes: int = 7):
    """Drives `Taa` like wookiee drives taa<3> (main.cpp:4122-4123, 4148; taa.hpp:894, 974) and mirrors it with the oracle."""
    CF = 3
    sc = SyntheticScene(W, H, pan_px=(1.5, 0.75))
    frames = [sc.frame(n) for n in range(nframes)]
    t = host.Taa(CF, flags=abi.TAA_FLAG_EXACT)
    p = configs.config3_full_chain()
    for i in range(2):
        C.memmove(C.addressof(t.mParameters[i]), C.addressof(p), C.sizeof(p))
    s = t.settings
    s.jitter.mSampleDistribution = 2
    s.mSharpener = sharpener
    s.mSharpenFactor = 0.5
    s.mPostProcessEnabled = 1
    dev = [dict(color=f.color.cuda(), depth=f.depth.cuda(), velocity=f.velocity.cuda()) for f in frames]
    slots = [dict(color=torch.empty_like(dev[0]["color"]), depth=torch.empty_like(dev[0]["depth"]), velocity=torch.empty_like(dev[0]["velocity"])) for _ in range(CF)]
    if host_frames:
        t.set_sizes_for_host_frames((W, H), (W, H))
    else:
        t.set_source_image_views((W, H), [x["color"] for x in slots], [x["depth"] for x in slots], None, [x["velocity"] for x in slots])
    hist = [np.zeros((H, W, 4), np.float16) for _ in range(CF)]
    outs = []
    for n, f in enumerate(frames):
        i, last = n % CF, (n + CF - 1) % CF
        if host_frames:
            fin = torch.empty(H, W, 4, dtype=torch.float16).pin_memory()
            hc, hd, hv = f.color.pin_memory(), f.depth.pin_memory(), f.velocity.pin_memory()  # kept alive until the frame has left the device
            t.frame_host(n, hc, hd, hv, f.view, f.proj, fin)
            t.wait(n)
            got = fin.numpy().copy()
        else:
            for k in slots[i]:
                slots[i][k].copy_(dev[n][k])
            proj, jit = t.get_jittered_projection_matrix(f.proj, n)
            t.save_history_proj_matrix(f.proj, n)
            t.update(n, f.view)
            ptr = t.render(n)
            torch.cuda.synchronize()
            got = t.image_by_ptr(ptr).cpu().numpy()
        # the mirror
        if n == 0:  # very first frame: blit colour -> result (taa.hpp:990, 1176)
            want = f.color.numpy()
        else:
            (jx, jy), _ = host.jitter_offset_for_frame(n, W, H, sample_distribution=2)
            u = configs.uniforms_for(p, (jx, jy))
            u.mSinTime[0] = u.mSinTime[1] = u.mSinTime[2] = u.mSinTime[3] = 0.0
            m = lambda a: (C.c_float * 16)(*a)
            abi.load_library().taa_reprojection_matrices(m(f.proj), m(f.view), m(frames[n - 1].proj), m(frames[n - 1].view),
                                                         u.mInverseViewProjMatrix, u.mHistoryViewProjMatrix)
            ch = chain_of(sharpener, 1)
            ins = np_inputs(f)
            ref, want = oracle_chain(oracle, u, ins, hist[last], frames[n - 1].depth.numpy(), ch)
            hist[i] = ref["history_out"]
        cmp = slice(0, 3) if (sharpener == 2 and n > 0) else slice(0, 4)
        r = mismatch_report(f"frame {n}", want[..., cmp], got[..., cmp])
        assert r is None, r
        outs.append(got)
    assert t.launch_count >= (nframes - 1) * 2
    if not host_frames:
        assert t.duration() > 0.0
    t.close()


@pytest.mark.parametrize("sharpener", [0, 2])
def test_invokee_render_sequence(oracle, sharpener):
    _run_invokee_sequence(oracle, host_frames=False, sharpener=sharpener)


def test_invokee_host_buffer_frames(oracle):
    _run_invokee_sequence(oracle, host_frames=True, sharpener=2)


def test_invokee_with_upsampling(oracle):
    """TAAU through the invokee (in_size != out_size, taa.hpp:292): the very first frame is avk::blit_image's nearest-texel scaling of the
    colour buffer (taa.hpp:1176, avk.cpp:7829), the following frames are resolved at the target resolution."""
    CF = 3
    iw, ih, ow, oh = 96, 54, 192, 108
    sc = SyntheticScene(iw, ih, pan_px=(0.75, 0.5))
    frames = [sc.frame(n) for n in range(4)]
    t = host.Taa(CF, flags=abi.TAA_FLAG_EXACT)
    p = configs.config2_resolve()
    for i in range(2):
        C.memmove(C.addressof(t.mParameters[i]), C.addressof(p), C.sizeof(p))
    t.settings.jitter.mSampleDistribution = 2
    t.settings.mPostProcessEnabled = 0
    slots = [dict(color=torch.empty_like(frames[0].color).cuda(), depth=torch.empty_like(frames[0].depth).cuda(),
                  velocity=torch.empty_like(frames[0].velocity).cuda()) for _ in range(CF)]
    t.set_source_image_views((ow, oh), [x["color"] for x in slots], [x["depth"] for x in slots], None, [x["velocity"] for x in slots])
    hist = [np.zeros((oh, ow, 4), np.float16) for _ in range(CF)]
    for n, f in enumerate(frames):
        i, last = n % CF, (n + CF - 1) % CF
        for k in slots[i]:
            slots[i][k].copy_(getattr(f, k))
        t.get_jittered_projection_matrix(f.proj, n)
        t.save_history_proj_matrix(f.proj, n)
        t.update(n, f.view)
        ptr = t.render(n)
        torch.cuda.synchronize()
        got = t.image_by_ptr(ptr).cpu().numpy()
        assert got.shape == (oh, ow, 4)
        if n == 0:
            ys = np.minimum(np.floor((np.arange(oh, dtype=np.float32) + np.float32(0.5)) * (np.float32(ih) / np.float32(oh))).astype(int), ih - 1)
            xs = np.minimum(np.floor((np.arange(ow, dtype=np.float32) + np.float32(0.5)) * (np.float32(iw) / np.float32(ow))).astype(int), iw - 1)
            want = f.color.numpy()[ys][:, xs]
        else:
            (jx, jy), _ = host.jitter_offset_for_frame(n, iw, ih, sample_distribution=2)  # the input resolution (taa.hpp:153-154)
            u = configs.uniforms_for(p, (jx, jy), upsampling=True)
            u.mSinTime[0] = u.mSinTime[1] = u.mSinTime[2] = u.mSinTime[3] = 0.0
            m = lambda a: (C.c_float * 16)(*a)
            abi.load_library().taa_reprojection_matrices(m(f.proj), m(f.view), m(frames[n - 1].proj), m(frames[n - 1].view),
                                                         u.mInverseViewProjMatrix, u.mHistoryViewProjMatrix)
            ins = np_inputs(f)
            ref = oracle.resolve(u, ins["color"], ins["depth"], ins["velocity"], hist[last], history_depth=frames[n - 1].depth.numpy(),
                                 out_size=(ow, oh), want=("history_out", "result"))
            want = ref["result"]
            hist[i] = ref["history_out"]
        r = mismatch_report(f"upsampling frame {n}", want, got)
        assert r is None, r
    t.close()


def test_invokee_new_size_after_host_frames():
    """set_source_image_views() after frame_host() has run (a resolution change): the host-frame pipeline is rebuilt for the new size."""
    t = host.Taa(3)
    p = configs.config2_resolve()
    for i in range(2):
        C.memmove(C.addressof(t.mParameters[i]), C.addressof(p), C.sizeof(p))
    t.settings.jitter.mSampleDistribution = 2
    for (w, h) in ((160, 90), (224, 126), (96, 54)):
        sc = SyntheticScene(w, h)
        t.set_sizes_for_host_frames((w, h), (w, h))
        outs = []
        for n in range(4):
            f = sc.frame(n)
            fin = torch.empty(h, w, 4, dtype=torch.float16).pin_memory()
            hc, hd, hv = f.color.pin_memory(), f.depth.pin_memory(), f.velocity.pin_memory()
            t.frame_host(n, hc, hd, hv, f.view, f.proj, fin)
            t.wait(n)
            outs.append(fin.numpy().copy())
        assert np.isfinite(outs[-1].astype(np.float32)).all() and 0.05 < float(outs[-1][..., :3].astype(np.float32).mean()) < 0.95
    t.close()


def test_right_hand_debug_setting_without_a_splitter(oracle):
    """param[1].mDebugToScreenOutput is copied into the post-process constants unconditionally (taa.hpp:955-959), but post_process.comp never
    looks at the right-hand settings without a splitter: the frame renders normally, no debug image needed."""
    sc = SyntheticScene(W, H)
    f0, f1 = sc.frame(1), sc.frame(2)
    ins, hist = np_inputs(f1), f0.color.numpy().copy()
    u = configs.uniforms_for(configs.config2_resolve(), f1.jitter_ndc)
    pp = host.postprocess_default(W, H)
    pp.debugR_show = 1
    ch = chain_of(0, 1, pp)
    ref, ref_final = oracle_chain(oracle, u, ins, hist, None, ch)
    ctx = host.TaaContext((W, H), flags=abi.TAA_FLAG_EXACT)
    final, ho = gpu_out(), gpu_out()
    ctx.frame(u, ch, final, color=to_dev(ins["color"]), depth=to_dev(ins["depth"]), velocity=to_dev(ins["velocity"]), history_in=to_dev(hist), history_out=ho)
    torch.cuda.synchronize()
    assert mismatch_report("final", ref_final, final.cpu().numpy()) is None


@pytest.mark.parametrize("size", [(200, 112), (2, 5), (62, 7), (130, 33), (258, 65), (61, 9)])
@pytest.mark.parametrize("sharpener,post", [(1, 0), (1, 1), (2, 0), (2, 1), (0, 1)])
def test_fused_chain_is_one_launch(oracle, size, sharpener, post):
    """Settings that leave nothing to the exact pass (BASELINE config 2): [sharpen | CAS] + an identity post-process ride in the resolve's
    epilogue (taa_resolve_stream.cu) — ONE launch, no intermediate image — and agree with the oracle's three-pass chain: the sharpeners
    amplify a difference of the resolved value by at most 3x (unsharp mask at factor 0.5) / ~1.6x (CAS). (Rows of an odd number of texels
    cannot be described by a tensor map: such images take the strip kernel and the separate sharpening launch.)"""
    w, h = size
    sc = SyntheticScene(w, h, pan_px=(2.5, -1.25))
    f0, f1 = sc.frame(3), sc.frame(4)
    ins, hist = np_inputs(f1), f0.color.numpy().copy()
    u = configs.uniforms_for(configs.config2_resolve(), f1.jitter_ndc)
    ch = abi.taa_post_chain()
    ch.sharpener = sharpener
    ch.sharpen.sharpeningFactor = 0.5
    ch.cas = host.cas_setup(0.5, w, h)
    ch.postprocess = post
    ch.pp = host.postprocess_default(w, h)
    ref, ref_final = oracle_chain(oracle, u, ins, hist, None, ch)
    ctx = host.TaaContext((w, h))
    final = torch.full((h, w, 4), float("nan"), dtype=torch.float16, device="cuda")
    ho = torch.full((h, w, 4), float("nan"), dtype=torch.float16, device="cuda")
    n0 = ctx.launch_count
    ctx.frame(u, ch, final, color=to_dev(ins["color"]), depth=to_dev(ins["depth"]), velocity=to_dev(ins["velocity"]), history_in=to_dev(hist), history_out=ho)
    torch.cuda.synchronize()
    fusable = (w % 2 == 0) or not sharpener
    assert ctx.launch_count - n0 == (1 if fusable else 2), "the chain must be one launch"
    got = final.cpu().numpy()
    cmp = slice(0, 3) if sharpener == 2 else slice(0, 4)
    tol = TOL_ABS * (3.0 if sharpener == 1 else 2.0 if sharpener == 2 else 1.0)
    d = np.abs(ref_final[..., cmp].astype(np.float32) - got[..., cmp].astype(np.float32))
    assert np.isfinite(got[..., cmp].astype(np.float32)).all()
    assert d.max() <= tol, f"final: max |d| = {d.max()} at {np.unravel_index(d.argmax(), d.shape)}"
    assert np.abs(ref["history_out"].astype(np.float32) - ho.cpu().numpy().astype(np.float32)).max() <= TOL_ABS
