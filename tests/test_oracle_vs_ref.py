"""Pins the hand-written CPU restatement (oracle/taa_oracle.cpp) to the reference's own shader text.

oracle/_ref/libtaa_ref.so is shaders/taa.comp, sharpen.comp and post_process.comp of the reference, rewritten lexically to C++
and compiled against oracle/glsl_shim.h by oracle/ref_build.py (in the build container, where /root/reference exists; the
prebuilt library travels with the repo snapshot). Every output the shader writes must be bit-identical to the oracle's.
CPU only; runs wherever the library exists.
"""
import numpy as np
import pytest

import oracle_py
from common import mismatch_report, np_inputs, random_history, bits16
from taa_star_b200 import abi, configs, host
from taa_star_b200.synth import SyntheticScene

pytestmark = pytest.mark.skipif(not oracle_py.ref_available(), reason="oracle/_ref/libtaa_ref.so not built (needs /root/reference at build time)")

W, H = 96, 54
OUTS = ("history_out", "result", "debug")


def with_params(base, **kw):
    p = abi.TaaParameters.from_buffer_copy(base)
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def both(u, ins, hist, want=OUTS, **kw):
    a = oracle_py.resolve(u, ins["color"], ins["depth"], ins["velocity"], hist, want=want, **kw)
    b = oracle_py.resolve(u, ins["color"], ins["depth"], ins["velocity"], hist, want=want, impl="ref", **kw)
    problems = []
    for name in want:
        if name == "segmask":
            n = int((a[name] != b[name]).sum())
            if n:
                problems.append(f"segmask: {n} values differ")
        else:
            r = mismatch_report(name, b[name], a[name])
            if r:
                problems.append("(reference shader vs oracle) " + r)
    assert not problems, "\n".join(problems)


def frames(n=3, **kw):
    sc = SyntheticScene(W, H, **kw)
    return sc.frame(n - 1), sc.frame(n)


def uniforms(p, f0, f1, **kw):
    import ctypes as C
    u = configs.uniforms_for(p, f1.jitter_ndc, **kw)
    m = lambda a: (C.c_float * 16)(*a)
    abi.load_library().taa_reprojection_matrices(m(f1.proj), m(f1.view), m(f0.proj), m(f0.view), u.mInverseViewProjMatrix, u.mHistoryViewProjMatrix)
    return u


@pytest.mark.parametrize("cfg", ["config1", "config2", "config3"])
def test_baseline_configs(cfg):
    f0, f1 = frames(5)
    p = {"config1": configs.config1_defaults, "config2": configs.config2_resolve, "config3": configs.config3_full_chain}[cfg]()
    both(uniforms(p, f0, f1), np_inputs(f1), random_history(H, W, 7), history_depth=f0.depth.numpy())


SWITCHES = [
    dict(mInterpolationMode=0), dict(mInterpolationMode=1), dict(mColorClampingOrClipping=0), dict(mColorClampingOrClipping=1), dict(mColorClampingOrClipping=3),
    dict(mUseVelocityVectors=0), dict(mUseVelocityVectors=1), dict(mUseVelocityVectors=2, mVelocitySampleMode=1), dict(mUseVelocityVectors=2, mVelocitySampleMode=2),
    dict(mVarianceClipping=0, mShapedNeighbourhood=1), dict(mVarianceClipping=0, mShapedNeighbourhood=0), dict(mVarClipGamma=0.75),
    dict(mUseYCoCg=0), dict(mUseYCoCg=1, mShrinkChromaAxis=1), dict(mToneMapLumaKaris=1), dict(mToneMapLumaKaris=1, mUseYCoCg=0),
    dict(mReduceBlendNearClamp=1), dict(mLumaWeightingLottes=1), dict(mLumaWeightingLottes=1, mUseYCoCg=0), dict(mVelBasedAlpha=1, mVelBasedAlphaFactor=40.0),
    dict(mDepthCulling=1), dict(mRejectOutside=1, mRejectionAlpha=0.5), dict(mDynamicAntiGhosting=1), dict(mUnjitterNeighbourhood=1),
    dict(mUnjitterCurrentSample=1, mUnjitterFactor=-1.0), dict(mPassThrough=1), dict(mAddNoise=1),
    dict(mDebugMode=1), dict(mDebugMode=2), dict(mDebugMode=3), dict(mDebugMode=4), dict(mDebugMode=5), dict(mDebugMode=6, mDebugCenter=1),
    dict(mDebugMode=7, mDebugScale=2.5), dict(mDebugMode=8),
]


@pytest.mark.parametrize("sw", SWITCHES, ids=lambda s: ",".join(f"{k}={v}" for k, v in s.items()))
def test_each_switch(sw):
    f0, f1 = frames(3, pan_px=(5.25, -2.5))
    u = uniforms(with_params(configs.config2_resolve(), **sw), f0, f1)
    u.mSinTime[0] = 0.37
    both(u, np_inputs(f1), random_history(H, W, 11), history_depth=f0.depth.numpy())


def test_uniform_flags_and_split_screen():
    f0, f1 = frames(2)
    ins, hist = np_inputs(f1), random_history(H, W, 3)
    both(uniforms(configs.config2_resolve(), f0, f1, reset_history=True), ins, hist)
    u = uniforms(configs.config2_resolve(), f0, f1)
    u.mBypassHistoryUpdate = 1
    both(u, ins, hist)
    both(uniforms(configs.config2_resolve(), f0, f1, params1=configs.config1_defaults(), split_x=W // 3), ins, hist)


def test_upsampling():
    for (iw, ih, ow, oh) in ((48, 27, 96, 54), (50, 30, 75, 45)):
        sc = SyntheticScene(iw, ih)
        f0, f1 = sc.frame(2), sc.frame(3)
        for p in (configs.config1_defaults(), configs.config2_resolve(), with_params(configs.config2_resolve(), mUnjitterFactor=-1.0)):
            both(uniforms(p, f0, f1, upsampling=True), np_inputs(f1), random_history(oh, ow, 2), out_size=(ow, oh))


def test_segmentation_mask():
    f0, f1 = frames(7)
    ins, hist = np_inputs(f1), random_history(H, W, 4)
    prev_seg = (np.random.default_rng(8).integers(0, 6, (H, W)).astype(np.uint32) << 16) | 2
    all_flags = 0xffffffff & ~(abi.TAA_RTFLAG_ALL | abi.TAA_RTFLAG_FXD)
    for flags in (abi.TAA_RTFLAG_OUT, abi.TAA_RTFLAG_DIS, abi.TAA_RTFLAG_NRM, abi.TAA_RTFLAG_DPT, abi.TAA_RTFLAG_MID, abi.TAA_RTFLAG_LUM,
                  abi.TAA_RTFLAG_CNT | abi.TAA_RTFLAG_MID, abi.TAA_RTFLAG_ALL, abi.TAA_RTFLAG_FXD | abi.TAA_RTFLAG_MID, all_flags):
        p = with_params(configs.config3_full_chain(), mRayTraceAugment=1, mRayTraceAugmentFlags=flags, mRayTraceHistoryCount=8,
                        mRayTraceAugment_WDpt=4.0, mRayTraceAugment_WLum=1.5, mRayTraceAugment_WNrm=40.0)
        both(uniforms(p, f0, f1), ins, hist, want=OUTS + ("segmask",), history_depth=f0.depth.numpy(), matid=ins["matid"], prev_matid=f0.matid.numpy(),
             prev_segmask=prev_seg, uvnrm=ins["uvnrm"])


def test_extreme_motion():
    f0, f1 = frames(2, pan_px=(0.0, 0.0))
    ins = np_inputs(f1)
    rng = np.random.default_rng(5)
    vel = ins["velocity"].astype(np.float32)
    vel[..., 0] = rng.uniform(-1.5, 1.5, (H, W))
    vel[..., 1] = rng.uniform(-1.5, 1.5, (H, W))
    vel[0:4, :, 0:2] = 0.0
    vel[10, :, 0] = 60000.0
    ins["velocity"] = vel.astype(np.float16)
    for p in (configs.config3_full_chain(), with_params(configs.config3_full_chain(), mInterpolationMode=0)):
        both(uniforms(p, f0, f1), ins, random_history(H, W, 9), history_depth=f1.depth.numpy())


def test_sequence_of_16_frames():
    sc = SyntheticScene(W, H)
    hist_a = np.zeros((H, W, 4), np.float16)
    hist_b = hist_a.copy()
    prev = None
    for n in range(16):
        f = sc.frame(n)
        ins = np_inputs(f)
        u = configs.uniforms_for(configs.config3_full_chain(), f.jitter_ndc, reset_history=(n == 0))
        hd = prev if prev is not None else ins["depth"]
        a = oracle_py.resolve(u, ins["color"], ins["depth"], ins["velocity"], hist_a, history_depth=hd)
        b = oracle_py.resolve(u, ins["color"], ins["depth"], ins["velocity"], hist_b, history_depth=hd, impl="ref")
        assert mismatch_report(f"history_out frame {n}", b["history_out"], a["history_out"]) is None
        hist_a, hist_b, prev = a["history_out"], b["history_out"], ins["depth"]


def test_sharpen_and_post_process():
    src = random_history(H, W, 21, alpha_binary=False)
    dbg = random_history(H, W, 22, alpha_binary=False)
    for f in (0.0, 0.5, 2.0):
        assert mismatch_report("sharpen", oracle_py.ref_sharpen(src, f), oracle_py.sharpen(src, f)) is None
    pp = host.postprocess_default(W, H)
    cases = [dict(), dict(splitX=W // 2), dict(zoom=1), dict(zoom=1, showZoomBox=0), dict(debugL_show=1), dict(splitX=W // 3, debugR_show=1)]
    for c in cases:
        pc = abi.TaaPostProcessPush.from_buffer_copy(pp)
        for k, v in c.items():
            setattr(pc, k, v)
        pc.debugR_mask[3] = 1.0
        assert mismatch_report(f"post_process {c}", oracle_py.ref_post_process(src, dbg, pc), oracle_py.post_process(src, dbg, pc)) is None


def test_cas_filter_body():
    """sharpen_cas.comp + ffx_cas.h's CasFilter (the reference's own text through the shim, with the functions of ffx_a.h it uses spliced in
    by name) against the hand restatement in taa_oracle.cpp: rgb bit for bit, for every sharpness, on ragged sizes (the 16x16 workgroup
    tiling hangs over the image; border loads go out of range), on values above 1 and on flat, zero and negative regions.
    Alpha: the shader stores an uninitialised `AF4 c` (undefined in GLSL); the restatement and the CUDA kernel write 1."""
    rng = np.random.default_rng(31)
    for (h, w, sharp) in [(72, 128, 0.5), (37, 53, 0.0), (16, 16, 1.0), (90, 160, 0.8), (1, 1, 0.5), (5, 300, 0.25)]:
        src = random_history(h, w, 40 + h, alpha_binary=False).astype(np.float32)
        src[: h // 3, : w // 3, :3] *= 3.0                      # above 1: the filter saturates
        src[h // 2:, w // 2:, :3] = 0.25                        # flat
        src[: h // 4, w // 2:, :3] = 0.0                        # zero: the reciprocal approximations see 0
        src[h // 2:, : w // 4, :3] -= 0.5                       # negative values
        src[..., :3] += (rng.random((h, w, 3), dtype=np.float32) < 0.02) * 8.0   # isolated peaks
        src = src.astype(np.float16)
        c0, c1 = oracle_py.cas_setup(sharp, w, h)
        a, b = oracle_py.cas(src, c0, c1), oracle_py.ref_cas(src, c0, c1)
        assert mismatch_report(f"cas {w}x{h} sharpness {sharp}", b[..., :3], a[..., :3]) is None


def fxaa_test_image(h, w, seed):
    """Hard edges at all orientations (long ones too, so that the end-of-edge search runs out its steps), thin lines, noise, flat areas."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    img = np.zeros((h, w, 4), np.float32)
    img[..., 0] = (yy > 0.37 * xx + 3.3)                                   # shallow diagonal edge: long spans
    img[..., 1] = ((xx - w / 2) ** 2 + (yy - h / 2) ** 2 < (0.3 * h) ** 2)  # circle
    img[..., 2] = (xx > 0.11 * yy + w / 3)                                 # steep diagonal
    img[h // 2, :, :3] = 1.0                                               # 1-px horizontal line
    img[:, w // 4, :3] = 0.0                                               # 1-px vertical line
    img[: h // 6, : w // 3, :3] = rng.random((h // 6, w // 3, 3), dtype=np.float32)  # noise block
    img[..., :3] = img[..., :3] * 0.8 + 0.1 * rng.random((h, w, 1), dtype=np.float32) * (yy[..., None] > 0.8 * h)
    img[..., 3] = rng.random((h, w), dtype=np.float32)                     # garbage alpha: fxaa_prepare must overwrite it
    return img.astype(np.float16)


def test_fxaa_prepare_and_fxaa():
    for (h, w, seed) in ((H, W, 31), (37, 53, 32), (1, 1, 33), (2, 64, 34)):
        src = fxaa_test_image(h, w, seed)
        prep = oracle_py.fxaa_prepare(src)
        assert mismatch_report("fxaa_prepare", oracle_py.fxaa_prepare(src, impl="ref"), prep) is None
        rng = np.random.default_rng(seed + 100)
        seg_all = np.full((h, w), 1 | (5 << 16), np.uint32)
        seg_mixed = rng.integers(0, 4, (h, w)).astype(np.uint32) | (rng.integers(0, 9, (h, w)).astype(np.uint32) << 16)
        for seg in (seg_all, seg_mixed):
            for pc in (oracle_py.fxaa_push(w, h), oracle_py.fxaa_push(w, h, subpix=0.25, edge_threshold=0.333, edge_threshold_min=0.0312),
                       oracle_py.fxaa_push(w, h, subpix=1.0, edge_threshold=0.063, edge_threshold_min=0.0)):
                for gather4 in (True, False):
                    a = oracle_py.fxaa(prep, seg, pc, gather4=gather4)
                    b = oracle_py.fxaa(prep, seg, pc, gather4=gather4, impl="ref")
                    assert mismatch_report(f"fxaa {h}x{w} gather4={gather4}", b, a) is None
        # FXAA must actually do something on this image, and the two texel-access variants must not be the same code path
        out = oracle_py.fxaa(prep, seg_all, oracle_py.fxaa_push(w, h))
        if h > 8:
            assert (bits16(out) != bits16(prep)).any()
