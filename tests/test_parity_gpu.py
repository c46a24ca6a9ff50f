"""GPU parity: the CUDA path (through the C-ABI) against the CPU oracle on the same seeded inputs.

Bar (BASELINE.json north_star): rgba16f outputs within 2^-10 per channel per frame, PSNR >= 60 dB after 64 accumulated
frames, integer masks bit-exact. The EXACT kernels are held to a stricter bar here: bit-identical fp16 outputs
(contexts are created with TAA_FLAG_EXACT). The default (tuned) path is tested in test_tuned_gpu.py.
"""
import ctypes as C

import numpy as np
import pytest
import torch

from common import (PSNR_MIN, TOL_ABS, mismatch_report, np_inputs, psnr, random_history, run_gpu_resolve, same_f16, to_dev)
from taa_star_b200 import abi, configs, host
from taa_star_b200.synth import SyntheticScene

pytestmark = pytest.mark.gpu

W, H = 256, 144


def scene(w=W, h=H, **kw):
    return SyntheticScene(w, h, **kw)


def check_exact(oracle, u, ins, hist, out_size=None, want=("history_out", "result", "mask"), flags=abi.TAA_FLAG_EXACT, hist_depth=None, prev_matid=None,
                prev_segmask=None, tol=None, ctx=None):
    in_h, in_w = ins["depth"].shape
    ref = oracle.resolve(u, ins["color"], ins["depth"], ins["velocity"], hist, history_depth=hist_depth, prev_segmask=prev_segmask,
                         matid=ins.get("matid"), prev_matid=prev_matid, uvnrm=ins.get("uvnrm"), out_size=out_size, want=want)
    own = ctx is None
    if own:
        ctx = host.TaaContext((in_w, in_h), out_size, flags=flags)
    got = run_gpu_resolve(ctx, u, ins, hist, hist_depth=hist_depth, prev_matid=prev_matid, prev_segmask=prev_segmask, out_size=out_size, want=want)
    if own:
        ctx.close()
    problems = []
    for name in want:
        if name in ("mask", "segmask"):
            bad = int((ref[name] != got[name]).sum())
            if bad:
                problems.append(f"{name}: {bad} of {ref[name].size} integer values differ")
        elif tol is None:
            r = mismatch_report(name, ref[name], got[name])
            if r:
                problems.append(r)
        else:
            d = np.abs(ref[name].astype(np.float32) - got[name].astype(np.float32)).max()
            if not d <= tol:
                problems.append(f"{name}: max |d| = {d} > {tol}")
    assert not problems, "\n".join(problems)
    return ref, got


def with_params(base, **kw):
    p = abi.TaaParameters.from_buffer_copy(base)
    for k, v in kw.items():
        setattr(p, k, v)
    return p


# ---- the three BASELINE configs ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cfg", ["config1", "config2", "config3"])
def test_baseline_configs_single_frame(oracle, cfg):
    sc = scene()
    f0, f1 = sc.frame(4), sc.frame(5)
    ins = np_inputs(f1)
    hist = random_history(H, W, 7)
    p = {"config1": configs.config1_defaults, "config2": configs.config2_resolve, "config3": configs.config3_full_chain}[cfg]()
    u = configs.uniforms_for(p, f1.jitter_ndc)
    lib = abi.load_library()
    m = lambda a: (C.c_float * 16)(*a)
    lib.taa_reprojection_matrices(m(f1.proj), m(f1.view), m(f0.proj), m(f0.view), u.mInverseViewProjMatrix, u.mHistoryViewProjMatrix)
    check_exact(oracle, u, ins, hist, hist_depth=f0.depth.numpy(), want=("history_out", "result", "mask", "debug"))


# ---- every switch of Parameters, one at a time on top of a base ---------------------------------------------------------------
SWITCHES = [
    dict(mInterpolationMode=0), dict(mInterpolationMode=1), dict(mInterpolationMode=2),
    dict(mColorClampingOrClipping=0), dict(mColorClampingOrClipping=1), dict(mColorClampingOrClipping=2), dict(mColorClampingOrClipping=3),
    dict(mUseVelocityVectors=0), dict(mUseVelocityVectors=1), dict(mUseVelocityVectors=2, mVelocitySampleMode=1), dict(mUseVelocityVectors=2, mVelocitySampleMode=2),
    dict(mUseVelocityVectors=1, mVelocitySampleMode=2),
    dict(mVarianceClipping=0, mShapedNeighbourhood=1), dict(mVarianceClipping=0, mShapedNeighbourhood=0), dict(mVarClipGamma=0.75),
    dict(mUseYCoCg=0), dict(mUseYCoCg=1, mShrinkChromaAxis=1), dict(mUseYCoCg=0, mShrinkChromaAxis=1),
    dict(mToneMapLumaKaris=1), dict(mToneMapLumaKaris=1, mUseYCoCg=0),
    dict(mReduceBlendNearClamp=1), dict(mLumaWeightingLottes=1), dict(mLumaWeightingLottes=1, mUseYCoCg=0),
    dict(mVelBasedAlpha=1, mVelBasedAlphaFactor=40.0), dict(mDepthCulling=1), dict(mRejectOutside=1), dict(mDynamicAntiGhosting=1),
    dict(mRejectOutside=1, mRejectionAlpha=0.5), dict(mUnjitterNeighbourhood=1), dict(mUnjitterCurrentSample=1, mUnjitterFactor=-1.0),
    dict(mUnjitterNeighbourhood=1, mUnjitterCurrentSample=1), dict(mPassThrough=1), dict(mAlpha=0.3),
    dict(mDebugMode=1), dict(mDebugMode=2), dict(mDebugMode=3), dict(mDebugMode=4), dict(mDebugMode=5), dict(mDebugMode=6, mDebugCenter=1),
    dict(mDebugMode=7, mDebugScale=2.5), dict(mDebugMode=8),
]


@pytest.mark.parametrize("sw", SWITCHES, ids=lambda s: ",".join(f"{k}={v}" for k, v in s.items()))
def test_each_switch(oracle, sw):
    sc = scene(pan_px=(5.25, -2.5))
    f0, f1 = sc.frame(2), sc.frame(3)
    ins = np_inputs(f1)
    hist = random_history(H, W, 11)
    p = with_params(configs.config2_resolve(), **sw)
    u = configs.uniforms_for(p, f1.jitter_ndc)
    lib = abi.load_library()
    m = lambda a: (C.c_float * 16)(*a)
    lib.taa_reprojection_matrices(m(f1.proj), m(f1.view), m(f0.proj), m(f0.view), u.mInverseViewProjMatrix, u.mHistoryViewProjMatrix)
    check_exact(oracle, u, ins, hist, hist_depth=f0.depth.numpy(), want=("history_out", "result", "mask", "debug"))


def test_uniform_flags_reset_bypass_split(oracle):
    sc = scene()
    f1 = sc.frame(1)
    ins = np_inputs(f1)
    hist = random_history(H, W, 3)
    base = configs.config2_resolve()
    u = configs.uniforms_for(base, f1.jitter_ndc, reset_history=True)
    check_exact(oracle, u, ins, hist)
    u = configs.uniforms_for(base, f1.jitter_ndc)
    u.mBypassHistoryUpdate = 1
    check_exact(oracle, u, ins, hist)
    # split screen: left = config 2, right = defaults; column splitX itself uses param[0] (taa.comp:717)
    u = configs.uniforms_for(base, f1.jitter_ndc, params1=configs.config1_defaults(), split_x=W // 3)
    check_exact(oracle, u, ins, hist)


def test_extreme_and_non_finite_motion(oracle):
    """History positions far outside the image, on the borders, and NaN/inf velocities (saturating conversions)."""
    sc = scene(pan_px=(0.0, 0.0))
    f1 = sc.frame(1)
    ins = np_inputs(f1)
    vel = ins["velocity"].astype(np.float32)
    rng = np.random.default_rng(5)
    vel[..., 0] = rng.uniform(-1.5, 1.5, (H, W))
    vel[..., 1] = rng.uniform(-1.5, 1.5, (H, W))
    vel[0:4, :, 0] = 0.0
    vel[0:4, :, 1] = 0.0
    vel[10, :, 0] = 60000.0
    vel[11, :, 1] = -60000.0
    ins["velocity"] = vel.astype(np.float16)
    ins["velocity"].view(np.uint16)[12, 0:8, 0] = 0x7e00  # NaN
    ins["velocity"].view(np.uint16)[12, 8:16, 1] = 0x7c00  # +inf
    ins["velocity"].view(np.uint16)[12, 16:24, 0] = 0xfc00  # -inf
    hist = random_history(H, W, 9)
    for p in (configs.config3_full_chain(), with_params(configs.config3_full_chain(), mInterpolationMode=0),
              with_params(configs.config3_full_chain(), mInterpolationMode=1)):
        u = configs.uniforms_for(p, f1.jitter_ndc)
        check_exact(oracle, u, ins, hist, hist_depth=f1.depth.numpy())


@pytest.mark.parametrize("size", [(1, 1), (2, 3), (17, 5), (31, 33), (130, 9)])
def test_tiny_and_ragged_sizes(oracle, size):
    w, h = size
    sc = scene(w, h)
    f1 = sc.frame(1)
    ins = np_inputs(f1)
    hist = random_history(h, w, 1)
    for p in (configs.config1_defaults(), configs.config3_full_chain()):
        u = configs.uniforms_for(p, f1.jitter_ndc)
        check_exact(oracle, u, ins, hist, hist_depth=f1.depth.numpy())


def test_taa_upsampling(oracle):
    """Lo-res inputs, hi-res history (taa.comp:222-257): 2x and a non-integer factor."""
    for (iw, ih, ow, oh) in ((128, 72, 256, 144), (100, 60, 150, 90)):
        sc = scene(iw, ih)
        f1 = sc.frame(3)
        ins = np_inputs(f1)
        hist = random_history(oh, ow, 2)
        for p in (configs.config1_defaults(), configs.config2_resolve(), with_params(configs.config2_resolve(), mUnjitterFactor=-1.0)):
            u = configs.uniforms_for(p, f1.jitter_ndc, upsampling=True)
            check_exact(oracle, u, ins, hist, out_size=(ow, oh))


def test_noise_statistically(oracle):
    """sin() of large arguments is implementation-dependent: outside the bit-exact gate, bounded instead."""
    sc = scene()
    f1 = sc.frame(1)
    ins = np_inputs(f1)
    hist = random_history(H, W, 3)
    p = with_params(configs.config2_resolve(), mAddNoise=1)
    u = configs.uniforms_for(p, f1.jitter_ndc)
    u.mSinTime[0] = 0.37
    ref, got = check_exact(oracle, u, ins, hist, want=("history_out",), tol=2.5 * p.mNoiseFactor)
    p0 = with_params(configs.config2_resolve(), mAddNoise=0)
    base = oracle.resolve(configs.uniforms_for(p0, f1.jitter_ndc), ins["color"], ins["depth"], ins["velocity"], hist, want=("history_out",))
    n = got["history_out"][..., :3].astype(np.float32) - base["history_out"][..., :3].astype(np.float32)
    assert abs(float(n.mean())) < 2e-4 and 0.3 * p.mNoiseFactor < float(n.std()) < 0.8 * p.mNoiseFactor


def test_segmentation_mask(oracle):
    sc = scene()
    f0, f1 = sc.frame(6), sc.frame(7)
    ins = np_inputs(f1)
    hist = random_history(H, W, 4)
    prev_seg = (np.random.default_rng(8).integers(0, 6, (H, W)).astype(np.uint32) << 16) | 2
    all_flags = 0xffffffff & ~(abi.TAA_RTFLAG_ALL | abi.TAA_RTFLAG_FXD)
    for flags in (abi.TAA_RTFLAG_OUT, abi.TAA_RTFLAG_DIS, abi.TAA_RTFLAG_DPT, abi.TAA_RTFLAG_MID, abi.TAA_RTFLAG_LUM,
                  abi.TAA_RTFLAG_CNT | abi.TAA_RTFLAG_MID, abi.TAA_RTFLAG_ALL, abi.TAA_RTFLAG_FXD | abi.TAA_RTFLAG_MID,
                  all_flags & ~abi.TAA_RTFLAG_NRM):
        p = with_params(configs.config3_full_chain(), mRayTraceAugment=1, mRayTraceAugmentFlags=flags, mRayTraceHistoryCount=8,
                        mRayTraceAugment_WDpt=4.0, mRayTraceAugment_WLum=1.5)
        u = configs.uniforms_for(p, f1.jitter_ndc)
        check_exact(oracle, u, ins, hist, hist_depth=f0.depth.numpy(), prev_matid=f0.matid.numpy(), prev_segmask=prev_seg,
                    want=("history_out", "result", "mask", "segmask"))
    # The normal indicator goes through sin/cos, which GLSL leaves to the implementation: oracle, reference-shader shim and CUDA kernel share one
    # software sincos (taa_sincos), so the integer mask is bit-exact with the normal indicator too — alone, in the reference's default flag
    # set (taa.hpp:61: all indicators) and with a heavy weight that puts many pixels near the threshold.
    for flags, wnrm in ((abi.TAA_RTFLAG_NRM, 40.0), (abi.TAA_RTFLAG_NRM, 1.0), (all_flags, 1.0), (all_flags, 40.0), (0xffffffff & ~abi.TAA_RTFLAG_FXD, 4.0)):
        p = with_params(configs.config3_full_chain(), mRayTraceAugment=1, mRayTraceAugmentFlags=flags, mRayTraceAugment_WNrm=wnrm, mRayTraceHistoryCount=8)
        u = configs.uniforms_for(p, f1.jitter_ndc)
        check_exact(oracle, u, ins, hist, hist_depth=f0.depth.numpy(), prev_matid=f0.matid.numpy(), prev_segmask=prev_seg,
                    want=("history_out", "result", "mask", "segmask"))
    p = with_params(configs.config3_full_chain(), mRayTraceAugment=1, mRayTraceAugmentFlags=abi.TAA_RTFLAG_NRM, mRayTraceAugment_WNrm=40.0)
    u = configs.uniforms_for(p, f1.jitter_ndc)
    ref = oracle.resolve(u, ins["color"], ins["depth"], ins["velocity"], hist, history_depth=f0.depth.numpy(), matid=ins["matid"],
                         uvnrm=ins["uvnrm"], want=("segmask",))
    assert 0.01 < (ref["segmask"] != 0).mean() < 0.99


# ---- accumulated sequence -------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cfg", ["config2", "config3"])
def test_64_frame_sequence(oracle, cfg):
    """Free-running 64 frames on both sides (each feeds its own history): bit-identical at every frame, hence PSNR = inf >= 60 dB."""
    w, h = 192, 108
    sc = scene(w, h)
    p = {"config2": configs.config2_resolve, "config3": configs.config3_full_chain}[cfg]()
    ctx = host.TaaContext((w, h), flags=abi.TAA_FLAG_EXACT)
    hist_ref = np.zeros((h, w, 4), np.float16)
    hist_gpu = hist_ref.copy()
    prev_depth = None
    worst = 0.0
    for n in range(64):
        f = sc.frame(n)
        ins = np_inputs(f)
        u = configs.uniforms_for(p, f.jitter_ndc, reset_history=(n == 0))
        hd = prev_depth if prev_depth is not None else ins["depth"]
        ref = oracle.resolve(u, ins["color"], ins["depth"], ins["velocity"], hist_ref, history_depth=hd, want=("history_out", "result", "mask"))
        got = run_gpu_resolve(ctx, u, ins, hist_gpu, hist_depth=hd)
        d = np.abs(ref["result"].astype(np.float32) - got["result"].astype(np.float32)).max()
        worst = max(worst, float(d))
        assert d <= TOL_ABS, f"frame {n}: max |d| = {d}"
        assert (ref["mask"] == got["mask"]).all(), f"frame {n}: mask differs"
        hist_ref, hist_gpu = ref["history_out"], got["history_out"]
        prev_depth = ins["depth"]
    assert psnr(ref["result"][..., :3], got["result"][..., :3]) >= PSNR_MIN
    assert same_f16(ref["history_out"], got["history_out"]).all(), f"history diverged (worst per-frame |d| = {worst})"
    rej = (ref["mask"] & 1).mean()
    if cfg == "config3":
        assert 0.001 < rej < 0.5, f"rejection rate {rej}: the sequence should exercise disocclusion"


def test_converges_to_supersampled_reference(oracle):
    """Sanity of the whole loop (not a parity test): a static jittered scene accumulates towards the jitter-free image."""
    w, h = 160, 90
    sc = scene(w, h, pan_px=(0.0, 0.0), mover_px=(0.0, 0.0))
    p = configs.config2_resolve()
    ctx = host.TaaContext((w, h), flags=abi.TAA_FLAG_EXACT)
    hist = np.zeros((h, w, 4), np.float16)
    frames = [sc.frame(n) for n in range(48)]
    mean = np.mean([f.color.numpy().astype(np.float32) for f in frames[:8]], axis=0)
    for n, f in enumerate(frames):
        u = configs.uniforms_for(p, f.jitter_ndc, reset_history=(n == 0))
        got = run_gpu_resolve(ctx, u, np_inputs(f), hist)
        hist = got["history_out"]
    single = np.abs(frames[-1].color.numpy().astype(np.float32)[..., :3] - mean[..., :3]).mean()
    accum = np.abs(got["result"].astype(np.float32)[..., :3] - mean[..., :3]).mean()
    assert accum < 0.6 * single, (accum, single)


# ---- the north-star call and argument checking -------------------------------------------------------------------------------------
def test_taa_resolve_simple_signature(oracle):
    sc = scene()
    f1 = sc.frame(2)
    ins = np_inputs(f1)
    hist = random_history(H, W, 21)
    u = configs.uniforms_for(configs.config2_resolve(), f1.jitter_ndc)
    ref = oracle.resolve(u, ins["color"], ins["depth"], ins["velocity"], hist, want=("history_out",))
    ctx = host.TaaContext((W, H), flags=abi.TAA_FLAG_EXACT)
    out = torch.zeros(H, W, 4, dtype=torch.float16, device="cuda")
    ctx.resolve_simple(to_dev(ins["color"]), to_dev(ins["depth"]), to_dev(ins["velocity"]), to_dev(hist), out, u)
    torch.cuda.synchronize()
    assert same_f16(ref["history_out"], out.cpu().numpy()).all()
    assert ctx.launch_count == 1


def test_argument_errors():
    ctx = host.TaaContext((64, 32))
    u = configs.uniforms_for(configs.config3_full_chain())
    z8 = torch.zeros(32, 64, 4, dtype=torch.float16, device="cuda")
    z4 = torch.zeros(32, 64, dtype=torch.float32, device="cuda")
    with pytest.raises(abi.TaaError) as e:  # depth culling without history depth (taa.comp:818)
        ctx.resolve(u, color=z8, depth=z4, velocity=z8, history_in=z8, history_out=z8.clone())
    assert e.value.status == abi.TAA_E_INVALID_ARG and "history_depth" in str(e.value)
    u = configs.uniforms_for(configs.config2_resolve())
    with pytest.raises(abi.TaaError) as e:  # in-place history
        ctx.resolve(u, color=z8, depth=z4, velocity=z8, history_in=z8, history_out=z8)
    assert "alias" in str(e.value)
    with pytest.raises(abi.TaaError):  # missing required image
        ctx.resolve(u, color=z8, depth=z4, history_in=z8, history_out=z8.clone())
    narrow = torch.zeros(32, 32, 4, dtype=torch.float16, device="cuda")
    with pytest.raises(abi.TaaError) as e:  # pitch too small
        ctx.resolve(u, color=narrow, depth=z4, velocity=z8, history_in=z8, history_out=z8.clone())
    assert "pitch" in str(e.value)
    with pytest.raises(abi.TaaError):
        ctx.resolve(u, color=z8.cpu(), depth=z4, velocity=z8, history_in=z8, history_out=z8.clone())


def test_pitched_images(oracle):
    """Row pitch larger than the row: images as sub-rectangles of wider allocations."""
    sc = scene()
    f1 = sc.frame(1)
    ins = np_inputs(f1)
    hist = random_history(H, W, 5)
    u = configs.uniforms_for(configs.config2_resolve(), f1.jitter_ndc)
    ref = oracle.resolve(u, ins["color"], ins["depth"], ins["velocity"], hist, want=("history_out", "result"))

    def padded(a, pad):
        t = torch.from_numpy(a).cuda()
        shape = list(t.shape)
        shape[1] += pad
        big = torch.full(shape, 7, dtype=t.dtype, device="cuda")
        big[:, :t.shape[1]] = t
        return big[:, :t.shape[1]]

    ctx = host.TaaContext((W, H), flags=abi.TAA_FLAG_EXACT)
    ho = padded(np.zeros((H, W, 4), np.float16), 3)
    res = padded(np.zeros((H, W, 4), np.float16), 16)
    ctx.resolve(u, color=padded(ins["color"], 5), depth=padded(ins["depth"], 9), velocity=padded(ins["velocity"], 1), history_in=padded(hist, 2),
                history_out=ho, result=res)
    torch.cuda.synchronize()
    assert same_f16(ref["history_out"], ho.cpu().numpy()).all() and same_f16(ref["result"], res.cpu().numpy()).all()


# ---- row bands ------------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("nbands", [2, 3])
def test_row_bands_equal_whole_frame(oracle, nbands):
    """Band contexts with a halo of history/input rows reproduce the whole-frame result bit for bit."""
    sc = scene(pan_px=(3.0, 4.5))
    f0, f1 = sc.frame(3), sc.frame(4)
    ins = np_inputs(f1)
    hist = random_history(H, W, 13)
    u = configs.uniforms_for(configs.config3_full_chain(), f1.jitter_ndc)
    ref = oracle.resolve(u, ins["color"], ins["depth"], ins["velocity"], hist, history_depth=f0.depth.numpy(), want=("history_out", "result", "mask"))
    halo = 12
    rows = [(b * H // nbands, (b + 1) * H // nbands) for b in range(nbands)]
    for (y0, y1) in rows:
        a, b = max(0, y0 - halo), min(H, y1 + halo)
        ctx = host.TaaContext((W, H), band=(y0, y1 - y0), flags=abi.TAA_FLAG_EXACT)
        out = {k: torch.zeros(y1 - y0, W, 4, dtype=torch.float16, device="cuda") for k in ("history_out", "result")}
        mask = torch.zeros(y1 - y0, W, dtype=torch.int32, device="cuda")
        sl = lambda arr: (to_dev(arr[a:b]), a)
        ctx.resolve(u, color=sl(ins["color"]), depth=sl(ins["depth"]), velocity=sl(ins["velocity"]), history_in=sl(hist),
                    history_depth=sl(f0.depth.numpy()), history_out=(out["history_out"], y0), result=(out["result"], y0), mask=(mask, y0))
        assert ctx.poll_status() == abi.TAA_OK
        for k in out:
            assert same_f16(ref[k][y0:y1], out[k].cpu().numpy()).all(), (k, y0, y1)
        assert (ref["mask"][y0:y1] == mask.cpu().numpy().view(np.uint32)).all()
        ctx.close()


def test_band_halo_overflow_is_reported():
    sc = scene(pan_px=(0.0, 40.0))
    f1 = sc.frame(1)
    ins = np_inputs(f1)
    hist = random_history(H, W, 13)
    u = configs.uniforms_for(configs.config2_resolve(), f1.jitter_ndc)
    y0, y1, halo = 48, 96, 4
    ctx = host.TaaContext((W, H), band=(y0, y1 - y0), flags=abi.TAA_FLAG_EXACT)
    a, b = y0 - halo, y1 + halo
    sl = lambda arr: (to_dev(arr[a:b]), a)
    out = torch.zeros(y1 - y0, W, 4, dtype=torch.float16, device="cuda")
    ctx.resolve(u, color=sl(ins["color"]), depth=sl(ins["depth"]), velocity=sl(ins["velocity"]), history_in=sl(hist), history_out=(out, y0))
    assert ctx.poll_status() == abi.TAA_E_HALO_OVERFLOW
    assert ctx.poll_status() == abi.TAA_OK  # cleared by the read


# ---- the reference's default settings (BASELINE configs[0]) on the specialised exact kernel -----------------------------------------------
@pytest.mark.parametrize("size", [(256, 144), (1, 1), (33, 9), (31, 17), (97, 41), (640, 360)])
def test_defaults_kernel_is_bit_identical(oracle, size):
    """Default context flags + the reference's default switch pattern = taa_resolve_defaults_kernel (shared-memory colour taps, folded
    switches). It is the general kernel's arithmetic: every output bit-identical to the oracle, masks and the debug image included; ragged
    sizes, float parameters away from their defaults, history reset."""
    w, h = size
    sc = scene(w, h, pan_px=(2.75, -1.5))
    f0, f1 = sc.frame(2), sc.frame(3)
    for p, reset in ((configs.config1_defaults(), False), (with_params(configs.config1_defaults(), mAlpha=0.2, mDebugMode=2, mDebugScale=3.0), False),
                     (configs.config1_defaults(), True)):
        u = configs.uniforms_for(p, f1.jitter_ndc, reset_history=reset)
        # matrices that move static pixels (the defaults reproject them with the matrices, taa.comp:426-430)
        u.mHistoryViewProjMatrix[12] = 0.004
        u.mHistoryViewProjMatrix[13] = -0.003
        ctx = host.TaaContext((w, h))  # default flags: the specialised kernel is allowed
        n0 = ctx.launch_count
        check_exact(oracle, u, np_inputs(f1), random_history(h, w, 11), want=("history_out", "result", "mask", "debug"), ctx=ctx)
        assert ctx.launch_count - n0 == 1
        ctx.close()


def test_defaults_kernel_equals_general_kernel_on_bands():
    """Row bands of the specialised kernel (tile rows clipped at the band) against the general kernel on the whole frame, bit for bit."""
    w, h = 320, 200
    dev = torch.device("cuda")
    sc = SyntheticScene(w, h, device=dev, with_aux=False)
    f1 = sc.frame(5)
    u = configs.uniforms_for(configs.config1_defaults(), f1.jitter_ndc)
    hist = sc.frame(4).color.clone()
    outs = {}
    ctx = host.TaaContext((w, h), flags=abi.TAA_FLAG_EXACT)
    outs["general"] = {k: torch.zeros(h, w, 4, dtype=torch.float16, device=dev) for k in ("history_out", "result")}
    ctx.resolve(u, color=f1.color, depth=f1.depth, velocity=f1.velocity, history_in=hist, **outs["general"])
    ctx.close()
    outs["bands"] = {k: torch.zeros(h, w, 4, dtype=torch.float16, device=dev) for k in ("history_out", "result")}
    for a, b in ((0, 67), (67, 70), (70, 200)):
        c = host.TaaContext((w, h), band=(a, b - a))
        c.resolve(u, color=f1.color, depth=f1.depth, velocity=f1.velocity, history_in=hist, **outs["bands"])
        c.close()
    torch.cuda.synchronize()
    for k in ("history_out", "result"):
        assert torch.equal(outs["general"][k].view(torch.int16), outs["bands"][k].view(torch.int16)), k
