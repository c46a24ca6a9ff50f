"""GPU parity of the DEFAULT path: the tuned shared-memory kernel + exact fix-up pass (taa_resolve_tuned.cu) against the CPU oracle.

Bar (BASELINE.json north_star): every rgba16f output within 2^-10 per channel per frame (same inputs, same history on both
sides), PSNR >= 60 dB after 64 free-running frames, integer masks bit-exact. Full-size (3840x2160 and a 15360-wide strip) checks
use the exact GPU kernel — itself bit-identical to the oracle (test_parity_gpu.py) — as the reference, plus oracle row samples.
"""
import numpy as np
import pytest
import torch

from common import PSNR_MIN, TOL_ABS, np_inputs, psnr, random_history, run_gpu_resolve, to_dev
from taa_star_b200 import abi, configs, host
from taa_star_b200.synth import SyntheticScene

pytestmark = pytest.mark.gpu

W, H = 256, 144
CFG = {"config2": configs.config2_resolve, "config3": configs.config3_full_chain}


def with_params(base, **kw):
    p = abi.TaaParameters.from_buffer_copy(base)
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def max_abs_nan_aware(a, b):
    fa, fb = a.astype(np.float32), b.astype(np.float32)
    both_nan = np.isnan(fa) & np.isnan(fb)
    both_same_inf = np.isinf(fa) & (fa == fb)
    d = np.abs(fa - fb)
    d[both_nan | both_same_inf] = 0.0
    return float(np.nan_to_num(d, nan=np.inf).max())


def check_tuned(oracle, u, ins, hist, hist_depth=None, ctx=None, want=("history_out", "result", "mask"), expect_tuned=True, expect_launches=None):
    in_h, in_w = ins["depth"].shape
    ref = oracle.resolve(u, ins["color"], ins["depth"], ins["velocity"], hist, history_depth=hist_depth, want=want)
    own = ctx is None
    if own:
        ctx = host.TaaContext((in_w, in_h))
    n0 = ctx.launch_count
    got = run_gpu_resolve(ctx, u, ins, hist, hist_depth=hist_depth, want=want)
    launches = ctx.launch_count - n0
    if expect_launches is None:
        expect_launches = 2 if expect_tuned else 1
    assert launches == expect_launches, f"{launches} launches: the {'tuned' if expect_tuned else 'general'} path ({expect_launches}) was expected"
    fix = ctx.fixup_pixels()
    if own:
        ctx.close()
    worst = 0.0
    for name in want:
        if name == "mask":
            bad = int((ref[name] != got[name]).sum())
            assert bad == 0, f"mask: {bad} of {ref[name].size} values differ (fix-up list held {fix} pixels)"
        else:
            d = max_abs_nan_aware(ref[name], got[name])
            worst = max(worst, d)
            assert d <= TOL_ABS, f"{name}: max |d| = {d} > 2^-10"
    return ref, got, worst, fix


@pytest.mark.parametrize("cfg", ["config2", "config3"])
def test_single_frame(oracle, cfg):
    sc = SyntheticScene(W, H)
    f0, f1 = sc.frame(4), sc.frame(5)
    u = configs.uniforms_for(CFG[cfg](), f1.jitter_ndc)
    for seed in (7, 8):
        check_tuned(oracle, u, np_inputs(f1), random_history(H, W, seed), hist_depth=f0.depth.numpy())


SWITCHES = [dict(mAlpha=0.3), dict(mVarClipGamma=0.75), dict(mVarClipGamma=1.5), dict(mRejectOutside=1), dict(mRejectOutside=1, mRejectionAlpha=0.5),
            dict(mDepthCulling=1), dict(mDynamicAntiGhosting=1), dict(mVelBasedAlpha=1, mVelBasedAlphaFactor=40.0), dict(mLumaWeightingLottes=1),
            dict(mReduceBlendNearClamp=1), dict(mReduceBlendNearClamp=1, mLumaWeightingLottes=1, mDepthCulling=1)]


@pytest.mark.parametrize("sw", SWITCHES, ids=lambda s: ",".join(f"{k}={v}" for k, v in s.items()))
def test_each_switch_of_the_family(oracle, sw):
    sc = SyntheticScene(W, H, pan_px=(5.25, -2.5))
    f0, f1 = sc.frame(2), sc.frame(3)
    u = configs.uniforms_for(with_params(configs.config2_resolve(), **sw), f1.jitter_ndc)
    check_tuned(oracle, u, np_inputs(f1), random_history(H, W, 11), hist_depth=f0.depth.numpy())


def test_reset_history(oracle):
    sc = SyntheticScene(W, H)
    f1 = sc.frame(1)
    u = configs.uniforms_for(configs.config2_resolve(), f1.jitter_ndc, reset_history=True)
    check_tuned(oracle, u, np_inputs(f1), random_history(H, W, 3))


@pytest.mark.parametrize("sw", [dict(mInterpolationMode=1), dict(mUseYCoCg=0), dict(mColorClampingOrClipping=1), dict(mUseVelocityVectors=1),
                                dict(mToneMapLumaKaris=1), dict(mVelocitySampleMode=2), dict(mUnjitterNeighbourhood=1)],
                         ids=lambda s: ",".join(f"{k}={v}" for k, v in s.items()))
def test_settings_outside_the_family_run_on_the_general_kernel(oracle, sw):
    sc = SyntheticScene(W, H)
    f1 = sc.frame(2)
    u = configs.uniforms_for(with_params(configs.config2_resolve(), **sw), f1.jitter_ndc)
    _, _, worst, _ = check_tuned(oracle, u, np_inputs(f1), random_history(H, W, 5), expect_tuned=False)
    assert worst == 0.0


@pytest.mark.parametrize("size", [(1, 1), (2, 3), (17, 5), (31, 33), (33, 31), (130, 9), (300, 70)])
def test_tiny_and_ragged_sizes(oracle, size):
    w, h = size
    sc = SyntheticScene(w, h)
    f1 = sc.frame(1)
    for p in (configs.config2_resolve(), configs.config3_full_chain()):
        u = configs.uniforms_for(p, f1.jitter_ndc)
        check_tuned(oracle, u, np_inputs(f1), random_history(h, w, 1), hist_depth=f1.depth.numpy())


def test_extreme_and_non_finite_motion(oracle):
    sc = SyntheticScene(W, H, pan_px=(0.0, 0.0))
    f1 = sc.frame(1)
    ins = np_inputs(f1)
    vel = ins["velocity"].astype(np.float32)
    rng = np.random.default_rng(5)
    vel[..., 0] = rng.uniform(-1.5, 1.5, (H, W))
    vel[..., 1] = rng.uniform(-1.5, 1.5, (H, W))
    vel[0:4, :, 0:2] = 0.0
    vel[4:8, :, 0] = rng.uniform(-3, 3, (4, W)) / W   # a few pixels of motion, sub-pixel phases everywhere
    vel[4:8, :, 1] = rng.uniform(-3, 3, (4, W)) / H
    vel[10, :, 0] = 60000.0
    vel[11, :, 1] = -60000.0
    ins["velocity"] = vel.astype(np.float16)
    ins["velocity"].view(np.uint16)[12, 0:8, 0] = 0x7e00   # NaN
    ins["velocity"].view(np.uint16)[12, 8:16, 1] = 0x7c00  # +inf
    ins["velocity"].view(np.uint16)[12, 16:24, 0] = 0xfc00  # -inf
    hist = random_history(H, W, 9)
    for p in (configs.config2_resolve(), configs.config3_full_chain()):
        u = configs.uniforms_for(p, f1.jitter_ndc)
        check_tuned(oracle, u, ins, hist, hist_depth=f1.depth.numpy())


def test_static_camera_subtexel_bleed(oracle):
    """Zero motion puts every history tap on a texel centre (f ~ 0 or ~ 1), where the sampler's coordinate rounding decides which neighbour
    bleeds in; a wide frame makes that rounding large (SURVEY 'sampler-exact semantics')."""
    w, h = 4000, 24
    sc = SyntheticScene(w, h, pan_px=(0.0, 0.0), mover_px=(0.0, 0.0))
    f1 = sc.frame(3)
    u = configs.uniforms_for(configs.config2_resolve(), f1.jitter_ndc)
    check_tuned(oracle, u, np_inputs(f1), random_history(h, w, 2))


@pytest.mark.parametrize("cfg", ["config2", "config3"])
def test_64_frame_sequence(oracle, cfg):
    """Per frame: same inputs and same history on both sides -> within 2^-10, masks bit-exact. Free-running on the GPU beside a free-running
    oracle: PSNR >= 60 dB after 64 accumulated frames."""
    w, h = 192, 108
    sc = SyntheticScene(w, h)
    p = CFG[cfg]()
    ctx = host.TaaContext((w, h))
    hist_ref = np.zeros((h, w, 4), np.float16)
    hist_free = hist_ref.copy()
    prev_depth = None
    worst, fix_total = 0.0, 0
    for n in range(64):
        f = sc.frame(n)
        ins = np_inputs(f)
        u = configs.uniforms_for(p, f.jitter_ndc, reset_history=(n == 0))
        hd = prev_depth if prev_depth is not None else ins["depth"]
        ref, got, d, fix = check_tuned(oracle, u, ins, hist_ref, hist_depth=hd, ctx=ctx)
        worst = max(worst, d)
        fix_total += fix
        free = run_gpu_resolve(ctx, u, ins, hist_free, hist_depth=hd)
        hist_ref, hist_free = ref["history_out"], free["history_out"]
        prev_depth = ins["depth"]
    q = psnr(ref["result"][..., :3], free["result"][..., :3])
    assert q >= PSNR_MIN, f"PSNR after 64 free-running frames: {q:.1f} dB"
    d_free = max_abs_nan_aware(ref["result"], free["result"])
    print(f"{cfg}: worst per-frame |d| = {worst:.3e}, free-running |d| after 64 frames = {d_free:.3e}, PSNR = {q:.1f} dB, "
          f"fix-up pixels per frame = {fix_total / 64:.0f} of {w * h}")
    if cfg == "config3":
        assert 0.001 < (ref["mask"] & 1).mean() < 0.5


@pytest.mark.parametrize("nbands", [2, 3])
def test_row_bands_equal_whole_frame(oracle, nbands):
    sc = SyntheticScene(W, H, pan_px=(3.0, 4.5))
    f0, f1 = sc.frame(3), sc.frame(4)
    ins = np_inputs(f1)
    hist = random_history(H, W, 13)
    u = configs.uniforms_for(configs.config3_full_chain(), f1.jitter_ndc)
    ref, whole, _, _ = check_tuned(oracle, u, ins, hist, hist_depth=f0.depth.numpy())
    halo = 12
    for b in range(nbands):
        y0, y1 = b * H // nbands, (b + 1) * H // nbands
        a, e = max(0, y0 - halo), min(H, y1 + halo)
        ctx = host.TaaContext((W, H), band=(y0, y1 - y0))
        out = {k: torch.zeros(y1 - y0, W, 4, dtype=torch.float16, device="cuda") for k in ("history_out", "result")}
        mask = torch.zeros(y1 - y0, W, dtype=torch.int32, device="cuda")
        sl = lambda arr: (to_dev(arr[a:e]), a)
        ctx.resolve(u, color=sl(ins["color"]), depth=sl(ins["depth"]), velocity=sl(ins["velocity"]), history_in=sl(hist),
                    history_depth=sl(f0.depth.numpy()), history_out=(out["history_out"], y0), result=(out["result"], y0), mask=(mask, y0))
        assert ctx.poll_status() == abi.TAA_OK
        assert ctx.launch_count == 2
        for k in out:  # a band runs the same per-pixel arithmetic as the whole frame
            assert (out[k].cpu().numpy().view(np.uint16) == whole[k][y0:y1].view(np.uint16)).all(), (k, y0, y1)
        assert (ref["mask"][y0:y1] == mask.cpu().numpy().view(np.uint32)).all()
        ctx.close()


def test_band_halo_overflow_is_reported():
    sc = SyntheticScene(W, H, pan_px=(0.0, 40.0))
    f1 = sc.frame(1)
    ins = np_inputs(f1)
    hist = random_history(H, W, 13)
    u = configs.uniforms_for(configs.config2_resolve(), f1.jitter_ndc)
    y0, y1, halo = 48, 96, 4
    ctx = host.TaaContext((W, H), band=(y0, y1 - y0))
    a, b = y0 - halo, y1 + halo
    sl = lambda arr: (to_dev(arr[a:b]), a)
    out = torch.zeros(y1 - y0, W, 4, dtype=torch.float16, device="cuda")
    ctx.resolve(u, color=sl(ins["color"]), depth=sl(ins["depth"]), velocity=sl(ins["velocity"]), history_in=sl(hist), history_out=(out, y0))
    assert ctx.poll_status() == abi.TAA_E_HALO_OVERFLOW


def _gpu_frame_pair(w, h, device, n=9, **kw):
    sc = SyntheticScene(w, h, device=device, with_aux=False, **kw)
    return sc.frame(n - 1), sc.frame(n)


@pytest.mark.parametrize("cfg", ["config2", "config3"])
def test_full_size_4k_against_exact_kernel_and_oracle_rows(oracle, cfg):
    """3840x2160 (BASELINE configs[1], [2]): tuned path vs the exact GPU kernel on the whole frame (masks identical, colour within 2^-10),
    and vs the oracle on sampled row bands."""
    w, h = 3840, 2160
    dev = torch.device("cuda")
    f0, f1 = _gpu_frame_pair(w, h, dev)
    u = configs.uniforms_for(CFG[cfg](), f1.jitter_ndc)
    hist = f0.color.clone()
    hist[..., 3] = (torch.rand(h, w, device=dev) > 0.8).to(torch.float16)
    outs = {}
    for name, flags in (("tuned", 0), ("exact", abi.TAA_FLAG_EXACT)):
        ctx = host.TaaContext((w, h), flags=flags)
        o = {k: torch.zeros(h, w, 4, dtype=torch.float16, device=dev) for k in ("history_out", "result")}
        o["mask"] = torch.zeros(h, w, dtype=torch.int32, device=dev)
        ctx.resolve(u, color=f1.color, depth=f1.depth, velocity=f1.velocity, history_in=hist, history_depth=f0.depth, **o)
        torch.cuda.synchronize()
        if name == "tuned":
            fix = ctx.fixup_pixels()
        outs[name] = o
        ctx.close()
    assert torch.equal(outs["tuned"]["mask"], outs["exact"]["mask"]), \
        f"{int((outs['tuned']['mask'] != outs['exact']['mask']).sum())} mask values differ at 4K ({fix} pixels went through the fix-up)"
    for k in ("history_out", "result"):
        d = float((outs["tuned"][k].float() - outs["exact"][k].float()).abs().max())
        assert d <= TOL_ABS, f"{k}: max |d| = {d}"
    print(f"{cfg} 4K: fix-up pixels = {fix} ({100.0 * fix / (w * h):.3f} %), rectified = {float((outs['exact']['mask'] & 2).bool().float().mean()):.3f}")
    # oracle on three row bands of the same frame
    ins = dict(color=f1.color.cpu().numpy(), depth=f1.depth.cpu().numpy(), velocity=f1.velocity.cpu().numpy())
    hist_np, hd_np = hist.cpu().numpy(), f0.depth.cpu().numpy()
    for ya in (0, 1000, h - 24):
        ref = oracle.resolve(u, ins["color"], ins["depth"], ins["velocity"], hist_np, history_depth=hd_np, want=("history_out", "mask"), rows=(ya, ya + 24))
        got = outs["tuned"]["history_out"][ya:ya + 24].cpu().numpy()
        assert max_abs_nan_aware(ref["history_out"][ya:ya + 24], got) <= TOL_ABS
        assert (ref["mask"][ya:ya + 24] == outs["tuned"]["mask"][ya:ya + 24].cpu().numpy().view(np.uint32)).all()


def test_16k_wide_strip(oracle):
    """A 15360-wide strip (BASELINE configs[3] width): the sampler's coordinate rounding is 4x that of 4K."""
    w, h = 15360, 40
    sc = SyntheticScene(w, h, pan_px=(3.0, 0.5))
    f0, f1 = sc.frame(4), sc.frame(5)
    u = configs.uniforms_for(configs.config2_resolve(), f1.jitter_ndc)
    check_tuned(oracle, u, np_inputs(f1), f0.color.numpy().copy())


@pytest.mark.parametrize("sw", [dict(), dict(mVarClipGamma=0.75, mAlpha=0.3), dict(mRejectOutside=1, mDepthCulling=1, mDynamicAntiGhosting=1),
                                dict(mVelBasedAlpha=1, mVelBasedAlphaFactor=40.0, mLumaWeightingLottes=1, mReduceBlendNearClamp=1), "config3"],
                         ids=lambda s: s if isinstance(s, str) else ",".join(f"{k}={v}" for k, v in s.items()) or "config2")
def test_fixup_pass_is_bit_exact(oracle, sw):
    """TAA_FLAG_FIXUP_ALL sends every pixel through the fix-up pass (the family-specialised exact arithmetic): all outputs must then be
    bit-identical to the oracle, including ragged sizes, borders, movers and non-finite motion."""
    from common import mismatch_report
    for (w, h, pan) in ((256, 144, (5.25, -2.5)), (67, 35, (0.0, 0.0)), (300, 70, (-17.0, 9.0))):
        sc = SyntheticScene(w, h, pan_px=pan)
        f0, f1 = sc.frame(2), sc.frame(3)
        p = configs.config3_full_chain() if sw == "config3" else with_params(configs.config2_resolve(), **sw)
        u = configs.uniforms_for(p, f1.jitter_ndc)
        ins, hist = np_inputs(f1), random_history(h, w, 11)
        if w == 300:
            ins["velocity"].view(np.uint16)[12, 0:8, 0] = 0x7e00
            ins["velocity"].view(np.uint16)[13, 8:16, 1] = 0x7c00
            ins["velocity"][14, :, 0] = 60000.0
        ref = oracle.resolve(u, ins["color"], ins["depth"], ins["velocity"], hist, history_depth=f0.depth.numpy(), want=("history_out", "result", "mask"))
        ctx = host.TaaContext((w, h), flags=abi.TAA_FLAG_FIXUP_ALL)
        got = run_gpu_resolve(ctx, u, ins, hist, hist_depth=f0.depth.numpy())
        assert ctx.launch_count == 2 and ctx.fixup_pixels() == w * h
        for k in ("history_out", "result"):
            r = mismatch_report(k, ref[k], got[k])
            assert r is None, r
        assert (ref["mask"] == got["mask"]).all()


# ---- paths of the strip kernel (taa_resolve_strip.cu) ------------------------------------------------------------------------
@pytest.mark.parametrize("cfg", ["config2", "config3"])
def test_smoothly_varying_motion_runs_the_general_strip_path(oracle, cfg):
    """No tile is uniform and no column keeps its history u: window restarts, per-pixel weights, the mover mask of config 3."""
    w, h = 320, 200
    sc = SyntheticScene(w, h)
    f0, f1 = sc.frame(2), sc.frame(3)
    ins = np_inputs(f1)
    yy, xx = np.meshgrid(np.arange(h, dtype=np.float32), np.arange(w, dtype=np.float32), indexing="ij")
    gain = 1.0 + 0.25 * np.sin(xx * 0.11) * np.cos(yy * 0.13)
    vel = ins["velocity"].astype(np.float32)
    vel[..., 0:2] *= gain[..., None]
    ins["velocity"] = vel.astype(np.float16)
    u = configs.uniforms_for(CFG[cfg](), f1.jitter_ndc)
    check_tuned(oracle, u, ins, random_history(h, w, 21), hist_depth=f0.depth.numpy())


@pytest.mark.parametrize("pan", [(0.0, 0.0), (3.0, 0.5), (-2.25, -1.75), (0.5, 7.0), (40.0, -30.0)])
def test_uniform_motion_tiles(oracle, pan):
    """Uniform-motion fast path: zero motion, sub-pixel phases, footprints that leave the image on every side, no mover in the scene."""
    w, h = 256, 160
    sc = SyntheticScene(w, h, pan_px=pan, mover_px=pan)  # the foreground quad moves with the background: one velocity everywhere
    f0, f1 = sc.frame(6), sc.frame(7)
    ins = np_inputs(f1)
    ins["velocity"][..., 0:2] = ins["velocity"][h // 2, w // 2, 0:2]  # bit-identical texels, whatever the generator rounds to
    ins["velocity"][..., 3] = 0
    for cfg in ("config2", "config3"):
        u = configs.uniforms_for(CFG[cfg](), f1.jitter_ndc)
        check_tuned(oracle, u, ins, random_history(h, w, 22), hist_depth=f0.depth.numpy())


def test_uniform_tiles_next_to_a_mover_and_history_alpha(oracle):
    """Config 3: a history whose alpha marks a moving region (anti-ghost ring test) under uniform motion, and velocity.w set in one tile."""
    w, h = 256, 160
    sc = SyntheticScene(w, h, pan_px=(2.0, 1.0), mover_px=(2.0, 1.0))
    f0, f1 = sc.frame(3), sc.frame(4)
    ins = np_inputs(f1)
    ins["velocity"][..., 0:2] = ins["velocity"][h // 2, w // 2, 0:2]
    ins["velocity"][..., 3] = 0
    ins["velocity"][40:56, 70:90, 3] = 1.0   # a mover patch: its tile and the neighbours within two texels take the general path
    hist = random_history(h, w, 23)
    hist[..., 3] = 0
    hist[60:100, 100:180, 3] = 1.0           # dynamic mask of the previous frame
    hist[10, 5:9, 3] = 1.0
    u = configs.uniforms_for(CFG["config3"](), f1.jitter_ndc)
    check_tuned(oracle, u, ins, hist, hist_depth=f0.depth.numpy())


@pytest.mark.parametrize("env", [dict(TAA_TUNED_VARIANT="strip"), dict(TAA_STREAM_REJ="1")], ids=["strip", "stream-rej"])
def test_other_kernels_of_the_family_in_a_subprocess(env):
    """The default dispatch sends the plain variants to the streaming kernel and the rejection variants to the strip kernel. The other
    pairing honours the same contract: the strip kernel on everything (TAA_TUNED_VARIANT=strip), the streaming kernel on the rejection
    variants too (TAA_STREAM_REJ=1)."""
    import os
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(here, "test_tuned_gpu.py"), "-q", "-x", "-m", "gpu", "-k",
                        "test_single_frame or test_each_switch_of_the_family or test_tiny_and_ragged_sizes or test_extreme_and_non_finite_motion or "
                        "test_fixup_pass_is_bit_exact or test_row_bands_equal_whole_frame or test_uniform_motion_tiles or test_64_frame_sequence"],
                       env=dict(os.environ, **env), capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("cfg,launches", [("config2", 1), ("config3", 2)])
def test_without_a_mask_binding(oracle, cfg, launches):
    """`rectified` is reported through the mask alone: without a mask binding the fix-up pass runs only where a rejection predicate can be
    undecided (dynamic anti-ghosting, config 3); config 2 is the tuned kernel alone. Colours stay within 2^-10 per frame and >= 60 dB
    free-running, on the uniform-motion path and on the general one."""
    w, h = 256, 144
    p = CFG[cfg]()
    for pan, vary in (((3.0, 0.5), False), ((2.0, -1.5), True)):
        sc = SyntheticScene(w, h, pan_px=pan)
        ctx = host.TaaContext((w, h))
        hist_ref = np.zeros((h, w, 4), np.float16)
        hist_free = hist_ref.copy()
        prev_depth = None
        for n in range(24):
            f = sc.frame(n)
            ins = np_inputs(f)
            if vary:
                yy, xx = np.meshgrid(np.arange(h, dtype=np.float32), np.arange(w, dtype=np.float32), indexing="ij")
                vel = ins["velocity"].astype(np.float32)
                vel[..., 0:2] *= (1.0 + 0.25 * np.sin(xx * 0.11 + n) * np.cos(yy * 0.13))[..., None]
                ins["velocity"] = vel.astype(np.float16)
            u = configs.uniforms_for(p, f.jitter_ndc, reset_history=(n == 0))
            hd = prev_depth if prev_depth is not None else ins["depth"]
            ref, got, d, _ = check_tuned(oracle, u, ins, hist_ref, hist_depth=hd, ctx=ctx, want=("history_out", "result"), expect_launches=launches)
            free = run_gpu_resolve(ctx, u, ins, hist_free, hist_depth=hd, want=("history_out", "result"))
            hist_ref, hist_free = ref["history_out"], free["history_out"]
            prev_depth = ins["depth"]
        q = psnr(ref["result"][..., :3], free["result"][..., :3])
        assert q >= PSNR_MIN, f"PSNR after 24 free-running frames: {q:.1f} dB"
        ctx.close()


def test_slow_units_first_changes_nothing_but_the_order():
    """The streaming kernel starts the units that were slow in the context's previous call first (hint buffers of the context). Whatever the
    hints hold, every unit is resolved exactly once: a context that has seen other frames (movers elsewhere, so its hints point at other units),
    a context that has seen the same frame, and a fresh context give the same bits; outputs are pre-filled with a sentinel so a unit that nobody
    resolved would show."""
    w, h = 1920, 1080
    dev = torch.device("cuda")
    p = configs.config2_resolve()
    scenes = [SyntheticScene(w, h, device=dev, with_aux=False), SyntheticScene(w, h, device=dev, with_aux=False, pan_px=(-2.0, 1.25), mover_px=(9.0, 3.0))]
    scenes[1].m0 = (0.3 * w, 0.7 * h)
    frames = [sc.frame(n) for sc in scenes for n in (3, 4)]
    hist = frames[0].color.clone()

    def run(ctx, f):
        o = {k: torch.full((h, w, 4), 777.0, dtype=torch.float16, device=dev) for k in ("history_out", "result")}
        ctx.resolve(configs.uniforms_for(p, f.jitter_ndc), color=f.color, depth=f.depth, velocity=f.velocity, history_in=hist, **o)
        torch.cuda.synchronize()
        assert not bool((o["history_out"] == 777.0).any()), "a unit was left unresolved"
        return o

    target = frames[1]
    fresh = host.TaaContext((w, h))
    ref = run(fresh, target)
    fresh.close()
    seasoned = host.TaaContext((w, h))
    for f in (frames[2], frames[3], frames[2], target, target, frames[3], target):
        got = run(seasoned, f)
        if f is target:
            for k in ref:
                assert torch.equal(ref[k], got[k]), f"{k} depends on the scheduling hints"
    seasoned.close()


def test_4k_eight_frames_free_running(oracle):
    """3840x2160, 8 free-running frames (the default path feeds on its own history) against the exact kernel doing the same: PSNR >= 60 dB and
    every frame within 2^-10 of the exact kernel run on the same history."""
    w, h = 3840, 2160
    dev = torch.device("cuda")
    sc = SyntheticScene(w, h, device=dev, with_aux=False)
    p = configs.config2_resolve()
    tuned, exact = host.TaaContext((w, h)), host.TaaContext((w, h), flags=abi.TAA_FLAG_EXACT)
    ht = [torch.zeros(h, w, 4, dtype=torch.float16, device=dev) for _ in range(2)]
    he = [torch.zeros(h, w, 4, dtype=torch.float16, device=dev) for _ in range(2)]
    rt, re_, chk = (torch.zeros(h, w, 4, dtype=torch.float16, device=dev) for _ in range(3))
    for n in range(8):
        f = sc.frame(n)
        u = configs.uniforms_for(p, f.jitter_ndc, reset_history=(n == 0))
        a, b = n & 1, (n & 1) ^ 1
        exact.resolve(u, color=f.color, depth=f.depth, velocity=f.velocity, history_in=ht[a], history_out=chk)   # exact arithmetic on the tuned path's own history
        tuned.resolve(u, color=f.color, depth=f.depth, velocity=f.velocity, history_in=ht[a], history_out=ht[b], result=rt)
        exact.resolve(u, color=f.color, depth=f.depth, velocity=f.velocity, history_in=he[a], history_out=he[b], result=re_)
        torch.cuda.synchronize()
        d = float((ht[b].float() - chk.float()).abs().max())
        assert d <= TOL_ABS, f"frame {n}: max |d| = {d} against the exact kernel on the same history"
    mse = float(((rt[..., :3].float() - re_[..., :3].float()) ** 2).mean())
    q = 99.0 if mse == 0 else 10.0 * np.log10(1.0 / mse)
    assert q >= PSNR_MIN, f"PSNR after 8 free-running 4K frames: {q:.1f} dB"
    tuned.close(); exact.close()


def test_band_peers_on_one_gpu():
    """taa_band_peers without a second GPU: two band contexts of one frame in ONE process, each registered as the other's neighbour (plain device
    pointers instead of IPC mappings), driven on two CUDA streams with no synchronisation between them. The resolve kernel stores each band's
    boundary rows into the other band's halo and the flags order the frames. After 7 frames both bands — halo rows included — equal the
    whole-frame resolve bit for bit (motion that is not a whole number of texels)."""
    import ctypes as C
    w, h, halo = 1920, 1080, 20
    dev = torch.device("cuda")
    lib = abi.load_library()
    sc = SyntheticScene(w, h, device=dev, with_aux=False, pan_px=(3.25, 7.5), mover_px=(-6.5, 5.25))
    p = configs.config2_resolve()
    bands = [(0, h // 2), (h // 2, h - h // 2)]
    hy = [(max(0, a - halo), min(h, a + n + halo)) for a, n in bands]
    hist = [[torch.zeros(b - a, w, 4, dtype=torch.float16, device=dev) for _ in range(2)] for a, b in hy]
    flags = [torch.zeros(abi.TAA_BAND_FLAG_WORDS, dtype=torch.int32, device=dev) for _ in bands]
    res = [torch.zeros(n, w, 4, dtype=torch.float16, device=dev) for _, n in bands]
    ctxs = [host.TaaContext((w, h), band=b) for b in bands]
    streams = [torch.cuda.Stream() for _ in bands]
    torch.cuda.synchronize()
    for r in range(2):
        o = 1 - r
        pb = abi.taa_band_peer()
        pb.history[0], pb.history[1] = hist[o][0].data_ptr(), hist[o][1].data_ptr()
        pb.row_pitch, pb.y0, pb.band_rows, pb.flags = w * 8, hy[o][0], bands[o][1], flags[o].data_ptr()
        up, dn = (C.byref(pb), None) if r == 1 else (None, C.byref(pb))  # band 1's neighbour lies above it, band 0's below
        st = lib.taa_band_peers(ctxs[r]._h, up, dn, hist[r][0].data_ptr(), hist[r][1].data_ptr(), flags[r].data_ptr(), halo)
        assert st == abi.TAA_OK, lib.taa_last_error_string(ctxs[r]._h).decode()
    whole = host.TaaContext((w, h))
    wh = [torch.zeros(h, w, 4, dtype=torch.float16, device=dev) for _ in range(2)]
    wres = torch.zeros(h, w, 4, dtype=torch.float16, device=dev)
    nframes = 7
    frames = [sc.frame(n) for n in range(nframes)]
    torch.cuda.synchronize()
    for n, f in enumerate(frames):
        u = configs.uniforms_for(p, f.jitter_ndc, reset_history=(n == 0))
        whole.resolve(u, color=f.color, depth=f.depth, velocity=f.velocity, history_in=wh[n & 1], history_out=wh[1 - (n & 1)], result=wres)
    torch.cuda.synchronize()
    for n, f in enumerate(frames):  # both bands enqueue all their frames; only the flags order them against each other
        u = configs.uniforms_for(p, f.jitter_ndc, reset_history=(n == 0))
        for r in (1, 0):
            a, cnt = bands[r]
            i0, i1 = max(0, a - 2), min(h, a + cnt + 2)
            with torch.cuda.stream(streams[r]):
                ctxs[r].resolve(u, stream=streams[r], color=(f.color[i0:i1], i0), depth=(f.depth[i0:i1], i0), velocity=(f.velocity[i0:i1], i0),
                                history_in=(hist[r][n & 1], hy[r][0]), history_out=(hist[r][1 - (n & 1)], hy[r][0]), result=(res[r], a))
    torch.cuda.synchronize()
    last = 1 - ((nframes - 1) & 1)
    for r in range(2):
        assert ctxs[r].poll_status(streams[r]) == abi.TAA_OK, lib.taa_last_error_string(ctxs[r]._h).decode()
        a, cnt = bands[r]
        assert torch.equal(hist[r][last].view(torch.int16), wh[last][hy[r][0]:hy[r][1]].view(torch.int16)), f"band {r}: history (band + halo rows) differs from the whole frame"
        assert torch.equal(res[r].view(torch.int16), wres[a:a + cnt].view(torch.int16)), f"band {r}: result differs from the whole frame"
    # a call the streaming kernel cannot serve alone is refused on such a context
    with pytest.raises(abi.TaaError):
        m = torch.zeros(bands[0][1], w, dtype=torch.int32, device=dev)
        f = frames[0]
        ctxs[0].resolve(configs.uniforms_for(p, f.jitter_ndc), color=(f.color[0:bands[0][1] + 2], 0), depth=(f.depth[0:bands[0][1] + 2], 0),
                        velocity=(f.velocity[0:bands[0][1] + 2], 0), history_in=(hist[0][0], 0), history_out=(hist[0][1], 0), mask=(m, 0))
    for c in ctxs + [whole]:
        c.close()
