"""CPU: the pure-host helpers of the library (jitter, CasSetup, matrices) against the oracle and the known answers."""
import ctypes as C
import json
import math
import os

import numpy as np
import pytest

from taa_star_b200 import abi, host

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def known():
    return json.load(open(os.path.join(GOLDEN, "known_answers.json")))


def test_cas_setup_known_answer(taalib, oracle):
    k = known()["cas_setup"]
    for case in k:
        pc = host.cas_setup(case["sharpness"], case["w"], case["h"])
        assert list(pc.const0) == case["const0"], case
        assert list(pc.const1) == case["const1"], case
        c0, c1 = oracle.cas_setup(case["sharpness"], case["w"], case["h"])
        assert c0 == case["const0"] and c1 == case["const1"], case


def test_halton_known_answer(taalib, oracle):
    k = known()["halton_2_3_8_px"]
    for i, (hx, hy) in enumerate(k):
        assert taalib.taa_halton(i + 1, 2) - 0.5 == pytest.approx(hx, abs=1e-6)
        assert taalib.taa_halton(i + 1, 3) - 0.5 == pytest.approx(hy, abs=1e-6)
        assert taalib.taa_halton(i + 1, 2) == oracle.halton(i + 1, 2)
        assert taalib.taa_halton(i + 1, 3) == oracle.halton(i + 1, 3)
    for b in (2, 3, 5):
        for i in range(1, 200):
            assert taalib.taa_halton(i, b) == oracle.halton(i, b)


@pytest.mark.parametrize("dist", [0, 1, 2, 3, 4, 5])
def test_jitter_patterns_match_oracle(oracle, dist):
    dbg = [(0.1, -0.2), (0.3, 0.4), (-0.45, 0.05)] if dist == 5 else None
    for (w, h) in ((1920, 1080), (3840, 2160), (333, 77)):
        for frame in list(range(0, 40)) + [12345]:
            for kw in (dict(), dict(fixed_index=2), dict(slow_motion=3), dict(rotate_degrees=30.0, extra_scale=1.5)):
                a, n = host.jitter_offset_for_frame(frame, w, h, sample_distribution=dist, debug_offsets=dbg, **kw)
                b, m = oracle.jitter(frame, w, h, sample_distribution=dist, debug_offsets=dbg, **kw)
                assert n == m == {0: 4, 1: 4, 2: 8, 3: 16, 4: 16, 5: 3}[dist]
                assert a == b, (dist, w, h, frame, kw)


def test_jitter_is_within_half_a_pixel_and_halton_ndc_scale():
    for frame in range(16):
        (x, y), n = host.jitter_offset_for_frame(frame, 1920, 1080, sample_distribution=2)
        assert abs(x) <= 1.0 / 1920 and abs(y) <= 1.0 / 1080
    (x, y), _ = host.jitter_offset_for_frame(1, 1920, 1080, sample_distribution=2)
    assert x == pytest.approx(-0.25 * 2 / 1920, rel=1e-6) and y == pytest.approx((1 / 3 + 1 / 9 - 0.5 + 0.0) * 0 + (2 / 3 - 0.5) * 2 / 1080, rel=1e-5)


def test_jittered_projection_is_translate_times_proj(taalib):
    rng = np.random.default_rng(1)
    P = rng.standard_normal((4, 4)).astype(np.float32)  # P[row, col]
    flat = P.T.reshape(-1)  # column-major
    out = (C.c_float * 16)()
    taalib.taa_jittered_projection((C.c_float * 16)(*flat), 0.25, -0.5, out)
    T = np.eye(4, dtype=np.float32)
    T[0, 3], T[1, 3] = 0.25, -0.5
    want = (T @ P).T.reshape(-1)  # glm::translate(...) * P  (taa.hpp:248)
    assert np.allclose(np.array(out[:]), want, rtol=1e-6, atol=1e-6)


def test_reprojection_matrices(taalib):
    from taa_star_b200.synth import SyntheticScene
    sc = SyntheticScene(64, 36, with_aux=False)
    P, V1, V0 = sc.proj_matrix(), sc.view_matrix(5), sc.view_matrix(4)
    inv, hist = (C.c_float * 16)(), (C.c_float * 16)()
    m = lambda a: (C.c_float * 16)(*a)
    assert taalib.taa_reprojection_matrices(m(P), m(V1), m(P), m(V0), inv, hist) == 0
    Pm, V1m, V0m = (np.array(a, dtype=np.float64).reshape(4, 4).T for a in (P, V1, V0))
    assert np.allclose(np.array(inv[:]).reshape(4, 4).T, np.linalg.inv(Pm @ V1m), rtol=1e-5, atol=1e-6)
    assert np.allclose(np.array(hist[:]).reshape(4, 4).T, Pm @ V0m, rtol=1e-6, atol=1e-6)
    # a point on the plane reprojects by the pan: history uv = uv - pan/res
    uv = np.array([0.3, 0.6])
    clip = np.array([uv[0] * 2 - 1, uv[1] * 2 - 1, sc.ndc_depth(sc.plane_depth), 1.0])
    world = np.array(inv[:]).reshape(4, 4).T @ clip
    h = np.array(hist[:]).reshape(4, 4).T @ world
    huv = h[:2] / h[3] * 0.5 + 0.5
    assert np.allclose(uv - huv, [sc.pan[0] / sc.W, sc.pan[1] / sc.H], atol=1e-6)
    sing = [0.0] * 16
    assert taalib.taa_reprojection_matrices(m(sing), m(V1), m(P), m(V0), inv, hist) == abi.TAA_E_INVALID_ARG


def test_postprocess_defaults(taalib):
    pp = host.postprocess_default(1920, 1080)  # taa.hpp:101-111 evaluated by :352-359
    assert list(pp.zoomSrcLTWH) == [(1920 - 20) // 2, (1080 - 20) // 2, 20, 20]
    assert list(pp.zoomDstLTWH) == [1920 - 200 - 10, 10, 200, 200]
    assert pp.splitX == -1 and pp.zoom == 0 and pp.showZoomBox == 1 and list(pp.debugL_mask) == [1, 1, 1, 0]


def test_invokee_settings_defaults(taalib):
    t = host.Taa(3)
    s = t.settings
    assert s.mTaaEnabled == 1 and s.mPostProcessEnabled == 1 and s.mSharpener == 0 and s.mSharpenFactor == 0.5  # taa.hpp:1351-1352,1418-1419
    assert s.jitter.mSampleDistribution == 1 and s.jitter.mFixedJitterIndex == -1 and s.jitter.mJitterSlowMotion == 1  # taa.hpp:1353,1404-1407
    assert s.mResetHistoryOnChange == 1
    assert t.mParameters[0].mAlpha == pytest.approx(0.05) and t.execution_order() == 100 and t.taa_enabled()
    t.mParameters[1].mAlpha = 0.25
    assert t.mParameters[1].mAlpha == 0.25
    with pytest.raises(Exception):
        host.Taa(1)  # static_assert(CF > 1), taa.hpp:1005
    t.close()


# ---- settings files: writeSettingsToIni / readSettingsFromIni, taa.hpp:1198-1339 ----------------------------------
def _settings_from(values):
    """The parameter blocks of tests/golden/taa_settings_values.json (the state make_ini_golden.py gave the reference's class)."""
    params = [abi.TaaParameters(), abi.TaaParameters()]
    for p, v in zip(params, values["param"]):
        for k, x in v.items():
            if k == "mDebugMask":
                for i in range(4):
                    p.mDebugMask[i] = x[i]
            else:
                setattr(p, k, x)
    s, pp = abi.taa_invokee_settings(), abi.TaaPostProcessPush()
    pr = values["primary"]
    for k in ("mTaaEnabled", "mSplitScreen", "mResetHistoryOnChange", "mPostProcessEnabled", "mSharpener", "mSplitX", "mSharpenFactor"):
        setattr(s, k, pr[k])
    for k in ("mSampleDistribution", "mFixedJitterIndex", "mJitterSlowMotion", "mJitterExtraScale", "mJitterRotateDegrees"):
        setattr(s.jitter, k, pr[k])
    offs = (C.c_float * (2 * len(pr["mDebugSampleOffsets"])))(*[c for o in pr["mDebugSampleOffsets"] for c in o])
    s.jitter.mDebugSampleOffsets = C.cast(offs, C.POINTER(C.c_float))
    s.jitter.mDebugSampleOffsetsCount = len(pr["mDebugSampleOffsets"])
    s._keep = offs
    po = values["post"]
    pp.zoom, pp.showZoomBox = po["zoom"], po["showZoomBox"]
    for i in range(4):
        pp.zoomSrcLTWH[i], pp.zoomDstLTWH[i] = po["zoomSrcLTWH"][i], po["zoomDstLTWH"][i]
    return params, s, pp


def _f32(x):
    return float(np.float32(x))


def test_write_settings_ini_matches_the_reference_text(taalib):
    """Byte for byte what the reference's writeSettingsToIni + mINI generate() wrote for the same values (make_ini_golden.py)."""
    values = json.load(open(os.path.join(GOLDEN, "taa_settings_values.json")))
    params, s, pp = _settings_from(values)
    text = host.write_settings_ini(params, s, pp)
    want = open(os.path.join(GOLDEN, "taa_settings_written.ini")).read()
    assert text == want


def test_read_settings_ini_matches_the_reference(taalib):
    """Every field as the reference's readSettingsFromIni left it after reading tests/golden/taa_settings_input.ini (mixed case, comments,
    empty and missing keys, duplicates, 'true'/'yes', trailing text after numbers, fewer sample offsets than before)."""
    values = json.load(open(os.path.join(GOLDEN, "taa_settings_values.json")))
    want = json.load(open(os.path.join(GOLDEN, "taa_settings_read.json")))
    params, s, pp = _settings_from(values)
    offs = host.read_settings_ini(open(os.path.join(GOLDEN, "taa_settings_input.ini")).read(), params, s, pp)
    for p, w in zip(params, want["param"]):
        for k, x in w.items():
            if k == "mDebugMask":
                assert [p.mDebugMask[i] for i in range(4)] == [_f32(c) for c in x], k
            else:
                got = getattr(p, k)
                assert got == (_f32(x) if isinstance(got, float) else x), (k, got, x)
    pr = want["primary"]
    for k in ("mTaaEnabled", "mSplitScreen", "mResetHistoryOnChange", "mPostProcessEnabled", "mSharpener", "mSplitX"):
        assert getattr(s, k) == pr[k], k
    assert s.mSharpenFactor == _f32(pr["mSharpenFactor"])
    for k in ("mSampleDistribution", "mFixedJitterIndex", "mJitterSlowMotion"):
        assert getattr(s.jitter, k) == pr[k], k
    for k in ("mJitterExtraScale", "mJitterRotateDegrees"):
        assert getattr(s.jitter, k) == _f32(pr[k]), k
    assert [(_f32(a), _f32(b)) for a, b in pr["mDebugSampleOffsets"]] == offs
    po = want["post"]
    assert (pp.zoom, pp.showZoomBox) == (po["zoom"], po["showZoomBox"])
    assert [pp.zoomSrcLTWH[i] for i in range(4)] == po["zoomSrcLTWH"] and [pp.zoomDstLTWH[i] for i in range(4)] == po["zoomDstLTWH"]


def test_settings_ini_round_trip_and_errors(taalib):
    values = json.load(open(os.path.join(GOLDEN, "taa_settings_values.json")))
    params, s, pp = _settings_from(values)
    text = host.write_settings_ini(params, s, pp)
    p2 = [abi.TaaParameters(), abi.TaaParameters()]
    s2, pp2 = abi.taa_invokee_settings(), abi.TaaPostProcessPush()
    host.read_settings_ini(text, p2, s2, pp2)
    # std::to_string(float) keeps six decimals: what survives the trip is the written text
    assert host.write_settings_ini(p2, s2, pp2) == host.write_settings_ini(*_reparse(text))
    assert p2[1].mInterpolationMode == params[1].mInterpolationMode and p2[0].mUseYCoCg == params[0].mUseYCoCg
    assert s2.jitter.mDebugSampleOffsetsCount == 3 and pp2.zoomDstLTWH[0] == -3
    with pytest.raises(host.TaaError) as e:
        host.read_settings_ini("[TAA_Param_0]\nmAlpha=fast\nmMinAlpha=0.25\n", p2, s2, pp2)
    assert "malpha" in str(e.value).lower()
    assert p2[0].mMinAlpha == 0.25  # the valid keys are applied all the same
    # an empty text changes nothing
    before = bytes(p2[0])
    host.read_settings_ini("", p2, s2, pp2)
    assert bytes(p2[0]) == before


def _reparse(text):
    p = [abi.TaaParameters(), abi.TaaParameters()]
    s, pp = abi.taa_invokee_settings(), abi.TaaPostProcessPush()
    host.read_settings_ini(text, p, s, pp)
    return p, s, pp


# ---- jitter patterns and the jittered projection against the reference's own code ------------------------------------------------
def _bits(x):
    return int(np.float32(x).view(np.uint32))


def _from_bits(u):
    return float(np.uint32(u).view(np.float32))


def test_jitter_and_projection_match_the_reference_bit_for_bit(taalib, oracle):
    """tests/golden/jitter_golden.json: taa<CF>::get_jitter_offset_for_frame for all six distributions x {slow motion, fixed index,
    rotation, extra scale} x 20 frames x four resolutions, and get_jittered_projection_matrix, produced by the reference's own function
    bodies compiled as-is (make_jitter_golden.py). The library and the oracle must reproduce every float bit."""
    g = json.load(open(os.path.join(GOLDEN, "jitter_golden.json")))
    dbg = [tuple(o) for o in g["debug_offsets"]]
    checked = 0
    for res in g["resolutions"]:
        w, h = res["w"], res["h"]
        for c in res["cases"]:
            scale, rot = _from_bits(c["scale"]), _from_bits(c["rot"])
            for f in range(20):
                want = (c["xy"][2 * f], c["xy"][2 * f + 1])
                (x, y), n = host.jitter_offset_for_frame(f, w, h, c["dist"], c["fixed"], scale, c["slow"], rot, dbg)
                assert n == c["n"], (w, h, c, n)
                assert (_bits(x), _bits(y)) == want, f"library: {w}x{h} dist {c['dist']} fixed {c['fixed']} slow {c['slow']} scale {scale} rot {rot} frame {f}"
                (ox, oy), on = oracle.jitter(f, w, h, c["dist"], c["fixed"], scale, c["slow"], rot, dbg)
                assert on == c["n"]
                assert (_bits(ox), _bits(oy)) == want, f"oracle: {w}x{h} dist {c['dist']} frame {f}"
                checked += 1
        for p in res["proj"]:
            out = (C.c_float * 16)()
            taalib.taa_jittered_projection((C.c_float * 16)(*[_from_bits(u) for u in p["in"]]), _from_bits(p["off"][0]), _from_bits(p["off"][1]), out)
            assert [_bits(v) for v in out] == p["out"], f"projection {w}x{h} frame {p['frame']}"
            (jx, jy), _ = host.jitter_offset_for_frame(p["frame"], w, h, 2)
            assert (_bits(jx), _bits(jy)) == tuple(p["off"])
    assert checked == 4 * 24 * 20


def test_reprojection_matrices_match_glm_bit_for_bit(taalib):
    """taa.hpp:993-994 — glm::inverse(P_cur * V_cur) and P_prev * V_prev — computed with the GLM the reference vendors for twelve camera
    poses (tests/golden/make_matrix_golden.py); taa_reprojection_matrices must reproduce every float bit."""
    g = json.load(open(os.path.join(GOLDEN, "matrix_golden.json")))
    for i, c in enumerate(g):
        m = {k: (C.c_float * 16)(*[_from_bits(u) for u in c[k]]) for k in ("proj_cur", "view_cur", "proj_prev", "view_prev")}
        inv, hist = (C.c_float * 16)(), (C.c_float * 16)()
        assert taalib.taa_reprojection_matrices(m["proj_cur"], m["view_cur"], m["proj_prev"], m["view_prev"], inv, hist) == 0
        assert [_bits(v) for v in hist] == c["history_view_proj"], f"pose {i}: P_prev * V_prev"
        assert [_bits(v) for v in inv] == c["inverse_view_proj"], f"pose {i}: inverse(P_cur * V_cur)"


def test_pinned_sincos_is_close_to_libm(oracle):
    """GLSL leaves sin / cos to the implementation; oracle, shim and kernels share one software version (taa_sincos). It has to be a sane
    sin / cos: within 2e-7 of libm on the range the spherical normals use, finite-argument results in [-1, 1], NaN for non-finite arguments."""
    import ctypes as C
    import numpy as np
    L = oracle.lib()
    L.taa_oracle_sincos.restype = None
    L.taa_oracle_sincos.argtypes = [C.POINTER(C.c_float)] * 3 + [C.c_int]
    x = np.concatenate([np.linspace(-7.0, 7.0, 200001), np.linspace(-1000.0, 1000.0, 20001), [0.0, np.pi / 2, np.pi, 1.5 * np.pi, 2 * np.pi]]).astype(np.float32)
    s, c = np.empty_like(x), np.empty_like(x)
    ptr = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))
    L.taa_oracle_sincos(ptr(x), ptr(s), ptr(c), x.size)
    small = np.abs(x) <= 7.0
    assert np.abs(s[small] - np.sin(x[small].astype(np.float64))).max() < 2e-7
    assert np.abs(c[small] - np.cos(x[small].astype(np.float64))).max() < 2e-7
    assert np.abs(s - np.sin(x.astype(np.float64))).max() < 1e-4 and np.abs(c - np.cos(x.astype(np.float64))).max() < 1e-4
    assert np.abs(s).max() <= 1.0 + 1e-6 and np.abs(c).max() <= 1.0 + 1e-6
    bad = np.array([np.inf, -np.inf, np.nan, 3e9], np.float32)
    sb, cb = np.empty_like(bad), np.empty_like(bad)
    L.taa_oracle_sincos(ptr(bad), ptr(sb), ptr(cb), bad.size)
    assert np.isnan(sb[:3]).all() and np.isnan(cb[:3]).all() and sb[3] == 0.0 and cb[3] == 1.0
