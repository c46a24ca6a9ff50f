"""CPU: the C-ABI library loads and exports every symbol include/taa_b200.h declares; layouts match the reference's blocks."""
import ctypes as C
import os
import re
import subprocess

import pytest

from taa_star_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "taa_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"TAA_API\s+[\w\s\*]+?\b(taa_\w+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound(taalib):
    names = declared_symbols()
    assert len(names) >= 35
    out = subprocess.check_output(["nm", "-D", "--defined-only", abi.LIB_PATH], text=True)
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    missing = [n for n in names if n not in exported]
    assert not missing, f"declared in taa_b200.h but not exported: {missing}"
    unbound = [n for n in names if n not in abi.SIGNATURES]
    assert not unbound, f"declared in taa_b200.h but not bound in abi.py: {unbound}"
    extra = [n for n in abi.SIGNATURES if n not in names]
    assert not extra, f"bound in abi.py but not declared in the header: {extra}"


def test_header_compiles_as_c_and_cpp(tmp_path):
    for comp, std, name in (("/usr/bin/gcc", "-std=c11", "t.c"), ("/usr/bin/g++", "-std=c++17", "t.cpp")):
        p = tmp_path / name
        p.write_text('#include "taa_b200.h"\nint main(void){ TaaUniforms u; (void)u; return sizeof(TaaParameters) == 176 ? 0 : 1; }\n')
        exe = tmp_path / (name + ".out")
        subprocess.check_call([comp, std, "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(p), "-o", str(exe)])
        assert subprocess.call([str(exe)]) == 0


def test_block_layouts_match_reference():
    # offsets from SURVEY.md A.2 (std140 of shaders/taa.comp:50-118, static_asserts at taa.hpp:78,130)
    P, U = abi.TaaParameters, abi.TaaUniforms
    expect_p = dict(mAlpha=0, mColorClampingOrClipping=4, mDepthCulling=8, mUnjitterNeighbourhood=12, mUnjitterCurrentSample=16,
                    mUnjitterFactor=20, mPassThrough=24, mUseYCoCg=28, mShrinkChromaAxis=32, mVarianceClipping=36, mShapedNeighbourhood=40,
                    mLumaWeightingLottes=44, mVarClipGamma=48, mMinAlpha=52, mMaxAlpha=56, mRejectionAlpha=60, mRejectOutside=64,
                    mUseVelocityVectors=68, mVelocitySampleMode=72, mInterpolationMode=76, mToneMapLumaKaris=80, mAddNoise=84,
                    mNoiseFactor=88, mReduceBlendNearClamp=92, mDynamicAntiGhosting=96, mVelBasedAlpha=100, mVelBasedAlphaMax=104,
                    mVelBasedAlphaFactor=108, mRayTraceAugment=112, mRayTraceAugmentFlags=116, mRayTraceAugment_WNrm=120,
                    mRayTraceAugment_WDpt=124, mRayTraceAugment_WMId=128, mRayTraceAugment_WLum=132, mRayTraceAugment_Thresh=136,
                    mRayTraceHistoryCount=140, mDebugMask=144, mDebugMode=160, mDebugScale=164, mDebugCenter=168, mDebugToScreenOutput=172)
    for k, off in expect_p.items():
        assert getattr(P, k).offset == off, k
    assert C.sizeof(P) == 176
    expect_u = dict(mHistoryViewProjMatrix=0, mInverseViewProjMatrix=64, param=128, mJitterNdc=480, mSinTime=496, splitScreen=512, splitX=516,
                    mUpsampling=520, mBypassHistoryUpdate=524, mResetHistory=528, mCamNearPlane=532, mCamFarPlane=536, pad1=540)
    for k, off in expect_u.items():
        assert getattr(U, k).offset == off, k
    assert C.sizeof(U) == 544
    assert C.sizeof(abi.TaaPostProcessPush) == 84 and C.sizeof(abi.TaaCasPush) == 32 and C.sizeof(abi.TaaFxaaPush) == 32


def test_parameter_defaults_match_taa_hpp(taalib):
    p = abi.default_parameters()  # taa.hpp:31-76
    assert p.mAlpha == pytest.approx(0.05) and p.mColorClampingOrClipping == 1 and p.mUnjitterFactor == 1.0
    assert p.mMinAlpha == pytest.approx(1 - 0.97, abs=1e-7) and p.mMaxAlpha == pytest.approx(1 - 0.88, abs=1e-7)
    assert p.mRejectionAlpha == 1.0 and p.mUseVelocityVectors == 1 and p.mVelocitySampleMode == 0 and p.mInterpolationMode == 0
    assert p.mNoiseFactor == pytest.approx(1 / 510) and p.mVelBasedAlphaMax == pytest.approx(0.2) and p.mVelBasedAlphaFactor == pytest.approx(1 / 40)
    assert p.mRayTraceAugmentFlags == (0xffffffff & ~(abi.TAA_RTFLAG_ALL | abi.TAA_RTFLAG_FXD))
    assert (p.mRayTraceAugment_WNrm, p.mRayTraceAugment_WMId, p.mRayTraceAugment_WLum, p.mRayTraceAugment_Thresh) == (0.5, 0.25, 0.5, 0.5)
    assert p.mRayTraceAugment_WDpt == pytest.approx(0.015) and p.mRayTraceHistoryCount == -1
    assert list(p.mDebugMask) == [1, 1, 1, 0] and p.mDebugMode == 0 and p.mDebugScale == 1.0
    for name, _ in abi.TaaParameters._fields_:
        if name.startswith("m") and name not in ("mAlpha", "mColorClampingOrClipping", "mUnjitterFactor", "mVarClipGamma", "mMinAlpha", "mMaxAlpha",
                                                  "mRejectionAlpha", "mUseVelocityVectors", "mNoiseFactor", "mVelBasedAlphaMax", "mVelBasedAlphaFactor",
                                                  "mRayTraceAugmentFlags", "mRayTraceAugment_WNrm", "mRayTraceAugment_WDpt", "mRayTraceAugment_WMId",
                                                  "mRayTraceAugment_WLum", "mRayTraceAugment_Thresh", "mRayTraceHistoryCount", "mDebugMask", "mDebugScale"):
            assert getattr(p, name) == 0, name


def test_no_gpu_means_loud_failure(taalib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    d = abi.taa_desc(C.sizeof(abi.taa_desc), abi.ABI_VERSION, 64, 64, 64, 64, 0, 64, -1, 0)
    h = C.c_void_p()
    st = taalib.taa_create(C.byref(h), C.byref(d))
    assert st == abi.TAA_E_CUDA and not h.value
    assert b"no CPU path" in taalib.taa_last_error_string(None)


def test_create_rejects_bad_descriptors(taalib):
    h = C.c_void_p()
    d = abi.taa_desc(C.sizeof(abi.taa_desc) - 4, abi.ABI_VERSION, 64, 64, 64, 64, 0, 64, -1, 0)
    assert taalib.taa_create(C.byref(h), C.byref(d)) == abi.TAA_E_INVALID_ARG
    d = abi.taa_desc(C.sizeof(abi.taa_desc), abi.ABI_VERSION, 64, 64, 64, 64, 32, 64, -1, 0)  # band leaves the frame
    assert taalib.taa_create(C.byref(h), C.byref(d)) == abi.TAA_E_INVALID_ARG
    d = abi.taa_desc(C.sizeof(abi.taa_desc), abi.ABI_VERSION, 0, 64, 64, 64, 0, 64, -1, 0)
    assert taalib.taa_create(C.byref(h), C.byref(d)) == abi.TAA_E_INVALID_ARG
    assert taalib.taa_create(None, C.byref(d)) == abi.TAA_E_INVALID_ARG
