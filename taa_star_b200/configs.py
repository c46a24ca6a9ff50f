"""The BASELINE.json configs expressed as `Parameters` blocks.

The reference ships no presets; these are this repo's reading of BASELINE.json `configs`.
"""
from __future__ import annotations

from . import abi


def config1_defaults() -> abi.TaaParameters:
    """configs[0]: the reference's defaults (taa.hpp:31-76): alpha 0.05, RGB min/max clamp, bilinear history,
    velocity for movers only (static pixels are reprojected with the matrices)."""
    return abi.default_parameters()


def config2_resolve() -> abi.TaaParameters:
    """configs[1]: YCoCg variance clip (gamma 1), fast clip, Catmull-Rom history, alpha 0.1, velocity for everything."""
    p = abi.default_parameters()
    p.mUseYCoCg = 1
    p.mVarianceClipping = 1
    p.mVarClipGamma = 1.0
    p.mColorClampingOrClipping = 2
    p.mInterpolationMode = 2
    p.mAlpha = 0.1
    p.mUseVelocityVectors = 2
    p.mRejectionAlpha = 1.0
    return p


def config3_full_chain() -> abi.TaaParameters:
    """configs[2]: config 2 + depth/outside/anti-ghost rejection, velocity alpha, Lottes luma weighting.
    (CAS 0.5 and post-process are host settings: mSharpener = 2, mPostProcessEnabled = 1.)"""
    p = config2_resolve()
    p.mDepthCulling = 1
    p.mRejectOutside = 1
    p.mDynamicAntiGhosting = 1
    p.mVelBasedAlpha = 1
    p.mLumaWeightingLottes = 1
    p.mMinAlpha = 0.03
    p.mMaxAlpha = 0.12
    return p


def uniforms_for(params: abi.TaaParameters, jitter_ndc=(0.0, 0.0), reset_history: bool = False, near: float = 0.1, far: float = 100.0,
                 upsampling: bool = False, params1: abi.TaaParameters | None = None, split_x: int | None = None) -> abi.TaaUniforms:
    u = abi.default_uniforms()
    u.param[0] = params
    u.param[1] = params1 if params1 is not None else params
    u.mJitterNdc[0], u.mJitterNdc[1] = jitter_ndc
    u.mResetHistory = 1 if reset_history else 0
    u.mCamNearPlane, u.mCamFarPlane = near, far
    u.mUpsampling = 1 if upsampling else 0
    if split_x is not None:
        u.splitScreen = 1
        u.splitX = split_x
    return u
