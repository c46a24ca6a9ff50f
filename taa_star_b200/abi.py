"""ctypes view of include/taa_b200.h — the C-ABI of libtaa_b200.so.

The structures mirror the reference's std140 blocks byte for byte
(source/taa.hpp:30-130 == shaders/taa.comp:50-118); field names are the reference's.
There is no fallback: if the shared library is missing, importing the compute entry points fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

ABI_VERSION = 2

TAA_OK, TAA_E_INVALID_ARG, TAA_E_UNSUPPORTED, TAA_E_CUDA, TAA_E_NCCL, TAA_E_HALO_OVERFLOW, TAA_E_PEER_TIMEOUT = 0, -1, -2, -3, -4, -5, -6
TAA_BAND_FLAG_WORDS, TAA_IPC_HANDLE_BYTES = 16, 64
TAA_FLAG_DEFAULT, TAA_FLAG_EXACT, TAA_FLAG_FIXUP_ALL = 0, 1, 2

# shaders/shader_cpu_common.h:31-40
TAA_RTFLAG_OUT, TAA_RTFLAG_DIS, TAA_RTFLAG_NRM, TAA_RTFLAG_DPT = 0x1, 0x2, 0x4, 0x8
TAA_RTFLAG_MID, TAA_RTFLAG_LUM, TAA_RTFLAG_CNT, TAA_RTFLAG_ALL = 0x10, 0x20, 0x40, 0x80
TAA_RTFLAG_FXD, TAA_RTFLAG_FXA = 0x100, 0x200

TAA_IMG_RESULT, TAA_IMG_HISTORY, TAA_IMG_DEBUG, TAA_IMG_POSTPROCESS, TAA_IMG_SEGMASK, TAA_IMG_TEMP0, TAA_IMG_TEMP1 = range(7)

_b32, _i32, _u32, _f32 = C.c_uint32, C.c_int32, C.c_uint32, C.c_float


class TaaParameters(C.Structure):
    """struct Parameters — taa.hpp:30-77 / taa.comp:50-95 (176 bytes)."""
    _fields_ = [
        ("mAlpha", _f32), ("mColorClampingOrClipping", _i32), ("mDepthCulling", _b32), ("mUnjitterNeighbourhood", _b32),
        ("mUnjitterCurrentSample", _b32), ("mUnjitterFactor", _f32), ("mPassThrough", _b32), ("mUseYCoCg", _b32),
        ("mShrinkChromaAxis", _b32), ("mVarianceClipping", _b32), ("mShapedNeighbourhood", _b32), ("mLumaWeightingLottes", _b32),
        ("mVarClipGamma", _f32), ("mMinAlpha", _f32), ("mMaxAlpha", _f32), ("mRejectionAlpha", _f32),
        ("mRejectOutside", _b32), ("mUseVelocityVectors", _i32), ("mVelocitySampleMode", _i32), ("mInterpolationMode", _i32),
        ("mToneMapLumaKaris", _b32), ("mAddNoise", _b32), ("mNoiseFactor", _f32), ("mReduceBlendNearClamp", _b32),
        ("mDynamicAntiGhosting", _b32), ("mVelBasedAlpha", _b32), ("mVelBasedAlphaMax", _f32), ("mVelBasedAlphaFactor", _f32),
        ("mRayTraceAugment", _b32), ("mRayTraceAugmentFlags", _u32), ("mRayTraceAugment_WNrm", _f32), ("mRayTraceAugment_WDpt", _f32),
        ("mRayTraceAugment_WMId", _f32), ("mRayTraceAugment_WLum", _f32), ("mRayTraceAugment_Thresh", _f32), ("mRayTraceHistoryCount", _i32),
        ("mDebugMask", _f32 * 4), ("mDebugMode", _i32), ("mDebugScale", _f32), ("mDebugCenter", _b32), ("mDebugToScreenOutput", _b32),
    ]


class TaaUniforms(C.Structure):
    """struct uniforms_for_taa — taa.hpp:113-129 / taa.comp:102-118 (544 bytes)."""
    _fields_ = [
        ("mHistoryViewProjMatrix", _f32 * 16), ("mInverseViewProjMatrix", _f32 * 16), ("param", TaaParameters * 2),
        ("mJitterNdc", _f32 * 4), ("mSinTime", _f32 * 4), ("splitScreen", _b32), ("splitX", _i32), ("mUpsampling", _b32),
        ("mBypassHistoryUpdate", _b32), ("mResetHistory", _b32), ("mCamNearPlane", _f32), ("mCamFarPlane", _f32), ("pad1", _f32),
    ]


class TaaSharpenPush(C.Structure):
    _fields_ = [("sharpeningFactor", _f32)]


class TaaCasPush(C.Structure):
    _fields_ = [("const0", _u32 * 4), ("const1", _u32 * 4)]


class TaaFxaaPush(C.Structure):
    _fields_ = [("fxaaQualityRcpFrame", _f32 * 2), ("fxaaQualitySubpix", _f32), ("fxaaQualityEdgeThreshold", _f32),
                ("fxaaQualityEdgeThresholdMin", _f32), ("pad1", _f32), ("pad2", _f32), ("pad3", _f32)]


class TaaPostProcessPush(C.Structure):
    _fields_ = [("zoomSrcLTWH", _i32 * 4), ("zoomDstLTWH", _i32 * 4), ("debugL_mask", _f32 * 4), ("debugR_mask", _f32 * 4),
                ("zoom", _b32), ("showZoomBox", _b32), ("splitX", _i32), ("debugL_show", _b32), ("debugR_show", _b32)]


class taa_image(C.Structure):
    _fields_ = [("data", C.c_void_p), ("pitch_bytes", C.c_int64), ("y0", _i32), ("rows", _i32)]


RESOLVE_IMAGE_NAMES = ("color", "depth", "velocity", "history_in", "history_depth", "history_out", "result", "debug",
                       "segmask", "prev_segmask", "matid", "prev_matid", "uvnrm", "mask")


class taa_resolve_images(C.Structure):
    _fields_ = [(n, taa_image) for n in RESOLVE_IMAGE_NAMES]


class taa_post_chain(C.Structure):
    _fields_ = [("sharpener", _i32), ("sharpen", TaaSharpenPush), ("cas", TaaCasPush), ("postprocess", _i32), ("pp", TaaPostProcessPush),
                ("fxaa", _i32), ("fxaa_pc", TaaFxaaPush)]


class taa_desc(C.Structure):
    _fields_ = [("struct_size", _u32), ("abi_version", _u32), ("in_width", _i32), ("in_height", _i32), ("out_width", _i32),
                ("out_height", _i32), ("band_y0", _i32), ("band_rows", _i32), ("device", _i32), ("flags", _u32)]


class taa_band_peer(C.Structure):
    """A neighbour band as mapped into this process (taa_band_peers)."""
    _fields_ = [("history", C.c_void_p * 2), ("row_pitch", C.c_int64), ("y0", _i32), ("band_rows", _i32), ("flags", C.c_void_p)]


class taa_jitter_settings(C.Structure):
    _fields_ = [("mSampleDistribution", _i32), ("mFixedJitterIndex", _i32), ("mJitterExtraScale", _f32), ("mJitterSlowMotion", _i32),
                ("mJitterRotateDegrees", _f32), ("mDebugSampleOffsets", C.POINTER(_f32)), ("mDebugSampleOffsetsCount", _i32)]


class taa_source_views(C.Structure):
    _fields_ = [("color", C.c_void_p), ("depth", C.c_void_p), ("uvnrm", C.c_void_p), ("velocity", C.c_void_p),
                ("matid", C.c_void_p), ("raytraced", C.c_void_p)]


class taa_invokee_settings(C.Structure):
    _fields_ = [("mTaaEnabled", _b32), ("mPostProcessEnabled", _b32), ("mResetHistory", _b32), ("mSplitScreen", _b32), ("mSplitX", _i32),
                ("mSharpener", _i32), ("mSharpenFactor", _f32), ("mResetHistoryOnChange", _b32), ("jitter", taa_jitter_settings)]


assert C.sizeof(TaaParameters) == 176 and C.sizeof(TaaUniforms) == 544
assert C.sizeof(TaaCasPush) == 32 and C.sizeof(TaaFxaaPush) == 32 and C.sizeof(TaaPostProcessPush) == 84

_P = C.POINTER
_vp, _ll = C.c_void_p, C.c_longlong

# name -> (restype, argtypes); every TAA_API symbol of include/taa_b200.h
SIGNATURES = {
    "taa_abi_version": (C.c_int, []),
    "taa_create": (C.c_int, [_P(_vp), _P(taa_desc)]),
    "taa_destroy": (None, [_vp]),
    "taa_last_error_string": (C.c_char_p, [_vp]),
    "taa_status_string": (C.c_char_p, [C.c_int]),
    "taa_resolve": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _P(TaaUniforms), _vp]),
    "taa_resolve_ex": (C.c_int, [_vp, _P(taa_resolve_images), _P(TaaUniforms), _vp]),
    "taa_frame": (C.c_int, [_vp, _P(taa_resolve_images), _P(TaaUniforms), _P(taa_post_chain), _P(taa_image), _vp]),
    "taa_sharpen": (C.c_int, [_vp, _P(taa_image), _P(taa_image), _P(TaaSharpenPush), _vp]),
    "taa_sharpen_cas": (C.c_int, [_vp, _P(taa_image), _P(taa_image), _P(TaaCasPush), _vp]),
    "taa_post_process": (C.c_int, [_vp, _P(taa_image), _P(taa_image), _P(taa_image), _P(TaaPostProcessPush), _vp]),
    "taa_fxaa_prepare": (C.c_int, [_vp, _P(taa_image), _P(taa_image), _vp]),
    "taa_fxaa": (C.c_int, [_vp, _P(taa_image), _P(taa_image), _P(taa_image), _P(TaaFxaaPush), _vp]),
    "taa_fxaa_fused": (C.c_int, [_vp, _P(taa_image), _P(taa_image), _P(taa_image), _P(TaaFxaaPush), _vp]),
    "taa_fxaa_default": (None, [_P(TaaFxaaPush), _i32, _i32]),
    "taa_launch_count": (_ll, [_vp]),
    "taa_poll_status": (C.c_int, [_vp, _vp]),
    "taa_fixup_pixels": (_ll, [_vp, _vp]),
    "taa_cas_setup": (None, [_P(TaaCasPush), _f32, _f32, _f32]),
    "taa_parameters_default": (None, [_P(TaaParameters)]),
    "taa_uniforms_default": (None, [_P(TaaUniforms)]),
    "taa_postprocess_default": (None, [_P(TaaPostProcessPush), _i32, _i32]),
    "taa_jitter_offset_for_frame": (C.c_int, [_P(taa_jitter_settings), _i32, _i32, C.c_int64, _P(_f32)]),
    "taa_halton": (_f32, [_i32, _i32]),
    "taa_jittered_projection": (None, [_P(_f32), _f32, _f32, _P(_f32)]),
    "taa_reprojection_matrices": (C.c_int, [_P(_f32)] * 6),
    "taa_invokee_create": (C.c_int, [_P(_vp), _i32, _i32, _u32]),
    "taa_invokee_destroy": (None, [_vp]),
    "taa_invokee_last_error": (C.c_char_p, [_vp]),
    "taa_invokee_set_source_image_views": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _P(taa_source_views)]),
    "taa_invokee_parameters": (_P(TaaParameters), [_vp, _i32]),
    "taa_invokee_settings_ptr": (_P(taa_invokee_settings), [_vp]),
    "taa_invokee_postprocess": (_P(TaaPostProcessPush), [_vp]),
    "taa_invokee_get_jittered_projection_matrix": (C.c_int, [_vp, _P(_f32), C.c_int64, _P(_f32), _P(_f32)]),
    "taa_invokee_save_history_proj_matrix": (C.c_int, [_vp, _P(_f32), C.c_int64]),
    "taa_invokee_update": (C.c_int, [_vp, C.c_int64, _P(_f32), _f32, _f32, _f32]),
    "taa_invokee_render": (C.c_int, [_vp, C.c_int64, _vp, _P(_vp)]),
    "taa_invokee_duration": (_f32, [_vp]),
    "taa_invokee_image": (_vp, [_vp, _i32, _i32]),
    "taa_invokee_uniforms": (_P(TaaUniforms), [_vp]),
    "taa_invokee_frame_host": (C.c_int, [_vp, C.c_int64, _P(taa_source_views), _P(_f32), _P(_f32), _f32, _f32, _f32, _vp]),
    "taa_invokee_wait": (C.c_int, [_vp, C.c_int64]),
    "taa_invokee_launch_count": (_ll, [_vp]),
    "taa_invokee_h2d_bytes": (_ll, [_vp]),
    "taa_settings_write_ini": (C.c_int32, [_P(TaaParameters), _P(taa_invokee_settings), _P(TaaPostProcessPush), C.c_char_p, C.c_int32]),
    "taa_settings_read_ini": (C.c_int, [C.c_char_p, _P(TaaParameters), _P(taa_invokee_settings), _P(TaaPostProcessPush), _P(_f32), C.c_int32]),
    "taa_settings_ini_last_error": (C.c_char_p, []),
    "taa_invokee_write_settings_ini": (C.c_int32, [_vp, C.c_char_p, C.c_int32]),
    "taa_invokee_read_settings_ini": (C.c_int, [_vp, C.c_char_p]),
    "taa_band_peers": (C.c_int, [_vp, _P(taa_band_peer), _P(taa_band_peer), _vp, _vp, _vp, _i32]),
    "taa_device_alloc": (_vp, [C.c_size_t]),
    "taa_device_free": (None, [_vp]),
    "taa_ipc_export": (C.c_int, [_vp, _vp]),
    "taa_ipc_open": (C.c_int, [_vp, _P(_vp)]),
    "taa_ipc_close": (C.c_int, [_vp]),
    "taa_host_alloc": (_vp, [C.c_size_t]),
    "taa_host_free": (None, [_vp]),
}

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libtaa_b200.so")
_lib = None


class TaaError(RuntimeError):
    def __init__(self, status: int, what: str):
        super().__init__(what)
        self.status = status


def load_library() -> C.CDLL:
    """Loads libtaa_b200.so and binds every symbol. Raises if the library or a symbol is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with ./build.sh (or __graft_entry__.build()). "
                          "There is no CPU or PyTorch fallback for the resolve path.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    if lib.taa_abi_version() != ABI_VERSION:
        raise ImportError(f"libtaa_b200.so has ABI {lib.taa_abi_version()}, bindings expect {ABI_VERSION}")
    _lib = lib
    return lib


def default_parameters() -> TaaParameters:
    p = TaaParameters()
    load_library().taa_parameters_default(C.byref(p))
    return p


def default_uniforms() -> TaaUniforms:
    u = TaaUniforms()
    load_library().taa_uniforms_default(C.byref(u))
    return u
