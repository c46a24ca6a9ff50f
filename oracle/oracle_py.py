"""ctypes loader of oracle/libtaa_oracle.so — TEST INFRASTRUCTURE ONLY.

May be imported by tests/, __graft_entry__.smoke() and the cpu_baseline / --impl reference legs of bench.py.
The product package (taa_star_b200/) must never import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
from taa_star_b200 import abi  # struct layouts only (POD mirrors of the reference's blocks)  # noqa: E402

LIB_PATH = os.path.join(_HERE, "libtaa_oracle.so")
_lib = None


def build():
    subprocess.check_call(["make", "-C", _HERE, "libtaa_oracle.so"], stdout=subprocess.DEVNULL)


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        P = C.POINTER
        L.taa_oracle_resolve.restype = C.c_int
        L.taa_oracle_resolve.argtypes = [P(abi.taa_resolve_images), P(abi.TaaUniforms)] + [C.c_int] * 7
        L.taa_oracle_sharpen.restype = C.c_int
        L.taa_oracle_sharpen.argtypes = [P(abi.taa_image), P(abi.taa_image), C.c_int, C.c_int, C.c_float, C.c_int]
        L.taa_oracle_cas.restype = C.c_int
        L.taa_oracle_cas.argtypes = [P(abi.taa_image), P(abi.taa_image), C.c_int, C.c_int, P(C.c_uint32), P(C.c_uint32), C.c_int]
        L.taa_oracle_cas_setup.restype = None
        L.taa_oracle_cas_setup.argtypes = [P(C.c_uint32), P(C.c_uint32)] + [C.c_float] * 5
        L.taa_oracle_post_process.restype = C.c_int
        L.taa_oracle_post_process.argtypes = [P(abi.taa_image), P(abi.taa_image), P(abi.taa_image), C.c_int, C.c_int, P(abi.TaaPostProcessPush), C.c_int]
        L.taa_oracle_fxaa_prepare.restype = C.c_int
        L.taa_oracle_fxaa_prepare.argtypes = [P(abi.taa_image), P(abi.taa_image), C.c_int, C.c_int, C.c_int]
        L.taa_oracle_fxaa.restype = C.c_int
        L.taa_oracle_fxaa.argtypes = [P(abi.taa_image), P(abi.taa_image), P(abi.taa_image), C.c_int, C.c_int, P(abi.TaaFxaaPush), C.c_int, C.c_int]
        L.taa_oracle_halton.restype = C.c_float
        L.taa_oracle_halton.argtypes = [C.c_int, C.c_int]
        L.taa_oracle_jitter.restype = C.c_int
        L.taa_oracle_jitter.argtypes = [C.c_int, C.c_int, C.c_float, C.c_int, C.c_float, P(C.c_float), C.c_int, C.c_int, C.c_int, C.c_longlong, P(C.c_float)]
        L.taa_oracle_max_threads.restype = C.c_int
        L.taa_oracle_f32_to_f16.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong]
        L.taa_oracle_f16_to_f32.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong]
        _lib = L
    return _lib


REF_LIB_PATH = os.path.join(_HERE, "_ref", "libtaa_ref.so")
_ref = None


def ref_available() -> bool:
    return os.path.exists(REF_LIB_PATH)


def ref_lib() -> C.CDLL:
    """oracle/_ref/libtaa_ref.so: the reference's own shader sources compiled through oracle/glsl_shim.h (oracle/ref_build.py)."""
    global _ref
    if _ref is None:
        L = C.CDLL(REF_LIB_PATH)
        P = C.POINTER
        L.taa_ref_resolve.restype = C.c_int
        L.taa_ref_resolve.argtypes = [P(abi.taa_resolve_images), P(abi.TaaUniforms)] + [C.c_int] * 7
        L.taa_ref_sharpen.restype = C.c_int
        L.taa_ref_sharpen.argtypes = [P(abi.taa_image), P(abi.taa_image), C.c_int, C.c_int, C.c_float]
        L.taa_ref_sharpen_cas.restype = C.c_int
        L.taa_ref_sharpen_cas.argtypes = [P(abi.taa_image), P(abi.taa_image), C.c_int, C.c_int, P(abi.TaaCasPush)]
        L.taa_ref_post_process.restype = C.c_int
        L.taa_ref_post_process.argtypes = [P(abi.taa_image), P(abi.taa_image), P(abi.taa_image), C.c_int, C.c_int, P(abi.TaaPostProcessPush)]
        L.taa_ref_fxaa_prepare.restype = C.c_int
        L.taa_ref_fxaa_prepare.argtypes = [P(abi.taa_image), P(abi.taa_image), C.c_int, C.c_int]
        for f in (L.taa_ref_fxaa_gather4, L.taa_ref_fxaa_offset):
            f.restype = C.c_int
            f.argtypes = [P(abi.taa_image), P(abi.taa_image), P(abi.taa_image), C.c_int, C.c_int, P(abi.TaaFxaaPush)]
        _ref = L
    return _ref


def _img(a):
    if a is None:
        return abi.taa_image(None, 0, 0, 0)
    assert a.flags["C_CONTIGUOUS"]
    return abi.taa_image(a.ctypes.data, a.strides[0], 0, a.shape[0])


# numpy dtypes/shapes of the bindings: rgba16f -> (H, W, 4) uint16|float16, D32 -> (H, W) float32, r32ui -> (H, W) uint32|int32
def resolve(uniforms: abi.TaaUniforms, color, depth, velocity, history_in, history_depth=None, prev_segmask=None, matid=None, prev_matid=None,
            uvnrm=None, out_size=None, want=("history_out", "result"), rows=None, nthreads=0, impl="oracle"):
    """taa.comp on the CPU. Returns a dict of freshly allocated outputs named in `want`
    (any of history_out, result, debug, segmask, mask)."""
    in_h, in_w = depth.shape
    out_w, out_h = out_size if out_size else (in_w, in_h)
    outs = {}
    for name in want:
        if name in ("history_out", "result", "debug"):
            outs[name] = np.zeros((out_h, out_w, 4), dtype=np.float16)
        elif name in ("segmask", "mask"):
            outs[name] = np.zeros((out_h, out_w), dtype=np.uint32)
        else:
            raise KeyError(name)
    im = abi.taa_resolve_images()
    im.color, im.depth, im.velocity, im.history_in = _img(color), _img(depth), _img(velocity), _img(history_in)
    im.history_depth, im.prev_segmask, im.matid, im.prev_matid, im.uvnrm = _img(history_depth), _img(prev_segmask), _img(matid), _img(prev_matid), _img(uvnrm)
    for name, a in outs.items():
        setattr(im, name, _img(a))
    y0, y1 = rows if rows else (0, out_h)
    if impl == "ref":  # the reference's shader text itself; it has no `mask` output (that image is this repo's addition)
        assert "mask" not in want
        st = ref_lib().taa_ref_resolve(C.byref(im), C.byref(uniforms), in_w, in_h, out_w, out_h, y0, y1, nthreads)
    else:
        st = lib().taa_oracle_resolve(C.byref(im), C.byref(uniforms), in_w, in_h, out_w, out_h, y0, y1, nthreads)
    if st != 0:
        raise RuntimeError(f"taa resolve ({impl}) failed: {st}")
    return outs


def ref_sharpen(src, factor):
    h, w = src.shape[:2]
    dst = np.zeros_like(src)
    a, b = _img(src), _img(dst)
    assert ref_lib().taa_ref_sharpen(C.byref(a), C.byref(b), w, h, factor) == 0
    return dst


def ref_cas(src, const0, const1):
    """sharpen_cas.comp (+ ffx_cas.h's CasFilter) of the reference itself, compiled through the GLSL shim."""
    h, w = src.shape[:2]
    dst = np.zeros_like(src)
    a, b = _img(src), _img(dst)
    pc = abi.TaaCasPush()
    for i in range(4):
        pc.const0[i], pc.const1[i] = const0[i], const1[i]
    assert ref_lib().taa_ref_sharpen_cas(C.byref(a), C.byref(b), w, h, C.byref(pc)) == 0
    return dst


def ref_post_process(src, debug, pc: abi.TaaPostProcessPush):
    h, w = src.shape[:2]
    dst = np.zeros_like(src)
    a, b, d = _img(src), _img(dst), _img(debug)
    assert ref_lib().taa_ref_post_process(C.byref(a), C.byref(d) if debug is not None else None, C.byref(b), w, h, C.byref(pc)) == 0
    return dst


def sharpen(src, factor, nthreads=0):
    h, w = src.shape[:2]
    dst = np.zeros_like(src)
    a, b = _img(src), _img(dst)
    assert lib().taa_oracle_sharpen(C.byref(a), C.byref(b), w, h, factor, nthreads) == 0
    return dst


def cas_setup(sharpness, w, h):
    c0, c1 = (C.c_uint32 * 4)(), (C.c_uint32 * 4)()
    lib().taa_oracle_cas_setup(c0, c1, sharpness, float(w), float(h), float(w), float(h))
    return list(c0), list(c1)


def cas(src, const0, const1, nthreads=0):
    h, w = src.shape[:2]
    dst = np.zeros_like(src)
    a, b = _img(src), _img(dst)
    c0, c1 = (C.c_uint32 * 4)(*const0), (C.c_uint32 * 4)(*const1)
    assert lib().taa_oracle_cas(C.byref(a), C.byref(b), w, h, c0, c1, nthreads) == 0
    return dst


def post_process(src, debug, pc: abi.TaaPostProcessPush, nthreads=0):
    h, w = src.shape[:2]
    dst = np.zeros_like(src)
    a, b, d = _img(src), _img(dst), _img(debug)
    assert lib().taa_oracle_post_process(C.byref(a), C.byref(d) if debug is not None else None, C.byref(b), w, h, C.byref(pc), nthreads) == 0
    return dst


def fxaa_push(w, h, subpix=0.75, edge_threshold=0.116, edge_threshold_min=0.0833):
    """push_constants_for_fxaa as update() fills it (taa.hpp:93-99, 953): rcpFrame = 1 / output resolution, in fp32."""
    pc = abi.TaaFxaaPush()
    pc.fxaaQualityRcpFrame[0] = np.float32(1.0) / np.float32(w)
    pc.fxaaQualityRcpFrame[1] = np.float32(1.0) / np.float32(h)
    pc.fxaaQualitySubpix, pc.fxaaQualityEdgeThreshold, pc.fxaaQualityEdgeThresholdMin = subpix, edge_threshold, edge_threshold_min
    return pc


def fxaa_prepare(src, nthreads=0, impl="oracle"):
    h, w = src.shape[:2]
    dst = np.zeros_like(src)
    a, b = _img(src), _img(dst)
    if impl == "ref":
        assert ref_lib().taa_ref_fxaa_prepare(C.byref(a), C.byref(b), w, h) == 0
    else:
        assert lib().taa_oracle_fxaa_prepare(C.byref(a), C.byref(b), w, h, nthreads) == 0
    return dst


def fxaa(src, segmask, pc: abi.TaaFxaaPush, gather4=True, nthreads=0, impl="oracle"):
    """antialias_fxaa.comp on an image prepared by fxaa_prepare (luma in alpha)."""
    h, w = src.shape[:2]
    dst = np.zeros_like(src)
    a, b, m = _img(src), _img(dst), _img(segmask)
    if impl == "ref":
        f = ref_lib().taa_ref_fxaa_gather4 if gather4 else ref_lib().taa_ref_fxaa_offset
        assert f(C.byref(a), C.byref(m), C.byref(b), w, h, C.byref(pc)) == 0
    else:
        assert lib().taa_oracle_fxaa(C.byref(a), C.byref(m), C.byref(b), w, h, C.byref(pc), 1 if gather4 else 0, nthreads) == 0
    return dst


def halton(i, b):
    return float(lib().taa_oracle_halton(i, b))


def jitter(frame_id, in_w, in_h, sample_distribution=1, fixed_index=-1, extra_scale=1.0, slow_motion=1, rotate_degrees=0.0, debug_offsets=None):
    out = (C.c_float * 2)()
    dbg, n = None, 0
    if debug_offsets:
        n = len(debug_offsets)
        dbg = (C.c_float * (2 * n))(*[v for xy in debug_offsets for v in xy])
    r = lib().taa_oracle_jitter(sample_distribution, fixed_index, extra_scale, slow_motion, rotate_degrees, dbg, n, in_w, in_h, frame_id, out)
    if r < 0:
        raise RuntimeError("taa_oracle_jitter failed")
    return (out[0], out[1]), r


def max_threads():
    return int(lib().taa_oracle_max_threads())
