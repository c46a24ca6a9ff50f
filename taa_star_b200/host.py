"""Python face of the C-ABI: `TaaContext` (taa_create / taa_resolve_ex / taa_frame ...) and `Taa`, the mirror of
`template<size_t CF> class taa : gvk::invokee` (reference source/taa.hpp:26-1427) with the reference's method names.

torch is used for device memory and streams only; every pixel is computed by libtaa_b200.so.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Sequence, Tuple, Union

import torch

from . import abi
from .abi import (TaaCasPush, TaaError, TaaFxaaPush, TaaParameters, TaaPostProcessPush, TaaSharpenPush, TaaUniforms, taa_desc, taa_image,
                  taa_post_chain, taa_resolve_images, taa_source_views)

# bytes per texel of every binding of taa.comp (shaders/shader_cpu_common.h:60-77)
_BPT = {"color": 8, "depth": 4, "velocity": 8, "history_in": 8, "history_depth": 4, "history_out": 8, "result": 8, "debug": 8,
        "segmask": 4, "prev_segmask": 4, "matid": 4, "prev_matid": 4, "uvnrm": 16, "mask": 4}

ImageArg = Union[torch.Tensor, Tuple[torch.Tensor, int]]


def _image(t: Optional[ImageArg], bpt: int, require_cuda: bool = True) -> taa_image:
    """taa_image of a (rows, width, ...) tensor; `(tensor, y0)` for a band buffer whose row 0 is global row y0."""
    if t is None:
        return taa_image(None, 0, 0, 0)
    y0 = 0
    if isinstance(t, tuple):
        t, y0 = t
    if require_cuda and not t.is_cuda:
        raise TaaError(abi.TAA_E_INVALID_ARG, "images must be CUDA tensors (there is no CPU path)")
    row_bytes = t.stride(0) * t.element_size()
    inner = t[0]
    if not inner.is_contiguous():
        raise TaaError(abi.TAA_E_INVALID_ARG, "image rows must be contiguous")
    if inner.numel() * inner.element_size() % bpt:
        raise TaaError(abi.TAA_E_INVALID_ARG, "image row is not a whole number of texels")
    return taa_image(t.data_ptr(), row_bytes, int(y0), int(t.shape[0]))


def _stream_ptr(stream: Optional[torch.cuda.Stream]) -> int:
    if stream is None:
        stream = torch.cuda.current_stream()
    return stream.cuda_stream


class TaaContext:
    """One taa_ctx: a whole frame, or one row band of a larger frame (band = (first_row, rows))."""

    def __init__(self, in_size: Tuple[int, int], out_size: Optional[Tuple[int, int]] = None, band: Optional[Tuple[int, int]] = None,
                 device: int = -1, flags: int = abi.TAA_FLAG_DEFAULT):
        self._lib = abi.load_library()
        out_size = out_size or in_size
        band = band or (0, out_size[1])
        self.in_w, self.in_h = in_size
        self.out_w, self.out_h = out_size
        self.band = band
        d = taa_desc(C.sizeof(taa_desc), abi.ABI_VERSION, self.in_w, self.in_h, self.out_w, self.out_h, band[0], band[1], device, flags)
        h = C.c_void_p()
        st = self._lib.taa_create(C.byref(h), C.byref(d))
        if st != abi.TAA_OK:
            raise TaaError(st, f"taa_create: {self._lib.taa_status_string(st).decode()}: {self._lib.taa_last_error_string(None).decode()}")
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._lib.taa_destroy(self._h)
            self._h = None

    __del__ = close

    def _check(self, st: int, what: str):
        if st != abi.TAA_OK:
            raise TaaError(st, f"{what}: {self._lib.taa_status_string(st).decode()}: {self._lib.taa_last_error_string(self._h).decode()}")

    @staticmethod
    def images(**kw: Optional[ImageArg]) -> taa_resolve_images:
        im = taa_resolve_images()
        for name, t in kw.items():
            if name not in _BPT:
                raise KeyError(name)
            setattr(im, name, _image(t, _BPT[name]))
        return im

    def resolve(self, uniforms: TaaUniforms, stream: Optional[torch.cuda.Stream] = None, **images: Optional[ImageArg]):
        """taa_resolve_ex: one dispatch of taa.comp. Keyword names are the fields of taa_resolve_images."""
        im = self.images(**images)
        self._check(self._lib.taa_resolve_ex(self._h, C.byref(im), C.byref(uniforms), _stream_ptr(stream)), "taa_resolve_ex")

    def resolve_prepared(self, im: taa_resolve_images, uniforms: TaaUniforms, stream_ptr: int):
        self._check(self._lib.taa_resolve_ex(self._h, C.byref(im), C.byref(uniforms), stream_ptr), "taa_resolve_ex")

    def resolve_simple(self, color, depth, motion, history_in, history_out, uniforms: TaaUniforms, stream=None):
        """taa_resolve(ctx, color, depth, motion, history_in, history_out, params) — the north-star signature."""
        for t in (color, depth, motion, history_in, history_out):
            if not (t.is_cuda and t.is_contiguous()):
                raise TaaError(abi.TAA_E_INVALID_ARG, "taa_resolve takes contiguous CUDA tensors")
        self._check(self._lib.taa_resolve(self._h, color.data_ptr(), depth.data_ptr(), motion.data_ptr(), history_in.data_ptr(),
                                          history_out.data_ptr(), C.byref(uniforms), _stream_ptr(stream)), "taa_resolve")

    def frame(self, uniforms: TaaUniforms, chain: taa_post_chain, final: torch.Tensor, stream=None, **images: Optional[ImageArg]):
        """taa_frame: taa.comp + [sharpen | CAS] + [post-process] as render() records them (taa.hpp:1008-1159)."""
        im = self.images(**images)
        fin = _image(final, 8)
        self._check(self._lib.taa_frame(self._h, C.byref(im), C.byref(uniforms), C.byref(chain), C.byref(fin), _stream_ptr(stream)), "taa_frame")

    def frame_prepared(self, im: taa_resolve_images, uniforms: TaaUniforms, chain: taa_post_chain, fin, stream_ptr: int):
        """taa_frame on pre-built argument blocks (fin = the taa_image of the final image): no per-call Python marshalling."""
        self._check(self._lib.taa_frame(self._h, C.byref(im), C.byref(uniforms), C.byref(chain), C.byref(fin), stream_ptr), "taa_frame")

    @staticmethod
    def image(t, bpt: int = 8):
        return _image(t, bpt)

    def sharpen(self, src, dst, factor: float, stream=None):
        pc = TaaSharpenPush(factor)
        a, b = _image(src, 8), _image(dst, 8)
        self._check(self._lib.taa_sharpen(self._h, C.byref(a), C.byref(b), C.byref(pc), _stream_ptr(stream)), "taa_sharpen")

    def sharpen_cas(self, src, dst, pc: TaaCasPush, stream=None):
        a, b = _image(src, 8), _image(dst, 8)
        self._check(self._lib.taa_sharpen_cas(self._h, C.byref(a), C.byref(b), C.byref(pc), _stream_ptr(stream)), "taa_sharpen_cas")

    def post_process(self, src, debug, dst, pc: TaaPostProcessPush, stream=None):
        a, d, b = _image(src, 8), _image(debug, 8), _image(dst, 8)
        self._check(self._lib.taa_post_process(self._h, C.byref(a), C.byref(d) if debug is not None else None, C.byref(b), C.byref(pc),
                                               _stream_ptr(stream)), "taa_post_process")

    def fxaa_prepare(self, src, dst, stream=None):
        a, b = _image(src, 8), _image(dst, 8)
        self._check(self._lib.taa_fxaa_prepare(self._h, C.byref(a), C.byref(b), _stream_ptr(stream)), "taa_fxaa_prepare")

    def fxaa(self, src, segmask, dst, pc: TaaFxaaPush, fused: bool = False, stream=None):
        """antialias_fxaa.comp on a prepared image; fused=True takes the unprepared screen result (prepare + fxaa in one launch)."""
        a, m, b = _image(src, 8), _image(segmask, 4), _image(dst, 8)
        f = self._lib.taa_fxaa_fused if fused else self._lib.taa_fxaa
        self._check(f(self._h, C.byref(a), C.byref(m), C.byref(b), C.byref(pc), _stream_ptr(stream)), "taa_fxaa")

    def poll_status(self, stream=None) -> int:
        return self._lib.taa_poll_status(self._h, _stream_ptr(stream))

    def fixup_pixels(self, stream=None) -> int:
        """Pixels the last resolve handed from the tuned kernel to the exact fix-up pass (synchronises)."""
        return int(self._lib.taa_fixup_pixels(self._h, _stream_ptr(stream)))

    @property
    def launch_count(self) -> int:
        return int(self._lib.taa_launch_count(self._h))


def write_settings_ini(params, settings: "abi.taa_invokee_settings", post: "abi.TaaPostProcessPush") -> str:
    """writeSettingsToIni (taa.hpp:1198-1265) on plain parameter blocks: params = (TaaParameters, TaaParameters). No GPU needed."""
    lib = abi.load_library()
    arr = (abi.TaaParameters * 2)(params[0], params[1])
    n = lib.taa_settings_write_ini(arr, C.byref(settings), C.byref(post), None, 0)
    buf = C.create_string_buffer(n)
    lib.taa_settings_write_ini(arr, C.byref(settings), C.byref(post), buf, n)
    return buf.value.decode()


def read_settings_ini(text: str, params, settings: "abi.taa_invokee_settings", post: "abi.TaaPostProcessPush", max_offsets: int = 64):
    """readSettingsFromIni (taa.hpp:1267-1339): updates params[0], params[1], settings and post in place; returns the list of
    mDebugSampleOffsets (vec2 tuples). Keys that are absent or empty keep their current value. No GPU needed."""
    lib = abi.load_library()
    arr = (abi.TaaParameters * 2)(params[0], params[1])
    offs = (C.c_float * (2 * max_offsets))()
    st = lib.taa_settings_read_ini(text.encode(), arr, C.byref(settings), C.byref(post), offs, max_offsets)
    C.memmove(C.byref(params[0]), C.byref(arr[0]), C.sizeof(abi.TaaParameters))
    C.memmove(C.byref(params[1]), C.byref(arr[1]), C.sizeof(abi.TaaParameters))
    n = settings.jitter.mDebugSampleOffsetsCount
    out = [(offs[2 * i], offs[2 * i + 1]) for i in range(n)]
    settings._offsets_keepalive = offs  # settings.jitter.mDebugSampleOffsets points into it
    if st != abi.TAA_OK:
        raise TaaError(st, lib.taa_settings_ini_last_error().decode())
    return out


def cas_setup(sharpness: float, out_w: int, out_h: int) -> TaaCasPush:
    """CasSetup as update() calls it (taa.hpp:965)."""
    pc = TaaCasPush()
    abi.load_library().taa_cas_setup(C.byref(pc), sharpness, float(out_w), float(out_h))
    return pc


def fxaa_default(w: int, h: int) -> TaaFxaaPush:
    """push_constants_for_fxaa as update() fills them (taa.hpp:93-99, 953)."""
    pc = TaaFxaaPush()
    abi.load_library().taa_fxaa_default(C.byref(pc), w, h)
    return pc


def postprocess_default(w: int, h: int) -> TaaPostProcessPush:
    pp = TaaPostProcessPush()
    abi.load_library().taa_postprocess_default(C.byref(pp), w, h)
    return pp


def jitter_offset_for_frame(frame_id: int, in_w: int, in_h: int, sample_distribution: int = 1, fixed_index: int = -1,
                            extra_scale: float = 1.0, slow_motion: int = 1, rotate_degrees: float = 0.0,
                            debug_offsets: Optional[Sequence[Tuple[float, float]]] = None) -> Tuple[Tuple[float, float], int]:
    """get_jitter_offset_for_frame (taa.hpp:150-233) -> ((x, y) in NDC, pattern length)."""
    s = abi.taa_jitter_settings(sample_distribution, fixed_index, extra_scale, slow_motion, rotate_degrees, None, 0)
    keep = None
    if debug_offsets:
        keep = (C.c_float * (2 * len(debug_offsets)))(*[v for xy in debug_offsets for v in xy])
        s.mDebugSampleOffsets = C.cast(keep, C.POINTER(C.c_float))
        s.mDebugSampleOffsetsCount = len(debug_offsets)
    out = (C.c_float * 2)()
    n = abi.load_library().taa_jitter_offset_for_frame(C.byref(s), in_w, in_h, frame_id, out)
    if n < 0:
        raise TaaError(n, "taa_jitter_offset_for_frame")
    return (out[0], out[1]), n


def _mat(m: Sequence[float]):
    return (C.c_float * 16)(*[float(v) for v in m])


class Taa:
    """`class taa<CF>` (taa.hpp). Owns result/history/temp/debug/post-process/seg-mask images x CF and the history ring."""

    def __init__(self, concurrent_frames: int = 3, device: int = -1, flags: int = abi.TAA_FLAG_DEFAULT):
        self._lib = abi.load_library()
        h = C.c_void_p()
        st = self._lib.taa_invokee_create(C.byref(h), concurrent_frames, device, flags)
        if st != abi.TAA_OK:
            raise TaaError(st, "taa_invokee_create failed")
        self._h = h
        self.CF = concurrent_frames
        self._keep = None
        self.out_size = None

    def close(self):
        if getattr(self, "_h", None):
            self._lib.taa_invokee_destroy(self._h)
            self._h = None

    __del__ = close

    def _check(self, st: int, what: str):
        if st != abi.TAA_OK:
            raise TaaError(st, f"{what}: {self._lib.taa_status_string(st).decode()}: {self._lib.taa_invokee_last_error(self._h).decode()}")

    # ---- settings surface (pointers into the C++ object; assign fields directly) ----
    @property
    def mParameters(self):
        return [self._lib.taa_invokee_parameters(self._h, 0).contents, self._lib.taa_invokee_parameters(self._h, 1).contents]

    @property
    def settings(self) -> abi.taa_invokee_settings:
        return self._lib.taa_invokee_settings_ptr(self._h).contents

    @property
    def mPostProcessPushConstants(self) -> TaaPostProcessPush:
        return self._lib.taa_invokee_postprocess(self._h).contents

    @property
    def uniforms(self) -> TaaUniforms:
        return self._lib.taa_invokee_uniforms(self._h).contents

    def taa_enabled(self) -> bool:  # taa.hpp:141
        return bool(self.settings.mTaaEnabled)

    def execution_order(self) -> int:  # taa.hpp:139
        return 100

    # ---- taa.hpp:263 ----
    def set_source_image_views(self, target_resolution: Tuple[int, int], color: Sequence[torch.Tensor], depth: Sequence[torch.Tensor],
                               uvnrm: Optional[Sequence[Optional[torch.Tensor]]], velocity: Sequence[torch.Tensor],
                               matid: Optional[Sequence[Optional[torch.Tensor]]] = None, raytraced=None):
        CF = self.CF
        assert len(color) == len(depth) == len(velocity) == CF
        in_h, in_w = depth[0].shape[:2]
        views = (taa_source_views * CF)()
        keep = []
        for i in range(CF):
            for name, seq in (("color", color), ("depth", depth), ("uvnrm", uvnrm), ("velocity", velocity), ("matid", matid)):
                t = seq[i] if seq is not None else None
                if t is not None and t.numel():
                    if not (t.is_cuda and t.is_contiguous()):
                        raise TaaError(abi.TAA_E_INVALID_ARG, f"{name}[{i}] must be a contiguous CUDA tensor")
                    setattr(views[i], name, t.data_ptr())
                    keep.append(t)
        self._keep = keep
        self.out_size = tuple(target_resolution)
        self.in_size = (in_w, in_h)
        self._check(self._lib.taa_invokee_set_source_image_views(self._h, target_resolution[0], target_resolution[1], in_w, in_h, views),
                    "set_source_image_views")

    def set_sizes_for_host_frames(self, target_resolution: Tuple[int, int], in_size: Tuple[int, int]):
        """set_source_image_views without device G-buffers: frames arrive through frame_host()."""
        self.out_size = tuple(target_resolution)
        self.in_size = tuple(in_size)
        self._check(self._lib.taa_invokee_set_source_image_views(self._h, target_resolution[0], target_resolution[1], in_size[0], in_size[1], None),
                    "set_source_image_views")

    # ---- taa.hpp:150 / 243 / 235 ----
    def get_jittered_projection_matrix(self, proj: Sequence[float], frame_id: int):
        out, jit = (C.c_float * 16)(), (C.c_float * 2)()
        self._check(self._lib.taa_invokee_get_jittered_projection_matrix(self._h, _mat(proj), frame_id, out, jit), "get_jittered_projection_matrix")
        return list(out), (jit[0], jit[1])

    def save_history_proj_matrix(self, proj: Sequence[float], frame_id: int):
        self._check(self._lib.taa_invokee_save_history_proj_matrix(self._h, _mat(proj), frame_id), "save_history_proj_matrix")

    # ---- invokee overrides: taa.hpp:894 / 974 ----
    def update(self, frame_id: int, view: Sequence[float], time_s: float = 0.0, cam_near: float = 0.1, cam_far: float = 100.0):
        self._check(self._lib.taa_invokee_update(self._h, frame_id, _mat(view), time_s, cam_near, cam_far), "update")

    def render(self, frame_id: int, stream: Optional[torch.cuda.Stream] = None) -> int:
        """Enqueues the frame; returns the device pointer of the image render() would blit to the swapchain."""
        out = C.c_void_p()
        self._check(self._lib.taa_invokee_render(self._h, frame_id, _stream_ptr(stream), C.byref(out)), "render")
        return out.value

    def writeSettingsToIni(self) -> str:  # taa.hpp:1198
        n = self._lib.taa_invokee_write_settings_ini(self._h, None, 0)
        buf = C.create_string_buffer(n)
        self._lib.taa_invokee_write_settings_ini(self._h, buf, n)
        return buf.value.decode()

    def readSettingsFromIni(self, text: str):  # taa.hpp:1267
        st = self._lib.taa_invokee_read_settings_ini(self._h, text.encode())
        if st != abi.TAA_OK:
            raise TaaError(st, self._lib.taa_settings_ini_last_error().decode())

    def duration(self) -> float:  # taa.hpp:367
        return float(self._lib.taa_invokee_duration(self._h))

    def image(self, which: int, slot: int, dtype=torch.float16) -> torch.Tensor:
        """A torch view of an owned image (TAA_IMG_*), for inspection."""
        ptr = self._lib.taa_invokee_image(self._h, which, slot)
        if not ptr:
            raise TaaError(abi.TAA_E_INVALID_ARG, "no such image")
        w, h = self.out_size
        return tensor_from_ptr(ptr, (h, w) if which == abi.TAA_IMG_SEGMASK else (h, w, 4), torch.int32 if which == abi.TAA_IMG_SEGMASK else dtype)

    def image_by_ptr(self, ptr: int) -> torch.Tensor:
        w, h = self.out_size
        return tensor_from_ptr(ptr, (h, w, 4), torch.float16)

    @property
    def launch_count(self) -> int:
        return int(self._lib.taa_invokee_launch_count(self._h))

    # ---- host-buffer frames ----
    @property
    def h2d_bytes(self) -> int:
        """Bytes frame_host has uploaded so far (depth travels only when a kernel of the frame reads it)."""
        return int(self._lib.taa_invokee_h2d_bytes(self._h))

    def frame_host(self, frame_id: int, color: torch.Tensor, depth: torch.Tensor, velocity: torch.Tensor, view, proj, out_final: torch.Tensor,
                   time_s: float = 0.0, cam_near: float = 0.1, cam_far: float = 100.0, uvnrm=None, matid=None):
        v = taa_source_views(color.data_ptr(), depth.data_ptr(), uvnrm.data_ptr() if uvnrm is not None else None, velocity.data_ptr(),
                             matid.data_ptr() if matid is not None else None, None)
        self._check(self._lib.taa_invokee_frame_host(self._h, frame_id, C.byref(v), _mat(view), _mat(proj), time_s, cam_near, cam_far,
                                                     out_final.data_ptr()), "frame_host")

    def wait(self, frame_id: int):
        self._check(self._lib.taa_invokee_wait(self._h, frame_id), "wait")


class _PtrHolder:
    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def tensor_from_ptr(ptr: int, shape, dtype) -> torch.Tensor:
    n = 1
    for s in shape:
        n *= s
    nbytes = n * torch.empty((), dtype=dtype).element_size()
    return torch.as_tensor(_PtrHolder(ptr, nbytes), device="cuda").view(dtype).view(*shape)
