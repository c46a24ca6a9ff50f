"""Synthetic jittered G-buffer sequences (SURVEY.md §8d "Synthetic inputs").

The reference's renderer (source/main.cpp, shaders/fwd_geometry.frag) is out of scope; what the resolve
consumes is its G-buffer contract, reproduced here analytically:
  colour   rgba16f  radiance in [0,1], evaluated at the jittered sample position (cf. shaders/testimage.frag:20-24)
  depth    D32      NDC depth of gvk's perspective projection (gears_vk/framework/src/camera.cpp:182-191), 0 = near
  velocity rgba16f  ((ndc - ndc_prev) * (0.5, 0.5, 1), movingObjectId), jitter removed (shaders/fwd_geometry.frag:289-295)
  matId    r32ui    (materialIndex + 1) | mover << 31 (fwd_geometry.frag:286)
  uvNrm    rgba32f  (uv, spherical view-space normal) (fwd_geometry.frag:280-283)
Scene: a textured plane at view depth 10 seen by a camera that translates so that the content moves by
`pan_px` per frame, plus one textured foreground quad at depth 4 that moves by `mover_px` per frame.
Everything is generated in fp32 and rounded (RTE) to the storage format. Runs on CPU or CUDA tensors.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch


def halton(i: int, b: int) -> float:
    """helpers::halton, source/helper_functions.hpp:9-17 (fp32 arithmetic emulated with torch scalars)."""
    f = torch.tensor(1.0, dtype=torch.float32)
    r = torch.tensor(0.0, dtype=torch.float32)
    bb = torch.tensor(float(b), dtype=torch.float32)
    while i > 0:
        f = f / bb
        r = r + f * torch.tensor(float(i % b), dtype=torch.float32)
        i //= b
    return float(r)


def halton_2_3_ndc(n: int, width: int, height: int):
    """helpers::halton_2_3<n>(2/res) — NDC jitter offsets of pattern 2/3 (taa.hpp:170-171)."""
    out = []
    for i in range(n):
        px = torch.tensor(2.0, dtype=torch.float32) / torch.tensor(float(width), dtype=torch.float32)
        py = torch.tensor(2.0, dtype=torch.float32) / torch.tensor(float(height), dtype=torch.float32)
        hx = torch.tensor(halton(i + 1, 2), dtype=torch.float32) - 0.5
        hy = torch.tensor(halton(i + 1, 3), dtype=torch.float32) - 0.5
        out.append((float(px * hx), float(py * hy)))
    return out


@dataclass
class Frame:
    index: int
    color: torch.Tensor      # (H, W, 4) float16
    depth: torch.Tensor      # (H, W) float32
    velocity: torch.Tensor   # (H, W, 4) float16
    matid: torch.Tensor      # (H, W) int32 (bit pattern of r32ui)
    uvnrm: torch.Tensor      # (H, W, 4) float32
    jitter_ndc: tuple        # (x, y)
    view: list               # 16 floats, column-major
    proj: list               # 16 floats, column-major, UN-jittered (what save_history_proj_matrix gets, main.cpp:4123)


class SyntheticScene:
    def __init__(self, width: int, height: int, pan_px=(3.0, 0.5), mover_px=(-6.0, 0.0), seed: int = 0x7AA57A2,
                 device="cpu", fov_deg: float = 60.0, near: float = 0.1, far: float = 100.0, jitter_len: int = 8,
                 plane_depth: float = 10.0, mover_depth: float = 4.0, with_aux: bool = True, rows=None):
        self.W, self.H = int(width), int(height)
        self.pan = (float(pan_px[0]), float(pan_px[1]))
        self.mover = (float(mover_px[0]), float(mover_px[1]))
        self.device = torch.device(device)
        self.near, self.far = float(near), float(far)
        self.fov = math.radians(fov_deg)
        self.aspect = self.W / self.H
        self.plane_depth, self.mover_depth = float(plane_depth), float(mover_depth)
        self.with_aux = with_aux
        self.jitter = halton_2_3_ndc(jitter_len, self.W, self.H)
        g = torch.Generator(device="cpu").manual_seed(seed & 0x7FFFFFFF)
        # 6 gratings: direction, spatial frequency (cycles / px, up to ~0.45), amplitude, per-channel phase
        ang = torch.rand(6, generator=g) * math.pi
        freq = torch.tensor([0.004, 0.011, 0.035, 0.09, 0.21, 0.43])
        self.kx = (torch.cos(ang) * freq * 2 * math.pi).to(self.device)
        self.ky = (torch.sin(ang) * freq * 2 * math.pi).to(self.device)
        self.amp = torch.tensor([0.16, 0.12, 0.09, 0.07, 0.06, 0.05]).to(self.device)
        self.phase = (torch.rand(6, 3, generator=g) * 2 * math.pi).to(self.device)
        self.m0 = (0.62 * self.W, 0.45 * self.H)          # mover centre at frame 0 (px)
        self.mhalf = (0.11 * self.W, 0.16 * self.H)        # mover half extent (px)
        # rows = (a, b): generate only image rows [a, b) of the W x H frame (one band of a sharded frame)
        self.row0, self.row1 = rows if rows is not None else (0, self.H)
        ys, xs = torch.meshgrid(torch.arange(self.row0, self.row1, dtype=torch.float32, device=self.device),
                                torch.arange(self.W, dtype=torch.float32, device=self.device), indexing="ij")
        self.xs, self.ys = xs + 0.5, ys + 0.5

    # ---- camera (gvk conventions) ------------------------------------------------------------
    def _xy_scale(self):
        return 1.0 / math.tan(self.fov / 2.0)

    def proj_matrix(self):
        s = self._xy_scale()
        z_scale = self.far / (self.far - self.near)
        m = [0.0] * 16  # column-major: m[col*4 + row]; P = M * diag(1,-1,-1,1) (camera.cpp:153-195)
        m[0] = s / self.aspect
        m[5] = -s
        m[10] = -z_scale
        m[11] = -1.0
        m[14] = -self.near * z_scale
        return m

    def cam_pos(self, n: int):
        s = self._xy_scale()
        d = self.plane_depth
        # content moves by +pan px per frame on screen
        cx = -self.pan[0] * n * 2.0 * d * self.aspect / (s * self.W)
        cy = self.pan[1] * n * 2.0 * d / (s * self.H)
        return (cx, cy, 0.0)

    def view_matrix(self, n: int):
        c = self.cam_pos(n)
        return [1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, -c[0], -c[1], -c[2], 1]

    def ndc_depth(self, d: float) -> float:
        z_scale = self.far / (self.far - self.near)
        return z_scale * (1.0 - self.near / d)

    # ---- content -----------------------------------------------------------------------------
    def _pattern(self, px, py, variant: int):
        """Procedural radiance in [0,1]: gratings + 1-px checker stripes + hard edges."""
        out = []
        for c in range(3):
            v = torch.full_like(px, 0.5)
            for i in range(6):
                v = v + self.amp[i] * torch.sin(self.kx[i] * px + self.ky[i] * py + self.phase[i, c] + 1.7 * variant)
            out.append(v)
        rgb = torch.stack(out, dim=-1)
        # hard vertical/horizontal edges every 97 / 61 px
        edge = ((torch.floor(px / 97.0) + torch.floor(py / 61.0)) % 2.0) * 0.18 - 0.09
        rgb = rgb + edge.unsqueeze(-1)
        # a band of 1-px checker stripes (maximum frequency content)
        checker = ((torch.floor(px) + torch.floor(py)) % 2.0) * 0.5 - 0.25
        band = ((torch.floor(py / 40.0) % 5.0) == 2.0).to(px.dtype)
        rgb = rgb + (checker * band).unsqueeze(-1) * torch.tensor([1.0, 0.8, 0.6], device=px.device)
        return rgb.clamp(0.0, 1.0)

    def frame(self, n: int) -> Frame:
        W, H = self.W, self.H
        j = self.jitter[n % len(self.jitter)]
        jpx, jpy = j[0] * W * 0.5, j[1] * H * 0.5  # jitter in px; the sample sees the point at ndc - jitter
        sx, sy = self.xs - jpx, self.ys - jpy
        # background: pattern anchored to the world, world moves by +pan px per frame
        bg = self._pattern(sx - self.pan[0] * n, sy - self.pan[1] * n, 0)
        mcx, mcy = self.m0[0] + self.mover[0] * n, self.m0[1] + self.mover[1] * n
        inside = ((sx - mcx).abs() <= self.mhalf[0]) & ((sy - mcy).abs() <= self.mhalf[1])
        fg = self._pattern((sx - mcx) * 1.3 + 1000.0, (sy - mcy) * 1.3 + 500.0, 1)
        rgb = torch.where(inside.unsqueeze(-1), fg, bg)
        color = torch.cat([rgb, torch.ones_like(rgb[..., :1])], dim=-1).to(torch.float16)
        zb, zm = self.ndc_depth(self.plane_depth), self.ndc_depth(self.mover_depth)
        depth = torch.where(inside, torch.full_like(sx, zm), torch.full_like(sx, zb)).to(torch.float32)
        vel = torch.zeros(self.row1 - self.row0, W, 4, dtype=torch.float32, device=self.device)
        vel[..., 0] = torch.where(inside, torch.full_like(sx, self.mover[0] / W), torch.full_like(sx, self.pan[0] / W))
        vel[..., 1] = torch.where(inside, torch.full_like(sx, self.mover[1] / H), torch.full_like(sx, self.pan[1] / H))
        vel[..., 3] = inside.to(torch.float32)
        velocity = vel.to(torch.float16)
        if self.with_aux:
            mat_bg = torch.where((torch.floor((sx - self.pan[0] * n) / 256.0) % 2.0) == 0.0, 1, 2).to(torch.int64)
            matid = torch.where(inside, torch.full_like(mat_bg, 3 | 0x80000000), mat_bg)
            matid = (matid & 0xFFFFFFFF).to(torch.int64)
            matid = torch.where(matid >= 2 ** 31, matid - 2 ** 32, matid).to(torch.int32)
            nx = 0.1 * torch.sin((sx - self.pan[0] * n) / 50.0)
            ny = 0.1 * torch.cos((sy - self.pan[1] * n) / 40.0)
            nz = torch.ones_like(nx)
            nx = torch.where(inside, torch.full_like(nx, 0.6), nx)
            ny = torch.where(inside, torch.zeros_like(ny), ny)
            nz = torch.where(inside, torch.full_like(nz, 0.8), nz)
            ln = torch.sqrt(nx * nx + ny * ny + nz * nz)
            nx, ny, nz = nx / ln, ny / ln, nz / ln
            l2 = torch.sqrt(nx * nx + ny * ny)
            az = torch.where(l2 == 0, torch.zeros_like(l2), torch.acos((nx / l2.clamp_min(1e-30)).clamp(-1, 1)))
            az = torch.where(ny < 0, 2 * math.pi - az, az)
            el = torch.asin(nz.clamp(-1, 1))
            uvnrm = torch.stack([self.xs / W, self.ys / H, az, el], dim=-1).to(torch.float32)
        else:
            matid = torch.zeros(0, dtype=torch.int32, device=self.device)
            uvnrm = torch.zeros(0, dtype=torch.float32, device=self.device)
        return Frame(n, color.contiguous(), depth.contiguous(), velocity.contiguous(), matid.contiguous(), uvnrm.contiguous(),
                     j, self.view_matrix(n), self.proj_matrix())
