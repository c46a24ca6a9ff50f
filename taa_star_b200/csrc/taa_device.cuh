// taa_device.cuh — device-side building blocks shared by the resolve kernels (sm_100a).
//
// Arithmetic contract of the EXACT kernels: every fp32 operation of shaders/taa.comp is performed
// once, in source order, with IEEE round-to-nearest (this file is compiled with --fmad=false; the
// default -prec-div/-prec-sqrt keep '/' and sqrtf correctly rounded). The sampler follows the
// Vulkan linear-filter equations in fp32 (taa.hpp:274 creates one bilinear clamp-to-edge sampler
// that every `texture()` call in taa.comp uses).
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>
#include "../../include/taa_b200.h"

namespace taa {

// one image binding as the kernels see it
struct Img {
	const unsigned char* __restrict__ p;
	long long pitch;
	int y0;    // global row stored in buffer row 0
	int rows;  // rows present
};
struct ImgW {
	unsigned char* __restrict__ p;
	long long pitch;
	int y0;
	int rows;
};

struct ResolveArgs {
	Img color, depth, velocity, history_in, history_depth, prev_segmask, matid, prev_matid, uvnrm;
	ImgW history_out, result, debug, segmask, mask;
	int in_w, in_h, out_w, out_h;
	int band_y0, band_rows;
	unsigned int* status;  // device word: bit0 = a read left the rows held by a band buffer
	TaaUniforms ubo;
	// the follow-on pass fused into the resolve (streaming kernel only; set by taa_frame when the chain allows it):
	ImgW final_img;        // what sharpen.comp | sharpen_cas.comp (+ an identity post_process.comp) would have written
	int epilogue;          // 0: none, 1: sharpen.comp, 2: sharpen_cas.comp
	float epilogue_k;      // sharpening factor | CAS peak (const1.x)
};

// ---- scalar helpers -----------------------------------------------------------------------------
__device__ __forceinline__ float mixf(float x, float y, float a) { return x * (1.0f - a) + y * a; }
__device__ __forceinline__ float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
__device__ __forceinline__ float lerpf(float p, float q, float w) { return p + w * (q - p); }
__device__ __forceinline__ int iclamp(int v, int lo, int hi) { return min(max(v, lo), hi); }

struct f3 { float x, y, z; };
__device__ __forceinline__ f3 mk3(float x, float y, float z) { f3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ f3 operator+(f3 a, f3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ f3 operator-(f3 a, f3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ f3 operator*(f3 a, f3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ f3 operator*(f3 a, float b) { return mk3(a.x * b, a.y * b, a.z * b); }
__device__ __forceinline__ f3 operator*(float a, f3 b) { return mk3(a * b.x, a * b.y, a * b.z); }
__device__ __forceinline__ f3 operator/(f3 a, float b) { return mk3(a.x / b, a.y / b, a.z / b); }
__device__ __forceinline__ f3 operator/(f3 a, f3 b) { return mk3(a.x / b.x, a.y / b.y, a.z / b.z); }
__device__ __forceinline__ f3 min3(f3 a, f3 b) { return mk3(fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z)); }
__device__ __forceinline__ f3 max3(f3 a, f3 b) { return mk3(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z)); }
__device__ __forceinline__ f3 abs3(f3 a) { return mk3(fabsf(a.x), fabsf(a.y), fabsf(a.z)); }
__device__ __forceinline__ f3 xyz(float4 v) { return mk3(v.x, v.y, v.z); }
__device__ __forceinline__ float4 mk4(f3 v, float w) { return make_float4(v.x, v.y, v.z, w); }
__device__ __forceinline__ float4 operator*(float4 a, float b) { return make_float4(a.x * b, a.y * b, a.z * b, a.w * b); }
__device__ __forceinline__ float4 operator+(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

// ---- raw texel access ---------------------------------------------------------------------------
// Row pointer of global row gy. Rows outside the buffer (a band whose halo is too small) are clamped
// into it and reported through the status word.
template <class I>
__device__ __forceinline__ auto row_ptr(const I& im, int gy, unsigned int* status) -> decltype(im.p) {
	int ly = gy - im.y0;
	if (ly < 0 || ly >= im.rows) {
		if (status) atomicOr(status, 1u);
		ly = ly < 0 ? 0 : im.rows - 1;
	}
	return im.p + (long long)ly * im.pitch;
}

__device__ __forceinline__ float4 unpack_rgba16f(uint2 raw) {
	__half2 lo = *reinterpret_cast<__half2*>(&raw.x);
	__half2 hi = *reinterpret_cast<__half2*>(&raw.y);
	float2 a = __half22float2(lo), b = __half22float2(hi);
	return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ uint2 pack_rgba16f(float4 v) {
	__half2 lo = __floats2half2_rn(v.x, v.y);
	__half2 hi = __floats2half2_rn(v.z, v.w);
	uint2 r;
	r.x = *reinterpret_cast<unsigned int*>(&lo);
	r.y = *reinterpret_cast<unsigned int*>(&hi);
	return r;
}

// in-range texel reads (caller guarantees 0 <= x < W, 0 <= y < H)
__device__ __forceinline__ float4 ld_rgba16f(const Img& im, int x, int y, unsigned int* st) {
	return unpack_rgba16f(__ldg(reinterpret_cast<const uint2*>(row_ptr(im, y, st)) + x));
}
__device__ __forceinline__ float ld_r32f(const Img& im, int x, int y, unsigned int* st) {
	return __ldg(reinterpret_cast<const float*>(row_ptr(im, y, st)) + x);
}
__device__ __forceinline__ unsigned int ld_r32ui(const Img& im, int x, int y, unsigned int* st) {
	return __ldg(reinterpret_cast<const unsigned int*>(row_ptr(im, y, st)) + x);
}
__device__ __forceinline__ float4 ld_rgba32f(const Img& im, int x, int y, unsigned int* st) {
	return __ldg(reinterpret_cast<const float4*>(row_ptr(im, y, st)) + x);
}
// texelFetch / imageLoad: out of range -> 0; a NULL binding reads as 0 as well
__device__ __forceinline__ float4 fetch_rgba16f(const Img& im, int w, int h, int x, int y, unsigned int* st) {
	if (!im.p || x < 0 || y < 0 || x >= w || y >= h) return make_float4(0.f, 0.f, 0.f, 0.f);
	return ld_rgba16f(im, x, y, st);
}
__device__ __forceinline__ float fetch_r32f(const Img& im, int w, int h, int x, int y, unsigned int* st) {
	if (!im.p || x < 0 || y < 0 || x >= w || y >= h) return 0.f;
	return ld_r32f(im, x, y, st);
}
__device__ __forceinline__ unsigned int fetch_r32ui(const Img& im, int w, int h, int x, int y, unsigned int* st) {
	if (!im.p || x < 0 || y < 0 || x >= w || y >= h) return 0u;
	return ld_r32ui(im, x, y, st);
}
__device__ __forceinline__ float4 fetch_rgba32f(const Img& im, int w, int h, int x, int y, unsigned int* st) {
	if (!im.p || x < 0 || y < 0 || x >= w || y >= h) return make_float4(0.f, 0.f, 0.f, 0.f);
	return ld_rgba32f(im, x, y, st);
}

__device__ __forceinline__ void st_rgba16f(const ImgW& im, int x, int y, float4 v) {
	if (!im.p) return;
	reinterpret_cast<uint2*>(row_ptr(im, y, (unsigned int*)nullptr))[x] = pack_rgba16f(v);
}
__device__ __forceinline__ void st_r32ui(const ImgW& im, int x, int y, unsigned int v) {
	if (!im.p) return;
	reinterpret_cast<unsigned int*>(row_ptr(im, y, (unsigned int*)nullptr))[x] = v;
}

// ---- the sampler --------------------------------------------------------------------------------
// One axis of the Vulkan linear footprint: u = s*size - 0.5, i0 = floor(u), a = u - i0, clamp-to-edge.
struct Lin { int i0, i1; float a; };
__device__ __forceinline__ Lin lin_coord(float s, int size) {
	float u = s * (float)size - 0.5f;
	float fl = floorf(u);
	Lin c;
	c.a = u - fl;
	int i0 = (int)fl;  // cvt.rzi saturates, NaN -> 0
	c.i0 = iclamp(i0, 0, size - 1);
	c.i1 = iclamp(i0 == 2147483647 ? i0 : i0 + 1, 0, size - 1);
	return c;
}
__device__ __forceinline__ float4 lerp4(float4 p, float4 q, float w) {
	return make_float4(lerpf(p.x, q.x, w), lerpf(p.y, q.y, w), lerpf(p.z, q.z, w), lerpf(p.w, q.w, w));
}
// texture(sampler2D(tex, uSampler), uv) on rgba16f
__device__ __forceinline__ float4 tex_rgba16f(const Img& im, int w, int h, float s, float t, unsigned int* st) {
	Lin cx = lin_coord(s, w), cy = lin_coord(t, h);
	const uint2* r0 = reinterpret_cast<const uint2*>(row_ptr(im, cy.i0, st));
	const uint2* r1 = reinterpret_cast<const uint2*>(row_ptr(im, cy.i1, st));
	float4 t00 = unpack_rgba16f(__ldg(r0 + cx.i0)), t10 = unpack_rgba16f(__ldg(r0 + cx.i1));
	float4 t01 = unpack_rgba16f(__ldg(r1 + cx.i0)), t11 = unpack_rgba16f(__ldg(r1 + cx.i1));
	return lerp4(lerp4(t00, t10, cx.a), lerp4(t01, t11, cx.a), cy.a);
}
// texture(sampler2D(uCurrentDepth, uSampler), uv).r on D32
__device__ __forceinline__ float tex_r32f(const Img& im, int w, int h, float s, float t, unsigned int* st) {
	Lin cx = lin_coord(s, w), cy = lin_coord(t, h);
	const float* r0 = reinterpret_cast<const float*>(row_ptr(im, cy.i0, st));
	const float* r1 = reinterpret_cast<const float*>(row_ptr(im, cy.i1, st));
	float t00 = __ldg(r0 + cx.i0), t10 = __ldg(r0 + cx.i1), t01 = __ldg(r1 + cx.i0), t11 = __ldg(r1 + cx.i1);
	return lerpf(lerpf(t00, t10, cx.a), lerpf(t01, t11, cx.a), cy.a);
}

// sin and cos as this repository pins them. GLSL leaves their precision to the implementation (and no Vulkan driver exists on either box), so the
// oracle, the reference-shader shim (oracle/glsl_shim.h) and the CUDA kernels (taa_device.cuh) all evaluate THIS function text: Cody-Waite
// reduction by pi/2 in three steps, the single-precision minimax polynomials of the Cephes library on [-pi/4, pi/4], one IEEE binary32 operation
// per written operation (the three files are compiled without contraction). Arguments too large to reduce (>= 1e9) read as 0; non-finite ones give NaN.
__device__ __forceinline__ void taa_sincos(float x, float* s, float* c) {
	if (!(fabsf(x) < 1.0e9f)) { *s = x - x; *c = (x - x) + 1.0f; return; }
	const float k = floorf(x * 0.636619772f + 0.5f);
	float r = x - k * 1.5703125f;
	r = r - k * 4.837512969970703125e-4f;
	r = r - k * 7.54978995489188216e-8f;
	const float z = r * r;
	const float ps = ((-1.9515295891e-4f * z + 8.3321608736e-3f) * z - 1.6666654611e-1f) * z * r + r;
	const float pc = ((2.443315711809948e-5f * z - 1.388731625493765e-3f) * z + 4.166664568298827e-2f) * z * z - 0.5f * z + 1.0f;
	const int q = (int)k & 3;
	*s = q == 0 ? ps : q == 1 ? pc : q == 2 ? -ps : -pc;
	*c = q == 0 ? pc : q == 1 ? -ps : q == 2 ? -pc : ps;
}
__device__ __forceinline__ float taa_sin(float x) { float s, c; taa_sincos(x, &s, &c); return s; }
__device__ __forceinline__ float taa_cos(float x) { float s, c; taa_sincos(x, &s, &c); return c; }

// ---- colour space (taa.comp:158-197) ------------------------------------------------------------
__device__ __forceinline__ f3 rgb_to_ycocg(f3 c) {
	return mk3(.25f * c.x + .5f * c.y + .25f * c.z, .5f * c.x - .5f * c.z, -.25f * c.x + .5f * c.y - .25f * c.z);
}
__device__ __forceinline__ f3 ycocg_to_rgb(f3 c) {
	float tmp = c.x - c.z;
	return mk3(tmp + c.y, c.x + c.z, tmp - c.y);
}
__device__ __forceinline__ f3 tonemap_karis(f3 hdr) {
	float luma = fmaxf(fmaxf(hdr.x, hdr.y), hdr.z);
	return hdr / (1.0f + luma);
}
__device__ __forceinline__ f3 un_tonemap_karis(f3 ldr) {
	float luma = fmaxf(fmaxf(ldr.x, ldr.y), ldr.z);
	return ldr / (1.0f - luma);
}

}  // namespace taa
