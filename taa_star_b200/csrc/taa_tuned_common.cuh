// taa_tuned_common.cuh — building blocks shared by the tuned resolve kernels (taa_resolve_stream.cu: one warp per 62-column strip, rows
// streamed through TMA; taa_resolve_strip.cu: 64-wide shared-memory tiles, long strips). Arithmetic contract: DESIGN.md "Why the tuned kernels
// are exact where it matters".
#pragma once
#include "taa_device.cuh"
#include <cmath>

namespace taa {
namespace tuned {

// |tuned - exact| of the clip distance is dominated by the two outer Catmull-Rom taps, whose sampler bleed (<= 0.074 * eps * texel
// contrast, eps <= ~6e-4 texel at x ~ 3840) the 4x4 footprint cannot hold: <= ~2e-5 on full-contrast edges at 4K. The band is twice
// that and grows with the frame size like eps does.
constexpr float FIXUP_BAND_4K = 4.0e-5f;

// sampler footprints that depend on the column or on the row only, as byte offsets into the image
struct ColX { unsigned int m, n; float p; };                 // colour tap: main texel, bleeding neighbour (byte offsets in a row), its weight
struct ColY { unsigned int m, n; float p; };  // same for rows (byte offsets of the rows in the buffer)
struct VelX { unsigned int o0, o1; float a, u; };            // velocity tap: exact bilinear footprint and the pixel's u (taa.comp:131)
struct VelY { unsigned int o0, o1; float a, v; };

__device__ __forceinline__ float sat(float x) { return __saturatef(x); }
__device__ __forceinline__ float rcp_approx(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float sqrt_approx(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ __half2 h2(unsigned int v) { return *reinterpret_cast<__half2*>(&v); }

// byte offset of global row gy in a (band) buffer; rows the buffer does not hold are clamped into it and reported.
// 32-bit: tuned_supports() admits only buffers smaller than 4 GB, so that addresses are uniform base + 32-bit offset.
template <class I>
__device__ __forceinline__ unsigned int row_off(const I& im, int gy, unsigned int* status) {
	int ly = gy - im.y0;
	if ((unsigned int)ly >= (unsigned int)im.rows) {
		if (status) atomicOr(status, 1u);
		ly = ly < 0 ? 0 : im.rows - 1;
	}
	return (unsigned int)ly * (unsigned int)im.pitch;
}

// taa.comp:207: offset + vec2(iuv + d + 0.5) * invsize with offset = 0, through the sampler (exact coordinate arithmetic)
__device__ __forceinline__ void colour_axis(int g, float inv, int size, int& m, int& n, float& p) {
	Lin L = lin_coord(((float)g + 0.5f) * inv, size);
	if (L.a < 0.5f) { m = L.i0; n = L.i1; p = L.a; } else { m = L.i1; n = L.i0; p = 1.0f - L.a; }
}

// One axis of sample_history_bicubic_catmullrom (taa.comp:441-514) as weights on the 4 texels k-1..k+2.
// The centre tap's sampler position is evaluated exactly as the shader + sampler do, ((tc + w2/wC) * invTexSize) * size - 0.5,
// and spread with the bilinear tent, so its rounding bleed lands on the right texel. The two outer taps
// (|w| <= 0.074) are taken at their texel centres.
struct AxisW { int k; float w[4]; };
__device__ __forceinline__ AxisW catmull_axis(float h, float size, float inv) {
	const float it = h * size;                        // iTc = uv * texSize
	const float kf = floorf(it - 0.5f);
	const float tc = kf + 0.5f;                       // round down to the nearest texel centre
	const float f = it - tc;
	const float f2 = f * f;
	const float w0 = f * fmaf(f, fmaf(-0.5f, f, 1.0f), -0.5f);
	const float w1 = fmaf(f2, fmaf(1.5f, f, -2.5f), 1.0f);
	const float w2 = f * fmaf(f, fmaf(-1.5f, f, 2.0f), 0.5f);
	const float w3 = f2 * fmaf(0.5f, f, -0.5f);
	const float wC = w1 + w2;
	const float r = w2 * rcp_approx(wC);
	const float uC = ((tc + r) * inv) * size - 0.5f;  // the sampler's unnormalised coordinate of the centre tap
	const float sC = uC - kf;                         // in [0,1] up to the coordinate rounding
	const float d1 = sC - 1.0f;
	AxisW o;
	o.w[0] = fmaf(wC, sat(-sC), w0);
	o.w[1] = wC * sat(1.0f - fabsf(sC));
	o.w[2] = wC * sat(1.0f - fabsf(d1));
	o.w[3] = fmaf(wC, sat(d1), w3);
	o.k = (int)fminf(fmaxf(kf, -8.0f), size + 8.0f);  // NaN -> 0; keeps k + 2 far from integer overflow
	return o;
}

struct Hist { float r, g, b, a; unsigned int abits; };

// the 16 texels of the 4x4 footprint. INTERIOR: no texel is clamped and all rows are in the buffer -> one base pointer, immediate offsets.
template <bool INTERIOR>
__device__ __forceinline__ void load_history(const Img& him, int kx, int ky, int W, int H, unsigned int* st, uint2 (&q)[16]) {
	if (INTERIOR) {
		const unsigned int pitch = (unsigned int)him.pitch;
		const unsigned int base = (unsigned int)(ky - 1 - him.y0) * pitch + (unsigned int)(kx - 1) * 8u;
#pragma unroll
		for (int i = 0; i < 4; ++i) {
			const uint2* hp = reinterpret_cast<const uint2*>(him.p + (base + i * pitch));
			q[4 * i] = __ldg(hp); q[4 * i + 1] = __ldg(hp + 1); q[4 * i + 2] = __ldg(hp + 2); q[4 * i + 3] = __ldg(hp + 3);
		}
	} else {
		const int c0 = iclamp(kx - 1, 0, W - 1), c1 = iclamp(kx, 0, W - 1), c2 = iclamp(kx + 1, 0, W - 1), c3 = iclamp(kx + 2, 0, W - 1);
#pragma unroll
		for (int i = 0; i < 4; ++i) {
			const uint2* hp = reinterpret_cast<const uint2*>(him.p + (size_t)row_off(him, iclamp(ky - 1 + i, 0, H - 1), st));
			q[4 * i] = __ldg(hp + c0); q[4 * i + 1] = __ldg(hp + c1); q[4 * i + 2] = __ldg(hp + c2); q[4 * i + 3] = __ldg(hp + c3);
		}
	}
}

// one history row of the footprint filtered horizontally (4 texels, weights of the x axis)
struct HRow { float r, g, b, a; unsigned int abits; };
template <bool REJ>
__device__ __forceinline__ HRow hfilter(const uint2 q0, const uint2 q1, const uint2 q2, const uint2 q3, const float (&w)[4]) {
	HRow o = {0.f, 0.f, 0.f, 0.f, 0u};
	const float2 e0 = __half22float2(h2(q0.x)), e1 = __half22float2(h2(q1.x)), e2 = __half22float2(h2(q2.x)), e3 = __half22float2(h2(q3.x));
	o.r = fmaf(w[3], e3.x, fmaf(w[2], e2.x, fmaf(w[1], e1.x, w[0] * e0.x)));
	o.g = fmaf(w[3], e3.y, fmaf(w[2], e2.y, fmaf(w[1], e1.y, w[0] * e0.y)));
	if (REJ) {
		const float2 g0 = __half22float2(h2(q0.y)), g1 = __half22float2(h2(q1.y)), g2 = __half22float2(h2(q2.y)), g3 = __half22float2(h2(q3.y));
		o.b = fmaf(w[3], g3.x, fmaf(w[2], g2.x, fmaf(w[1], g1.x, w[0] * g0.x)));
		o.a = fmaf(w[3], g3.y, fmaf(w[2], g2.y, fmaf(w[1], g1.y, w[0] * g0.y)));
		o.abits = (q0.y | q1.y) | (q2.y | q3.y);
	} else {
		const float g0 = __low2float(h2(q0.y)), g1 = __low2float(h2(q1.y)), g2 = __low2float(h2(q2.y)), g3 = __low2float(h2(q3.y));
		o.b = fmaf(w[3], g3, fmaf(w[2], g2, fmaf(w[1], g1, w[0] * g0)));
	}
	return o;
}

// d = a * b + c with a, b fp16 and c, d fp32 (FHFMA): the product of two halves is exact in fp32
__device__ __forceinline__ float fhfma(__half a, __half b, float c) {
	float r;
	asm("fma.rn.f32.f16 %0, %1, %2, %3;" : "=f"(r) : "h"(__half_as_ushort(a)), "h"(__half_as_ushort(b)), "f"(c));
	return r;
}

// One sampled colour tap (taa.comp:207 through the sampler) converted to YCoCg: T[m] + px (T[nx] - T[m]) + py (T[ny] - T[m]).
// px, py are ~1e-4 (the sampler coordinate lands that far beside the texel centre); the px*py cross term (< 1e-6) is dropped.
__device__ __forceinline__ float3 sample_ycocg(const uint2 M, const uint2 NX, const uint2 NY, const __half px, const __half py) {
	const __half2 m01 = h2(M.x), m23 = h2(M.y);
	const __half2 dx01 = __hsub2(h2(NX.x), m01), dx23 = __hsub2(h2(NX.y), m23);
	const __half2 dy01 = __hsub2(h2(NY.x), m01), dy23 = __hsub2(h2(NY.y), m23);
	const float cr = fhfma(__low2half(dx01), px, fhfma(__low2half(dy01), py, __low2float(m01)));
	const float cg = fhfma(__high2half(dx01), px, fhfma(__high2half(dy01), py, __high2float(m01)));
	const float cb = fhfma(__low2half(dx23), px, fhfma(__low2half(dy23), py, __low2float(m23)));
	const float t = cr + cb, hg = 0.5f * cg;
	return make_float3(fmaf(0.25f, t, hg), 0.5f * (cr - cb), fmaf(-0.25f, t, hg));
}

// both halves of a packed pair are finite (exponent field != 31)
__device__ __forceinline__ bool finite2(unsigned int v) { return (((v & 0x7c007c00u) + 0x04000400u) & 0x80008000u) == 0u; }

__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned int)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(void* dst, const void* src) {
	asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned int)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}

// ---- bulk (TMA) staging: one thread copies one whole tile row (544 bytes) global -> shared, completion counted on an mbarrier ----
__device__ __forceinline__ unsigned int smem_u32(const void* p) { return (unsigned int)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned int count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned int bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned int parity) {
	asm volatile(
	    "{\n"
	    ".reg .pred p;\n"
	    "WAIT_%=:\n"
	    "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
	    "@p bra DONE_%=;\n"
	    "bra WAIT_%=;\n"
	    "DONE_%=:\n"
	    "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_row(void* dst, const void* src, unsigned int bytes, unsigned long long* bar) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes),
	             "r"(smem_u32(bar)) : "memory");
}

}  // namespace tuned
}  // namespace taa
