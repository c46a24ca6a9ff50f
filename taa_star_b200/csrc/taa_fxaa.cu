// taa_fxaa.cu — the FXAA branch of taa<CF>::render() (source/taa.hpp:1061-1107): antialias_fxaa_prepare.comp (luma into alpha)
// and antialias_fxaa.comp (FxaaPixelShader of shaders/Fxaa3_11_mod.h:884-1243, FXAA_PC, quality preset 12, on the pixels whose
// segmentation mask says 1; every other pixel is copied).
//
// Exact arithmetic like the other follow-on passes (this file is compiled with --fmad=false): one fp32 operation per GLSL
// operation in source order, the bilinear sampler of taa_device.cuh, and the FXAA_GATHER4_ALPHA == 1 texel access that a GLSL front
// end predefining GL_ARB_gpu_shader5 (glslang) compiles (Fxaa3_11_mod.h:309-326, 888-912): the 2x2 gather footprint is chosen after
// snapping the unnormalised coordinate to 1/256 texel (subTexelPrecisionBits = 8), like oracle/taa_oracle.cpp.
//
// PREPARED = false fuses the two dispatches: the luma the prepare pass would have stored in alpha is computed where it is read,
// rounded to fp16 as that store rounds it, so the result is bit-identical and the intermediate image is never written.
#include "taa_device.cuh"
#include "taa_kernels.h"

namespace taa {

namespace {

// antialias_fxaa_prepare.comp:22-24
__device__ __forceinline__ float fxaa_luma(float r, float g, float b) { return r * 0.299f + g * 0.587f + b * 0.114f; }

// one texel of the image FXAA samples: (rgb, luma)
template <bool PREPARED>
__device__ __forceinline__ float4 fx_texel(const Img& im, int x, int y) {
	float4 t = unpack_rgba16f(__ldg(reinterpret_cast<const uint2*>(im.p + (long long)(y - im.y0) * im.pitch) + x));
	if (!PREPARED) t.w = __half2float(__float2half_rn(fxaa_luma(t.x, t.y, t.z)));
	return t;
}

struct Foot { int x0, x1, y0, y1; float a, b; };
// bilinear footprint of uv, moved by (ox, oy) texels before clamp-to-edge (textureLodOffset)
__device__ __forceinline__ Foot foot_linear(float s, float t, int w, int h, int ox, int oy) {
	const float u = s * (float)w - 0.5f, v = t * (float)h - 0.5f;
	const float fu = floorf(u), fv = floorf(v);
	Foot f;
	f.a = u - fu; f.b = v - fv;
	const int i0 = (int)fu + ox, j0 = (int)fv + oy;  // |coordinates| stay far below 2^31: pos is within a few texels of the image
	f.x0 = iclamp(i0, 0, w - 1); f.x1 = iclamp(i0 + 1, 0, w - 1);
	f.y0 = iclamp(j0, 0, h - 1); f.y1 = iclamp(j0 + 1, 0, h - 1);
	return f;
}
__device__ __forceinline__ Foot foot_gather(float s, float t, int w, int h, int ox, int oy) {
	const float u = s * (float)w - 0.5f, v = t * (float)h - 0.5f;
	const float fu = floorf(floorf(u * 256.0f + 0.5f) * (1.0f / 256.0f)), fv = floorf(floorf(v * 256.0f + 0.5f) * (1.0f / 256.0f));
	Foot f;
	f.a = 0.f; f.b = 0.f;
	const int i0 = (int)fu + ox, j0 = (int)fv + oy;
	f.x0 = iclamp(i0, 0, w - 1); f.x1 = iclamp(i0 + 1, 0, w - 1);
	f.y0 = iclamp(j0, 0, h - 1); f.y1 = iclamp(j0 + 1, 0, h - 1);
	return f;
}
template <bool PREPARED>
__device__ __forceinline__ float4 fx_sample(const Img& im, int w, int h, float s, float t, int ox = 0, int oy = 0) {
	const Foot f = foot_linear(s, t, w, h, ox, oy);
	return lerp4(lerp4(fx_texel<PREPARED>(im, f.x0, f.y0), fx_texel<PREPARED>(im, f.x1, f.y0), f.a),
	             lerp4(fx_texel<PREPARED>(im, f.x0, f.y1), fx_texel<PREPARED>(im, f.x1, f.y1), f.a), f.b);
}
// only the luma of a bilinear sample (FxaaLuma(FxaaTexTop(..)), Fxaa3_11_mod.h:710)
template <bool PREPARED>
__device__ __forceinline__ float fx_sample_luma(const Img& im, int w, int h, float s, float t, int ox = 0, int oy = 0) {
	const Foot f = foot_linear(s, t, w, h, ox, oy);
	return lerpf(lerpf(fx_texel<PREPARED>(im, f.x0, f.y0).w, fx_texel<PREPARED>(im, f.x1, f.y0).w, f.a),
	             lerpf(fx_texel<PREPARED>(im, f.x0, f.y1).w, fx_texel<PREPARED>(im, f.x1, f.y1).w, f.a), f.b);
}

template <bool PREPARED>
__device__ float4 fxaa_pixel(const Img& im, int w, int h, float px, float py, const TaaFxaaPush& pc) {
	const float rx = pc.fxaaQualityRcpFrame[0], ry = pc.fxaaQualityRcpFrame[1];
	float pmx = px, pmy = py;
	const float4 rgbyM = fx_sample<PREPARED>(im, w, h, pmx, pmy);
	const float lumaM = rgbyM.w;
	float lumaS, lumaE, lumaN, lumaW, lumaNW, lumaSE;
	{
		const Foot A = foot_gather(pmx, pmy, w, h, 0, 0), B = foot_gather(pmx, pmy, w, h, -1, -1);
		lumaE = fx_texel<PREPARED>(im, A.x1, A.y0).w;   // luma4A.z
		lumaS = fx_texel<PREPARED>(im, A.x0, A.y1).w;   // luma4A.x
		lumaSE = fx_texel<PREPARED>(im, A.x1, A.y1).w;  // luma4A.y
		lumaNW = fx_texel<PREPARED>(im, B.x0, B.y0).w;  // luma4B.w
		lumaN = fx_texel<PREPARED>(im, B.x1, B.y0).w;   // luma4B.z
		lumaW = fx_texel<PREPARED>(im, B.x0, B.y1).w;   // luma4B.x
	}
	const float maxSM = fmaxf(lumaS, lumaM), minSM = fminf(lumaS, lumaM);
	const float maxESM = fmaxf(lumaE, maxSM), minESM = fminf(lumaE, minSM);
	const float maxWN = fmaxf(lumaN, lumaW), minWN = fminf(lumaN, lumaW);
	const float rangeMax = fmaxf(maxWN, maxESM), rangeMin = fminf(minWN, minESM);
	const float rangeMaxScaled = rangeMax * pc.fxaaQualityEdgeThreshold;
	const float range = rangeMax - rangeMin;
	const float rangeMaxClamped = fmaxf(pc.fxaaQualityEdgeThresholdMin, rangeMaxScaled);
	if (range < rangeMaxClamped) return rgbyM;
	const float lumaNE = fx_sample_luma<PREPARED>(im, w, h, pmx, pmy, 1, -1);
	const float lumaSW = fx_sample_luma<PREPARED>(im, w, h, pmx, pmy, -1, 1);
	const float lumaNS = lumaN + lumaS, lumaWE = lumaW + lumaE;
	const float subpixRcpRange = 1.0f / range;
	const float subpixNSWE = lumaNS + lumaWE;
	const float edgeHorz1 = (-2.0f * lumaM) + lumaNS, edgeVert1 = (-2.0f * lumaM) + lumaWE;
	const float lumaNESE = lumaNE + lumaSE, lumaNWNE = lumaNW + lumaNE;
	const float edgeHorz2 = (-2.0f * lumaE) + lumaNESE, edgeVert2 = (-2.0f * lumaN) + lumaNWNE;
	const float lumaNWSW = lumaNW + lumaSW, lumaSWSE = lumaSW + lumaSE;
	const float edgeHorz4 = (fabsf(edgeHorz1) * 2.0f) + fabsf(edgeHorz2), edgeVert4 = (fabsf(edgeVert1) * 2.0f) + fabsf(edgeVert2);
	const float edgeHorz3 = (-2.0f * lumaW) + lumaNWSW, edgeVert3 = (-2.0f * lumaS) + lumaSWSE;
	const float edgeHorz = fabsf(edgeHorz3) + edgeHorz4, edgeVert = fabsf(edgeVert3) + edgeVert4;
	const float subpixNWSWNESE = lumaNWSW + lumaNESE;
	float lengthSign = rx;
	const bool horzSpan = edgeHorz >= edgeVert;
	const float subpixA = subpixNSWE * 2.0f + subpixNWSWNESE;
	if (!horzSpan) { lumaN = lumaW; lumaS = lumaE; }
	if (horzSpan) lengthSign = ry;
	const float subpixB = (subpixA * (1.0f / 12.0f)) - lumaM;
	const float gradientN = lumaN - lumaM, gradientS = lumaS - lumaM;
	float lumaNN = lumaN + lumaM;
	const float lumaSS = lumaS + lumaM;
	const bool pairN = fabsf(gradientN) >= fabsf(gradientS);
	const float gradient = fmaxf(fabsf(gradientN), fabsf(gradientS));
	if (pairN) lengthSign = -lengthSign;
	const float subpixC = clampf(fabsf(subpixB) * subpixRcpRange, 0.0f, 1.0f);
	float pbx = pmx, pby = pmy;
	const float offx = (!horzSpan) ? 0.0f : rx, offy = horzSpan ? 0.0f : ry;
	if (!horzSpan) pbx += lengthSign * 0.5f;
	if (horzSpan) pby += lengthSign * 0.5f;
	const float P[5] = {1.0f, 1.5f, 2.0f, 4.0f, 12.0f};  // FXAA_QUALITY_P0..P4 of preset 12 (Fxaa3_11_mod.h:433-440)
	float pnx = pbx - offx * P[0], pny = pby - offy * P[0];
	float ppx = pbx + offx * P[0], ppy = pby + offy * P[0];
	const float subpixD = ((-2.0f) * subpixC) + 3.0f;
	float lumaEndN = fx_sample_luma<PREPARED>(im, w, h, pnx, pny);
	const float subpixE = subpixC * subpixC;
	float lumaEndP = fx_sample_luma<PREPARED>(im, w, h, ppx, ppy);
	if (!pairN) lumaNN = lumaSS;
	const float gradientScaled = gradient * 1.0f / 4.0f;
	const float lumaMM = lumaM - lumaNN * 0.5f;
	const float subpixF = subpixD * subpixE;
	const bool lumaMLTZero = lumaMM < 0.0f;
	lumaEndN -= lumaNN * 0.5f;
	lumaEndP -= lumaNN * 0.5f;
	bool doneN = fabsf(lumaEndN) >= gradientScaled, doneP = fabsf(lumaEndP) >= gradientScaled;
	// the nested `if(doneNP)` blocks of Fxaa3_11_mod.h:1030-1190 for FXAA_QUALITY_PS == 5
#pragma unroll 1
	for (int i = 1;; ++i) {
		if (!doneN) { pnx -= offx * P[i]; pny -= offy * P[i]; }
		const bool doneNP = (!doneN) || (!doneP);
		if (!doneP) { ppx += offx * P[i]; ppy += offy * P[i]; }
		if (!doneNP || i == 4) break;
		if (!doneN) lumaEndN = fx_sample_luma<PREPARED>(im, w, h, pnx, pny);
		if (!doneP) lumaEndP = fx_sample_luma<PREPARED>(im, w, h, ppx, ppy);
		if (!doneN) lumaEndN = lumaEndN - lumaNN * 0.5f;
		if (!doneP) lumaEndP = lumaEndP - lumaNN * 0.5f;
		doneN = fabsf(lumaEndN) >= gradientScaled;
		doneP = fabsf(lumaEndP) >= gradientScaled;
	}
	float dstN = pmx - pnx, dstP = ppx - pmx;
	if (!horzSpan) { dstN = pmy - pny; dstP = ppy - pmy; }
	const bool goodSpanN = (lumaEndN < 0.0f) != lumaMLTZero;
	const float spanLength = dstP + dstN;
	const bool goodSpanP = (lumaEndP < 0.0f) != lumaMLTZero;
	const float spanLengthRcp = 1.0f / spanLength;
	const bool directionN = dstN < dstP;
	const float dst = fminf(dstN, dstP);
	const bool goodSpan = directionN ? goodSpanN : goodSpanP;
	const float subpixG = subpixF * subpixF;
	const float pixelOffset = (dst * (-spanLengthRcp)) + 0.5f;
	const float subpixH = subpixG * pc.fxaaQualitySubpix;
	const float pixelOffsetGood = goodSpan ? pixelOffset : 0.0f;
	const float pixelOffsetSubpix = fmaxf(pixelOffsetGood, subpixH);
	if (!horzSpan) pmx += pixelOffsetSubpix * lengthSign;
	if (horzSpan) pmy += pixelOffsetSubpix * lengthSign;
	const float4 o = fx_sample<PREPARED>(im, w, h, pmx, pmy);
	return make_float4(o.x, o.y, o.z, lumaM);
}

__global__ void __launch_bounds__(256) fxaa_prepare_kernel(const __grid_constant__ PostImg io) {
	const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
	if (x >= io.w || y >= io.h) return;
	float4 c = fx_texel<true>(io.src, x, y);
	c.w = fxaa_luma(c.x, c.y, c.z);
	st_rgba16f(io.dst, x, y, c);
}

// antialias_fxaa.comp:30-64. io.debug carries the segmentation mask (r32ui) here.
template <bool PREPARED>
__global__ void __launch_bounds__(256) fxaa_kernel(const __grid_constant__ PostImg io, const __grid_constant__ TaaFxaaPush pc) {
	const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
	if (x >= io.w || y >= io.h) return;
	const unsigned int seg = __ldg(reinterpret_cast<const unsigned int*>(io.debug.p + (long long)(y - io.debug.y0) * io.debug.pitch) + x) & 3u;
	float4 color;
	if (seg == 1u) color = fxaa_pixel<PREPARED>(io.src, io.w, io.h, ((float)x + 0.5f) * pc.fxaaQualityRcpFrame[0], ((float)y + 0.5f) * pc.fxaaQualityRcpFrame[1], pc);
	else color = fx_texel<PREPARED>(io.src, x, y);
	st_rgba16f(io.dst, x, y, color);
}

inline dim3 grid2d(int w, int h, dim3 b) { return dim3((w + b.x - 1) / b.x, (h + b.y - 1) / b.y); }

}  // namespace

cudaError_t launch_fxaa_prepare(const PostImg& io, cudaStream_t stream) {
	dim3 b(32, 8);
	fxaa_prepare_kernel<<<grid2d(io.w, io.h, b), b, 0, stream>>>(io);
	return cudaGetLastError();
}
cudaError_t launch_fxaa(const PostImg& io, const TaaFxaaPush& pc, bool prepared, cudaStream_t stream) {
	dim3 b(32, 8);
	if (prepared) fxaa_kernel<true><<<grid2d(io.w, io.h, b), b, 0, stream>>>(io, pc);
	else fxaa_kernel<false><<<grid2d(io.w, io.h, b), b, 0, stream>>>(io, pc);
	return cudaGetLastError();
}

}  // namespace taa
