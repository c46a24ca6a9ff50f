// taa_fxaa.cu — the FXAA branch of taa<CF>::render() (source/taa.hpp:1061-1107): antialias_fxaa_prepare.comp (luma into alpha) and
// antialias_fxaa.comp (the pixels whose segmentation mask says 1 are filtered, every other pixel is copied).
//
// The kernel is this repository's own: a CTA owns a 32 x 16 tile; pixels the mask does not select are copied and done; if the tile has
// any selected pixel, the luma of the tile and a two-texel apron is staged in shared memory once (computed on the fly from rgb, rounded
// to fp16 as the prepare pass would have stored it, unless the source is already prepared), the local-contrast gate runs out of that
// tile, and only the pixels that pass the gate are COMPACTED into a work list that consecutive threads then walk along their edges —
// the expensive part (up to ten bilinear luma taps per pixel) runs without idle lanes.
//
// What is NOT ours is the arithmetic of the filter itself — contrast gate, edge orientation, end-of-edge search with the step table of
// quality preset 12, sub-pixel blend: it is FXAA 3.11 (PC quality path) as vendored by the reference in shaders/Fxaa3_11_mod.h:884-1243,
// evaluated here in the order that header writes it so that the result is bit-identical to the reference's shader (tests/test_fxaa_gpu.py;
// one fp32 operation per written operation, this file is compiled with --fmad=false). That part is derived from:
//
//     NVIDIA FXAA 3.11 by TIMOTHY LOTTES
//     COPYRIGHT (C) 2010, 2011 NVIDIA CORPORATION. ALL RIGHTS RESERVED.
//     TO THE MAXIMUM EXTENT PERMITTED BY APPLICABLE LAW, THIS SOFTWARE IS PROVIDED *AS IS* AND NVIDIA AND ITS SUPPLIERS DISCLAIM ALL
//     WARRANTIES, EITHER EXPRESS OR IMPLIED, INCLUDING, BUT NOT LIMITED TO, IMPLIED WARRANTIES OF MERCHANTABILITY AND FITNESS FOR A
//     PARTICULAR PURPOSE. IN NO EVENT SHALL NVIDIA OR ITS SUPPLIERS BE LIABLE FOR ANY SPECIAL, INCIDENTAL, INDIRECT, OR CONSEQUENTIAL
//     DAMAGES WHATSOEVER (INCLUDING, WITHOUT LIMITATION, DAMAGES FOR LOSS OF BUSINESS PROFITS, BUSINESS INTERRUPTION, LOSS OF BUSINESS
//     INFORMATION, OR ANY OTHER PECUNIARY LOSS) ARISING OUT OF THE USE OF OR INABILITY TO USE THIS SOFTWARE, EVEN IF NVIDIA HAS BEEN
//     ADVISED OF THE POSSIBILITY OF SUCH DAMAGES.
//
// Texel access follows the FXAA_GATHER4_ALPHA == 1 path that a GLSL front end predefining GL_ARB_gpu_shader5 compiles (Fxaa3_11_mod.h:309-326,
// 888-912): the 2 x 2 gather footprint is chosen after snapping the unnormalised coordinate to 1/256 texel (subTexelPrecisionBits = 8), like
// oracle/taa_oracle.cpp; everything else goes through the bilinear clamp-to-edge sampler of taa.hpp:274.
#include "taa_device.cuh"
#include "taa_kernels.h"

namespace taa {

namespace {

constexpr int FTW = 32, FTH = 16;   // pixels per CTA
constexpr int FAP = 2;              // apron of the luma tile: the corner taps of the gate reach two texels out
constexpr int FLW = FTW + 2 * FAP, FLH = FTH + 2 * FAP;
constexpr int FNT = 256;            // threads per CTA: 32 x 8, two tile rows each

// antialias_fxaa_prepare.comp:22-24
__device__ __forceinline__ float luma_of(float r, float g, float b) { return r * 0.299f + g * 0.587f + b * 0.114f; }

// one texel of the image the filter samples: (rgb, luma)
template <bool PREPARED>
__device__ __forceinline__ float4 src_texel(const Img& im, int x, int y) {
	float4 t = unpack_rgba16f(__ldg(reinterpret_cast<const uint2*>(im.p + (long long)(y - im.y0) * im.pitch) + x));
	if (!PREPARED) t.w = __half2float(__float2half_rn(luma_of(t.x, t.y, t.z)));
	return t;
}

// 2 x 2 texel footprint of a sampler access at normalised (s, t), moved by (ox, oy) texels before clamp-to-edge
struct Quad { int x0, x1, y0, y1; float fx, fy; };
__device__ __forceinline__ Quad quad_bilinear(float s, float t, int w, int h, int ox, int oy) {
	const float u = s * (float)w - 0.5f, v = t * (float)h - 0.5f;
	const float fu = floorf(u), fv = floorf(v);
	Quad q;
	q.fx = u - fu; q.fy = v - fv;
	const int i0 = (int)fu + ox, j0 = (int)fv + oy;  // (positions stay within a few texels of the image)
	q.x0 = iclamp(i0, 0, w - 1); q.x1 = iclamp(i0 + 1, 0, w - 1);
	q.y0 = iclamp(j0, 0, h - 1); q.y1 = iclamp(j0 + 1, 0, h - 1);
	return q;
}
__device__ __forceinline__ Quad quad_gather(float s, float t, int w, int h, int ox, int oy) {
	const float u = s * (float)w - 0.5f, v = t * (float)h - 0.5f;
	const float fu = floorf(floorf(u * 256.0f + 0.5f) * (1.0f / 256.0f)), fv = floorf(floorf(v * 256.0f + 0.5f) * (1.0f / 256.0f));
	Quad q;
	q.fx = 0.f; q.fy = 0.f;
	const int i0 = (int)fu + ox, j0 = (int)fv + oy;
	q.x0 = iclamp(i0, 0, w - 1); q.x1 = iclamp(i0 + 1, 0, w - 1);
	q.y0 = iclamp(j0, 0, h - 1); q.y1 = iclamp(j0 + 1, 0, h - 1);
	return q;
}

// the CTA's luma tile: entry (j, i) holds the luma of texel (clamp(x0 - FAP + i), clamp(y0 - FAP + j))
struct LumaTile {
	float v[FLH][FLW + 1];
	int x0, y0;  // image coordinates of entry (FAP, FAP)
	__device__ __forceinline__ float at(int gx, int gy) const { return v[gy - y0 + FAP][gx - x0 + FAP]; }  // (gx, gy): clamped image coordinates inside the tile's reach
	__device__ __forceinline__ float bilinear(const Quad& q) const {
		return lerpf(lerpf(at(q.x0, q.y0), at(q.x1, q.y0), q.fx), lerpf(at(q.x0, q.y1), at(q.x1, q.y1), q.fx), q.fy);
	}
};

// luma of a bilinear sample anywhere in the image (the end-of-edge search leaves the tile)
template <bool PREPARED>
__device__ __forceinline__ float luma_tap(const Img& im, int w, int h, float s, float t) {
	const Quad q = quad_bilinear(s, t, w, h, 0, 0);
	return lerpf(lerpf(src_texel<PREPARED>(im, q.x0, q.y0).w, src_texel<PREPARED>(im, q.x1, q.y0).w, q.fx),
	             lerpf(src_texel<PREPARED>(im, q.x0, q.y1).w, src_texel<PREPARED>(im, q.x1, q.y1).w, q.fx), q.fy);
}
template <bool PREPARED>
__device__ __forceinline__ float4 colour_tap(const Img& im, int w, int h, float s, float t) {
	const Quad q = quad_bilinear(s, t, w, h, 0, 0);
	return lerp4(lerp4(src_texel<PREPARED>(im, q.x0, q.y0), src_texel<PREPARED>(im, q.x1, q.y0), q.fx),
	             lerp4(src_texel<PREPARED>(im, q.x0, q.y1), src_texel<PREPARED>(im, q.x1, q.y1), q.fx), q.fy);
}

// the lumas the gate and the orientation test look at: centre sample, the plus (gathered texels), the corners (NE / SW are offset samples)
struct Around { float m, n, s, w, e, nw, ne, sw, se; };

__device__ __forceinline__ Around lumas_around(const LumaTile& T, float s, float t, int w, int h) {
	Around a;
	a.m = T.bilinear(quad_bilinear(s, t, w, h, 0, 0));
	const Quad lo = quad_gather(s, t, w, h, 0, 0), hi = quad_gather(s, t, w, h, -1, -1);
	a.e = T.at(lo.x1, lo.y0); a.s = T.at(lo.x0, lo.y1); a.se = T.at(lo.x1, lo.y1);
	a.nw = T.at(hi.x0, hi.y0); a.n = T.at(hi.x1, hi.y0); a.w = T.at(hi.x0, hi.y1);
	a.ne = T.bilinear(quad_bilinear(s, t, w, h, 1, -1));
	a.sw = T.bilinear(quad_bilinear(s, t, w, h, -1, 1));
	return a;
}

// local contrast below the thresholds: the pixel keeps its centre sample. Returns the contrast range.
__device__ __forceinline__ bool gate_closed(const Around& a, const TaaFxaaPush& pc, float& range) {
	const float hiSM = fmaxf(a.s, a.m), loSM = fminf(a.s, a.m);
	const float hiESM = fmaxf(a.e, hiSM), loESM = fminf(a.e, loSM);
	const float hiWN = fmaxf(a.n, a.w), loWN = fminf(a.n, a.w);
	const float hi = fmaxf(hiWN, hiESM), lo = fminf(loWN, loESM);
	const float scaled = hi * pc.fxaaQualityEdgeThreshold;
	range = hi - lo;
	return range < fmaxf(pc.fxaaQualityEdgeThresholdMin, scaled);
}

// A pixel that passed the gate: orientation of the edge, search for its two ends, sub-pixel shift, final tap.
template <bool PREPARED>
__device__ float4 walk_edge(const Img& im, int w, int h, const float cs, const float ct, Around a, const float range, const TaaFxaaPush& pc) {
	const float texel_x = pc.fxaaQualityRcpFrame[0], texel_y = pc.fxaaQualityRcpFrame[1];
	// ---- is the edge horizontal or vertical? second differences along rows and columns, the centre line weighted twice ----
	const float sumNS = a.n + a.s, sumWE = a.w + a.e;
	const float inv_range = 1.0f / range;
	const float sumPlus = sumNS + sumWE;
	const float rowM = (-2.0f * a.m) + sumNS, colM = (-2.0f * a.m) + sumWE;
	const float sumEastCorners = a.ne + a.se, sumNorthCorners = a.nw + a.ne;
	const float rowE = (-2.0f * a.e) + sumEastCorners, colN = (-2.0f * a.n) + sumNorthCorners;
	const float sumWestCorners = a.nw + a.sw, sumSouthCorners = a.sw + a.se;
	const float rowME = (fabsf(rowM) * 2.0f) + fabsf(rowE), colMN = (fabsf(colM) * 2.0f) + fabsf(colN);
	const float rowW = (-2.0f * a.w) + sumWestCorners, colS = (-2.0f * a.s) + sumSouthCorners;
	const float horizontalness = fabsf(rowW) + rowME, verticalness = fabsf(colS) + colMN;
	const float sumCorners = sumWestCorners + sumEastCorners;
	const bool horizontal = horizontalness >= verticalness;
	const float lowpass = sumPlus * 2.0f + sumCorners;
	// the two neighbours across the edge, and one texel across it
	float across = texel_x;
	if (!horizontal) { a.n = a.w; a.s = a.e; }
	if (horizontal) across = texel_y;
	const float contrast_lp = (lowpass * (1.0f / 12.0f)) - a.m;
	const float gradA = a.n - a.m, gradB = a.s - a.m;
	float pairLuma = a.n + a.m;
	const float otherPair = a.s + a.m;
	const bool steeperA = fabsf(gradA) >= fabsf(gradB);
	const float gradient = fmaxf(fabsf(gradA), fabsf(gradB));
	if (steeperA) across = -across;
	const float subpix0 = clampf(fabsf(contrast_lp) * inv_range, 0.0f, 1.0f);
	// ---- start half a texel across the edge, step along it in both directions ----
	float bx = cs, by = ct;
	const float along_x = (!horizontal) ? 0.0f : texel_x, along_y = horizontal ? 0.0f : texel_y;
	if (!horizontal) bx += across * 0.5f;
	if (horizontal) by += across * 0.5f;
	const float STEP[5] = {1.0f, 1.5f, 2.0f, 4.0f, 12.0f};  // quality preset 12 (Fxaa3_11_mod.h:433-440)
	float ex[2], ey[2], endLuma[2];  // the two cursors: [0] against, [1] along the positive direction
	bool found[2];
	ex[0] = bx - along_x * STEP[0]; ey[0] = by - along_y * STEP[0];
	ex[1] = bx + along_x * STEP[0]; ey[1] = by + along_y * STEP[0];
	const float subpix1 = ((-2.0f) * subpix0) + 3.0f;
	endLuma[0] = luma_tap<PREPARED>(im, w, h, ex[0], ey[0]);
	const float subpix2 = subpix0 * subpix0;
	endLuma[1] = luma_tap<PREPARED>(im, w, h, ex[1], ey[1]);
	if (!steeperA) pairLuma = otherPair;
	const float threshold = gradient * 1.0f / 4.0f;
	const float centre_rel = a.m - pairLuma * 0.5f;
	const float subpix3 = subpix1 * subpix2;
	const bool centre_below = centre_rel < 0.0f;
	endLuma[0] -= pairLuma * 0.5f;
	endLuma[1] -= pairLuma * 0.5f;
	found[0] = fabsf(endLuma[0]) >= threshold;
	found[1] = fabsf(endLuma[1]) >= threshold;
#pragma unroll 1
	for (int k = 1;; ++k) {
		if (!found[0]) { ex[0] -= along_x * STEP[k]; ey[0] -= along_y * STEP[k]; }
		const bool searching = (!found[0]) || (!found[1]);
		if (!found[1]) { ex[1] += along_x * STEP[k]; ey[1] += along_y * STEP[k]; }
		if (!searching || k == 4) break;
#pragma unroll
		for (int d = 0; d < 2; ++d)
			if (!found[d]) endLuma[d] = luma_tap<PREPARED>(im, w, h, ex[d], ey[d]);
#pragma unroll
		for (int d = 0; d < 2; ++d)
			if (!found[d]) endLuma[d] = endLuma[d] - pairLuma * 0.5f;
		found[0] = fabsf(endLuma[0]) >= threshold;
		found[1] = fabsf(endLuma[1]) >= threshold;
	}
	// ---- where between the two ends does the pixel sit? ----
	float dist0 = cs - ex[0], dist1 = ex[1] - cs;
	if (!horizontal) { dist0 = ct - ey[0]; dist1 = ey[1] - ct; }
	const bool good0 = (endLuma[0] < 0.0f) != centre_below;
	const float span = dist1 + dist0;
	const bool good1 = (endLuma[1] < 0.0f) != centre_below;
	const float inv_span = 1.0f / span;
	const bool nearer0 = dist0 < dist1;
	const float nearest = fminf(dist0, dist1);
	const bool good = nearer0 ? good0 : good1;
	const float subpix4 = subpix3 * subpix3;
	const float edge_shift = (nearest * (-inv_span)) + 0.5f;
	const float subpix_shift = subpix4 * pc.fxaaQualitySubpix;
	const float shift = fmaxf(good ? edge_shift : 0.0f, subpix_shift);
	float fs = cs, ft = ct;
	if (!horizontal) fs += shift * across;
	if (horizontal) ft += shift * across;
	const float4 o = colour_tap<PREPARED>(im, w, h, fs, ft);
	return make_float4(o.x, o.y, o.z, a.m);
}

__global__ void __launch_bounds__(256) fxaa_prepare_kernel(const __grid_constant__ PostImg io) {
	const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
	if (x >= io.w || y >= io.h) return;
	float4 c = src_texel<true>(io.src, x, y);
	c.w = luma_of(c.x, c.y, c.z);
	st_rgba16f(io.dst, x, y, c);
}

// antialias_fxaa.comp:30-64. io.debug carries the segmentation mask (r32ui) here.
template <bool PREPARED>
__global__ void __launch_bounds__(FNT) fxaa_tile_kernel(const __grid_constant__ PostImg io, const __grid_constant__ TaaFxaaPush pc) {
	__shared__ LumaTile T;
	__shared__ unsigned short work[FTW * FTH];
	__shared__ int nwork;
	const int w = io.w, h = io.h;
	const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
	const int x0 = blockIdx.x * FTW, y0 = blockIdx.y * FTH;
	if (tid == 0) { nwork = 0; T.x0 = x0; T.y0 = y0; }
	// ---- the pixels the mask does not select are copied; is there anything else in this tile? ----
	bool selected[2];
#pragma unroll
	for (int r = 0; r < 2; ++r) {
		const int x = x0 + tx, y = y0 + ty + 8 * r;
		selected[r] = false;
		if (x < w && y < h) {
			const unsigned int seg = __ldg(reinterpret_cast<const unsigned int*>(io.debug.p + (long long)(y - io.debug.y0) * io.debug.pitch) + x) & 3u;
			selected[r] = seg == 1u;
			if (!selected[r]) st_rgba16f(io.dst, x, y, src_texel<PREPARED>(io.src, x, y));
		}
	}
	if (!__syncthreads_or((selected[0] || selected[1]) ? 1 : 0)) return;
	// ---- luma of the tile and its apron ----
	for (int k = tid; k < FLW * FLH; k += FNT) {
		const int i = k % FLW, j = k / FLW;
		T.v[j][i] = src_texel<PREPARED>(io.src, iclamp(x0 - FAP + i, 0, w - 1), iclamp(y0 - FAP + j, 0, h - 1)).w;
	}
	__syncthreads();
	// ---- the gate; what passes it goes to the work list ----
#pragma unroll
	for (int r = 0; r < 2; ++r) {
		if (!selected[r]) continue;
		const int x = x0 + tx, y = y0 + ty + 8 * r;
		const float cs = ((float)x + 0.5f) * pc.fxaaQualityRcpFrame[0], ct = ((float)y + 0.5f) * pc.fxaaQualityRcpFrame[1];
		const Around a = lumas_around(T, cs, ct, w, h);
		float range;
		if (gate_closed(a, pc, range)) st_rgba16f(io.dst, x, y, colour_tap<PREPARED>(io.src, w, h, cs, ct));
		else work[atomicAdd(&nwork, 1)] = (unsigned short)((ty + 8 * r) * FTW + tx);
	}
	__syncthreads();
	// ---- consecutive threads walk the edges of the listed pixels (any order: each pixel is independent) ----
	const int n = nwork;
	for (int k = tid; k < n; k += FNT) {
		const int p = work[k], x = x0 + (p % FTW), y = y0 + (p / FTW);
		const float cs = ((float)x + 0.5f) * pc.fxaaQualityRcpFrame[0], ct = ((float)y + 0.5f) * pc.fxaaQualityRcpFrame[1];
		const Around a = lumas_around(T, cs, ct, w, h);
		float range;
		gate_closed(a, pc, range);
		st_rgba16f(io.dst, x, y, walk_edge<PREPARED>(io.src, w, h, cs, ct, a, range, pc));
	}
}

inline dim3 grid2d(int w, int h, dim3 b) { return dim3((w + b.x - 1) / b.x, (h + b.y - 1) / b.y); }

}  // namespace

cudaError_t launch_fxaa_prepare(const PostImg& io, cudaStream_t stream) {
	dim3 b(32, 8);
	fxaa_prepare_kernel<<<grid2d(io.w, io.h, b), b, 0, stream>>>(io);
	return cudaGetLastError();
}
cudaError_t launch_fxaa(const PostImg& io, const TaaFxaaPush& pc, bool prepared, cudaStream_t stream) {
	const dim3 grid((io.w + FTW - 1) / FTW, (io.h + FTH - 1) / FTH);
	if (prepared) fxaa_tile_kernel<true><<<grid, FNT, 0, stream>>>(io, pc);
	else fxaa_tile_kernel<false><<<grid, FNT, 0, stream>>>(io, pc);
	return cudaGetLastError();
}

}  // namespace taa
