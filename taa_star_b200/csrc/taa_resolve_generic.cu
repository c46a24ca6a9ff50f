// taa_resolve_generic.cu — the EXACT, fully general resolve kernel: every switch of `Parameters`
// (shaders/taa.comp:50-95) is a warp-uniform runtime branch. One thread per output pixel.
// It exists so that ANY settings block the reference accepts runs on the GPU; the tuned kernels in
// taa_resolve_stream.cu / taa_resolve_strip.cu cover the BASELINE configs. Compiled with --fmad=false (see taa_device.cuh).
#include "taa_device.cuh"
#include "taa_kernels.h"
#include <cstdlib>

namespace taa {

namespace {

struct Px {
	const ResolveArgs& A;
	const TaaParameters& P;
	unsigned int* st;
	__device__ Px(const ResolveArgs& a, const TaaParameters& p) : A(a), P(p), st(a.status) {}

	// #define JITTER_UV (ubo.mJitterNdc.xy * 0.5 * params.mUnjitterFactor)           taa.comp:47
	__device__ float2 jitter_uv() const {
		return make_float2((A.ubo.mJitterNdc[0] * 0.5f) * P.mUnjitterFactor, (A.ubo.mJitterNdc[1] * 0.5f) * P.mUnjitterFactor);
	}
	__device__ f3 to_work_space(f3 rgb) const {  // maybe_rgb_to_ycocg(tonemap_rgb(.))   taa.comp:177,182
		if (P.mToneMapLumaKaris) rgb = tonemap_karis(rgb);
		return P.mUseYCoCg ? rgb_to_ycocg(rgb) : rgb;
	}
	__device__ float luminance(f3 c) const { return P.mUseYCoCg ? c.x : rgb_to_ycocg(c).x; }  // taa.comp:179

	// one tap of getNeighbourhood / getCurrentColor                                  taa.comp:204-219
	__device__ f3 colour_tap(float2 offset, int x, int y, float invw, float invh) const {
		float s = offset.x + ((float)x + 0.5f) * invw;
		float t = offset.y + ((float)y + 0.5f) * invh;
		return to_work_space(xyz(tex_rgba16f(A.color, A.in_w, A.in_h, s, t, st)));
	}

	// getCurrentUpsampledColor                                                       taa.comp:222-257
	__device__ f3 upsampled_colour(int cx, int cy, float& beta) const {
		const float lox = (float)A.in_w, loy = (float)A.in_h, hix = (float)A.out_w, hiy = (float)A.out_h;
		const float scx = hix / lox, scy = hiy / loy;
		float2 j = jitter_uv();
		const float tjx = (j.x * lox) * -1.0f, tjy = (j.y * loy) * -1.0f;
		const float almostOne = 0.999999f;
		float fx = -1.f, fy = -1.f;
#pragma unroll
		for (int i = 0; i < 4; ++i) {
			float px = (i & 1) ? almostOne : 0.f, py = (i & 2) ? almostOne : 0.f;
			float sx = (floorf((lox * ((float)cx + px)) / hix) + 0.5f) + tjx;
			float sy = (floorf((loy * ((float)cy + py)) / hiy) + 0.5f) + tjy;
			if ((int)(sx * scx) == cx && (int)(sy * scy) == cy) { fx = sx; fy = sy; }
		}
		if (fx >= 0.0f) {
			beta = 1.0f;
			float s = (floorf(fx) + 0.5f) / lox, t = (floorf(fy) + 0.5f) / loy;
			return to_work_space(xyz(tex_rgba16f(A.color, A.in_w, A.in_h, s, t, st)));
		}
		beta = 0.0f;
		return mk3(0.f, 0.f, 0.f);
	}

	// findClosestUvAndZ_3x3                                                          taa.comp:371-389
	__device__ float2 closest_uv_3x3(float u, float v) const {
		const float tx = 1.0f / (float)A.in_w, ty = 1.0f / (float)A.in_h;
		float cox = tx * -1.f, coy = ty * -1.f;
		float dClosest = tex_r32f(A.depth, A.in_w, A.in_h, u + cox, v + coy, st);
#pragma unroll
		for (int i = 1; i < 9; ++i) {
			float ox = tx * (float)(i % 3 - 1), oy = ty * (float)(i / 3 - 1);
			float d = tex_r32f(A.depth, A.in_w, A.in_h, u + ox, v + oy, st);
			if (d < dClosest) { cox = ox; coy = oy; dClosest = d; }
		}
		return make_float2(u + cox, v + coy);
	}

	// sample_history_rgba                                                            taa.comp:441-549
	__device__ float4 hist(float s, float t) const { return tex_rgba16f(A.history_in, A.out_w, A.out_h, s, t, st); }
	__device__ float4 sample_history(float u, float v) const {
		if (P.mInterpolationMode == 0) return hist(u, v);
		const float W = (float)A.out_w, H = (float)A.out_h;
		const float iw = 1.0f / W, ih = 1.0f / H;
		const float ix = u * W, iy = v * H;
		const float tcx = floorf(ix - 0.5f) + 0.5f, tcy = floorf(iy - 0.5f) + 0.5f;
		const float fx = ix - tcx, fy = iy - tcy;
		const float fx2 = fx * fx, fy2 = fy * fy, fx3 = fx2 * fx, fy3 = fy2 * fy;
		if (P.mInterpolationMode == 1) {  // b-spline, 4 bilinear taps                     taa.comp:517-543
			float w0x = fx2 - 0.5f * (fx3 + fx), w0y = fy2 - 0.5f * (fy3 + fy);
			float w1x = 1.5f * fx3 - 2.5f * fx2 + 1.0f, w1y = 1.5f * fy3 - 2.5f * fy2 + 1.0f;
			float w3x = 0.5f * (fx3 - fx2), w3y = 0.5f * (fy3 - fy2);
			float w2x = 1.0f - w0x - w1x - w3x, w2y = 1.0f - w0y - w1y - w3y;
			float s0x = w0x + w1x, s0y = w0y + w1y, s1x = w2x + w3x, s1y = w2y + w3y;
			float f0x = w1x / (w0x + w1x), f0y = w1y / (w0y + w1y);
			float f1x = w3x / (w2x + w3x), f1y = w3y / (w2y + w3y);
			float t0x = (tcx - 1.0f + f0x) * iw, t0y = (tcy - 1.0f + f0y) * ih;
			float t1x = (tcx + 1.0f + f1x) * iw, t1y = (tcy + 1.0f + f1y) * ih;
			return (hist(t0x, t0y) * s0x + hist(t1x, t0y) * s1x) * s0y + (hist(t0x, t1y) * s0x + hist(t1x, t1y) * s1x) * s1y;
		}
		// catmull-rom, 9 bilinear taps                                                   taa.comp:441-514
		float w0x = -0.5f * fx3 + fx2 - 0.5f * fx, w0y = -0.5f * fy3 + fy2 - 0.5f * fy;
		float w1x = 1.5f * fx3 - 2.5f * fx2 + 1.0f, w1y = 1.5f * fy3 - 2.5f * fy2 + 1.0f;
		float w2x = -1.5f * fx3 + 2.0f * fx2 + 0.5f * fx, w2y = -1.5f * fy3 + 2.0f * fy2 + 0.5f * fy;
		float w3x = 0.5f * fx3 - 0.5f * fx2, w3y = 0.5f * fy3 - 0.5f * fy2;
		float wCx = w1x + w2x, wCy = w1y + w2y;
		float t0x = (tcx - 1.0f) * iw, t0y = (tcy - 1.0f) * ih;
		float tCx = (tcx + w2x / wCx) * iw, tCy = (tcy + w2y / wCy) * ih;
		float t3x = (tcx + 2.0f) * iw, t3y = (tcy + 2.0f) * ih;
		float4 r = hist(t0x, t0y) * w0x * w0y;
		r = r + hist(tCx, t0y) * wCx * w0y;
		r = r + hist(t3x, t0y) * w3x * w0y;
		r = r + hist(t0x, tCy) * w0x * wCy;
		r = r + hist(tCx, tCy) * wCx * wCy;
		r = r + hist(t3x, tCy) * w3x * wCy;
		r = r + hist(t0x, t3y) * w0x * w3y;
		r = r + hist(tCx, t3y) * wCx * w3y;
		r = r + hist(t3x, t3y) * w3x * w3y;
		return r;
	}

	// ---- segmentation mask helpers                                                  taa.comp:559-587
	__device__ float lin_depth(int x, int y) const {
		float d = fetch_r32f(A.depth, A.in_w, A.in_h, x, y, st);
		float n = A.ubo.mCamNearPlane, f = A.ubo.mCamFarPlane;
		return n * f / (f + d * (n - f));
	}
	__device__ float luma_at(int x, int y) const { return rgb_to_ycocg(xyz(fetch_rgba16f(A.color, A.in_w, A.in_h, x, y, st))).x; }
	__device__ f3 normal_at(int x, int y) const {
		float4 n = fetch_rgba32f(A.uvnrm, A.in_w, A.in_h, x, y, st);
		return mk3(taa_cos(n.z) * taa_cos(n.w), taa_sin(n.z) * taa_cos(n.w), taa_sin(n.w));
	}
	__device__ static float dot3(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
	template <class F>
	__device__ float sobel_len(int x, int y, F f) const {
		const int xl = iclamp(x - 1, 0, A.in_w - 1), xr = iclamp(x + 1, 0, A.in_w - 1), xc = iclamp(x, 0, A.in_w - 1);
		const int yt = iclamp(y - 1, 0, A.in_h - 1), yb = iclamp(y + 1, 0, A.in_h - 1), yc = iclamp(y, 0, A.in_h - 1);
		float c00 = f(xl, yt), c01 = f(xc, yt), c02 = f(xr, yt), c10 = f(xl, yc), c12 = f(xr, yc), c20 = f(xl, yb), c21 = f(xc, yb), c22 = f(xr, yb);
		float gx = c00 - c20 + 2.0f * c01 - 2.0f * c21 + c02 - c22;
		float gy = c00 - c02 + 2.0f * c10 - 2.0f * c12 + c20 - c22;
		return sqrtf(gx * gx + gy * gy);
	}
	// calc_segmentation_value                                                        taa.comp:589-702
	__device__ unsigned int segmentation(int x, int y, float hu, float hv) const {
		const unsigned int flags = P.mRayTraceAugmentFlags;
		if (flags & TAA_RTFLAG_FXD) {
			const int b = 100;
			if (x < b || y < b || x >= A.in_w - b || y >= A.in_h - b) return 1u;
		}
		if (flags & TAA_RTFLAG_ALL) return 2u;
		const bool useCnt = (flags & TAA_RTFLAG_CNT) != 0;
		const unsigned int newCnt = useCnt ? ((unsigned int)P.mRayTraceHistoryCount << 16) : 0u;
		if (flags & TAA_RTFLAG_OUT) {
			if (hu < 0.f || hv < 0.f || hu >= 1.f || hv >= 1.f) return 1u;
		}
		const unsigned int matId = fetch_r32ui(A.matid, A.in_w, A.in_h, x, y, st);
		if (flags & TAA_RTFLAG_DIS) {
			int px = (int)(hu * (float)A.in_w), py = (int)(hv * (float)A.in_h);
			if (px >= 0 && py >= 0 && px < A.in_w && py < A.in_h) {
				unsigned int prev = fetch_r32ui(A.prev_matid, A.in_w, A.in_h, px, py, st);
				if (prev != matId && (prev & 0x80000000u)) return 2u | newCnt;
			}
		}
		float nrm = 0.f, dpt = 0.f, mat = 0.f, lum = 0.f;
		const int xl = iclamp(x - 1, 0, A.in_w - 1), xr = iclamp(x + 1, 0, A.in_w - 1), xc = iclamp(x, 0, A.in_w - 1);
		const int yt = iclamp(y - 1, 0, A.in_h - 1), yb = iclamp(y + 1, 0, A.in_h - 1), yc = iclamp(y, 0, A.in_h - 1);
		if (flags & TAA_RTFLAG_NRM) {
			f3 nC = normal_at(x, y), nL = normal_at(xl, yc), nR = normal_at(xr, yc), nT = normal_at(xc, yt), nB = normal_at(xc, yb);
			float mind = fmaxf(0.f, fminf(fminf(fminf(dot3(nC, nL), dot3(nC, nR)), dot3(nC, nT)), dot3(nC, nB)));
			nrm = 1.0f - mind;
		}
		if (flags & TAA_RTFLAG_DPT) dpt = sobel_len(x, y, [&](int a, int b) { return lin_depth(a, b); });
		if (flags & TAA_RTFLAG_MID) {
			if (matId != fetch_r32ui(A.matid, A.in_w, A.in_h, xl, yc, st) || matId != fetch_r32ui(A.matid, A.in_w, A.in_h, xr, yc, st) ||
			    matId != fetch_r32ui(A.matid, A.in_w, A.in_h, xc, yt, st) || matId != fetch_r32ui(A.matid, A.in_w, A.in_h, xc, yb, st))
				mat = 1.0f;
		}
		if (flags & TAA_RTFLAG_LUM) lum = sobel_len(x, y, [&](int a, int b) { return luma_at(a, b); });
		float total = nrm * P.mRayTraceAugment_WNrm + dpt * P.mRayTraceAugment_WDpt + mat * P.mRayTraceAugment_WMId + lum * P.mRayTraceAugment_WLum;
		if (total >= P.mRayTraceAugment_Thresh) return 2u | newCnt;
		if (useCnt) {
			unsigned int oldCnt = (fetch_r32ui(A.prev_segmask, A.out_w, A.out_h, x, y, st) & 0xffff0000u) >> 16;
			if (oldCnt > 0) return 2u | ((oldCnt - 1) << 16);
		}
		return 0u;
	}
};

}  // namespace

// main() for one output pixel                                                         taa.comp:708-960
// WRITE_SCREEN = false: the screen result is left alone (fix-up of a fused frame, where the follow-on passes
// already consumed the tuned kernel's on-chip result).
// where the nine neighbourhood taps (and the current colour) come from: straight through the software sampler ...
struct DirectTaps {
	__device__ __forceinline__ f3 operator()(const Px& px, float2 off, int x, int y, float invw, float invh) const { return px.colour_tap(off, x, y, invw, invh); }
};
// ... or out of a shared-memory tile in which every texel's tap was evaluated once (SPEC = 1: the tap of texel (x, y) is a function of (x, y)
// alone there — no unjitter offset, no upsampling —, so nine pixels share it; same function, same bits)
struct TileTaps {
	const float* tile;  // [3][TILE_H + 2][TILE_W + 2]
	int x0, y0, w, h;   // texel of tile entry (1, 1); tile extent without the apron
	__device__ __forceinline__ f3 operator()(const Px& px, float2 off, int x, int y, float invw, float invh) const {
		// (x, y) is the pixel's own texel or a neighbour: uv_to_tc(tc_to_uv(x)) == x for every frame size below 2^22, so the tap is in the tile;
		// the clamp only keeps an impossible index inside the array
		const int i = min(max(x - x0 + 1, 0), w + 1), j = min(max(y - y0 + 1, 0), h + 1);
		const int n = (w + 2) * (h + 2), k = j * (w + 2) + i;
		return mk3(tile[k], tile[n + k], tile[2 * n + k]);
	}
};

// SPEC = 1: the switch pattern of the reference's DEFAULT settings (taa.hpp:31-76: RGB, min / max box, clamp, bilinear history, velocity for
// movers only and matrix reprojection for the rest, no rejection, no alpha modulation) folded at compile time. The code is this very function:
// the folded switches only remove what those settings never execute.
__device__ __forceinline__ void fold_reference_defaults(TaaParameters& P) {
	P.mPassThrough = 0; P.mUseYCoCg = 0; P.mShrinkChromaAxis = 0; P.mVarianceClipping = 0; P.mShapedNeighbourhood = 0; P.mColorClampingOrClipping = 1;
	P.mUnjitterNeighbourhood = 0; P.mUnjitterCurrentSample = 0; P.mToneMapLumaKaris = 0; P.mAddNoise = 0; P.mRayTraceAugment = 0;
	P.mUseVelocityVectors = 1; P.mVelocitySampleMode = 0; P.mInterpolationMode = 0; P.mRejectOutside = 0; P.mDepthCulling = 0;
	P.mDynamicAntiGhosting = 0; P.mVelBasedAlpha = 0; P.mLumaWeightingLottes = 0; P.mReduceBlendNearClamp = 0;
}

template <bool WRITE_SCREEN, int SPEC = 0, class TAPS = DirectTaps>
__device__ __forceinline__ void resolve_pixel_exact(const ResolveArgs& A, const int x, const int y, const TAPS taps = TAPS()) {
	unsigned int* st = A.status;

	TaaParameters Pl;
	const TaaParameters* Pp = &A.ubo.param[(SPEC == 0 && A.ubo.splitScreen && x > A.ubo.splitX) ? 1 : 0];
	if (SPEC == 1) { Pl = *Pp; fold_reference_defaults(Pl); Pp = &Pl; }
	const TaaParameters& P = *Pp;
	Px px(A, P);

	const float u = ((float)x + 0.5f) / (float)A.out_w;  // tc_to_uv                       taa.comp:131
	const float v = ((float)y + 0.5f) / (float)A.out_h;
	const int lx = (int)(u * (float)A.in_w), ly = (int)(v * (float)A.in_h);  // uv_to_tc  taa.comp:724

	if (P.mPassThrough) {  // taa.comp:726-731
		float4 c = fetch_rgba16f(A.color, A.in_w, A.in_h, lx, ly, st);
		float4 h = fetch_rgba16f(A.history_in, A.out_w, A.out_h, x, y, st);
		if (WRITE_SCREEN) st_rgba16f(A.result, x, y, make_float4(c.x, c.y, c.z, 1.f));
		st_rgba16f(A.history_out, x, y, make_float4(h.x, h.y, h.z, 1.f));
		st_rgba16f(A.debug, x, y, make_float4(0.f, 0.f, 0.f, 0.f));
		st_r32ui(A.mask, x, y, 0u);
		return;
	}
	if (SPEC == 0 && A.ubo.mBypassHistoryUpdate) {  // taa.comp:732-737
		float4 h = fetch_rgba16f(A.history_in, A.out_w, A.out_h, x, y, st);
		if (WRITE_SCREEN) st_rgba16f(A.result, x, y, make_float4(h.x, h.y, h.z, 1.f));
		st_rgba16f(A.history_out, x, y, make_float4(h.x, h.y, h.z, 1.f));
		st_rgba16f(A.debug, x, y, make_float4(0.f, 0.f, 0.f, 0.f));
		st_r32ui(A.mask, x, y, 0u);
		return;
	}

	// ---- getColorAndAabb                                                             taa.comp:259-319
	const float invw = 1.0f / (float)A.in_w, invh = 1.0f / (float)A.in_h;
	f3 colMin, colMax, clipTowards;
	{
		float2 off = P.mUnjitterNeighbourhood ? px.jitter_uv() : make_float2(0.f, 0.f);
		f3 cC = taps(px, off, lx, ly, invw, invh);
		f3 c1 = taps(px, off, lx - 1, ly - 1, invw, invh);
		f3 c2 = taps(px, off, lx, ly - 1, invw, invh);
		f3 c3 = taps(px, off, lx + 1, ly - 1, invw, invh);
		f3 c4 = taps(px, off, lx - 1, ly, invw, invh);
		f3 c5 = taps(px, off, lx + 1, ly, invw, invh);
		f3 c6 = taps(px, off, lx - 1, ly + 1, invw, invh);
		f3 c7 = taps(px, off, lx, ly + 1, invw, invh);
		f3 c8 = taps(px, off, lx + 1, ly + 1, invw, invh);
		if (P.mVarianceClipping) {
			const float N = 9.0f;
			f3 m1 = cC + c1 + c2 + c3 + c4 + c5 + c6 + c7 + c8;
			f3 m2 = cC * cC + c1 * c1 + c2 * c2 + c3 * c3 + c4 * c4 + c5 * c5 + c6 * c6 + c7 * c7 + c8 * c8;
			f3 mean = m1 / N;
			f3 var = max3(mk3(0.f, 0.f, 0.f), m2 / N - mean * mean);
			f3 sigma = mk3(sqrtf(var.x), sqrtf(var.y), sqrtf(var.z));
			colMin = mean - P.mVarClipGamma * sigma;
			colMax = mean + P.mVarClipGamma * sigma;
			clipTowards = mean;
		} else if (P.mShapedNeighbourhood) {
			f3 mn9 = min3(min3(min3(min3(min3(min3(min3(min3(cC, c1), c2), c3), c4), c5), c6), c7), c8);
			f3 mx9 = max3(max3(max3(max3(max3(max3(max3(max3(cC, c1), c2), c3), c4), c5), c6), c7), c8);
			f3 mn5 = min3(min3(min3(min3(cC, c2), c4), c5), c7);
			f3 mx5 = max3(max3(max3(max3(cC, c2), c4), c5), c7);
			colMin = (mn9 + mn5) * 0.5f;
			colMax = (mx9 + mx5) * 0.5f;
			clipTowards = cC;
		} else {
			colMin = min3(min3(min3(min3(min3(min3(min3(min3(cC, c1), c2), c3), c4), c5), c6), c7), c8);
			colMax = max3(max3(max3(max3(max3(max3(max3(max3(cC, c1), c2), c3), c4), c5), c6), c7), c8);
			clipTowards = cC;
		}
		if (P.mUseYCoCg && P.mShrinkChromaAxis) {
			f3 halfSize = mk3(0.5f * 1.0f, 0.5f * 0.5f, 0.5f * 0.5f) * (colMax - colMin);
			f3 center = (colMin + colMax) * 0.5f;
			colMin = center - halfSize;
			colMax = center + halfSize;
			if (clipTowards.x < colMin.x || clipTowards.y < colMin.y || clipTowards.z < colMin.z ||
			    clipTowards.x > colMax.x || clipTowards.y > colMax.y || clipTowards.z > colMax.z)
				clipTowards = center;
		}
	}

	// ---- current colour                                                              taa.comp:750-756
	f3 cur;
	float beta;
	if (SPEC == 0 && A.ubo.mUpsampling) {
		cur = px.upsampled_colour(x, y, beta);
	} else {
		float2 off = P.mUnjitterCurrentSample ? px.jitter_uv() : make_float2(0.f, 0.f);
		cur = taps(px, off, lx, ly, invw, invh);
		beta = 1.0f;
	}
	const float depth = fetch_r32f(A.depth, A.in_w, A.in_h, lx, ly, st);

	// ---- getHistoryPosition                                                          taa.comp:391-438
	float hu, hv, expectedHistoryDepth;
	{
		float4 vel = tex_rgba16f(A.velocity, A.in_w, A.in_h, u, v, st);
		bool canUseVelocity = !(P.mUseVelocityVectors == 0 || (P.mUseVelocityVectors == 1 && vel.w < 0.5f));
		if (canUseVelocity) {
			if (P.mVelocitySampleMode == 1) {
				const int ox[8] = {1, -1, -1, 0, -1, 0, 1, 1}, oy[8] = {-1, 0, -1, -1, 1, 1, 0, 1};
				float mvx = vel.x, mvy = vel.y, sx = vel.x, sy = vel.y;
#pragma unroll
				for (int i = 0; i < 8; ++i) {
					float4 s4 = tex_rgba16f(A.velocity, A.in_w, A.in_h, u + invw * (float)ox[i], v + invh * (float)oy[i], st);
					sx = s4.x;
					sy = s4.y;
					if (sx * sx + sy * sy > mvx * mvx + mvy * mvy) { mvx = sx; mvy = sy; }
				}
				vel.x = sx;  // taa.comp:412 (sic)
				vel.y = sy;
			} else if (P.mVelocitySampleMode == 2) {
				float2 c = px.closest_uv_3x3(u, v);
				vel = tex_rgba16f(A.velocity, A.in_w, A.in_h, c.x, c.y, st);
			}
			hu = u - vel.x;
			hv = v - vel.y;
			expectedHistoryDepth = depth - vel.z;
		} else {
			const float* Mi = A.ubo.mInverseViewProjMatrix;
			const float* Mh = A.ubo.mHistoryViewProjMatrix;
			float cx = u * 2.0f - 1.0f, cy = v * 2.0f - 1.0f, cz = depth, cw = 1.0f;
			float wx = ((Mi[0] * cx + Mi[4] * cy) + Mi[8] * cz) + Mi[12] * cw;
			float wy = ((Mi[1] * cx + Mi[5] * cy) + Mi[9] * cz) + Mi[13] * cw;
			float wz = ((Mi[2] * cx + Mi[6] * cy) + Mi[10] * cz) + Mi[14] * cw;
			float ww = ((Mi[3] * cx + Mi[7] * cy) + Mi[11] * cz) + Mi[15] * cw;
			float hx = ((Mh[0] * wx + Mh[4] * wy) + Mh[8] * wz) + Mh[12] * ww;
			float hy = ((Mh[1] * wx + Mh[5] * wy) + Mh[9] * wz) + Mh[13] * ww;
			float hz = ((Mh[2] * wx + Mh[6] * wy) + Mh[10] * wz) + Mh[14] * ww;
			float hw = ((Mh[3] * wx + Mh[7] * wy) + Mh[11] * wz) + Mh[15] * ww;
			hu = (hx / hw) * 0.5f + 0.5f;
			hv = (hy / hw) * 0.5f + 0.5f;
			expectedHistoryDepth = hz / hw;
		}
	}
	const float du = u - hu, dv = v - hv;
	const float pixelSpeed = sqrtf(du * du + dv * dv);

	const float4 historyRaw = px.sample_history(hu, hv);
	f3 hist = P.mUseYCoCg ? rgb_to_ycocg(xyz(historyRaw)) : xyz(historyRaw);

	float alpha = P.mAlpha;
	bool rejected = false;

	unsigned int seg = 0u;
	const bool genSeg = P.mRayTraceAugment != 0;
	if (genSeg) {  // taa.comp:775-784
		seg = px.segmentation(x, y, hu, hv);
		if (seg & 3u) { alpha = P.mRejectionAlpha; rejected = true; }
	}

	// ---- history rejection                                                           taa.comp:787-823
	if (P.mRejectOutside) {
		if (hu < 0.f || hv < 0.f || hu >= 1.f || hv >= 1.f) { alpha = P.mRejectionAlpha; rejected = true; }
	}
	float writeDynamicMask = 0.f;
	if (P.mDynamicAntiGhosting) {
		const float eps = 1e-5f;
		auto mov = [&](float s, float t) {
			float4 q = tex_rgba16f(A.velocity, A.in_w, A.in_h, s, t, st);
			return (fabsf(q.x) > eps || fabsf(q.y) > eps) && (fabsf(q.w) >= 0.5f);
		};
		bool movL = mov(u + invw * -1.f, v + invh * 0.f);
		bool movR = mov(u + invw * 1.f, v + invh * 0.f);
		bool movT = mov(u + invw * 0.f, v + invh * -1.f);
		bool movB = mov(u + invw * 0.f, v + invh * 1.f);
		bool movC = mov(u, v);
		if (!(movL || movR || movT || movB || movC) && historyRaw.w > 0.0f) rejected = true;
		writeDynamicMask = movC ? 1.0f : 0.0f;
	}
	if (P.mDepthCulling) {
		int tx = (int)(hu * (float)A.in_w), ty = (int)(hv * (float)A.in_h);
		float hd = fetch_r32f(A.history_depth, A.in_w, A.in_h, tx, ty, st);
		float depthEpsilon = 0.1f * (1.0f - hd);
		if (fabsf(hd - expectedHistoryDepth) > depthEpsilon) rejected = true;
	}

	// ---- rectification                                                               taa.comp:826-845
	const f3 origHist = hist;
	switch (P.mColorClampingOrClipping) {
		case 1: hist = min3(max3(hist, colMin), colMax); break;
		case 2: {  // clipAabb(colMin, colMax, vec4(0,0,0,1), vec4(hist,1))               taa.comp:323-345
			const float eps = 1e-7f;
			f3 pClip = 0.5f * (colMax + colMin);
			f3 e = 0.5f * (colMax - colMin);
			f3 eClip = mk3(e.x + eps, e.y + eps, e.z + eps);
			f3 vClip = hist - pClip;
			f3 aUnit = abs3(vClip / eClip);
			float maUnit = fmaxf(aUnit.x, fmaxf(aUnit.y, aUnit.z));
			if (maUnit > 1.0f) hist = pClip + vClip / maUnit;
			break;
		}
		case 3: {  // clipAabbSlow(colMin, colMax, vec4(clipTowards,1), vec4(hist,1))     taa.comp:348-369
			const float eps = 1e-7f;
			f3 p = clipTowards;
			f3 r = hist - p;
			f3 rmax = colMax - p, rmin = colMin - p;
			if (r.x > rmax.x + eps) r = r * (rmax.x / r.x);
			if (r.y > rmax.y + eps) r = r * (rmax.y / r.y);
			if (r.z > rmax.z + eps) r = r * (rmax.z / r.z);
			if (r.x < rmin.x - eps) r = r * (rmin.x / r.x);
			if (r.y < rmin.y - eps) r = r * (rmin.y / r.y);
			if (r.z < rmin.z - eps) r = r * (rmin.z / r.z);
			hist = p + r;
			break;
		}
		default: break;
	}
	const f3 rdiff = hist - origHist;
	const bool rectified = fabsf(rdiff.x) > 0.001f || fabsf(rdiff.y) > 0.001f || fabsf(rdiff.z) > 0.001f;

	// ---- blending                                                                    taa.comp:848-900
	if (rejected) {
		alpha = P.mRejectionAlpha;
		beta = 1.0f;
	} else {
		if (P.mVelBasedAlpha) alpha = fmaxf(alpha, mixf(alpha, P.mVelBasedAlphaMax, clampf(pixelSpeed * P.mVelBasedAlphaFactor, 0.f, 1.f)));
		if (P.mLumaWeightingLottes) {
			float lc = px.luminance(cur), lh = px.luminance(hist);
			float diff = fabsf(lc - lh) / fmaxf(fmaxf(lc, lh), 0.2f);
			float w = 1.0f - diff;
			alpha = mixf(P.mMaxAlpha, P.mMinAlpha, w * w);
		}
		if (P.mReduceBlendNearClamp) {
			float lmin = px.luminance(colMin), lmax = px.luminance(colMax), lh = px.luminance(origHist);
			float distToClamp = 2.0f * fabsf(fminf(lh - lmin, lmax - lh)) / (lmax - lmin);
			if (lmax - lmin < 0.001f) distToClamp = 1.0f;
			alpha *= clampf(4.0f * distToClamp, 0.f, 1.f);
		}
	}
	if (A.ubo.mResetHistory) { alpha = 1.0f; beta = 1.0f; }

	const float ab = alpha * beta;
	f3 aa = mk3(mixf(hist.x, cur.x, ab), mixf(hist.y, cur.y, ab), mixf(hist.z, cur.z, ab));
	if (P.mUseYCoCg) aa = ycocg_to_rgb(aa);
	if (P.mAddNoise) {  // noise()                                                        taa.comp:551-556
		float sx = u + A.ubo.mSinTime[0] + 0.6959174f, sy = v + A.ubo.mSinTime[0] + 0.6959174f;
		float s = taa_sin(sx * 12.9898f + sy * 78.233f);
		float n0 = s * 43758.5453f, n1 = s * 28001.8384f, n2 = s * 50849.4141f;
		n0 = n0 - floorf(n0); n1 = n1 - floorf(n1); n2 = n2 - floorf(n2);
		aa.x = aa.x + (n0 * 2.0f - 1.0f) * P.mNoiseFactor;
		aa.y = aa.y + (n1 * 2.0f - 1.0f) * P.mNoiseFactor;
		aa.z = aa.z + (n2 * 2.0f - 1.0f) * P.mNoiseFactor;
	}
	const float4 toHistory = mk4(aa, writeDynamicMask);
	const float4 toScreen = mk4(P.mToneMapLumaKaris ? un_tonemap_karis(aa) : aa, 1.0f);

	st_rgba16f(A.history_out, x, y, toHistory);
	if (WRITE_SCREEN) st_rgba16f(A.result, x, y, toScreen);

	if (A.debug.p) {  // taa.comp:912-942, 957
		float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
		switch (P.mDebugMode) {
			case 0: d = mk4(colMax - colMin, 0.f); break;
			case 1: { f3 t = colMax - colMin; float q = t.x * t.y * t.z; d = make_float4(q, q, q, 0.f); break; }
			case 2: d = make_float4(rejected ? 1.f : 0.f, sqrtf(rdiff.x * rdiff.x + rdiff.y * rdiff.y + rdiff.z * rdiff.z), 0.f, 0.f); break;
			case 3: d = make_float4(alpha, alpha, alpha, 0.f); break;
			case 4: d = fetch_rgba16f(A.velocity, A.in_w, A.in_h, x, y, st); break;
			case 5: d = make_float4(pixelSpeed, 0.f, 0.f, 0.f); break;
			case 6: d = toScreen; break;
			case 7: d = toHistory; break;
			case 8:
				switch (seg & 3u) {
					case 0: d = make_float4(0.f, 0.f, 1.f, 0.f); break;
					case 1: d = make_float4(1.f, 0.f, 0.f, 0.f); break;
					case 2: d = make_float4(1.f, 1.f, 0.f, 0.f); break;
					default: break;
				}
				break;
			default: break;
		}
		const float sc = P.mDebugScale;
		d = make_float4(d.x * (P.mDebugMask[0] * sc), d.y * (P.mDebugMask[1] * sc), d.z * (P.mDebugMask[2] * sc), d.w * (P.mDebugMask[3] * sc));
		if (P.mDebugCenter) d = make_float4(d.x * 0.5f + 0.5f, d.y * 0.5f + 0.5f, d.z * 0.5f + 0.5f, d.w * 0.5f + 0.5f);
		st_rgba16f(A.debug, x, y, d);
	}
	if (genSeg) st_r32ui(A.segmask, x, y, seg);
	st_r32ui(A.mask, x, y, (rejected ? 1u : 0u) | (rectified ? 2u : 0u) | (((unsigned int)P.mColorClampingOrClipping & 3u) << 2));
}

__global__ void __launch_bounds__(256) taa_resolve_generic_kernel(const __grid_constant__ ResolveArgs A) {
	const int x = blockIdx.x * blockDim.x + threadIdx.x;
	const int y = A.band_y0 + blockIdx.y * blockDim.y + threadIdx.y;
	if (x >= A.out_w || y >= A.band_y0 + A.band_rows || y >= A.out_h) return;
	resolve_pixel_exact<true>(A, x, y);
}

// The same arithmetic as resolve_pixel_exact — operation for operation — for the settings family of the tuned kernel
// (tuned_supports(): YCoCg, variance box, clipAabb, Catmull-Rom, velocity for everything, no tonemap / unjitter / noise / TAAU /
// seg-mask / debug), with the constant switches folded and the sampler coordinates of the 3x3 and Catmull-Rom tap grids
// evaluated once per row and per column instead of once per tap. Used by the fix-up pass, where it is ~2.5x cheaper.
struct RowPair { const uint2* r0; const uint2* r1; };
__device__ __forceinline__ f3 lerp3(f3 p, f3 q, float w) { return mk3(lerpf(p.x, q.x, w), lerpf(p.y, q.y, w), lerpf(p.z, q.z, w)); }
__device__ __forceinline__ f3 rgb_of(uint2 raw) {
	float4 t = unpack_rgba16f(raw);
	return mk3(t.x, t.y, t.z);
}
__device__ __forceinline__ f3 tap3(const RowPair& R, const Lin& X, float ya) {
	return lerp3(lerp3(rgb_of(__ldg(R.r0 + X.i0)), rgb_of(__ldg(R.r0 + X.i1)), X.a), lerp3(rgb_of(__ldg(R.r1 + X.i0)), rgb_of(__ldg(R.r1 + X.i1)), X.a), ya);
}
__device__ __forceinline__ float4 tap4(const RowPair& R, const Lin& X, float ya) {
	return lerp4(lerp4(unpack_rgba16f(__ldg(R.r0 + X.i0)), unpack_rgba16f(__ldg(R.r0 + X.i1)), X.a),
	             lerp4(unpack_rgba16f(__ldg(R.r1 + X.i0)), unpack_rgba16f(__ldg(R.r1 + X.i1)), X.a), ya);
}

template <bool WRITE_SCREEN>
__device__ __forceinline__ void resolve_pixel_exact_family(const ResolveArgs& A, const int x, const int y) {
	unsigned int* st = A.status;
	const TaaParameters& P = A.ubo.param[0];
	const int W = A.out_w, H = A.out_h;  // input and output sizes are equal in this family
	const float fW = (float)W, fH = (float)H;
	const float u = ((float)x + 0.5f) / fW, v = ((float)y + 0.5f) / fH;
	const int lx = (int)(u * fW), ly = (int)(v * fH);
	const float invw = 1.0f / fW, invh = 1.0f / fH;

	// ---- getColorAndAabb, variance branch (taa.comp:259-277); offset = vec2(0) is still added, as in the shader ----
	f3 colMin, colMax, cur;
	{
		Lin cx[3], cy[3];
		RowPair rows[3];
#pragma unroll
		for (int d = 0; d < 3; ++d) {
			cx[d] = lin_coord(0.0f + ((float)(lx + d - 1) + 0.5f) * invw, W);
			cy[d] = lin_coord(0.0f + ((float)(ly + d - 1) + 0.5f) * invh, H);
			rows[d].r0 = reinterpret_cast<const uint2*>(row_ptr(A.color, cy[d].i0, st));
			rows[d].r1 = reinterpret_cast<const uint2*>(row_ptr(A.color, cy[d].i1, st));
		}
		const f3 cC = rgb_to_ycocg(tap3(rows[1], cx[1], cy[1].a));
		const f3 c1 = rgb_to_ycocg(tap3(rows[0], cx[0], cy[0].a)), c2 = rgb_to_ycocg(tap3(rows[0], cx[1], cy[0].a)), c3 = rgb_to_ycocg(tap3(rows[0], cx[2], cy[0].a));
		const f3 c4 = rgb_to_ycocg(tap3(rows[1], cx[0], cy[1].a)), c5 = rgb_to_ycocg(tap3(rows[1], cx[2], cy[1].a));
		const f3 c6 = rgb_to_ycocg(tap3(rows[2], cx[0], cy[2].a)), c7 = rgb_to_ycocg(tap3(rows[2], cx[1], cy[2].a)), c8 = rgb_to_ycocg(tap3(rows[2], cx[2], cy[2].a));
		const float N = 9.0f;
		f3 m1 = cC + c1 + c2 + c3 + c4 + c5 + c6 + c7 + c8;
		f3 m2 = cC * cC + c1 * c1 + c2 * c2 + c3 * c3 + c4 * c4 + c5 * c5 + c6 * c6 + c7 * c7 + c8 * c8;
		f3 mean = m1 / N;
		f3 var = max3(mk3(0.f, 0.f, 0.f), m2 / N - mean * mean);
		f3 sigma = mk3(sqrtf(var.x), sqrtf(var.y), sqrtf(var.z));
		colMin = mean - P.mVarClipGamma * sigma;
		colMax = mean + P.mVarClipGamma * sigma;
		cur = cC;  // getCurrentColor (taa.comp:215-220) repeats the centre tap with the same (zero) offset
	}
	const float depth = fetch_r32f(A.depth, W, H, lx, ly, st);

	// ---- getHistoryPosition, velocity for everything, simple sample (taa.comp:391-438) ----
	const float4 vel = tex_rgba16f(A.velocity, W, H, u, v, st);
	const float hu = u - vel.x, hv = v - vel.y;
	const float expectedHistoryDepth = depth - vel.z;
	const float du = u - hu, dv = v - hv;
	const float pixelSpeed = sqrtf(du * du + dv * dv);

	// ---- sample_history_bicubic_catmullrom (taa.comp:441-514) ----
	float4 historyRaw;
	{
		const float iw = 1.0f / fW, ih = 1.0f / fH;
		const float ix = hu * fW, iy = hv * fH;
		const float tcx = floorf(ix - 0.5f) + 0.5f, tcy = floorf(iy - 0.5f) + 0.5f;
		const float fx = ix - tcx, fy = iy - tcy;
		const float fx2 = fx * fx, fy2 = fy * fy, fx3 = fx2 * fx, fy3 = fy2 * fy;
		const float w0x = -0.5f * fx3 + fx2 - 0.5f * fx, w0y = -0.5f * fy3 + fy2 - 0.5f * fy;
		const float w1x = 1.5f * fx3 - 2.5f * fx2 + 1.0f, w1y = 1.5f * fy3 - 2.5f * fy2 + 1.0f;
		const float w2x = -1.5f * fx3 + 2.0f * fx2 + 0.5f * fx, w2y = -1.5f * fy3 + 2.0f * fy2 + 0.5f * fy;
		const float w3x = 0.5f * fx3 - 0.5f * fx2, w3y = 0.5f * fy3 - 0.5f * fy2;
		const float wCx = w1x + w2x, wCy = w1y + w2y;
		const Lin X0 = lin_coord((tcx - 1.0f) * iw, W), XC = lin_coord((tcx + w2x / wCx) * iw, W), X3 = lin_coord((tcx + 2.0f) * iw, W);
		const Lin Y0 = lin_coord((tcy - 1.0f) * ih, H), YC = lin_coord((tcy + w2y / wCy) * ih, H), Y3 = lin_coord((tcy + 2.0f) * ih, H);
		RowPair R0, RC, R3;
		R0.r0 = reinterpret_cast<const uint2*>(row_ptr(A.history_in, Y0.i0, st)); R0.r1 = reinterpret_cast<const uint2*>(row_ptr(A.history_in, Y0.i1, st));
		RC.r0 = reinterpret_cast<const uint2*>(row_ptr(A.history_in, YC.i0, st)); RC.r1 = reinterpret_cast<const uint2*>(row_ptr(A.history_in, YC.i1, st));
		R3.r0 = reinterpret_cast<const uint2*>(row_ptr(A.history_in, Y3.i0, st)); R3.r1 = reinterpret_cast<const uint2*>(row_ptr(A.history_in, Y3.i1, st));
		float4 r = tap4(R0, X0, Y0.a) * w0x * w0y;
		r = r + tap4(R0, XC, Y0.a) * wCx * w0y;
		r = r + tap4(R0, X3, Y0.a) * w3x * w0y;
		r = r + tap4(RC, X0, YC.a) * w0x * wCy;
		r = r + tap4(RC, XC, YC.a) * wCx * wCy;
		r = r + tap4(RC, X3, YC.a) * w3x * wCy;
		r = r + tap4(R3, X0, Y3.a) * w0x * w3y;
		r = r + tap4(R3, XC, Y3.a) * wCx * w3y;
		r = r + tap4(R3, X3, Y3.a) * w3x * w3y;
		historyRaw = r;
	}
	f3 hist = rgb_to_ycocg(xyz(historyRaw));

	float alpha = P.mAlpha;
	bool rejected = false;
	// ---- history rejection (taa.comp:787-823) ----
	if (P.mRejectOutside) {
		if (hu < 0.f || hv < 0.f || hu >= 1.f || hv >= 1.f) { alpha = P.mRejectionAlpha; rejected = true; }
	}
	float writeDynamicMask = 0.f;
	if (P.mDynamicAntiGhosting) {
		const float eps = 1e-5f;
		auto mov = [&](float s, float t) {
			float4 q = tex_rgba16f(A.velocity, W, H, s, t, st);
			return (fabsf(q.x) > eps || fabsf(q.y) > eps) && (fabsf(q.w) >= 0.5f);
		};
		const bool movC = (fabsf(vel.x) > eps || fabsf(vel.y) > eps) && (fabsf(vel.w) >= 0.5f);  // the same sample as getHistoryPosition's
		const bool movement = movC || mov(u + invw * -1.f, v + invh * 0.f) || mov(u + invw * 1.f, v + invh * 0.f) ||
		                      mov(u + invw * 0.f, v + invh * -1.f) || mov(u + invw * 0.f, v + invh * 1.f);
		if (!movement && historyRaw.w > 0.0f) rejected = true;
		writeDynamicMask = movC ? 1.0f : 0.0f;
	}
	if (P.mDepthCulling) {
		int tx = (int)(hu * fW), ty = (int)(hv * fH);
		float hd = fetch_r32f(A.history_depth, W, H, tx, ty, st);
		float depthEpsilon = 0.1f * (1.0f - hd);
		if (fabsf(hd - expectedHistoryDepth) > depthEpsilon) rejected = true;
	}

	// ---- clipAabb(colMin, colMax, vec4(0,0,0,1), vec4(hist,1)) (taa.comp:323-345) ----
	const f3 origHist = hist;
	{
		const float eps = 1e-7f;
		f3 pClip = 0.5f * (colMax + colMin);
		f3 e = 0.5f * (colMax - colMin);
		f3 eClip = mk3(e.x + eps, e.y + eps, e.z + eps);
		f3 vClip = hist - pClip;
		f3 aUnit = abs3(vClip / eClip);
		float maUnit = fmaxf(aUnit.x, fmaxf(aUnit.y, aUnit.z));
		if (maUnit > 1.0f) hist = pClip + vClip / maUnit;
	}
	const f3 rdiff = hist - origHist;
	const bool rectified = fabsf(rdiff.x) > 0.001f || fabsf(rdiff.y) > 0.001f || fabsf(rdiff.z) > 0.001f;

	// ---- blending (taa.comp:848-900); luminance() is .x in YCoCg ----
	if (rejected) {
		alpha = P.mRejectionAlpha;
	} else {
		if (P.mVelBasedAlpha) alpha = fmaxf(alpha, mixf(alpha, P.mVelBasedAlphaMax, clampf(pixelSpeed * P.mVelBasedAlphaFactor, 0.f, 1.f)));
		if (P.mLumaWeightingLottes) {
			float lc = cur.x, lh = hist.x;
			float diff = fabsf(lc - lh) / fmaxf(fmaxf(lc, lh), 0.2f);
			float w = 1.0f - diff;
			alpha = mixf(P.mMaxAlpha, P.mMinAlpha, w * w);
		}
		if (P.mReduceBlendNearClamp) {
			float lmin = colMin.x, lmax = colMax.x, lh = origHist.x;
			float distToClamp = 2.0f * fabsf(fminf(lh - lmin, lmax - lh)) / (lmax - lmin);
			if (lmax - lmin < 0.001f) distToClamp = 1.0f;
			alpha *= clampf(4.0f * distToClamp, 0.f, 1.f);
		}
	}
	float beta = 1.0f;
	if (A.ubo.mResetHistory) { alpha = 1.0f; beta = 1.0f; }
	const float ab = alpha * beta;
	const f3 aa = ycocg_to_rgb(mk3(mixf(hist.x, cur.x, ab), mixf(hist.y, cur.y, ab), mixf(hist.z, cur.z, ab)));
	st_rgba16f(A.history_out, x, y, mk4(aa, writeDynamicMask));
	if (WRITE_SCREEN) st_rgba16f(A.result, x, y, mk4(aa, 1.0f));
	st_r32ui(A.mask, x, y, (rejected ? 1u : 0u) | (rectified ? 2u : 0u) | (2u << 2));
}

// Fix-up pass of the tuned kernels (taa_resolve_stream.cu, taa_resolve_strip.cu): the pixels whose `rectified` predicate
// (taa.comp:845) the re-associated arithmetic could not decide safely are recomputed here with the
// exact arithmetic, from the inputs alone. `list` holds pixels packed as y * out_w + x.
template <bool WRITE_SCREEN>
__global__ void __launch_bounds__(128) taa_resolve_fixup_kernel(const __grid_constant__ ResolveArgs A, const unsigned int* __restrict__ list,
                                                                const unsigned int* __restrict__ count) {
	// launched with programmatic stream serialisation right behind the tuned kernel: set-up overlaps that kernel's tail, and
	// nothing it wrote (list, count, images) is read before it has completed and flushed
	asm volatile("griddepcontrol.wait;" ::: "memory");
	const unsigned int n = *count;
	for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		const unsigned int p = list[i];
		const int y = (int)(p / (unsigned int)A.out_w), x = (int)(p - (unsigned int)y * (unsigned int)A.out_w);
		resolve_pixel_exact_family<WRITE_SCREEN>(A, x, y);
	}
}

// The reference's default settings (BASELINE configs[0]) on the exact arithmetic: CTA = 32 x 8 pixels; the sampled colour of every texel of the
// tile and its one-texel apron is evaluated once into shared memory (1.33 sampler evaluations per pixel instead of 10), then every pixel runs
// resolve_pixel_exact with the default switches folded. Bit-identical to the general kernel by construction (tests/test_parity_gpu.py).
constexpr int SPEC_W = 32, SPEC_H = 8, SPEC_N = (SPEC_W + 2) * (SPEC_H + 2);
__global__ void __launch_bounds__(SPEC_W * SPEC_H) taa_resolve_defaults_kernel(const __grid_constant__ ResolveArgs A) {
	__shared__ float tile[3 * SPEC_N];
	const int x0 = blockIdx.x * SPEC_W, y0 = A.band_y0 + blockIdx.y * SPEC_H;
	const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
	{
		TaaParameters Pl = A.ubo.param[0];
		fold_reference_defaults(Pl);
		Px px(A, Pl);
		const float invw = 1.0f / (float)A.in_w, invh = 1.0f / (float)A.in_h;
		const int yend = min(A.band_y0 + A.band_rows, A.out_h);  // rows below the band / image are nobody's neighbours but the last row's
		for (int k = threadIdx.y * SPEC_W + threadIdx.x; k < SPEC_N; k += SPEC_W * SPEC_H) {
			const int j = k / (SPEC_W + 2), i = k - j * (SPEC_W + 2);
			const int tx = x0 - 1 + i, ty = y0 - 1 + j;
			f3 c = mk3(0.f, 0.f, 0.f);
			if (tx <= A.out_w && ty <= yend) c = px.colour_tap(make_float2(0.f, 0.f), tx, ty, invw, invh);
			tile[k] = c.x; tile[SPEC_N + k] = c.y; tile[2 * SPEC_N + k] = c.z;
		}
	}
	__syncthreads();
	if (x >= A.out_w || y >= A.band_y0 + A.band_rows || y >= A.out_h) return;
	TileTaps taps = {tile, x0, y0, SPEC_W, SPEC_H};
	resolve_pixel_exact<true, 1, TileTaps>(A, x, y, taps);
}

// the call's settings are the reference's default switch pattern (fold_reference_defaults) and nothing else is asked for
bool defaults_kernel_supports(const ResolveArgs& A) {
	static const bool off = [] { const char* v = getenv("TAA_DEFAULTS_KERNEL"); return v && v[0] == '0'; }();  // A/B aid
	const TaaUniforms& U = A.ubo;
	const TaaParameters& P = U.param[0];
	if (off || U.splitScreen || U.mUpsampling || U.mBypassHistoryUpdate) return false;
	if (A.in_w != A.out_w || A.in_h != A.out_h || A.segmask.p) return false;
	return !P.mPassThrough && !P.mUseYCoCg && !P.mVarianceClipping && !P.mShapedNeighbourhood && P.mColorClampingOrClipping == 1 && !P.mUnjitterNeighbourhood &&
	       !P.mUnjitterCurrentSample && !P.mToneMapLumaKaris && !P.mAddNoise && !P.mRayTraceAugment && P.mUseVelocityVectors == 1 && P.mVelocitySampleMode == 0 &&
	       P.mInterpolationMode == 0 && !P.mRejectOutside && !P.mDepthCulling && !P.mDynamicAntiGhosting && !P.mVelBasedAlpha && !P.mLumaWeightingLottes &&
	       !P.mReduceBlendNearClamp;
}

cudaError_t launch_resolve_generic(const ResolveArgs& args, bool allow_specialised, cudaStream_t stream) {
	dim3 block(32, 8);
	dim3 grid((args.out_w + block.x - 1) / block.x, (args.band_rows + block.y - 1) / block.y);
	if (allow_specialised && defaults_kernel_supports(args)) taa_resolve_defaults_kernel<<<grid, block, 0, stream>>>(args);
	else taa_resolve_generic_kernel<<<grid, block, 0, stream>>>(args);
	return cudaGetLastError();
}

cudaError_t launch_resolve_fixup(const ResolveArgs& args, const unsigned int* list, const unsigned int* count, bool write_screen, int num_sms,
                                 cudaStream_t stream) {
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = dim3(num_sms * 8);  // the count lives on the device: surplus CTAs find nothing to do and exit
	cfg.blockDim = dim3(128);
	cfg.dynamicSmemBytes = 0;
	cfg.stream = stream;
	cudaLaunchAttribute attr[1];
	attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
	attr[0].val.programmaticStreamSerializationAllowed = 1;
	cfg.attrs = attr;
	cfg.numAttrs = 1;
	if (write_screen) return cudaLaunchKernelEx(&cfg, taa_resolve_fixup_kernel<true>, args, list, count);
	return cudaLaunchKernelEx(&cfg, taa_resolve_fixup_kernel<false>, args, list, count);
}

}  // namespace taa
