// taa_resolve_tuned.cu — the 32x32-TILE tuned resolve kernel for the BASELINE configs 2-5 family of settings. Since round r01_d the default
// is the strip kernel (taa_resolve_strip.cu, same contract, 1.4-1.9x faster); this one is its A/B partner (TAA_TUNED_VARIANT=tile) and
// the place where the arithmetic contract both share is written down.
// (SURVEY A.8): YCoCg variance clipping, clipAabb rectification, Catmull-Rom history, velocity
// reprojection, optional outside / depth / anti-ghost rejection, velocity alpha, Lottes weighting,
// near-clamp anti-flicker. Everything else runs on the exact generic kernel (taa_dispatch.cu decides).
//
// Structure (one CTA = 32 x 32 output pixels, 8 warps, each thread owns a 4-pixel column strip):
//   phase 0  per-tile tables of the sampler coordinates, which depend on the column or the row only
//   phase 1  the 34 x 34 tile of sampled, YCoCg-converted current colour -> shared memory (each of
//            the 9 neighbourhood taps of taa.comp:204-212 is the centre tap of some pixel)
//   phase 2  per pixel: 3x3 moments from shared memory (row sums slide down the strip), velocity
//            sample, history gather on the 4x4 Catmull-Rom footprint, clip, blend, store
//
// Arithmetic contract (this file is compiled with --fmad=false like the rest; fmaf is explicit):
//   EXACT, in the oracle's operation order: pixel uv, the velocity sample, history uv, every
//     rejection predicate (taa.comp:789-823) and therefore bit 0 of the mask, the history texel
//     coordinates and the sampler's sub-texel offsets.
//   RE-ASSOCIATED / CONTRACTED: colour filtering (neighbourhood moments, history footprint evaluated
//     with separable weights on 4x4 texels instead of 9 bilinear taps, clip, blend). The sampler's
//     fp32 coordinate rounding (a tap "at a texel centre" lands up to ~1e-3 texel beside it and bleeds
//     the neighbour in) is kept for the colour taps and for the dominant centre tap of the history
//     filter, so the result stays within ~1e-5 of the exact kernel, far inside the 2^-10 gate.
//   The `rectified` predicate (taa.comp:845, bit 1 of the mask) compares a colour difference with
//     0.001; pixels where the re-associated value is within FIXUP_BAND_4K (scaled with the frame size) of that threshold are appended
//     to a list and recomputed by the exact arithmetic in a second, tiny launch (launch_resolve_fixup),
//     which makes the mask bit-exact.
#include "taa_tuned_common.cuh"
#include "taa_kernels.h"

namespace taa {

namespace {

using namespace tuned;

constexpr int TW = 32;           // tile width: one warp
constexpr int NWARP = 8;
constexpr int RPT = 4;           // rows per thread
constexpr int TH = NWARP * RPT;  // tile height
constexpr int SW = TW + 2, SH = TH + 2;
constexpr int NT = TW * NWARP;
constexpr int S_CHUNK = 5;        // sampled-colour texels per thread whose loads are in flight together (phase 1)

struct __align__(16) Smem {
	float4 S[SH][SW];
	ColY crow[SH];
	VelY vrow[TH];
	ColX ccol[SW];
	VelX vcol[TW];
	unsigned long long wmask[TH + 4];  // bit c of row r: velocity.w != 0 at texel (x0 - 2 + c, y0 - 2 + r), clamped to the image
};

template <bool REJ, bool ALPHA>
__global__ void __launch_bounds__(NT, 3)
taa_resolve_tuned_kernel(const __grid_constant__ ResolveArgs A, unsigned int* __restrict__ fix_list, unsigned int* __restrict__ fix_count,
                         unsigned int* __restrict__ fix_count_next, const float fix_band) {
	__shared__ Smem sm;
	const TaaParameters& P = A.ubo.param[0];
	unsigned int* st = A.status;
	const int W = A.out_w, H = A.out_h;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int x0 = blockIdx.x * TW;
	const int y0 = A.band_y0 + blockIdx.y * TH;
	const int rows_valid = min(TH, A.band_y0 + A.band_rows - y0);
	const float fW = (float)W, fH = (float)H;
	const float invw = 1.0f / fW, invh = 1.0f / fH;

	if (blockIdx.x == 0 && blockIdx.y == 0 && tid == 0 && fix_count_next) *fix_count_next = 0u;  // the counter the next frame appends to

	// ---- phase 0: coordinate tables ----------------------------------------------------------------
	static_assert(SW + SH + TW + TH <= NT, "one thread per table entry");
	{
		int i = tid;
		if (i < SW) {
			ColX t;
			int m, n;
			colour_axis(x0 - 1 + i, invw, W, m, n, t.p);
			t.m = (unsigned int)m * 8u; t.n = (unsigned int)n * 8u;
			sm.ccol[i] = t;
		} else if ((i -= SW) < rows_valid + 2) {
			ColY t;
			int m, n;
			colour_axis(y0 - 1 + i, invh, H, m, n, t.p);
			t.m = row_off(A.color, m, st); t.n = row_off(A.color, n, st);
			sm.crow[i] = t;
		} else if ((i -= SH) >= 0 && i < TW) {
			const int x = min(x0 + i, W - 1);
			const float u = ((float)x + 0.5f) / fW;  // tc_to_uv, taa.comp:131
			Lin L = lin_coord(u, W);
			VelX t = {(unsigned int)L.i0 * 8u, (unsigned int)L.i1 * 8u, L.a, u};
			sm.vcol[i] = t;
		} else if ((i -= TW) >= 0 && i < rows_valid) {
			const int y = y0 + i;
			const float v = ((float)y + 0.5f) / fH;
			Lin L = lin_coord(v, H);
			VelY t = {row_off(A.velocity, L.i0, st), row_off(A.velocity, L.i1, st), L.a, v};
			sm.vrow[i] = t;
		}
	}
	// Movers (velocity.w != 0, fwd_geometry.frag:289-295) are what the 5-tap anti-ghosting test looks for (taa.comp:796-811). Every tap's
	// bilinear footprint lies inside the 5x5 texels around the pixel; where all of them have w == +-0 the taps return w == 0 exactly.
	if (REJ && P.mDynamicAntiGhosting) {
		for (int r = warp; r < rows_valid + 4; r += NWARP) {
			const uint2* vp = reinterpret_cast<const uint2*>(A.velocity.p + (size_t)row_off(A.velocity, iclamp(y0 - 2 + r, 0, H - 1), st));
			const unsigned int wa = __ldg(vp + iclamp(x0 - 2 + lane, 0, W - 1)).y & 0x7fff0000u;
			const unsigned int wb = lane < 4 ? (__ldg(vp + iclamp(x0 + 30 + lane, 0, W - 1)).y & 0x7fff0000u) : 0u;
			const unsigned int lo = __ballot_sync(0xffffffffu, wa != 0u), hi = __ballot_sync(0xffffffffu, wb != 0u);
			if (lane == 0) sm.wmask[r] = (unsigned long long)lo | ((unsigned long long)hi << 32);
		}
	}
	__syncthreads();

	// ---- phase 1: sampled current colour in YCoCg --------------------------------------------------
	// S = T[m] + px (T[nx] - T[m]) + py (T[ny] - T[m]); the corrections are ~1e-4 of a texel difference and are
	// evaluated in packed fp16 (their own rounding error is below 1e-7); the px*py cross term (< 1e-6) is dropped.
	{
		const int n_s = (rows_valid + 2) * SW;
		for (int base = 0; base < n_s; base += S_CHUNK * NT) {  // loads of a chunk are all in flight before the first is used
			uint2 M[S_CHUNK], NX[S_CHUNK], NY[S_CHUNK];
			float pxs[S_CHUNK], pys[S_CHUNK];
#pragma unroll
			for (int k = 0; k < S_CHUNK; ++k) {
				const int idx = min(base + tid + k * NT, n_s - 1);
				const int r = idx / SW, c = idx - r * SW;
				const ColX cx = sm.ccol[c];
				const ColY cy = sm.crow[r];
				M[k] = __ldg(reinterpret_cast<const uint2*>(A.color.p + (cy.m + cx.m)));
				NX[k] = __ldg(reinterpret_cast<const uint2*>(A.color.p + (cy.m + cx.n)));
				NY[k] = __ldg(reinterpret_cast<const uint2*>(A.color.p + (cy.n + cx.m)));
				pxs[k] = cx.p; pys[k] = cy.p;
			}
#pragma unroll
			for (int k = 0; k < S_CHUNK; ++k) {
				const int idx = base + tid + k * NT;
				if (idx < n_s) {
					const int r = idx / SW, c = idx - r * SW;
					const __half2 px = __float2half2_rn(pxs[k]), py = __float2half2_rn(pys[k]);
					const __half2 m01 = h2(M[k].x), m23 = h2(M[k].y);
					const __half2 c01 = __hfma2(px, __hsub2(h2(NX[k].x), m01), __hmul2(py, __hsub2(h2(NY[k].x), m01)));
					const __half2 c23 = __hfma2(px, __hsub2(h2(NX[k].y), m23), __hmul2(py, __hsub2(h2(NY[k].y), m23)));
					const float2 a01 = __half22float2(m01), b01 = __half22float2(c01);
					const float cr = a01.x + b01.x, cg = a01.y + b01.y, cb = __low2float(m23) + __low2float(c23);
					const float t = cr + cb, hg = 0.5f * cg;
					sm.S[r][c] = make_float4(fmaf(0.25f, t, hg), 0.5f * (cr - cb), fmaf(-0.25f, t, hg), 0.0f);
				}
			}
		}
	}
	__syncthreads();

	// ---- phase 2: one column strip per thread ------------------------------------------------------
	const int x = x0 + lane;
	const bool xvalid = x < W;
	const VelX vc = sm.vcol[lane];
	const float u = vc.u;
	const int r0 = warp * RPT;
	if (r0 >= rows_valid) return;

	// history rows this buffer holds, for the interior test of the gather
	const int hlo = max(0, A.history_in.y0), hhi = min(H - 1, A.history_in.y0 + A.history_in.rows - 1);
	// output pointers of this thread's column, walking down the strip
	const int xs = min(x, W - 1);
	unsigned int o_hist = (unsigned int)(y0 + r0 - A.history_out.y0) * (unsigned int)A.history_out.pitch + (unsigned int)xs * 8u;
	unsigned int o_res = (unsigned int)(y0 + r0 - A.result.y0) * (unsigned int)A.result.pitch + (unsigned int)xs * 8u;
	unsigned int o_mask = (unsigned int)(y0 + r0 - A.mask.y0) * (unsigned int)A.mask.pitch + (unsigned int)xs * 4u;
	unsigned int o_depth = 0u;
	if (REJ) o_depth = row_off(A.depth, y0 + r0, st) + (unsigned int)xs * 4u;

	const float gg = P.mVarClipGamma * P.mVarClipGamma, gg9 = gg * (1.0f / 9.0f);
	// row sums of the first two neighbourhood rows of the strip
	float3 s1a, s2a, s1b, s2b, cur_next;
	{
		const float4 a = sm.S[r0][lane], b = sm.S[r0][lane + 1], c = sm.S[r0][lane + 2];
		s1a = make_float3(a.x + b.x + c.x, a.y + b.y + c.y, a.z + b.z + c.z);
		s2a = make_float3(fmaf(a.x, a.x, fmaf(b.x, b.x, c.x * c.x)), fmaf(a.y, a.y, fmaf(b.y, b.y, c.y * c.y)), fmaf(a.z, a.z, fmaf(b.z, b.z, c.z * c.z)));
	}
	{
		const float4 a = sm.S[r0 + 1][lane], b = sm.S[r0 + 1][lane + 1], c = sm.S[r0 + 1][lane + 2];
		s1b = make_float3(a.x + b.x + c.x, a.y + b.y + c.y, a.z + b.z + c.z);
		s2b = make_float3(fmaf(a.x, a.x, fmaf(b.x, b.x, c.x * c.x)), fmaf(a.y, a.y, fmaf(b.y, b.y, c.y * c.y)), fmaf(a.z, a.z, fmaf(b.z, b.z, c.z * c.z)));
		cur_next = make_float3(b.x, b.y, b.z);
	}

	// Horizontally filtered history rows are shared down the strip: with the same history u (same column, same velocity.x) the
	// x weights are identical, and consecutive pixels' footprints overlap in 3 of their 4 rows. hr0..hr3 hold the filtered rows
	// sh_K .. sh_K + 3 of the previous pixel; if this pixel's footprint starts exactly one row further down, only one row is new.
	bool sh_valid = false;
	float sh_hu = 0.f;
	int sh_K = 0;
	AxisW axs;
	axs.k = 0; axs.w[0] = axs.w[1] = axs.w[2] = axs.w[3] = 0.f;
	HRow hr0 = {0.f, 0.f, 0.f, 0.f, 0u}, hr1 = hr0, hr2 = hr0, hr3 = hr0;
	const unsigned int hpitch = (unsigned int)A.history_in.pitch;

	// velocity footprint of the first pixel of the strip; the next one is requested while the current pixel is filtered
	uint2 vt00, vt10, vt01, vt11;
	{
		const VelY vr = sm.vrow[r0];
		vt00 = __ldg(reinterpret_cast<const uint2*>(A.velocity.p + (vr.o0 + vc.o0))); vt10 = __ldg(reinterpret_cast<const uint2*>(A.velocity.p + (vr.o0 + vc.o1)));
		vt01 = __ldg(reinterpret_cast<const uint2*>(A.velocity.p + (vr.o1 + vc.o0))); vt11 = __ldg(reinterpret_cast<const uint2*>(A.velocity.p + (vr.o1 + vc.o1)));
	}

#pragma unroll
	for (int rr = 0; rr < RPT; ++rr) {
		const int rt = r0 + rr;  // tile row of this pixel
		if (rt >= rows_valid) break;
		const int y = y0 + rt;

		// ---- getHistoryPosition (taa.comp:391-438), exact ----
		const VelY vr = sm.vrow[rt];
		const float v = vr.v;
		float velx, vely, velz = 0.f;
		bool movC = false;
		unsigned long long near_movers = 0ull;  // any velocity.w != 0 among the 5x5 texels around the pixel
		if (REJ && P.mDynamicAntiGhosting)
			near_movers = ((sm.wmask[rt] | sm.wmask[rt + 1] | sm.wmask[rt + 2] | sm.wmask[rt + 3] | sm.wmask[rt + 4]) >> lane) & 0x1full;
		{
			const float2 a00 = __half22float2(h2(vt00.x)), a10 = __half22float2(h2(vt10.x)), a01 = __half22float2(h2(vt01.x)), a11 = __half22float2(h2(vt11.x));
			velx = lerpf(lerpf(a00.x, a10.x, vc.a), lerpf(a01.x, a11.x, vc.a), vr.a);
			vely = lerpf(lerpf(a00.y, a10.y, vc.a), lerpf(a01.y, a11.y, vc.a), vr.a);
			if (REJ) {
				const float2 b00 = __half22float2(h2(vt00.y)), b10 = __half22float2(h2(vt10.y)), b01 = __half22float2(h2(vt01.y)), b11 = __half22float2(h2(vt11.y));
				velz = lerpf(lerpf(b00.x, b10.x, vc.a), lerpf(b01.x, b11.x, vc.a), vr.a);
				if (near_movers) {  // otherwise all four texels carry w == +-0 and the sample's w is exactly 0
					const float velw = lerpf(lerpf(b00.y, b10.y, vc.a), lerpf(b01.y, b11.y, vc.a), vr.a);
					movC = (fabsf(velx) > 1e-5f || fabsf(vely) > 1e-5f) && (fabsf(velw) >= 0.5f);
				}
			}
		}
		const float hu = u - velx, hv = v - vely;

		// ---- history: request the 4x4 Catmull-Rom footprint, then do the neighbourhood statistics while it arrives ----
		const AxisW ay = catmull_axis(hv, fH, invh);
		const bool shared = sh_valid && hu == sh_hu && ay.k - 1 == sh_K + 1 && sh_K + 4 <= hhi;
		uint2 q0, q1, q2, q3;
		if (shared) {  // one new row, sh_K + 4: requested now, filtered after the neighbourhood statistics
			const uint2* hp = reinterpret_cast<const uint2*>(A.history_in.p + ((unsigned int)(sh_K + 4 - A.history_in.y0) * hpitch + (unsigned int)(axs.k - 1) * 8u));
			q0 = __ldg(hp); q1 = __ldg(hp + 1); q2 = __ldg(hp + 2); q3 = __ldg(hp + 3);
		} else {  // (re)start the window with this pixel's four rows
			axs = catmull_axis(hu, fW, invw);
			const bool interior = (unsigned int)(axs.k - 1) <= (unsigned int)(W - 4) && ay.k - 1 >= hlo && ay.k + 2 <= hhi;
			uint2 q[16];
			if (interior) load_history<true>(A.history_in, axs.k, ay.k, W, H, st, q);
			else load_history<false>(A.history_in, axs.k, ay.k, W, H, st, q);
			hr0 = hfilter<REJ>(q[0], q[1], q[2], q[3], axs.w);
			hr1 = hfilter<REJ>(q[4], q[5], q[6], q[7], axs.w);
			hr2 = hfilter<REJ>(q[8], q[9], q[10], q[11], axs.w);
			hr3 = hfilter<REJ>(q[12], q[13], q[14], q[15], axs.w);
			sh_K = ay.k - 1;
			sh_hu = hu;
			sh_valid = interior;
			q0 = q1 = q2 = q3 = make_uint2(0u, 0u);
		}
		const AxisW& ax = axs;
		if (rr + 1 < RPT && rt + 1 < rows_valid) {
			const VelY vn = sm.vrow[rt + 1];
			vt00 = __ldg(reinterpret_cast<const uint2*>(A.velocity.p + (vn.o0 + vc.o0))); vt10 = __ldg(reinterpret_cast<const uint2*>(A.velocity.p + (vn.o0 + vc.o1)));
			vt01 = __ldg(reinterpret_cast<const uint2*>(A.velocity.p + (vn.o1 + vc.o0))); vt11 = __ldg(reinterpret_cast<const uint2*>(A.velocity.p + (vn.o1 + vc.o1)));
		}

		const float3 cur = cur_next;
		float3 s1c, s2c;
		{
			const float4 a = sm.S[rt + 2][lane], b = sm.S[rt + 2][lane + 1], c = sm.S[rt + 2][lane + 2];
			s1c = make_float3(a.x + b.x + c.x, a.y + b.y + c.y, a.z + b.z + c.z);
			s2c = make_float3(fmaf(a.x, a.x, fmaf(b.x, b.x, c.x * c.x)), fmaf(a.y, a.y, fmaf(b.y, b.y, c.y * c.y)), fmaf(a.z, a.z, fmaf(b.z, b.z, c.z * c.z)));
			cur_next = make_float3(b.x, b.y, b.z);
		}
		// ---- variance box (taa.comp:266-277) ----
		// mean = m1 / 9, extent = gamma * sqrt(max(0, m2 / 9 - mean^2)) = sqrt(max(0, (gamma^2 / 9) m2 - gamma^2 mean^2))
		const float ninth = 1.0f / 9.0f;
		const float3 mean = make_float3((s1a.x + s1b.x + s1c.x) * ninth, (s1a.y + s1b.y + s1c.y) * ninth, (s1a.z + s1b.z + s1c.z) * ninth);
		const float3 ext = make_float3(sqrt_approx(fmaxf(0.f, fmaf(-gg * mean.x, mean.x, (s2a.x + s2b.x + s2c.x) * gg9))),
		                               sqrt_approx(fmaxf(0.f, fmaf(-gg * mean.y, mean.y, (s2a.y + s2b.y + s2c.y) * gg9))),
		                               sqrt_approx(fmaxf(0.f, fmaf(-gg * mean.z, mean.z, (s2a.z + s2b.z + s2c.z) * gg9))));
		s1a = s1b; s2a = s2b; s1b = s1c; s2b = s2c;

		if (shared) {
			hr0 = hr1; hr1 = hr2; hr2 = hr3;
			hr3 = hfilter<REJ>(q0, q1, q2, q3, axs.w);
			sh_K += 1;
		}
		Hist hs;
		{
			const HRow &a0 = hr0, &a1 = hr1, &a2 = hr2, &a3 = hr3;
			hs.r = fmaf(ay.w[3], a3.r, fmaf(ay.w[2], a2.r, fmaf(ay.w[1], a1.r, ay.w[0] * a0.r)));
			hs.g = fmaf(ay.w[3], a3.g, fmaf(ay.w[2], a2.g, fmaf(ay.w[1], a1.g, ay.w[0] * a0.g)));
			hs.b = fmaf(ay.w[3], a3.b, fmaf(ay.w[2], a2.b, fmaf(ay.w[1], a1.b, ay.w[0] * a0.b)));
			hs.a = REJ ? fmaf(ay.w[3], a3.a, fmaf(ay.w[2], a2.a, fmaf(ay.w[1], a1.a, ay.w[0] * a0.a))) : 0.f;
			hs.abits = REJ ? ((a0.abits | a1.abits) | (a2.abits | a3.abits)) : 0u;
		}
		float3 hist;  // maybe_rgb_to_ycocg(historyRaw.rgb), taa.comp:769
		{
			const float t = hs.r + hs.b, hg2 = 0.5f * hs.g;
			hist = make_float3(fmaf(0.25f, t, hg2), 0.5f * (hs.r - hs.b), fmaf(-0.25f, t, hg2));
		}

		// ---- rejection (taa.comp:787-823), exact predicates ----
		bool rejected = false, uncertain = fix_band > 3.0e38f;  // TAA_FLAG_FIXUP_ALL
		float writeDynamicMask = 0.f;
		if (REJ) {
			if (P.mRejectOutside && (hu < 0.f || hv < 0.f || hu >= 1.f || hv >= 1.f)) rejected = true;
			if (P.mDynamicAntiGhosting) {
				auto mov = [&](float s, float t) {
					float4 q = tex_rgba16f(A.velocity, W, H, s, t, st);
					return (fabsf(q.x) > 1e-5f || fabsf(q.y) > 1e-5f) && (fabsf(q.w) >= 0.5f);
				};
				bool movement = movC;
				if (!movement && near_movers) movement = mov(u + invw * -1.f, v + invh * 0.f) || mov(u + invw * 1.f, v + invh * 0.f) ||
				                                         mov(u + invw * 0.f, v + invh * -1.f) || mov(u + invw * 0.f, v + invh * 1.f);
				if (!movement) {
					if (hs.a > 0.0f) rejected = true;
					if ((hs.abits & 0x7fff0000u) != 0u) {
						// the sign of a filtered 0/1 mask that cancels to ~0 is not safe under re-association
						if (fabsf(hs.a) < 2.5f * fix_band) uncertain = true;
					} else {
						// all 16 texels carry alpha == 0, so hs.a == 0 exactly; the exact 9-tap sum can still be != 0 when an outer
						// tap's sampler bleed reaches a texel of the 6x6 ring with alpha != 0
						unsigned int ring = 0x7fff0000u;
						if ((unsigned int)(ax.k - 2) <= (unsigned int)(W - 6) && ay.k - 2 >= hlo && ay.k + 3 <= hhi) {
							const unsigned int hp = (unsigned int)A.history_in.pitch;
							const unsigned char* rb = A.history_in.p + ((unsigned int)(ay.k - 2 - A.history_in.y0) * hp + (unsigned int)(ax.k - 2) * 8u + 4u);
							ring = 0u;
#pragma unroll
							for (int j = 0; j < 6; ++j) ring |= __ldg(reinterpret_cast<const unsigned int*>(rb + j * 8)) | __ldg(reinterpret_cast<const unsigned int*>(rb + 5 * hp + j * 8));
#pragma unroll
							for (int i = 1; i < 5; ++i) ring |= __ldg(reinterpret_cast<const unsigned int*>(rb + i * hp)) | __ldg(reinterpret_cast<const unsigned int*>(rb + i * hp + 40));
						}
						if (ring & 0x7fff0000u) uncertain = true;
					}
				}
				writeDynamicMask = movC ? 1.0f : 0.0f;
			}
			if (P.mDepthCulling) {
				const float depth = __ldg(reinterpret_cast<const float*>(A.depth.p + o_depth));
				const float expected = depth - velz;
				const int tx = (int)(hu * fW), ty = (int)(hv * fH);
				const float hd = fetch_r32f(A.history_depth, W, H, tx, ty, st);
				if (fabsf(hd - expected) > 0.1f * (1.0f - hd)) rejected = true;
			}
			o_depth += (unsigned int)A.depth.pitch;
		}

		// ---- clipAabb towards the box centre (taa.comp:323-345) ----
		const float3 vcl = make_float3(hist.x - mean.x, hist.y - mean.y, hist.z - mean.z);
		const float ma = fmaxf(fabsf(vcl.x) * rcp_approx(ext.x + 1e-7f), fmaxf(fabsf(vcl.y) * rcp_approx(ext.y + 1e-7f), fabsf(vcl.z) * rcp_approx(ext.z + 1e-7f)));
		float3 hc = hist;
		bool rectified = false;
		if (ma > 1.0f) {
			const float s = rcp_approx(ma);
			hc = make_float3(fmaf(vcl.x, s, mean.x), fmaf(vcl.y, s, mean.y), fmaf(vcl.z, s, mean.z));
			const float dx = fabsf(hc.x - hist.x), dy = fabsf(hc.y - hist.y), dz = fabsf(hc.z - hist.z);
			// any(greaterThan(abs(diff), 0.001)) == (largest component > 0.001): only the largest component can flip the decision
			const float dmax = fmaxf(dx, fmaxf(dy, dz));
			rectified = dmax > 0.001f;
			// (`rectified` is only reported through the mask: without a mask binding there is nothing to decide exactly; the colours of the
			// two arithmetics agree to ~1e-5 either way)
			if (A.mask.p != nullptr && fabsf(dmax - 0.001f) < fix_band) uncertain = true;
		}

		// ---- blend (taa.comp:848-900) ----
		float alpha = P.mAlpha;
		if (rejected) {
			alpha = P.mRejectionAlpha;
		} else if (ALPHA) {
			if (P.mVelBasedAlpha) {
				const float du = u - hu, dv = v - hv;
				const float speed = sqrt_approx(fmaf(du, du, dv * dv));
				alpha = fmaxf(alpha, mixf(alpha, P.mVelBasedAlphaMax, sat(speed * P.mVelBasedAlphaFactor)));
			}
			if (P.mLumaWeightingLottes) {
				const float lc = cur.x, lh = hc.x;
				const float w = 1.0f - fabsf(lc - lh) * rcp_approx(fmaxf(fmaxf(lc, lh), 0.2f));
				alpha = mixf(P.mMaxAlpha, P.mMinAlpha, w * w);
			}
			if (P.mReduceBlendNearClamp) {
				const float lmin = mean.x - ext.x, lmax = mean.x + ext.x, lh = hist.x;
				float dist = 2.0f * fabsf(fminf(lh - lmin, lmax - lh)) * rcp_approx(lmax - lmin);
				if (lmax - lmin < 0.001f) dist = 1.0f;
				alpha *= sat(4.0f * dist);
			}
		}
		if (A.ubo.mResetHistory) alpha = 1.0f;
		const float om = 1.0f - alpha;
		const float oy = fmaf(hc.x, om, cur.x * alpha), oco = fmaf(hc.y, om, cur.y * alpha), ocg = fmaf(hc.z, om, cur.z * alpha);
		const float tmp = oy - ocg;
		const float outr = tmp + oco, outg = oy + ocg, outb = tmp - oco;

		// ---- stores (taa.comp:908-909, 955-956) ----
		if (xvalid) {
			const __half2 rg = __floats2half2_rn(outr, outg);
			const __half2 bm = __floats2half2_rn(outb, writeDynamicMask), b1 = __floats2half2_rn(outb, 1.0f);
			*reinterpret_cast<uint2*>(A.history_out.p + o_hist) = make_uint2(*reinterpret_cast<const unsigned int*>(&rg), *reinterpret_cast<const unsigned int*>(&bm));
			if (A.result.p) *reinterpret_cast<uint2*>(A.result.p + o_res) = make_uint2(*reinterpret_cast<const unsigned int*>(&rg), *reinterpret_cast<const unsigned int*>(&b1));
			if (A.mask.p) *reinterpret_cast<unsigned int*>(A.mask.p + o_mask) = (rejected ? 1u : 0u) | (rectified ? 2u : 0u) | (2u << 2);
		}
		o_hist += (unsigned int)A.history_out.pitch;
		o_res += (unsigned int)A.result.pitch;
		o_mask += (unsigned int)A.mask.pitch;

		// ---- hand the undecidable pixels to the exact pass (one atomic per warp) ----
		const unsigned int um = __ballot_sync(0xffffffffu, uncertain && xvalid && fix_list != nullptr);
		if (um) {
			const int leader = __ffs(um) - 1;
			unsigned int slot = 0;
			if (lane == leader) slot = atomicAdd(fix_count, (unsigned int)__popc(um));
			slot = __shfl_sync(0xffffffffu, slot, leader);
			if ((um >> lane) & 1u) fix_list[slot + __popc(um & ((1u << lane) - 1u))] = (unsigned int)y * (unsigned int)W + (unsigned int)x;
		}
	}
}

}  // namespace

bool tuned_supports(const ResolveArgs& A) {
	const TaaUniforms& U = A.ubo;
	const TaaParameters& P = U.param[0];
	if (U.splitScreen || U.mUpsampling || U.mBypassHistoryUpdate) return false;
	if (A.in_w != A.out_w || A.in_h != A.out_h) return false;
	if ((long long)A.out_w * A.out_h >= (1ll << 32)) return false;
	const Img* ins[] = {&A.color, &A.depth, &A.velocity, &A.history_in};
	const ImgW* outs[] = {&A.history_out, &A.result, &A.mask};
	for (const Img* i : ins) if (i->p && (long long)i->rows * i->pitch >= (1ll << 32)) return false;  // 32-bit offsets in the kernel
	for (const ImgW* o : outs) if (o->p && ((long long)o->rows * o->pitch >= (1ll << 32) || o->y0 > A.band_y0)) return false;
	if (A.debug.p || A.segmask.p) return false;
	if (P.mPassThrough || !P.mUseYCoCg || P.mShrinkChromaAxis || !P.mVarianceClipping || P.mColorClampingOrClipping != 2) return false;
	if (P.mUnjitterNeighbourhood || P.mUnjitterCurrentSample || P.mToneMapLumaKaris || P.mAddNoise || P.mRayTraceAugment) return false;
	if (P.mUseVelocityVectors != 2 || P.mVelocitySampleMode != 0 || P.mInterpolationMode != 2) return false;
	if (P.mDepthCulling && !A.history_depth.p) return false;
	return true;
}

cudaError_t launch_resolve_tuned(const ResolveArgs& A, unsigned int* fix_list, unsigned int* fix_count, unsigned int* fix_count_next, bool fixup_all,
                                 cudaStream_t stream) {
	const TaaParameters& P = A.ubo.param[0];
	const float band = fixup_all ? INFINITY : FIXUP_BAND_4K * fmaxf(1.0f, fmaxf((float)A.out_w / 3840.0f, (float)A.out_h / 3840.0f));
	const bool rej = P.mDepthCulling || P.mRejectOutside || P.mDynamicAntiGhosting;
	const bool alp = P.mVelBasedAlpha || P.mLumaWeightingLottes || P.mReduceBlendNearClamp;
	dim3 block(TW * NWARP);
	dim3 grid((A.out_w + TW - 1) / TW, (A.band_rows + TH - 1) / TH);
	if (rej) {
		if (alp) taa_resolve_tuned_kernel<true, true><<<grid, block, 0, stream>>>(A, fix_list, fix_count, fix_count_next, band);
		else taa_resolve_tuned_kernel<true, false><<<grid, block, 0, stream>>>(A, fix_list, fix_count, fix_count_next, band);
	} else {
		if (alp) taa_resolve_tuned_kernel<false, true><<<grid, block, 0, stream>>>(A, fix_list, fix_count, fix_count_next, band);
		else taa_resolve_tuned_kernel<false, false><<<grid, block, 0, stream>>>(A, fix_list, fix_count, fix_count_next, band);
	}
	return cudaGetLastError();
}

}  // namespace taa
