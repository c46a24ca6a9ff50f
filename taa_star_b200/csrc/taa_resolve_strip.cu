// taa_resolve_strip.cu — the staged, long-strip tuned resolve kernel (settings family: tuned_supports() in taa_dispatch.cu; arithmetic contract:
// exact coordinates and predicates, re-associated colour filtering, undecidable pixels handed to the exact fix-up list).
// Since round 2 it serves the REJECTION variants of the family (config 3) and images that tensor maps cannot describe; everything else runs on
// the streaming kernel (taa_resolve_stream.cu). The 32x32-tile kernel it replaced in round 1 is gone (its numbers: profiles/r01_b_*, r01_c_*).
//
// What is different is the shape of the work, chosen against what ncu shows for the tile kernel (issue-bound at ~470 warp
// instructions per pixel, and behind that stalled on the latency of first-touch loads):
//   * one CTA = 64 x 32 output pixels, 8 warps laid out 2 x 4; a thread walks down 8 rows of one column, so the per-strip
//     set-up (history window start, first row sums, pointers) is paid once per 8 pixels;
//   * the raw colour and velocity tiles (36 x 68 texels each, 2-texel halo; plus the 32 x 64 depth tile of the rejection variants) are staged into shared memory with
//     cp.async — every first-touch DRAM access of the CTA is in flight at once, none goes through registers, and the coordinate
//     tables are built while they arrive. Phase 1 then samples the colour tile out of shared memory (in place: the sampled
//     YCoCg tile overwrites the raw one) and the per-pixel velocity footprint is four LDS instead of four global loads;
//   * the history window runs one row AHEAD: while pixel y is evaluated from four filtered rows already in registers, the
//     row the next pixel adds is in flight, and is filtered horizontally at the end of the iteration. With coherent motion
//     (same history u down the column, footprint advancing one row per pixel) that is 4 texel loads per pixel instead of 16;
//   * the anti-ghost "ring" test (does any texel of the 6x6 block around the footprint carry alpha?) slides with that
//     window: two extra 4-byte loads per row instead of 20 per pixel;
//   * uniform velocity footprints (all four texels bit-identical and finite: lerp(p, p, w) == p exactly) skip the bilinear arithmetic;
//   * the sampler's sub-texel bleed of the colour taps is applied with mixed-precision FMAs (fma.rn.f32.f16: f16 x f16 + f32);
//   * warps whose tile has UNIFORM motion (every staged velocity texel bit-identical and finite, no mover within two texels; per warp:
//     all history footprints interior and advancing one row per pixel row) take a path (FAST) in which history coordinates and
//     Catmull-Rom weights come from per-row / per-column tables built once per tile — the very same functions of the same inputs, so the
//     results are bit-identical to the general path — and the window never restarts;
//   * the pixels handed to the exact pass are collected per strip and appended with one atomic per warp; without a mask binding and
//     without a fix-up list (DIAG = false) the `rectified` bookkeeping is not evaluated at all.
// Measured history of these choices: DESIGN.md section 5 ("What the profiles say").
#include "taa_tuned_common.cuh"
#include "taa_kernels.h"
#include <cstdlib>

namespace taa {

namespace {

using namespace tuned;

constexpr int TW = 64;   // tile width: two warps
constexpr int WY = 4;    // warps stacked vertically
constexpr int RPT = 8;   // rows per thread
constexpr int TH = WY * RPT;
constexpr int NWARP = 2 * WY;
constexpr int NT = 32 * NWARP;
constexpr int SW = TW + 2, SH = TH + 2;  // sampled-colour tile: 1-texel halo
constexpr int RW = TW + 4;               // raw tiles: columns x0 - 2 .. x0 + 65 (16-byte aligned rows)
constexpr int CRH = TH + 4;              // raw colour rows y0 - 2 .. y0 + 33
constexpr int VRH = TH + 4;              // raw velocity rows y0 - 2 .. y0 + 33 (the mover mask looks two texels out)
constexpr unsigned int RROW = RW * 8u;   // bytes per raw tile row

struct ColT { unsigned int m, n; float p; };                          // colour tap along one axis: byte offsets of the main texel and of its bleeding neighbour in the raw tile, weight
struct __align__(16) VelT { unsigned int o0, o1; float a, c; };       // velocity tap along one axis: byte offsets of the bilinear footprint in the raw tile, weight, the pixel's uv coordinate

// Uniform-motion tiles: Catmull-Rom weights per tile row / tile column (the history position is separable then)
struct __align__(16) AxisT { float w[4]; int k; int tc; unsigned int outside; float h; };  // weights, first-tap texel k, (int)(h * size), h outside [0, 1), h

template <bool REJ>
struct __align__(16) StripSmemT {
	union {
		float4 S[SH][SW];      // sampled current colour in YCoCg (phase 1 output)
		uint2 craw[CRH][RW];   // raw colour texels (cp.async target, phase 1 input)
	} u;
	uint2 vraw[VRH][RW];
	float dtile[REJ ? TH : 1][TW];  // current depth of the tile's pixels (rejection variants with depth culling)
	VelT vrow[TH];
	VelT vcol[TW];
	AxisT roww[TH];
	AxisT colw[TW];
	ColT crow[SH];
	ColT ccol[SW];
	unsigned int colok[2], rowok;         // uniform-motion votes: footprint columns of each tile half interior; bit wy: rows of warp row wy interior and regular
	unsigned long long mbar;              // completion barrier of the bulk copies that stage the raw tiles
	unsigned long long wmask[2][TH + 4];  // [half][r], bit c: velocity.w != 0 at texel (x0 + 32 half - 2 + c, y0 - 2 + r), clamped to the image
};
static_assert(sizeof(StripSmemT<true>) <= 75 * 1024, "three CTAs per SM");

// source row of tile row r (rows gy0 ..): clamped to the image, then silently into the rows the buffer holds
__device__ __forceinline__ const unsigned char* tile_src_row(const Img& im, int gy, int x0, int H) {
	const int ly = iclamp(iclamp(gy, 0, H - 1) - im.y0, 0, im.rows - 1);
	return im.p + (size_t)ly * (size_t)im.pitch + (size_t)(x0 - 2) * 8u;
}
__device__ __forceinline__ bool bulk_ok(const Img& im, int x0, int W) {  // CTA-uniform: no column is clamped, 16-byte aligned rows
	return x0 >= 2 && x0 + TW + 1 <= W - 1 && (((unsigned long long)im.p | (unsigned long long)im.pitch) & 15ull) == 0ull;
}

// Stage rows gy0 .. gy0 + nrow - 1 (clamped to the image, then silently into the rows the buffer holds: the tables of phase 0 report
// the rows that are really used), columns x0 - 2 .. x0 + 65 (clamped to the image) of an 8-byte-texel image into a raw tile.
// A warp takes whole rows: 34 16-byte chunks (interior tiles) or 68 texels (tiles that touch the left / right image border).
__device__ __forceinline__ void stage_tile(const Img& im, uint2 (*dst)[RW], int gy0, int nrow, int x0, int W, int H, int warp, int lane) {
	const bool fast = x0 >= 2 && x0 + TW + 1 <= W - 1 && (((unsigned long long)im.p | (unsigned long long)im.pitch) & 15ull) == 0ull;  // CTA-uniform
	if (fast) {
		const unsigned char* base = im.p + ((size_t)(x0 - 2) * 8u + (size_t)lane * 16u);
		unsigned char* d = reinterpret_cast<unsigned char*>(&dst[warp][0]) + lane * 16;
		const int lo = max(0, im.y0), hi = min(H - 1, im.y0 + im.rows - 1);
		if (gy0 >= lo && gy0 + nrow - 1 <= hi) {  // no row is clamped: walk the pointer
			const unsigned char* src = base + (size_t)(gy0 + warp - im.y0) * (size_t)im.pitch;
			const size_t step = (size_t)NWARP * (size_t)im.pitch;
			for (int r = warp; r < nrow; r += NWARP, src += step, d += NWARP * RROW) {
				cp_async16(d, src);
				if (lane < RW / 2 - 32) cp_async16(d + 512, src + 512);
			}
		} else {
			for (int r = warp; r < nrow; r += NWARP, d += NWARP * RROW) {
				const int ly = iclamp(iclamp(gy0 + r, 0, H - 1) - im.y0, 0, im.rows - 1);
				const unsigned char* src = base + (size_t)ly * (size_t)im.pitch;
				cp_async16(d, src);
				if (lane < RW / 2 - 32) cp_async16(d + 512, src + 512);
			}
		}
	} else {
		const int c0 = iclamp(x0 - 2 + lane, 0, W - 1), c1 = iclamp(x0 + 30 + lane, 0, W - 1), c2 = iclamp(x0 + 62 + lane, 0, W - 1);
		for (int r = warp; r < nrow; r += NWARP) {
			const int ly = iclamp(iclamp(gy0 + r, 0, H - 1) - im.y0, 0, im.rows - 1);
			const unsigned char* src = im.p + (size_t)ly * (size_t)im.pitch;
			cp_async8(&dst[r][lane], src + (size_t)c0 * 8u);
			cp_async8(&dst[r][lane + 32], src + (size_t)c1 * 8u);
			if (lane < RW - 64) cp_async8(&dst[r][lane + 64], src + (size_t)c2 * 8u);
		}
	}
}

// The 4 x 4 (with REJ: 6 x 6 alpha ring) history texels a strip's window starts from, requested early (before the barrier that
// ends phase 1) on uniform-motion warps so that their DRAM latency is hidden behind the tile write-back.
template <bool REJ>
struct Win0 {
	uint2 q[16];
	unsigned int e[8], t[6];  // (unused without REJ)
};
template <bool REJ>
__device__ __forceinline__ void load_win0(Win0<REJ>& w, const unsigned char* p, const unsigned int hpitch) {
#pragma unroll
	for (int i = 0; i < 4; ++i) {
		const uint2* hp = reinterpret_cast<const uint2*>(p + i * hpitch);
		w.q[4 * i] = __ldg(hp); w.q[4 * i + 1] = __ldg(hp + 1); w.q[4 * i + 2] = __ldg(hp + 2); w.q[4 * i + 3] = __ldg(hp + 3);
	}
	if (REJ) {
#pragma unroll
		for (int i = 0; i < 4; ++i) {
			w.e[2 * i] = __ldg(reinterpret_cast<const unsigned int*>(p + i * hpitch - 4));
			w.e[2 * i + 1] = __ldg(reinterpret_cast<const unsigned int*>(p + i * hpitch + 36));
		}
#pragma unroll
		for (int j = 0; j < 6; ++j) w.t[j] = __ldg(reinterpret_cast<const unsigned int*>(p - hpitch - 4 + 8 * j));
	}
}

// ---- phase 2: one column strip per thread ---------------------------------------------------------------------------------
// FAST = the tile's motion is uniform (every staged velocity texel bit-identical, no mover near, all footprints interior, footprint
// rows advancing one per pixel row): history coordinates and Catmull-Rom weights come from the per-row / per-column tables, and the
// window never restarts. The arithmetic is the very same as in the general path (same functions of the same inputs).
// DIAG = the call reports something beyond the colours: the mask is bound or a fix-up list is kept. Without it the `rectified` bookkeeping
// (taa.comp:845 feeds the mask and the debug views only) is not evaluated at all.
template <bool REJ, bool ALPHA, bool FAST, bool DIAG, int UNR>
__device__ __forceinline__ void strip_phase2(const ResolveArgs& A, StripSmemT<REJ>& sm, unsigned int* __restrict__ fix_list, unsigned int* __restrict__ fix_count,
                                             const float fix_band, const int x0, const int y0, const int rows_valid, const int warp, const int lane,
                                             const Win0<REJ>& w0) {
	// Request a uniform strip's first window before the barrier that ends phase 1? Measured on B200: +0.7 % for the rejection variants
	// (128 registers), -1.5 % for the plain ones (the 32 registers it pins do not fit the 80-register cap).
	constexpr bool EARLY_WIN = REJ;
	const TaaParameters& P = A.ubo.param[0];
	unsigned int* st = A.status;
	const int W = A.out_w, H = A.out_h;
	const float fW = (float)W, fH = (float)H;
	const float invw = 1.0f / fW, invh = 1.0f / fH;
	const int wx = warp & 1, wy = warp >> 1;
	const int lx = wx * 32 + lane;  // tile column of this thread
	const int r0 = wy * RPT;
	if (r0 >= rows_valid) return;
	const int nr = min(RPT, rows_valid - r0);
	const int x = x0 + lx;
	const bool xvalid = x < W;
	const VelT vc = sm.vcol[lx];
	const float u = vc.c;
	const unsigned char* vraw = reinterpret_cast<const unsigned char*>(&sm.vraw[0][0]);

	// history rows this buffer holds, for the interior test of the gather
	const int hlo = max(0, A.history_in.y0), hhi = min(H - 1, A.history_in.y0 + A.history_in.rows - 1);
	// output pointers of this thread's column, walking down the strip
	const int xs = min(x, W - 1);
	unsigned int o_hist = (unsigned int)(y0 + r0 - A.history_out.y0) * (unsigned int)A.history_out.pitch + (unsigned int)xs * 8u;
	unsigned int o_res = (unsigned int)(y0 + r0 - A.result.y0) * (unsigned int)A.result.pitch + (unsigned int)xs * 8u;
	unsigned int o_mask = (unsigned int)(y0 + r0 - A.mask.y0) * (unsigned int)A.mask.pitch + (unsigned int)xs * 4u;
	if (REJ && P.mDepthCulling) { row_off(A.depth, y0 + r0, st); row_off(A.depth, y0 + r0 + nr - 1, st); }  // reports rows a band buffer does not hold

	const float gg = P.mVarClipGamma * P.mVarClipGamma, gg9 = gg * (1.0f / 9.0f);
	// row sums of the first two neighbourhood rows of the strip
	float3 s1a, s2a, s1b, s2b, cur_next;
	{
		const float4 a = sm.u.S[r0][lx], b = sm.u.S[r0][lx + 1], c = sm.u.S[r0][lx + 2];
		s1a = make_float3(a.x + b.x + c.x, a.y + b.y + c.y, a.z + b.z + c.z);
		s2a = make_float3(fmaf(a.x, a.x, fmaf(b.x, b.x, c.x * c.x)), fmaf(a.y, a.y, fmaf(b.y, b.y, c.y * c.y)), fmaf(a.z, a.z, fmaf(b.z, b.z, c.z * c.z)));
	}
	{
		const float4 a = sm.u.S[r0 + 1][lx], b = sm.u.S[r0 + 1][lx + 1], c = sm.u.S[r0 + 1][lx + 2];
		s1b = make_float3(a.x + b.x + c.x, a.y + b.y + c.y, a.z + b.z + c.z);
		s2b = make_float3(fmaf(a.x, a.x, fmaf(b.x, b.x, c.x * c.x)), fmaf(a.y, a.y, fmaf(b.y, b.y, c.y * c.y)), fmaf(a.z, a.z, fmaf(b.z, b.z, c.z * c.z)));
		cur_next = make_float3(b.x, b.y, b.z);
	}

	// The history window: hr0..hr3 = horizontally filtered rows sh_K .. sh_K + 3 for history u == sh_hu, complete before a pixel
	// starts; or0..or4 = OR of the alpha words of the 6 texels (columns k-2 .. k+3) of rows sh_K - 1 .. sh_K + 3 (REJ only).
	bool sh_valid = false;
	float sh_hu = 0.f;
	int sh_K = 0;
	AxisW axs;
	axs.k = 0; axs.w[0] = axs.w[1] = axs.w[2] = axs.w[3] = 0.f;
	HRow hr0 = {0.f, 0.f, 0.f, 0.f, 0u}, hr1 = hr0, hr2 = hr0, hr3 = hr0;
	unsigned int or0 = 0u, or1 = 0u, or2 = 0u, or3 = 0u, or4 = 0u;
	const unsigned int hpitch = (unsigned int)A.history_in.pitch;
	const unsigned char* hbase = A.history_in.p;
	unsigned int fixbits = 0u;

	// FAST: constants of the column, the uniform velocity, and the window of the first pixel
	float f_hu = 0.f, f_velz = 0.f;
	int f_tx = 0;
	bool f_outx = false;
	float f_hd = 0.f;        // previous depth at the history position of the pixel (requested one pixel ahead)
	unsigned int hoff = 0u;  // byte offset of the row in flight (first tap column) in the history buffer
	if (FAST) {
		const AxisT cw = sm.colw[lx];
		axs.k = cw.k; axs.w[0] = cw.w[0]; axs.w[1] = cw.w[1]; axs.w[2] = cw.w[2]; axs.w[3] = cw.w[3];
		f_hu = cw.h; f_tx = cw.tc; f_outx = cw.outside != 0u;
		if (REJ) f_velz = __low2float(h2(sm.vraw[0][0].y));
		const int K = sm.roww[r0].k - 1;
		hoff = (unsigned int)(K - A.history_in.y0) * hpitch + (unsigned int)(cw.k - 1) * 8u;
		Win0<REJ> wl;
		if (!EARLY_WIN) load_win0<REJ>(wl, hbase + hoff, hpitch);
		const Win0<REJ>& ww = EARLY_WIN ? w0 : wl;  // EARLY_WIN: requested before the barrier that ends phase 1
		const uint2* q = ww.q;
		if (REJ) {
			const unsigned int *e = ww.e, *t = ww.t;
			or0 = (t[0] | t[1] | t[2]) | (t[3] | t[4] | t[5]);
			or1 = (q[0].y | q[1].y | q[2].y) | (q[3].y | e[0] | e[1]);
			or2 = (q[4].y | q[5].y | q[6].y) | (q[7].y | e[2] | e[3]);
			or3 = (q[8].y | q[9].y | q[10].y) | (q[11].y | e[4] | e[5]);
			or4 = (q[12].y | q[13].y | q[14].y) | (q[15].y | e[6] | e[7]);
		}
		hr0 = hfilter<REJ>(q[0], q[1], q[2], q[3], axs.w);
		hr1 = hfilter<REJ>(q[4], q[5], q[6], q[7], axs.w);
		hr2 = hfilter<REJ>(q[8], q[9], q[10], q[11], axs.w);
		hr3 = hfilter<REJ>(q[12], q[13], q[14], q[15], axs.w);
		hoff += 4u * hpitch;
		if (REJ && P.mDepthCulling) {
			f_hd = fetch_r32f(A.history_depth, W, H, f_tx, sm.roww[r0].tc, st);
			// consume the load HERE: a raw load result carried into the loop makes its first use in the loop body wait on the load's
			// scoreboard in every iteration, and that scoreboard is shared with the look-ahead loads issued just before (measured: 19 %
			// of the kernel's stall samples on that one FADD)
			f_hd = __fadd_rn(f_hd, 0.0f);  // (an arithmetic no-op ptxas keeps: -0 + 0 = +0 is the only change, immaterial to the depth test)
		}
	}

#pragma unroll UNR
	for (int rr = 0; rr < nr; ++rr) {
		const int rt = r0 + rr;  // tile row of this pixel

		float depth = 0.f, f_hd_next = 0.f;
		if (REJ && P.mDepthCulling) depth = sm.dtile[rt][lx];
		float v, hu, hv, velz = 0.f;
		float ayw0, ayw1, ayw2, ayw3;
		int K = 0, f_ty = 0;
		bool f_outy = false;
		bool movC = false;
		unsigned long long near_movers = 0ull;  // any velocity.w != 0 among the 5x5 texels around the pixel
		bool ahead;
		if (FAST) {
			const AxisT rw = sm.roww[rt];
			v = sm.vrow[rt].c;
			hu = f_hu; hv = rw.h; velz = f_velz;
			ayw0 = rw.w[0]; ayw1 = rw.w[1]; ayw2 = rw.w[2]; ayw3 = rw.w[3];
			f_ty = rw.tc; f_outy = rw.outside != 0u;
			ahead = true;  // (the row after the strip's last footprint is requested too: it is in the buffer, see the vote)
			if (REJ && P.mDepthCulling && rr + 1 < nr) f_hd_next = fetch_r32f(A.history_depth, W, H, f_tx, sm.roww[rt + 1].tc, st);
		} else {
			// ---- getHistoryPosition (taa.comp:391-438), exact ----
			const VelT vr = sm.vrow[rt];
			v = vr.c;
			const uint2 vt00 = *reinterpret_cast<const uint2*>(vraw + (vr.o0 + vc.o0)), vt10 = *reinterpret_cast<const uint2*>(vraw + (vr.o0 + vc.o1));
			const uint2 vt01 = *reinterpret_cast<const uint2*>(vraw + (vr.o1 + vc.o0)), vt11 = *reinterpret_cast<const uint2*>(vraw + (vr.o1 + vc.o1));
			float velx, vely;
			if (REJ && P.mDynamicAntiGhosting) {
				const unsigned long long* wm = &sm.wmask[wx][rt];
				near_movers = ((wm[0] | wm[1] | wm[2] | wm[3] | wm[4]) >> lane) & 0x1full;
			}
			{
				bool uni = vt00.x == vt10.x && vt00.x == vt01.x && vt00.x == vt11.x && finite2(vt00.x);
				if (REJ) uni = uni && vt00.y == vt10.y && vt00.y == vt01.y && vt00.y == vt11.y && finite2(vt00.y);
				if (uni) {  // lerp(p, p, w) == p + w * 0 == p for finite p
					const float2 a = __half22float2(h2(vt00.x));
					velx = a.x; vely = a.y;
					if (REJ) {
						const float2 b = __half22float2(h2(vt00.y));
						velz = b.x;
						if (near_movers) movC = (fabsf(velx) > 1e-5f || fabsf(vely) > 1e-5f) && (fabsf(b.y) >= 0.5f);
					}
				} else {
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
			}
			hu = u - velx; hv = v - vely;

			// ---- history window ----
			const AxisW ay = catmull_axis(hv, fH, invh);
			ayw0 = ay.w[0]; ayw1 = ay.w[1]; ayw2 = ay.w[2]; ayw3 = ay.w[3];
			K = ay.k - 1;
			const bool steady = sh_valid && hu == sh_hu && K == sh_K && (!REJ || K + 4 <= hhi);
			if (!steady) {  // (re)start the window with this pixel's rows
				axs = catmull_axis(hu, fW, invw);
				const int kx = axs.k;
				const bool interior = REJ ? ((unsigned int)(kx - 2) <= (unsigned int)(W - 6) && K - 1 >= hlo && K + 4 <= hhi)
				                          : ((unsigned int)(kx - 1) <= (unsigned int)(W - 4) && K >= hlo && K + 3 <= hhi);
				if (interior) {
					const unsigned char* p = hbase + ((unsigned int)(K - A.history_in.y0) * hpitch + (unsigned int)(kx - 1) * 8u);
					uint2 q[16];
#pragma unroll
					for (int i = 0; i < 4; ++i) {
						const uint2* hp = reinterpret_cast<const uint2*>(p + i * hpitch);
						q[4 * i] = __ldg(hp); q[4 * i + 1] = __ldg(hp + 1); q[4 * i + 2] = __ldg(hp + 2); q[4 * i + 3] = __ldg(hp + 3);
					}
					if (REJ) {
						unsigned int e[8], t[6];
#pragma unroll
						for (int i = 0; i < 4; ++i) {
							e[2 * i] = __ldg(reinterpret_cast<const unsigned int*>(p + i * hpitch - 4));
							e[2 * i + 1] = __ldg(reinterpret_cast<const unsigned int*>(p + i * hpitch + 36));
						}
#pragma unroll
						for (int j = 0; j < 6; ++j) t[j] = __ldg(reinterpret_cast<const unsigned int*>(p - hpitch - 4 + 8 * j));
						or0 = (t[0] | t[1] | t[2]) | (t[3] | t[4] | t[5]);
						or1 = (q[0].y | q[1].y | q[2].y) | (q[3].y | e[0] | e[1]);
						or2 = (q[4].y | q[5].y | q[6].y) | (q[7].y | e[2] | e[3]);
						or3 = (q[8].y | q[9].y | q[10].y) | (q[11].y | e[4] | e[5]);
						or4 = (q[12].y | q[13].y | q[14].y) | (q[15].y | e[6] | e[7]);
					}
					hr0 = hfilter<REJ>(q[0], q[1], q[2], q[3], axs.w);
					hr1 = hfilter<REJ>(q[4], q[5], q[6], q[7], axs.w);
					hr2 = hfilter<REJ>(q[8], q[9], q[10], q[11], axs.w);
					hr3 = hfilter<REJ>(q[12], q[13], q[14], q[15], axs.w);
				} else {
					uint2 q[16];
					load_history<false>(A.history_in, kx, ay.k, W, H, st, q);
					hr0 = hfilter<REJ>(q[0], q[1], q[2], q[3], axs.w);
					hr1 = hfilter<REJ>(q[4], q[5], q[6], q[7], axs.w);
					hr2 = hfilter<REJ>(q[8], q[9], q[10], q[11], axs.w);
					hr3 = hfilter<REJ>(q[12], q[13], q[14], q[15], axs.w);
					or0 = 0x7fff0000u;  // the ring is not all in reach: treat it as carrying alpha (the exact pass decides)
				}
				sh_hu = hu;
				sh_valid = interior;
			}
			// the row the next pixel adds (and, with REJ, the bottom row of this pixel's ring): K + 4
			ahead = sh_valid && K + 4 <= hhi && (REJ || rr + 1 < nr);
			hoff = (unsigned int)(K + 4 - A.history_in.y0) * hpitch + (unsigned int)(axs.k - 1) * 8u;
		}
		uint2 q0 = make_uint2(0u, 0u), q1 = q0, q2 = q0, q3 = q0;
		unsigned int e0 = 0u, e1 = 0u;
		if (ahead) {
			const unsigned char* p = hbase + hoff;
			const uint2* hp = reinterpret_cast<const uint2*>(p);
			q0 = __ldg(hp); q1 = __ldg(hp + 1); q2 = __ldg(hp + 2); q3 = __ldg(hp + 3);
			if (REJ) {
				e0 = __ldg(reinterpret_cast<const unsigned int*>(p - 4));
				e1 = __ldg(reinterpret_cast<const unsigned int*>(p + 36));
			}
		}
		if (FAST) {
			hoff += hpitch;
			// hint for the row the NEXT iteration requests: even lanes touch the first, odd lanes the last texel of their four, which
			// together cover every 32-byte sector of the warp's row segment
			asm volatile("prefetch.global.L1 [%0];" ::"l"(hbase + (hoff + ((lane & 1) ? 24u : 0u))));
		}

		const float3 cur = cur_next;
		float3 s1c, s2c;
		{
			const float4 a = sm.u.S[rt + 2][lx], b = sm.u.S[rt + 2][lx + 1], c = sm.u.S[rt + 2][lx + 2];
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

		// ---- the footprint, filtered vertically ----
		float hsr, hsg, hsb, hsa = 0.f;
		hsr = fmaf(ayw3, hr3.r, fmaf(ayw2, hr2.r, fmaf(ayw1, hr1.r, ayw0 * hr0.r)));
		hsg = fmaf(ayw3, hr3.g, fmaf(ayw2, hr2.g, fmaf(ayw1, hr1.g, ayw0 * hr0.g)));
		hsb = fmaf(ayw3, hr3.b, fmaf(ayw2, hr2.b, fmaf(ayw1, hr1.b, ayw0 * hr0.b)));
		if (REJ) hsa = fmaf(ayw3, hr3.a, fmaf(ayw2, hr2.a, fmaf(ayw1, hr1.a, ayw0 * hr0.a)));
		float3 hist;  // maybe_rgb_to_ycocg(historyRaw.rgb), taa.comp:769
		{
			const float t = hsr + hsb, hg2 = 0.5f * hsg;
			hist = make_float3(fmaf(0.25f, t, hg2), 0.5f * (hsr - hsb), fmaf(-0.25f, t, hg2));
		}

		// ---- rejection (taa.comp:787-823), exact predicates ----
		bool rejected = false, uncertain = DIAG && fix_band > 3.0e38f;  // TAA_FLAG_FIXUP_ALL
		bool check_ring = false;
		float writeDynamicMask = 0.f;
		if (REJ) {
			if (FAST) {
				if (P.mRejectOutside && (f_outx || f_outy)) rejected = true;
			} else {
				if (P.mRejectOutside && (hu < 0.f || hv < 0.f || hu >= 1.f || hv >= 1.f)) rejected = true;
			}
			if (P.mDynamicAntiGhosting) {
				bool movement = false;
				if (!FAST) {  // (FAST: no texel within two of the tile carries velocity.w, so none of the five taps sees a mover)
					auto mov = [&](float s, float t) {
						float4 q = tex_rgba16f(A.velocity, W, H, s, t, st);
						return (fabsf(q.x) > 1e-5f || fabsf(q.y) > 1e-5f) && (fabsf(q.w) >= 0.5f);
					};
					movement = movC;
					if (!movement && near_movers) movement = mov(u + invw * -1.f, v + invh * 0.f) || mov(u + invw * 1.f, v + invh * 0.f) ||
					                                         mov(u + invw * 0.f, v + invh * -1.f) || mov(u + invw * 0.f, v + invh * 1.f);
				}
				if (!movement) {
					if (hsa > 0.0f) rejected = true;
					// The sign of a filtered 0/1 mask that cancels to ~0 is not safe under re-association, and an outer tap's sampler bleed
					// can reach a texel of the 6x6 ring: undecided only if some texel of that block carries alpha at all (checked below,
					// when the bottom row of the ring has arrived).
					check_ring = fabsf(hsa) < 2.5f * fix_band;
				}
				writeDynamicMask = movC ? 1.0f : 0.0f;
			}
			if (P.mDepthCulling) {
				const float expected = depth - velz;
				const int tx = FAST ? f_tx : (int)(hu * fW), ty = FAST ? f_ty : (int)(hv * fH);
				const float hd = FAST ? f_hd : fetch_r32f(A.history_depth, W, H, tx, ty, st);
				if (fabsf(hd - expected) > 0.1f * (1.0f - hd)) rejected = true;
			}
		}

		// ---- clipAabb towards the box centre (taa.comp:323-345) ----
		const float3 vcl = make_float3(hist.x - mean.x, hist.y - mean.y, hist.z - mean.z);
		const float ma = fmaxf(fabsf(vcl.x) * rcp_approx(ext.x + 1e-7f), fmaxf(fabsf(vcl.y) * rcp_approx(ext.y + 1e-7f), fabsf(vcl.z) * rcp_approx(ext.z + 1e-7f)));
		float3 hc = hist;
		bool rectified = false;
		if (ma > 1.0f) {
			const float s = rcp_approx(ma);
			hc = make_float3(fmaf(vcl.x, s, mean.x), fmaf(vcl.y, s, mean.y), fmaf(vcl.z, s, mean.z));
			if (DIAG) {
				const float dx = fabsf(hc.x - hist.x), dy = fabsf(hc.y - hist.y), dz = fabsf(hc.z - hist.z);
				// any(greaterThan(abs(diff), 0.001)) == (largest component > 0.001): only the largest component can flip the decision
				const float dmax = fmaxf(dx, fmaxf(dy, dz));
				rectified = dmax > 0.001f;
				// (`rectified` is only reported through the mask: without a mask binding there is nothing to decide exactly; the colours of the
				// two arithmetics agree to ~1e-5 either way)
				if (A.mask.p != nullptr && fabsf(dmax - 0.001f) < fix_band) uncertain = true;
			}
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
			if (DIAG && A.mask.p) *reinterpret_cast<unsigned int*>(A.mask.p + o_mask) = (rejected ? 1u : 0u) | (rectified ? 2u : 0u) | (2u << 2);
		}
		o_hist += (unsigned int)A.history_out.pitch;
		o_res += (unsigned int)A.result.pitch;
		o_mask += (unsigned int)A.mask.pitch;

		// ---- slide the history window: the row that was in flight joins it (garbage if none was: the window is invalid then) ----
		const unsigned int or5 = (q0.y | q1.y | q2.y) | (q3.y | e0 | e1);
		if (REJ && check_ring && (((or0 | or1 | or2) | (or3 | or4 | or5)) & 0x7fff0000u)) uncertain = true;
		hr0 = hr1; hr1 = hr2; hr2 = hr3;
		hr3 = hfilter<REJ>(q0, q1, q2, q3, axs.w);
		if (REJ) { or0 = or1; or1 = or2; or2 = or3; or3 = or4; or4 = or5; }
		if (!FAST) {
			sh_K = K + 1;
			sh_valid = ahead;
		} else {
			f_hd = f_hd_next;
		}
		if (DIAG && uncertain && xvalid) fixbits |= 1u << rr;
	}

	// ---- hand the undecidable pixels of the strip to the exact pass (one atomic per warp) ----
	if (DIAG && fix_list != nullptr && __ballot_sync(0xffffffffu, fixbits != 0u)) {
		const int n = __popc(fixbits);
		int pre = n;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			const int t = __shfl_up_sync(0xffffffffu, pre, d);
			if (lane >= d) pre += t;
		}
		unsigned int base = 0u;
		if (lane == 31) base = atomicAdd(fix_count, (unsigned int)pre);
		base = __shfl_sync(0xffffffffu, base, 31);
		unsigned int slot = base + (unsigned int)(pre - n);
		while (fixbits) {
			const int b = __ffs(fixbits) - 1;
			fixbits &= fixbits - 1u;
			fix_list[slot++] = (unsigned int)(y0 + r0 + b) * (unsigned int)W + (unsigned int)x;
		}
	}
}

// Stage the tile's own depth texels (4 bytes each): rows y0 .. y0 + nrow - 1, columns x0 .. x0 + 63.
__device__ __forceinline__ void stage_depth(const Img& im, float (*dst)[TW], int y0, int nrow, int x0, int W, int warp, int lane) {
	const bool fast = x0 + TW <= W && (((unsigned long long)im.p | (unsigned long long)im.pitch) & 15ull) == 0ull;  // CTA-uniform
	for (int r = warp; r < nrow; r += NWARP) {
		const int ly = iclamp(y0 + r - im.y0, 0, im.rows - 1);
		const unsigned char* src = im.p + (size_t)ly * (size_t)im.pitch;
		if (fast) {
			if (lane < TW / 4) cp_async16(&dst[r][4 * lane], src + (size_t)x0 * 4u + (size_t)lane * 16u);
		} else {
			asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(&dst[r][lane])), "l"(src + (size_t)min(x0 + lane, W - 1) * 4u) : "memory");
			asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(&dst[r][lane + 32])), "l"(src + (size_t)min(x0 + lane + 32, W - 1) * 4u) : "memory");
		}
	}
}

template <bool REJ, bool ALPHA, bool DIAG, int MINB, int UNR>
__global__ void __launch_bounds__(NT, MINB)
taa_resolve_strip_kernel(const __grid_constant__ ResolveArgs A, unsigned int* __restrict__ fix_list, unsigned int* __restrict__ fix_count,
                         unsigned int* __restrict__ fix_count_next, const float fix_band, const bool use_bulk) {
	extern __shared__ __align__(16) unsigned char smem_raw[];
	using StripSmem = StripSmemT<REJ>;
	constexpr bool EARLY_WIN = REJ;  // see strip_phase2
	StripSmem& sm = *reinterpret_cast<StripSmem*>(smem_raw);
	const TaaParameters& P = A.ubo.param[0];
	unsigned int* st = A.status;
	const int W = A.out_w, H = A.out_h;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int x0 = blockIdx.x * TW;
	const int y0 = A.band_y0 + blockIdx.y * TH;
	const int rows_valid = min(TH, A.band_y0 + A.band_rows - y0);
	const float fW = (float)W, fH = (float)H;
	const float invw = 1.0f / fW, invh = 1.0f / fH;

	asm volatile("griddepcontrol.wait;" ::: "memory");  // see launch_variant
	if (blockIdx.x == 0 && blockIdx.y == 0 && tid == 0 && fix_count_next) *fix_count_next = 0u;  // the counter the next frame appends to

	// ---- stage the raw tiles: everything the CTA touches for the first time is requested here, at once ----
	// Default: per-thread cp.async (16 bytes per lane and row). use_bulk (TAA_STRIP_BULK=1): warp 0 issues one bulk (TMA) copy per tile row
	// (70 copies of 544 bytes) on interior tiles instead — measured SLOWER on B200 (0.145 vs 0.137 ms per 4K frame: the CTA then sits a
	// quarter of its stall samples at the mbarrier), so it stays an A/B knob.
	const bool bulk = use_bulk && bulk_ok(A.color, x0, W) && bulk_ok(A.velocity, x0, W);
	if (bulk) {
		if (warp == 0) {
			const int nc = rows_valid + 4, nv = rows_valid + 4;
			if (lane == 0) {
				mbar_init(&sm.mbar, 1u);
				mbar_expect_tx(&sm.mbar, (unsigned int)(nc + nv) * RROW);
			}
			__syncwarp();
			for (int r = lane; r < nc + nv; r += 32) {
				if (r < nc) bulk_row(&sm.u.craw[r][0], tile_src_row(A.color, y0 - 2 + r, x0, H), RROW, &sm.mbar);
				else bulk_row(&sm.vraw[r - nc][0], tile_src_row(A.velocity, y0 - 2 + (r - nc), x0, H), RROW, &sm.mbar);
			}
		}
		if (REJ && P.mDepthCulling) {
			stage_depth(A.depth, sm.dtile, y0, rows_valid, x0, W, warp, lane);
			asm volatile("cp.async.commit_group;" ::: "memory");
		}
	} else {
		stage_tile(A.color, sm.u.craw, y0 - 2, rows_valid + 4, x0, W, H, warp, lane);
		stage_tile(A.velocity, sm.vraw, y0 - 2, rows_valid + 4, x0, W, H, warp, lane);
		if (REJ && P.mDepthCulling) stage_depth(A.depth, sm.dtile, y0, rows_valid, x0, W, warp, lane);
		asm volatile("cp.async.commit_group;" ::: "memory");
	}

	// ---- phase 0: coordinate tables (while the tiles arrive) ---------------------------------------
	for (int i = tid; i < SW + (rows_valid + 2) + TW + rows_valid; i += NT) {
		int j = i;
		if (j < SW) {  // tile column j = image column x0 - 1 + j; raw tile column of image column g (clamped): g - (x0 - 2)
			ColT t;
			int m, n;
			colour_axis(x0 - 1 + j, invw, W, m, n, t.p);
			t.m = (unsigned int)(iclamp(m - (x0 - 2), 0, RW - 1)) * 8u; t.n = (unsigned int)(iclamp(n - (x0 - 2), 0, RW - 1)) * 8u;
			sm.ccol[j] = t;
		} else if ((j -= SW) < rows_valid + 2) {
			ColT t;
			int m, n;
			colour_axis(y0 - 1 + j, invh, H, m, n, t.p);
			row_off(A.color, m, st); row_off(A.color, n, st);  // reports rows a band buffer does not hold
			t.m = (unsigned int)(iclamp(m - (y0 - 2), 0, CRH - 1)) * RROW; t.n = (unsigned int)(iclamp(n - (y0 - 2), 0, CRH - 1)) * RROW;
			sm.crow[j] = t;
		} else if ((j -= rows_valid + 2) < TW) {
			const int x = min(x0 + j, W - 1);
			const float u = ((float)x + 0.5f) / fW;  // tc_to_uv, taa.comp:131
			Lin L = lin_coord(u, W);
			VelT t = {(unsigned int)(iclamp(L.i0 - (x0 - 2), 0, RW - 1)) * 8u, (unsigned int)(iclamp(L.i1 - (x0 - 2), 0, RW - 1)) * 8u, L.a, u};
			sm.vcol[j] = t;
		} else {
			j -= TW;
			const int y = y0 + j;
			const float v = ((float)y + 0.5f) / fH;
			Lin L = lin_coord(v, H);
			row_off(A.velocity, L.i0, st); row_off(A.velocity, L.i1, st);
			VelT t = {(unsigned int)(iclamp(L.i0 - (y0 - 2), 0, VRH - 1)) * RROW, (unsigned int)(iclamp(L.i1 - (y0 - 2), 0, VRH - 1)) * RROW, L.a, v};
			sm.vrow[j] = t;
		}
	}
	asm volatile("cp.async.wait_group 0;" ::: "memory");
	__syncthreads();  // tables, (mbarrier initialised,) cp.async data of all threads
	if (bulk) mbar_wait(&sm.mbar, 0u);

	// ---- is the motion uniform? (votes, taken at the barrier inside phase 1) -------------------------
	// Tile-wide: every staged velocity texel bit-identical and finite, no mover within two texels of the tile. Per warp (32 columns x 8
	// rows): every history footprint interior to the image / band buffer, footprint rows advancing by exactly one per pixel row.
	bool vote = true;
	{
		const uint2 vref = sm.vraw[0][0];
		const uint4* vr4 = reinterpret_cast<const uint4*>(&sm.vraw[0][0]);
		for (int i = tid; i < (rows_valid + 4) * (RW / 2); i += NT) {
			const uint4 t = vr4[i];
			vote = vote && t.x == vref.x && t.z == vref.x && (!REJ || (t.y == vref.y && t.w == vref.y));
		}
		vote = vote && finite2(vref.x) && (!REJ || finite2(vref.y));
		if (REJ && P.mDynamicAntiGhosting) vote = vote && (vref.y & 0x7fff0000u) == 0u;  // no mover within two texels of the tile
		const int hlo = max(0, A.history_in.y0), hhi = min(H - 1, A.history_in.y0 + A.history_in.rows - 1);
		const int ring = REJ ? 1 : 0;
		if (tid < TW + TH) {  // warps 0, 1: the columns of the two tile halves; warp 2: the rows
			const float2 vxy = __half22float2(h2(vref.x));
			AxisT t;
			if (tid < TW) {
				const float hu = sm.vcol[tid].c - vxy.x;
				const AxisW a = catmull_axis(hu, fW, invw);
				t.w[0] = a.w[0]; t.w[1] = a.w[1]; t.w[2] = a.w[2]; t.w[3] = a.w[3];
				t.k = a.k; t.tc = (int)(hu * fW); t.outside = (hu < 0.f || hu >= 1.f) ? 1u : 0u; t.h = hu;
				const bool ok = a.k - 1 - ring >= 0 && a.k + 2 + ring <= W - 1;
				sm.colw[tid] = t;
				const bool all_ok = __all_sync(0xffffffffu, ok);
				if (lane == 0) sm.colok[warp] = all_ok ? 1u : 0u;
			} else {
				const int r = tid - TW;
				bool ok = true;
				if (r < rows_valid) {
					const float hv = sm.vrow[r].c - vxy.y;
					const AxisW a = catmull_axis(hv, fH, invh);
					t.w[0] = a.w[0]; t.w[1] = a.w[1]; t.w[2] = a.w[2]; t.w[3] = a.w[3];
					t.k = a.k; t.tc = (int)(hv * fH); t.outside = (hv < 0.f || hv >= 1.f) ? 1u : 0u; t.h = hv;
					ok = a.k - 1 - ring >= hlo && a.k + 3 <= hhi;  // (row k + 3 is the look-ahead row of the window)
					if (r + 1 < rows_valid && ((r + 1) & (RPT - 1)) != 0) ok = ok && catmull_axis(sm.vrow[r + 1].c - vxy.y, fH, invh).k == a.k + 1;
					sm.roww[r] = t;
				}
				const unsigned int bad = __ballot_sync(0xffffffffu, !ok);
				if (lane == 0) {
					unsigned int m = 0u;
#pragma unroll
					for (int w = 0; w < WY; ++w) m |= ((bad >> (w * RPT)) & ((1u << RPT) - 1u)) == 0u ? (1u << w) : 0u;
					sm.rowok = m;
				}
			}
		}
	}

	bool fast;
	Win0<REJ> w0;
	// ---- phase 1: sampled current colour in YCoCg, in place ----------------------------------------
	// A warp takes whole tile rows (warp, warp + 8, ...); a lane owns tile columns lane and lane + 32; columns 64 and 65 go to the
	// first threads. Everything is read into registers first: the sampled tile overwrites the raw one.
	{
		constexpr int NR = (SH + NWARP - 1) / NWARP;  // rows per warp
		const int nrows = rows_valid + 2;
		const ColT ca = sm.ccol[lane], cb = sm.ccol[lane + 32];
		const __half pxa = __float2half_rn(ca.p), pxb = __float2half_rn(cb.p);
		const unsigned char* cr = reinterpret_cast<const unsigned char*>(&sm.u.craw[0][0]);
		float3 va[NR], vb[NR], ve = make_float3(0.f, 0.f, 0.f);
#pragma unroll
		for (int k = 0; k < NR; ++k) {
			const int r = warp + k * NWARP;
			if (r >= nrows) break;  // warp-uniform
			const ColT cy = sm.crow[r];
			const __half py = __float2half_rn(cy.p);
			va[k] = sample_ycocg(*reinterpret_cast<const uint2*>(cr + (cy.m + ca.m)), *reinterpret_cast<const uint2*>(cr + (cy.m + ca.n)),
			                     *reinterpret_cast<const uint2*>(cr + (cy.n + ca.m)), pxa, py);
			vb[k] = sample_ycocg(*reinterpret_cast<const uint2*>(cr + (cy.m + cb.m)), *reinterpret_cast<const uint2*>(cr + (cy.m + cb.n)),
			                     *reinterpret_cast<const uint2*>(cr + (cy.n + cb.m)), pxb, py);
		}
		if (tid < 2 * nrows) {
			const int r = tid >> 1, c = TW + (tid & 1);
			const ColT cx = sm.ccol[c];
			const ColT cy = sm.crow[r];
			ve = sample_ycocg(*reinterpret_cast<const uint2*>(cr + (cy.m + cx.m)), *reinterpret_cast<const uint2*>(cr + (cy.m + cx.n)),
			                  *reinterpret_cast<const uint2*>(cr + (cy.n + cx.m)), __float2half_rn(cx.p), __float2half_rn(cy.p));
		}
		const bool uni = __syncthreads_and(vote ? 1 : 0) != 0;
		fast = uni && sm.colok[warp & 1] != 0u && ((sm.rowok >> (warp >> 1)) & 1u) != 0u;  // this warp's strips
		const bool cta_fast = uni && sm.colok[0] != 0u && sm.colok[1] != 0u && sm.rowok == (1u << WY) - 1u;
		// Movers (velocity.w != 0, fwd_geometry.frag:289-295) are what the 5-tap anti-ghosting test looks for (taa.comp:796-811). Every tap's
		// bilinear footprint lies inside the 5x5 texels around the pixel; where all of them have w == +-0 the taps return w == 0 exactly.
		// Only the general path needs the mask (a uniform tile has none); it is read after the next barrier.
		if (REJ && !cta_fast && P.mDynamicAntiGhosting) {
			for (int r = warp; r < rows_valid + 4; r += NWARP) {
				row_off(A.velocity, iclamp(y0 - 2 + r, 0, H - 1), st);  // reports rows a band buffer does not hold
				const unsigned int wa = sm.vraw[r][lane].y & 0x7fff0000u, wb = sm.vraw[r][lane + 32].y & 0x7fff0000u;
				const unsigned int wc = lane < 4 ? (sm.vraw[r][lane + 64].y & 0x7fff0000u) : 0u;
				const unsigned int b0 = __ballot_sync(0xffffffffu, wa != 0u), b1 = __ballot_sync(0xffffffffu, wb != 0u), b2 = __ballot_sync(0xffffffffu, wc != 0u);
				if (lane == 0) {
					sm.wmask[0][r] = (unsigned long long)b0 | ((unsigned long long)b1 << 32);
					sm.wmask[1][r] = (unsigned long long)b1 | ((unsigned long long)b2 << 32);
				}
			}
		}
#pragma unroll
		for (int k = 0; k < NR; ++k) {
			const int r = warp + k * NWARP;
			if (r < nrows) {
				sm.u.S[r][lane] = make_float4(va[k].x, va[k].y, va[k].z, 0.f);
				sm.u.S[r][lane + 32] = make_float4(vb[k].x, vb[k].y, vb[k].z, 0.f);
			}
		}
		if (tid < 2 * nrows) sm.u.S[tid >> 1][TW + (tid & 1)] = make_float4(ve.x, ve.y, ve.z, 0.f);
	}
	// the window a uniform-motion strip starts from: requested before the barrier, consumed after it
	if (EARLY_WIN && fast && (warp >> 1) * RPT < rows_valid)
		load_win0<REJ>(w0, A.history_in.p + ((unsigned int)(sm.roww[(warp >> 1) * RPT].k - 1 - A.history_in.y0) * (unsigned int)A.history_in.pitch +
		                                    (unsigned int)(sm.colw[(warp & 1) * 32 + lane].k - 1) * 8u), (unsigned int)A.history_in.pitch);
	__syncthreads();
	// keep the compiler from consuming (= waiting for) the window texels before the barrier
#pragma unroll
	for (int i = 0; i < 16; i += 4)
		asm volatile("" : "+r"(w0.q[i].x), "+r"(w0.q[i].y), "+r"(w0.q[i + 1].x), "+r"(w0.q[i + 1].y), "+r"(w0.q[i + 2].x), "+r"(w0.q[i + 2].y), "+r"(w0.q[i + 3].x),
		             "+r"(w0.q[i + 3].y));
	if (REJ) {
		asm volatile("" : "+r"(w0.e[0]), "+r"(w0.e[1]), "+r"(w0.e[2]), "+r"(w0.e[3]), "+r"(w0.e[4]), "+r"(w0.e[5]), "+r"(w0.e[6]), "+r"(w0.e[7]));
		asm volatile("" : "+r"(w0.t[0]), "+r"(w0.t[1]), "+r"(w0.t[2]), "+r"(w0.t[3]), "+r"(w0.t[4]), "+r"(w0.t[5]));
	}

	if (fast) strip_phase2<REJ, ALPHA, true, DIAG, UNR>(A, sm, fix_list, fix_count, fix_band, x0, y0, rows_valid, warp, lane, w0);
	else strip_phase2<REJ, ALPHA, false, DIAG, UNR>(A, sm, fix_list, fix_count, fix_band, x0, y0, rows_valid, warp, lane, w0);
}

template <bool REJ, bool ALPHA, bool DIAG, int MINB, int UNR>
cudaError_t launch_variant(const ResolveArgs& A, unsigned int* fix_list, unsigned int* fix_count, unsigned int* fix_count_next, float band, cudaStream_t stream) {
	using StripSmem = StripSmemT<REJ>;
	auto kern = taa_resolve_strip_kernel<REJ, ALPHA, DIAG, MINB, UNR>;
	static bool configured[64] = {false};  // per variant and per device: the attribute belongs to the (function, device) pair
	int dev = 0;
	cudaGetDevice(&dev);
	dev &= 63;
	if (!configured[dev]) {
		cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(StripSmem));
		if (e != cudaSuccess) return e;
		configured[dev] = true;
	}
	dim3 grid((A.out_w + TW - 1) / TW, (A.band_rows + TH - 1) / TH);
	static const bool use_bulk = [] { const char* v = getenv("TAA_STRIP_BULK"); return v && v[0] == '1'; }();
	// programmatic stream serialisation: the launch set-up overlaps the tail of whatever kernel precedes it in the stream (normally the
	// previous frame's fix-up pass); the kernel itself waits for that kernel's completion before it touches memory
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = grid;
	cfg.blockDim = dim3(NT);
	cfg.dynamicSmemBytes = sizeof(StripSmem);
	cfg.stream = stream;
	cudaLaunchAttribute attr[1];
	attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
	attr[0].val.programmaticStreamSerializationAllowed = 1;
	cfg.attrs = attr;
	cfg.numAttrs = 1;
	return cudaLaunchKernelEx(&cfg, kern, A, fix_list, fix_count, fix_count_next, band, use_bulk);
}

template <int MINB, int UNR>
cudaError_t launch_minb(const ResolveArgs& A, unsigned int* fix_list, unsigned int* fix_count, unsigned int* fix_count_next, float band, cudaStream_t stream) {
	const TaaParameters& P = A.ubo.param[0];
	const bool rej = P.mDepthCulling || P.mRejectOutside || P.mDynamicAntiGhosting;
	const bool alp = P.mVelBasedAlpha || P.mLumaWeightingLottes || P.mReduceBlendNearClamp;
	const bool diag = fix_list != nullptr || A.mask.p != nullptr;  // (the rejection variants are built with DIAG only)
	if (rej) return alp ? launch_variant<true, true, true, MINB, UNR>(A, fix_list, fix_count, fix_count_next, band, stream)
	                    : launch_variant<true, false, true, MINB, UNR>(A, fix_list, fix_count, fix_count_next, band, stream);
	if (diag) return alp ? launch_variant<false, true, true, MINB, UNR>(A, fix_list, fix_count, fix_count_next, band, stream)
	                     : launch_variant<false, false, true, MINB, UNR>(A, fix_list, fix_count, fix_count_next, band, stream);
	return alp ? launch_variant<false, true, false, MINB, UNR>(A, fix_list, fix_count, fix_count_next, band, stream)
	           : launch_variant<false, false, false, MINB, UNR>(A, fix_list, fix_count, fix_count_next, band, stream);
}

}  // namespace

cudaError_t launch_resolve_strip(const ResolveArgs& A, unsigned int* fix_list, unsigned int* fix_count, unsigned int* fix_count_next, bool fixup_all,
                                 cudaStream_t stream) {
	const float band = fixup_all ? INFINITY : FIXUP_BAND_4K * fmaxf(1.0f, fmaxf((float)A.out_w / 3840.0f, (float)A.out_h / 3840.0f));
	// Defaults from the A/B on B200 (scripts/strip_ab.sh): the strip loop not unrolled (no spills at the 80-register cap); three CTAs/SM
	// for the plain variants, two (128 registers) for the rejection variants, whose window state does not fit 80 registers.
	const TaaParameters& P = A.ubo.param[0];
	const bool rej = P.mDepthCulling || P.mRejectOutside || P.mDynamicAntiGhosting;
	static const int minb_env = [] { const char* s = getenv("TAA_STRIP_MINB"); return s ? atoi(s) : 0; }();  // tuning aids
	static const int unr = [] { const char* s = getenv("TAA_STRIP_UNROLL"); return s ? atoi(s) : 1; }();
	const int minb = minb_env ? minb_env : (rej ? 2 : 3);
#define TAA_STRIP_GO(MB, UN) return launch_minb<MB, UN>(A, fix_list, fix_count, fix_count_next, band, stream)
	if (minb == 2) { if (unr == 4) TAA_STRIP_GO(2, 4); TAA_STRIP_GO(2, 1); }
	if (unr == 4) TAA_STRIP_GO(3, 4);
	TAA_STRIP_GO(3, 1);
#undef TAA_STRIP_GO
}

}  // namespace taa
