// taa_resolve_stream.cu — the streaming resolve kernel: the default for the BASELINE configs 2-5 family since round 2
// (same settings family and the same arithmetic contract as taa_resolve_strip.cu: exact coordinates and predicates, re-associated colour
// filtering, undecidable pixels handed to the exact fix-up pass).
//
// Shape of the work (chosen against the ncu captures of the strip kernel: issue-bound at 353 thread-instructions per pixel, three
// block-wide barriers and an un-overlapped staging -> phase 1 -> phase 2 sequence per tile):
//   * A WARP is the unit of work, there is no block-wide barrier anywhere. A warp owns a strip of 62 output columns and walks down R rows.
//     A lane owns two ADJACENT columns (c0 even): one 16-byte store per output image and row, the partner column's values are already in
//     registers, and only one neighbour per side comes over a warp shuffle. 64 sampled columns give 62 outputs (the two outermost columns
//     have no neighbour in the warp).
//   * The raw colour, velocity (and depth) rows stream through a per-warp RING in shared memory, filled by 2-D tensor-map TMA boxes
//     (cp.async.bulk.tensor.2d, 72 texels x 2 rows per box) that complete on per-slot mbarriers; the warp itself re-arms a slot as soon as
//     it has consumed it, so the rows of the next steps are in flight while the current ones are evaluated. TMA zero-fills what lies
//     outside the image; the sampler's clamp-to-edge is applied through the coordinate tables (they never point at a filled texel).
//   * Nothing is staged twice: the sampled current colour (taa.comp:207, sampler bleed kept) is evaluated once per texel straight out of
//     the ring, converted to YCoCg and folded into ROLLING 3-row sums of the variance box (taa.comp:266-277); there is no sampled tile.
//   * Uniform motion (every velocity texel a strip touches is bit-identical and finite; checked row by row as the rows arrive): the
//     4 x 4 Catmull-Rom footprints of the two columns share five texels per row, the horizontally filtered rows slide down the strip in
//     registers (one new row per pixel row, requested one row ahead), the per-row weights come from a table the lanes build in parallel
//     once per unit. The footprint start is FORCED to advance by exactly one texel per column / row (the natural floor() can jitter by one
//     where the fractional position is within rounding of 0; the Catmull-Rom weights are continuous there, so the two choices agree to
//     O(eps^2)) — a static camera stays on this path.
//   * Anything else (varying motion, footprints that leave the image, movers) takes the general rows: exact bilinear velocity sample out
//     of the ring, per-pixel 4 x 4 gather through L1. A unit falls back from uniform to general rows one way, at a two-row step.
// Rows per unit R and the grid are chosen on the host so that the units fill whole waves of resident warps. Units stay SHORT (<= 30 rows, about
// three waves per 4K frame): units differ in cost by a factor of ~4 (a unit that meets a mover's edge finishes on the general rows), and the
// hardware's CTA dispatch is the only load balancer. Measured on B200 with 128-entry row tables (4K, config 2): R = 26: 0.104 ms, 39: 0.129,
// 52: 0.133, 78 (one wave): 0.183 ms with the SMs idle 56 % of the time behind the few general units.
#include "taa_tuned_common.cuh"
#include "taa_kernels.h"
#include <cuda.h>
#include <cstdlib>
#include <cstring>
#include <type_traits>

namespace taa {

namespace {

using namespace tuned;

constexpr int OWS = 62;                     // output columns per warp strip (60 with a fused sharpening epilogue: it needs resolved neighbours)
constexpr int RWT = 72;                     // ring row: image columns Xb .. Xb + 71, Xb = 62 s - 4 (one TMA box row, 576 bytes)
constexpr unsigned int ROWB = RWT * 8u;
constexpr int SROWS = 2;                    // rows per slot = per TMA box
constexpr int RMAX = 30;                    // output rows per unit: one 32-entry row table covers rows Y0 - 1 .. Y0 + 30
#ifndef TAA_STREAM_NWARP
#define TAA_STREAM_NWARP 2
#endif
constexpr int NWARP = TAA_STREAM_NWARP;     // warps (independent strips) per CTA; MINB below counts PAIRS of warps per SM. (One warp per CTA frees a slot the
                                            // moment its strip is done, but measured slower on B200: 0.1025 against 0.0990 ms, 4K pan + mover.)
constexpr int PFD = 3;                      // L2 prefetch distance of the history rows, in rows beyond the one requested (0: off)
constexpr int DW = 80;                      // depth ring row: columns Xd .. Xd + 79, Xd = Xs rounded down to a multiple of four (a box starts on 16 bytes)
constexpr unsigned int DROWB = DW * 4u;

// ---- slow units first --------------------------------------------------------------------------------------------------------------------
// Units differ in cost by a factor of ~3 (a unit that meets a mover's edge finishes on the general rows) and the hardware dispatches CTAs in
// index order: slow units that start in the last wave are the tail of the launch (measured: 4K pan + one mover, 78 of 5208 units take 70-88 us
// against 27 us; the frame ends at 108 us with most SMs idle from 87 us on). TAA is temporal: a unit that was slow in the previous frame is
// very likely slow in this one. Every CTA that left the uniform rows appends itself to a hint buffer of the context; the next call launches
// HINT_N extra CTAs in FRONT of the grid that take those units first (a regular CTA whose unit a front CTA has taken exits at once).
// Correctness does not depend on what the buffer holds: both sides evaluate the same predicate on data that is read-only during the launch
// (list[flag[u] - 1] == u, flag[u] - 1 < count), so every unit is resolved exactly once whatever the hints say. Three buffers rotate: a call
// reads one, fills the next and clears the third (calls on one context are stream-ordered, as the fix-up counters already require).
constexpr unsigned int HINT_N = 512u;          // front CTAs = list entries
constexpr unsigned int HINT_FLAGS = 32768u;    // CTAs per launch the flags cover (more: hints off)
constexpr unsigned int HINT_WORDS = 2u + HINT_N + HINT_FLAGS;  // count, geometry signature, list, flag per CTA (= list slot + 1)

template <bool REJ>
struct Cfg {
	static constexpr int LOOK = REJ ? 1 : 0;        // extra slots of look-ahead: the rejection variants look two velocity rows further down
	static constexpr int NSLOT = REJ ? 5 : 4;       // ring slots: 2 (3) live + 2 in flight
	static constexpr int NR = NSLOT * SROWS;        // ring rows
};

template <bool REJ>
struct __align__(128) WarpSmem {
	unsigned char craw[Cfg<REJ>::NR * ROWB];          // colour rows:   ring row (g - (Y0 - 2)) % NR holds image row g
	unsigned char vraw[Cfg<REJ>::NR * ROWB];          // velocity rows: ring row (g - (Y0 - 3)) % NR holds image row g
	unsigned char dt[REJ ? Cfg<REJ>::NR * DROWB : 128];  // depth rows, as colour
	uint4 tabA[32];                                   // uniform motion: Catmull-Rom y weights of output row Y0 - 1 + t
	uint4 tabB[32];                                   // output row: velocity footprint rows (ring offsets lo | hi << 16), weight, v, history v
	uint4 tabC[32];                                   // x, y: sampled colour row: ring offsets m | n << 16, bleed weight (half bits); y bit 31, z: output row: hv outside [0, 1), (int)(hv * H);
	                                                  // w: uniform motion: byte offset of the history row requested while output row t - 2 is evaluated
	unsigned long long mbar[8];
};

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, int x, int y, unsigned long long* bar) {
	asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(dst)),
	             "l"(tm), "r"(x), "r"(y), "r"(smem_u32(bar))
	             : "memory");
}

// wait for a phase of a slot barrier; a wait that never ends (a box that cannot arrive) traps instead of hanging the device
__device__ __forceinline__ void slot_wait(unsigned long long* bar, unsigned int parity) {
	unsigned int ok = 0u;
	for (unsigned int spins = 0u;; ++spins) {
		asm volatile(
		    "{\n"
		    ".reg .pred p;\n"
		    "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
		    "selp.u32 %0, 1, 0, p;\n"
		    "}\n"
		    : "=r"(ok)
		    : "r"(smem_u32(bar)), "r"(parity)
		    : "memory");
		if (ok) return;
		if (spins > (1u << 24)) asm volatile("trap;");
	}
}

__device__ __forceinline__ float rsqrt_approx(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

// catmull_axis with the footprint start given (kf = index of the texel left of the sample): the weights of tuned::catmull_axis, which
// are continuous in the position, evaluated for a footprint that may start one texel beside the natural one
__device__ __forceinline__ AxisW catmull_axis_k(float h, float size, float inv, float kf) {
	const float it = h * size;
	const float tc = kf + 0.5f;
	const float f = it - tc;
	const float f2 = f * f;
	const float w0 = f * fmaf(f, fmaf(-0.5f, f, 1.0f), -0.5f);
	const float w1 = fmaf(f2, fmaf(1.5f, f, -2.5f), 1.0f);
	const float w2 = f * fmaf(f, fmaf(-1.5f, f, 2.0f), 0.5f);
	const float w3 = f2 * fmaf(0.5f, f, -0.5f);
	const float wC = w1 + w2;
	const float r = w2 * rcp_approx(wC);
	const float uC = ((tc + r) * inv) * size - 0.5f;
	const float sC = uC - kf;
	const float d1 = sC - 1.0f;
	AxisW o;
	o.w[0] = fmaf(wC, sat(-sC), w0);
	o.w[1] = wC * sat(1.0f - fabsf(sC));
	o.w[2] = wC * sat(1.0f - fabsf(d1));
	o.w[3] = fmaf(wC, sat(d1), w3);
	o.k = (int)fminf(fmaxf(kf, -8.0f), size + 8.0f);
	return o;
}

// global row gy as a row the buffer holds (clamped into it; reported when the row was really asked for)
__device__ __forceinline__ int buf_row(const Img& im, int gy, unsigned int* st, bool used, unsigned int which = 0u) {
	int ly = gy - im.y0;
	if ((unsigned int)ly >= (unsigned int)im.rows) {
		if (st && used) atomicOr(st, 1u | which);  // (bits 4 .. 6 say which input it was: diagnostics)
		ly = ly < 0 ? 0 : im.rows - 1;
	}
	return ly + im.y0;
}
__device__ __forceinline__ unsigned int ring_row(int g, int base, int nr) { return (unsigned int)(((g - base) % nr + nr) % nr); }

struct F3 { float x, y, z; };
struct RowSum { F3 s1, s2; };     // sum and sum of squares of three horizontally adjacent sampled texels
struct HR { float r, g, b, a; };  // one horizontally filtered history row

__device__ __forceinline__ F3 shfl_up3(F3 v) { F3 o; o.x = __shfl_up_sync(0xffffffffu, v.x, 1); o.y = __shfl_up_sync(0xffffffffu, v.y, 1); o.z = __shfl_up_sync(0xffffffffu, v.z, 1); return o; }
__device__ __forceinline__ F3 shfl_down3(F3 v) { F3 o; o.x = __shfl_down_sync(0xffffffffu, v.x, 1); o.y = __shfl_down_sync(0xffffffffu, v.y, 1); o.z = __shfl_down_sync(0xffffffffu, v.z, 1); return o; }

// row sums of the two columns of a lane: column A sees (left neighbour, a, b), column B sees (a, b, right neighbour)
__device__ __forceinline__ void row_sums(const F3 a, const F3 b, RowSum& ra, RowSum& rb) {
	const F3 L = shfl_up3(b), R = shfl_down3(a);
	const float abx = a.x + b.x, aby = a.y + b.y, abz = a.z + b.z;
	const float qx = fmaf(a.x, a.x, b.x * b.x), qy = fmaf(a.y, a.y, b.y * b.y), qz = fmaf(a.z, a.z, b.z * b.z);
	ra.s1.x = abx + L.x; ra.s1.y = aby + L.y; ra.s1.z = abz + L.z;
	rb.s1.x = abx + R.x; rb.s1.y = aby + R.y; rb.s1.z = abz + R.z;
	ra.s2.x = fmaf(L.x, L.x, qx); ra.s2.y = fmaf(L.y, L.y, qy); ra.s2.z = fmaf(L.z, L.z, qz);
	rb.s2.x = fmaf(R.x, R.x, qx); rb.s2.y = fmaf(R.y, R.y, qy); rb.s2.z = fmaf(R.z, R.z, qz);
}

// five adjacent history texels of one row filtered for the two columns of a lane: column A uses texels 0..3, column B texels 1..4
template <bool REJ>
__device__ __forceinline__ void hfilter_pair(const uint2 q0, const uint2 q1, const uint2 q2, const uint2 q3, const uint2 q4, const float (&wa)[4], const float (&wb)[4],
                                             HR& a, HR& b) {
	const float2 e0 = __half22float2(h2(q0.x)), e1 = __half22float2(h2(q1.x)), e2 = __half22float2(h2(q2.x)), e3 = __half22float2(h2(q3.x)), e4 = __half22float2(h2(q4.x));
	a.r = fmaf(wa[3], e3.x, fmaf(wa[2], e2.x, fmaf(wa[1], e1.x, wa[0] * e0.x)));
	a.g = fmaf(wa[3], e3.y, fmaf(wa[2], e2.y, fmaf(wa[1], e1.y, wa[0] * e0.y)));
	b.r = fmaf(wb[3], e4.x, fmaf(wb[2], e3.x, fmaf(wb[1], e2.x, wb[0] * e1.x)));
	b.g = fmaf(wb[3], e4.y, fmaf(wb[2], e3.y, fmaf(wb[1], e2.y, wb[0] * e1.y)));
	if (REJ) {
		const float2 g0 = __half22float2(h2(q0.y)), g1 = __half22float2(h2(q1.y)), g2 = __half22float2(h2(q2.y)), g3 = __half22float2(h2(q3.y)), g4 = __half22float2(h2(q4.y));
		a.b = fmaf(wa[3], g3.x, fmaf(wa[2], g2.x, fmaf(wa[1], g1.x, wa[0] * g0.x)));
		a.a = fmaf(wa[3], g3.y, fmaf(wa[2], g2.y, fmaf(wa[1], g1.y, wa[0] * g0.y)));
		b.b = fmaf(wb[3], g4.x, fmaf(wb[2], g3.x, fmaf(wb[1], g2.x, wb[0] * g1.x)));
		b.a = fmaf(wb[3], g4.y, fmaf(wb[2], g3.y, fmaf(wb[1], g2.y, wb[0] * g1.y)));
	} else {
		const float g0 = __low2float(h2(q0.y)), g1 = __low2float(h2(q1.y)), g2 = __low2float(h2(q2.y)), g3 = __low2float(h2(q3.y)), g4 = __low2float(h2(q4.y));
		a.b = fmaf(wa[3], g3, fmaf(wa[2], g2, fmaf(wa[1], g1, wa[0] * g0)));
		b.b = fmaf(wb[3], g4, fmaf(wb[2], g3, fmaf(wb[1], g2, wb[0] * g1)));
		a.a = 0.f; b.a = 0.f;
	}
}

// The switches of `Parameters` the family leaves open. FX = 0: read from the parameter block (warp-uniform branches); FX = 1: the BASELINE
// config 3 combination, fixed at compile time (reject outside + depth culling + dynamic anti-ghosting, velocity alpha + Lottes luma weighting,
// no near-clamp reduction, no history reset) — the host picks the variant.
template <int FX>
struct Sw {
	__device__ __forceinline__ static bool outside(const TaaParameters& P) { return FX ? true : P.mRejectOutside != 0; }
	__device__ __forceinline__ static bool depth(const TaaParameters& P) { return FX ? true : P.mDepthCulling != 0; }
	__device__ __forceinline__ static bool antighost(const TaaParameters& P) { return FX ? true : P.mDynamicAntiGhosting != 0; }
	__device__ __forceinline__ static bool velalpha(const TaaParameters& P) { return FX ? true : P.mVelBasedAlpha != 0; }
	__device__ __forceinline__ static bool lottes(const TaaParameters& P) { return FX ? true : P.mLumaWeightingLottes != 0; }
	__device__ __forceinline__ static bool nearclamp(const TaaParameters& P) { return FX ? false : P.mReduceBlendNearClamp != 0; }
	__device__ __forceinline__ static bool reset(const TaaUniforms& U) { return FX ? false : U.mResetHistory != 0; }
};

// warp-uniform constants of a dispatch
struct KConst {
	float gg9, rg9, tiny;  // gamma^2 / 9 (at least 1e-12), its inverse square root, 1e-14 / gg9
	float alpha0;          // mAlpha, or 1 with mResetHistory
	bool reset;
};

struct PixOut {
	unsigned int rg, bh, br;  // packed halves: (r, g), (b, history alpha), (b, 1)
	unsigned int mask;
	bool check_ring, uncertain;
};

// Everything of taa.comp::main after the history sample (taa.comp:769-909) for one pixel. cur = sampled current colour (YCoCg), S1 / S2 =
// sum and sum of squares over the 3 x 3 neighbourhood, hs* = filtered history (rgb, alpha), rejected = outside / depth decisions of the
// caller (exact predicates), movement / movC = the 5-tap velocity test and its centre tap, du / dv = uv - history uv.
template <bool REJ, bool ALPHA, bool DIAG, int FX>
__device__ __forceinline__ PixOut resolve_pixel(const ResolveArgs& A, const KConst& kc, const float fix_band, const F3 cur, const F3 S1, const F3 S2, const float hsr,
                                                const float hsg, const float hsb, const float hsa, bool rejected, const bool movement, const bool movC, const float du,
                                                const float dv) {
	const TaaParameters& P = A.ubo.param[0];
	PixOut o;
	o.check_ring = false;
	o.uncertain = DIAG && fix_band > 3.0e38f;  // TAA_FLAG_FIXUP_ALL
	// ---- variance box (taa.comp:266-277): mean = S1 / 9, extent = gamma sqrt(max(0, S2 / 9 - mean^2)) = sqrt(gg9 e2) with e2 = S2 - S1 mean.
	// The clip needs 1 / (extent + 1e-7), taken as rsqrt(max(extent^2, 1e-14)) = rg9 rsqrt(max(e2, tiny)): the clipped colour moves by at most
	// 1e-7 against the reference formulation. (rg9 multiplies the largest ratio once, below.)
	const float ninth = 1.0f / 9.0f;
	const float mx = S1.x * ninth, my = S1.y * ninth, mz = S1.z * ninth;
	const float e2x = fmaf(-mx, S1.x, S2.x), e2y = fmaf(-my, S1.y, S2.y), e2z = fmaf(-mz, S1.z, S2.z);
	const float ix = rsqrt_approx(fmaxf(e2x, kc.tiny)), iy = rsqrt_approx(fmaxf(e2y, kc.tiny)), iz = rsqrt_approx(fmaxf(e2z, kc.tiny));
	// ---- maybe_rgb_to_ycocg(historyRaw.rgb), taa.comp:769
	const float t = hsr + hsb, hg2 = 0.5f * hsg;
	const float hx = fmaf(0.25f, t, hg2), hy = 0.5f * (hsr - hsb), hz = fmaf(-0.25f, t, hg2);
	// ---- rejection by history alpha (taa.comp:796-811)
	float wdm = 0.f;
	if (REJ && Sw<FX>::antighost(P)) {
		if (!movement) {
			if (hsa > 0.0f) rejected = true;
			// the sign of a filtered 0/1 mask that cancels to ~0 is not safe under re-association: undecided if the 6x6 ring carries alpha at all
			o.check_ring = fabsf(hsa) < 2.5f * fix_band;
		}
		wdm = movC ? 1.0f : 0.0f;
	}
	// ---- clipAabb towards the box centre (taa.comp:323-345)
	const float vx = hx - mx, vy = hy - my, vz = hz - mz;
	const float ma = fmaxf(fabsf(vx) * ix, fmaxf(fabsf(vy) * iy, fabsf(vz) * iz)) * kc.rg9;
	const float s = rcp_approx(fmaxf(ma, 1.0f));
	const float cx = fmaf(vx, s, mx), cy = fmaf(vy, s, my), cz = fmaf(vz, s, mz);
	bool rectified = false;
	if (DIAG) {
		// |clipped - history| = |v| (1 - s) per component; any(greaterThan(abs(diff), 0.001)) == (largest component > 0.001)
		const float dmax = fmaxf(fabsf(vx), fmaxf(fabsf(vy), fabsf(vz))) * (1.0f - s);
		rectified = ma > 1.0f && dmax > 0.001f;
		if (A.mask.p != nullptr && ma > 1.0f && fabsf(dmax - 0.001f) < fix_band) o.uncertain = true;
	}
	// ---- blend (taa.comp:848-900)
	float alpha = kc.alpha0;
	const bool reset = FX ? false : kc.reset;
	if (REJ && rejected) alpha = reset ? 1.0f : P.mRejectionAlpha;
	else if (ALPHA && !reset) {
		if (Sw<FX>::velalpha(P)) {
			const float speed = sqrt_approx(fmaf(du, du, dv * dv));
			alpha = fmaxf(alpha, mixf(alpha, P.mVelBasedAlphaMax, sat(speed * P.mVelBasedAlphaFactor)));
		}
		if (Sw<FX>::lottes(P)) {
			const float lc = cur.x, lh = cx;
			const float w = 1.0f - fabsf(lc - lh) * rcp_approx(fmaxf(fmaxf(lc, lh), 0.2f));
			alpha = mixf(P.mMaxAlpha, P.mMinAlpha, w * w);
		}
		if (Sw<FX>::nearclamp(P)) {
			const float ex = (kc.gg9 * fmaxf(e2x, 0.f)) * (ix * kc.rg9);  // the extent of the luma axis: sqrt(gg9 e2)
			const float lmin = mx - ex, lmax = mx + ex, lh = hx;
			float dist = 2.0f * fabsf(fminf(lh - lmin, lmax - lh)) * rcp_approx(lmax - lmin);
			if (lmax - lmin < 0.001f) dist = 1.0f;
			alpha *= sat(4.0f * dist);
		}
	}
	const float om = 1.0f - alpha;
	const float oy = fmaf(cx, om, cur.x * alpha), oco = fmaf(cy, om, cur.y * alpha), ocg = fmaf(cz, om, cur.z * alpha);
	const float tmp = oy - ocg;
	const float outr = tmp + oco, outg = oy + ocg, outb = tmp - oco;
	const __half2 rg = __floats2half2_rn(outr, outg), bm = __floats2half2_rn(outb, wdm);
	o.rg = *reinterpret_cast<const unsigned int*>(&rg);
	o.bh = *reinterpret_cast<const unsigned int*>(&bm);
	o.br = (o.bh & 0xffffu) | 0x3c000000u;
	o.mask = (rejected ? 1u : 0u) | (rectified ? 2u : 0u) | (2u << 2);
	return o;
}

// The general history sample (taa.comp:441-514 through the sampler, re-associated): the 4 x 4 footprint gathered through L1.
// ring = OR of the alpha words of the 6 x 6 texels around the footprint (all ones where they are not all in reach).
template <bool REJ>
__device__ __forceinline__ void gather_history(const ResolveArgs& A, const AxisW& ax, const AxisW& ay, const int W, const int H, const int hlo, const int hhi, float& hsr,
                                               float& hsg, float& hsb, float& hsa, unsigned int& ring) {
	unsigned int* st = A.status;
	const int kx = ax.k, K = ay.k - 1;
	const int rg = REJ ? 1 : 0;
	const bool interior = kx - 1 - rg >= 0 && kx + 2 + rg <= W - 1 && K - rg >= hlo && K + 3 + rg <= hhi;
	uint2 q[16];
	ring = 0u;
	if (interior) {
		const unsigned int hpitch = (unsigned int)A.history_in.pitch;
		const unsigned char* p = A.history_in.p + ((unsigned int)(K - A.history_in.y0) * hpitch + (unsigned int)(kx - 1) * 8u);
#pragma unroll
		for (int i = 0; i < 4; ++i) {
			const uint2* hp = reinterpret_cast<const uint2*>(p + i * hpitch);
			q[4 * i] = __ldg(hp); q[4 * i + 1] = __ldg(hp + 1); q[4 * i + 2] = __ldg(hp + 2); q[4 * i + 3] = __ldg(hp + 3);
		}
		if (REJ) {
#pragma unroll
			for (int i = 0; i < 4; ++i) {
				ring |= __ldg(reinterpret_cast<const unsigned int*>(p + i * hpitch - 4)) | __ldg(reinterpret_cast<const unsigned int*>(p + i * hpitch + 36));
				ring |= (q[4 * i].y | q[4 * i + 1].y) | (q[4 * i + 2].y | q[4 * i + 3].y);
			}
#pragma unroll
			for (int j = 0; j < 6; ++j)
				ring |= __ldg(reinterpret_cast<const unsigned int*>(p - hpitch - 4 + 8 * j)) | __ldg(reinterpret_cast<const unsigned int*>(p + 4 * hpitch - 4 + 8 * j));
		}
	} else {
		load_history<false>(A.history_in, kx, ay.k, W, H, st, q);
		ring = 0x7fff0000u;
	}
	const HRow r0 = hfilter<REJ>(q[0], q[1], q[2], q[3], ax.w), r1 = hfilter<REJ>(q[4], q[5], q[6], q[7], ax.w);
	const HRow r2 = hfilter<REJ>(q[8], q[9], q[10], q[11], ax.w), r3 = hfilter<REJ>(q[12], q[13], q[14], q[15], ax.w);
	hsr = fmaf(ay.w[3], r3.r, fmaf(ay.w[2], r2.r, fmaf(ay.w[1], r1.r, ay.w[0] * r0.r)));
	hsg = fmaf(ay.w[3], r3.g, fmaf(ay.w[2], r2.g, fmaf(ay.w[1], r1.g, ay.w[0] * r0.g)));
	hsb = fmaf(ay.w[3], r3.b, fmaf(ay.w[2], r2.b, fmaf(ay.w[1], r1.b, ay.w[0] * r0.b)));
	hsa = REJ ? fmaf(ay.w[3], r3.a, fmaf(ay.w[2], r2.a, fmaf(ay.w[1], r1.a, ay.w[0] * r0.a))) : 0.f;
}

// ---- the sharpening pass fused into the resolve (EPI = 1: sharpen.comp:23-38, EPI = 2: sharpen_cas.comp + CasFilter, ffx_cas.h:408-537) --------
// The reference writes the resolved image, reads it back through a five-texel plus and writes the sharpened image (taa.hpp:1111-1139): 16 bytes
// of DRAM traffic per pixel that a resolve which still has the neighbours in registers does not need. Values are taken as the rgba16f image
// would have returned them (rounded to fp16), the arithmetic is the one of taa_post.cu (this file is compiled without contraction).
__device__ __forceinline__ float e_lo_sqrt(float a) { return __uint_as_float((__float_as_uint(a) >> 1) + 0x1fbc4639u); }  // ffx_a.h:1455-1457
__device__ __forceinline__ float e_lo_rcp(float a) { return __uint_as_float(0x7ef07ebbu - __float_as_uint(a)); }
__device__ __forceinline__ float e_med_rcp(float a) { float b = __uint_as_float(0x7ef19fffu - __float_as_uint(a)); return b * (-b * a + 2.0f); }
template <int EPI>
__device__ __forceinline__ F3 sharpen_px(const F3 up, const F3 lf, const F3 ce, const F3 rt, const F3 dn, const float k) {
	F3 o;
	if (EPI == 1) {  // C + (4C - L - R - T - B) * f, clamped to [0, 1]
		o.x = clampf(ce.x + ((((4.0f * ce.x - lf.x) - rt.x) - up.x) - dn.x) * k, 0.f, 1.f);
		o.y = clampf(ce.y + ((((4.0f * ce.y - lf.y) - rt.y) - up.y) - dn.y) * k, 0.f, 1.f);
		o.z = clampf(ce.z + ((((4.0f * ce.z - lf.z) - rt.z) - up.z) - dn.z) * k, 0.f, 1.f);
	} else {  // b = up, d = left, e = centre, f = right, h = down; only the green weight survives (ffx_cas.h:514-522)
		const float mn = fminf(fminf(lf.y, fminf(ce.y, rt.y)), fminf(up.y, dn.y));
		const float mx = fmaxf(fmaxf(lf.y, fmaxf(ce.y, rt.y)), fmaxf(up.y, dn.y));
		float amp = clampf(fminf(mn, 1.0f - mx) * e_lo_rcp(mx), 0.f, 1.f);
		amp = e_lo_sqrt(amp);
		const float w = amp * k;
		const float rw = e_med_rcp(1.0f + 4.0f * w);
		o.x = clampf((up.x * w + lf.x * w + rt.x * w + dn.x * w + ce.x) * rw, 0.f, 1.f);
		o.y = clampf((up.y * w + lf.y * w + rt.y * w + dn.y * w + ce.y) * rw, 0.f, 1.f);
		o.z = clampf((up.z * w + lf.z * w + rt.z * w + dn.z * w + ce.z) * rw, 0.f, 1.f);
	}
	return o;
}

#ifdef TAA_STREAM_TRACE
// debugging aid (not in the product build): per unit { start ns, end ns, SM id | first general row << 16, Y0 | strip << 16 }
__device__ unsigned long long g_stream_trace[4 * 16384];
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ unsigned int smid() { unsigned int t; asm volatile("mov.u32 %0, %smid;" : "=r"(t)); return t; }
#endif

// How a launch cuts its band into units: row blocks 0 .. nbig - 1 are R rows high, the blocks after them Rs (<= R) rows: the CTAs are dispatched in
// index order, so the last wave consists of short units and the launch ends on a finer grain (decreasing chunk sizes, as in guided self-scheduling).
struct UnitGeo { int nx, R, nbig, Rs, ny, nbt, nbb; };  // ny row blocks; the first nbt / last nbb of them are the band's boundary blocks (peer mode)

// ---- row bands without a per-frame collective (PEER variants) ------------------------------------------------------------------------------
// A band's first / last `halo` rows of history_out are what the neighbour above / below reads as the halo of its history_in in the next frame.
// The units that resolve those rows (the boundary blocks) store them a second time, straight into the neighbour's buffer (peer-mapped memory
// over NVLink), and then signal a counter in the neighbour's flag block: st ... ; fence.sys ; red.add on the peer word. The same units are the
// only ones of the neighbour's NEXT frame that read those rows: they wait (ld.acquire.sys) for the counter to reach the number of signalling
// warps before they touch the history. Boundary blocks are dispatched first, so that the neighbour's flag is complete long before its next
// frame starts; in the steady state nobody spins. The words are indexed by the parity q of the history buffer being written:
//   flags[2 s + q]     signals received from side s (0: the band above, 1: the band below) for buffers of parity q
//   flags[4 + 2 s + q] this band's own boundary warps of side s that are done with the frame writing parity q
// The last boundary warp of a side resets the word its side waited on (parity q ^ 1) before it signals: the neighbour cannot signal that word
// again before it has seen all of this frame's signals. A neighbour's stores into the parity-q halo can only begin after it has seen this
// band's signals of the frame before, i.e. after every reader of that halo is done (write-after-read).
struct PeerArgs {
	unsigned char* nb_hist[2];   // the neighbours' history buffer of the parity being written, mapped into this process (nullptr: no neighbour)
	long long nb_pitch[2];
	int nb_y0[2];
	unsigned int* nb_flags[2];   // the neighbours' flag blocks
	unsigned int* flags;         // this band's flag block
	unsigned int expect[2];      // signals per frame from side s
	unsigned int mine[2];        // boundary warps of this band on side s
	int halo, q, wait;
};

__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
	unsigned int v;
	asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}

// (the kernel body; the __global__ entry points below differ in how the register budget is given)
template <bool REJ, bool ALPHA, bool DIAG, int FX, int EPI, bool PEER>
__device__ __forceinline__ void stream_body(const ResolveArgs& A, const CUtensorMap& tmC, const CUtensorMap& tmV, const CUtensorMap& tmD, unsigned int* __restrict__ fix_list,
                                            unsigned int* __restrict__ fix_count, unsigned int* __restrict__ fix_count_next, const float fix_band, const UnitGeo& geo,
                                            const unsigned int rt_zero, const unsigned int* __restrict__ hint_in, unsigned int* __restrict__ hint_out,
                                            unsigned int* __restrict__ hint_clear, const unsigned int hint_sig, const PeerArgs& peer) {
	extern __shared__ __align__(128) unsigned char smem_raw[];
	using C = Cfg<REJ>;
	constexpr int NSLOT = C::NSLOT, NR = C::NR, LOOK = C::LOOK;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	WarpSmem<REJ>& sm = reinterpret_cast<WarpSmem<REJ>*>(smem_raw)[warp];
	const TaaParameters& P = A.ubo.param[0];
	unsigned int* st = A.status;
	const int W = A.out_w, H = A.out_h;
	const float fW = (float)W, fH = (float)H;
	const float invw = 1.0f / fW, invh = 1.0f / fH;
	asm volatile("griddepcontrol.wait;" ::: "memory");  // launched with programmatic stream serialisation: nothing is touched before the predecessor is done
	if (blockIdx.x == 0 && threadIdx.x == 0) {
		if (fix_count_next) *fix_count_next = 0u;  // the counter the next frame appends to
		if (hint_out) { hint_out[1] = hint_sig; hint_clear[0] = 0u; }
	}
	// ---- which unit is this CTA's? (1-D grid: [HINT_N front CTAs |] nx x ny regular CTAs) ----
	unsigned int cta = blockIdx.x;
	if (hint_out) {
		const unsigned int n = hint_in[1] == hint_sig ? min(hint_in[0], HINT_N) : 0u;
		if (cta < HINT_N) {  // front: the unit in slot `cta` of the list, if the hints are consistent about it
			if (cta >= n) return;
			const unsigned int u = hint_in[2u + cta];
			if (u >= gridDim.x - HINT_N || hint_in[2u + HINT_N + u] != cta + 1u) return;
			cta = u;
		} else {             // regular: unless a front CTA has taken the unit
			cta -= HINT_N;
			const unsigned int f = hint_in[2u + HINT_N + cta];
			if (f >= 1u && f - 1u < n && hint_in[2u + f - 1u] == cta) return;
		}
		if (threadIdx.x == 0) hint_clear[2u + HINT_N + cta] = 0u;
	}
	const int bx = (int)(cta % (unsigned int)geo.nx);
	int by = (int)(cta / (unsigned int)geo.nx);
	// (dispatch order = top-down, the short units last, measured best: bottom-up 0.1053 ms, short units first then top-down 0.1056, against 0.0979)
	if (PEER) {  // dispatch order: the top boundary blocks, the bottom boundary blocks, then the interior (measured in the single-GPU emulation: their
	             // system-scope fences cost ~14 us of a 55 us band when the boundary blocks run last, ~4 us when they run first)
		if (by >= geo.nbt) by = by < geo.nbt + geo.nbb ? geo.ny - geo.nbb + (by - geo.nbt) : by - geo.nbb;
	}
	const bool side_top = PEER && peer.nb_hist[0] != nullptr && by < geo.nbt, side_bot = PEER && peer.nb_hist[1] != nullptr && by >= geo.ny - geo.nbb;
	const int strip = bx * NWARP + warp;
	// rows this unit owns (writes); with a sharpening epilogue it also resolves the row above and the row below them, whose values the
	// plus-shaped stencil needs (whole frames only: the follow-on passes do not run on bands)
	const int Yo = A.band_y0 + (by < geo.nbig ? by * geo.R : geo.nbig * geo.R + (by - geo.nbig) * geo.Rs);
	const int no = min(by < geo.nbig ? geo.R : geo.Rs, A.band_y0 + A.band_rows - Yo);
	const int Y0 = EPI ? max(Yo - 1, 0) : Yo;
	const int nr = EPI ? min(Yo + no, H - 1) - Y0 + 1 : no;
	constexpr int STEP_X = EPI ? OWS - 2 : OWS;
	const int Xs = strip * STEP_X - 2, Xb = Xs - 2;  // first sampled column, first ring column

	if (Xs + (EPI ? 2 : 1) > W - 1 || no <= 0) return;  // (warp-uniform; there is no block-wide barrier in this kernel)
	if (PEER && peer.wait && (side_top || side_bot)) {
		// the halo rows this unit may read were stored by the neighbour's boundary warps during the previous frame: all of them must have signalled
		if (lane == 0) {
			const long long t0 = clock64();
			for (int sd = 0; sd < 2; ++sd) {
				if (!(sd == 0 ? side_top : side_bot)) continue;
				const unsigned int* f = peer.flags + (2 * sd + (peer.q ^ 1));
				while (ld_acquire_sys(f) < peer.expect[sd]) {
					__nanosleep(64);
					if (clock64() - t0 > (1ll << 35)) { if (st) atomicOr(st, 2u); break; }  // (~17 s: reported as a peer time-out, never a hang)
				}
			}
		}
		__syncwarp();
	}
#ifdef TAA_STREAM_TRACE
	const unsigned long long tr_t0 = gtime();
	int tr_general = 0xffff;
#endif

	const bool use_depth = REJ && Sw<FX>::depth(P);
	const int nslots = (nr + 4 + LOOK + 1) / 2;  // ring rows Y0 - 2 .. Y0 + nr + 1 (+ LOOK) in boxes of two
	const unsigned int slot_bytes = 2u * SROWS * ROWB + (use_depth ? SROWS * DROWB : 0u);
	auto issue = [&](int k) {  // lane 0: arm slot k and request its boxes
		const int pos = k % NSLOT;
		unsigned long long* bar = &sm.mbar[pos];
		mbar_expect_tx(bar, slot_bytes);
		tma_load_2d(sm.craw + pos * (SROWS * ROWB), &tmC, 2 * Xb, Y0 - 2 + SROWS * k - A.color.y0, bar);
		tma_load_2d(sm.vraw + pos * (SROWS * ROWB), &tmV, 2 * Xb, Y0 - 3 + SROWS * k - A.velocity.y0, bar);
		if (use_depth) tma_load_2d(sm.dt + pos * (SROWS * DROWB), &tmD, Xs - (Xs & 3), Y0 - 2 + SROWS * k - A.depth.y0, bar);
	};
	auto wait_slot = [&](int k) { slot_wait(&sm.mbar[k % NSLOT], (unsigned int)((k / NSLOT) & 1)); };

	if (lane == 0) {
#pragma unroll
		for (int p = 0; p < NSLOT; ++p) mbar_init(&sm.mbar[p], 1u);
		asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
		for (int k = 0; k < min(NSLOT, nslots); ++k) issue(k);
	}
	__syncwarp();

	// ---- constants of the lane's two columns c0 = Xs + 2 lane, c1 = c0 + 1 (while the first boxes arrive) ----
	const int c0 = Xs + 2 * lane, c1 = c0 + 1;
	const bool cv0 = c0 >= 0 && c0 < W, cv1 = c1 >= 0 && c1 < W;  // the column exists in the image
	unsigned int cm0, cn0, cm1, cn1;  // colour taps: byte offsets of the main texel and of its bleeding neighbour in a ring row
	__half px0, px1;
	{
		int m, n;
		float p;
		colour_axis(c0, invw, W, m, n, p);
		cm0 = (unsigned int)iclamp(m - Xb, 0, RWT - 1) * 8u; cn0 = (unsigned int)iclamp(n - Xb, 0, RWT - 1) * 8u; px0 = __float2half_rn(p);
		colour_axis(c1, invw, W, m, n, p);
		cm1 = (unsigned int)iclamp(m - Xb, 0, RWT - 1) * 8u; cn1 = (unsigned int)iclamp(n - Xb, 0, RWT - 1) * 8u; px1 = __float2half_rn(p);
	}
	const int x0c = iclamp(c0, 0, W - 1), x1c = iclamp(c1, 0, W - 1);
	const float u0 = ((float)x0c + 0.5f) / fW, u1 = ((float)x1c + 0.5f) / fW;  // tc_to_uv, taa.comp:131
	const bool so0 = lane >= 1 && (!EPI || lane <= 30) && cv0, so1 = lane <= 30 && (!EPI || lane >= 1) && cv1;  // the column is an output of this strip
	const bool al16 = (((unsigned long long)A.history_out.p | (unsigned long long)A.history_out.pitch | (unsigned long long)A.result.p | (unsigned long long)A.result.pitch) & 15ull) == 0ull;
	const bool st16 = so0 && so1 && al16;

	KConst kc;
	{
		kc.gg9 = fmaxf(P.mVarClipGamma * P.mVarClipGamma * (1.0f / 9.0f), 1e-12f);
		kc.rg9 = rsqrt_approx(kc.gg9);
		kc.tiny = 1e-14f * rcp_approx(kc.gg9);
		kc.reset = Sw<FX>::reset(A.ubo);
		kc.alpha0 = kc.reset ? 1.0f : P.mAlpha;
	}
	const int hlo = max(0, A.history_in.y0), hhi = min(H - 1, A.history_in.y0 + A.history_in.rows - 1);
	const unsigned int hpitch = (unsigned int)A.history_in.pitch;
	const unsigned char* hbase = A.history_in.p;
	// While the first boxes are in flight: pull the history rows around the unit's first rows into L2 (where they lie if the motion is small). The
	// first window is requested only after the velocity rows have arrived, i.e. a second memory latency in a row; this makes it an L2 hit.
	if (PFD > 0) {
		const unsigned int xb = ((unsigned int)max(Xb, 0) * 8u) & ~127u;
		for (int q = lane; q < 60; q += 32) {
			const int r = iclamp(Y0 - 3 + q / 6, hlo, hhi);
			const unsigned int off = (unsigned int)(r - A.history_in.y0) * hpitch + min(xb + (unsigned int)(q % 6) * 128u, (unsigned int)(W - 1) * 8u);
			asm volatile("prefetch.global.L2 [%0];" ::"l"(hbase + off));
		}
	}

	// ---- row tables, the part that does not depend on the motion: entry t = image row Y0 - 1 + t ----
	{
		const int g = Y0 - 1 + lane;
		int m, n;
		float p;
		colour_axis(g, invh, H, m, n, p);
		const bool cused = lane <= nr + 1;  // sampled rows Y0 - 1 .. Y0 + nr
		const unsigned int cm = ring_row(buf_row(A.color, m, st, cused, 16u), Y0 - 2, NR) * ROWB, cn = ring_row(buf_row(A.color, n, st, cused, 16u), Y0 - 2, NR) * ROWB;
		const __half ph = __float2half_rn(p);
		sm.tabC[lane] = make_uint4(cm | (cn << 16), (unsigned int)__half_as_ushort(ph), 0u, 0u);
		const float v = ((float)g + 0.5f) / fH;
		const Lin L = lin_coord(v, H);
		const bool vused = lane >= 1 && lane <= nr;  // output rows
		const unsigned int o0 = ring_row(buf_row(A.velocity, L.i0, st, vused, 32u), Y0 - 3, NR) * ROWB, o1 = ring_row(buf_row(A.velocity, L.i1, st, vused, 32u), Y0 - 3, NR) * ROWB;
		sm.tabB[lane] = make_uint4(o0 | (o1 << 16), __float_as_uint(L.a), __float_as_uint(v), 0u);
		if (use_depth && vused) buf_row(A.depth, g, st, true, 64u);
	}
	__syncwarp();

	// ---- the first rows: slots 0, 1 (and 2) ----
	wait_slot(0);
	if (nslots > 1) wait_slot(1);
	if (LOOK && nslots > 2) wait_slot(2);

	// ---- is the motion uniform? ----
	// Reference texel: velocity row Y0 at a column that exists. A row passes if the lane's two texels (and, for the rejection variants, the
	// two texels beside the strip's outputs) are bit-identical to it. Rows outside the image are never sampled as themselves; rows the buffer
	// does not hold end the uniform path (the general rows report them).
	const uint2 vref = *reinterpret_cast<const uint2*>(sm.vraw + ring_row(Y0, Y0 - 3, NR) * ROWB + (unsigned int)(iclamp(Xs + 2, 0, W - 1) - Xb) * 8u);
	bool uni = finite2(vref.x) && (!REJ || finite2(vref.y));
	if (REJ && Sw<FX>::antighost(P)) uni = uni && (vref.y & 0x7fff0000u) == 0u;  // the strip's own motion is not a mover's
	// velocity rows the unit samples that exist in the image must be in the buffer (else: general rows, which report them)
	uni = uni && max(0, Y0 - 1 - LOOK) >= A.velocity.y0 && min(H - 1, Y0 + nr + LOOK) <= A.velocity.y0 + A.velocity.rows - 1;
	unsigned int wrows = 0u;  // bit r: velocity row (newest voted - r) carries velocity.w != 0 somewhere under / beside the strip
	const unsigned char* vown = sm.vraw + (unsigned int)(2 + 2 * lane) * 8u;                                    // the lane's two texels in a ring row
	const int cex = lane == 0 ? Xs - 1 : Xs + 64;                                                              // (rejection variants) the column beside the strip's outputs
	const bool cev = REJ && (lane == 0 || lane == 31) && cex >= 0 && cex < W;
	const unsigned char* vext = sm.vraw + (unsigned int)(iclamp(cex - Xb, 0, RWT - 1)) * 8u;
	auto vote_row = [&](int g) -> bool {  // the lane's verdict on velocity row g (rows outside the image are never sampled as themselves)
		if (g < 0 || g > H - 1) return true;
		const unsigned int ro = ring_row(g, Y0 - 3, NR) * ROWB;
		const uint4 t = *reinterpret_cast<const uint4*>(vown + ro);
		bool ok = (!cv0 || (t.x == vref.x && (!REJ || t.y == vref.y))) && (!cv1 || (t.z == vref.x && (!REJ || t.w == vref.y)));
		if (REJ) {
			unsigned int wb = ((cv0 ? t.y : 0u) | (cv1 ? t.w : 0u)) & 0x7fff0000u;
			if (cev) {
				const uint2 e = *reinterpret_cast<const uint2*>(vext + ro);
				ok = ok && e.x == vref.x && e.y == vref.y;
				wb |= e.y & 0x7fff0000u;
			}
			wrows = (wrows << 1) | (__any_sync(0xffffffffu, wb != 0u) ? 1u : 0u);
		}
		return ok;
	};
	{
		bool ok = vote_row(Y0 - 1) & vote_row(Y0);
		if (REJ) ok = ok & vote_row(Y0 - 2) & vote_row(Y0 + 1) & vote_row(Y0 + 2);
		uni = uni && __all_sync(0xffffffffu, ok);
	}

	// ---- uniform motion: column constants and the motion-dependent part of the row tables ----
	float wA[4] = {0.f, 0.f, 0.f, 0.f}, wB[4] = {0.f, 0.f, 0.f, 0.f};
	float hu0 = 0.f, hu1 = 0.f, f_velz = 0.f;
	int tx0 = 0, tx1 = 0;
	bool outx0 = false, outx1 = false;
	const unsigned int hmax_off = (unsigned int)(hhi - A.history_in.y0) * hpitch;  // last history row that is both in the image and in the buffer
	unsigned int pf_off = 0u;                                        // lanes 0 .. 5: the 128-byte lines of the strip's history row segment, PFD rows further down
	unsigned int oq[5] = {0u, 0u, 0u, 0u, 0u}, oe0 = 0u, oe1 = 0u;  // byte offsets in a history row: the lane's five texels (clamped to the image), alpha words beside them
	int K0 = 0;                                                      // first footprint row of output row Y0
	auto hrow = [&](int r) { return (unsigned int)(iclamp(r, 0, H - 1) - A.history_in.y0) * hpitch; };  // byte offset of history row r (clamp-to-edge)
	if (uni) {
		const float2 vxy = __half22float2(h2(vref.x));
		if (REJ) f_velz = __low2float(h2(vref.y));
		hu0 = u0 - vxy.x; hu1 = u1 - vxy.x;
		const AxisW a = catmull_axis(hu0, fW, invw);
		const AxisW b = catmull_axis_k(hu1, fW, invw, (float)(a.k + 1));
#pragma unroll
		for (int i = 0; i < 4; ++i) { wA[i] = a.w[i]; wB[i] = b.w[i]; }
		tx0 = (int)(hu0 * fW); tx1 = (int)(hu1 * fW);
		outx0 = hu0 < 0.f || hu0 >= 1.f; outx1 = hu1 < 0.f || hu1 >= 1.f;
#pragma unroll
		for (int j = 0; j < 5; ++j) oq[j] = (unsigned int)iclamp(a.k - 1 + j, 0, W - 1) * 8u;
		oe0 = (unsigned int)iclamp(a.k - 2, 0, W - 1) * 8u + 4u;
		oe1 = (unsigned int)iclamp(a.k + 4, 0, W - 1) * 8u + 4u;
		{
			const unsigned int first = __shfl_sync(0xffffffffu, oq[0], 0);  // lane 0's first texel: the segment starts there (or a few texels further left at the image border)
			pf_off = min((first & ~127u) + 128u * (unsigned int)lane, (unsigned int)(W - 1) * 8u);
		}
		// rows: the footprint of output row Y0 + i starts at K0 + i; rows outside the image are clamped to its edge, the rest must be in the buffer
		const float v_first = ((float)Y0 + 0.5f) / fH;
		K0 = catmull_axis(v_first - vxy.y, fH, invh).k - 1;
		const int rg = REJ ? 1 : 0;
		uni = iclamp(K0 - rg, 0, H - 1) >= hlo && iclamp(K0 + nr + 3, 0, H - 1) <= hhi;
		const int g = Y0 - 1 + lane, i = lane - 1;
		const float v = ((float)g + 0.5f) / fH;
		const float hv = v - vxy.y;
		const AxisW wy = catmull_axis_k(hv, fH, invh, (float)(K0 + 1 + i));
		sm.tabA[lane] = make_uint4(__float_as_uint(wy.w[0]), __float_as_uint(wy.w[1]), __float_as_uint(wy.w[2]), __float_as_uint(wy.w[3]));
		uint4 tb = sm.tabB[lane];
		tb.w = __float_as_uint(hv);
		sm.tabB[lane] = tb;
		uint4 tc = sm.tabC[lane];
		tc.y |= (hv < 0.f || hv >= 1.f) ? 0x80000000u : 0u;
		tc.z = (unsigned int)(int)(hv * fH);
		tc.w = hrow(K0 + lane + 2);  // entry t is read while output row t - 2 is evaluated: the row that joins the window then is K0 + (t - 2) + 4
		sm.tabC[lane] = tc;
	}
	__syncwarp();

	// ---- sampled colour rows Y0 - 1 and Y0 (taa.comp:207 through the sampler, YCoCg), their row sums ----
	const unsigned char* cr = sm.craw;
	auto sample_pair = [&](const uint4 tc, F3& a, F3& b) {
		const unsigned int cm = tc.x & 0xffffu, cn = tc.x >> 16;
		const __half py = __ushort_as_half((unsigned short)(tc.y & 0xffffu));
		const float3 va = sample_ycocg(*reinterpret_cast<const uint2*>(cr + (cm + cm0)), *reinterpret_cast<const uint2*>(cr + (cm + cn0)),
		                               *reinterpret_cast<const uint2*>(cr + (cn + cm0)), px0, py);
		const float3 vb = sample_ycocg(*reinterpret_cast<const uint2*>(cr + (cm + cm1)), *reinterpret_cast<const uint2*>(cr + (cm + cn1)),
		                               *reinterpret_cast<const uint2*>(cr + (cn + cm1)), px1, py);
		a.x = va.x; a.y = va.y; a.z = va.z;
		b.x = vb.x; b.y = vb.y; b.z = vb.z;
	};
	// Rolling state of the variance box, per column: Q[e] = row sums of sampled row y (e = parity of the unit row), PQ = sums of rows y - 1 and
	// y added up, CUR[e] = sampled colour of row y. A row adds sampled row y + 1 into slot e ^ 1: no register is ever moved.
	RowSum QA[2], QB[2], PQA, PQB;
	F3 CURA[2], CURB[2];
	{
		F3 a, b;
		RowSum rA, rB;
		sample_pair(sm.tabC[0], a, b);
		row_sums(a, b, rA, rB);
		sample_pair(sm.tabC[1], CURA[0], CURB[0]);
		row_sums(CURA[0], CURB[0], QA[0], QB[0]);
		PQA.s1.x = rA.s1.x + QA[0].s1.x; PQA.s1.y = rA.s1.y + QA[0].s1.y; PQA.s1.z = rA.s1.z + QA[0].s1.z;
		PQA.s2.x = rA.s2.x + QA[0].s2.x; PQA.s2.y = rA.s2.y + QA[0].s2.y; PQA.s2.z = rA.s2.z + QA[0].s2.z;
		PQB.s1.x = rB.s1.x + QB[0].s1.x; PQB.s1.y = rB.s1.y + QB[0].s1.y; PQB.s1.z = rB.s1.z + QB[0].s1.z;
		PQB.s2.x = rB.s2.x + QB[0].s2.x; PQB.s2.y = rB.s2.y + QB[0].s2.y; PQB.s2.z = rB.s2.z + QB[0].s2.z;
		QA[1] = QA[0]; QB[1] = QB[0]; CURA[1] = CURA[0]; CURB[1] = CURB[0];
	}
	// one row of the rolling box: samples row y + 1 (table entry tc), returns the 3 x 3 sums of row y
	auto advance_box = [&](auto EC, const uint4 tc, F3& S1a, F3& S2a, F3& S1b, F3& S2b) {
		constexpr int E = decltype(EC)::value;
		sample_pair(tc, CURA[E ^ 1], CURB[E ^ 1]);
		row_sums(CURA[E ^ 1], CURB[E ^ 1], QA[E ^ 1], QB[E ^ 1]);
		const RowSum &ra = QA[E ^ 1], &rb = QB[E ^ 1], &qa = QA[E], &qb = QB[E];
		S1a.x = PQA.s1.x + ra.s1.x; S1a.y = PQA.s1.y + ra.s1.y; S1a.z = PQA.s1.z + ra.s1.z;
		S2a.x = PQA.s2.x + ra.s2.x; S2a.y = PQA.s2.y + ra.s2.y; S2a.z = PQA.s2.z + ra.s2.z;
		S1b.x = PQB.s1.x + rb.s1.x; S1b.y = PQB.s1.y + rb.s1.y; S1b.z = PQB.s1.z + rb.s1.z;
		S2b.x = PQB.s2.x + rb.s2.x; S2b.y = PQB.s2.y + rb.s2.y; S2b.z = PQB.s2.z + rb.s2.z;
		PQA.s1.x = qa.s1.x + ra.s1.x; PQA.s1.y = qa.s1.y + ra.s1.y; PQA.s1.z = qa.s1.z + ra.s1.z;
		PQA.s2.x = qa.s2.x + ra.s2.x; PQA.s2.y = qa.s2.y + ra.s2.y; PQA.s2.z = qa.s2.z + ra.s2.z;
		PQB.s1.x = qb.s1.x + rb.s1.x; PQB.s1.y = qb.s1.y + rb.s1.y; PQB.s1.z = qb.s1.z + rb.s1.z;
		PQB.s2.x = qb.s2.x + rb.s2.x; PQB.s2.y = qb.s2.y + rb.s2.y; PQB.s2.z = qb.s2.z + rb.s2.z;
	};
	__syncwarp();
	if (lane == 0 && NSLOT < nslots) issue(NSLOT);  // slot 0 (rows Y0 - 2, Y0 - 1) is consumed

	// ---- output pointers (32-bit offsets: tuned_supports() admits only buffers below 4 GB) ----
	unsigned int o_hist = (unsigned int)(Y0 - A.history_out.y0) * (unsigned int)A.history_out.pitch + (unsigned int)x0c * 8u;
	unsigned int o_res = (unsigned int)(Y0 - A.result.y0) * (unsigned int)A.result.pitch + (unsigned int)x0c * 8u;
	unsigned int o_mask = (unsigned int)(Y0 - A.mask.y0) * (unsigned int)A.mask.pitch + (unsigned int)x0c * 4u;
	unsigned int fixA = 0u, fixB = 0u;  // rows of the unit whose pixel goes to the exact pass
	// stores as predicated instructions (no divergent branches around them): lanes 1 .. 30 write both pixels with one 16-byte store
	const unsigned int p16h = st16 ? 1u : 0u, p0h = (so0 && !st16) ? 1u : 0u, p1h = (so1 && !st16) ? 1u : 0u;
	const unsigned int hasres = A.result.p ? 1u : 0u, p16r = p16h & hasres, p0r = p0h & hasres, p1r = p1h & hasres;
	auto store_px = [&](unsigned char* base, const unsigned int off, const unsigned int q16, const unsigned int q0, const unsigned int q1, const unsigned int arg,
	                    const unsigned int ab, const unsigned int brg, const unsigned int bb) {
		unsigned char* ptr = base + off;
		asm volatile(
		    "{\n"
		    ".reg .pred a, b, c;\n"
		    "setp.ne.u32 a, %1, 0;\n"
		    "setp.ne.u32 b, %2, 0;\n"
		    "setp.ne.u32 c, %3, 0;\n"
		    "@a st.global.v4.u32 [%0], {%4, %5, %6, %7};\n"
		    "@b st.global.v2.u32 [%0], {%4, %5};\n"
		    "@c st.global.v2.u32 [%0 + 8], {%6, %7};\n"
		    "}\n" ::"l"(ptr), "r"(q16), "r"(q0), "r"(q1), "r"(arg), "r"(ab), "r"(brg), "r"(bb)
		    : "memory");
	};
	// ---- the sharpening epilogue: the resolved rows y - 1 (up) and y (centre) of the lane's two columns, as the rgba16f image holds them ----
	F3 eUpA = {0.f, 0.f, 0.f}, eUpB = eUpA, eCeA = eUpA, eCeB = eUpA;
	unsigned int o_fin = (unsigned int)(Y0 - A.final_img.y0) * (unsigned int)A.final_img.pitch + (unsigned int)x0c * 8u;  // offset of the CENTRE row
	const float ek = A.epilogue_k;
	const bool e16 = so0 && so1 && (((unsigned long long)A.final_img.p | (unsigned long long)A.final_img.pitch) & 15ull) == 0ull;
	const unsigned int p16f = e16 ? 1u : 0u, p0f = (so0 && !e16) ? 1u : 0u, p1f = (so1 && !e16) ? 1u : 0u;
	auto as_f3 = [](const unsigned int rg, const unsigned int b) { const float2 x = __half22float2(h2(rg)); F3 o; o.x = x.x; o.y = x.y; o.z = __low2float(h2(b)); return o; };
	// the centre row gets its sharpened value: `dn` = the row below it (zero below the image, as an image load out of range returns)
	auto emit_final = [&](const F3 dnA, const F3 dnB, const int yc) {
		const F3 zero = {0.f, 0.f, 0.f};
		// neighbours across lanes; at the image's left edge sharpen.comp clamps the coordinate (its own texel), CasFilter loads out of range (0);
		// right of the last column both read out of range
		F3 lfA, rtB;
		lfA.x = __shfl_up_sync(0xffffffffu, eCeB.x, 1); lfA.y = __shfl_up_sync(0xffffffffu, eCeB.y, 1); lfA.z = __shfl_up_sync(0xffffffffu, eCeB.z, 1);
		rtB.x = __shfl_down_sync(0xffffffffu, eCeA.x, 1); rtB.y = __shfl_down_sync(0xffffffffu, eCeA.y, 1); rtB.z = __shfl_down_sync(0xffffffffu, eCeA.z, 1);
		if (c0 == 0) lfA = EPI == 1 ? eCeA : zero;
		if (c1 == W - 1) rtB = zero;
		const F3 rtA = c1 <= W - 1 ? eCeB : zero;
		const F3 upA = yc == 0 ? (EPI == 1 ? eCeA : zero) : eUpA, upB = yc == 0 ? (EPI == 1 ? eCeB : zero) : eUpB;
		if (yc >= Yo && yc < Yo + no) {
			const F3 fa = sharpen_px<EPI>(upA, lfA, eCeA, rtA, dnA, ek), fb = sharpen_px<EPI>(upB, eCeA, eCeB, rtB, dnB, ek);
			const __half2 arg = __floats2half2_rn(fa.x, fa.y), ab = __floats2half2_rn(fa.z, 1.0f), brg = __floats2half2_rn(fb.x, fb.y), bb = __floats2half2_rn(fb.z, 1.0f);
			store_px(A.final_img.p, o_fin, p16f, p0f, p1f, *reinterpret_cast<const unsigned int*>(&arg), *reinterpret_cast<const unsigned int*>(&ab),
			         *reinterpret_cast<const unsigned int*>(&brg), *reinterpret_cast<const unsigned int*>(&bb));
		}
	};
	auto store_row = [&](const PixOut& a, const PixOut& b, const int i) {
		const bool own = !EPI || (Y0 + i >= Yo && Y0 + i < Yo + no);  // (the extra rows of the epilogue are resolved but not written)
		store_px(A.history_out.p, o_hist, own ? p16h : 0u, own ? p0h : 0u, own ? p1h : 0u, a.rg, a.bh, b.rg, b.bh);
		if (PEER) {  // the band's first / last halo rows go to the neighbour's halo as well
			const int g = Y0 + i;
			if (side_top && g < A.band_y0 + peer.halo)
				store_px(peer.nb_hist[0], (unsigned int)(g - peer.nb_y0[0]) * (unsigned int)peer.nb_pitch[0] + (unsigned int)x0c * 8u, p16h, p0h, p1h, a.rg, a.bh, b.rg, b.bh);
			if (side_bot && g >= A.band_y0 + A.band_rows - peer.halo)
				store_px(peer.nb_hist[1], (unsigned int)(g - peer.nb_y0[1]) * (unsigned int)peer.nb_pitch[1] + (unsigned int)x0c * 8u, p16h, p0h, p1h, a.rg, a.bh, b.rg, b.bh);
		}
		store_px(A.result.p, o_res, own ? p16r : 0u, own ? p0r : 0u, own ? p1r : 0u, a.rg, a.br, b.rg, b.br);
		if (EPI) {
			const F3 nA = as_f3(a.rg, a.br), nB = as_f3(b.rg, b.br);
			if (i > 0) { emit_final(nA, nB, Y0 + i - 1); o_fin += (unsigned int)A.final_img.pitch; }
			eUpA = eCeA; eUpB = eCeB; eCeA = nA; eCeB = nB;
			if (i == nr - 1 && Y0 + i == H - 1) { const F3 zero = {0.f, 0.f, 0.f}; emit_final(zero, zero, Y0 + i); }  // the image's last row
		}
		if (DIAG) {
			if (A.mask.p) {
				if (so0) *reinterpret_cast<unsigned int*>(A.mask.p + o_mask) = a.mask;
				if (so1) *reinterpret_cast<unsigned int*>(A.mask.p + o_mask + 4u) = b.mask;
			}
			if (a.uncertain && so0) fixA |= 1u << i;
			if (b.uncertain && so1) fixB |= 1u << i;
		}
		o_hist += (unsigned int)A.history_out.pitch;
		o_res += (unsigned int)A.result.pitch;
		o_mask += (unsigned int)A.mask.pitch;
	};
	// begin a two-row step: the slot with the rows it adds has landed; the velocity rows it adds are put to the vote
	auto step_begin = [&](const int i) {
		const int k = i / 2 + 2 + LOOK;
		if (k < nslots) {
			wait_slot(k);
			const bool ok = vote_row(Y0 + i + 1 + 2 * LOOK) & vote_row(Y0 + i + 2 + 2 * LOOK);
			uni = uni && __all_sync(0xffffffffu, ok);
		}
	};
	// end it: ring rows (Y0 - 2) + i + 2, + 3 are consumed, their slot is re-armed with the rows NSLOT slots further down
	auto step_end = [&](const int i) {
		__syncwarp();
		const int k = i / 2 + 1 + NSLOT;
		if (lane == 0 && k < nslots) issue(k);
	};

	int i = 0;
	bool begun = false;  // the step at row i has begun (its slot awaited, its rows voted) and then left the uniform path
	if (uni) {
		// ================================ uniform motion ================================
		HR hA[4], hB[4];
		unsigned int orw[4] = {0u, 0u, 0u, 0u}, or_prev = 0u;  // OR of the alpha words of the window's rows / of the row above it (columns k - 2 .. k + 4)
		{
			if (REJ) {
				const unsigned char* p = hbase + hrow(K0 - 1) + 4;
				or_prev = (__ldg(reinterpret_cast<const unsigned int*>(p + oq[0])) | __ldg(reinterpret_cast<const unsigned int*>(p + oq[1])) |
				           __ldg(reinterpret_cast<const unsigned int*>(p + oq[2]))) |
				          (__ldg(reinterpret_cast<const unsigned int*>(p + oq[3])) | __ldg(reinterpret_cast<const unsigned int*>(p + oq[4]))) |
				          (__ldg(reinterpret_cast<const unsigned int*>(p - 4 + oe0)) | __ldg(reinterpret_cast<const unsigned int*>(p - 4 + oe1)));
			}
#pragma unroll
			for (int j = 0; j < 4; ++j) {
				const unsigned char* p = hbase + hrow(K0 + j);
				const uint2 q0 = __ldg(reinterpret_cast<const uint2*>(p + oq[0])), q1 = __ldg(reinterpret_cast<const uint2*>(p + oq[1])),
				            q2 = __ldg(reinterpret_cast<const uint2*>(p + oq[2])), q3 = __ldg(reinterpret_cast<const uint2*>(p + oq[3])),
				            q4 = __ldg(reinterpret_cast<const uint2*>(p + oq[4]));
				if (REJ) {
					const unsigned int e0 = __ldg(reinterpret_cast<const unsigned int*>(p + oe0)), e1 = __ldg(reinterpret_cast<const unsigned int*>(p + oe1));
					orw[j] = (q0.y | q1.y | q2.y) | (q3.y | q4.y) | (e0 | e1);
				}
				hfilter_pair<REJ>(q0, q1, q2, q3, q4, wA, wB, hA[j], hB[j]);
			}
		}
		float hdA = 0.f, hdB = 0.f;  // previous depth at the history position of the next pixel row (requested one row ahead)
		if (use_depth) {
			const int ty = (int)sm.tabC[1].z;
			hdA = fetch_r32f(A.history_depth, W, H, tx0, ty, st);
			hdB = fetch_r32f(A.history_depth, W, H, tx1, ty, st);
			hdA = __fadd_rn(hdA, 0.0f); hdB = __fadd_rn(hdB, 0.0f);  // consumed here, not by the first row of the loop (see taa_resolve_strip.cu)
		}

		auto fast_row = [&](auto PHC, const int i) {
			constexpr int PH = decltype(PHC)::value;  // the register slot the row in flight replaces: i & 3
			const int t = i + 1;
			// the history row the next pixel row adds, K0 + i + 4, is requested now and consumed at the end of this row
			const uint4 tcn = sm.tabC[t + 1];
			const unsigned char* p = hbase + tcn.w;
			uint2 q0 = __ldg(reinterpret_cast<const uint2*>(p + oq[0])), q1 = __ldg(reinterpret_cast<const uint2*>(p + oq[1])),
			      q2 = __ldg(reinterpret_cast<const uint2*>(p + oq[2])), q3 = __ldg(reinterpret_cast<const uint2*>(p + oq[3])),
			      q4 = __ldg(reinterpret_cast<const uint2*>(p + oq[4]));
			unsigned int e0 = 0u, e1 = 0u;
			if (REJ) {
				e0 = __ldg(reinterpret_cast<const unsigned int*>(p + oe0));
				e1 = __ldg(reinterpret_cast<const unsigned int*>(p + oe1));
			}
			// the rows further down are pulled into L2 meanwhile: one row of look-ahead covers an L2 hit, not a DRAM access under load
			if (PFD > 0 && lane < 6) asm volatile("prefetch.global.L2 [%0];" ::"l"(hbase + (min(tcn.w + (unsigned int)PFD * hpitch, hmax_off) + pf_off)));
			float hdA_n = 0.f, hdB_n = 0.f;
			if (use_depth && i + 1 < nr) {
				const int ty = (int)sm.tabC[t + 1].z;
				hdA_n = fetch_r32f(A.history_depth, W, H, tx0, ty, st);
				hdB_n = fetch_r32f(A.history_depth, W, H, tx1, ty, st);
			}
			const uint4 wyu = sm.tabA[t];
			const float wy0 = __uint_as_float(wyu.x), wy1 = __uint_as_float(wyu.y), wy2 = __uint_as_float(wyu.z), wy3 = __uint_as_float(wyu.w);
			// sampled row y + 1 joins the rolling box
			F3 S1a, S2a, S1b, S2b;
			advance_box(std::integral_constant<int, PH & 1>{}, tcn, S1a, S2a, S1b, S2b);
			// the footprints, filtered vertically: history row K0 + i + j sits in register slot (i + j) & 3 = (PH + j) & 3. Summed in the order of
			// the rows, like the general rows do: the two paths give the same bits
			constexpr int J0 = PH & 3, J1 = (PH + 1) & 3, J2 = (PH + 2) & 3, J3 = (PH + 3) & 3;
			const float ar = fmaf(wy3, hA[J3].r, fmaf(wy2, hA[J2].r, fmaf(wy1, hA[J1].r, wy0 * hA[J0].r)));
			const float ag = fmaf(wy3, hA[J3].g, fmaf(wy2, hA[J2].g, fmaf(wy1, hA[J1].g, wy0 * hA[J0].g)));
			const float ab = fmaf(wy3, hA[J3].b, fmaf(wy2, hA[J2].b, fmaf(wy1, hA[J1].b, wy0 * hA[J0].b)));
			const float br = fmaf(wy3, hB[J3].r, fmaf(wy2, hB[J2].r, fmaf(wy1, hB[J1].r, wy0 * hB[J0].r)));
			const float bg = fmaf(wy3, hB[J3].g, fmaf(wy2, hB[J2].g, fmaf(wy1, hB[J1].g, wy0 * hB[J0].g)));
			const float bb = fmaf(wy3, hB[J3].b, fmaf(wy2, hB[J2].b, fmaf(wy1, hB[J1].b, wy0 * hB[J0].b)));
			float aa = 0.f, ba = 0.f;
			if (REJ) {
				aa = fmaf(wy3, hA[J3].a, fmaf(wy2, hA[J2].a, fmaf(wy1, hA[J1].a, wy0 * hA[J0].a)));
				ba = fmaf(wy3, hB[J3].a, fmaf(wy2, hB[J2].a, fmaf(wy1, hB[J1].a, wy0 * hB[J0].a)));
			}
			// rejection (taa.comp:787-823), exact predicates
			bool rejA = false, rejB = false;
			float dv = 0.f;
			if (REJ || ALPHA) {
				const uint4 tb = sm.tabB[t];
				dv = __uint_as_float(tb.z) - __uint_as_float(tb.w);
			}
			if (REJ) {
				const uint4 tc = sm.tabC[t];
				if (Sw<FX>::outside(P)) { rejA = outx0 || (tc.y >> 31) != 0u; rejB = outx1 || (tc.y >> 31) != 0u; }
				if (use_depth) {
					const float2 d = *reinterpret_cast<const float2*>(sm.dt + ring_row(Y0 + i, Y0 - 2, NR) * DROWB + (unsigned int)(2 * lane + (Xs & 3)) * 4u);
					const float ea = d.x - f_velz, eb = d.y - f_velz;
					if (fabsf(hdA - ea) > 0.1f * (1.0f - hdA)) rejA = true;
					if (fabsf(hdB - eb) > 0.1f * (1.0f - hdB)) rejB = true;
				}
			}
			PixOut oa = resolve_pixel<REJ, ALPHA, DIAG, FX>(A, kc, fix_band, CURA[PH & 1], S1a, S2a, ar, ag, ab, aa, rejA, false, false, u0 - hu0, dv);
			PixOut ob = resolve_pixel<REJ, ALPHA, DIAG, FX>(A, kc, fix_band, CURB[PH & 1], S1b, S2b, br, bg, bb, ba, rejB, false, false, u1 - hu1, dv);
			// slide the window: the row that was in flight replaces the oldest. The texels are not to be touched (= waited for) any earlier:
			// left alone, ptxas packs the blue halves of the five texels into spare register halves right behind the loads to save registers, and
			// the look-ahead is gone (ncu: 40 % of the stall samples sat on those PRMTs).
			// What keeps them: the second word of each texel is XORed with a zero that exists only once the row's colours do (rt_zero is a kernel
			// argument that is always 0; ptxas cannot know).
			if (!REJ) {
				store_row(oa, ob, i);
				const unsigned int z = (oa.rg ^ ob.bh) & rt_zero;
				q0.y ^= z; q1.y ^= z; q2.y ^= z; q3.y ^= z; q4.y ^= z;
			}
			if (REJ) {
				const unsigned int or_new = (q0.y | q1.y | q2.y) | (q3.y | q4.y) | (e0 | e1);
				if ((oa.check_ring || ob.check_ring) && (((or_prev | or_new) | (orw[0] | orw[1]) | (orw[2] | orw[3])) & 0x7fff0000u)) {
					if (oa.check_ring) oa.uncertain = true;
					if (ob.check_ring) ob.uncertain = true;
				}
				or_prev = orw[PH];
				orw[PH] = or_new;
				hdA = hdA_n; hdB = hdB_n;
				store_row(oa, ob, i);
			}
			hfilter_pair<REJ>(q0, q1, q2, q3, q4, wA, wB, hA[PH], hB[PH]);
		};

		// (The row counter advances through an opaque instruction that stays behind the rows' stores. With the plain `i += 2` the increment is
		// hoisted to the top of the loop body, and in variants that spill ptxas 12.9 was seen to re-materialise `i + 3`, the table index of the
		// third row, from the ALREADY incremented register: that row then read the table entries of four rows further down.)
		while (i < nr) {
			step_begin(i);
			if (!uni) { begun = true; break; }
			fast_row(std::integral_constant<int, 0>{}, i);
			if (i + 1 < nr) fast_row(std::integral_constant<int, 1>{}, i + 1);
			step_end(i);
			asm volatile("add.s32 %0, %0, 2;" : "+r"(i));
			if (i >= nr) break;
			step_begin(i);
			if (!uni) { begun = true; break; }
			fast_row(std::integral_constant<int, 2>{}, i);
			if (i + 1 < nr) fast_row(std::integral_constant<int, 3>{}, i + 1);
			step_end(i);
			asm volatile("add.s32 %0, %0, 2;" : "+r"(i));
		}
	}

	if (i < nr) {
		// ================================ general rows ================================
		// a slow unit: the next frame starts it first — and its left and right neighbours with it (an edge that moves on by a strip is then
		// expected there; a hinted unit that turns out fast costs nothing, it just runs early)
		if (hint_out && lane < 3) {
			const int nb = bx + (lane == 0 ? 0 : lane == 1 ? -1 : 1);
			const unsigned int u = (cta / (unsigned int)geo.nx) * (unsigned int)geo.nx + (unsigned int)nb;  // hint units are numbered in DISPATCH order (by may be remapped)
			if (nb >= 0 && nb < geo.nx && atomicCAS(&hint_out[2u + HINT_N + u], 0u, 0xffffffffu) == 0u) {
				const unsigned int slot = atomicAdd(&hint_out[0], 1u);
				if (slot < HINT_N) hint_out[2u + slot] = u;
				hint_out[2u + HINT_N + u] = slot < HINT_N ? slot + 1u : 0u;
			}
		}
#ifdef TAA_STREAM_TRACE
		tr_general = i;
#endif
		const unsigned char* vraw = sm.vraw;
		unsigned int vo00, vo01, vo10, vo11;  // velocity footprint columns of the lane's two pixels (byte offsets in a ring row)
		float va0, va1;
		{
			const Lin L0 = lin_coord(u0, W), L1 = lin_coord(u1, W);
			vo00 = (unsigned int)iclamp(L0.i0 - Xb, 0, RWT - 1) * 8u; vo01 = (unsigned int)iclamp(L0.i1 - Xb, 0, RWT - 1) * 8u; va0 = L0.a;
			vo10 = (unsigned int)iclamp(L1.i0 - Xb, 0, RWT - 1) * 8u; vo11 = (unsigned int)iclamp(L1.i1 - Xb, 0, RWT - 1) * 8u; va1 = L1.a;
		}
		struct Pos { float hu, hv, v, velz; bool movC; };  // where the pixel's history lies
		// ---- getHistoryPosition (taa.comp:391-438), exact ----
		auto history_pos = [&](const int i, const unsigned int o0, const unsigned int o1, const float a, const float u) -> Pos {
			const uint4 tb = sm.tabB[i + 1];
			const unsigned int r0 = tb.x & 0xffffu, r1 = tb.x >> 16;
			const float ra = __uint_as_float(tb.y);
			Pos o;
			o.v = __uint_as_float(tb.z);
			const uint2 vt00 = *reinterpret_cast<const uint2*>(vraw + (r0 + o0)), vt10 = *reinterpret_cast<const uint2*>(vraw + (r0 + o1));
			const uint2 vt01 = *reinterpret_cast<const uint2*>(vraw + (r1 + o0)), vt11 = *reinterpret_cast<const uint2*>(vraw + (r1 + o1));
			float velx, vely;
			o.velz = 0.f;
			o.movC = false;
			bool same = vt00.x == vt10.x && vt00.x == vt01.x && vt00.x == vt11.x && finite2(vt00.x);
			if (REJ) same = same && vt00.y == vt10.y && vt00.y == vt01.y && vt00.y == vt11.y && finite2(vt00.y);
			if (same) {  // lerp(p, p, w) == p + w * 0 == p for finite p
				const float2 xy = __half22float2(h2(vt00.x));
				velx = xy.x; vely = xy.y;
				if (REJ) {
					const float2 zw = __half22float2(h2(vt00.y));
					o.velz = zw.x;
					o.movC = (fabsf(velx) > 1e-5f || fabsf(vely) > 1e-5f) && (fabsf(zw.y) >= 0.5f);
				}
			} else {
				const float2 a00 = __half22float2(h2(vt00.x)), a10 = __half22float2(h2(vt10.x)), a01 = __half22float2(h2(vt01.x)), a11 = __half22float2(h2(vt11.x));
				velx = lerpf(lerpf(a00.x, a10.x, a), lerpf(a01.x, a11.x, a), ra);
				vely = lerpf(lerpf(a00.y, a10.y, a), lerpf(a01.y, a11.y, a), ra);
				if (REJ) {
					const float2 b00 = __half22float2(h2(vt00.y)), b10 = __half22float2(h2(vt10.y)), b01 = __half22float2(h2(vt01.y)), b11 = __half22float2(h2(vt11.y));
					o.velz = lerpf(lerpf(b00.x, b10.x, a), lerpf(b01.x, b11.x, a), ra);
					const float velw = lerpf(lerpf(b00.y, b10.y, a), lerpf(b01.y, b11.y, a), ra);
					o.movC = (fabsf(velx) > 1e-5f || fabsf(vely) > 1e-5f) && (fabsf(velw) >= 0.5f);
				}
			}
			o.hu = u - velx; o.hv = o.v - vely;
			return o;
		};
		struct HS { float r, g, b, a; unsigned int ring; };  // the filtered history sample
		// everything after the history sample
		auto finish_pixel = [&](const Pos& ps, const HS& hs, const float u, const F3 cur, const F3 S1, const F3 S2, const float depth, const bool movers_near) -> PixOut {
			const float hu = ps.hu, hv = ps.hv, v = ps.v;
			bool rejected = false, movement = false;
			if (REJ) {
				if (Sw<FX>::outside(P) && (hu < 0.f || hv < 0.f || hu >= 1.f || hv >= 1.f)) rejected = true;
				if (Sw<FX>::antighost(P)) {
					movement = ps.movC;
					// the four other taps of the movement test (taa.comp:796-806): where no texel within two of the strip's recent rows carries
					// velocity.w, every tap's w is exactly 0
					if (!movement && movers_near) {
						auto mov = [&](float s, float t) {
							const float4 q = tex_rgba16f(A.velocity, W, H, s, t, st);
							return (fabsf(q.x) > 1e-5f || fabsf(q.y) > 1e-5f) && (fabsf(q.w) >= 0.5f);
						};
						movement = mov(u + invw * -1.f, v + invh * 0.f) || mov(u + invw * 1.f, v + invh * 0.f) || mov(u + invw * 0.f, v + invh * -1.f) ||
						           mov(u + invw * 0.f, v + invh * 1.f);
					}
				}
				if (use_depth) {
					const float expected = depth - ps.velz;
					const float hd = fetch_r32f(A.history_depth, W, H, (int)(hu * fW), (int)(hv * fH), st);
					if (fabsf(hd - expected) > 0.1f * (1.0f - hd)) rejected = true;
				}
			}
			PixOut o = resolve_pixel<REJ, ALPHA, DIAG, FX>(A, kc, fix_band, cur, S1, S2, hs.r, hs.g, hs.b, hs.a, rejected, movement, ps.movC, u - hu, v - hv);
			if (REJ && o.check_ring && (hs.ring & 0x7fff0000u)) o.uncertain = true;
			return o;
		};
		auto general_row = [&](const int i) {  // (the rolling box is kept in slot 0: the general rows move it there, a few MOVs do not matter here)
			const int t = i + 1;
			F3 S1a, S2a, S1b, S2b;
			advance_box(std::integral_constant<int, 0>{}, sm.tabC[t + 1], S1a, S2a, S1b, S2b);
			const F3 ca = CURA[0], cb = CURB[0];
			QA[0] = QA[1]; QB[0] = QB[1]; CURA[0] = CURA[1]; CURB[0] = CURB[1];
			float2 d = make_float2(0.f, 0.f);
			if (use_depth) d = *reinterpret_cast<const float2*>(sm.dt + ring_row(Y0 + i, Y0 - 2, NR) * DROWB + (unsigned int)(2 * lane + (Xs & 3)) * 4u);
			// velocity rows up to Y0 + i + 4 (+ 1) have been voted: bits 0 .. 7 cover rows y - 2 .. y + 2 of both rows of the step
			const bool movers_near = REJ && (wrows & 0xffu) != 0u;
			const Pos pa = history_pos(i, vo00, vo01, va0, u0), pb = history_pos(i, vo10, vo11, va1, u1);
			HS ha, hb;
			// The usual case under smoothly varying motion: the two footprints are interior, start on the same row and one column apart (columns
			// k - 1 .. k + 3 cover both). If that holds for every lane, the 4 x 5 texels are requested in one batch and each is converted once; the
			// arithmetic per pixel is the one of gather_history (same weights, same order of summation: same bits).
			const AxisW ax0 = catmull_axis(pa.hu, fW, invw), ay0 = catmull_axis(pa.hv, fH, invh);
			const AxisW ax1 = catmull_axis(pb.hu, fW, invw), ay1 = catmull_axis(pb.hv, fH, invh);
			const int kx = ax0.k, K = ay0.k - 1;
			const bool paired = !REJ && ax1.k == kx + 1 && ay1.k == ay0.k && kx - 1 >= 0 && kx + 3 <= W - 1 && K >= hlo && K + 3 <= hhi;
			if (__all_sync(0xffffffffu, paired)) {
				const unsigned char* p = hbase + ((unsigned int)(K - A.history_in.y0) * hpitch + (unsigned int)(kx - 1) * 8u);
				uint2 q[20];
#pragma unroll
				for (int r = 0; r < 4; ++r) {
					const uint2* hp = reinterpret_cast<const uint2*>(p + r * hpitch);
#pragma unroll
					for (int c = 0; c < 5; ++c) q[5 * r + c] = __ldg(hp + c);
				}
				HR fa[4], fb[4];
#pragma unroll
				for (int r = 0; r < 4; ++r) hfilter_pair<false>(q[5 * r], q[5 * r + 1], q[5 * r + 2], q[5 * r + 3], q[5 * r + 4], ax0.w, ax1.w, fa[r], fb[r]);
				ha.r = fmaf(ay0.w[3], fa[3].r, fmaf(ay0.w[2], fa[2].r, fmaf(ay0.w[1], fa[1].r, ay0.w[0] * fa[0].r)));
				ha.g = fmaf(ay0.w[3], fa[3].g, fmaf(ay0.w[2], fa[2].g, fmaf(ay0.w[1], fa[1].g, ay0.w[0] * fa[0].g)));
				ha.b = fmaf(ay0.w[3], fa[3].b, fmaf(ay0.w[2], fa[2].b, fmaf(ay0.w[1], fa[1].b, ay0.w[0] * fa[0].b)));
				hb.r = fmaf(ay1.w[3], fb[3].r, fmaf(ay1.w[2], fb[2].r, fmaf(ay1.w[1], fb[1].r, ay1.w[0] * fb[0].r)));
				hb.g = fmaf(ay1.w[3], fb[3].g, fmaf(ay1.w[2], fb[2].g, fmaf(ay1.w[1], fb[1].g, ay1.w[0] * fb[0].g)));
				hb.b = fmaf(ay1.w[3], fb[3].b, fmaf(ay1.w[2], fb[2].b, fmaf(ay1.w[1], fb[1].b, ay1.w[0] * fb[0].b)));
				ha.a = 0.f; hb.a = 0.f; ha.ring = 0u; hb.ring = 0u;
			} else {
				gather_history<REJ>(A, ax0, ay0, W, H, hlo, hhi, ha.r, ha.g, ha.b, ha.a, ha.ring);
				gather_history<REJ>(A, ax1, ay1, W, H, hlo, hhi, hb.r, hb.g, hb.b, hb.a, hb.ring);
			}
			const PixOut oa = finish_pixel(pa, ha, u0, ca, S1a, S2a, d.x, movers_near);
			const PixOut ob = finish_pixel(pb, hb, u1, cb, S1b, S2b, d.y, movers_near);
			store_row(oa, ob, i);
		};
		if (i & 1) { QA[0] = QA[1]; QB[0] = QB[1]; CURA[0] = CURA[1]; CURB[0] = CURB[1]; }  // (never: the uniform rows are left at a step boundary)
		while (i < nr) {
			if (!begun) step_begin(i);
			begun = false;
#pragma unroll 1
			for (int r = 0; r < 2; ++r)  // (one copy of the row in the instruction stream: the general rows were short of instruction cache)
				if (i + r < nr) general_row(i + r);
			step_end(i);
			i += 2;
		}
	}

	if (PEER && (side_top || side_bot)) {
		__threadfence_system();  // every lane's stores (local and peer) are ordered before the signal (measured: 1-3 % of a band's time against
		                         // relying on the warp barrier + lane 0's cumulative release alone; kept)
		__syncwarp();
		if (lane == 0) {
			for (int sd = 0; sd < 2; ++sd) {
				if (!(sd == 0 ? side_top : side_bot)) continue;
				unsigned int* done = peer.flags + (4 + 2 * sd + peer.q);
				if (atomicAdd(done, 1u) == peer.mine[sd] - 1u) {  // the side's last warp: the word the side waited on is free for the frame after the next
					peer.flags[2 * sd + (peer.q ^ 1)] = 0u;
					*done = 0u;
					__threadfence_system();
				}
				// this band is side (1 - sd) of that neighbour
				asm volatile("red.release.sys.global.add.u32 [%0], 1;" ::"l"(peer.nb_flags[sd] + (2 * (1 - sd) + peer.q)) : "memory");
			}
		}
	}
#ifdef TAA_STREAM_TRACE
	if (lane == 0) {
		const unsigned int u = (cta * NWARP + warp) & 16383u;
		g_stream_trace[4 * u] = tr_t0; g_stream_trace[4 * u + 1] = gtime();
		g_stream_trace[4 * u + 2] = smid() | ((unsigned long long)tr_general << 16); g_stream_trace[4 * u + 3] = (unsigned int)Y0 | ((unsigned long long)strip << 16);
	}
#endif
	// ---- hand the undecidable pixels of the unit to the exact pass (one atomic per warp) ----
	if (DIAG && fix_list != nullptr && __ballot_sync(0xffffffffu, (fixA | fixB) != 0u)) {
		const int n = __popc(fixA) + __popc(fixB);
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
		while (fixA) {
			const int b = __ffs(fixA) - 1;
			fixA &= fixA - 1u;
			fix_list[slot++] = (unsigned int)(Y0 + b) * (unsigned int)W + (unsigned int)c0;
		}
		while (fixB) {
			const int b = __ffs(fixB) - 1;
			fixB &= fixB - 1u;
			fix_list[slot++] = (unsigned int)(Y0 + b) * (unsigned int)W + (unsigned int)c1;
		}
	}
}

#define TAA_STREAM_PARAMS                                                                                                                              \
	const __grid_constant__ ResolveArgs A, const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmD, \
	    unsigned int *__restrict__ fix_list, unsigned int *__restrict__ fix_count, unsigned int *__restrict__ fix_count_next, const float fix_band,            \
	    const __grid_constant__ UnitGeo geo, const unsigned int rt_zero, const unsigned int *__restrict__ hint_in, unsigned int *__restrict__ hint_out,        \
	    unsigned int *__restrict__ hint_clear, const unsigned int hint_sig, const __grid_constant__ PeerArgs peer
#define TAA_STREAM_ARGS A, tmC, tmV, tmD, fix_list, fix_count, fix_count_next, fix_band, geo, rt_zero, hint_in, hint_out, hint_clear, hint_sig, peer

// register budget = 64 K / (MINB pairs of warps): MINB = 6 -> 168 registers, 12 warps per SM
template <bool REJ, bool ALPHA, bool DIAG, int FX, int MINB, int EPI, bool PEER>
__global__ void __launch_bounds__(32 * NWARP, MINB * 2 / NWARP) taa_resolve_stream_kernel(TAA_STREAM_PARAMS) {
	stream_body<REJ, ALPHA, DIAG, FX, EPI, PEER>(TAA_STREAM_ARGS);
}
// (A register budget given directly, __maxnreg__(144) = 7 pairs of warps per SM with 40 bytes of spills, measured 0.1085 ms against 0.0979 at
// 168 registers / 6 pairs; __launch_bounds__(64, 8) = 128 registers, 112 bytes of spills: 0.193 against 0.179 at the time.)

// ---- host side -------------------------------------------------------------------------------------------------------------------------

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
	static EncodeTiledFn fn = [] {
		void* p = nullptr;
		cudaDriverEntryPointQueryResult q;
		if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
		return (EncodeTiledFn)p;
	}();
	return fn;
}

// a row-major image of 4-byte words (an 8-byte texel is two), boxes of box_w words x SROWS rows, out-of-range words read as 0
bool make_map(CUtensorMap* tm, const Img& im, int words_per_row, int box_w) {
	EncodeTiledFn fn = encode_fn();
	if (!fn) return false;
	const cuuint64_t dims[2] = {(cuuint64_t)words_per_row, (cuuint64_t)im.rows};
	const cuuint64_t strides[1] = {(cuuint64_t)im.pitch};
	const cuuint32_t box[2] = {(cuuint32_t)box_w, (cuuint32_t)SROWS};
	const cuuint32_t estr[2] = {1u, 1u};
	return fn(tm, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, (void*)im.p, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
	          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

bool tma_able(const Img& im) { return im.p && (((unsigned long long)im.p | (unsigned long long)im.pitch) & 15ull) == 0ull && im.rows > 0; }

// Rows per unit: the grid is (strip pairs) x ceil(band / R) CTAs on `resident` CTA slots; take the R whose last wave is fullest, weighted
// by the per-unit overhead (tables, two extra sampled rows, the first window: ~2.5 rows' worth of instructions)
int pick_rows(int nx, int band_rows, int resident, int rmax) {
	static const int forced = [] { const char* s = getenv("TAA_STREAM_R"); return s ? atoi(s) : 0; }();
	if (forced >= 2 && forced <= rmax) return forced;
	int best = rmax;
	double best_eff = -1.0;
	for (int r = 12; r <= rmax; ++r) {
		const double n = (double)nx * ((band_rows + r - 1) / r);
		const double waves = n / resident;
		const double full = waves / (double)(long long)(waves + 0.999999);
		const double eff = full * r / (r + 2.5);
		if (eff > best_eff + 1e-9) { best_eff = eff; best = r; }
	}
	return best;
}

// The unit geometry of a band of `band_rows` rows (see UnitGeo); halo_top / halo_bot > 0: the band has a neighbour on that side whose halo is that high
UnitGeo unit_geometry(int nx, int band_rows, int resident, int rmax, bool hints_on, int halo_top, int halo_bot, bool epilogue = false) {
	const int R = pick_rows(nx, band_rows, resident, rmax);
	// the last rows of the band in short units
	// (measured on B200, 4K pan + one mover, R = 26: no hints, no tail 0.1033 ms; hints alone 0.1043; tail alone 0.1054; hints + 15 % tail in units of 14
	// rows 0.0958; 25 %: 0.0960; 35 %: 0.0994; units of 8 rows: 0.0968 .. 0.1015)
	static const int tail_env = [] { const char* v = getenv("TAA_STREAM_TAIL"); return v ? atoi(v) : -1; }();   // tuning aids
	// (a unit of the sharpening-epilogue variants resolves two extra rows: short units cost more there — measured 0.1733 ms at 10 %, 0.1742 at 20 %, 0.1809 at 0)
	const int tail_pct = tail_env >= 0 ? tail_env : (hints_on ? (epilogue ? 10 : 20) : 0);
	static const int rs_env = [] { const char* v = getenv("TAA_STREAM_RS"); return v ? atoi(v) : 0; }();
	UnitGeo geo;
	geo.nx = nx; geo.R = R;

	geo.Rs = rs_env >= 2 && rs_env <= R ? rs_env : max(2, (R / 2 + 1) & ~1);
	geo.nbig = tail_pct > 0 ? (int)(((long long)band_rows * (100 - min(tail_pct, 100)) / 100) / R) : (band_rows + R - 1) / R;
	const int rest = max(0, band_rows - geo.nbig * R);
	geo.ny = geo.nbig + (rest + geo.Rs - 1) / geo.Rs;
	geo.nbt = geo.nbb = 0;
	for (int b = 0; b < geo.ny; ++b) {
		const int y = b < geo.nbig ? b * R : geo.nbig * R + (b - geo.nbig) * geo.Rs;
		const int e = min(band_rows, y + (b < geo.nbig ? R : geo.Rs));
		if (halo_top > 0 && y < halo_top) ++geo.nbt;
		if (halo_bot > 0 && e > band_rows - halo_bot) ++geo.nbb;
	}
	return geo;
}

template <bool REJ, bool ALPHA, bool DIAG, int FX, int MINB, int EPI, bool PEER = false>
cudaError_t launch_variant(const ResolveArgs& A, const CUtensorMap& tmC, const CUtensorMap& tmV, const CUtensorMap& tmD, unsigned int* fix_list, unsigned int* fix_count,
                           unsigned int* fix_count_next, float band, int num_sms, unsigned int* hints, int hint_phase, const StreamPeers* peers, cudaStream_t stream) {
	auto kern = taa_resolve_stream_kernel<REJ, ALPHA, DIAG, FX, MINB, EPI, PEER>;
	const int smem = (int)sizeof(WarpSmem<REJ>) * NWARP;
	static int resident_per_sm[64] = {0};  // per device (the attribute and the occupancy are per device)
	int dev = 0;
	cudaGetDevice(&dev);
	dev &= 63;
	if (!resident_per_sm[dev]) {
		cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
		if (e != cudaSuccess) return e;
		cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100 * (MINB * 2 / NWARP) * (smem + 1024) / (228 * 1024) + 2);
		int nb = 0;
		e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, 32 * NWARP, smem);
		if (e != cudaSuccess) return e;
		resident_per_sm[dev] = nb > 0 ? nb : 1;
	}
	// strips: 62 output columns starting at column -1 (the first strip's first column does not exist), or 60 starting at 0 with an epilogue
	const int nstrips = EPI ? (A.out_w + OWS - 3) / (OWS - 2) : (A.out_w + 1 + OWS - 1) / OWS;
	const int nx = (nstrips + NWARP - 1) / NWARP;
	const int resident = resident_per_sm[dev] * num_sms, rmax = EPI ? RMAX - 2 : RMAX;
	cudaLaunchConfig_t cfg = {};
	static const bool hints_off = [] { const char* v = getenv("TAA_STREAM_HINTS"); return v && v[0] == '0'; }();  // A/B aid
	const bool hints_on = hints && !hints_off;
	const bool up = PEER && peers->nb_hist[0], dn = PEER && peers->nb_hist[1];
	const UnitGeo geo = unit_geometry(nx, A.band_rows, resident, rmax, hints_on, up ? peers->halo : 0, dn ? peers->halo : 0, EPI != 0);
	const int ny = geo.ny, R = geo.R;
	PeerArgs pa = {};
	if (PEER) {
		if (geo.nbt + geo.nbb > ny || peers->halo > A.band_rows) return cudaErrorInvalidConfiguration;  // (a band lower than its two halos)
		for (int sd = 0; sd < 2; ++sd) {
			pa.nb_hist[sd] = peers->nb_hist[sd]; pa.nb_pitch[sd] = peers->nb_pitch[sd]; pa.nb_y0[sd] = peers->nb_y0[sd]; pa.nb_flags[sd] = peers->nb_flags[sd];
			if (!peers->nb_hist[sd]) continue;
			// what the neighbour signals: the warps of ITS boundary blocks that face this band (same function of its band height)
			const UnitGeo ng = unit_geometry(nx, peers->nb_band_rows[sd], resident, rmax, hints_on, peers->halo, peers->halo);
			pa.expect[sd] = (unsigned int)((sd == 0 ? ng.nbb : ng.nbt) * nstrips);
			pa.mine[sd] = (unsigned int)((sd == 0 ? geo.nbt : geo.nbb) * nstrips);
		}
		pa.flags = peers->flags; pa.halo = peers->halo; pa.q = peers->q; pa.wait = peers->wait;

	}
	const bool hinted = hints_on && (long long)nx * ny <= (long long)HINT_FLAGS;
	const unsigned int* hin = hinted ? hints + (size_t)(hint_phase % 3) * HINT_WORDS : nullptr;
	unsigned int* hout = hinted ? hints + (size_t)((hint_phase + 1) % 3) * HINT_WORDS : nullptr;
	unsigned int* hclr = hinted ? hints + (size_t)((hint_phase + 2) % 3) * HINT_WORDS : nullptr;
	const unsigned int sig = ((unsigned int)nx * 2654435761u) ^ ((unsigned int)ny * 40503u) ^ ((unsigned int)R << 24) ^ ((unsigned int)geo.Rs << 18) ^ ((unsigned int)geo.nbig * 977u) ^ (unsigned int)A.band_y0 ^ ((unsigned int)A.out_w << 8) ^ 0x5eedu;
	cfg.gridDim = dim3(nx * ny + (hinted ? (int)HINT_N : 0));
	cfg.blockDim = dim3(32 * NWARP);
	cfg.dynamicSmemBytes = smem;
	cfg.stream = stream;
	cudaLaunchAttribute attr[1];
	attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
	attr[0].val.programmaticStreamSerializationAllowed = 1;
	cfg.attrs = attr;
	cfg.numAttrs = 1;
	return cudaLaunchKernelEx(&cfg, kern, A, tmC, tmV, tmD, fix_list, fix_count, fix_count_next, band, geo, 0u, hin, hout, hclr, sig, pa);
}

}  // namespace

#ifdef TAA_STREAM_TRACE
extern "C" __attribute__((visibility("default"))) int taa_debug_stream_trace(void* dst, size_t bytes) {
	return (int)cudaMemcpyFromSymbol(dst, g_stream_trace, bytes < sizeof(g_stream_trace) ? bytes : sizeof(g_stream_trace));
}
#endif

bool stream_supports(const ResolveArgs& A) {
	static const bool off = [] { const char* v = getenv("TAA_TUNED_VARIANT"); return v && v[0] == 's'; }();  // "strip": the A/B partner
	if (off || !encode_fn()) return false;
	const TaaParameters& P = A.ubo.param[0];
	// The rejection variants (config 3) still run faster on the strip kernel (0.216 against 0.369 ms per 4K frame on B200): their uniform-motion
	// rows carry too much predicate and addressing work here yet. TAA_STREAM_REJ=1 routes them through this kernel (tests, tuning).
	static const bool rej_too = [] { const char* v = getenv("TAA_STREAM_REJ"); return v && v[0] == '1'; }();
	if (!rej_too && (P.mDepthCulling || P.mRejectOutside || P.mDynamicAntiGhosting)) return false;
	if (!tma_able(A.color) || !tma_able(A.velocity)) return false;
	if (P.mDepthCulling && !tma_able(A.depth)) return false;
	return true;
}

// The sharpening pass can ride in the resolve's epilogue when the call takes a plain variant of this kernel (no rejection switch) on a whole
// frame and leaves nothing to the exact fix-up pass (no mask bound), which would have to patch the sharpened neighbours of the pixels it rewrites.
bool stream_epilogue_ok(const ResolveArgs& A, bool fixup_all) {
	const TaaParameters& P = A.ubo.param[0];
	if (fixup_all || A.mask.p || P.mDepthCulling || P.mRejectOutside || P.mDynamicAntiGhosting) return false;
	if (A.band_y0 != 0 || A.band_rows != A.out_h) return false;
	return tuned_supports(A) && stream_supports(A);
}

size_t stream_hint_bytes() { return 3u * (size_t)HINT_WORDS * sizeof(unsigned int); }

cudaError_t launch_resolve_stream(const ResolveArgs& A, unsigned int* fix_list, unsigned int* fix_count, unsigned int* fix_count_next, bool fixup_all, int num_sms,
                                  unsigned int* hints, int hint_phase, const StreamPeers* peers, cudaStream_t stream) {
	const float band = fixup_all ? INFINITY : FIXUP_BAND_4K * fmaxf(1.0f, fmaxf((float)A.out_w / 3840.0f, (float)A.out_h / 3840.0f));
	const TaaParameters& P = A.ubo.param[0];
	const bool rej = P.mDepthCulling || P.mRejectOutside || P.mDynamicAntiGhosting;
	const bool alp = P.mVelBasedAlpha || P.mLumaWeightingLottes || P.mReduceBlendNearClamp;
	const bool diag = fix_list != nullptr || A.mask.p != nullptr;
	CUtensorMap tmC, tmV, tmD;
	if (!make_map(&tmC, A.color, 2 * A.in_w, 2 * RWT) || !make_map(&tmV, A.velocity, 2 * A.in_w, 2 * RWT)) return cudaErrorInvalidValue;
	if (rej && P.mDepthCulling) { if (!make_map(&tmD, A.depth, A.in_w, DW)) return cudaErrorInvalidValue; }
	else tmD = tmC;
	// Resident CTAs per SM (= the register budget): the A/B on B200 decides the defaults; TAA_STREAM_MINB overrides (tuning aid).
	static const int minb_env = [] { const char* v = getenv("TAA_STREAM_MINB"); return v ? atoi(v) : 0; }();
	const bool fx3 = rej && alp && P.mDepthCulling && P.mRejectOutside && P.mDynamicAntiGhosting && P.mVelBasedAlpha && P.mLumaWeightingLottes &&
	                 !P.mReduceBlendNearClamp && !A.ubo.mResetHistory;
#define TAA_STREAM_GO(RJ, AL, DG, FX, MB) return launch_variant<RJ, AL, DG, FX, MB, 0>(A, tmC, tmV, tmD, fix_list, fix_count, fix_count_next, band, num_sms, hints, hint_phase, nullptr, stream)
	if (peers) {  // boundary rows stored into the neighbours' halos, completion by flags: plain variants only (a fix-up pass would rewrite stored pixels)
		if (rej || diag || A.epilogue) return cudaErrorNotSupported;
		if (alp) return launch_variant<false, true, false, 0, 6, 0, true>(A, tmC, tmV, tmD, fix_list, fix_count, fix_count_next, band, num_sms, hints, hint_phase, peers, stream);
		return launch_variant<false, false, false, 0, 6, 0, true>(A, tmC, tmV, tmD, fix_list, fix_count, fix_count_next, band, num_sms, hints, hint_phase, peers, stream);
	}
	if (A.epilogue) {  // (stream_epilogue_ok() has admitted the call: a plain variant, nothing for the exact pass to decide)
#define TAA_STREAM_EPI(AL, MB, EP) return launch_variant<false, AL, false, 0, MB, EP>(A, tmC, tmV, tmD, fix_list, fix_count, fix_count_next, band, num_sms, hints, hint_phase, nullptr, stream)
		static const int epi_minb = [] { const char* v = getenv("TAA_STREAM_EPI_MINB"); return v ? atoi(v) : 0; }();  // tuning aid
		if (A.epilogue == 1) { if (alp) TAA_STREAM_EPI(true, 5, 1); TAA_STREAM_EPI(false, 5, 1); }
		if (alp) TAA_STREAM_EPI(true, 5, 2);
		if (epi_minb == 6) TAA_STREAM_EPI(false, 6, 2);
		if (epi_minb == 5) TAA_STREAM_EPI(false, 5, 2);
		TAA_STREAM_EPI(false, 4, 2);  // (measured on B200, 4K: 0.168 ms at 4 pairs of warps per SM (203 registers), 0.173 at 6 (168 + spills), 0.187 at 5)
#undef TAA_STREAM_EPI
	}
	if (rej) {
		if (fx3) { if (minb_env == 4) TAA_STREAM_GO(true, true, true, 1, 4); if (minb_env == 5) TAA_STREAM_GO(true, true, true, 1, 5); TAA_STREAM_GO(true, true, true, 1, 6); }
		if (alp) TAA_STREAM_GO(true, true, true, 0, 6);
		TAA_STREAM_GO(true, false, true, 0, 6);
	}
	if (diag) { if (alp) TAA_STREAM_GO(false, true, true, 0, 6); TAA_STREAM_GO(false, false, true, 0, 6); }
	if (alp) TAA_STREAM_GO(false, true, false, 0, 6);
	if (minb_env == 8) TAA_STREAM_GO(false, false, false, 0, 8);
	if (minb_env == 5) TAA_STREAM_GO(false, false, false, 0, 5);
	TAA_STREAM_GO(false, false, false, 0, 6);
#undef TAA_STREAM_GO
}

}  // namespace taa
