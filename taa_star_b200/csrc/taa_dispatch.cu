// taa_dispatch.cu — chooses the resolve kernel for a settings block (every switch of
// `Parameters` is uniform across a dispatch except the split-screen select).
#include "taa_ctx.h"
#include <cstdlib>

namespace taa {

// The settings family of the tuned kernels (BASELINE configs 2-5): YCoCg, variance box, clipAabb, Catmull-Rom history, velocity for everything;
// whole-frame or band, equal input and output size, buffers below 4 GB (the kernels address with 32-bit offsets).
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

static cudaError_t ensure_fix_buffers(taa_ctx* c) {
	if (c->fix_list) return cudaSuccess;
	cudaError_t e = cudaMalloc(&c->fix_list, (size_t)c->desc.out_width * c->desc.band_rows * sizeof(unsigned int));
	if (e != cudaSuccess) return e;
	e = cudaMalloc(&c->fix_count, 2 * sizeof(unsigned int));
	if (e != cudaSuccess) return e;
	return cudaMemset(c->fix_count, 0, 2 * sizeof(unsigned int));
}

// Returns the number of kernels launched through *launched.
cudaError_t dispatch_resolve(taa_ctx* c, const ResolveArgs& A, cudaStream_t s, int* launched) {
	*launched = 0;
	if (!(c->desc.flags & TAA_FLAG_EXACT) && tuned_supports(A)) {
		c->last_was_tuned = true;
		// The exact fix-up pass has something to decide only if (a) the mask is bound (`rectified`, taa.comp:845, is reported through it
		// alone), (b) dynamic anti-ghosting is on (the sign of a filtered alpha that cancels to ~0 flips `rejected`, and with it the colour),
		// or (c) TAA_FLAG_FIXUP_ALL asks for it. Otherwise the tuned kernel runs alone: its colours are within ~1e-5 of the exact ones.
		const TaaParameters& P = A.ubo.param[0];
		const bool fixup_all = (c->desc.flags & TAA_FLAG_FIXUP_ALL) != 0;
		const bool need_fixup = fixup_all || A.mask.p != nullptr || P.mDynamicAntiGhosting;
		unsigned int *list = nullptr, *cnt = nullptr, *cnt_next = nullptr;
		if (need_fixup) {
			cudaError_t e = ensure_fix_buffers(c);
			if (e != cudaSuccess) return e;
			list = c->fix_list;
			cnt = c->fix_count + c->fix_parity;
			cnt_next = c->fix_count + (c->fix_parity ^ 1);
			c->fix_parity ^= 1;
		}
		const bool streaming = stream_supports(A);
		StreamPeers sp, *peers = nullptr;
		if (c->peers.on) {
			// the boundary rows are stored into the neighbours' buffers by the resolve kernel itself: only a call that the streaming kernel serves alone
			// can do that (an exact fix-up pass would rewrite pixels the neighbour already holds)
			const int q = A.history_out.p == c->peers.own_hist[0] ? 0 : A.history_out.p == c->peers.own_hist[1] ? 1 : -1;
			if (!streaming || need_fixup || q < 0) {
				set_error(c, "taa_band_peers is set: the call must run on the streaming kernel alone (config 2 family, no mask, no dynamic anti-ghosting) and write one of the two registered history buffers");
				return cudaErrorNotSupported;
			}
			for (int sd = 0; sd < 2; ++sd) {
				const taa_band_peer& n = c->peers.side[sd];
				sp.nb_hist[sd] = c->peers.has[sd] ? (unsigned char*)n.history[q] : nullptr;
				sp.nb_pitch[sd] = n.row_pitch; sp.nb_y0[sd] = n.y0; sp.nb_band_rows[sd] = n.band_rows; sp.nb_flags[sd] = n.flags;
			}
			sp.flags = c->peers.flags; sp.halo = c->peers.halo; sp.q = q; sp.wait = c->peers.first ? 0 : 1;
			c->peers.first = false;
			peers = &sp;
		}
		cudaError_t e = streaming ? launch_resolve_stream(A, list, cnt, cnt_next, fixup_all, c->num_sms, c->hints, c->hint_phase++, peers, s)
		                          : launch_resolve_strip(A, list, cnt, cnt_next, fixup_all, s);
		if (c->hint_phase >= 3 * 1024) c->hint_phase -= 3 * 1024;
		if (e != cudaSuccess) return e;
		*launched = 1;
		if (need_fixup) {
			e = launch_resolve_fixup(A, list, cnt, A.result.p != nullptr, c->num_sms, s);
			if (e == cudaSuccess) *launched = 2;
		}
		return e;
	}
	c->last_was_tuned = false;
	cudaError_t e = launch_resolve_generic(A, !(c->desc.flags & TAA_FLAG_EXACT), s);
	if (e == cudaSuccess) *launched = 1;
	return e;
}

}  // namespace taa
