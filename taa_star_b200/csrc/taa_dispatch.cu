// taa_dispatch.cu — chooses the resolve kernel for a settings block (every switch of
// `Parameters` is uniform across a dispatch except the split-screen select).
#include "taa_ctx.h"
#include <cstdlib>

namespace taa {

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
		static const bool tile = [] { const char* v = getenv("TAA_TUNED_VARIANT"); return v && v[0] == 't'; }();  // A/B aid: the 32x32-tile kernel
		unsigned int *list = nullptr, *cnt = nullptr, *cnt_next = nullptr;
		if (need_fixup) {
			cudaError_t e = ensure_fix_buffers(c);
			if (e != cudaSuccess) return e;
			list = c->fix_list;
			cnt = c->fix_count + c->fix_parity;
			cnt_next = c->fix_count + (c->fix_parity ^ 1);
			c->fix_parity ^= 1;
		}
		const bool streaming = !tile && stream_supports(A);
		if (streaming && !c->hints) {  // (a failed allocation only costs the ordering hint)
			if (cudaMalloc(&c->hints, stream_hint_bytes()) == cudaSuccess) cudaMemsetAsync(c->hints, 0, stream_hint_bytes(), s);
			else { c->hints = nullptr; cudaGetLastError(); }
		}
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
		cudaError_t e = tile ? launch_resolve_tuned(A, list, cnt, cnt_next, fixup_all, s)
		                     : streaming ? launch_resolve_stream(A, list, cnt, cnt_next, fixup_all, c->num_sms, c->hints, c->hint_phase++, peers, s)
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
