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
		cudaError_t e = ensure_fix_buffers(c);
		if (e != cudaSuccess) return e;
		unsigned int* cnt = c->fix_count + c->fix_parity;
		unsigned int* cnt_next = c->fix_count + (c->fix_parity ^ 1);
		c->fix_parity ^= 1;
		// TAA_TUNED_VARIANT=tile selects the 32x32-tile kernel (A/B aid); both honour the same contract
		static const bool tile = [] { const char* v = getenv("TAA_TUNED_VARIANT"); return v && v[0] == 't'; }();
		e = tile ? launch_resolve_tuned(A, c->fix_list, cnt, cnt_next, (c->desc.flags & TAA_FLAG_FIXUP_ALL) != 0, s)
		         : launch_resolve_strip(A, c->fix_list, cnt, cnt_next, (c->desc.flags & TAA_FLAG_FIXUP_ALL) != 0, s);
		if (e != cudaSuccess) return e;
		*launched = 1;
		e = launch_resolve_fixup(A, c->fix_list, cnt, A.result.p != nullptr, c->num_sms, s);
		if (e == cudaSuccess) *launched = 2;
		return e;
	}
	c->last_was_tuned = false;
	cudaError_t e = launch_resolve_generic(A, s);
	if (e == cudaSuccess) *launched = 1;
	return e;
}

}  // namespace taa
