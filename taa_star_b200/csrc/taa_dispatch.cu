// taa_dispatch.cu — chooses the resolve kernel for a settings block (every switch of
// `Parameters` is uniform across a dispatch except the split-screen select).
#include "taa_ctx.h"

namespace taa {

cudaError_t dispatch_resolve(taa_ctx* c, const ResolveArgs& A, cudaStream_t s) {
	(void)c;
	return launch_resolve_generic(A, s);
}

}  // namespace taa
