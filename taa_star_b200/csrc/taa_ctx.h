// taa_ctx.h — the context object behind the C-ABI (internal).
#pragma once
#include <cuda_runtime.h>
#include <string>
#include "../../include/taa_b200.h"
#include "taa_kernels.h"

struct taa_ctx {
	taa_desc desc{};
	int num_sms = 0;
	unsigned int* d_status = nullptr;  // device status word written by the kernels (halo overflow)
	void* scratch[2] = {nullptr, nullptr};  // rgba16f out-res images for the unfused follow-on passes
	unsigned int* fix_list = nullptr;  // pixels the tuned kernel hands to the exact fix-up pass (one slot per band pixel)
	unsigned int* fix_count = nullptr; // two counters, used alternately: call n appends to [n & 1] and zeroes [(n + 1) & 1]
	int fix_parity = 0;
	unsigned int* hints = nullptr;     // streaming kernel: units that were slow in the previous call (three rotating buffers)
	int hint_phase = 0;
	// row-band sharding without a per-frame collective (taa_band_peers): side 0 = the band above, 1 = the band below
	struct {
		bool on = false, first = true;
		void* own_hist[2] = {nullptr, nullptr};
		taa_band_peer side[2] = {};
		bool has[2] = {false, false};
		uint32_t* flags = nullptr;
		int halo = 0;
	} peers;
	bool last_was_tuned = false;
	long long launches = 0;            // kernels launched through this context
	std::string last_error;
};

namespace taa {
void set_error(taa_ctx* c, const char* fmt, ...);
int cuda_fail(taa_ctx* c, cudaError_t e, const char* what);
int build_resolve_args(taa_ctx* c, const taa_resolve_images* im, const TaaUniforms* u, ResolveArgs& A);
int run_resolve(taa_ctx* c, const ResolveArgs& A, cudaStream_t s);
// picks the kernel for this settings block: a tuned variant when one covers it, else the generic one
cudaError_t dispatch_resolve(taa_ctx* c, const ResolveArgs& A, cudaStream_t s, int* launched);
}  // namespace taa
