// taa_abi.cu — the C-ABI core of include/taa_b200.h: context, resolve / frame entry points and the
// pure-host helpers that mirror source/taa.hpp (CasSetup call, defaults, jitter, matrices).
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>

#include "taa_ctx.h"

namespace taa {

thread_local std::string g_create_error;

void set_error(taa_ctx* c, const char* fmt, ...) {
	char buf[512];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof buf, fmt, ap);
	va_end(ap);
	if (c) c->last_error = buf; else g_create_error = buf;
}

int cuda_fail(taa_ctx* c, cudaError_t e, const char* what) {
	set_error(c, "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
	return TAA_E_CUDA;
}

static bool fill_img(Img& d, const taa_image& s, int default_rows) {
	d.p = (const unsigned char*)s.data;
	d.pitch = s.pitch_bytes;
	d.y0 = s.y0;
	d.rows = s.rows > 0 ? s.rows : default_rows;
	return s.data != nullptr;
}
static bool fill_imgw(ImgW& d, const taa_image& s, int default_rows) {
	d.p = (unsigned char*)s.data;
	d.pitch = s.pitch_bytes;
	d.y0 = s.y0;
	d.rows = s.rows > 0 ? s.rows : default_rows;
	return s.data != nullptr;
}

static int check_pitch(taa_ctx* c, const taa_image& im, int width, int bpt, const char* name) {
	if (!im.data) return TAA_OK;
	if (im.pitch_bytes < (int64_t)width * bpt || (im.pitch_bytes % bpt) != 0) {
		set_error(c, "image '%s': pitch %lld is smaller than %d texels of %d bytes or not a multiple of the texel size", name,
		          (long long)im.pitch_bytes, width, bpt);
		return TAA_E_INVALID_ARG;
	}
	if (((uintptr_t)im.data % bpt) != 0) {
		set_error(c, "image '%s': base pointer is not aligned to the texel size (%d)", name, bpt);
		return TAA_E_INVALID_ARG;
	}
	return TAA_OK;
}

int build_resolve_args(taa_ctx* c, const taa_resolve_images* im, const TaaUniforms* u, ResolveArgs& A) {
	if (!c || !im || !u) return TAA_E_INVALID_ARG;
	const taa_desc& d = c->desc;
	struct { const taa_image* im; int w; int bpt; const char* name; bool required; } checks[] = {
		{&im->color, d.in_width, 8, "color", true},         {&im->depth, d.in_width, 4, "depth", true},
		{&im->velocity, d.in_width, 8, "velocity", true},   {&im->history_in, d.out_width, 8, "history_in", true},
		{&im->history_out, d.out_width, 8, "history_out", true}, {&im->history_depth, d.in_width, 4, "history_depth", false},
		{&im->result, d.out_width, 8, "result", false},     {&im->debug, d.out_width, 8, "debug", false},
		{&im->segmask, d.out_width, 4, "segmask", false},   {&im->prev_segmask, d.out_width, 4, "prev_segmask", false},
		{&im->matid, d.in_width, 4, "matid", false},        {&im->prev_matid, d.in_width, 4, "prev_matid", false},
		{&im->uvnrm, d.in_width, 16, "uvnrm", false},       {&im->mask, d.out_width, 4, "mask", false},
	};
	for (auto& k : checks) {
		if (k.required && !k.im->data) { set_error(c, "image '%s' is required", k.name); return TAA_E_INVALID_ARG; }
		int r = check_pitch(c, *k.im, k.w, k.bpt, k.name);
		if (r != TAA_OK) return r;
	}
	if (im->history_in.data == im->history_out.data) { set_error(c, "history_in and history_out must not alias (taa.hpp:1013,1018)"); return TAA_E_INVALID_ARG; }
	for (int i = 0; i < 2; ++i) {
		const TaaParameters& p = u->param[i];
		if (i == 1 && !u->splitScreen) break;
		if (p.mDepthCulling && !im->history_depth.data) { set_error(c, "mDepthCulling needs history_depth (taa.comp:818)"); return TAA_E_INVALID_ARG; }
		if (p.mRayTraceAugment && !im->segmask.data) { set_error(c, "mRayTraceAugment needs segmask (taa.comp:959)"); return TAA_E_INVALID_ARG; }
	}
	fill_img(A.color, im->color, d.in_height);
	fill_img(A.depth, im->depth, d.in_height);
	fill_img(A.velocity, im->velocity, d.in_height);
	fill_img(A.history_in, im->history_in, d.out_height);
	fill_img(A.history_depth, im->history_depth, d.in_height);
	fill_img(A.prev_segmask, im->prev_segmask, d.out_height);
	fill_img(A.matid, im->matid, d.in_height);
	fill_img(A.prev_matid, im->prev_matid, d.in_height);
	fill_img(A.uvnrm, im->uvnrm, d.in_height);
	fill_imgw(A.history_out, im->history_out, d.out_height);
	fill_imgw(A.result, im->result, d.out_height);
	fill_imgw(A.debug, im->debug, d.out_height);
	fill_imgw(A.segmask, im->segmask, d.out_height);
	fill_imgw(A.mask, im->mask, d.out_height);
	A.in_w = d.in_width;
	A.in_h = d.in_height;
	A.out_w = d.out_width;
	A.out_h = d.out_height;
	A.band_y0 = d.band_y0;
	A.band_rows = d.band_rows;
	A.status = c->d_status;
	A.ubo = *u;
	A.final_img = ImgW{nullptr, 0, 0, 0};
	A.epilogue = 0;
	A.epilogue_k = 0.f;
	return TAA_OK;
}

int run_resolve(taa_ctx* c, const ResolveArgs& A, cudaStream_t s) {
	int launched = 0;
	int cur = -1;  // a process may hold contexts on several devices: launch on the context's
	if (cudaGetDevice(&cur) == cudaSuccess && cur != c->desc.device) {
		cudaError_t e = cudaSetDevice(c->desc.device);
		if (e != cudaSuccess) return cuda_fail(c, e, "cudaSetDevice");
	}
	cudaError_t e = dispatch_resolve(c, A, s, &launched);
	c->launches += launched;
	if (e != cudaSuccess) return cuda_fail(c, e, "taa resolve launch");
	return TAA_OK;
}

static int ensure_scratch(taa_ctx* c, int idx) {
	if (c->scratch[idx]) return TAA_OK;
	size_t bytes = (size_t)c->desc.out_width * c->desc.out_height * 8;
	cudaError_t e = cudaMalloc(&c->scratch[idx], bytes);
	if (e != cudaSuccess) return cuda_fail(c, e, "cudaMalloc(scratch)");
	return TAA_OK;
}

}  // namespace taa

using namespace taa;

extern "C" {

int taa_abi_version(void) { return TAA_B200_ABI_VERSION; }

const char* taa_status_string(int s) {
	switch (s) {
		case TAA_OK: return "TAA_OK";
		case TAA_E_INVALID_ARG: return "TAA_E_INVALID_ARG";
		case TAA_E_UNSUPPORTED: return "TAA_E_UNSUPPORTED";
		case TAA_E_CUDA: return "TAA_E_CUDA";
		case TAA_E_NCCL: return "TAA_E_NCCL";
		case TAA_E_HALO_OVERFLOW: return "TAA_E_HALO_OVERFLOW";
		case TAA_E_PEER_TIMEOUT: return "TAA_E_PEER_TIMEOUT";
		default: return "TAA_E_UNKNOWN";
	}
}

const char* taa_last_error_string(const taa_ctx* ctx) { return ctx ? ctx->last_error.c_str() : g_create_error.c_str(); }

int taa_create(taa_ctx** out_ctx, const taa_desc* desc) {
	if (!out_ctx) return TAA_E_INVALID_ARG;
	*out_ctx = nullptr;
	if (!desc || desc->struct_size != sizeof(taa_desc) || desc->abi_version != TAA_B200_ABI_VERSION) {
		set_error(nullptr, "taa_desc: struct_size/abi_version mismatch (expected %zu / %d)", sizeof(taa_desc), TAA_B200_ABI_VERSION);
		return TAA_E_INVALID_ARG;
	}
	if (desc->in_width <= 0 || desc->in_height <= 0 || desc->out_width <= 0 || desc->out_height <= 0 || desc->band_y0 < 0 ||
	    desc->band_rows <= 0 || desc->band_y0 + desc->band_rows > desc->out_height) {
		set_error(nullptr, "taa_desc: invalid sizes in %dx%d out %dx%d band [%d,+%d)", desc->in_width, desc->in_height, desc->out_width,
		          desc->out_height, desc->band_y0, desc->band_rows);
		return TAA_E_INVALID_ARG;
	}
	int ndev = 0;
	cudaError_t e = cudaGetDeviceCount(&ndev);
	if (e != cudaSuccess || ndev == 0) {
		set_error(nullptr, "no CUDA device: %s — this library has no CPU path", e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
		return TAA_E_CUDA;
	}
	int dev = desc->device;
	if (dev < 0) { e = cudaGetDevice(&dev); if (e != cudaSuccess) return cuda_fail(nullptr, e, "cudaGetDevice"); }
	if (dev >= ndev) { set_error(nullptr, "device %d out of range (%d devices)", dev, ndev); return TAA_E_INVALID_ARG; }
	e = cudaSetDevice(dev);
	if (e != cudaSuccess) return cuda_fail(nullptr, e, "cudaSetDevice");
	cudaDeviceProp prop;
	e = cudaGetDeviceProperties(&prop, dev);
	if (e != cudaSuccess) return cuda_fail(nullptr, e, "cudaGetDeviceProperties");
	if (prop.major != 10) {
		set_error(nullptr, "device %d is sm_%d%d; this library only carries sm_100a code", dev, prop.major, prop.minor);
		return TAA_E_UNSUPPORTED;
	}
	taa_ctx* c = new (std::nothrow) taa_ctx();
	if (!c) return TAA_E_INVALID_ARG;
	c->desc = *desc;
	c->desc.device = dev;
	c->num_sms = prop.multiProcessorCount;
	e = cudaMalloc(&c->d_status, sizeof(unsigned int));
	if (e == cudaSuccess) e = cudaMemset(c->d_status, 0, sizeof(unsigned int));
	if (e != cudaSuccess) { int r = cuda_fail(nullptr, e, "cudaMalloc(status)"); delete c; return r; }
	// the streaming kernel's scheduling hints (allocated here, not at the first resolve: that call may be part of a stream capture, where an
	// allocation is not allowed). A failure only costs the ordering hint.
	if (cudaMalloc(&c->hints, stream_hint_bytes()) == cudaSuccess) cudaMemset(c->hints, 0, stream_hint_bytes());
	else { c->hints = nullptr; cudaGetLastError(); }
	*out_ctx = c;
	return TAA_OK;
}

void taa_destroy(taa_ctx* c) {
	if (!c) return;
	cudaSetDevice(c->desc.device);
	if (c->d_status) cudaFree(c->d_status);
	for (void* p : c->scratch) if (p) cudaFree(p);
	if (c->fix_list) cudaFree(c->fix_list);
	if (c->fix_count) cudaFree(c->fix_count);
	if (c->hints) cudaFree(c->hints);
	delete c;
}

int taa_resolve_ex(taa_ctx* c, const taa_resolve_images* im, const TaaUniforms* u, void* stream) {
	ResolveArgs A;
	int r = build_resolve_args(c, im, u, A);
	if (r != TAA_OK) return r;
	return run_resolve(c, A, (cudaStream_t)stream);
}

int taa_resolve(taa_ctx* c, const void* color, const void* depth, const void* motion, const void* history_in, void* history_out,
                const TaaUniforms* params, void* stream) {
	if (!c) return TAA_E_INVALID_ARG;
	const taa_desc& d = c->desc;
	if (d.band_y0 != 0 || d.band_rows != d.out_height) { set_error(c, "taa_resolve is for whole-frame contexts; use taa_resolve_ex for bands"); return TAA_E_INVALID_ARG; }
	taa_resolve_images im;
	memset(&im, 0, sizeof im);
	im.color = {(void*)color, (int64_t)d.in_width * 8, 0, d.in_height};
	im.depth = {(void*)depth, (int64_t)d.in_width * 4, 0, d.in_height};
	im.velocity = {(void*)motion, (int64_t)d.in_width * 8, 0, d.in_height};
	im.history_in = {(void*)history_in, (int64_t)d.out_width * 8, 0, d.out_height};
	im.history_out = {history_out, (int64_t)d.out_width * 8, 0, d.out_height};
	return taa_resolve_ex(c, &im, params, stream);
}

static int make_post(taa_ctx* c, const taa_image* src, const taa_image* dbg, const taa_image* dst, PostImg& io) {
	if (!c || !src || !dst || !src->data || !dst->data) { set_error(c, "post pass: src and dst are required"); return TAA_E_INVALID_ARG; }
	const taa_desc& d = c->desc;
	if (d.band_y0 != 0 || d.band_rows != d.out_height) { set_error(c, "follow-on passes run on whole frames only"); return TAA_E_UNSUPPORTED; }
	int r = check_pitch(c, *src, d.out_width, 8, "src");
	if (r == TAA_OK) r = check_pitch(c, *dst, d.out_width, 8, "dst");
	if (r == TAA_OK && dbg) r = check_pitch(c, *dbg, d.out_width, 8, "debug");
	if (r != TAA_OK) return r;
	if (src->data == dst->data) { set_error(c, "post pass: src and dst must differ (stencil read)"); return TAA_E_INVALID_ARG; }
	fill_img(io.src, *src, d.out_height);
	taa_image none = {nullptr, 0, 0, 0};
	fill_img(io.debug, dbg ? *dbg : none, d.out_height);
	fill_imgw(io.dst, *dst, d.out_height);
	io.w = d.out_width;
	io.h = d.out_height;
	return TAA_OK;
}

int taa_sharpen(taa_ctx* c, const taa_image* src, const taa_image* dst, const TaaSharpenPush* pc, void* stream) {
	PostImg io;
	if (!pc) return TAA_E_INVALID_ARG;
	int r = make_post(c, src, nullptr, dst, io);
	if (r != TAA_OK) return r;
	cudaError_t e = launch_sharpen(io, pc->sharpeningFactor, (cudaStream_t)stream);
	if (e != cudaSuccess) return cuda_fail(c, e, "sharpen launch");
	c->launches++;
	return TAA_OK;
}
int taa_sharpen_cas(taa_ctx* c, const taa_image* src, const taa_image* dst, const TaaCasPush* pc, void* stream) {
	PostImg io;
	if (!pc) return TAA_E_INVALID_ARG;
	int r = make_post(c, src, nullptr, dst, io);
	if (r != TAA_OK) return r;
	cudaError_t e = launch_cas(io, *pc, (cudaStream_t)stream);
	if (e != cudaSuccess) return cuda_fail(c, e, "cas launch");
	c->launches++;
	return TAA_OK;
}
int taa_post_process(taa_ctx* c, const taa_image* src, const taa_image* debug, const taa_image* dst, const TaaPostProcessPush* pc, void* stream) {
	PostImg io;
	if (!pc) return TAA_E_INVALID_ARG;
	// (post_process.comp:69-75 never looks at the right-hand settings without a splitter: splitX < 0)
	if ((pc->debugL_show || (pc->splitX >= 0 && pc->debugR_show)) && !(debug && debug->data)) { set_error(c, "post_process: debug image required when debug*_show is set"); return TAA_E_INVALID_ARG; }
	int r = make_post(c, src, debug, dst, io);
	if (r != TAA_OK) return r;
	cudaError_t e = launch_post_process(io, *pc, (cudaStream_t)stream);
	if (e != cudaSuccess) return cuda_fail(c, e, "post_process launch");
	c->launches++;
	return TAA_OK;
}

static int make_fxaa(taa_ctx* c, const taa_image* src, const taa_image* seg, const taa_image* dst, const TaaFxaaPush* pc, PostImg& io) {
	if (!pc || !seg || !seg->data) { set_error(c, "fxaa: push constants and the segmentation mask are required"); return TAA_E_INVALID_ARG; }
	int r = make_post(c, src, nullptr, dst, io);
	if (r == TAA_OK) r = check_pitch(c, *seg, c->desc.out_width, 4, "segmask");
	if (r != TAA_OK) return r;
	fill_img(io.debug, *seg, c->desc.out_height);
	return TAA_OK;
}
int taa_fxaa_prepare(taa_ctx* c, const taa_image* src, const taa_image* dst, void* stream) {
	PostImg io;
	int r = make_post(c, src, nullptr, dst, io);
	if (r != TAA_OK) return r;
	cudaError_t e = launch_fxaa_prepare(io, (cudaStream_t)stream);
	if (e != cudaSuccess) return cuda_fail(c, e, "fxaa_prepare launch");
	c->launches++;
	return TAA_OK;
}
static int fxaa_impl(taa_ctx* c, const taa_image* src, const taa_image* seg, const taa_image* dst, const TaaFxaaPush* pc, bool prepared, void* stream) {
	PostImg io;
	int r = make_fxaa(c, src, seg, dst, pc, io);
	if (r != TAA_OK) return r;
	cudaError_t e = launch_fxaa(io, *pc, prepared, (cudaStream_t)stream);
	if (e != cudaSuccess) return cuda_fail(c, e, "fxaa launch");
	c->launches++;
	return TAA_OK;
}
int taa_fxaa(taa_ctx* c, const taa_image* src, const taa_image* seg, const taa_image* dst, const TaaFxaaPush* pc, void* stream) {
	return fxaa_impl(c, src, seg, dst, pc, true, stream);
}
int taa_fxaa_fused(taa_ctx* c, const taa_image* src, const taa_image* seg, const taa_image* dst, const TaaFxaaPush* pc, void* stream) {
	return fxaa_impl(c, src, seg, dst, pc, false, stream);
}

// render() after taa.comp (taa.hpp:1029-1169): [fxaa_prepare + fxaa] -> [sharpen | CAS] -> [post-process]. Each arrow of the reference is a
// full-frame image; here FXAA is one launch and the sharpener is evaluated inside the post-process launch.
int taa_frame(taa_ctx* c, const taa_resolve_images* images, const TaaUniforms* u, const taa_post_chain* chain, const taa_image* final_img, void* stream) {
	if (!c || !images || !u || !chain || !final_img || !final_img->data) { set_error(c, "taa_frame: NULL argument"); return TAA_E_INVALID_ARG; }
	const taa_desc& d = c->desc;
	const bool fxaa = chain->fxaa != 0, sharpen = chain->sharpener != 0, post = chain->postprocess != 0;
	if (chain->sharpener < 0 || chain->sharpener > 2) { set_error(c, "mSharpener must be 0, 1 or 2 (taa.hpp:1418)"); return TAA_E_INVALID_ARG; }
	if (fxaa && !images->segmask.data) { set_error(c, "taa_frame: FXAA needs the segmentation mask image (mRayTraceAugment)"); return TAA_E_INVALID_ARG; }
	if (post && (chain->pp.debugL_show || (chain->pp.splitX >= 0 && chain->pp.debugR_show)) && !images->debug.data) { set_error(c, "post_process: debug image required when debug*_show is set"); return TAA_E_INVALID_ARG; }
	taa_resolve_images im = *images;
	const int64_t pitch = (int64_t)d.out_width * 8;
	int stages = (fxaa ? 1 : 0) + ((sharpen || post) ? 1 : 0);  // launches after the resolve
	int next_scratch = 0;
	auto scratch = [&](taa_image& out) -> int {
		int r = ensure_scratch(c, next_scratch);
		if (r != TAA_OK) return r;
		out = {c->scratch[next_scratch], pitch, 0, d.out_height};
		next_scratch ^= 1;
		return TAA_OK;
	};
	// stage 0: resolve. Its screen result goes to the caller's image, or to `final` when nothing follows, or to scratch.
	const bool result_given = im.result.data != nullptr;
	auto prepare_result = [&]() -> int {
		if (result_given) return TAA_OK;
		if (stages == 0) { im.result = *final_img; return TAA_OK; }
		return scratch(im.result);
	};
	// ---- the fused chain: [sharpen | CAS] (+ an identity post-process) in the resolve's epilogue, one launch and no intermediate image ----
	// Admissible when post_process.comp would only copy (no zoom box, no splitter, no debug view), FXAA is off and the call takes a plain
	// variant of the streaming kernel with nothing left to the exact pass (stream_epilogue_ok). Without a sharpener an identity post-process
	// is a copy: the resolve then writes its screen result straight into `final`.
	const bool pp_identity = !post || (!chain->pp.zoom && chain->pp.splitX < 0 && !chain->pp.debugL_show);
	static const int fuse_env = [] { const char* v = getenv("TAA_FUSED_CHAIN"); return v ? atoi(v) : -1; }();  // 0 / 1: A/B aid (default: see below)
	if (!fxaa && pp_identity && (sharpen || post) && check_pitch(c, *final_img, d.out_width, 8, "final") == TAA_OK && !(c->desc.flags & TAA_FLAG_EXACT)) {
		taa_resolve_images fim = *images;
		if (!sharpen) {
			if (!fim.result.data) {  // (a caller that wants both images still gets the copy below)
				fim.result = *final_img;
				return taa_resolve_ex(c, &fim, u, stream);
			}
		} else {
			ResolveArgs A;
			int r = build_resolve_args(c, &fim, u, A);
			if (r != TAA_OK) return r;
			if (fuse_env != 0 && stream_epilogue_ok(A, (c->desc.flags & TAA_FLAG_FIXUP_ALL) != 0) && final_img->data != fim.result.data && final_img->data != fim.history_out.data) {
				fill_imgw(A.final_img, *final_img, d.out_height);
				A.epilogue = chain->sharpener;
				if (chain->sharpener == 1) A.epilogue_k = chain->sharpen.sharpeningFactor;
				else memcpy(&A.epilogue_k, &chain->cas.const1[0], 4);
				return run_resolve(c, A, (cudaStream_t)stream);
			}
		}
	}
	int r = prepare_result();
	if (r != TAA_OK) return r;
	r = taa_resolve_ex(c, &im, u, stream);
	if (r != TAA_OK) return r;
	taa_image last = im.result;
	if (fxaa) {
		taa_image dst = *final_img;
		if (--stages > 0) { r = scratch(dst); if (r != TAA_OK) return r; }
		r = taa_fxaa_fused(c, &last, &im.segmask, &dst, &chain->fxaa_pc, stream);
		if (r != TAA_OK) return r;
		last = dst;
	}
	if (sharpen && post) {  // [sharpen | CAS] evaluated inside the post-process pass: one launch, no intermediate image
		PostImg io;
		r = make_post(c, &last, images->debug.data ? &images->debug : nullptr, final_img, io);
		if (r != TAA_OK) return r;
		cudaError_t e = launch_sharpen_post(io, chain->sharpener, chain->sharpen.sharpeningFactor, chain->cas, chain->pp, (cudaStream_t)stream);
		if (e != cudaSuccess) return cuda_fail(c, e, "sharpen + post_process launch");
		c->launches++;
	} else if (sharpen) {
		r = chain->sharpener == 1 ? taa_sharpen(c, &last, final_img, &chain->sharpen, stream) : taa_sharpen_cas(c, &last, final_img, &chain->cas, stream);
		if (r != TAA_OK) return r;
	} else if (post) {
		r = taa_post_process(c, &last, images->debug.data ? &images->debug : nullptr, final_img, &chain->pp, stream);
		if (r != TAA_OK) return r;
	} else if (last.data != final_img->data) {  // caller wanted both: plain copy (the reference would blit, taa.hpp:1169)
		cudaError_t e = cudaMemcpy2DAsync(final_img->data, final_img->pitch_bytes, last.data, last.pitch_bytes, (size_t)d.out_width * 8, d.out_height,
		                                  cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
		if (e != cudaSuccess) return cuda_fail(c, e, "copy result -> final");
	}
	return TAA_OK;
}

int taa_poll_status(taa_ctx* c, void* stream) {
	if (!c) return TAA_E_INVALID_ARG;
	unsigned int h = 0;
	cudaError_t e = cudaMemcpyAsync(&h, c->d_status, sizeof h, cudaMemcpyDeviceToHost, (cudaStream_t)stream);
	if (e == cudaSuccess) e = cudaMemsetAsync(c->d_status, 0, sizeof h, (cudaStream_t)stream);
	if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t)stream);
	if (e != cudaSuccess) return cuda_fail(c, e, "taa_poll_status");
	if (h & 1u) {
		set_error(c, "a read left the rows held by a band buffer (status word 0x%x: 0x10 colour, 0x20 velocity, 0x40 depth apron too small; none of them: history halo too small for this motion)", h);
		return TAA_E_HALO_OVERFLOW;
	}
	if (h & 2u) { set_error(c, "a neighbour band's boundary rows did not arrive (taa_band_peers)"); return TAA_E_PEER_TIMEOUT; }
	return TAA_OK;
}

// ---- row bands without a per-frame collective (see include/taa_b200.h and "PEER variants" in taa_resolve_stream.cu) ----
int taa_band_peers(taa_ctx* c, const taa_band_peer* up, const taa_band_peer* down, void* own_history0, void* own_history1, uint32_t* own_flags, int32_t halo_rows) {
	if (!c) return TAA_E_INVALID_ARG;
	if (!up && !down) { c->peers.on = false; return TAA_OK; }
	if (!own_history0 || !own_history1 || own_history0 == own_history1 || !own_flags || halo_rows <= 0 || halo_rows > c->desc.band_rows) {
		set_error(c, "taa_band_peers: two distinct own history buffers, a flag block and 0 < halo_rows <= band_rows are required");
		return TAA_E_INVALID_ARG;
	}
	const taa_band_peer* in[2] = {up, down};
	for (int sd = 0; sd < 2; ++sd) {
		c->peers.has[sd] = in[sd] != nullptr;
		if (!in[sd]) continue;
		const taa_band_peer& n = *in[sd];
		if (!n.history[0] || !n.history[1] || !n.flags || n.band_rows < halo_rows || n.row_pitch < (int64_t)c->desc.out_width * 8 ||
		    (((uintptr_t)n.history[0] | (uintptr_t)n.history[1] | (uintptr_t)n.row_pitch) & 15u)) {
			set_error(c, "taa_band_peers: neighbour %d needs two mapped history buffers (16-byte aligned base and pitch), a mapped flag block and band_rows >= halo_rows", sd);
			return TAA_E_INVALID_ARG;
		}
		c->peers.side[sd] = n;
	}
	int cur = -1;
	if (cudaGetDevice(&cur) == cudaSuccess && cur != c->desc.device) cudaSetDevice(c->desc.device);
	cudaError_t e = cudaMemset(own_flags, 0, TAA_BAND_FLAG_WORDS * sizeof(uint32_t));
	if (e == cudaSuccess) e = cudaDeviceSynchronize();
	if (e != cudaSuccess) return cuda_fail(c, e, "taa_band_peers");
	c->peers.own_hist[0] = own_history0;
	c->peers.own_hist[1] = own_history1;
	c->peers.flags = own_flags;
	c->peers.halo = halo_rows;
	c->peers.first = true;
	c->peers.on = true;
	return TAA_OK;
}

void* taa_device_alloc(size_t bytes) {
	void* p = nullptr;
	if (cudaMalloc(&p, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
	cudaMemset(p, 0, bytes);
	return p;
}
void taa_device_free(void* p) { if (p) cudaFree(p); }
static_assert(sizeof(cudaIpcMemHandle_t) == TAA_IPC_HANDLE_BYTES, "IPC handle size");
int taa_ipc_export(const void* p, void* handle_out) {
	if (!p || !handle_out) return TAA_E_INVALID_ARG;
	cudaIpcMemHandle_t h;
	cudaError_t e = cudaIpcGetMemHandle(&h, const_cast<void*>(p));
	if (e != cudaSuccess) return cuda_fail(nullptr, e, "cudaIpcGetMemHandle");
	memcpy(handle_out, &h, sizeof h);
	return TAA_OK;
}
int taa_ipc_open(const void* handle, void** mapped_out) {
	if (!handle || !mapped_out) return TAA_E_INVALID_ARG;
	cudaIpcMemHandle_t h;
	memcpy(&h, handle, sizeof h);
	cudaError_t e = cudaIpcOpenMemHandle(mapped_out, h, cudaIpcMemLazyEnablePeerAccess);
	if (e != cudaSuccess) return cuda_fail(nullptr, e, "cudaIpcOpenMemHandle");
	return TAA_OK;
}
int taa_ipc_close(void* mapped) {
	if (!mapped) return TAA_OK;
	cudaError_t e = cudaIpcCloseMemHandle(mapped);
	return e == cudaSuccess ? TAA_OK : cuda_fail(nullptr, e, "cudaIpcCloseMemHandle");
}

long long taa_launch_count(const taa_ctx* c) { return c ? c->launches : 0; }

long long taa_fixup_pixels(taa_ctx* c, void* stream) {
	if (!c) return TAA_E_INVALID_ARG;
	if (!c->fix_count || !c->last_was_tuned) return 0;
	unsigned int h = 0;
	cudaError_t e = cudaMemcpyAsync(&h, c->fix_count + (c->fix_parity ^ 1), sizeof h, cudaMemcpyDeviceToHost, (cudaStream_t)stream);
	if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t)stream);
	if (e != cudaSuccess) return cuda_fail(c, e, "taa_fixup_pixels");
	return (long long)h;
}

// ================================ pure host helpers ================================================

static uint32_t fbits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

// fp32 -> fp16 the way AU1_AH1_AF1 does on the CPU (shaders/ffx_a.h:470-544): truncate, clamp to 65504
static uint32_t cas_half_bits(float f) {
	uint32_t u = fbits(f), sign = (u >> 16) & 0x8000u, m = u & 0x7fffffu;
	int e = (int)((u >> 23) & 0xffu);
	if (e < 103) return sign;
	if (e < 113) return sign | ((1u << (e - 103)) + (m >> (126 - e)));
	if (e < 143) return sign | (((uint32_t)(e - 112) << 10) + (m >> 13));
	return sign | 0x7bffu;
}

// CasSetup(const0, const1, sharpness, w, h, w, h) — shaders/ffx_cas.h:375-394 as called at taa.hpp:965
void taa_cas_setup(TaaCasPush* out, float sharpness, float w, float h) {
	if (!out) return;
	out->const0[0] = fbits(w * (1.0f / w));
	out->const0[1] = fbits(h * (1.0f / h));
	out->const0[2] = fbits(0.5f * w * (1.0f / w) - 0.5f);
	out->const0[3] = fbits(0.5f * h * (1.0f / h) - 0.5f);
	float s = fminf(1.0f, fmaxf(0.0f, sharpness));
	float sharp = -(1.0f / (5.0f * s + (-8.0f * s + 8.0f)));
	out->const1[0] = fbits(sharp);
	out->const1[1] = cas_half_bits(sharp) + (cas_half_bits(0.0f) << 16);
	out->const1[2] = fbits(8.0f * w * (1.0f / w));
	out->const1[3] = 0;
}

void taa_parameters_default(TaaParameters* p) {  // taa.hpp:31-76
	if (!p) return;
	memset(p, 0, sizeof *p);
	p->mAlpha = 0.05f;
	p->mColorClampingOrClipping = 1;
	p->mUnjitterFactor = 1.0f;
	p->mVarClipGamma = 1.0f;
	p->mMinAlpha = 1.0f - 0.97f;
	p->mMaxAlpha = 1.0f - 0.88f;
	p->mRejectionAlpha = 1.0f;
	p->mUseVelocityVectors = 1;
	p->mNoiseFactor = 1.f / 510.f;
	p->mVelBasedAlphaMax = 0.2f;
	p->mVelBasedAlphaFactor = 1.f / 40.f;
	p->mRayTraceAugmentFlags = 0xffffffffu & ~(uint32_t)(TAA_RTFLAG_ALL | TAA_RTFLAG_FXD);
	p->mRayTraceAugment_WNrm = 0.5f;
	p->mRayTraceAugment_WDpt = 0.015f;
	p->mRayTraceAugment_WMId = 0.25f;
	p->mRayTraceAugment_WLum = 0.5f;
	p->mRayTraceAugment_Thresh = 0.5f;
	p->mRayTraceHistoryCount = -1;
	p->mDebugMask[0] = p->mDebugMask[1] = p->mDebugMask[2] = 1.0f;
	p->mDebugScale = 1.0f;
}

void taa_uniforms_default(TaaUniforms* u) {
	if (!u) return;
	memset(u, 0, sizeof *u);
	for (int i = 0; i < 4; ++i) u->mHistoryViewProjMatrix[i * 5] = u->mInverseViewProjMatrix[i * 5] = 1.0f;
	taa_parameters_default(&u->param[0]);
	taa_parameters_default(&u->param[1]);
}

void taa_fxaa_default(TaaFxaaPush* pc, int32_t w, int32_t h) {  // taa.hpp:93-99, 953
	if (!pc) return;
	memset(pc, 0, sizeof *pc);
	pc->fxaaQualityRcpFrame[0] = 1.f / (float)w;
	pc->fxaaQualityRcpFrame[1] = 1.f / (float)h;
	pc->fxaaQualitySubpix = 0.75f;
	pc->fxaaQualityEdgeThreshold = 0.116f;
	pc->fxaaQualityEdgeThresholdMin = 0.0833f;
}

void taa_postprocess_default(TaaPostProcessPush* pp, int32_t w, int32_t h) {  // taa.hpp:101-111, 352-359
	if (!pp) return;
	memset(pp, 0, sizeof *pp);
	int dZoomSrc = (int)roundf((float)w / 96.f), dZoomDst = (int)roundf((float)w / 9.6f), dZoomBrd = (int)roundf((float)w / (96.f * 2.f));
	pp->zoomSrcLTWH[0] = (w - dZoomSrc) / 2; pp->zoomSrcLTWH[1] = (h - dZoomSrc) / 2; pp->zoomSrcLTWH[2] = dZoomSrc; pp->zoomSrcLTWH[3] = dZoomSrc;
	pp->zoomDstLTWH[0] = w - dZoomDst - dZoomBrd; pp->zoomDstLTWH[1] = dZoomBrd; pp->zoomDstLTWH[2] = dZoomDst; pp->zoomDstLTWH[3] = dZoomDst;
	pp->debugL_mask[0] = pp->debugL_mask[1] = pp->debugL_mask[2] = 1.f;
	pp->debugR_mask[0] = pp->debugR_mask[1] = pp->debugR_mask[2] = 1.f;
	pp->zoom = 0;
	pp->showZoomBox = 1;
	pp->splitX = -1;
}

float taa_halton(int32_t i, int32_t b) {  // helper_functions.hpp:9-17
	float f = 1.0f, r = 0.0f;
	for (; i > 0; i /= b) {
		f = f / (float)b;
		r = r + f * (float)(i % b);
	}
	return r;
}

int taa_jitter_offset_for_frame(const taa_jitter_settings* s, int32_t in_w, int32_t in_h, int64_t frame, float out_ndc[2]) {
	if (!s || !out_ndc || in_w <= 0 || in_h <= 0 || frame < 0) return TAA_E_INVALID_ARG;
	const float px[2] = {2.0f / (float)in_w, 2.0f / (float)in_h};  // sPxSizeNDC, taa.hpp:155
	float pat[16][2];
	int n;
	static const float quad[4][2] = {{-.25f, -.25f}, {.25f, -.25f}, {.25f, .25f}, {-.25f, .25f}};   // taa.hpp:158-163
	static const float helix[4][2] = {{-.25f, -.25f}, {.25f, .25f}, {.25f, -.25f}, {-.25f, .25f}};  // taa.hpp:164-169
	switch (s->mSampleDistribution) {
		case 0: n = 4; for (int i = 0; i < n; ++i) for (int k = 0; k < 2; ++k) pat[i][k] = px[k] * quad[i][k]; break;
		case 1: n = 4; for (int i = 0; i < n; ++i) for (int k = 0; k < 2; ++k) pat[i][k] = px[k] * helix[i][k]; break;
		case 2:
		case 3:
			n = s->mSampleDistribution == 2 ? 8 : 16;  // halton_2_3<N>, helper_functions.hpp:19-26
			for (int i = 0; i < n; ++i) { pat[i][0] = px[0] * (taa_halton(i + 1, 2) - 0.5f); pat[i][1] = px[1] * (taa_halton(i + 1, 3) - 0.5f); }
			break;
		case 4: {  // taa.hpp:173-179
			n = 16;
			const float eighth = 1.f / 8.f;
			for (int i = 0; i < 16; ++i) { pat[i][0] = px[0] * ((float)(2 * (i % 4) - 3) * eighth); pat[i][1] = px[1] * ((float)(2 * (i / 4) - 3) * eighth); }
			break;
		}
		case 5: n = 0; break;
		default: return TAA_E_INVALID_ARG;
	}
	if (s->mJitterSlowMotion > 1) frame /= s->mJitterSlowMotion;       // taa.hpp:219
	if (s->mFixedJitterIndex >= 0) frame = s->mFixedJitterIndex;      // taa.hpp:220
	float pos[2];
	if (s->mSampleDistribution == 5) {  // custom offsets are stored in pixel units and scaled here (taa.hpp:207-211, 223)
		n = s->mDebugSampleOffsetsCount;
		if (n <= 0 || !s->mDebugSampleOffsets) return TAA_E_INVALID_ARG;
		int idx = (int)(frame % n);
		pos[0] = s->mDebugSampleOffsets[2 * idx] * px[0];
		pos[1] = s->mDebugSampleOffsets[2 * idx + 1] * px[1];
	} else {
		int idx = (int)(frame % n);
		pos[0] = pat[idx][0];
		pos[1] = pat[idx][1];
	}
	if (s->mJitterRotateDegrees != 0.f) {  // taa.hpp:225-229
		float rad = s->mJitterRotateDegrees * 0.01745329251994329576923690768489f;
		float sn = sinf(rad), cs = cosf(rad);
		float x = pos[0] * cs - pos[1] * sn, y = pos[0] * sn + pos[1] * cs;
		pos[0] = x;
		pos[1] = y;
	}
	out_ndc[0] = pos[0] * s->mJitterExtraScale;
	out_ndc[1] = pos[1] * s->mJitterExtraScale;
	return n;
}

// column-major 4x4 helpers (glm conventions)
static void mat4_mul(const float a[16], const float b[16], float o[16]) {
	float r[16];
	for (int c = 0; c < 4; ++c)
		for (int rr = 0; rr < 4; ++rr) {
			float s = 0.f;
			for (int k = 0; k < 4; ++k) s += a[k * 4 + rr] * b[c * 4 + k];
			r[c * 4 + rr] = s;
		}
	memcpy(o, r, sizeof r);
}
// glm::inverse(mat4) as the reference computes it (taa.hpp:993 through the GLM 0.9.9.9 it vendors,
// gears_vk/external/universal/include/glm/detail/func_matrix.inl:294-352): the same fp32 operations in the same order, so the uploaded
// mInverseViewProjMatrix is bit-identical (tests/golden/matrix_golden.json, made with that GLM). m and out are column-major: M(c, r) = m[4c + r].
static int mat4_inverse(const float m[16], float out[16]) {
#define M(c, r) m[4 * (c) + (r)]
	const float c00 = M(2, 2) * M(3, 3) - M(3, 2) * M(2, 3), c02 = M(1, 2) * M(3, 3) - M(3, 2) * M(1, 3), c03 = M(1, 2) * M(2, 3) - M(2, 2) * M(1, 3);
	const float c04 = M(2, 1) * M(3, 3) - M(3, 1) * M(2, 3), c06 = M(1, 1) * M(3, 3) - M(3, 1) * M(1, 3), c07 = M(1, 1) * M(2, 3) - M(2, 1) * M(1, 3);
	const float c08 = M(2, 1) * M(3, 2) - M(3, 1) * M(2, 2), c10 = M(1, 1) * M(3, 2) - M(3, 1) * M(1, 2), c11 = M(1, 1) * M(2, 2) - M(2, 1) * M(1, 2);
	const float c12 = M(2, 0) * M(3, 3) - M(3, 0) * M(2, 3), c14 = M(1, 0) * M(3, 3) - M(3, 0) * M(1, 3), c15 = M(1, 0) * M(2, 3) - M(2, 0) * M(1, 3);
	const float c16 = M(2, 0) * M(3, 2) - M(3, 0) * M(2, 2), c18 = M(1, 0) * M(3, 2) - M(3, 0) * M(1, 2), c19 = M(1, 0) * M(2, 2) - M(2, 0) * M(1, 2);
	const float c20 = M(2, 0) * M(3, 1) - M(3, 0) * M(2, 1), c22 = M(1, 0) * M(3, 1) - M(3, 0) * M(1, 1), c23 = M(1, 0) * M(2, 1) - M(2, 0) * M(1, 1);
	const float fac0[4] = {c00, c00, c02, c03}, fac1[4] = {c04, c04, c06, c07}, fac2[4] = {c08, c08, c10, c11};
	const float fac3[4] = {c12, c12, c14, c15}, fac4[4] = {c16, c16, c18, c19}, fac5[4] = {c20, c20, c22, c23};
	const float v0[4] = {M(1, 0), M(0, 0), M(0, 0), M(0, 0)}, v1[4] = {M(1, 1), M(0, 1), M(0, 1), M(0, 1)};
	const float v2[4] = {M(1, 2), M(0, 2), M(0, 2), M(0, 2)}, v3[4] = {M(1, 3), M(0, 3), M(0, 3), M(0, 3)};
	const float sa[4] = {1.f, -1.f, 1.f, -1.f}, sb[4] = {-1.f, 1.f, -1.f, 1.f};
	float inv[16];
	for (int i = 0; i < 4; ++i) {
		const float i0 = (v1[i] * fac0[i] - v2[i] * fac1[i]) + v3[i] * fac2[i];
		const float i1 = (v0[i] * fac0[i] - v2[i] * fac3[i]) + v3[i] * fac4[i];
		const float i2 = (v0[i] * fac1[i] - v1[i] * fac3[i]) + v3[i] * fac5[i];
		const float i3 = (v0[i] * fac2[i] - v1[i] * fac4[i]) + v2[i] * fac5[i];
		inv[0 + i] = i0 * sa[i]; inv[4 + i] = i1 * sb[i]; inv[8 + i] = i2 * sa[i]; inv[12 + i] = i3 * sb[i];
	}
	const float d0 = M(0, 0) * inv[0], d1 = M(0, 1) * inv[4], d2 = M(0, 2) * inv[8], d3 = M(0, 3) * inv[12];  // m[0] * Row0
	const float det = (d0 + d1) + (d2 + d3);
	const float ood = 1.0f / det;
	for (int i = 0; i < 16; ++i) out[i] = inv[i] * ood;
#undef M
	return det == 0.0f ? TAA_E_INVALID_ARG : TAA_OK;  // (GLM returns the infinities; so does `out`)
}

void taa_jittered_projection(const float proj[16], float jx, float jy, float out[16]) {  // taa.hpp:248
	float t[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, jx, jy, 0, 1};
	mat4_mul(t, proj, out);
}

int taa_reprojection_matrices(const float pc[16], const float vc[16], const float pp[16], const float vp[16], float out_inv[16], float out_hist[16]) {
	if (!pc || !vc || !pp || !vp || !out_inv || !out_hist) return TAA_E_INVALID_ARG;
	float pv[16];
	mat4_mul(pc, vc, pv);
	int r = mat4_inverse(pv, out_inv);  // taa.hpp:993
	mat4_mul(pp, vp, out_hist);         // taa.hpp:994
	return r;
}

void* taa_host_alloc(size_t bytes) {
	void* p = nullptr;
	if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) return nullptr;
	return p;
}
void taa_host_free(void* p) { if (p) cudaFreeHost(p); }

}  // extern "C"
