// taa_invokee.cu — the host side of `template<size_t CF> class taa : gvk::invokee` (source/taa.hpp:26-1427)
// without Vulkan / ImGui: settings surface, jitter sequence, history ring, per-frame uniforms and the
// dispatch order of render(). The G-buffers stay owned by the caller (taa.hpp:279-284); this class owns
// result / history / temp[2] / debug / post-process / seg-mask images x CF (taa.hpp:294-340).
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "taa_ctx.h"

struct taa_invokee {
	int CF = 3;
	int device = 0;
	uint32_t flags = 0;
	taa_ctx* ctx = nullptr;
	std::string last_error;

	// "Settings, which can be modified via ImGui" (taa.hpp:1350-1424)
	TaaParameters mParameters[2];
	taa_invokee_settings S{};
	TaaPostProcessPush mPostProcessPushConstants{};
	TaaSharpenPush mSharpenerPushConstants{1.f};
	TaaCasPush mCasPushConstants{};
	TaaFxaaPush mFxaaPushConstants{};
	TaaUniforms mTaaUniforms{};
	std::vector<float> debugOffsets{0.f, 0.f};

	bool mUpsampling = false;
	int in_w = 0, in_h = 0, out_w = 0, out_h = 0;
	std::vector<taa_source_views> src;                  // mSrcColor/... per frame in flight (taa.hpp:1357-1362)
	std::vector<void*> img[7];                          // owned images, indexed by TAA_IMG_* then slot
	std::vector<float> mHistoryProjMatrices, mHistoryViewMatrices;  // CF x 16 (taa.hpp:1375-1376)

	// function-local statics of the reference, made members
	bool isVeryFirstFrame = true;       // taa.hpp:989
	float prevSharpenFactor = -1.f;     // taa.hpp:961
	TaaParameters oldParams[2];         // taa.hpp:915
	bool oldParamsValid = false;        // taa.hpp:916
	long long lastJitterIndex = 0;      // taa.hpp:908
	size_t numJitterSamples = 0;

	// timing (replaces the Vulkan timestamp queries, taa.hpp:997,1171,367-374)
	std::vector<cudaEvent_t> evStart, evStop;
	std::vector<char> evValid;
	float lastDurationMs = 0.f;

	// host-buffer pipeline (CF frames in flight, main.cpp:341)
	bool pipe_ready = false;
	cudaStream_t sUp = nullptr, sCompute = nullptr, sDown = nullptr;
	std::vector<cudaEvent_t> evUploaded, evComputed, evDone;
	std::vector<char> slotBusy;
	long long h2dBytes = 0;             // bytes taa_invokee_frame_host has uploaded so far (depth only when a kernel of the frame reads it)
	std::vector<taa_source_views> dsrc;  // device copies of host G-buffers
	bool owns_dsrc = false;
};

namespace {

void inv_error(taa_invokee* t, const char* fmt, ...) {
	char buf[512];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof buf, fmt, ap);
	va_end(ap);
	if (t) t->last_error = buf;
}
int inv_cuda(taa_invokee* t, cudaError_t e, const char* what) {
	inv_error(t, "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
	return TAA_E_CUDA;
}

// the host-frame pipeline (streams, events, device copies of the G-buffers) is sized for one input size: torn down with the images
void free_pipe(taa_invokee* t) {
	if (!t->pipe_ready) return;
	cudaStreamSynchronize(t->sUp); cudaStreamSynchronize(t->sCompute); cudaStreamSynchronize(t->sDown);
	for (auto e : t->evUploaded) cudaEventDestroy(e);
	for (auto e : t->evComputed) cudaEventDestroy(e);
	for (auto e : t->evDone) cudaEventDestroy(e);
	t->evUploaded.clear(); t->evComputed.clear(); t->evDone.clear();
	cudaStreamDestroy(t->sUp); cudaStreamDestroy(t->sCompute); cudaStreamDestroy(t->sDown);
	t->sUp = t->sCompute = t->sDown = nullptr;
	t->slotBusy.clear();
	t->pipe_ready = false;
}

void free_images(taa_invokee* t) {
	for (auto& v : t->img) {
		for (void* p : v) if (p) cudaFree(p);
		v.clear();
	}
	if (t->owns_dsrc) {
		for (auto& s : t->dsrc) {
			cudaFree((void*)s.color); cudaFree((void*)s.depth); cudaFree((void*)s.velocity);
			if (s.uvnrm) cudaFree((void*)s.uvnrm);
			if (s.matid) cudaFree((void*)s.matid);
		}
		t->dsrc.clear();
		t->owns_dsrc = false;
	}
}

taa_image whole(void* p, int w, int h, int bpt) { return taa_image{p, (int64_t)w * bpt, 0, h}; }

int slot_of(const taa_invokee* t, int64_t frame) { return (int)(frame % t->CF); }  // window::in_flight_index_for_frame, window.hpp:177

}  // namespace

extern "C" {

int taa_invokee_create(taa_invokee** out, int32_t concurrent_frames, int32_t device, uint32_t flags) {
	if (!out) return TAA_E_INVALID_ARG;
	*out = nullptr;
	if (concurrent_frames < 2) return TAA_E_INVALID_ARG;  // static_assert(CF > 1), taa.hpp:1005
	taa_invokee* t = new (std::nothrow) taa_invokee();
	if (!t) return TAA_E_INVALID_ARG;
	t->CF = concurrent_frames;
	t->flags = flags;
	if (device < 0) { if (cudaGetDevice(&device) != cudaSuccess) device = 0; }
	t->device = device;
	taa_parameters_default(&t->mParameters[0]);
	taa_parameters_default(&t->mParameters[1]);
	taa_uniforms_default(&t->mTaaUniforms);
	t->S.mTaaEnabled = 1;             // taa.hpp:1351
	t->S.mPostProcessEnabled = 1;     // taa.hpp:1352
	t->S.mResetHistory = 0;
	t->S.mSplitScreen = 0;
	t->S.mSplitX = 0;
	t->S.mSharpener = 0;              // taa.hpp:1418
	t->S.mSharpenFactor = 0.5f;       // taa.hpp:1419
	t->S.mResetHistoryOnChange = 1;   // taa.hpp:1421
	t->S.jitter.mSampleDistribution = 1;   // taa.hpp:1353
	t->S.jitter.mFixedJitterIndex = -1;
	t->S.jitter.mJitterExtraScale = 1.0f;
	t->S.jitter.mJitterSlowMotion = 1;
	t->S.jitter.mJitterRotateDegrees = 0.f;
	t->S.jitter.mDebugSampleOffsets = t->debugOffsets.data();
	t->S.jitter.mDebugSampleOffsetsCount = 1;
	t->mFxaaPushConstants.fxaaQualitySubpix = 0.75f;
	t->mFxaaPushConstants.fxaaQualityEdgeThreshold = 0.116f;
	t->mFxaaPushConstants.fxaaQualityEdgeThresholdMin = 0.0833f;
	t->mHistoryProjMatrices.assign((size_t)t->CF * 16, 0.f);
	t->mHistoryViewMatrices.assign((size_t)t->CF * 16, 0.f);
	*out = t;
	return TAA_OK;
}

void taa_invokee_destroy(taa_invokee* t) {
	if (!t) return;
	cudaSetDevice(t->device);
	free_pipe(t);
	for (auto e : t->evStart) cudaEventDestroy(e);
	for (auto e : t->evStop) cudaEventDestroy(e);
	free_images(t);
	if (t->ctx) taa_destroy(t->ctx);
	delete t;
}

const char* taa_invokee_last_error(const taa_invokee* t) { return t ? t->last_error.c_str() : ""; }

// set_source_image_views — taa.hpp:263-360
int taa_invokee_set_source_image_views(taa_invokee* t, int32_t target_w, int32_t target_h, int32_t in_w, int32_t in_h, const taa_source_views* views) {
	if (!t || target_w <= 0 || target_h <= 0 || in_w <= 0 || in_h <= 0) return TAA_E_INVALID_ARG;
	cudaError_t e = cudaSetDevice(t->device);
	if (e != cudaSuccess) return inv_cuda(t, e, "cudaSetDevice");
	free_pipe(t);  // (a host-frame pipeline of the old size: frames in flight are waited for, ensure_pipe() rebuilds it)
	free_images(t);
	if (t->ctx) { taa_destroy(t->ctx); t->ctx = nullptr; }
	taa_desc d{};
	d.struct_size = sizeof d;
	d.abi_version = TAA_B200_ABI_VERSION;
	d.in_width = in_w; d.in_height = in_h; d.out_width = target_w; d.out_height = target_h;
	d.band_y0 = 0; d.band_rows = target_h; d.device = t->device; d.flags = t->flags;
	int r = taa_create(&t->ctx, &d);
	if (r != TAA_OK) { inv_error(t, "taa_create: %s", taa_last_error_string(nullptr)); return r; }
	t->in_w = in_w; t->in_h = in_h; t->out_w = target_w; t->out_h = target_h;
	t->mUpsampling = (target_w != in_w || target_h != in_h);  // taa.hpp:292
	t->src.assign(t->CF, taa_source_views{});
	if (views) for (int i = 0; i < t->CF; ++i) t->src[i] = views[i];
	const size_t rgba = (size_t)target_w * target_h * 8, r32 = (size_t)target_w * target_h * 4;
	for (int k = 0; k < 7; ++k) {
		t->img[k].assign(t->CF, nullptr);
		for (int i = 0; i < t->CF; ++i) {
			size_t bytes = (k == TAA_IMG_SEGMASK) ? r32 : rgba;
			e = cudaMalloc(&t->img[k][i], bytes);
			if (e == cudaSuccess) e = cudaMemset(t->img[k][i], 0, bytes);  // the reference leaves them undefined
			if (e != cudaSuccess) return inv_cuda(t, e, "cudaMalloc(owned image)");
		}
	}
	if (t->evStart.empty()) {
		t->evStart.resize(t->CF); t->evStop.resize(t->CF); t->evValid.assign(t->CF, 0);
		for (int i = 0; i < t->CF; ++i) { cudaEventCreate(&t->evStart[i]); cudaEventCreate(&t->evStop[i]); }
	}
	// "also initialize some gui elements based on image dimensions" (taa.hpp:351-359)
	t->S.mSplitX = target_w / 2;
	taa_postprocess_default(&t->mPostProcessPushConstants, target_w, target_h);
	t->isVeryFirstFrame = true;
	t->prevSharpenFactor = -1.f;
	t->oldParamsValid = false;
	return TAA_OK;
}

TaaParameters* taa_invokee_parameters(taa_invokee* t, int32_t i) { return (t && (i == 0 || i == 1)) ? &t->mParameters[i] : nullptr; }
taa_invokee_settings* taa_invokee_settings_ptr(taa_invokee* t) { return t ? &t->S : nullptr; }
TaaPostProcessPush* taa_invokee_postprocess(taa_invokee* t) { return t ? &t->mPostProcessPushConstants : nullptr; }
const TaaUniforms* taa_invokee_uniforms(const taa_invokee* t) { return t ? &t->mTaaUniforms : nullptr; }
long long taa_invokee_launch_count(const taa_invokee* t) { return (t && t->ctx) ? taa_launch_count(t->ctx) : 0; }
long long taa_invokee_h2d_bytes(const taa_invokee* t) { return t ? t->h2dBytes : 0; }

// writeSettingsToIni / readSettingsFromIni (taa.hpp:1198-1339) on the invokee's own members; taa_ini.cu holds the format
int32_t taa_invokee_write_settings_ini(taa_invokee* t, char* out, int32_t cap) {
	if (!t) return TAA_E_INVALID_ARG;
	return taa_settings_write_ini(t->mParameters, &t->S, &t->mPostProcessPushConstants, out, cap);
}
int taa_invokee_read_settings_ini(taa_invokee* t, const char* text) {
	if (!t || !text) return TAA_E_INVALID_ARG;
	std::vector<float> tmp(2 * 4096, 0.f);
	const int r = taa_settings_read_ini(text, t->mParameters, &t->S, &t->mPostProcessPushConstants, tmp.data(), 4096);
	const int n = t->S.jitter.mDebugSampleOffsets == tmp.data() ? t->S.jitter.mDebugSampleOffsetsCount : 0;
	if (n > 0) t->debugOffsets.assign(tmp.begin(), tmp.begin() + 2 * n);
	t->S.jitter.mDebugSampleOffsets = t->debugOffsets.data();
	t->S.jitter.mDebugSampleOffsetsCount = (int32_t)(t->debugOffsets.size() / 2);
	if (r != TAA_OK) t->last_error = taa_settings_ini_last_error();
	return r;
}

// get_jittered_projection_matrix — taa.hpp:243-259
int taa_invokee_get_jittered_projection_matrix(taa_invokee* t, const float proj[16], int64_t frame, float out_proj[16], float out_jitter[2]) {
	if (!t || !proj || !out_proj || t->in_w <= 0) return TAA_E_INVALID_ARG;
	float j[2] = {0.f, 0.f};
	if (t->S.mTaaEnabled) {
		int n = taa_jitter_offset_for_frame(&t->S.jitter, t->in_w, t->in_h, frame, j);
		if (n < 0) return n;
		taa_jittered_projection(proj, j[0], j[1], out_proj);
	} else {
		memcpy(out_proj, proj, 64);
	}
	if (out_jitter) { out_jitter[0] = j[0]; out_jitter[1] = j[1]; }
	return TAA_OK;
}

// save_history_proj_matrix — taa.hpp:235-240 (stores the UN-jittered projection, main.cpp:4123)
int taa_invokee_save_history_proj_matrix(taa_invokee* t, const float proj[16], int64_t frame) {
	if (!t || !proj || frame < 0) return TAA_E_INVALID_ARG;
	memcpy(&t->mHistoryProjMatrices[(size_t)slot_of(t, frame) * 16], proj, 64);
	return TAA_OK;
}

// update() — taa.hpp:894-971 (minus handle_input)
int taa_invokee_update(taa_invokee* t, int64_t frame, const float view[16], float time_s, float cam_near, float cam_far) {
	if (!t || !view || frame < 0 || t->in_w <= 0) return TAA_E_INVALID_ARG;
	const int i = slot_of(t, frame);
	bool bypassHistUpdate = false;  // taa.hpp:906-912
	if (t->S.jitter.mJitterSlowMotion > 1) {
		long long thisJitterIndex = frame / t->S.jitter.mJitterSlowMotion;
		bypassHistUpdate = (thisJitterIndex == t->lastJitterIndex);
		t->lastJitterIndex = thisJitterIndex;
	}
	if (t->S.mResetHistoryOnChange && t->oldParamsValid) {  // taa.hpp:914-931
		for (int k = 0; k < 2; ++k) {
			TaaParameters cmp[2] = {t->mParameters[k], t->oldParams[k]};
			for (auto& c : cmp) {
				c.mDebugMode = 0; c.mDebugScale = 0.f; c.mDebugCenter = 0; c.mDebugToScreenOutput = 0;
				c.mDebugMask[0] = c.mDebugMask[1] = c.mDebugMask[2] = c.mDebugMask[3] = 0.f;
			}
			if (memcmp(&cmp[0], &cmp[1], sizeof(TaaParameters)) != 0) t->S.mResetHistory = 1;
		}
	}
	memcpy(&t->mHistoryViewMatrices[(size_t)i * 16], view, 64);  // taa.hpp:935
	float jitter[2];
	int n = taa_jitter_offset_for_frame(&t->S.jitter, t->in_w, t->in_h, frame, jitter);  // taa.hpp:937
	if (n < 0) { inv_error(t, "invalid jitter settings"); return n; }
	t->numJitterSamples = (size_t)n;
	TaaUniforms& U = t->mTaaUniforms;
	for (int k = 0; k < 2; ++k) {  // taa.hpp:939-942
		U.param[k] = t->mParameters[k];
		if (t->mParameters[k].mRayTraceHistoryCount < 0) U.param[k].mRayTraceHistoryCount = n;
	}
	U.mJitterNdc[0] = jitter[0]; U.mJitterNdc[1] = jitter[1]; U.mJitterNdc[2] = 0.f; U.mJitterNdc[3] = 0.f;
	const float tm[4] = {.125f, .25f, .5f, 1.f};
	for (int k = 0; k < 4; ++k) U.mSinTime[k] = sinf(tm[k] * time_s);  // taa.hpp:944
	U.mUpsampling = t->mUpsampling;
	U.splitScreen = t->S.mSplitScreen;
	U.splitX = t->S.mSplitX;
	U.mBypassHistoryUpdate = bypassHistUpdate;
	U.mResetHistory = t->S.mResetHistory;
	U.mCamNearPlane = cam_near;
	U.mCamFarPlane = cam_far;
	t->mFxaaPushConstants.fxaaQualityRcpFrame[0] = 1.f / (float)t->out_w;  // taa.hpp:953
	t->mFxaaPushConstants.fxaaQualityRcpFrame[1] = 1.f / (float)t->out_h;
	TaaPostProcessPush& pp = t->mPostProcessPushConstants;  // taa.hpp:955-959
	pp.splitX = t->S.mSplitScreen ? t->S.mSplitX : -1;
	memcpy(pp.debugL_mask, t->mParameters[0].mDebugMask, 16);
	memcpy(pp.debugR_mask, t->mParameters[1].mDebugMask, 16);
	pp.debugL_show = t->mParameters[0].mDebugToScreenOutput;
	pp.debugR_show = t->mParameters[1].mDebugToScreenOutput;
	if (t->S.mSharpenFactor != t->prevSharpenFactor) {  // taa.hpp:961-966
		t->prevSharpenFactor = t->S.mSharpenFactor;
		t->mSharpenerPushConstants.sharpeningFactor = t->S.mSharpenFactor;
		taa_cas_setup(&t->mCasPushConstants, t->S.mSharpenFactor, (float)t->out_w, (float)t->out_h);
	}
	t->S.mResetHistory = 0;  // taa.hpp:968
	t->oldParams[0] = t->mParameters[0];
	t->oldParams[1] = t->mParameters[1];
	t->oldParamsValid = true;
	return TAA_OK;
}

static int render_with_sources(taa_invokee* t, int64_t frame, const std::vector<taa_source_views>& src, cudaStream_t stream, const void** out_final) {
	const int i = slot_of(t, frame), last = (i + t->CF - 1) % t->CF;  // taa.hpp:980-981
	const taa_source_views& cur = src[i];
	const taa_source_views& prev = src[last];
	if (!cur.color || !cur.depth || !cur.velocity) { inv_error(t, "frame slot %d has no colour/depth/velocity view", i); return TAA_E_INVALID_ARG; }
	const int W = t->out_w, H = t->out_h;
	cudaEventRecord(t->evStart[i], stream);
	const void* final_img = nullptr;
	if (t->S.mTaaEnabled && !t->isVeryFirstFrame) {  // taa.hpp:990
		TaaUniforms& U = t->mTaaUniforms;
		int r = taa_reprojection_matrices(&t->mHistoryProjMatrices[(size_t)i * 16], &t->mHistoryViewMatrices[(size_t)i * 16],
		                                  &t->mHistoryProjMatrices[(size_t)last * 16], &t->mHistoryViewMatrices[(size_t)last * 16],
		                                  U.mInverseViewProjMatrix, U.mHistoryViewProjMatrix);  // taa.hpp:993-994
		if (r != TAA_OK) { inv_error(t, "projection*view of slot %d is singular", i); return r; }
		taa_resolve_images im;
		memset(&im, 0, sizeof im);  // bindings as at taa.hpp:1009-1025
		im.color = whole((void*)cur.color, t->in_w, t->in_h, 8);
		im.depth = whole((void*)cur.depth, t->in_w, t->in_h, 4);
		im.velocity = whole((void*)cur.velocity, t->in_w, t->in_h, 8);
		im.history_in = whole(t->img[TAA_IMG_HISTORY][last], W, H, 8);
		if (prev.depth) im.history_depth = whole((void*)prev.depth, t->in_w, t->in_h, 4);
		im.history_out = whole(t->img[TAA_IMG_HISTORY][i], W, H, 8);
		im.result = whole(t->img[TAA_IMG_RESULT][i], W, H, 8);
		const bool anyDebug = t->mParameters[0].mDebugToScreenOutput || (t->S.mSplitScreen && t->mParameters[1].mDebugToScreenOutput);
		if (anyDebug) im.debug = whole(t->img[TAA_IMG_DEBUG][i], W, H, 8);  // the reference writes it every frame (taa.comp:957); we only when shown
		const bool rt = t->mParameters[0].mRayTraceAugment || (t->S.mSplitScreen && t->mParameters[1].mRayTraceAugment);  // needRayTraceAssist, taa.hpp:1344
		if (rt) {
			im.segmask = whole(t->img[TAA_IMG_SEGMASK][i], W, H, 4);
			im.prev_segmask = whole(t->img[TAA_IMG_SEGMASK][last], W, H, 4);
			if (cur.matid) im.matid = whole((void*)cur.matid, t->in_w, t->in_h, 4);
			if (prev.matid) im.prev_matid = whole((void*)prev.matid, t->in_w, t->in_h, 4);
			if (cur.uvnrm) im.uvnrm = whole((void*)cur.uvnrm, t->in_w, t->in_h, 16);
		}
		taa_post_chain chain;
		memset(&chain, 0, sizeof chain);
		chain.sharpener = t->S.mSharpener;
		chain.sharpen = t->mSharpenerPushConstants;
		chain.cas = t->mCasPushConstants;
		chain.postprocess = t->S.mPostProcessEnabled ? 1 : 0;
		chain.pp = t->mPostProcessPushConstants;
		// FXAA on the pixels the seg-mask marks (taa.hpp:1033, 1061). The sparse ray-trace callback between the two (taa.hpp:1036-1058)
		// belongs to the ray tracer and is out of scope: pixels marked 2 keep their TAA colour.
		chain.fxaa = rt && ((t->mParameters[0].mRayTraceAugmentFlags & TAA_RTFLAG_FXA) || (t->S.mSplitScreen && (t->mParameters[1].mRayTraceAugmentFlags & TAA_RTFLAG_FXA)));
		chain.fxaa_pc = t->mFxaaPushConstants;
		// which image ends up on screen: result -> temp[0] (sharpener) -> postprocess (taa.hpp:1029-1161)
		void* fin = t->img[TAA_IMG_RESULT][i];
		if (chain.postprocess) fin = t->img[TAA_IMG_POSTPROCESS][i];
		else if (chain.sharpener) fin = t->img[chain.fxaa ? TAA_IMG_TEMP1 : TAA_IMG_TEMP0][i];
		else if (chain.fxaa) fin = t->img[TAA_IMG_TEMP0][i];
		taa_image fimg = whole(fin, W, H, 8);
		r = taa_frame(t->ctx, &im, &U, &chain, &fimg, stream);
		if (r != TAA_OK) { inv_error(t, "taa_frame: %s", taa_last_error_string(t->ctx)); return r; }
		final_img = fin;
	} else {
		// "blit" colour -> result (taa.hpp:1176): a plain copy for equal sizes, the nearest-texel scaling of vkCmdBlitImage with upsampling
		cudaError_t e;
		if (t->mUpsampling) {
			taa::PostImg io{};
			io.src = taa::Img{(const unsigned char*)cur.color, (long long)t->in_w * 8, 0, t->in_h};
			io.dst = taa::ImgW{(unsigned char*)t->img[TAA_IMG_RESULT][i], (long long)W * 8, 0, H};
			io.w = W; io.h = H;
			e = taa::launch_blit_nearest(io, t->in_w, t->in_h, stream);
			if (e == cudaSuccess) t->ctx->launches++;
		} else {
			e = cudaMemcpyAsync(t->img[TAA_IMG_RESULT][i], cur.color, (size_t)W * H * 8, cudaMemcpyDeviceToDevice, stream);
		}
		if (e != cudaSuccess) return inv_cuda(t, e, "blit colour -> result");
		final_img = t->img[TAA_IMG_RESULT][i];
	}
	cudaEventRecord(t->evStop[i], stream);
	t->evValid[i] = 1;
	t->isVeryFirstFrame = false;  // taa.hpp:1183
	if (out_final) *out_final = final_img;
	return TAA_OK;
}

// render() — taa.hpp:974-1192
int taa_invokee_render(taa_invokee* t, int64_t frame, void* stream, const void** out_final) {
	if (!t || !t->ctx || frame < 0) return TAA_E_INVALID_ARG;
	return render_with_sources(t, frame, t->src, (cudaStream_t)stream, out_final);
}

// duration() — taa.hpp:367-374. Returns the device time of the most recent render() that has completed.
float taa_invokee_duration(taa_invokee* t) {
	if (!t || !t->S.mTaaEnabled) return 0.0f;
	for (int i = 0; i < (int)t->evValid.size(); ++i) {
		if (!t->evValid[i]) continue;
		if (cudaEventQuery(t->evStop[i]) == cudaSuccess) {
			float ms = 0.f;
			if (cudaEventElapsedTime(&ms, t->evStart[i], t->evStop[i]) == cudaSuccess) t->lastDurationMs = ms;
			t->evValid[i] = 0;
		}
	}
	return t->lastDurationMs;
}

void* taa_invokee_image(taa_invokee* t, int32_t which, int32_t slot) {
	if (!t || which < 0 || which >= 7 || slot < 0 || slot >= (int)t->img[which].size()) return nullptr;
	return t->img[which][slot];
}

// ---- host-buffer frames --------------------------------------------------------------------------
static int ensure_pipe(taa_invokee* t, bool uvnrm, bool matid) {
	if (t->pipe_ready) return TAA_OK;
	cudaError_t e = cudaSetDevice(t->device);
	if (e != cudaSuccess) return inv_cuda(t, e, "cudaSetDevice");
	if ((e = cudaStreamCreateWithFlags(&t->sUp, cudaStreamNonBlocking)) != cudaSuccess) return inv_cuda(t, e, "stream");
	if ((e = cudaStreamCreateWithFlags(&t->sCompute, cudaStreamNonBlocking)) != cudaSuccess) return inv_cuda(t, e, "stream");
	if ((e = cudaStreamCreateWithFlags(&t->sDown, cudaStreamNonBlocking)) != cudaSuccess) return inv_cuda(t, e, "stream");
	t->evUploaded.resize(t->CF); t->evComputed.resize(t->CF); t->evDone.resize(t->CF);
	t->slotBusy.assign(t->CF, 0);
	for (int i = 0; i < t->CF; ++i) {
		cudaEventCreateWithFlags(&t->evUploaded[i], cudaEventDisableTiming);
		cudaEventCreateWithFlags(&t->evComputed[i], cudaEventDisableTiming);
		cudaEventCreateWithFlags(&t->evDone[i], cudaEventDisableTiming);
	}
	t->dsrc.assign(t->CF, taa_source_views{});
	t->owns_dsrc = true;   // (set first: a failure below leaves a partly allocated set that free_images() / free_pipe() release)
	t->pipe_ready = true;
	const size_t px = (size_t)t->in_w * t->in_h;
	for (int i = 0; i < t->CF; ++i) {
		void *c = nullptr, *d = nullptr, *v = nullptr, *n = nullptr, *m = nullptr;
		if ((e = cudaMalloc(&c, px * 8)) == cudaSuccess && (e = cudaMalloc(&d, px * 4)) == cudaSuccess) e = cudaMalloc(&v, px * 8);
		if (e == cudaSuccess && uvnrm) e = cudaMalloc(&n, px * 16);
		if (e == cudaSuccess && matid) e = cudaMalloc(&m, px * 4);
		t->dsrc[i] = taa_source_views{c, d, n, v, m, nullptr};
		if (e != cudaSuccess) {
			int r = inv_cuda(t, e, "cudaMalloc(device G-buffer)");
			free_pipe(t);
			for (auto& sv : t->dsrc) { cudaFree((void*)sv.color); cudaFree((void*)sv.depth); cudaFree((void*)sv.velocity); cudaFree((void*)sv.uvnrm); cudaFree((void*)sv.matid); }
			t->dsrc.clear();
			t->owns_dsrc = false;
			return r;
		}
	}
	return TAA_OK;
}

int taa_invokee_frame_host(taa_invokee* t, int64_t frame, const taa_source_views* hv, const float view[16], const float proj[16],
                           float time_s, float cam_near, float cam_far, void* out_final_host) {
	if (!t || !t->ctx || !hv || !view || !proj || !out_final_host || frame < 0) return TAA_E_INVALID_ARG;
	if (!hv->color || !hv->depth || !hv->velocity) { inv_error(t, "host views need colour, depth and velocity"); return TAA_E_INVALID_ARG; }
	int r = ensure_pipe(t, hv->uvnrm != nullptr, hv->matid != nullptr);
	if (r != TAA_OK) return r;
	const int i = slot_of(t, frame);
	cudaError_t e;
	// the slot's previous frame (frame - CF) must have left the device: the fence wait of composition.hpp:242-244
	if (t->slotBusy[i]) { if ((e = cudaEventSynchronize(t->evDone[i])) != cudaSuccess) return inv_cuda(t, e, "wait slot"); t->slotBusy[i] = 0; }
	// slot i's buffers were last read by frame-CF (current) and frame-CF+1 (as history depth / previous material)
	const int reader = (i + 1) % t->CF;
	if (frame >= t->CF - 1) cudaStreamWaitEvent(t->sUp, t->evComputed[reader], 0);
	const size_t px = (size_t)t->in_w * t->in_h;
	const taa_source_views& d = t->dsrc[i];
	if ((e = cudaMemcpyAsync((void*)d.color, hv->color, px * 8, cudaMemcpyHostToDevice, t->sUp)) != cudaSuccess) return inv_cuda(t, e, "H2D colour");
	// Depth is uploaded when a dispatch of this frame can read it: taa.comp touches uCurrentDepth for the matrix reprojection of pixels
	// without a usable velocity (taa.comp:421-430), the closest-depth velocity mode (:399-404), depth culling (:815-823, also as the NEXT
	// frame's history depth) and the segmentation mask (:640). With velocity for everything and none of those (BASELINE config 2) no kernel reads
	// it, and 4 of the 20 bytes per pixel stay off the PCIe link. After a settings change that starts to need depth the first frame
	// sees the last uploaded depth as its history depth — the reference resets the history on such a change (mResetHistoryOnChange).
	bool need_depth = false;
	for (int k = 0; k < (t->S.mSplitScreen ? 2 : 1); ++k) {
		const TaaParameters& P = t->mParameters[k];
		need_depth = need_depth || P.mDepthCulling || P.mUseVelocityVectors != 2 || P.mVelocitySampleMode == 2 || P.mRayTraceAugment;
	}
	t->h2dBytes += (long long)px * (8 + 8 + (need_depth ? 4 : 0)) + ((d.uvnrm && hv->uvnrm) ? (long long)px * 16 : 0) + ((d.matid && hv->matid) ? (long long)px * 4 : 0);
	if (need_depth && (e = cudaMemcpyAsync((void*)d.depth, hv->depth, px * 4, cudaMemcpyHostToDevice, t->sUp)) != cudaSuccess) return inv_cuda(t, e, "H2D depth");
	if ((e = cudaMemcpyAsync((void*)d.velocity, hv->velocity, px * 8, cudaMemcpyHostToDevice, t->sUp)) != cudaSuccess) return inv_cuda(t, e, "H2D velocity");
	if (d.uvnrm && hv->uvnrm && (e = cudaMemcpyAsync((void*)d.uvnrm, hv->uvnrm, px * 16, cudaMemcpyHostToDevice, t->sUp)) != cudaSuccess) return inv_cuda(t, e, "H2D uvnrm");
	if (d.matid && hv->matid && (e = cudaMemcpyAsync((void*)d.matid, hv->matid, px * 4, cudaMemcpyHostToDevice, t->sUp)) != cudaSuccess) return inv_cuda(t, e, "H2D matid");
	cudaEventRecord(t->evUploaded[i], t->sUp);
	// what wookiee::update / taa::update do per frame (main.cpp:4122-4123, taa.hpp:894)
	if ((r = taa_invokee_save_history_proj_matrix(t, proj, frame)) != TAA_OK) return r;
	if ((r = taa_invokee_update(t, frame, view, time_s, cam_near, cam_far)) != TAA_OK) return r;
	cudaStreamWaitEvent(t->sCompute, t->evUploaded[i], 0);
	const void* fin = nullptr;
	if ((r = render_with_sources(t, frame, t->dsrc, t->sCompute, &fin)) != TAA_OK) return r;
	cudaEventRecord(t->evComputed[i], t->sCompute);
	cudaStreamWaitEvent(t->sDown, t->evComputed[i], 0);
	if ((e = cudaMemcpyAsync(out_final_host, fin, (size_t)t->out_w * t->out_h * 8, cudaMemcpyDeviceToHost, t->sDown)) != cudaSuccess) return inv_cuda(t, e, "D2H final");
	cudaEventRecord(t->evDone[i], t->sDown);
	t->slotBusy[i] = 1;
	return TAA_OK;
}

int taa_invokee_wait(taa_invokee* t, int64_t frame) {
	if (!t || !t->pipe_ready || frame < 0) return TAA_E_INVALID_ARG;
	const int i = slot_of(t, frame);
	if (!t->slotBusy[i]) return TAA_OK;
	cudaError_t e = cudaEventSynchronize(t->evDone[i]);
	if (e != cudaSuccess) return inv_cuda(t, e, "taa_invokee_wait");
	t->slotBusy[i] = 0;
	return TAA_OK;
}

}  // extern "C"
