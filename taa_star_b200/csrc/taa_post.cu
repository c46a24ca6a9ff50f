// taa_post.cu — the follow-on full-screen passes of taa<CF>::render() (source/taa.hpp:1111-1159) as
// stand-alone kernels: sharpen.comp, sharpen_cas.comp (FidelityFX CAS, sharpen-only), post_process.comp.
// Out-of-range image reads return 0 (the reference leaves them undefined: sharpen.comp:21 clamps to
// `size` instead of `size-1`; CasFilter loads at -1 / w / h unguarded, ffx_cas.h:429-437).
#include "taa_device.cuh"
#include "taa_kernels.h"

namespace taa {

namespace {

__device__ __forceinline__ f3 rgb_at(const Img& im, int w, int h, int x, int y) { return xyz(fetch_rgba16f(im, w, h, x, y, nullptr)); }

// sharpen.comp:23-38 at one pixel
__device__ __forceinline__ float4 sharpen_at(const PostImg& io, int x, int y, float factor) {
	f3 L = rgb_at(io.src, io.w, io.h, iclamp(x - 1, 0, io.w), iclamp(y, 0, io.h));
	f3 R = rgb_at(io.src, io.w, io.h, iclamp(x + 1, 0, io.w), iclamp(y, 0, io.h));
	f3 T = rgb_at(io.src, io.w, io.h, iclamp(x, 0, io.w), iclamp(y - 1, 0, io.h));
	f3 B = rgb_at(io.src, io.w, io.h, iclamp(x, 0, io.w), iclamp(y + 1, 0, io.h));
	f3 C = rgb_at(io.src, io.w, io.h, x, y);
	f3 val = C + ((((4.0f * C - L) - R) - T) - B) * factor;
	val = min3(max3(val, mk3(0.f, 0.f, 0.f)), mk3(1.f, 1.f, 1.f));
	return mk4(val, 1.f);
}
__global__ void __launch_bounds__(256) sharpen_kernel(const __grid_constant__ PostImg io, float factor) {
	const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
	if (x >= io.w || y >= io.h) return;
	st_rgba16f(io.dst, x, y, sharpen_at(io, x, y, factor));
}

// ffx_a.h:1455-1457
__device__ __forceinline__ float prx_lo_sqrt(float a) { return __uint_as_float((__float_as_uint(a) >> 1) + 0x1fbc4639u); }
__device__ __forceinline__ float prx_lo_rcp(float a) { return __uint_as_float(0x7ef07ebbu - __float_as_uint(a)); }
__device__ __forceinline__ float prx_med_rcp(float a) { float b = __uint_as_float(0x7ef19fffu - __float_as_uint(a)); return b * (-b * a + 2.0f); }
__device__ __forceinline__ float min3f(float x, float y, float z) { return fminf(x, fminf(y, z)); }
__device__ __forceinline__ float max3f(float x, float y, float z) { return fmaxf(x, fmaxf(y, z)); }

// sharpen_cas.comp:30-53 + CasFilter(noScaling), ffx_cas.h:408-537, at one pixel. Only the green weight survives (ffx_cas.h:514-522).
__device__ __forceinline__ float4 cas_at(const PostImg& io, int x, int y, float peak) {
	f3 b = rgb_at(io.src, io.w, io.h, x, y - 1);
	f3 d = rgb_at(io.src, io.w, io.h, x - 1, y);
	f3 e = rgb_at(io.src, io.w, io.h, x, y);
	f3 f = rgb_at(io.src, io.w, io.h, x + 1, y);
	f3 h = rgb_at(io.src, io.w, io.h, x, y + 1);
	float mnG = min3f(min3f(d.y, e.y, f.y), b.y, h.y);
	float mxG = max3f(max3f(d.y, e.y, f.y), b.y, h.y);
	float ampG = clampf(fminf(mnG, 1.0f - mxG) * prx_lo_rcp(mxG), 0.f, 1.f);
	ampG = prx_lo_sqrt(ampG);
	float wG = ampG * peak;
	float rcpWeight = prx_med_rcp(1.0f + 4.0f * wG);
	float pr = clampf((b.x * wG + d.x * wG + f.x * wG + h.x * wG + e.x) * rcpWeight, 0.f, 1.f);
	float pg = clampf((b.y * wG + d.y * wG + f.y * wG + h.y * wG + e.y) * rcpWeight, 0.f, 1.f);
	float pb = clampf((b.z * wG + d.z * wG + f.z * wG + h.z * wG + e.z) * rcpWeight, 0.f, 1.f);
	return make_float4(pr, pg, pb, 1.0f);  // alpha is undefined in the reference (sharpen_cas.comp:38)
}
__global__ void __launch_bounds__(256) cas_kernel(const __grid_constant__ PostImg io, float peak) {
	const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
	if (x >= io.w || y >= io.h) return;
	st_rgba16f(io.dst, x, y, cas_at(io, x, y, peak));
}

// The sharpener's value at (x, y) as the next pass would read it back from the rgba16f image it was stored to.
template <int SHARPENER>
__device__ __forceinline__ float4 stage_at(const PostImg& io, int w, int h, int x, int y, float k) {
	if (SHARPENER == 0) return fetch_rgba16f(io.src, w, h, x, y, nullptr);
	if (x < 0 || y < 0 || x >= w || y >= h) return make_float4(0.f, 0.f, 0.f, 0.f);
	return unpack_rgba16f(pack_rgba16f(SHARPENER == 1 ? sharpen_at(io, x, y, k) : cas_at(io, x, y, k)));
}

// post_process.comp:29-88. SHARPENER != 0: the sharpening pass (sharpen.comp | sharpen_cas.comp) is evaluated on the fly at the pixel
// post-process would fetch, instead of being written to and read back from an intermediate image (taa.hpp:1111-1159 does that).
template <int SHARPENER>
__global__ void __launch_bounds__(256) post_process_kernel(const __grid_constant__ PostImg io, const __grid_constant__ TaaPostProcessPush pc, const float k) {
	const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
	if (x >= io.w || y >= io.h) return;
	int fx = x, fy = y;
	if (x == pc.splitX) { st_rgba16f(io.dst, x, y, make_float4(0.f, 0.f, 0.f, 0.f)); return; }
	const int* S = pc.zoomSrcLTWH;
	const int* D = pc.zoomDstLTWH;
	if (pc.zoom && pc.showZoomBox) {
		if ((((x == S[0] - 1) || (x == S[0] + S[2])) && (y >= S[1] - 1) && (y <= S[1] + S[3])) ||
		    (((y == S[1] - 1) || (y == S[1] + S[3])) && (x >= S[0] - 1) && (x <= S[0] + S[2]))) {
			st_rgba16f(io.dst, x, y, make_float4(1.f, 0.f, 0.f, 0.f));
			return;
		}
	}
	if (pc.zoom && x >= D[0] && y >= D[1] && x < D[0] + D[2] && y < D[1] + D[3]) {
		if ((((x == D[0]) || (x == D[0] + D[2] - 1)) && (y >= D[1]) && (y <= D[1] + D[3] - 1)) ||
		    (((y == D[1]) || (y == D[1] + D[3] - 1)) && (x >= D[0]) && (x <= D[0] + D[2] - 1))) {
			st_rgba16f(io.dst, x, y, make_float4(1.f, 1.f, 1.f, 0.f));
			return;
		}
		float zu = ((float)(x - D[0]) + 0.5f) / (float)D[2], zv = ((float)(y - D[1]) + 0.5f) / (float)D[3];
		fx = (int)((float)S[0] + zu * (float)S[2]);
		fy = (int)((float)S[1] + zv * (float)S[3]);
	}
	const bool leftside = (pc.splitX < 0) || (x < pc.splitX);
	const bool showdebug = leftside ? (pc.debugL_show != 0) : (pc.debugR_show != 0);
	const float maskA = leftside ? pc.debugL_mask[3] : pc.debugR_mask[3];
	float4 val;
	if (showdebug) {
		float4 dbg = fetch_rgba16f(io.debug, io.w, io.h, fx, fy, nullptr);
		val = make_float4(dbg.x, dbg.y, dbg.z, 1.f);
		if (maskA > 0.f) { val.x += dbg.w; val.z += dbg.w; }
	} else {
		val = stage_at<SHARPENER>(io, io.w, io.h, fx, fy, k);
	}
	st_rgba16f(io.dst, x, y, val);
}

// vkCmdBlitImage with VK_FILTER_NEAREST over the whole images (avk::blit_image, gears_vk/auto_vk/src/avk.cpp:7803-7829, as taa.hpp:1176 calls it when
// anti-aliasing is off or on the very first frame): destination texel (x, y) takes source texel floor((x + 0.5) * sw / dw), floor((y + 0.5) * sh / dh)
__global__ void __launch_bounds__(256) blit_nearest_kernel(const __grid_constant__ PostImg io, const int sw, const int sh) {
	const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
	if (x >= io.w || y >= io.h) return;
	const int sx = min((int)floorf(((float)x + 0.5f) * ((float)sw / (float)io.w)), sw - 1);
	const int sy = min((int)floorf(((float)y + 0.5f) * ((float)sh / (float)io.h)), sh - 1);
	st_rgba16f(io.dst, x, y, ld_rgba16f(io.src, sx, sy, nullptr));
}

inline dim3 grid2d(int w, int h, dim3 b) { return dim3((w + b.x - 1) / b.x, (h + b.y - 1) / b.y); }

}  // namespace

cudaError_t launch_blit_nearest(const PostImg& io, int src_w, int src_h, cudaStream_t stream) {
	dim3 b(32, 8);
	blit_nearest_kernel<<<grid2d(io.w, io.h, b), b, 0, stream>>>(io, src_w, src_h);
	return cudaGetLastError();
}
cudaError_t launch_sharpen(const PostImg& io, float factor, cudaStream_t stream) {
	dim3 b(32, 8);
	sharpen_kernel<<<grid2d(io.w, io.h, b), b, 0, stream>>>(io, factor);
	return cudaGetLastError();
}
cudaError_t launch_cas(const PostImg& io, const TaaCasPush& pc, cudaStream_t stream) {
	dim3 b(32, 8);
	float peak;
	memcpy(&peak, &pc.const1[0], 4);
	cas_kernel<<<grid2d(io.w, io.h, b), b, 0, stream>>>(io, peak);
	return cudaGetLastError();
}
cudaError_t launch_post_process(const PostImg& io, const TaaPostProcessPush& pc, cudaStream_t stream) {
	dim3 b(32, 8);
	post_process_kernel<0><<<grid2d(io.w, io.h, b), b, 0, stream>>>(io, pc, 0.f);
	return cudaGetLastError();
}
cudaError_t launch_sharpen_post(const PostImg& io, int sharpener, float sharpeningFactor, const TaaCasPush& cas, const TaaPostProcessPush& pc, cudaStream_t stream) {
	dim3 b(32, 8);
	if (sharpener == 1) post_process_kernel<1><<<grid2d(io.w, io.h, b), b, 0, stream>>>(io, pc, sharpeningFactor);
	else {
		float peak;
		memcpy(&peak, &cas.const1[0], 4);
		post_process_kernel<2><<<grid2d(io.w, io.h, b), b, 0, stream>>>(io, pc, peak);
	}
	return cudaGetLastError();
}

}  // namespace taa
