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

// ---- [sharpen | CAS] + identity post-process as a streaming pass --------------------------------------------------------------------------
// The common case of the follow-on launch (no zoom box, no splitter, no debug view: post_process.comp only copies): a warp owns 64 columns,
// a lane two adjacent texels (one 16-byte load and one 16-byte store per row), and walks down a run of rows with the rows above and below
// the current one in registers — every texel is loaded once (plus two halo rows per run and one halo texel per warp edge), the horizontal
// neighbours come over warp shuffles. The arithmetic is sharpen_at / cas_at's, value for value (the exact tests compare bit for bit).
struct Px3 { float x, y, z; };
__device__ __forceinline__ Px3 px3_of(unsigned int rg, unsigned int ba) {
	const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&rg));
	Px3 o; o.x = a.x; o.y = a.y; o.z = __low2float(*reinterpret_cast<const __half2*>(&ba));
	return o;
}
struct RowRec { Px3 a, b, ext; };  // the lane's two texels; ext: the texel beside the warp's 64 columns (lane 0: left of them, lane 31: right of them)

template <int SHARPENER>
__device__ __forceinline__ Px3 sharpened(const Px3 up, const Px3 lf, const Px3 ce, const Px3 rt, const Px3 dn, const float k) {
	Px3 o;
	if (SHARPENER == 1) {
		o.x = fminf(fmaxf(ce.x + ((((4.0f * ce.x - lf.x) - rt.x) - up.x) - dn.x) * k, 0.f), 1.f);
		o.y = fminf(fmaxf(ce.y + ((((4.0f * ce.y - lf.y) - rt.y) - up.y) - dn.y) * k, 0.f), 1.f);
		o.z = fminf(fmaxf(ce.z + ((((4.0f * ce.z - lf.z) - rt.z) - up.z) - dn.z) * k, 0.f), 1.f);
	} else {
		const float mnG = min3f(min3f(lf.y, ce.y, rt.y), up.y, dn.y);
		const float mxG = max3f(max3f(lf.y, ce.y, rt.y), up.y, dn.y);
		float ampG = clampf(fminf(mnG, 1.0f - mxG) * prx_lo_rcp(mxG), 0.f, 1.f);
		ampG = prx_lo_sqrt(ampG);
		const float wG = ampG * k;
		const float rcpWeight = prx_med_rcp(1.0f + 4.0f * wG);
		o.x = clampf((up.x * wG + lf.x * wG + rt.x * wG + dn.x * wG + ce.x) * rcpWeight, 0.f, 1.f);
		o.y = clampf((up.y * wG + lf.y * wG + rt.y * wG + dn.y * wG + ce.y) * rcpWeight, 0.f, 1.f);
		o.z = clampf((up.z * wG + lf.z * wG + rt.z * wG + dn.z * wG + ce.z) * rcpWeight, 0.f, 1.f);
	}
	return o;
}

template <int SHARPENER>
__global__ void __launch_bounds__(128) sharpen_rows_kernel(const __grid_constant__ PostImg io, const float k, const int run) {
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int w = io.w, h = io.h;
	const int xw = (blockIdx.x * 4 + warp) * 64;  // first column of the warp
	if (xw >= w) return;
	const int x = xw + 2 * lane;                  // the lane's first column (w is even: both of its texels exist or neither does)
	const bool live = x < w;
	const int y0 = blockIdx.y * run, y1 = min(y0 + run, h);
	const Px3 zero = {0.f, 0.f, 0.f};
	// one image row as the stencil sees it: out of range reads as 0 (CasFilter loads unguarded, ffx_cas.h:429-437; sharpen.comp:21 clamps
	// to `size`, so the row above the first is the first row itself, the row below the last is out of range)
	auto load_row = [&](int y) -> RowRec {
		RowRec r;
		r.a = zero; r.b = zero; r.ext = zero;
		if (SHARPENER == 1 && y < 0) y = 0;
		if (y < 0 || y >= h) return r;
		const unsigned char* row = io.src.p + (long long)(y - io.src.y0) * io.src.pitch;
		if (live) {
			const uint4 t = __ldg(reinterpret_cast<const uint4*>(row + (size_t)x * 8u));
			r.a = px3_of(t.x, t.y); r.b = px3_of(t.z, t.w);
		}
		const int xe = lane == 0 ? xw - 1 : xw + 64;
		if ((lane == 0 || lane == 31) && xe >= 0 && xe < w) {
			const uint2 t = __ldg(reinterpret_cast<const uint2*>(row + (size_t)xe * 8u));
			r.ext = px3_of(t.x, t.y);
		}
		return r;
	};
	RowRec up = load_row(y0 - 1), ce = load_row(y0);
	for (int y = y0; y < y1; ++y) {
		const RowRec dn = load_row(y + 1);
		Px3 lf, rt;
		lf.x = __shfl_up_sync(0xffffffffu, ce.b.x, 1); lf.y = __shfl_up_sync(0xffffffffu, ce.b.y, 1); lf.z = __shfl_up_sync(0xffffffffu, ce.b.z, 1);
		rt.x = __shfl_down_sync(0xffffffffu, ce.a.x, 1); rt.y = __shfl_down_sync(0xffffffffu, ce.a.y, 1); rt.z = __shfl_down_sync(0xffffffffu, ce.a.z, 1);
		if (lane == 0) lf = (SHARPENER == 1 && x == 0) ? ce.a : ce.ext;  // (sharpen.comp clamps the column to 0: its own texel)
		if (lane == 31) rt = ce.ext;
		if (x + 2 >= w) rt = zero;  // right of the last column: out of range
		if (live) {
			const Px3 fa = sharpened<SHARPENER>(up.a, lf, ce.a, ce.b, dn.a, k), fb = sharpened<SHARPENER>(up.b, ce.a, ce.b, rt, dn.b, k);
			const __half2 arg = __floats2half2_rn(fa.x, fa.y), ab = __floats2half2_rn(fa.z, 1.0f), brg = __floats2half2_rn(fb.x, fb.y), bb = __floats2half2_rn(fb.z, 1.0f);
			*reinterpret_cast<uint4*>(io.dst.p + (long long)(y - io.dst.y0) * io.dst.pitch + (size_t)x * 8u) =
			    make_uint4(*reinterpret_cast<const unsigned int*>(&arg), *reinterpret_cast<const unsigned int*>(&ab), *reinterpret_cast<const unsigned int*>(&brg),
			               *reinterpret_cast<const unsigned int*>(&bb));
		}
		up = ce; ce = dn;
	}
}

inline dim3 grid2d(int w, int h, dim3 b) { return dim3((w + b.x - 1) / b.x, (h + b.y - 1) / b.y); }

}  // namespace

cudaError_t launch_blit_nearest(const PostImg& io, int src_w, int src_h, cudaStream_t stream) {
	dim3 b(32, 8);
	blit_nearest_kernel<<<grid2d(io.w, io.h, b), b, 0, stream>>>(io, src_w, src_h);
	return cudaGetLastError();
}
// images whose rows can be moved 16 bytes (two texels) at a time take the streaming pass (same values, alpha = 1 like sharpen_at / cas_at)
static bool rows_kernel_ok(const PostImg& io) {
	return (io.w & 1) == 0 && (((unsigned long long)io.src.p | (unsigned long long)io.src.pitch | (unsigned long long)io.dst.p | (unsigned long long)io.dst.pitch) & 15ull) == 0ull;
}
constexpr int ROWS_RUN = 32;  // rows per warp of the streaming pass: two halo rows per 32
cudaError_t launch_sharpen(const PostImg& io, float factor, cudaStream_t stream) {
	if (rows_kernel_ok(io)) {
		sharpen_rows_kernel<1><<<dim3((io.w + 255) / 256, (io.h + ROWS_RUN - 1) / ROWS_RUN), 128, 0, stream>>>(io, factor, ROWS_RUN);
		return cudaGetLastError();
	}
	dim3 b(32, 8);
	sharpen_kernel<<<grid2d(io.w, io.h, b), b, 0, stream>>>(io, factor);
	return cudaGetLastError();
}
cudaError_t launch_cas(const PostImg& io, const TaaCasPush& pc, cudaStream_t stream) {
	float peak;
	memcpy(&peak, &pc.const1[0], 4);
	if (rows_kernel_ok(io)) {
		sharpen_rows_kernel<2><<<dim3((io.w + 255) / 256, (io.h + ROWS_RUN - 1) / ROWS_RUN), 128, 0, stream>>>(io, peak, ROWS_RUN);
		return cudaGetLastError();
	}
	dim3 b(32, 8);
	cas_kernel<<<grid2d(io.w, io.h, b), b, 0, stream>>>(io, peak);
	return cudaGetLastError();
}
cudaError_t launch_post_process(const PostImg& io, const TaaPostProcessPush& pc, cudaStream_t stream) {
	dim3 b(32, 8);
	post_process_kernel<0><<<grid2d(io.w, io.h, b), b, 0, stream>>>(io, pc, 0.f);
	return cudaGetLastError();
}
cudaError_t launch_sharpen_post(const PostImg& io, int sharpener, float sharpeningFactor, const TaaCasPush& cas, const TaaPostProcessPush& pc, cudaStream_t stream) {
	// post_process.comp only copies (no zoom, no splitter, no debug view) and the rows can be moved 16 bytes at a time: the streaming pass
	const bool identity = !pc.zoom && pc.splitX < 0 && !pc.debugL_show;
	if (identity && rows_kernel_ok(io) && (sharpener == 1 || sharpener == 2)) {
		const int run = ROWS_RUN;
		const dim3 grid((io.w + 255) / 256, (io.h + run - 1) / run);
		if (sharpener == 1) sharpen_rows_kernel<1><<<grid, 128, 0, stream>>>(io, sharpeningFactor, run);
		else {
			float peak;
			memcpy(&peak, &cas.const1[0], 4);
			sharpen_rows_kernel<2><<<grid, 128, 0, stream>>>(io, peak, run);
		}
		return cudaGetLastError();
	}
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
