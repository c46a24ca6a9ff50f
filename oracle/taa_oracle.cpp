/*
 * taa_oracle.cpp — CPU restatement of TAA-STAR's temporal resolve and its follow-on passes.
 *
 * THIS IS TEST INFRASTRUCTURE. Only tests/, __graft_entry__.smoke() and the cpu_baseline /
 * --impl reference legs of bench.py may load it. The product (taa_star_b200/) never does.
 *
 * What it restates (all paths relative to /root/reference):
 *   shaders/taa.comp                 — whole file, function by function (line cites inline)
 *   shaders/sharpen.comp             — main():23-38
 *   shaders/sharpen_cas.comp         — main():30-53, with shaders/ffx_cas.h:375-394 (CasSetup),
 *                                      :408-537 (CasFilter, noScaling branch) and shaders/ffx_a.h:1455-1457
 *   shaders/post_process.comp        — main():29-88
 *   shaders/antialias_fxaa_prepare.comp:15-26, shaders/antialias_fxaa.comp:30-64 with
 *                                      shaders/Fxaa3_11_mod.h:433-440 (preset 12), :884-1243 (FxaaPixelShader, PC quality)
 *   source/helper_functions.hpp:9-26 — halton / halton_2_3
 *   source/taa.hpp:150-233           — get_jitter_offset_for_frame
 *
 * PARITY PINNING. The reference ships no tests or golden vectors, its device code is
 * GLSL and neither glslang nor a Vulkan ICD exists in the build container. The pins are:
 *   (1) oracle/_ref: the reference's own shader sources, mechanically rewritten to C++ at build time
 *       and compiled against the GLM the reference vendors (oracle/ref_build.py + oracle/glsl_shim.h, oracle/Makefile).
 *       tests/test_oracle_vs_ref.py compares this file against it bit for bit: taa.comp, sharpen.comp, sharpen_cas.comp + ffx_cas.h's
 *       CasFilter, post_process.comp, antialias_fxaa_prepare.comp, antialias_fxaa.comp + Fxaa3_11_mod.h.
 *   (2) known answers derived from reference code that compiles as-is: CasSetup (ffx_cas.h under A_CPU)
 *       and halton_2_3<8> (tests/golden/known_answers.json, made by tests/golden/make_known_answers.py).
 *
 * ARITHMETIC MODEL (what "the reference's result" means where GLSL leaves it to the implementation):
 *   - every GLSL operator/builtin is one IEEE-754 binary32 operation, round-to-nearest-even, evaluated
 *     left to right as written; no contraction (compile with -ffp-contract=off), no reassociation.
 *   - mix(x,y,a) = x*(1-a) + y*a (GLSL spec); clamp(x,lo,hi) = min(max(x,lo),hi); fract(x) = x - floor(x);
 *     length(v) = sqrt(dot(v,v)); dot is a left-to-right sum of products; mat4*vec4 accumulates
 *     column 0 first: ((M0*v.x + M1*v.y) + M2*v.z) + M3*v.w.
 *   - min/max: IEEE-754-2019 minimum/maximumNumber (-0 < +0, a NaN operand is dropped) = CUDA fminf/fmaxf.
 *   - float -> int conversion truncates, saturates, NaN -> 0.
 *   - the sampler (taa.hpp:274: linear, clamp-to-edge, normalised coordinates, LOD 0) follows the Vulkan
 *     texel-filtering equations in fp32: u = s*W - 0.5, i0 = floor(u), a = u - i0, indices clamped to
 *     [0, W-1], and the 2x2 footprint is reduced as lerp(lerp(t00,t10,a), lerp(t01,t11,a), b) with
 *     lerp(p,q,w) = p + w*(q - p)  (the form Mesa's llvmpipe/lavapipe — the CPU Vulkan driver BASELINE.json
 *     names — uses for float formats). fp16 texels are widened to fp32 exactly.
 *   - imageStore to rgba16f rounds fp32 -> fp16 to nearest even.
 *   - out-of-range texelFetch / imageLoad return 0.
 *   - CAS leaves the output alpha undefined in the reference (sharpen_cas.comp:38); we write 1.0.
 *   - sin/cos are libm's; they only feed the optional noise and the segmentation-mask normals.
 */
#include <cmath>
#include <cstdint>
#include <cstring>
#include <cstdlib>
#include <algorithm>
#include "../include/taa_b200.h"

#ifdef _OPENMP
#include <omp.h>
#endif

// sin and cos as this repository pins them. GLSL leaves their precision to the implementation (and no Vulkan driver exists on either box), so the
// oracle, the reference-shader shim (oracle/glsl_shim.h) and the CUDA kernels (taa_device.cuh) all evaluate THIS function text: Cody-Waite
// reduction by pi/2 in three steps, the single-precision minimax polynomials of the Cephes library on [-pi/4, pi/4], one IEEE binary32 operation
// per written operation (the three files are compiled without contraction). Arguments too large to reduce (>= 1e9) read as 0; non-finite ones give NaN.
static inline void taa_sincos(float x, float* s, float* c) {
	if (!(std::fabs(x) < 1.0e9f)) { *s = x - x; *c = (x - x) + 1.0f; return; }
	const float k = std::floor(x * 0.636619772f + 0.5f);
	float r = x - k * 1.5703125f;
	r = r - k * 4.837512969970703125e-4f;
	r = r - k * 7.54978995489188216e-8f;
	const float z = r * r;
	const float ps = ((-1.9515295891e-4f * z + 8.3321608736e-3f) * z - 1.6666654611e-1f) * z * r + r;
	const float pc = ((2.443315711809948e-5f * z - 1.388731625493765e-3f) * z + 4.166664568298827e-2f) * z * z - 0.5f * z + 1.0f;
	const int q = (int)k & 3;
	*s = q == 0 ? ps : q == 1 ? pc : q == 2 ? -ps : -pc;
	*c = q == 0 ? pc : q == 1 ? -ps : q == 2 ? -pc : ps;
}
static inline float taa_sin(float x) { float s, c; taa_sincos(x, &s, &c); return s; }
static inline float taa_cos(float x) { float s, c; taa_sincos(x, &s, &c); return c; }


namespace {

// ------------------------------------------------------------------------------------------------
// scalar helpers
// ------------------------------------------------------------------------------------------------
inline float gmin(float a, float b) {
	if (a != a) return b;
	if (b != b) return a;
	if (a == b) return std::signbit(a) ? a : b;
	return a < b ? a : b;
}
inline float gmax(float a, float b) {
	if (a != a) return b;
	if (b != b) return a;
	if (a == b) return std::signbit(a) ? b : a;
	return a > b ? a : b;
}
inline float gclamp(float x, float lo, float hi) { return gmin(gmax(x, lo), hi); }
inline float gfract(float x) { return x - floorf(x); }
inline float gmix(float x, float y, float a) { return x * (1.0f - a) + y * a; }
inline int f2i(float f) {
	if (f != f) return 0;
	if (f >= 2147483648.0f) return 2147483647;
	if (f <= -2147483648.0f) return (-2147483647 - 1);
	return (int)f;
}
inline int iclamp(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
inline uint32_t f2u_bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
inline float u2f_bits(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

// fp16 <-> fp32, IEEE, round-to-nearest-even
inline float half_to_float(uint16_t h) {
	uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
	uint32_t exp = (h >> 10) & 0x1fu;
	uint32_t man = h & 0x3ffu;
	uint32_t out;
	if (exp == 0) {
		if (man == 0) out = sign;
		else {
			// subnormal: normalise
			int e = -1;
			do { e++; man <<= 1; } while ((man & 0x400u) == 0);
			man &= 0x3ffu;
			out = sign | ((uint32_t)(127 - 15 - e) << 23) | (man << 13);
		}
	} else if (exp == 31) {
		out = sign | 0x7f800000u | (man << 13);
	} else {
		out = sign | ((exp + (127 - 15)) << 23) | (man << 13);
	}
	return u2f_bits(out);
}
inline uint16_t float_to_half(float f) {
	uint32_t x = f2u_bits(f);
	uint32_t sign = (x >> 16) & 0x8000u;
	uint32_t absx = x & 0x7fffffffu;
	if (absx >= 0x7f800000u) { // inf / nan
		if (absx > 0x7f800000u) return (uint16_t)(sign | 0x7e00u | ((absx >> 13) & 0x3ffu)); // quiet NaN, keep payload top bits
		return (uint16_t)(sign | 0x7c00u);
	}
	if (absx >= 0x477ff000u) { // >= 65520 rounds to inf
		return (uint16_t)(sign | 0x7c00u);
	}
	if (absx < 0x33000001u) { // < 2^-25 (and exactly 2^-25 ties to even = 0)
		return (uint16_t)sign;
	}
	int32_t e = (int32_t)(absx >> 23) - 127;
	uint32_t m = (absx & 0x7fffffu) | 0x800000u; // 24-bit significand
	int shift;
	uint32_t base;
	if (e < -14) { // subnormal half
		shift = 13 + (-14 - e);
		base = 0;
	} else {
		shift = 13;
		base = (uint32_t)(e + 15) << 10;
		m &= 0x7fffffu;
	}
	uint32_t q = m >> shift;
	uint32_t rem = m & ((1u << shift) - 1u);
	uint32_t halfway = 1u << (shift - 1);
	if (rem > halfway || (rem == halfway && (q & 1u))) q++;
	return (uint16_t)(sign | (base + q)); // carry into exponent is correct by construction
}

// ------------------------------------------------------------------------------------------------
// tiny GLSL-like vectors; every operator is one fp32 operation per component
// ------------------------------------------------------------------------------------------------
struct vec2 { float x, y; };
struct vec3 { float x, y, z; };
struct vec4 { float x, y, z, w; };
struct ivec2 { int x, y; };

inline vec2 operator+(vec2 a, vec2 b) { return {a.x + b.x, a.y + b.y}; }
inline vec2 operator-(vec2 a, vec2 b) { return {a.x - b.x, a.y - b.y}; }
inline vec2 operator*(vec2 a, vec2 b) { return {a.x * b.x, a.y * b.y}; }
inline vec2 operator/(vec2 a, vec2 b) { return {a.x / b.x, a.y / b.y}; }
inline vec2 operator+(vec2 a, float b) { return {a.x + b, a.y + b}; }
inline vec2 operator-(vec2 a, float b) { return {a.x - b, a.y - b}; }
inline vec2 operator*(vec2 a, float b) { return {a.x * b, a.y * b}; }
inline vec2 operator*(float a, vec2 b) { return {a * b.x, a * b.y}; }
inline vec2 operator/(float a, vec2 b) { return {a / b.x, a / b.y}; }
inline vec2 toVec2(ivec2 a) { return {(float)a.x, (float)a.y}; }
inline vec2 vfloor(vec2 a) { return {floorf(a.x), floorf(a.y)}; }
inline ivec2 toIvec2(vec2 a) { return {f2i(a.x), f2i(a.y)}; }
inline float dot(vec2 a, vec2 b) { return a.x * b.x + a.y * b.y; }

inline vec3 operator+(vec3 a, vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline vec3 operator-(vec3 a, vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline vec3 operator*(vec3 a, vec3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline vec3 operator/(vec3 a, vec3 b) { return {a.x / b.x, a.y / b.y, a.z / b.z}; }
inline vec3 operator*(vec3 a, float b) { return {a.x * b, a.y * b, a.z * b}; }
inline vec3 operator*(float a, vec3 b) { return {a * b.x, a * b.y, a * b.z}; }
inline vec3 operator/(vec3 a, float b) { return {a.x / b, a.y / b, a.z / b}; }
inline vec3 operator+(vec3 a, float b) { return {a.x + b, a.y + b, a.z + b}; }
inline vec3 vmin(vec3 a, vec3 b) { return {gmin(a.x, b.x), gmin(a.y, b.y), gmin(a.z, b.z)}; }
inline vec3 vmax(vec3 a, vec3 b) { return {gmax(a.x, b.x), gmax(a.y, b.y), gmax(a.z, b.z)}; }
inline vec3 vabs(vec3 a) { return {fabsf(a.x), fabsf(a.y), fabsf(a.z)}; }
inline vec3 vsqrt(vec3 a) { return {sqrtf(a.x), sqrtf(a.y), sqrtf(a.z)}; }
inline vec3 vclamp(vec3 v, vec3 lo, vec3 hi) { return vmin(vmax(v, lo), hi); }
inline vec3 vmix(vec3 x, vec3 y, float a) { return {gmix(x.x, y.x, a), gmix(x.y, y.y, a), gmix(x.z, y.z, a)}; }
inline float dot(vec3 a, vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float length(vec3 a) { return sqrtf(dot(a, a)); }
inline float length(vec2 a) { return sqrtf(dot(a, a)); }
inline bool anyLess(vec3 a, vec3 b) { return a.x < b.x || a.y < b.y || a.z < b.z; }
inline bool anyGreater(vec3 a, vec3 b) { return a.x > b.x || a.y > b.y || a.z > b.z; }

inline vec4 operator+(vec4 a, vec4 b) { return {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }
inline vec4 operator-(vec4 a, vec4 b) { return {a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w}; }
inline vec4 operator*(vec4 a, vec4 b) { return {a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w}; }
inline vec4 operator*(vec4 a, float b) { return {a.x * b, a.y * b, a.z * b, a.w * b}; }
inline vec4 operator/(vec4 a, float b) { return {a.x / b, a.y / b, a.z / b, a.w / b}; }
inline vec4 operator+(vec4 a, float b) { return {a.x + b, a.y + b, a.z + b, a.w + b}; }
inline vec4 operator-(vec4 a, float b) { return {a.x - b, a.y - b, a.z - b, a.w - b}; }
inline vec4 vabs(vec4 a) { return {fabsf(a.x), fabsf(a.y), fabsf(a.z), fabsf(a.w)}; }
inline vec3 rgb(vec4 a) { return {a.x, a.y, a.z}; }
inline vec4 mkvec4(vec3 a, float w) { return {a.x, a.y, a.z, w}; }
inline vec2 xy(vec4 a) { return {a.x, a.y}; }

// column-major mat4 * vec4: ((M0*v.x + M1*v.y) + M2*v.z) + M3*v.w
inline vec4 mat4_mul(const float* m, vec4 v) {
	vec4 r;
	r.x = ((m[0] * v.x + m[4] * v.y) + m[8] * v.z) + m[12] * v.w;
	r.y = ((m[1] * v.x + m[5] * v.y) + m[9] * v.z) + m[13] * v.w;
	r.z = ((m[2] * v.x + m[6] * v.y) + m[10] * v.z) + m[14] * v.w;
	r.w = ((m[3] * v.x + m[7] * v.y) + m[11] * v.z) + m[15] * v.w;
	return r;
}

// ------------------------------------------------------------------------------------------------
// images and the sampler
// ------------------------------------------------------------------------------------------------
struct Tex {
	const uint8_t* data = nullptr;
	int64_t pitch = 0;
	int w = 0, h = 0;
	bool valid() const { return data != nullptr; }
};

inline Tex mkTex(const taa_image& im, int w, int h) {
	Tex t;
	t.data = (const uint8_t*)im.data;
	t.pitch = im.pitch_bytes;
	t.w = w;
	t.h = h;
	return t;
}

// texelFetch / imageLoad on rgba16f; out of range -> 0
inline vec4 fetch_rgba16f(const Tex& t, int x, int y) {
	if (x < 0 || y < 0 || x >= t.w || y >= t.h) return {0, 0, 0, 0};
	const uint16_t* p = (const uint16_t*)(t.data + (int64_t)y * t.pitch) + (int64_t)x * 4;
	return {half_to_float(p[0]), half_to_float(p[1]), half_to_float(p[2]), half_to_float(p[3])};
}
inline float fetch_r32f(const Tex& t, int x, int y) {
	if (x < 0 || y < 0 || x >= t.w || y >= t.h) return 0.0f;
	return ((const float*)(t.data + (int64_t)y * t.pitch))[x];
}
inline uint32_t fetch_r32ui(const Tex& t, int x, int y) {
	if (x < 0 || y < 0 || x >= t.w || y >= t.h) return 0u;
	return ((const uint32_t*)(t.data + (int64_t)y * t.pitch))[x];
}
inline vec4 fetch_rgba32f(const Tex& t, int x, int y) {
	if (x < 0 || y < 0 || x >= t.w || y >= t.h) return {0, 0, 0, 0};
	const float* p = (const float*)(t.data + (int64_t)y * t.pitch) + (int64_t)x * 4;
	return {p[0], p[1], p[2], p[3]};
}

struct LinearCoord { int i0, i1; float a; };
// Vulkan unnormalised-coordinate + linear footprint selection for one axis, clamp-to-edge
inline LinearCoord linear_coord(float s, int size) {
	float u = s * (float)size - 0.5f;
	float fl = floorf(u);
	LinearCoord c;
	c.a = u - fl;
	int i0 = f2i(fl);
	c.i0 = iclamp(i0, 0, size - 1);
	c.i1 = iclamp(i0 == 2147483647 ? i0 : i0 + 1, 0, size - 1);
	return c;
}
inline float lerp1(float p, float q, float w) { return p + w * (q - p); }
inline vec4 lerp4(vec4 p, vec4 q, float w) { return {lerp1(p.x, q.x, w), lerp1(p.y, q.y, w), lerp1(p.z, q.z, w), lerp1(p.w, q.w, w)}; }

// texture(sampler2D(tex, uSampler), uv) on rgba16f
inline vec4 sample_rgba16f(const Tex& t, vec2 uv) {
	LinearCoord cx = linear_coord(uv.x, t.w);
	LinearCoord cy = linear_coord(uv.y, t.h);
	vec4 t00 = fetch_rgba16f(t, cx.i0, cy.i0);
	vec4 t10 = fetch_rgba16f(t, cx.i1, cy.i0);
	vec4 t01 = fetch_rgba16f(t, cx.i0, cy.i1);
	vec4 t11 = fetch_rgba16f(t, cx.i1, cy.i1);
	return lerp4(lerp4(t00, t10, cx.a), lerp4(t01, t11, cx.a), cy.a);
}
// texture(sampler2D(uCurrentDepth, uSampler), uv).r on D32 (taa.comp:378-386)
inline float sample_r32f(const Tex& t, vec2 uv) {
	LinearCoord cx = linear_coord(uv.x, t.w);
	LinearCoord cy = linear_coord(uv.y, t.h);
	float t00 = fetch_r32f(t, cx.i0, cy.i0);
	float t10 = fetch_r32f(t, cx.i1, cy.i0);
	float t01 = fetch_r32f(t, cx.i0, cy.i1);
	float t11 = fetch_r32f(t, cx.i1, cy.i1);
	return lerp1(lerp1(t00, t10, cx.a), lerp1(t01, t11, cx.a), cy.a);
}

inline void store_rgba16f(const taa_image& im, int x, int y, vec4 v) {
	if (!im.data) return;
	uint16_t* p = (uint16_t*)((uint8_t*)im.data + (int64_t)y * im.pitch_bytes) + (int64_t)x * 4;
	p[0] = float_to_half(v.x);
	p[1] = float_to_half(v.y);
	p[2] = float_to_half(v.z);
	p[3] = float_to_half(v.w);
}
inline void store_r32ui(const taa_image& im, int x, int y, uint32_t v) {
	if (!im.data) return;
	((uint32_t*)((uint8_t*)im.data + (int64_t)y * im.pitch_bytes))[x] = v;
}

// ------------------------------------------------------------------------------------------------
// taa.comp
// ------------------------------------------------------------------------------------------------
struct Shader {
	// bindings (taa.comp:24-38)
	Tex uCurrentFrame, uCurrentDepth, uCurrentVelocity, uCurrentUvNrm, uCurrentMaterial, uPreviousMaterial;
	Tex uHistoryFrame, uHistoryDepth, uPreviousSegMask;
	taa_image uResultScreen, uResultHistory, uSegMask, uDebug, uMask;
	const TaaUniforms* ubo;
	TaaParameters params; // `Parameters params;` taa.comp:100
	ivec2 textureSize_loRes, textureSize_hiRes; // taa.comp:122-123
	vec4 gDebugValue; // taa.comp:126

	// taa.comp:131-132
	static vec2 tc_to_uv(ivec2 tc, ivec2 texSize) { return (toVec2(tc) + 0.5f) / toVec2(texSize); }
	static ivec2 uv_to_tc(vec2 uv, ivec2 texSize) { return toIvec2(uv * toVec2(texSize)); }

	// taa.comp:47  #define JITTER_UV (ubo.mJitterNdc.xy * 0.5 * params.mUnjitterFactor)
	vec2 JITTER_UV() const { return (vec2{ubo->mJitterNdc[0], ubo->mJitterNdc[1]} * 0.5f) * params.mUnjitterFactor; }

	// taa.comp:158-164
	static vec3 rgb_to_ycocg(vec3 c) {
		return {
			.25f * c.x + .5f * c.y + .25f * c.z,
			.5f * c.x - .5f * c.z,
			-.25f * c.x + .5f * c.y - .25f * c.z};
	}
	// taa.comp:167-174
	static vec3 ycocg_to_rgb(vec3 c) {
		float tmp = c.x - c.z;
		return {tmp + c.y, c.x + c.z, tmp - c.y};
	}
	// taa.comp:177-179
	vec3 maybe_rgb_to_ycocg(vec3 c) const { return params.mUseYCoCg ? rgb_to_ycocg(c) : c; }
	vec3 maybe_ycocg_to_rgb(vec3 c) const { return params.mUseYCoCg ? ycocg_to_rgb(c) : c; }
	float luminance(vec3 c) const { return params.mUseYCoCg ? c.x : rgb_to_ycocg(c).x; }
	// taa.comp:182-197
	vec3 tonemap_rgb(vec3 hdr) const {
		if (params.mToneMapLumaKaris) {
			float luma = gmax(gmax(hdr.x, hdr.y), hdr.z);
			return hdr / (1.0f + luma);
		}
		return hdr;
	}
	vec3 un_tonemap_rgb(vec3 ldr) const {
		if (params.mToneMapLumaKaris) {
			float luma = gmax(gmax(ldr.x, ldr.y), ldr.z);
			return ldr / (1.0f - luma);
		}
		return ldr;
	}

	// one tap of taa.comp:204-212 / :219
	vec3 colourTap(vec2 offset, ivec2 iuv, int dx, int dy, vec2 invsize) const {
		vec2 p = toVec2(ivec2{iuv.x + dx, iuv.y + dy}) + 0.5f; // vec2(iuv + ivec2(dx,dy) + 0.5)
		return maybe_rgb_to_ycocg(tonemap_rgb(rgb(sample_rgba16f(uCurrentFrame, offset + p * invsize))));
	}
	// taa.comp:199-213
	void getNeighbourhood(ivec2 iuv, vec3& cC, vec3& c1, vec3& c2, vec3& c3, vec3& c4, vec3& c5, vec3& c6, vec3& c7, vec3& c8) const {
		vec2 offset = params.mUnjitterNeighbourhood ? JITTER_UV() : vec2{0, 0};
		vec2 invsize = 1.0f / toVec2(textureSize_loRes);
		cC = colourTap(offset, iuv, 0, 0, invsize);
		c1 = colourTap(offset, iuv, -1, -1, invsize);
		c2 = colourTap(offset, iuv, 0, -1, invsize);
		c3 = colourTap(offset, iuv, 1, -1, invsize);
		c4 = colourTap(offset, iuv, -1, 0, invsize);
		c5 = colourTap(offset, iuv, 1, 0, invsize);
		c6 = colourTap(offset, iuv, -1, 1, invsize);
		c7 = colourTap(offset, iuv, 0, 1, invsize);
		c8 = colourTap(offset, iuv, 1, 1, invsize);
	}
	// taa.comp:216-220
	vec3 getCurrentColor(ivec2 iuv) const {
		vec2 offset = params.mUnjitterCurrentSample ? JITTER_UV() : vec2{0, 0};
		vec2 invsize = 1.0f / toVec2(textureSize_loRes);
		return colourTap(offset, iuv, 0, 0, invsize);
	}
	// taa.comp:222-257
	vec3 getCurrentUpsampledColor(ivec2 currentTc, vec2 /*currentUv*/, float& beta) const {
		vec2 lo = toVec2(textureSize_loRes), hi = toVec2(textureSize_hiRes);
		vec2 scale = hi / lo;
		vec2 texelJitter = (JITTER_UV() * lo) * -1.0f;
		vec2 inTcSample;
		vec2 foundTc = {-1, -1};
		const float almostOne = 0.999999f;
		const vec2 probes[4] = {{0, 0}, {almostOne, 0}, {0, almostOne}, {almostOne, almostOne}};
		for (int i = 0; i < 4; ++i) {
			inTcSample = (vfloor((lo * (toVec2(currentTc) + probes[i])) / hi) + 0.5f) + texelJitter;
			ivec2 back = toIvec2(inTcSample * scale);
			if (back.x == currentTc.x && back.y == currentTc.y) foundTc = inTcSample;
		}
		if (foundTc.x >= 0.0f) {
			beta = 1.0f;
			vec2 texUv = (vfloor(foundTc) + 0.5f) / lo;
			return maybe_rgb_to_ycocg(tonemap_rgb(rgb(sample_rgba16f(uCurrentFrame, texUv))));
		} else {
			beta = 0.0f;
			return {0, 0, 0};
		}
	}
	// taa.comp:259-319
	void getColorAndAabb(ivec2 iuv, vec3& centerCol, vec3& minCol, vec3& maxCol, vec3& cliptowardsCol) const {
		const float N = 9.0f;
		vec3 c1, c2, c3, c4, c5, c6, c7, c8;
		getNeighbourhood(iuv, centerCol, c1, c2, c3, c4, c5, c6, c7, c8);
		if (params.mVarianceClipping) {
			vec3 m1 = centerCol + c1 + c2 + c3 + c4 + c5 + c6 + c7 + c8;
			vec3 m2 = centerCol * centerCol + c1 * c1 + c2 * c2 + c3 * c3 + c4 * c4 + c5 * c5 + c6 * c6 + c7 * c7 + c8 * c8;
			vec3 mean = m1 / N;
			vec3 sigma = vsqrt(vmax(vec3{0, 0, 0}, m2 / N - mean * mean));
			minCol = mean - params.mVarClipGamma * sigma;
			maxCol = mean + params.mVarClipGamma * sigma;
			cliptowardsCol = mean;
		} else if (params.mShapedNeighbourhood) {
			vec3 minCol_3x3 = vmin(vmin(vmin(vmin(vmin(vmin(vmin(vmin(centerCol, c1), c2), c3), c4), c5), c6), c7), c8);
			vec3 maxCol_3x3 = vmax(vmax(vmax(vmax(vmax(vmax(vmax(vmax(centerCol, c1), c2), c3), c4), c5), c6), c7), c8);
			vec3 minCol_5tap = vmin(vmin(vmin(vmin(centerCol, c2), c4), c5), c7);
			vec3 maxCol_5tap = vmax(vmax(vmax(vmax(centerCol, c2), c4), c5), c7);
			minCol = (minCol_3x3 + minCol_5tap) * 0.5f;
			maxCol = (maxCol_3x3 + maxCol_5tap) * 0.5f;
			cliptowardsCol = centerCol;
		} else {
			minCol = vmin(vmin(vmin(vmin(vmin(vmin(vmin(vmin(centerCol, c1), c2), c3), c4), c5), c6), c7), c8);
			maxCol = vmax(vmax(vmax(vmax(vmax(vmax(vmax(vmax(centerCol, c1), c2), c3), c4), c5), c6), c7), c8);
			cliptowardsCol = centerCol;
		}
		if (params.mUseYCoCg && params.mShrinkChromaAxis) {
			const vec3 scaleYCoCg = {1.0f, 0.5f, 0.5f};
			vec3 halfSize = (0.5f * scaleYCoCg) * (maxCol - minCol);
			vec3 center = (minCol + maxCol) * 0.5f;
			minCol = center - halfSize;
			maxCol = center + halfSize;
			if (anyLess(cliptowardsCol, minCol) || anyGreater(cliptowardsCol, maxCol)) cliptowardsCol = center;
		}
	}
	// taa.comp:323-345
	static vec4 clipAabb(vec3 aabbMin, vec3 aabbMax, vec4 p, vec4 q) {
		const float eps = 1e-7f;
		vec3 pClip = 0.5f * (aabbMax + aabbMin);
		vec3 eClip = 0.5f * (aabbMax - aabbMin) + eps;
		vec4 vClip = q - mkvec4(pClip, p.w);
		vec3 vUnit = rgb(vClip) / eClip;
		vec3 aUnit = vabs(vUnit);
		float maUnit = gmax(aUnit.x, gmax(aUnit.y, aUnit.z));
		if (maUnit > 1.0f) return mkvec4(pClip, p.w) + vClip / maUnit;
		return q;
	}
	// taa.comp:348-369
	static vec4 clipAabbSlow(vec3 aabbMin, vec3 aabbMax, vec4 p, vec4 q) {
		vec4 r = q - p;
		vec3 rmax = aabbMax - rgb(p);
		vec3 rmin = aabbMin - rgb(p);
		const float eps = 1e-7f;
		if (r.x > rmax.x + eps) r = r * (rmax.x / r.x);
		if (r.y > rmax.y + eps) r = r * (rmax.y / r.y);
		if (r.z > rmax.z + eps) r = r * (rmax.z / r.z);
		if (r.x < rmin.x - eps) r = r * (rmin.x / r.x);
		if (r.y < rmin.y - eps) r = r * (rmin.y / r.y);
		if (r.z < rmin.z - eps) r = r * (rmin.z / r.z);
		return p + r;
	}
	// taa.comp:371-389
	vec3 findClosestUvAndZ_3x3(vec2 uv) const {
		vec2 toUv = vec2{1, 1} / toVec2(ivec2{uCurrentDepth.w, uCurrentDepth.h});
		vec2 offset, closestOffset;
		float d, dClosest;
		const vec2 taps[9] = {{-1, -1}, {0, -1}, {1, -1}, {-1, 0}, {0, 0}, {1, 0}, {-1, 1}, {0, 1}, {1, 1}};
		offset = toUv * taps[0];
		d = sample_r32f(uCurrentDepth, uv + offset);
		closestOffset = offset;
		dClosest = d;
		for (int i = 1; i < 9; ++i) {
			offset = toUv * taps[i];
			d = sample_r32f(uCurrentDepth, uv + offset);
			if (d < dClosest) { closestOffset = offset; dClosest = d; }
		}
		vec2 r = uv + closestOffset;
		return {r.x, r.y, dClosest};
	}
	// taa.comp:391-438
	void getHistoryPosition(vec2 currentUv, float currentDepth, vec2& historyUv, float& historyDepth, float& outputPixelSpeed) const {
		vec4 velocitySample = sample_rgba16f(uCurrentVelocity, currentUv);
		bool canUseVelocity = true;
		if (params.mUseVelocityVectors == 0 || (params.mUseVelocityVectors == 1 && velocitySample.w < 0.5f)) canUseVelocity = false;
		if (canUseVelocity) {
			if (params.mVelocitySampleMode == 1) {
				vec2 toUv = vec2{1, 1} / toVec2(ivec2{uCurrentVelocity.w, uCurrentVelocity.h});
				vec2 maxVel = xy(velocitySample);
				vec2 sam = maxVel;
				const vec2 taps[8] = {{1, -1}, {-1, 0}, {-1, -1}, {0, -1}, {-1, 1}, {0, 1}, {1, 0}, {1, 1}};
				for (int i = 0; i < 8; ++i) {
					sam = xy(sample_rgba16f(uCurrentVelocity, currentUv + toUv * taps[i]));
					if (dot(sam, sam) > dot(maxVel, maxVel)) maxVel = sam;
				}
				velocitySample.x = sam.x; // taa.comp:412 (sic: last tap, not maxVel)
				velocitySample.y = sam.y;
			} else if (params.mVelocitySampleMode == 2) {
				vec3 closestUvAndZ = findClosestUvAndZ_3x3(currentUv);
				velocitySample = sample_rgba16f(uCurrentVelocity, vec2{closestUvAndZ.x, closestUvAndZ.y});
			}
			historyUv = currentUv - xy(velocitySample);
			historyDepth = currentDepth - velocitySample.z;
		} else {
			vec2 c2 = currentUv * 2.0f - 1.0f;
			vec4 clipSpace = {c2.x, c2.y, currentDepth, 1.0f};
			vec4 worldSpace = mat4_mul(ubo->mInverseViewProjMatrix, clipSpace);
			vec4 historyClipSpace = mat4_mul(ubo->mHistoryViewProjMatrix, worldSpace);
			historyUv = (vec2{historyClipSpace.x, historyClipSpace.y} / vec2{historyClipSpace.w, historyClipSpace.w}) * 0.5f + 0.5f;
			historyDepth = historyClipSpace.z / historyClipSpace.w;
		}
		vec2 velUv = currentUv - historyUv;
		outputPixelSpeed = sqrtf(dot(velUv, velUv));
	}
	// taa.comp:441-514
	vec4 sample_history_bicubic_catmullrom(vec2 uv) const {
		vec2 texSize = toVec2(ivec2{uHistoryFrame.w, uHistoryFrame.h});
		vec2 invTexSize = 1.0f / texSize;
		vec2 iTc = uv * texSize;
		vec2 tc = vfloor(iTc - 0.5f) + 0.5f;
		vec2 f = iTc - tc;
		vec2 f2 = f * f;
		vec2 f3 = f2 * f;
		vec2 w0 = -0.5f * f3 + f2 - 0.5f * f;
		vec2 w1 = 1.5f * f3 - 2.5f * f2 + 1.0f;
		vec2 w2 = -1.5f * f3 + 2.0f * f2 + 0.5f * f;
		vec2 w3 = 0.5f * f3 - 0.5f * f2;
		vec2 wC = w1 + w2;
		vec2 tc0 = (tc - 1.0f) * invTexSize;
		vec2 tcC = (tc + w2 / wC) * invTexSize;
		vec2 tc3 = (tc + 2.0f) * invTexSize;
		const Tex& H = uHistoryFrame;
		return sample_rgba16f(H, vec2{tc0.x, tc0.y}) * w0.x * w0.y
		     + sample_rgba16f(H, vec2{tcC.x, tc0.y}) * wC.x * w0.y
		     + sample_rgba16f(H, vec2{tc3.x, tc0.y}) * w3.x * w0.y
		     + sample_rgba16f(H, vec2{tc0.x, tcC.y}) * w0.x * wC.y
		     + sample_rgba16f(H, vec2{tcC.x, tcC.y}) * wC.x * wC.y
		     + sample_rgba16f(H, vec2{tc3.x, tcC.y}) * w3.x * wC.y
		     + sample_rgba16f(H, vec2{tc0.x, tc3.y}) * w0.x * w3.y
		     + sample_rgba16f(H, vec2{tcC.x, tc3.y}) * wC.x * w3.y
		     + sample_rgba16f(H, vec2{tc3.x, tc3.y}) * w3.x * w3.y;
	}
	// taa.comp:517-543
	vec4 sample_history_bicubic_b_spline(vec2 uv) const {
		vec2 texSize = toVec2(ivec2{uHistoryFrame.w, uHistoryFrame.h});
		vec2 invTexSize = 1.0f / texSize;
		vec2 iTc = uv * texSize;
		vec2 tc = vfloor(iTc - 0.5f) + 0.5f;
		vec2 f = iTc - tc;
		vec2 f2 = f * f;
		vec2 f3 = f2 * f;
		vec2 w0 = f2 - 0.5f * (f3 + f);
		vec2 w1 = 1.5f * f3 - 2.5f * f2 + 1.0f;
		vec2 w3 = 0.5f * (f3 - f2);
		vec2 w2 = (vec2{1.0f, 1.0f} - w0) - w1 - w3;
		vec2 s0 = w0 + w1;
		vec2 s1 = w2 + w3;
		vec2 f0 = w1 / (w0 + w1);
		vec2 f1 = w3 / (w2 + w3);
		vec2 t0 = ((tc - 1.0f) + f0) * invTexSize;
		vec2 t1 = ((tc + 1.0f) + f1) * invTexSize;
		const Tex& H = uHistoryFrame;
		return (sample_rgba16f(H, vec2{t0.x, t0.y}) * s0.x + sample_rgba16f(H, vec2{t1.x, t0.y}) * s1.x) * s0.y
		     + (sample_rgba16f(H, vec2{t0.x, t1.y}) * s0.x + sample_rgba16f(H, vec2{t1.x, t1.y}) * s1.x) * s1.y;
	}
	// taa.comp:545-549
	vec4 sample_history_rgba(vec2 uv) const {
		if (params.mInterpolationMode == 0) return sample_rgba16f(uHistoryFrame, uv);
		else if (params.mInterpolationMode == 1) return sample_history_bicubic_b_spline(uv);
		else return sample_history_bicubic_catmullrom(uv);
	}
	// taa.comp:551-556
	vec4 noise(vec2 uv) const {
		vec2 seed = uv + ubo->mSinTime[0] + 0.6959174f;
		float s = taa_sin(dot(seed, vec2{12.9898f, 78.233f}));
		vec4 nRand = {gfract(s * 43758.5453f), gfract(s * 28001.8384f), gfract(s * 50849.4141f), gfract(s * 12996.89f)};
		vec4 sRand = nRand * 2.0f - 1.0f;
		return sRand * params.mNoiseFactor;
	}
	// taa.comp:559-561
	static float linearize_depth(float d, float zNear, float zFar) { return zNear * zFar / (zFar + d * (zNear - zFar)); }
	// taa.comp:566-568
	float sample_linear_depth(ivec2 iuv) const { return linearize_depth(fetch_r32f(uCurrentDepth, iuv.x, iuv.y), ubo->mCamNearPlane, ubo->mCamFarPlane); }
	// taa.comp:570-572
	float sample_luminance(ivec2 iuv) const { return rgb_to_ycocg(rgb(fetch_rgba16f(uCurrentFrame, iuv.x, iuv.y))).x; }
	// taa.comp:574-580
	vec3 sample_normal(ivec2 iuv) const {
		vec4 uvNormal = fetch_rgba32f(uCurrentUvNrm, iuv.x, iuv.y);
		return {taa_cos(uvNormal.z) * taa_cos(uvNormal.w), taa_sin(uvNormal.z) * taa_cos(uvNormal.w), taa_sin(uvNormal.w)};
	}
	// taa.comp:582-587
	static vec2 sobel(float c00, float c01, float c02, float c10, float c12, float c20, float c21, float c22) {
		vec2 g;
		g.x = c00 - c20 + 2.0f * c01 - 2.0f * c21 + c02 - c22;
		g.y = c00 - c02 + 2.0f * c10 - 2.0f * c12 + c20 - c22;
		return g;
	}
	// #define CLAMP_TO_TEX(v) clamp((v), ivec2(0), textureSize(uCurrentFrame,0)-1)   taa.comp:42
	ivec2 CLAMP_TO_TEX(ivec2 v) const { return {iclamp(v.x, 0, uCurrentFrame.w - 1), iclamp(v.y, 0, uCurrentFrame.h - 1)}; }

	// taa.comp:589-702
	uint32_t calc_segmentation_value(ivec2 iuv, vec2 /*uv*/, vec2 historyUv, float /*historyDepth*/) const {
		const uint32_t flags = params.mRayTraceAugmentFlags;
		if ((flags & TAA_RTFLAG_FXD) != 0) {
			const int b = 100;
			if (iuv.x < b || iuv.y < b || iuv.x >= textureSize_loRes.x - b || iuv.y >= textureSize_loRes.y - b) return 1;
		}
		if ((flags & TAA_RTFLAG_ALL) != 0) return 2;
		bool useHistoryCount = (flags & TAA_RTFLAG_CNT) != 0;
		uint32_t newCountValue = useHistoryCount ? ((uint32_t)params.mRayTraceHistoryCount << 16) : 0u;
		if ((flags & TAA_RTFLAG_OUT) != 0) {
			if (historyUv.x < 0.0f || historyUv.y < 0.0f || historyUv.x >= 1.0f || historyUv.y >= 1.0f) return 1;
		}
		uint32_t matId = fetch_r32ui(uCurrentMaterial, iuv.x, iuv.y);
		if ((flags & TAA_RTFLAG_DIS) != 0) {
			ivec2 tcPrev = uv_to_tc(historyUv, textureSize_loRes);
			if (tcPrev.x >= 0 && tcPrev.y >= 0 && tcPrev.x < textureSize_loRes.x && tcPrev.y < textureSize_loRes.y) {
				uint32_t prevMatId = fetch_r32ui(uPreviousMaterial, tcPrev.x, tcPrev.y);
				if (prevMatId != matId && (prevMatId & 0x80000000u) != 0) return 2u | newCountValue;
			}
		}
		float nrmValue = 0.0f, dptValue = 0.0f, matValue = 0.0f, lumValue = 0.0f;
		auto off = [&](int dx, int dy) { return CLAMP_TO_TEX(ivec2{iuv.x + dx, iuv.y + dy}); };
		if ((flags & TAA_RTFLAG_NRM) != 0) {
			vec3 nC = sample_normal(iuv);
			vec3 nL = sample_normal(off(-1, 0));
			vec3 nR = sample_normal(off(1, 0));
			vec3 nT = sample_normal(off(0, -1));
			vec3 nB = sample_normal(off(0, 1));
			float mind = gmax(0.0f, gmin(gmin(gmin(dot(nC, nL), dot(nC, nR)), dot(nC, nT)), dot(nC, nB)));
			nrmValue = 1.0f - mind;
		}
		if ((flags & TAA_RTFLAG_DPT) != 0) {
			dptValue = length(sobel(sample_linear_depth(off(-1, -1)), sample_linear_depth(off(0, -1)), sample_linear_depth(off(1, -1)),
			                        sample_linear_depth(off(-1, 0)), sample_linear_depth(off(1, 0)),
			                        sample_linear_depth(off(-1, 1)), sample_linear_depth(off(0, 1)), sample_linear_depth(off(1, 1))));
		}
		if ((flags & TAA_RTFLAG_MID) != 0) {
			auto m = [&](int dx, int dy) { ivec2 p = off(dx, dy); return fetch_r32ui(uCurrentMaterial, p.x, p.y); };
			if (matId != m(-1, 0) || matId != m(1, 0) || matId != m(0, -1) || matId != m(0, 1)) matValue = 1.0f;
		}
		if ((flags & TAA_RTFLAG_LUM) != 0) {
			lumValue = length(sobel(sample_luminance(off(-1, -1)), sample_luminance(off(0, -1)), sample_luminance(off(1, -1)),
			                        sample_luminance(off(-1, 0)), sample_luminance(off(1, 0)),
			                        sample_luminance(off(-1, 1)), sample_luminance(off(0, 1)), sample_luminance(off(1, 1))));
		}
		float total = nrmValue * params.mRayTraceAugment_WNrm
		            + dptValue * params.mRayTraceAugment_WDpt
		            + matValue * params.mRayTraceAugment_WMId
		            + lumValue * params.mRayTraceAugment_WLum;
		if (total >= params.mRayTraceAugment_Thresh) return 2u | newCountValue;
		if (useHistoryCount) {
			uint32_t oldCnt = (fetch_r32ui(uPreviousSegMask, iuv.x, iuv.y) & 0xffff0000u) >> 16;
			if (oldCnt > 0) return 2u | ((oldCnt - 1) << 16);
		}
		return 0;
	}

	// taa.comp:708-960, one invocation
	void main(ivec2 iuv) {
		gDebugValue = {0, 0, 0, 0};
		vec2 uv = tc_to_uv(iuv, textureSize_hiRes);
		if (iuv.x >= textureSize_hiRes.x || iuv.y >= textureSize_hiRes.y) return;

		int paramsIdx = (ubo->splitScreen && iuv.x > ubo->splitX) ? 1 : 0;
		params = ubo->param[paramsIdx];
		bool generateSegmentationMask = params.mRayTraceAugment != 0;
		ivec2 iuv_lores = uv_to_tc(uv, textureSize_loRes);

		if (params.mPassThrough) {
			store_rgba16f(uResultScreen, iuv.x, iuv.y, mkvec4(rgb(fetch_rgba16f(uCurrentFrame, iuv_lores.x, iuv_lores.y)), 1));
			store_rgba16f(uResultHistory, iuv.x, iuv.y, mkvec4(rgb(fetch_rgba16f(uHistoryFrame, iuv.x, iuv.y)), 1));
			store_rgba16f(uDebug, iuv.x, iuv.y, {0, 0, 0, 0});
			store_r32ui(uMask, iuv.x, iuv.y, 0);
			return;
		}
		if (ubo->mBypassHistoryUpdate) {
			store_rgba16f(uResultScreen, iuv.x, iuv.y, mkvec4(rgb(fetch_rgba16f(uHistoryFrame, iuv.x, iuv.y)), 1));
			store_rgba16f(uResultHistory, iuv.x, iuv.y, mkvec4(rgb(fetch_rgba16f(uHistoryFrame, iuv.x, iuv.y)), 1));
			store_rgba16f(uDebug, iuv.x, iuv.y, {0, 0, 0, 0});
			store_r32ui(uMask, iuv.x, iuv.y, 0);
			return;
		}

		bool rejected = false;
		bool rectified = false;
		vec3 rectified_diff;
		vec3 currentColor, colMin, colMax, colClipTowards;
		float beta;
		getColorAndAabb(iuv_lores, currentColor, colMin, colMax, colClipTowards);
		if (ubo->mUpsampling) {
			currentColor = getCurrentUpsampledColor(iuv, uv, beta);
		} else {
			currentColor = getCurrentColor(iuv_lores);
			beta = 1.0f;
		}
		float depth = fetch_r32f(uCurrentDepth, iuv_lores.x, iuv_lores.y);

		vec2 historyUv;
		float expectedHistoryDepth, pixelSpeed;
		getHistoryPosition(uv, depth, historyUv, expectedHistoryDepth, pixelSpeed);

		vec4 historyRaw = sample_history_rgba(historyUv);
		vec3 historyColor = maybe_rgb_to_ycocg(rgb(historyRaw));

		float alpha = params.mAlpha;

		uint32_t segMaskValue;
		if (generateSegmentationMask) {
			segMaskValue = calc_segmentation_value(iuv, uv, historyUv, expectedHistoryDepth);
			if ((segMaskValue & 3u) != 0) {
				alpha = params.mRejectionAlpha;
				rejected = true;
			}
		} else {
			segMaskValue = 0;
		}

		// ---- history rejection ---- taa.comp:787-823
		if (params.mRejectOutside) {
			if (historyUv.x < 0.0f || historyUv.y < 0.0f || historyUv.x >= 1.0f || historyUv.y >= 1.0f) {
				alpha = params.mRejectionAlpha;
				rejected = true;
			}
		}
		float writeDynamicMask = 0;
		if (params.mDynamicAntiGhosting) {
			vec2 toUv = vec2{1, 1} / toVec2(ivec2{uCurrentVelocity.w, uCurrentVelocity.h});
			const float eps = 1e-5f;
			auto mov = [&](vec2 tap) {
				vec4 v = vabs(sample_rgba16f(uCurrentVelocity, tap));
				return (v.x > eps || v.y > eps) && (v.w >= 0.5f);
			};
			bool movL = mov(uv + toUv * vec2{-1, 0});
			bool movR = mov(uv + toUv * vec2{1, 0});
			bool movT = mov(uv + toUv * vec2{0, -1});
			bool movB = mov(uv + toUv * vec2{0, 1});
			bool movC = mov(uv);
			bool movement = movL || movR || movT || movB || movC;
			if (!movement && historyRaw.w > 0.0f) rejected = true;
			writeDynamicMask = movC ? 1.0f : 0.0f;
		}
		if (params.mDepthCulling) {
			ivec2 tcd = uv_to_tc(historyUv, textureSize_loRes);
			float historyDepth = fetch_r32f(uHistoryDepth, tcd.x, tcd.y);
			float depthEpsilon = 0.1f * (1.0f - historyDepth);
			if (fabsf(historyDepth - expectedHistoryDepth) > depthEpsilon) rejected = true;
		}

		// ---- history rectification ---- taa.comp:826-845
		vec3 origHistorColor = historyColor;
		switch (params.mColorClampingOrClipping) {
			case 1: historyColor = vclamp(historyColor, colMin, colMax); break;
			case 2: historyColor = rgb(clipAabb(colMin, colMax, vec4{0, 0, 0, 1}, mkvec4(historyColor, 1.0f))); break;
			case 3: historyColor = rgb(clipAabbSlow(colMin, colMax, mkvec4(colClipTowards, 1.0f), mkvec4(historyColor, 1.0f))); break;
			default: break;
		}
		rectified_diff = historyColor - origHistorColor;
		vec3 ad = vabs(rectified_diff);
		rectified = ad.x > 0.001f || ad.y > 0.001f || ad.z > 0.001f;

		// ---- blending ---- taa.comp:848-900
		if (rejected) {
			alpha = params.mRejectionAlpha;
			beta = 1.0f;
		} else {
			if (params.mVelBasedAlpha) {
				alpha = gmax(alpha, gmix(alpha, params.mVelBasedAlphaMax, gclamp(pixelSpeed * params.mVelBasedAlphaFactor, 0.0f, 1.0f)));
			}
			if (params.mLumaWeightingLottes) {
				float lumaCurrent = luminance(currentColor);
				float lumaHistory = luminance(historyColor);
				float diff = fabsf(lumaCurrent - lumaHistory) / gmax(gmax(lumaCurrent, lumaHistory), 0.2f);
				float w = 1.0f - diff;
				float ww = w * w;
				alpha = gmix(params.mMaxAlpha, params.mMinAlpha, ww);
			}
			if (params.mReduceBlendNearClamp) {
				float colMin_lum = luminance(colMin);
				float colMax_lum = luminance(colMax);
				float history_lum = luminance(origHistorColor);
				float distToClamp = 2.0f * fabsf(gmin(history_lum - colMin_lum, colMax_lum - history_lum)) / (colMax_lum - colMin_lum);
				if (colMax_lum - colMin_lum < 0.001f) distToClamp = 1.0f;
				alpha *= gclamp(4.0f * distToClamp, 0.0f, 1.0f);
			}
		}
		if (ubo->mResetHistory) { alpha = 1.0f; beta = 1.0f; }

		vec3 antiAliased = maybe_ycocg_to_rgb(vmix(historyColor, currentColor, alpha * beta));
		if (params.mAddNoise) {
			vec4 n = noise(uv);
			antiAliased = antiAliased + rgb(n);
		}
		vec4 output_to_history = mkvec4(antiAliased, writeDynamicMask);
		vec4 output_to_screen = mkvec4(un_tonemap_rgb(antiAliased), 1.0f);

		// ---- debugging ---- taa.comp:912-942
		if (params.mDebugMode == 0) {
			vec3 tmp = colMax - colMin;
			gDebugValue = mkvec4(tmp, 0);
		} else if (params.mDebugMode == 1) {
			vec3 tmp = colMax - colMin;
			float v = tmp.x * tmp.y * tmp.z;
			gDebugValue = {v, v, v, 0};
		} else if (params.mDebugMode == 2) {
			gDebugValue = {rejected ? 1.0f : 0.0f, length(rectified_diff), 0, 0};
		} else if (params.mDebugMode == 3) {
			gDebugValue = {alpha, alpha, alpha, 0};
		} else if (params.mDebugMode == 4) {
			gDebugValue = fetch_rgba16f(uCurrentVelocity, iuv.x, iuv.y);
		} else if (params.mDebugMode == 5) {
			gDebugValue = {pixelSpeed, 0, 0, 0};
		} else if (params.mDebugMode == 6) {
			gDebugValue = output_to_screen;
		} else if (params.mDebugMode == 7) {
			gDebugValue = output_to_history;
		} else if (params.mDebugMode == 8) {
			switch (segMaskValue & 3u) {
				case 0: gDebugValue = {0, 0, 1, 0}; break;
				case 1: gDebugValue = {1, 0, 0, 0}; break;
				case 2: gDebugValue = {1, 1, 0, 0}; break;
				default: break;
			}
		}
		// gDebugValue *= params.mDebugScale * params.mDebugMask;   (scalar * vec4 first, then vec4 * vec4)
		vec4 sm = vec4{params.mDebugMask[0], params.mDebugMask[1], params.mDebugMask[2], params.mDebugMask[3]} * params.mDebugScale;
		gDebugValue = gDebugValue * sm;
		if (params.mDebugCenter) gDebugValue = gDebugValue * 0.5f + 0.5f;

		// ---- stores ---- taa.comp:955-959
		store_rgba16f(uResultHistory, iuv.x, iuv.y, output_to_history);
		store_rgba16f(uResultScreen, iuv.x, iuv.y, output_to_screen);
		store_rgba16f(uDebug, iuv.x, iuv.y, gDebugValue);
		if (generateSegmentationMask) store_r32ui(uSegMask, iuv.x, iuv.y, segMaskValue);
		store_r32ui(uMask, iuv.x, iuv.y,
		            (rejected ? 1u : 0u) | (rectified ? 2u : 0u) | (((uint32_t)params.mColorClampingOrClipping & 3u) << 2));
	}
};

// ------------------------------------------------------------------------------------------------
// FidelityFX CAS pieces (ffx_a.h:1455-1457, GLSL definitions :654-732)
// ------------------------------------------------------------------------------------------------
inline float APrxLoSqrtF1(float a) { return u2f_bits((f2u_bits(a) >> 1) + 0x1fbc4639u); }
inline float APrxLoRcpF1(float a) { return u2f_bits(0x7ef07ebbu - f2u_bits(a)); }
inline float APrxMedRcpF1(float a) { float b = u2f_bits(0x7ef19fffu - f2u_bits(a)); return b * (-b * a + 2.0f); }
// AU1_AH1_AF1 (ffx_a.h:470-544, A_CPU): table-driven fp32 -> fp16 that TRUNCATES the mantissa, flushes values
// below the half denormal range to 0 and maps everything >= 2^16 (incl. inf/NaN) to +-65504 (0x7bff).
inline uint32_t amd_f32_to_f16(float f) {
	uint32_t u = f2u_bits(f);
	uint32_t sign = (u >> 16) & 0x8000u;
	int e = (int)((u >> 23) & 0xffu);
	uint32_t m = u & 0x7fffffu;
	if (e < 103) return sign;                                       // base 0, shift 24
	if (e < 113) return sign | ((1u << (e - 103)) + (m >> (126 - e))); // half denormals
	if (e < 143) return sign | (((uint32_t)(e - 112) << 10) + (m >> 13));
	return sign | 0x7bffu;                                          // shift 24
}
inline float AMin3F1(float x, float y, float z) { return gmin(x, gmin(y, z)); }
inline float AMax3F1(float x, float y, float z) { return gmax(x, gmax(y, z)); }
inline float ASatF1(float x) { return gclamp(x, 0.0f, 1.0f); }

} // namespace

// ================================================================================================
// C interface (loaded with ctypes by tests/, bench.py cpu legs and __graft_entry__.smoke())
// ================================================================================================
extern "C" {

int taa_oracle_version(void) { return 1; }

int taa_oracle_max_threads(void) {
#ifdef _OPENMP
	return omp_get_max_threads();
#else
	return 1;
#endif
}

// taa.comp over output rows [y_begin, y_end). Images are HOST pointers; y0 fields are ignored (whole frames).
int taa_oracle_resolve(const taa_resolve_images* im, const TaaUniforms* ubo, int in_w, int in_h, int out_w, int out_h,
                       int y_begin, int y_end, int nthreads) {
	if (!im || !ubo || in_w <= 0 || in_h <= 0 || out_w <= 0 || out_h <= 0) return TAA_E_INVALID_ARG;
	if (!im->color.data || !im->depth.data || !im->velocity.data || !im->history_in.data) return TAA_E_INVALID_ARG;
	Shader proto;
	proto.uCurrentFrame = mkTex(im->color, in_w, in_h);
	proto.uCurrentDepth = mkTex(im->depth, in_w, in_h);
	proto.uCurrentVelocity = mkTex(im->velocity, in_w, in_h);
	proto.uCurrentUvNrm = mkTex(im->uvnrm, in_w, in_h);
	proto.uCurrentMaterial = mkTex(im->matid, in_w, in_h);
	proto.uPreviousMaterial = mkTex(im->prev_matid, in_w, in_h);
	proto.uHistoryFrame = mkTex(im->history_in, out_w, out_h);
	proto.uHistoryDepth = mkTex(im->history_depth, in_w, in_h);
	proto.uPreviousSegMask = mkTex(im->prev_segmask, out_w, out_h);
	// a NULL optional image behaves like an all-out-of-range one
	Tex* optional[] = {&proto.uCurrentUvNrm, &proto.uCurrentMaterial, &proto.uPreviousMaterial, &proto.uHistoryDepth, &proto.uPreviousSegMask};
	for (Tex* t : optional) if (!t->data) { t->w = 0; t->h = 0; }
	proto.uResultScreen = im->result;
	proto.uResultHistory = im->history_out;
	proto.uSegMask = im->segmask;
	proto.uDebug = im->debug;
	proto.uMask = im->mask;
	proto.ubo = ubo;
	proto.textureSize_hiRes = {out_w, out_h};
	proto.textureSize_loRes = {in_w, in_h};
	if (y_begin < 0) y_begin = 0;
	if (y_end > out_h) y_end = out_h;
#ifdef _OPENMP
	if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 4) num_threads(nthreads)
#endif
	for (int y = y_begin; y < y_end; ++y) {
		Shader s = proto;
		for (int x = 0; x < out_w; ++x) s.main(ivec2{x, y});
	}
	(void)nthreads;
	return TAA_OK;
}

// sharpen.comp:23-38.  CLAMP_TO_TEX clamps to `size`, not `size-1` (sharpen.comp:21) -> OOB fetch returns 0.
int taa_oracle_sharpen(const taa_image* src, const taa_image* dst, int w, int h, float sharpeningFactor, int nthreads) {
	if (!src || !dst || !src->data || !dst->data) return TAA_E_INVALID_ARG;
	Tex in = mkTex(*src, w, h);
#ifdef _OPENMP
	if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel for schedule(static) num_threads(nthreads)
#endif
	for (int y = 0; y < h; ++y) {
		for (int x = 0; x < w; ++x) {
			auto C2T = [&](int px, int py) { return ivec2{iclamp(px, 0, w), iclamp(py, 0, h)}; };
			ivec2 pl = C2T(x - 1, y), pr = C2T(x + 1, y), pt = C2T(x, y - 1), pb = C2T(x, y + 1);
			vec3 L = rgb(fetch_rgba16f(in, pl.x, pl.y));
			vec3 R = rgb(fetch_rgba16f(in, pr.x, pr.y));
			vec3 T = rgb(fetch_rgba16f(in, pt.x, pt.y));
			vec3 B = rgb(fetch_rgba16f(in, pb.x, pb.y));
			vec3 C = rgb(fetch_rgba16f(in, x, y));
			vec3 val = C + ((((4.0f * C - L) - R) - T) - B) * sharpeningFactor;
			val = vclamp(val, vec3{0, 0, 0}, vec3{1, 1, 1});
			store_rgba16f(*dst, x, y, mkvec4(val, 1));
		}
	}
	(void)nthreads;
	return TAA_OK;
}

// CasSetup, ffx_cas.h:375-394, as called at taa.hpp:965 (input size == output size)
void taa_oracle_cas_setup(uint32_t const0[4], uint32_t const1[4], float sharpness, float inX, float inY, float outX, float outY) {
	const0[0] = f2u_bits(inX * (1.0f / outX));
	const0[1] = f2u_bits(inY * (1.0f / outY));
	const0[2] = f2u_bits(0.5f * inX * (1.0f / outX) - 0.5f);
	const0[3] = f2u_bits(0.5f * inY * (1.0f / outY) - 0.5f);
	float s = gmin(1.0f, gmax(0.0f, sharpness));       // ASatF1 (CPU: AMinF1(1, AMaxF1(0, a)), ffx_a.h:366)
	float lerp = 5.0f * s + (-8.0f * s + 8.0f);         // ALerpF1(8,5,s) = b*c+(-a*c+a), ffx_a.h:302
	float sharp = -(1.0f / lerp);
	const1[0] = f2u_bits(sharp);
	const1[1] = amd_f32_to_f16(sharp) + (amd_f32_to_f16(0.0f) << 16); // AU1_AH2_AF2, ffx_a.h:545
	const1[2] = f2u_bits(8.0f * inX * (1.0f / outX));
	const1[3] = 0;
}

// sharpen_cas.comp:30-53 + CasFilter(noScaling = true), ffx_cas.h:408-537. Unguarded imageLoad at -1 / w / h -> 0.
int taa_oracle_cas(const taa_image* src, const taa_image* dst, int w, int h, const uint32_t const0[4], const uint32_t const1[4], int nthreads) {
	if (!src || !dst || !src->data || !dst->data) return TAA_E_INVALID_ARG;
	(void)const0;
	Tex in = mkTex(*src, w, h);
	const float peak = u2f_bits(const1[0]);
#ifdef _OPENMP
	if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel for schedule(static) num_threads(nthreads)
#endif
	for (int y = 0; y < h; ++y) {
		for (int x = 0; x < w; ++x) {
			// the dispatch covers ceil(w/16) x ceil(h/16) tiles of 16x16 (taa.hpp:1134); imageStore outside the image is dropped
			vec3 b = rgb(fetch_rgba16f(in, x, y - 1));
			vec3 d = rgb(fetch_rgba16f(in, x - 1, y));
			vec3 e = rgb(fetch_rgba16f(in, x, y));
			vec3 f = rgb(fetch_rgba16f(in, x + 1, y));
			vec3 hh = rgb(fetch_rgba16f(in, x, y + 1));
			float mnR = AMin3F1(AMin3F1(d.x, e.x, f.x), b.x, hh.x);
			float mnG = AMin3F1(AMin3F1(d.y, e.y, f.y), b.y, hh.y);
			float mnB = AMin3F1(AMin3F1(d.z, e.z, f.z), b.z, hh.z);
			float mxR = AMax3F1(AMax3F1(d.x, e.x, f.x), b.x, hh.x);
			float mxG = AMax3F1(AMax3F1(d.y, e.y, f.y), b.y, hh.y);
			float mxB = AMax3F1(AMax3F1(d.z, e.z, f.z), b.z, hh.z);
			float rcpMR = APrxLoRcpF1(mxR), rcpMG = APrxLoRcpF1(mxG), rcpMB = APrxLoRcpF1(mxB);
			float ampR = ASatF1(gmin(mnR, 1.0f - mxR) * rcpMR);
			float ampG = ASatF1(gmin(mnG, 1.0f - mxG) * rcpMG);
			float ampB = ASatF1(gmin(mnB, 1.0f - mxB) * rcpMB);
			ampR = APrxLoSqrtF1(ampR);
			ampG = APrxLoSqrtF1(ampG);
			ampB = APrxLoSqrtF1(ampB);
			(void)ampR; (void)ampB; (void)mnR; (void)mnB; // only the green weight is used (ffx_cas.h:514-522)
			float wG = ampG * peak;
			float rcpWeight = APrxMedRcpF1(1.0f + 4.0f * wG);
			float pixR = ASatF1((b.x * wG + d.x * wG + f.x * wG + hh.x * wG + e.x) * rcpWeight);
			float pixG = ASatF1((b.y * wG + d.y * wG + f.y * wG + hh.y * wG + e.y) * rcpWeight);
			float pixB = ASatF1((b.z * wG + d.z * wG + f.z * wG + hh.z * wG + e.z) * rcpWeight);
			store_rgba16f(*dst, x, y, vec4{pixR, pixG, pixB, 1.0f});
		}
	}
	(void)nthreads;
	return TAA_OK;
}

// post_process.comp:29-88
int taa_oracle_post_process(const taa_image* src, const taa_image* debug, const taa_image* dst, int w, int h,
                            const TaaPostProcessPush* pc, int nthreads) {
	if (!src || !dst || !pc || !src->data || !dst->data) return TAA_E_INVALID_ARG;
	Tex in = mkTex(*src, w, h);
	Tex dbgT;
	if (debug && debug->data) dbgT = mkTex(*debug, w, h);
#ifdef _OPENMP
	if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel for schedule(static) num_threads(nthreads)
#endif
	for (int y = 0; y < h; ++y) {
		for (int x = 0; x < w; ++x) {
			ivec2 iuv = {x, y};
			ivec2 iuvFetchFrom = iuv;
			if (iuv.x == pc->splitX) { store_rgba16f(*dst, x, y, {0, 0, 0, 0}); continue; }
			const int32_t* S = pc->zoomSrcLTWH;
			const int32_t* D = pc->zoomDstLTWH;
			if (pc->zoom && pc->showZoomBox) {
				if ((((iuv.x == S[0] - 1) || (iuv.x == S[0] + S[2])) && (iuv.y >= S[1] - 1) && (iuv.y <= S[1] + S[3])) ||
				    (((iuv.y == S[1] - 1) || (iuv.y == S[1] + S[3])) && (iuv.x >= S[0] - 1) && (iuv.x <= S[0] + S[2]))) {
					store_rgba16f(*dst, x, y, {1, 0, 0, 0});
					continue;
				}
			}
			if (pc->zoom && iuv.x >= D[0] && iuv.y >= D[1] && iuv.x < D[0] + D[2] && iuv.y < D[1] + D[3]) {
				if ((((iuv.x == D[0]) || (iuv.x == D[0] + D[2] - 1)) && (iuv.y >= D[1]) && (iuv.y <= D[1] + D[3] - 1)) ||
				    (((iuv.y == D[1]) || (iuv.y == D[1] + D[3] - 1)) && (iuv.x >= D[0]) && (iuv.x <= D[0] + D[2] - 1))) {
					store_rgba16f(*dst, x, y, {1, 1, 1, 0});
					continue;
				}
				// vec2 zoomUv = (iuv - zoomDst.xy + 0.5) / vec2(zoomDst.zw);
				vec2 zoomUv = (toVec2(ivec2{iuv.x - D[0], iuv.y - D[1]}) + 0.5f) / toVec2(ivec2{D[2], D[3]});
				// iuvFetchFrom = ivec2(zoomSrc.xy + zoomUv * zoomSrc.zw);
				iuvFetchFrom = toIvec2(toVec2(ivec2{S[0], S[1]}) + zoomUv * toVec2(ivec2{S[2], S[3]}));
			}
			bool leftside = (pc->splitX < 0) || (iuv.x < pc->splitX);
			bool showdebug = leftside ? (pc->debugL_show != 0) : (pc->debugR_show != 0);
			const float* debugMask = leftside ? pc->debugL_mask : pc->debugR_mask;
			vec4 val;
			if (showdebug) {
				vec4 dbg = fetch_rgba16f(dbgT, iuvFetchFrom.x, iuvFetchFrom.y);
				val = {dbg.x, dbg.y, dbg.z, 1};
				if (debugMask[3] > 0.0f) { val.x += dbg.w; val.z += dbg.w; }
			} else {
				val = fetch_rgba16f(in, iuvFetchFrom.x, iuvFetchFrom.y);
			}
			store_rgba16f(*dst, x, y, val);
		}
	}
	(void)nthreads;
	return TAA_OK;
}

// antialias_fxaa_prepare.comp:15-26: rgb copied, alpha = luma (stored as fp16 like every rgba16f texel)
int taa_oracle_fxaa_prepare(const taa_image* src, const taa_image* dst, int w, int h, int nthreads) {
	if (!src || !dst || !src->data || !dst->data) return TAA_E_INVALID_ARG;
	Tex in = mkTex(*src, w, h);
#ifdef _OPENMP
	if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel for schedule(static) num_threads(nthreads)
#endif
	for (int y = 0; y < h; ++y)
		for (int x = 0; x < w; ++x) {
			vec3 c = rgb(fetch_rgba16f(in, x, y));
			store_rgba16f(*dst, x, y, mkvec4(c, dot(c, vec3{0.299f, 0.587f, 0.114f})));
		}
	(void)nthreads;
	return TAA_OK;
}

} // extern "C" (reopened below)

namespace {
// textureLodOffset(tex, p, 0, o): the integer offset is added to the footprint's texel indices before clamp-to-edge
inline vec4 sample_rgba16f_off(const Tex& t, vec2 uv, int ox, int oy) {
	float u = uv.x * (float)t.w - 0.5f, v = uv.y * (float)t.h - 0.5f;
	float fu = floorf(u), fv = floorf(v);
	float a = u - fu, b = v - fv;
	int i0 = f2i(fu) + ox, j0 = f2i(fv) + oy;
	int x0 = iclamp(i0, 0, t.w - 1), x1 = iclamp(i0 + 1, 0, t.w - 1), y0 = iclamp(j0, 0, t.h - 1), y1 = iclamp(j0 + 1, 0, t.h - 1);
	return lerp4(lerp4(fetch_rgba16f(t, x0, y0), fetch_rgba16f(t, x1, y0), a), lerp4(fetch_rgba16f(t, x0, y1), fetch_rgba16f(t, x1, y1), a), b);
}
// textureGather(Offset)(tex, p, comp = 3). The 2x2 footprint is the linear filter's; FXAA gathers AT a texel centre, where
// floor(u) flips with the last bit of the coordinate. Like hardware samplers (VkPhysicalDeviceLimits::subTexelPrecisionBits = 8
// on every desktop driver and on lavapipe) the unnormalised coordinate is snapped to 1/256 texel before the footprint is chosen.
// Returned order (Vulkan spec, "Texel Gathering"): x = (i0,j1), y = (i1,j1), z = (i1,j0), w = (i0,j0).
inline vec4 gather_alpha(const Tex& t, vec2 uv, int ox, int oy) {
	float u = uv.x * (float)t.w - 0.5f, v = uv.y * (float)t.h - 0.5f;
	float fu = floorf(floorf(u * 256.0f + 0.5f) * (1.0f / 256.0f)), fv = floorf(floorf(v * 256.0f + 0.5f) * (1.0f / 256.0f));
	int i0 = f2i(fu) + ox, j0 = f2i(fv) + oy;
	int x0 = iclamp(i0, 0, t.w - 1), x1 = iclamp(i0 + 1, 0, t.w - 1), y0 = iclamp(j0, 0, t.h - 1), y1 = iclamp(j0 + 1, 0, t.h - 1);
	return {fetch_rgba16f(t, x0, y1).w, fetch_rgba16f(t, x1, y1).w, fetch_rgba16f(t, x1, y0).w, fetch_rgba16f(t, x0, y0).w};
}

// FxaaPixelShader, FXAA_PC == 1, FXAA_QUALITY_PRESET 12, FXAA_DISCARD 0, FXAA_GREEN_AS_LUMA 0 (Fxaa3_11_mod.h:433-440, 884-1243;
// configured at antialias_fxaa.comp:23-27). gather4 selects the FXAA_GATHER4_ALPHA == 1 variant (:888-912), which is what a
// GLSL front end that predefines GL_ARB_gpu_shader5 (glslang does) compiles; 0 = the textureLodOffset variant (:913-925, 946-951).
// The nested doneNP blocks of the header (:1030-1190) are one loop over the step table here.
vec4 fxaa_pixel(const Tex& tex, vec2 pos, vec2 rcpFrame, float subpix, float edgeThreshold, float edgeThresholdMin, bool gather4) {
	static const float P[5] = {1.0f, 1.5f, 2.0f, 4.0f, 12.0f};  // FXAA_QUALITY_P0..P4, FXAA_QUALITY_PS = 5
	vec2 posM = pos;
	vec4 rgbyM = sample_rgba16f(tex, posM);
	const float lumaM = rgbyM.w;
	float lumaS, lumaE, lumaN, lumaW, lumaNW = 0, lumaSE = 0, lumaNE, lumaSW;
	if (gather4) {
		vec4 A = gather_alpha(tex, posM, 0, 0), B = gather_alpha(tex, posM, -1, -1);
		lumaE = A.z; lumaS = A.x; lumaSE = A.y; lumaNW = B.w; lumaN = B.z; lumaW = B.x;
	} else {
		lumaS = sample_rgba16f_off(tex, posM, 0, 1).w;
		lumaE = sample_rgba16f_off(tex, posM, 1, 0).w;
		lumaN = sample_rgba16f_off(tex, posM, 0, -1).w;
		lumaW = sample_rgba16f_off(tex, posM, -1, 0).w;
	}
	float maxSM = gmax(lumaS, lumaM), minSM = gmin(lumaS, lumaM);
	float maxESM = gmax(lumaE, maxSM), minESM = gmin(lumaE, minSM);
	float maxWN = gmax(lumaN, lumaW), minWN = gmin(lumaN, lumaW);
	float rangeMax = gmax(maxWN, maxESM), rangeMin = gmin(minWN, minESM);
	float rangeMaxScaled = rangeMax * edgeThreshold;
	float range = rangeMax - rangeMin;
	float rangeMaxClamped = gmax(edgeThresholdMin, rangeMaxScaled);
	if (range < rangeMaxClamped) return rgbyM;  // earlyExit
	if (gather4) {
		lumaNE = sample_rgba16f_off(tex, posM, 1, -1).w;
		lumaSW = sample_rgba16f_off(tex, posM, -1, 1).w;
	} else {
		lumaNW = sample_rgba16f_off(tex, posM, -1, -1).w;
		lumaSE = sample_rgba16f_off(tex, posM, 1, 1).w;
		lumaNE = sample_rgba16f_off(tex, posM, 1, -1).w;
		lumaSW = sample_rgba16f_off(tex, posM, -1, 1).w;
	}
	float lumaNS = lumaN + lumaS, lumaWE = lumaW + lumaE;
	float subpixRcpRange = 1.0f / range;
	float subpixNSWE = lumaNS + lumaWE;
	float edgeHorz1 = (-2.0f * lumaM) + lumaNS, edgeVert1 = (-2.0f * lumaM) + lumaWE;
	float lumaNESE = lumaNE + lumaSE, lumaNWNE = lumaNW + lumaNE;
	float edgeHorz2 = (-2.0f * lumaE) + lumaNESE, edgeVert2 = (-2.0f * lumaN) + lumaNWNE;
	float lumaNWSW = lumaNW + lumaSW, lumaSWSE = lumaSW + lumaSE;
	float edgeHorz4 = (fabsf(edgeHorz1) * 2.0f) + fabsf(edgeHorz2), edgeVert4 = (fabsf(edgeVert1) * 2.0f) + fabsf(edgeVert2);
	float edgeHorz3 = (-2.0f * lumaW) + lumaNWSW, edgeVert3 = (-2.0f * lumaS) + lumaSWSE;
	float edgeHorz = fabsf(edgeHorz3) + edgeHorz4, edgeVert = fabsf(edgeVert3) + edgeVert4;
	float subpixNWSWNESE = lumaNWSW + lumaNESE;
	float lengthSign = rcpFrame.x;
	bool horzSpan = edgeHorz >= edgeVert;
	float subpixA = subpixNSWE * 2.0f + subpixNWSWNESE;
	if (!horzSpan) lumaN = lumaW;
	if (!horzSpan) lumaS = lumaE;
	if (horzSpan) lengthSign = rcpFrame.y;
	float subpixB = (subpixA * (1.0f / 12.0f)) - lumaM;
	float gradientN = lumaN - lumaM, gradientS = lumaS - lumaM;
	float lumaNN = lumaN + lumaM, lumaSS = lumaS + lumaM;
	bool pairN = fabsf(gradientN) >= fabsf(gradientS);
	float gradient = gmax(fabsf(gradientN), fabsf(gradientS));
	if (pairN) lengthSign = -lengthSign;
	float subpixC = gclamp(fabsf(subpixB) * subpixRcpRange, 0.0f, 1.0f);
	vec2 posB = posM;
	vec2 offNP = {(!horzSpan) ? 0.0f : rcpFrame.x, horzSpan ? 0.0f : rcpFrame.y};
	if (!horzSpan) posB.x += lengthSign * 0.5f;
	if (horzSpan) posB.y += lengthSign * 0.5f;
	vec2 posN = {posB.x - offNP.x * P[0], posB.y - offNP.y * P[0]};
	vec2 posP = {posB.x + offNP.x * P[0], posB.y + offNP.y * P[0]};
	float subpixD = ((-2.0f) * subpixC) + 3.0f;
	float lumaEndN = sample_rgba16f(tex, posN).w;
	float subpixE = subpixC * subpixC;
	float lumaEndP = sample_rgba16f(tex, posP).w;
	if (!pairN) lumaNN = lumaSS;
	float gradientScaled = gradient * 1.0f / 4.0f;
	float lumaMM = lumaM - lumaNN * 0.5f;
	float subpixF = subpixD * subpixE;
	bool lumaMLTZero = lumaMM < 0.0f;
	lumaEndN -= lumaNN * 0.5f;
	lumaEndP -= lumaNN * 0.5f;
	bool doneN = fabsf(lumaEndN) >= gradientScaled, doneP = fabsf(lumaEndP) >= gradientScaled;
	for (int i = 1;; ++i) {
		if (!doneN) { posN.x -= offNP.x * P[i]; posN.y -= offNP.y * P[i]; }
		bool doneNP = (!doneN) || (!doneP);
		if (!doneP) { posP.x += offNP.x * P[i]; posP.y += offNP.y * P[i]; }
		if (!doneNP || i == 4) break;
		if (!doneN) lumaEndN = sample_rgba16f(tex, posN).w;
		if (!doneP) lumaEndP = sample_rgba16f(tex, posP).w;
		if (!doneN) lumaEndN = lumaEndN - lumaNN * 0.5f;
		if (!doneP) lumaEndP = lumaEndP - lumaNN * 0.5f;
		doneN = fabsf(lumaEndN) >= gradientScaled;
		doneP = fabsf(lumaEndP) >= gradientScaled;
	}
	float dstN = posM.x - posN.x, dstP = posP.x - posM.x;
	if (!horzSpan) dstN = posM.y - posN.y;
	if (!horzSpan) dstP = posP.y - posM.y;
	bool goodSpanN = (lumaEndN < 0.0f) != lumaMLTZero;
	float spanLength = dstP + dstN;
	bool goodSpanP = (lumaEndP < 0.0f) != lumaMLTZero;
	float spanLengthRcp = 1.0f / spanLength;
	bool directionN = dstN < dstP;
	float dst = gmin(dstN, dstP);
	bool goodSpan = directionN ? goodSpanN : goodSpanP;
	float subpixG = subpixF * subpixF;
	float pixelOffset = (dst * (-spanLengthRcp)) + 0.5f;
	float subpixH = subpixG * subpix;
	float pixelOffsetGood = goodSpan ? pixelOffset : 0.0f;
	float pixelOffsetSubpix = gmax(pixelOffsetGood, subpixH);
	if (!horzSpan) posM.x += pixelOffsetSubpix * lengthSign;
	if (horzSpan) posM.y += pixelOffsetSubpix * lengthSign;
	return mkvec4(rgb(sample_rgba16f(tex, posM)), lumaM);
}
}  // namespace

extern "C" {

// the pinned sin / cos, for the unit test that bounds their error against libm
void taa_oracle_sincos(const float* x, float* s, float* c, int n) {
	for (int i = 0; i < n; ++i) taa_sincos(x[i], &s[i], &c[i]);
}

// antialias_fxaa.comp:30-64: pixels whose seg-mask says 1 (FXAA) run FxaaPixelShader on the prepared image (luma in alpha), all
// others are copied. gather4: see fxaa_pixel.
int taa_oracle_fxaa(const taa_image* src, const taa_image* segmask, const taa_image* dst, int w, int h, const TaaFxaaPush* pc, int gather4, int nthreads) {
	if (!src || !segmask || !dst || !pc || !src->data || !segmask->data || !dst->data) return TAA_E_INVALID_ARG;
	Tex in = mkTex(*src, w, h), seg = mkTex(*segmask, w, h);
#ifdef _OPENMP
	if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 4) num_threads(nthreads)
#endif
	for (int y = 0; y < h; ++y)
		for (int x = 0; x < w; ++x) {
			vec4 color;
			if ((fetch_r32ui(seg, x, y) & 3u) == 1u) {
				vec2 rcp = {pc->fxaaQualityRcpFrame[0], pc->fxaaQualityRcpFrame[1]};
				color = fxaa_pixel(in, (toVec2(ivec2{x, y}) + 0.5f) * rcp, rcp, pc->fxaaQualitySubpix, pc->fxaaQualityEdgeThreshold, pc->fxaaQualityEdgeThresholdMin,
				                   gather4 != 0);
			} else {
				color = fetch_rgba16f(in, x, y);
			}
			store_rgba16f(*dst, x, y, color);
		}
	(void)nthreads;
	return TAA_OK;
}

// helpers::halton, helper_functions.hpp:9-17
float taa_oracle_halton(int i, int b) {
	float f = 1.0f, r = 0.0f;
	while (i > 0) {
		f = f / (float)b;
		r = r + f * (float)(i % b);
		i = i / b;
	}
	return r;
}

// get_jitter_offset_for_frame, taa.hpp:150-233. Returns the pattern length.
int taa_oracle_jitter(int sampleDistribution, int fixedJitterIndex, float extraScale, int slowMotion, float rotateDegrees,
                      const float* debugOffsets, int debugCount, int in_w, int in_h, long long frameId, float out_ndc[2]) {
	const float pxx = 2.0f / (float)in_w, pxy = 2.0f / (float)in_h; // sPxSizeNDC, taa.hpp:155
	float pat[16][2];
	int n = 0;
	float scx = 1.0f, scy = 1.0f;
	const float eighth = 1.f / 8.f;
	switch (sampleDistribution) {
		case 0: { const float q[4][2] = {{-0.25f, -0.25f}, {0.25f, -0.25f}, {0.25f, 0.25f}, {-0.25f, 0.25f}};
			n = 4; for (int i = 0; i < 4; ++i) { pat[i][0] = pxx * q[i][0]; pat[i][1] = pxy * q[i][1]; } break; }
		case 1: { const float q[4][2] = {{-0.25f, -0.25f}, {0.25f, 0.25f}, {0.25f, -0.25f}, {-0.25f, 0.25f}};
			n = 4; for (int i = 0; i < 4; ++i) { pat[i][0] = pxx * q[i][0]; pat[i][1] = pxy * q[i][1]; } break; }
		case 2: case 3: {
			n = sampleDistribution == 2 ? 8 : 16;
			for (int i = 0; i < n; ++i) { pat[i][0] = pxx * (taa_oracle_halton(i + 1, 2) - 0.5f); pat[i][1] = pxy * (taa_oracle_halton(i + 1, 3) - 0.5f); }
			break; }
		case 4: {
			n = 16;
			const float c[4] = {-3.f * eighth, -1.f * eighth, 1.f * eighth, 3.f * eighth};
			for (int j = 0; j < 4; ++j) for (int i = 0; i < 4; ++i) { pat[j * 4 + i][0] = pxx * c[i]; pat[j * 4 + i][1] = pxy * c[j]; }
			break; }
		case 5: {
			if (!debugOffsets || debugCount <= 0) return TAA_E_INVALID_ARG;
			if (slowMotion > 1) frameId /= slowMotion;
			if (fixedJitterIndex >= 0) frameId = fixedJitterIndex;
			int idx = (int)(frameId % debugCount);
			float px = debugOffsets[idx * 2] * pxx, py = debugOffsets[idx * 2 + 1] * pxy;
			if (rotateDegrees != 0.f) {
				float rad = rotateDegrees * 0.01745329251994329576923690768489f; // glm::radians
				float s = sinf(rad), c = cosf(rad);
				float nx = px * c - py * s, ny = px * s + py * c;
				px = nx; py = ny;
			}
			out_ndc[0] = px * extraScale; out_ndc[1] = py * extraScale;
			return debugCount; }
		default: return TAA_E_INVALID_ARG;
	}
	(void)scx; (void)scy;
	if (slowMotion > 1) frameId /= slowMotion;
	if (fixedJitterIndex >= 0) frameId = fixedJitterIndex;
	int idx = (int)(frameId % n);
	float px = pat[idx][0], py = pat[idx][1];
	if (rotateDegrees != 0.f) {
		float rad = rotateDegrees * 0.01745329251994329576923690768489f;
		float s = sinf(rad), c = cosf(rad);
		float nx = px * c - py * s, ny = px * s + py * c;
		px = nx; py = ny;
	}
	out_ndc[0] = px * extraScale;
	out_ndc[1] = py * extraScale;
	return n;
}

// bulk fp16 <-> fp32 (RTE), for building test inputs without numpy's float16 path
void taa_oracle_f32_to_f16(const float* src, uint16_t* dst, long long n) { for (long long i = 0; i < n; ++i) dst[i] = float_to_half(src[i]); }
void taa_oracle_f16_to_f32(const uint16_t* src, float* dst, long long n) { for (long long i = 0; i < n; ++i) dst[i] = half_to_float(src[i]); }

} // extern "C"
