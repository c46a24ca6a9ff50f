/*
 * glsl_shim.h — just enough GLSL 4.60 compute semantics in C++17 to compile the REFERENCE's own shader sources
 * (shaders/taa.comp, sharpen.comp, sharpen_cas.comp + ffx_cas.h + the used part of ffx_a.h, post_process.comp, antialias_fxaa_prepare.comp, antialias_fxaa.comp + Fxaa3_11_mod.h, read where they lie under /root/reference by
 * oracle/ref_build.py) into oracle/_ref/libtaa_ref.so. TEST INFRASTRUCTURE, like everything under oracle/.
 *
 * This file holds no algorithm of the reference: it is the "GLSL machine" (vector types, built-ins, the sampler
 * and image access). What the machine leaves to the implementation follows the same arithmetic model as
 * oracle/taa_oracle.cpp (its header lists the choices): one IEEE binary32 operation per GLSL operation, no
 * contraction (-ffp-contract=off), mat4*vec4 accumulated column by column, sampler = Vulkan linear /
 * clamp-to-edge / normalised coordinates in fp32 with lerp(p,q,w) = p + w*(q-p), fp32->fp16 stores round to
 * nearest even, out-of-range texelFetch / imageLoad return 0, float->int conversions truncate and saturate.
 *
 * ref_build.py rewrites the shader text mechanically: comments stripped, `layout(...)` resource declarations turned
 * into thread_local globals, in/out/inout parameters into values/references, floating literals suffixed with f,
 * multi-component swizzles `.rgb` into `._rgb()`, `bool` inside interface blocks into the 4-byte bool32, `main`
 * renamed. Nothing else of the source is touched. sharpen_cas.comp: of its include ffx_a.h (AMD's ~1900-line portability header) only the
 * type macros and the functions the shader uses are spliced in by name, of ffx_cas.h the non-packed GPU section; its imageLoad calls (float
 * image) become imageLoadF; members of its instance-less push-constant block become globals.
 */
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace glsl {

typedef unsigned int uint;

// ---- scalar conversions ---------------------------------------------------------------------------------
inline int f2i(float f) {
	if (f != f) return 0;
	if (f >= 2147483648.0f) return 2147483647;
	if (f <= -2147483648.0f) return (-2147483647 - 1);
	return (int)f;
}
struct bool32 {  // std140 bool
	uint32_t v;
	operator bool() const { return v != 0; }
};

// ---- vectors ----------------------------------------------------------------------------------------------
struct ivec2;
struct uvec2;
struct vec2 {
	union { float x, r, s; };
	union { float y, g, t; };
	vec2() : x(0), y(0) {}
	template <class A> explicit vec2(A a) : x((float)a), y((float)a) {}
	template <class A, class B> vec2(A a, B b) : x((float)a), y((float)b) {}
	vec2(const ivec2& v);  // GLSL implicit int -> float conversion
	explicit vec2(const uvec2& v);
	vec2(const vec2& o) : x(o.x), y(o.y) {}
	vec2& operator=(const vec2& o) { x = o.x; y = o.y; return *this; }
	vec2 _xy() const { return vec2(x, y); }
};
struct Swz2 {  // an l-value swizzle of two components
	float &a, &b;
	operator vec2() const { return vec2(a, b); }
	Swz2& operator=(const vec2& v) { a = v.x; b = v.y; return *this; }
	Swz2& operator=(const Swz2& o) { const float p = o.a, q = o.b; a = p; b = q; return *this; }
	Swz2& operator+=(const vec2& v) { a = a + v.x; b = b + v.y; return *this; }
	Swz2& operator+=(float v) { a = a + v; b = b + v; return *this; }
};
struct vec3 {
	union { float x, r; };
	union { float y, g; };
	union { float z, b; };
	vec3() : x(0), y(0), z(0) {}
	template <class A> explicit vec3(A a) : x((float)a), y((float)a), z((float)a) {}
	template <class A, class B, class C> vec3(A a, B b_, C c) : x((float)a), y((float)b_), z((float)c) {}
	vec3(const vec2& v, float c) : x(v.x), y(v.y), z(c) {}
	vec3(const vec3& o) : x(o.x), y(o.y), z(o.z) {}
	vec3& operator=(const vec3& o) { x = o.x; y = o.y; z = o.z; return *this; }
	Swz2 _xy() { return Swz2{x, y}; }
	vec2 _xy() const { return vec2(x, y); }
	Swz2 _gb() { return Swz2{y, z}; }
	vec2 _gb() const { return vec2(y, z); }
	Swz2 _rb() { return Swz2{x, z}; }
};
struct Swz3 {
	float &a, &b, &c;
	operator vec3() const { return vec3(a, b, c); }
	Swz3& operator=(const vec3& v) { a = v.x; b = v.y; c = v.z; return *this; }
	Swz3& operator=(const Swz3& o) { const float p = o.a, q = o.b, r = o.c; a = p; b = q; c = r; return *this; }
	Swz3& operator+=(const vec3& v) { a = a + v.x; b = b + v.y; c = c + v.z; return *this; }
};
struct vec4 {
	union { float x, r; };
	union { float y, g; };
	union { float z, b; };
	union { float w, a; };
	vec4() : x(0), y(0), z(0), w(0) {}
	template <class A> explicit vec4(A v) : x((float)v), y((float)v), z((float)v), w((float)v) {}
	template <class A, class B, class C, class D> vec4(A a_, B b_, C c, D d) : x((float)a_), y((float)b_), z((float)c), w((float)d) {}
	template <class D> vec4(const vec3& v, D d) : x(v.x), y(v.y), z(v.z), w((float)d) {}
	template <class C, class D> vec4(const vec2& v, C c, D d) : x(v.x), y(v.y), z((float)c), w((float)d) {}
	template <class A> vec4(A a_, const vec2& v, float d) : x((float)a_), y(v.x), z(v.y), w(d) {}
	vec4(const vec4& o) : x(o.x), y(o.y), z(o.z), w(o.w) {}
	vec4& operator=(const vec4& o) { x = o.x; y = o.y; z = o.z; w = o.w; return *this; }
	Swz3 _rgb() { return Swz3{x, y, z}; }
	vec3 _rgb() const { return vec3(x, y, z); }
	Swz3 _xyz() { return Swz3{x, y, z}; }
	vec3 _xyz() const { return vec3(x, y, z); }
	Swz2 _xy() { return Swz2{x, y}; }
	vec2 _xy() const { return vec2(x, y); }
	Swz2 _rg() { return Swz2{x, y}; }
	Swz2 _rb() { return Swz2{x, z}; }
};
struct ivec2 {
	int x, y;
	ivec2() : x(0), y(0) {}
	explicit ivec2(int a) : x(a), y(a) {}
	ivec2(int a, int b) : x(a), y(b) {}
	explicit ivec2(const vec2& v) : x(f2i(v.x)), y(f2i(v.y)) {}
	explicit ivec2(const uvec2& v);
};
struct ivec4 {
	int x, y, z, w;
	ivec2 _xy() const { return ivec2(x, y); }
	ivec2 _zw() const { return ivec2(z, w); }
};
struct uvec2 {
	uint x, y;
	uvec2() : x(0), y(0) {}
	uvec2(uint a, uint b) : x(a), y(b) {}
};
struct uvec3 {
	uint x, y, z;
	uvec3() : x(0), y(0), z(0) {}
	uvec3(uint a, uint b, uint c) : x(a), y(b), z(c) {}
	uvec2 _xy() const { return uvec2(x, y); }
};
struct uvec4 {
	union { uint x, r; };
	uint y, z, w;
	uvec4() : x(0), y(0), z(0), w(0) {}
	uvec4(uint a, uint b, uint c, uint d) : x(a), y(b), z(c), w(d) {}
	uvec2 _xy() const { return uvec2(x, y); }
	uvec2 _zw() const { return uvec2(z, w); }
};
struct bvec2 { bool x, y; };
struct bvec3 { bool x, y, z; };
inline vec2::vec2(const ivec2& v) : x((float)v.x), y((float)v.y) {}
inline vec2::vec2(const uvec2& v) : x((float)v.x), y((float)v.y) {}
inline uvec2 operator+(const uvec2& a, const uvec2& b) { return uvec2(a.x + b.x, a.y + b.y); }
inline ivec2::ivec2(const uvec2& v) : x((int)v.x), y((int)v.y) {}

// every operator is one fp32 operation per component
#define GLSL_VEC_OPS(V, N, ...)                                                                                                          \
	inline V operator+(const V& a, const V& b) { V r; for (int i = 0; i < N; ++i) (&r.x)[i] = (&a.x)[i] + (&b.x)[i]; return r; }         \
	inline V operator-(const V& a, const V& b) { V r; for (int i = 0; i < N; ++i) (&r.x)[i] = (&a.x)[i] - (&b.x)[i]; return r; }         \
	inline V operator*(const V& a, const V& b) { V r; for (int i = 0; i < N; ++i) (&r.x)[i] = (&a.x)[i] * (&b.x)[i]; return r; }         \
	inline V operator/(const V& a, const V& b) { V r; for (int i = 0; i < N; ++i) (&r.x)[i] = (&a.x)[i] / (&b.x)[i]; return r; }         \
	inline V operator+(const V& a, float b) { V r; for (int i = 0; i < N; ++i) (&r.x)[i] = (&a.x)[i] + b; return r; }                    \
	inline V operator-(const V& a, float b) { V r; for (int i = 0; i < N; ++i) (&r.x)[i] = (&a.x)[i] - b; return r; }                    \
	inline V operator*(const V& a, float b) { V r; for (int i = 0; i < N; ++i) (&r.x)[i] = (&a.x)[i] * b; return r; }                    \
	inline V operator/(const V& a, float b) { V r; for (int i = 0; i < N; ++i) (&r.x)[i] = (&a.x)[i] / b; return r; }                    \
	inline V operator+(float a, const V& b) { V r; for (int i = 0; i < N; ++i) (&r.x)[i] = a + (&b.x)[i]; return r; }                    \
	inline V operator-(float a, const V& b) { V r; for (int i = 0; i < N; ++i) (&r.x)[i] = a - (&b.x)[i]; return r; }                    \
	inline V operator*(float a, const V& b) { V r; for (int i = 0; i < N; ++i) (&r.x)[i] = a * (&b.x)[i]; return r; }                    \
	inline V operator/(float a, const V& b) { V r; for (int i = 0; i < N; ++i) (&r.x)[i] = a / (&b.x)[i]; return r; }                    \
	inline V operator-(const V& a) { V r; for (int i = 0; i < N; ++i) (&r.x)[i] = -(&a.x)[i]; return r; }                               \
	inline V& operator+=(V& a, const V& b) { a = a + b; return a; }                                                                      \
	inline V& operator-=(V& a, const V& b) { a = a - b; return a; }                                                                      \
	inline V& operator*=(V& a, const V& b) { a = a * b; return a; }                                                                      \
	inline V& operator*=(V& a, float b) { a = a * b; return a; }                                                                         \
	inline V abs(const V& a) { V r; for (int i = 0; i < N; ++i) (&r.x)[i] = std::fabs((&a.x)[i]); return r; }                            \
	inline V floor(const V& a) { V r; for (int i = 0; i < N; ++i) (&r.x)[i] = std::floor((&a.x)[i]); return r; }                         \
	inline V sqrt(const V& a) { V r; for (int i = 0; i < N; ++i) (&r.x)[i] = std::sqrt((&a.x)[i]); return r; }                           \
	inline V fract(const V& a) { V r; for (int i = 0; i < N; ++i) (&r.x)[i] = (&a.x)[i] - std::floor((&a.x)[i]); return r; }             \
	inline V min(const V& a, const V& b) { V r; for (int i = 0; i < N; ++i) (&r.x)[i] = fmin_((&a.x)[i], (&b.x)[i]); return r; }         \
	inline V max(const V& a, const V& b) { V r; for (int i = 0; i < N; ++i) (&r.x)[i] = fmax_((&a.x)[i], (&b.x)[i]); return r; }         \
	inline V clamp(const V& v, const V& lo, const V& hi) { return min(max(v, lo), hi); }                                                 \
	inline V mix(const V& a, const V& b, float t) { return a * (1.0f - t) + b * t; }                                                     \
	inline float dot(const V& a, const V& b) { float s = a.x * b.x; for (int i = 1; i < N; ++i) s = s + (&a.x)[i] * (&b.x)[i]; return s; } \
	inline float length(const V& a) { return std::sqrt(dot(a, a)); }

// IEEE-754-2019 minimumNumber / maximumNumber, like the oracle (and CUDA fminf / fmaxf)
inline float fmin_(float a, float b) {
	if (a != a) return b;
	if (b != b) return a;
	if (a == b) return std::signbit(a) ? a : b;
	return a < b ? a : b;
}
inline float fmax_(float a, float b) {
	if (a != a) return b;
	if (b != b) return a;
	if (a == b) return std::signbit(a) ? b : a;
	return a > b ? a : b;
}
GLSL_VEC_OPS(vec2, 2)
GLSL_VEC_OPS(vec3, 3)
GLSL_VEC_OPS(vec4, 4)

// scalar built-ins (int arguments convert to float, as GLSL's implicit conversions do)
inline float min(float a, float b) { return fmin_(a, b); }
inline float max(float a, float b) { return fmax_(a, b); }
inline float clamp(float v, float lo, float hi) { return fmin_(fmax_(v, lo), hi); }
inline float abs(float a) { return std::fabs(a); }
inline float floor(float a) { return std::floor(a); }
inline float sqrt(float a) { return std::sqrt(a); }
inline float fract(float a) { return a - std::floor(a); }
inline float sin(float a);
inline float cos(float a);
inline float mix(float a, float b, float t) { return a * (1.0f - t) + b * t; }

// sin and cos as this repository pins them. GLSL leaves their precision to the implementation (and no Vulkan driver exists on either box), so the
// oracle, the reference-shader shim (oracle/glsl_shim.h) and the CUDA kernels (taa_device.cuh) all evaluate THIS function text: Cody-Waite
// reduction by pi/2 in three steps, the single-precision minimax polynomials of the Cephes library on [-pi/4, pi/4], one IEEE binary32 operation
// per written operation (the three files are compiled without contraction). Arguments too large to reduce (>= 1e9) read as 0; non-finite ones give NaN.
inline void taa_sincos(float x, float* s, float* c) {
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
inline float taa_sin(float x) { float s, c; taa_sincos(x, &s, &c); return s; }
inline float taa_cos(float x) { float s, c; taa_sincos(x, &s, &c); return c; }
inline float sin(float a) { return taa_sin(a); }
inline float cos(float a) { return taa_cos(a); }

// integer vectors
inline ivec2 operator+(const ivec2& a, const ivec2& b) { return ivec2(a.x + b.x, a.y + b.y); }
inline ivec2 operator-(const ivec2& a, const ivec2& b) { return ivec2(a.x - b.x, a.y - b.y); }
inline ivec2 operator-(const ivec2& a, int b) { return ivec2(a.x - b, a.y - b); }
inline ivec2 operator+(const ivec2& a, int b) { return ivec2(a.x + b, a.y + b); }
// int vector (op) float scalar: GLSL converts the vector to float first (exact-match overloads, so that the int ones above are not picked)
inline vec2 operator+(const ivec2& a, float b) { return vec2(a) + b; }
inline vec2 operator-(const ivec2& a, float b) { return vec2(a) - b; }
inline vec2 operator*(const ivec2& a, float b) { return vec2(a) * b; }
inline bool operator==(const ivec2& a, const ivec2& b) { return a.x == b.x && a.y == b.y; }
inline ivec2 min(const ivec2& a, const ivec2& b) { return ivec2(a.x < b.x ? a.x : b.x, a.y < b.y ? a.y : b.y); }
inline ivec2 max(const ivec2& a, const ivec2& b) { return ivec2(a.x > b.x ? a.x : b.x, a.y > b.y ? a.y : b.y); }
inline ivec2 clamp(const ivec2& v, const ivec2& lo, const ivec2& hi) { return min(max(v, lo), hi); }

// relational
inline bvec2 lessThan(const vec2& a, const vec2& b) { return bvec2{a.x < b.x, a.y < b.y}; }
inline bvec2 greaterThanEqual(const vec2& a, const vec2& b) { return bvec2{a.x >= b.x, a.y >= b.y}; }
inline bvec2 lessThan(const ivec2& a, const ivec2& b) { return bvec2{a.x < b.x, a.y < b.y}; }
inline bvec2 greaterThanEqual(const ivec2& a, const ivec2& b) { return bvec2{a.x >= b.x, a.y >= b.y}; }
inline bvec3 lessThan(const vec3& a, const vec3& b) { return bvec3{a.x < b.x, a.y < b.y, a.z < b.z}; }
inline bvec3 greaterThan(const vec3& a, const vec3& b) { return bvec3{a.x > b.x, a.y > b.y, a.z > b.z}; }
inline bool any(const bvec2& v) { return v.x || v.y; }
inline bool any(const bvec3& v) { return v.x || v.y || v.z; }
inline bool all(const bvec2& v) { return v.x && v.y; }

// mat4: column-major; M * v accumulates column by column, ((M0 v.x + M1 v.y) + M2 v.z) + M3 v.w
struct mat4 { vec4 c[4]; };
inline vec4 operator*(const mat4& m, const vec4& v) { return ((m.c[0] * v.x + m.c[1] * v.y) + m.c[2] * v.z) + m.c[3] * v.w; }

// bit casts and bit-field built-ins (GLSL 4.60 8.3, 8.8)
inline float uintBitsToFloat(uint v) { float f; memcpy(&f, &v, 4); return f; }
inline uint floatBitsToUint(float f) { uint v; memcpy(&v, &f, 4); return v; }
inline vec2 uintBitsToFloat(const uvec2& v) { return vec2(uintBitsToFloat(v.x), uintBitsToFloat(v.y)); }
inline uint bitfieldExtract(uint value, int offset, int bits) { return bits == 0 ? 0u : (bits >= 32 ? value >> offset : (value >> offset) & ((1u << bits) - 1u)); }
inline uint bitfieldInsert(uint base, uint insert, int offset, int bits) {
	if (bits == 0) return base;
	const uint mask = (bits >= 32 ? 0xffffffffu : ((1u << bits) - 1u)) << offset;
	return (base & ~mask) | ((insert << offset) & mask);
}

// ---- fp16 -------------------------------------------------------------------------------------------------
inline float half_to_float(uint16_t h) {
	uint32_t sign = (uint32_t)(h & 0x8000u) << 16, exp = (h >> 10) & 0x1fu, man = h & 0x3ffu, out;
	if (exp == 0) {
		if (man == 0) out = sign;
		else {
			int e = -1;
			do { e++; man <<= 1; } while ((man & 0x400u) == 0);
			out = sign | ((uint32_t)(127 - 15 - e) << 23) | ((man & 0x3ffu) << 13);
		}
	} else if (exp == 31) out = sign | 0x7f800000u | (man << 13);
	else out = sign | ((exp + 112) << 23) | (man << 13);
	float f;
	memcpy(&f, &out, 4);
	return f;
}
inline uint16_t float_to_half(float f) {  // round to nearest even
	uint32_t x;
	memcpy(&x, &f, 4);
	uint32_t sign = (x >> 16) & 0x8000u, ax = x & 0x7fffffffu;
	if (ax >= 0x7f800000u) return (uint16_t)(ax > 0x7f800000u ? (sign | 0x7e00u | ((ax >> 13) & 0x3ffu)) : (sign | 0x7c00u));
	if (ax >= 0x477ff000u) return (uint16_t)(sign | 0x7c00u);
	if (ax < 0x33000001u) return (uint16_t)sign;
	int e = (int)(ax >> 23) - 127;
	uint32_t m = (ax & 0x7fffffu) | 0x800000u, base;
	int shift;
	if (e < -14) { shift = 13 + (-14 - e); base = 0; } else { shift = 13; base = (uint32_t)(e + 15) << 10; m &= 0x7fffffu; }
	uint32_t q = m >> shift, rem = m & ((1u << shift) - 1u), half = 1u << (shift - 1);
	if (rem > half || (rem == half && (q & 1u))) q++;
	return (uint16_t)(sign | (base + q));
}

// ---- resources --------------------------------------------------------------------------------------------
enum Format { F_NONE = 0, F_RGBA16F, F_R32F, F_RGBA32F, F_R32UI };
struct Image {
	unsigned char* data = nullptr;
	long long pitch = 0;
	int w = 0, h = 0;
	Format fmt = F_NONE;
};
typedef Image texture2D;
typedef Image image2D;
typedef Image uimage2D;
struct sampler {};
struct sampler2D {  // `sampler2D(tex, smp)` (separate texture + sampler, taa.comp) or a combined binding that the harness points at an image
	const Image* t = nullptr;
	sampler2D() {}
	sampler2D(const Image& tex, const sampler&) : t(&tex) {}
};
typedef sampler2D sampler2D_t;
inline ivec2 textureSize(const sampler2D& s, int) { return ivec2(s.t->w, s.t->h); }
inline ivec2 textureSize(const Image& t, int) { return ivec2(t.w, t.h); }
inline ivec2 imageSize(const Image& t) { return ivec2(t.w, t.h); }

inline vec4 texel(const Image& t, int x, int y) {  // in range
	const unsigned char* p = t.data + (long long)y * t.pitch;
	switch (t.fmt) {
		case F_RGBA16F: { const uint16_t* q = (const uint16_t*)p + 4 * x; return vec4(half_to_float(q[0]), half_to_float(q[1]), half_to_float(q[2]), half_to_float(q[3])); }
		case F_R32F: return vec4(((const float*)p)[x], 0.f, 0.f, 1.f);
		case F_RGBA32F: { const float* q = (const float*)p + 4 * x; return vec4(q[0], q[1], q[2], q[3]); }
		default: return vec4(0);
	}
}
inline vec4 texelFetch(const Image& t, const ivec2& c, int) {
	if (!t.data || c.x < 0 || c.y < 0 || c.x >= t.w || c.y >= t.h) return vec4(0);
	return texel(t, c.x, c.y);
}
inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
inline vec4 lerp4(const vec4& p, const vec4& q, float w) { return p + w * (q - p); }
inline vec4 texture(const sampler2D_t& s, const vec2& uv) {
	const Image& t = *s.t;
	const float u = uv.x * (float)t.w - 0.5f, v = uv.y * (float)t.h - 0.5f;
	const float fu = std::floor(u), fv = std::floor(v);
	const float a = u - fu, b = v - fv;
	const int i0 = f2i(fu), j0 = f2i(fv);
	const int x0 = clampi(i0, 0, t.w - 1), x1 = clampi(i0 == 2147483647 ? i0 : i0 + 1, 0, t.w - 1);
	const int y0 = clampi(j0, 0, t.h - 1), y1 = clampi(j0 == 2147483647 ? j0 : j0 + 1, 0, t.h - 1);
	return lerp4(lerp4(texel(t, x0, y0), texel(t, x1, y0), a), lerp4(texel(t, x0, y1), texel(t, x1, y1), a), b);
}
inline vec4 texelFetch(const sampler2D& s, const ivec2& c, int) { return texelFetch(*s.t, c, 0); }
inline vec4 textureLod(const sampler2D_t& s, const vec2& uv, float) { return texture(s, uv); }
// textureLodOffset: the texel offset is added to the footprint's integer coordinates before clamp-to-edge
inline vec4 textureLodOffset(const sampler2D_t& s, const vec2& uv, float, const ivec2& o) {
	const Image& t = *s.t;
	const float u = uv.x * (float)t.w - 0.5f, v = uv.y * (float)t.h - 0.5f;
	const float fu = std::floor(u), fv = std::floor(v);
	const float a = u - fu, b = v - fv;
	const int i0 = f2i(fu) + o.x, j0 = f2i(fv) + o.y;
	const int x0 = clampi(i0, 0, t.w - 1), x1 = clampi(i0 + 1, 0, t.w - 1), y0 = clampi(j0, 0, t.h - 1), y1 = clampi(j0 + 1, 0, t.h - 1);
	return lerp4(lerp4(texel(t, x0, y0), texel(t, x1, y0), a), lerp4(texel(t, x0, y1), texel(t, x1, y1), a), b);
}
// textureGather(Offset): the linear filter's 2x2 footprint, chosen after snapping the unnormalised coordinate to 1/256 texel
// (subTexelPrecisionBits = 8), returned as (i0,j1), (i1,j1), (i1,j0), (i0,j0)
inline vec4 textureGatherOffset(const sampler2D_t& s, const vec2& uv, const ivec2& o, int comp) {
	const Image& t = *s.t;
	const float u = uv.x * (float)t.w - 0.5f, v = uv.y * (float)t.h - 0.5f;
	const float fu = std::floor(std::floor(u * 256.0f + 0.5f) * (1.0f / 256.0f)), fv = std::floor(std::floor(v * 256.0f + 0.5f) * (1.0f / 256.0f));
	const int i0 = f2i(fu) + o.x, j0 = f2i(fv) + o.y;
	const int x0 = clampi(i0, 0, t.w - 1), x1 = clampi(i0 + 1, 0, t.w - 1), y0 = clampi(j0, 0, t.h - 1), y1 = clampi(j0 + 1, 0, t.h - 1);
	const vec4 a = texel(t, x0, y1), b = texel(t, x1, y1), c = texel(t, x1, y0), d = texel(t, x0, y0);
	return vec4((&a.x)[comp], (&b.x)[comp], (&c.x)[comp], (&d.x)[comp]);
}
inline vec4 textureGather(const sampler2D_t& s, const vec2& uv, int comp) { return textureGatherOffset(s, uv, ivec2(0, 0), comp); }
inline void imageStore(const Image& t, const ivec2& c, const vec4& v) {
	if (!t.data || c.x < 0 || c.y < 0 || c.x >= t.w || c.y >= t.h) return;
	uint16_t* q = (uint16_t*)(t.data + (long long)c.y * t.pitch) + 4 * c.x;
	q[0] = float_to_half(v.x); q[1] = float_to_half(v.y); q[2] = float_to_half(v.z); q[3] = float_to_half(v.w);
}
inline void imageStore(const Image& t, const ivec2& c, const uvec4& v) {
	if (!t.data || c.x < 0 || c.y < 0 || c.x >= t.w || c.y >= t.h) return;
	((uint32_t*)(t.data + (long long)c.y * t.pitch))[c.x] = v.x;
}
inline uvec4 imageLoad(const Image& t, const ivec2& c) {
	if (!t.data || c.x < 0 || c.y < 0 || c.x >= t.w || c.y >= t.h) return uvec4();
	return uvec4(((const uint32_t*)(t.data + (long long)c.y * t.pitch))[c.x], 0, 0, 0);
}

// imageLoad on a float image (sharpen_cas.comp; ref_build.py renames the call): out of range -> 0, like texelFetch
inline vec4 imageLoadF(const Image& t, const ivec2& c) { return texelFetch(t, c, 0); }

static thread_local uvec3 gl_GlobalInvocationID;
static thread_local uvec3 gl_LocalInvocationID;
static thread_local uvec3 gl_WorkGroupID;

}  // namespace glsl
