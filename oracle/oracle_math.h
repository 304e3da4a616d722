/*
 * oracle_math.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU oracle for the Granite clustered-deferred-lighting + HDR-post hot path.
 * This directory is a plain-C restatement of the reference's GLSL / host C++
 * for that path.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may build, load or call it.  The
 * product (granite_b200/) never includes or links anything from here.
 *
 * PARITY STATUS: pinned to the reference itself.  The reference holds no golden
 * vectors or numeric asserts for K1..K13 and its
 * Vulkan path cannot run in the build container, but its own shaders can be
 * executed on the CPU: GLSL (untouched, under /root/reference) -> SPIR-V with
 * the reference's vendored glslang -> C++ with its vendored spirv-cross
 * (`make -C oracle ref-shaders`, ref_{shader,post,light}_shim.cpp).  Every
 * kernel of this oracle is compared with those executables (K1-K4, K7-K13 bit
 * for bit on stored values; K5+K6 <= 1 B10G11R11 code, > 99.9 % identical) in
 * tests/test_oracle_ref_*shaders.py, and against fixtures they wrote
 * (tests/golden/ref*.npz) where /root/reference does not exist.  The host-math
 * helpers (perspective / inverse / look_at / floatToHalf / frustum) are checked
 * bit for bit against the reference's own math/ compiled into oracle/_ref
 * (`make ref`, tests/test_oracle_cpu.py).  DESIGN.md section 2 lists what the
 * shims supply (sampler filtering, storage formats: the things the reference
 * leaves to the Vulkan implementation).
 *
 * Arithmetic contract (SURVEY.md §8c "Oracle definition we adopt"):
 *   - fp32 everywhere ("mediump" is a no-op on desktop GPUs), evaluated
 *     strictly left-to-right with NO fused multiply-add (build with
 *     -ffp-contract=off);
 *   - sqrt and division are IEEE correctly rounded;
 *   - log2/exp2/pow are glibc's float versions;
 *   - GLSL built-ins are expanded as documented on each helper below.
 */
#ifndef ORACLE_MATH_H_
#define ORACLE_MATH_H_

#include <math.h>
#include <stdint.h>
#include <string.h>

typedef struct { float x, y; } vec2;
typedef struct { float x, y, z; } vec3;
typedef struct { float x, y, z, w; } vec4;

static inline float f_min(float a, float b) { return a < b ? a : b; }  /* GLSL min: y < x ? y : x ; symmetric for non-NaN */
static inline float f_max(float a, float b) { return a > b ? a : b; }
static inline float f_clamp(float x, float lo, float hi) { return f_min(f_max(x, lo), hi); }
static inline float f_mix(float a, float b, float t) { return a * (1.0f - t) + b * t; } /* GLSL spec: x*(1-a)+y*a */
static inline float f_sign(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); }
static inline float f_smoothstep(float e0, float e1, float x)
{
	float t = f_clamp((x - e0) / (e1 - e0), 0.0f, 1.0f);
	return t * t * (3.0f - 2.0f * t);
}

static inline uint32_t f_bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float bits_f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

static inline vec3 v3(float x, float y, float z) { vec3 r = { x, y, z }; return r; }
static inline vec3 v3_add(vec3 a, vec3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline vec3 v3_sub(vec3 a, vec3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline vec3 v3_mul(vec3 a, vec3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline vec3 v3_scale(vec3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
static inline vec3 v3_neg(vec3 a) { return v3(-a.x, -a.y, -a.z); }
static inline float v3_dot(vec3 a, vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline float v3_length(vec3 a) { return sqrtf(v3_dot(a, a)); }
/* GLSL normalize(v) := v * (1 / sqrt(dot(v,v))) */
static inline vec3 v3_normalize(vec3 a) { float inv = 1.0f / sqrtf(v3_dot(a, a)); return v3_scale(a, inv); }
static inline vec3 v3_mixf(vec3 a, vec3 b, float t) { return v3(f_mix(a.x, b.x, t), f_mix(a.y, b.y, t), f_mix(a.z, b.z, t)); }
static inline vec3 v3_min(vec3 a, vec3 b) { return v3(f_min(a.x, b.x), f_min(a.y, b.y), f_min(a.z, b.z)); }
static inline vec3 v3_max(vec3 a, vec3 b) { return v3(f_max(a.x, b.x), f_max(a.y, b.y), f_max(a.z, b.z)); }

static inline vec2 v2(float x, float y) { vec2 r = { x, y }; return r; }
static inline float v2_length(vec2 a) { return sqrtf(a.x * a.x + a.y * a.y); }

static inline vec4 v4(float x, float y, float z, float w) { vec4 r = { x, y, z, w }; return r; }
static inline vec4 v4_mixf(vec4 a, vec4 b, float t)
{
	return v4(f_mix(a.x, b.x, t), f_mix(a.y, b.y, t), f_mix(a.z, b.z, t), f_mix(a.w, b.w, t));
}

/* Column-major mat4 (m[col*4+row]), as muglm / GLSL. M*v summed over columns left to right. */
static inline vec4 m4_mul_v4(const float *m, vec4 v)
{
	vec4 r;
	r.x = m[0] * v.x + m[4] * v.y + m[8] * v.z + m[12] * v.w;
	r.y = m[1] * v.x + m[5] * v.y + m[9] * v.z + m[13] * v.w;
	r.z = m[2] * v.x + m[6] * v.y + m[10] * v.z + m[14] * v.w;
	r.w = m[3] * v.x + m[7] * v.y + m[11] * v.z + m[15] * v.w;
	return r;
}

/* ---- storage formats (SURVEY.md §7 "Format conversions", §8c) ---- */

/* IEEE binary16 <-> binary32, round-to-nearest-even (the imageStore rgba16f rule we adopt). */
static inline uint16_t f32_to_f16_rne(float f)
{
	uint32_t x = f_bits(f);
	uint32_t sign = (x >> 16) & 0x8000u;
	uint32_t ax = x & 0x7fffffffu;
	if (ax >= 0x7f800000u) /* inf / nan */
		return (uint16_t)(sign | 0x7c00u | ((ax > 0x7f800000u) ? 0x200u : 0u));
	if (ax >= 0x477ff000u) /* >= 65520 rounds to inf */
		return (uint16_t)(sign | 0x7c00u);
	if (ax < 0x33000001u) /* <= 2^-25: rounds to zero (2^-25 itself ties to even = 0) */
		return (uint16_t)sign;
	int e = (int)(ax >> 23) - 127;
	uint32_t m = (ax & 0x7fffffu) | 0x800000u;
	int shift;
	uint32_t half;
	if (e < -14)
	{
		shift = 13 + (-14 - e); /* subnormal half */
		uint32_t q = m >> shift;
		uint32_t rem = m & ((1u << shift) - 1u);
		uint32_t halfway = 1u << (shift - 1);
		if (rem > halfway || (rem == halfway && (q & 1u)))
			q++;
		return (uint16_t)(sign | q);
	}
	shift = 13;
	half = ((uint32_t)(e + 15) << 10) | ((m >> shift) & 0x3ffu);
	{
		uint32_t rem = m & 0x1fffu;
		if (rem > 0x1000u || (rem == 0x1000u && (half & 1u)))
			half++;
	}
	return (uint16_t)(sign | half);
}

static inline float f16_to_f32(uint16_t h)
{
	uint32_t sign = ((uint32_t)h & 0x8000u) << 16;
	uint32_t e = (h >> 10) & 0x1fu;
	uint32_t m = h & 0x3ffu;
	if (e == 0)
	{
		if (m == 0)
			return bits_f(sign);
		/* subnormal: m * 2^-24 */
		float v = (float)m * 5.9604644775390625e-8f;
		return sign ? -v : v;
	}
	if (e == 31)
		return bits_f(sign | 0x7f800000u | (m << 13));
	return bits_f(sign | ((e + 112u) << 23) | (m << 13));
}

/* Unsigned small floats of B10G11R11_UFLOAT_PACK32: 5-bit exponent (bias 15), MBITS mantissa,
 * no sign.  Conversion rule adopted: negative -> 0, NaN -> NaN, +inf -> inf, finite values
 * truncate toward zero (the D3D11 / Vulkan-permitted rule; never rounds up to inf). */
static inline uint32_t f32_to_ufloat(float f, int mbits)
{
	uint32_t x = f_bits(f);
	uint32_t max_finite = (30u << mbits) | ((1u << mbits) - 1u);
	if ((x & 0x7fffffffu) > 0x7f800000u)
		return (31u << mbits) | 1u; /* NaN */
	if (x & 0x80000000u)
		return 0; /* negative (incl. -inf, -0) */
	if (x == 0x7f800000u)
		return 31u << mbits;
	int e = (int)(x >> 23) - 127;
	uint32_t m = (x & 0x7fffffu) | 0x800000u;
	if (e > 15)
		return max_finite;
	if (e >= -14)
		return ((uint32_t)(e + 15) << mbits) | ((m >> (23 - mbits)) & ((1u << mbits) - 1u));
	/* denormal: value * 2^(14+mbits), truncated */
	int shift = (23 - mbits) + (-14 - e);
	if (shift > 24)
		return 0;
	return m >> shift;
}

static inline float ufloat_to_f32(uint32_t v, int mbits)
{
	uint32_t e = v >> mbits;
	uint32_t m = v & ((1u << mbits) - 1u);
	if (e == 0)
		return (float)m * (mbits == 6 ? 9.5367431640625e-7f /* 2^-20 */ : 1.9073486328125e-6f /* 2^-19 */);
	if (e == 31)
		return bits_f(0x7f800000u | (m << (23 - mbits)));
	return bits_f(((e + 112u) << 23) | (m << (23 - mbits)));
}

static inline uint32_t pack_r11g11b10(vec3 c)
{
	return f32_to_ufloat(c.x, 6) | (f32_to_ufloat(c.y, 6) << 11) | (f32_to_ufloat(c.z, 5) << 22);
}

static inline vec3 unpack_r11g11b10(uint32_t p)
{
	return v3(ufloat_to_f32(p & 0x7ffu, 6), ufloat_to_f32((p >> 11) & 0x7ffu, 6), ufloat_to_f32(p >> 22, 5));
}

/* sRGB8 <-> linear.  Decode is the exact EOTF evaluated in double then rounded to float. */
static inline float srgb8_to_linear(uint32_t v)
{
	double c = (double)v / 255.0;
	double l = c <= 0.04045 ? c / 12.92 : pow((c + 0.055) / 1.055, 2.4);
	return (float)l;
}

/* Attachment store to an *_SRGB 8-bit format: clamp, exact OETF in fp32, round-half-up. */
static inline uint32_t linear_to_srgb8(float c)
{
	if (!(c > 0.0f))
		c = 0.0f; /* also NaN -> 0 */
	if (c > 1.0f)
		c = 1.0f;
	float s = c <= 0.0031308f ? c * 12.92f : 1.055f * powf(c, 1.0f / 2.4f) - 0.055f;
	int q = (int)floorf(s * 255.0f + 0.5f);
	if (q < 0) q = 0;
	if (q > 255) q = 255;
	return (uint32_t)q;
}

static inline uint32_t float_to_unorm8(float c)
{
	if (!(c > 0.0f))
		c = 0.0f;
	if (c > 1.0f)
		c = 1.0f;
	return (uint32_t)floorf(c * 255.0f + 0.5f);
}

/* ---- images ---- */

typedef struct
{
	const uint16_t *data; /* RGBA16F, tightly packed rows */
	int w, h;
} img16f;

static inline vec4 fetch16f(img16f im, int x, int y)
{
	if (x < 0) x = 0;
	if (y < 0) y = 0;
	if (x > im.w - 1) x = im.w - 1;
	if (y > im.h - 1) y = im.h - 1;
	const uint16_t *p = im.data + ((size_t)y * im.w + x) * 4;
	return v4(f16_to_f32(p[0]), f16_to_f32(p[1]), f16_to_f32(p[2]), f16_to_f32(p[3]));
}

/* StockSampler::LinearClamp (vulkan/device.cpp:1077-1170): bilinear, clamp-to-edge, texel
 * centres at +0.5, exact fp32 weights.  u,v are NORMALISED coordinates exactly as the shader
 * computed them; the un-normalisation u*W-0.5 is done here in fp32. */
typedef struct { int x0, x1, y0, y1; float a, b; } bilin_t;

static inline bilin_t bilin_setup(float u, float v, int w, int h)
{
	bilin_t s;
	float fx = u * (float)w - 0.5f;
	float fy = v * (float)h - 0.5f;
	float flx = floorf(fx), fly = floorf(fy);
	s.a = fx - flx;
	s.b = fy - fly;
	/* clamp in float first so huge / NaN coordinates do not overflow the int conversion */
	flx = f_clamp(flx, -2.0f, (float)w + 1.0f);
	fly = f_clamp(fly, -2.0f, (float)h + 1.0f);
	if (!(flx == flx)) flx = 0.0f;
	if (!(fly == fly)) fly = 0.0f;
	s.x0 = (int)flx; s.y0 = (int)fly;
	s.x1 = s.x0 + 1; s.y1 = s.y0 + 1;
	return s;
}

static inline float bilin_mix(float t00, float t10, float t01, float t11, float a, float b)
{
	float top = t00 * (1.0f - a) + t10 * a;
	float bot = t01 * (1.0f - a) + t11 * a;
	return top * (1.0f - b) + bot * b;
}

/* A bilinear weight within 2^-9 of 0 or 1 becomes exactly 0 or 1.  Used ONLY by the history taps of
 * sample_catmull_rom (reprojection.h:286-334): the shader aims taps t0 and t3 at texel centres through
 * normalised coordinates, whose fp32 rounding is ~1e-4 texel at 4K; a GPU's sampler converts the
 * position to fixed point with 8 fractional bits and therefore fetches the texel, whereas an exact-weight
 * evaluation would blend in 1e-4 of a neighbour (decision table). */
static inline float snap_weight(float f) { return f <= 0.001953125f ? 0.0f : (f >= 1.0f - 0.001953125f ? 1.0f : f); }

static inline vec4 sample16f_linear_snap(img16f im, float u, float v)
{
	bilin_t s = bilin_setup(u, v, im.w, im.h);
	s.a = snap_weight(s.a);
	s.b = snap_weight(s.b);
	vec4 t00 = fetch16f(im, s.x0, s.y0), t10 = fetch16f(im, s.x1, s.y0);
	vec4 t01 = fetch16f(im, s.x0, s.y1), t11 = fetch16f(im, s.x1, s.y1);
	return v4(bilin_mix(t00.x, t10.x, t01.x, t11.x, s.a, s.b),
	          bilin_mix(t00.y, t10.y, t01.y, t11.y, s.a, s.b),
	          bilin_mix(t00.z, t10.z, t01.z, t11.z, s.a, s.b),
	          bilin_mix(t00.w, t10.w, t01.w, t11.w, s.a, s.b));
}

static inline vec4 sample16f_linear(img16f im, float u, float v)
{
	bilin_t s = bilin_setup(u, v, im.w, im.h);
	vec4 t00 = fetch16f(im, s.x0, s.y0), t10 = fetch16f(im, s.x1, s.y0);
	vec4 t01 = fetch16f(im, s.x0, s.y1), t11 = fetch16f(im, s.x1, s.y1);
	return v4(bilin_mix(t00.x, t10.x, t01.x, t11.x, s.a, s.b),
	          bilin_mix(t00.y, t10.y, t01.y, t11.y, s.a, s.b),
	          bilin_mix(t00.z, t10.z, t01.z, t11.z, s.a, s.b),
	          bilin_mix(t00.w, t10.w, t01.w, t11.w, s.a, s.b));
}

static inline void store16f(uint16_t *dst, int w, int x, int y, vec4 v)
{
	uint16_t *p = dst + ((size_t)y * w + x) * 4;
	p[0] = f32_to_f16_rne(v.x);
	p[1] = f32_to_f16_rne(v.y);
	p[2] = f32_to_f16_rne(v.z);
	p[3] = f32_to_f16_rne(v.w);
}

#endif
