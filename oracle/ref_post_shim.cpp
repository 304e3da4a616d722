// oracle/ref_post_shim.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Host driver for ONE of the reference's own post-processing shaders, executed on the CPU:
//   GLSL (/root/reference/assets/shaders/post/*.{comp,frag}, untouched)
//     -> SPIR-V by the reference's vendored glslang -> C++ by its vendored spirv-cross (`--cpp`)
//     -> this translation unit #includes that generated C++ (GEN_CPP).
// The arithmetic that runs is the reference shader's, statement for statement, on GLM vector types
// with plain IEEE fp32 (-ffp-contract=off).  `oracle/Makefile ref-shaders` is the recipe; outputs
// live in oracle/_ref/ (git-ignored); nothing of the reference is copied into the repository.
//
//   KERNEL=7   post/bloom_threshold.comp (DYNAMIC_EXPOSURE=1)   -> refk7_bloom_threshold
//   KERNEL=8   post/bloom_downsample.comp (FEEDBACK=0)          -> refk8_bloom_downsample
//   KERNEL=18  post/bloom_downsample.comp (FEEDBACK=1)          -> refk8_bloom_downsample_feedback
//   KERNEL=9   post/bloom_upsample.comp                         -> refk9_bloom_upsample
//   KERNEL=10  post/luminance.comp                              -> refk10_luminance
//   KERNEL=11  post/tonemap.frag (DYNAMIC_EXPOSURE=1)           -> refk11_tonemap
//   KERNEL=12  post/fxaa.frag (FXAA_TARGET_SRGB=0/1 -> KERNEL 12 / 22) -> refk12_fxaa
//   KERNEL=13  post/taa_resolve.frag (TAA_QUALITY=2, REPROJECTION_HISTORY=1; 23: quality 0, 33: quality 1,
//              43: no history)                                   -> refk13_taa_resolve
//   KERNEL=14  post/pq10_encode.frag (HDR10 / ST.2084 output encoding) -> refk14_pq10_encode
//
// What the shim supplies -- and the reference leaves to the Vulkan implementation -- is exactly
// the list DESIGN.md section 2 fixes once for oracle and product alike:
//   * StockSampler::LinearClamp / NearestClamp (vulkan/device.cpp:1077-1170): bilinear with exact
//     fp32 weights, texel centres at +0.5, clamp-to-edge; textureGather per the Vulkan spec
//     (texel order (i0,j1) (i1,j1) (i1,j0) (i0,j0), footprint of the bilinear sample);
//   * storage formats: RGBA16F stores round to nearest even, B10G11R11 / sRGB8 / UNORM8 through
//     the oracle's format helpers (liboracle.so: orc_pack_r11g11b10 ...), which are themselves
//     tested against numpy;
//   * GLSL built-ins the deprecated C++ backend's runtime lacks (mix in the specification's form).
// The C++ backend's own sampler.hpp / image.hpp are skeletons (linear filtering returns a constant),
// so their include guards are claimed here and replaced.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#define GLM_FORCE_PURE
#ifndef GLM_SWIZZLE
#define GLM_SWIZZLE
#endif
#ifndef GLM_FORCE_RADIANS
#define GLM_FORCE_RADIANS
#endif
#include <glm/glm.hpp>

extern "C" {
uint32_t orc_pack_r11g11b10(float r, float g, float b);
void orc_unpack_r11g11b10(uint32_t p, float *rgb);
uint16_t orc_f32_to_f16(float f);
float orc_f16_to_f32(uint16_t h);
uint32_t orc_linear_to_srgb8(float c);
float orc_srgb8_to_linear(uint32_t v);
}

#define SPIRV_CROSS_SAMPLER_HPP
#define SPIRV_CROSS_IMAGE_HPP
namespace spirv_cross
{
enum ShimFormat
{
	FMT_R11G11B10 = 0,
	FMT_RGBA16F = 1,
	FMT_RGBA8_UNORM = 2,
	FMT_RGBA8_SRGB = 3,
	FMT_D32F = 4,
	FMT_RG16F = 5,
	FMT_RG8_UNORM = 6,
	FMT_R8_UNORM = 7,
};

// A texture view combined with a StockSampler.
struct sampler2D
{
	const void *data = nullptr;
	int w = 0, h = 0, format = 0;
	bool linear = true;
	// Same-size sampling at the fragment's own centre (tonemap uHDR, FXAA centre / integer-offset taps):
	// a texel fetch on hardware, whose bilinear weights have 8 fractional bits; fp32 rounding of
	// (x + 0.5) / W * W - 0.5 would leak ~1e-7 of a neighbour instead.  DESIGN.md section 2: "same-size
	// sampling at a pixel centre = exact fetch".  Set by the rasteriser loop; null = never.
	const glm::vec2 *frag_uv = nullptr;
	const glm::ivec2 *frag_px = nullptr;
	// The TAA history sampler only: a bilinear weight within 2^-9 of 0 or 1 is exactly 0 or 1 (oracle_math.h
	// snap_weight: Catmull-Rom taps aimed at texel centres through normalised coordinates).
	bool snap_centres = false;

	glm::vec4 texel(int x, int y) const
	{
		x = x < 0 ? 0 : (x > w - 1 ? w - 1 : x);
		y = y < 0 ? 0 : (y > h - 1 ? h - 1 : y);
		const size_t i = (size_t)y * w + x;
		switch (format)
		{
		case FMT_R11G11B10:
		{
			float c[3];
			orc_unpack_r11g11b10(static_cast<const uint32_t *>(data)[i], c);
			return glm::vec4(c[0], c[1], c[2], 1.0f);
		}
		case FMT_RGBA16F:
		{
			const uint16_t *p = static_cast<const uint16_t *>(data) + 4 * i;
			return glm::vec4(orc_f16_to_f32(p[0]), orc_f16_to_f32(p[1]), orc_f16_to_f32(p[2]), orc_f16_to_f32(p[3]));
		}
		case FMT_RGBA8_UNORM:
		{
			const uint32_t v = static_cast<const uint32_t *>(data)[i];
			return glm::vec4((float)(v & 255u) / 255.0f, (float)((v >> 8) & 255u) / 255.0f, (float)((v >> 16) & 255u) / 255.0f, (float)(v >> 24) / 255.0f);
		}
		case FMT_RGBA8_SRGB:
		{
			const uint32_t v = static_cast<const uint32_t *>(data)[i];
			return glm::vec4(orc_srgb8_to_linear(v & 255u), orc_srgb8_to_linear((v >> 8) & 255u), orc_srgb8_to_linear((v >> 16) & 255u), (float)(v >> 24) / 255.0f);
		}
		case FMT_D32F:
			return glm::vec4(static_cast<const float *>(data)[i], 0.0f, 0.0f, 1.0f);
		case FMT_RG16F:
		{
			const uint16_t *p = static_cast<const uint16_t *>(data) + 2 * i;
			return glm::vec4(orc_f16_to_f32(p[0]), orc_f16_to_f32(p[1]), 0.0f, 1.0f);
		}
		case FMT_RG8_UNORM:
		{
			const uint8_t *p = static_cast<const uint8_t *>(data) + 2 * i;
			return glm::vec4((float)p[0] / 255.0f, (float)p[1] / 255.0f, 0.0f, 1.0f);
		}
		case FMT_R8_UNORM:
			return glm::vec4((float)static_cast<const uint8_t *>(data)[i] / 255.0f, 0.0f, 0.0f, 1.0f);
		}
		return glm::vec4(0.0f);
	}

	struct Foot
	{
		int x0, y0;
		float a, b;
	};
	Foot footprint(glm::vec2 uv) const
	{
		const float fx = uv.x * (float)w - 0.5f, fy = uv.y * (float)h - 0.5f;
		float flx = std::floor(fx), fly = std::floor(fy);
		Foot f;
		f.a = fx - flx;
		f.b = fy - fly;
		flx = std::fmin(std::fmax(flx, -2.0f), (float)w + 1.0f);
		fly = std::fmin(std::fmax(fly, -2.0f), (float)h + 1.0f);
		if (!(flx == flx)) flx = 0.0f;
		if (!(fly == fly)) fly = 0.0f;
		f.x0 = (int)flx;
		f.y0 = (int)fly;
		return f;
	}

	glm::vec4 sample(glm::vec2 uv, glm::ivec2 off) const
	{
		if (!linear)
		{
			const int x = (int)std::floor(uv.x * (float)w), y = (int)std::floor(uv.y * (float)h);
			return texel(x + off.x, y + off.y);
		}
		if (frag_uv && uv.x == frag_uv->x && uv.y == frag_uv->y)
			return texel(frag_px->x + off.x, frag_px->y + off.y);
		Foot f = footprint(uv);
		if (snap_centres)
		{
			f.a = f.a <= 0.001953125f ? 0.0f : (f.a >= 1.0f - 0.001953125f ? 1.0f : f.a);
			f.b = f.b <= 0.001953125f ? 0.0f : (f.b >= 1.0f - 0.001953125f ? 1.0f : f.b);
		}
		const glm::vec4 t00 = texel(f.x0 + off.x, f.y0 + off.y), t10 = texel(f.x0 + 1 + off.x, f.y0 + off.y);
		const glm::vec4 t01 = texel(f.x0 + off.x, f.y0 + 1 + off.y), t11 = texel(f.x0 + 1 + off.x, f.y0 + 1 + off.y);
		const float ia = 1.0f - f.a, ib = 1.0f - f.b;
		const glm::vec4 top = t00 * ia + t10 * f.a;
		const glm::vec4 bot = t01 * ia + t11 * f.a;
		return top * ib + bot * f.b;
	}
};

inline glm::vec4 textureLod(const sampler2D &s, const glm::vec2 &uv, float) { return s.sample(uv, glm::ivec2(0)); }
inline glm::vec4 texture(const sampler2D &s, const glm::vec2 &uv) { return s.sample(uv, glm::ivec2(0)); }
inline glm::vec4 textureLodOffset(const sampler2D &s, const glm::vec2 &uv, float, const glm::ivec2 &off) { return s.sample(uv, off); }
inline glm::vec4 texelFetch(const sampler2D &s, const glm::ivec2 &p, int) { return s.texel(p.x, p.y); }
inline glm::vec4 textureGatherOffset(const sampler2D &s, const glm::vec2 &uv, const glm::ivec2 &off, int comp = 0)
{
	const sampler2D::Foot f = s.footprint(uv);
	const int x0 = f.x0 + off.x, y0 = f.y0 + off.y;
	return glm::vec4(s.texel(x0, y0 + 1)[comp], s.texel(x0 + 1, y0 + 1)[comp], s.texel(x0 + 1, y0)[comp], s.texel(x0, y0)[comp]);
}
inline glm::vec4 textureGather(const sampler2D &s, const glm::vec2 &uv, int comp = 0) { return textureGatherOffset(s, uv, glm::ivec2(0), comp); }
inline glm::ivec2 textureSize(const sampler2D &s, int) { return glm::ivec2(s.w, s.h); }
typedef sampler2D texture2D; // separate images are only ever texelFetch'ed

// RGBA16F volumes (fog_accumulate.comp): a NearestClamp sampler3D -- nearest texel of the normalised coordinate, integer
// offset, clamp to edge -- and a storage image3D
struct sampler3D
{
	const uint16_t *data = nullptr;
	int w = 0, h = 0, d = 0;
};
inline glm::vec4 textureLodOffset(const sampler3D &s, const glm::vec3 &uvw, float, const glm::ivec3 &off)
{
	int x = (int)std::floor(uvw.x * (float)s.w) + off.x, y = (int)std::floor(uvw.y * (float)s.h) + off.y, z = (int)std::floor(uvw.z * (float)s.d) + off.z;
	x = x < 0 ? 0 : (x > s.w - 1 ? s.w - 1 : x);
	y = y < 0 ? 0 : (y > s.h - 1 ? s.h - 1 : y);
	z = z < 0 ? 0 : (z > s.d - 1 ? s.d - 1 : z);
	const uint16_t *t = s.data + 4 * (((size_t)z * s.h + y) * s.w + x);
	return glm::vec4(orc_f16_to_f32(t[0]), orc_f16_to_f32(t[1]), orc_f16_to_f32(t[2]), orc_f16_to_f32(t[3]));
}
struct image3D
{
	uint16_t *data = nullptr;
	int w = 0, h = 0, d = 0;
};
inline void imageStore(image3D &im, const glm::ivec3 &p, const glm::vec4 &v)
{
	if (p.x < 0 || p.y < 0 || p.z < 0 || p.x >= im.w || p.y >= im.h || p.z >= im.d)
		return;
	uint16_t *o = im.data + 4 * (((size_t)p.z * im.h + p.y) * im.w + p.x);
	o[0] = orc_f32_to_f16(v.x);
	o[1] = orc_f32_to_f16(v.y);
	o[2] = orc_f32_to_f16(v.z);
	o[3] = orc_f32_to_f16(v.w);
}

// rgba16f storage image
struct image2D
{
	uint16_t *data = nullptr;
	int w = 0, h = 0;
};
inline void imageStore(image2D &im, const glm::ivec2 &p, const glm::vec4 &v)
{
	if (p.x < 0 || p.y < 0 || p.x >= im.w || p.y >= im.h)
		return;
	uint16_t *d = im.data + 4 * ((size_t)p.y * im.w + p.x);
	d[0] = orc_f32_to_f16(v.x);
	d[1] = orc_f32_to_f16(v.y);
	d[2] = orc_f32_to_f16(v.z);
	d[3] = orc_f32_to_f16(v.w);
}
} // namespace spirv_cross

#include "spirv_cross/internal_interface.hpp"

// GLSL mix(): the Vulkan specification's form x * (1 - a) + y * a (see ref_shader_shim.cpp).
inline float mix(const float &x, const float &y, const float &a) { return x * (1.0f - a) + y * a; }
inline glm::vec2 mix(const glm::vec2 &x, const glm::vec2 &y, const glm::vec2 &a) { return x * (glm::vec2(1.0f) - a) + y * a; }
inline glm::vec3 mix(const glm::vec3 &x, const glm::vec3 &y, const glm::vec3 &a) { return x * (glm::vec3(1.0f) - a) + y * a; }
inline glm::vec4 mix(const glm::vec4 &x, const glm::vec4 &y, const glm::vec4 &a) { return x * (glm::vec4(1.0f) - a) + y * a; }
inline glm::vec2 mix(const glm::vec2 &x, const glm::vec2 &y, const float &a) { return x * (1.0f - a) + y * a; }
inline glm::vec3 mix(const glm::vec3 &x, const glm::vec3 &y, const float &a) { return x * (1.0f - a) + y * a; }
inline glm::vec4 mix(const glm::vec4 &x, const glm::vec4 &y, const float &a) { return x * (1.0f - a) + y * a; }
// mix(x, y, bvec) = component select
inline glm::vec3 mix(const glm::vec3 &x, const glm::vec3 &y, const glm::bvec3 &a) { return glm::vec3(a.x ? y.x : x.x, a.y ? y.y : x.y, a.z ? y.z : x.z); }
inline float mix(const float &x, const float &y, const bool &a) { return a ? y : x; }

// mat4 * vec4.  GLM adds the four column products pairwise, (c0 x + c1 y) + (c2 z + c3 w), which is a
// property of GLM, not of the reference; OpMatrixTimesVector is lowered by GPU compilers to the
// linear chain c0 x + c1 y + c2 z + c3 w.  Push-constant matrices of the generated code use this type.
struct ShimMat4
{
	glm::vec4 c[4];
	const glm::vec4 &operator[](int i) const { return c[i]; }
};
inline glm::vec4 operator*(const ShimMat4 &m, const glm::vec4 &v) { return ((m.c[0] * v.x + m.c[1] * v.y) + m.c[2] * v.z) + m.c[3] * v.w; }
#define mat4 ShimMat4

#if KERNEL >= 150
// SMAA (SMAA.hlsl with SMAA_GLSL_4): mad(a, b, c) is GLSL fma().  GLM's fma is a * b + c with two roundings; a GPU
// executes it fused, and so do oracle and product.  Non-template overloads win over GLM's.
inline float fma(const float &a, const float &b, const float &c) { return std::fma(a, b, c); }
inline glm::vec2 fma(const glm::vec2 &a, const glm::vec2 &b, const glm::vec2 &c) { return glm::vec2(std::fma(a.x, b.x, c.x), std::fma(a.y, b.y, c.y)); }
inline glm::vec3 fma(const glm::vec3 &a, const glm::vec3 &b, const glm::vec3 &c)
{
	return glm::vec3(std::fma(a.x, b.x, c.x), std::fma(a.y, b.y, c.y), std::fma(a.z, b.z, c.z));
}
inline glm::vec4 fma(const glm::vec4 &a, const glm::vec4 &b, const glm::vec4 &c)
{
	return glm::vec4(std::fma(a.x, b.x, c.x), std::fma(a.y, b.y, c.y), std::fma(a.z, b.z, c.z), std::fma(a.w, b.w, c.w));
}
// the edge-detection shader discards where it finds no edge: the attachment keeps its clear colour, 0 (smaa.cpp:131-135)
#define discard return vec2(0.0f)
#endif

#if KERNEL == 24 || KERNEL == 25 || KERNEL == 26
// FSR 1: scalar min / max as a GPU executes them (FMNMX / v_max_f32 return the non-NaN operand, IEEE minNum / maxNum);
// GLM's (x < y) ? y : x would keep a NaN first operand.  RCAS meets one by design: a channel that is 0 over the whole
// ring gives hitMin = 0 * (1 / 0), and max(-hitMin, hitMax) must fall through to hitMax (ffx_fsr1.h:753-755).
inline float max(const float &a, const float &b) { return std::fmax(a, b); }
inline float min(const float &a, const float &b) { return std::fmin(a, b); }
#endif

#include GEN_CPP
#undef mat4
#ifdef discard
#undef discard
#endif

namespace
{
using Sh = Impl::Shader;
using spirv_cross::image2D;
using spirv_cross::sampler2D;

struct Runner
{
	spirv_cross_shader_t *sh;
	const spirv_cross_interface *itf;
	Runner() : sh(nullptr), itf(spirv_cross_get_interface()) { sh = itf->construct(); }
	~Runner() { itf->destruct(sh); }
	void resource(unsigned set, unsigned binding, void *ptr)
	{
		void *p = ptr;
		spirv_cross_set_resource(sh, set, binding, &p, sizeof(p));
	}
	void push(void *data, size_t size) { spirv_cross_set_push_constant(sh, data, size); }
	void dispatch(unsigned gx, unsigned gy)
	{
		glm::uvec3 num(gx, gy, 1), id(0);
		spirv_cross_set_builtin(sh, SPIRV_CROSS_BUILTIN_NUM_WORK_GROUPS, &num, sizeof(num));
		spirv_cross_set_builtin(sh, SPIRV_CROSS_BUILTIN_WORK_GROUP_ID, &id, sizeof(id));
		for (unsigned y = 0; y < gy; y++)
			for (unsigned x = 0; x < gx; x++)
			{
				id = glm::uvec3(x, y, 0);
				itf->invoke(sh);
			}
	}
	// Full-screen triangle (vulkan/command_buffer.cpp:4605-4616, quad.vert:7-12): vUV = (x + 0.5) / W.
	// `emit(x, y)` is called after every invocation.
	glm::ivec2 pixel = glm::ivec2(0);
	template <typename F>
	void raster(int w, int h, int y0, int y1, glm::vec2 *uv_slot, F &&emit, glm::vec2 *pixel_uv_slot = nullptr)
	{
		glm::vec4 frag(0.0f);
		spirv_cross_set_builtin(sh, SPIRV_CROSS_BUILTIN_FRAG_COORD, &frag, sizeof(frag));
		const float inv_w = 1.0f / (float)w, inv_h = 1.0f / (float)h;
		for (int y = y0; y < y1; y++)
			for (int x = 0; x < w; x++)
			{
				frag = glm::vec4((float)x + 0.5f, (float)y + 0.5f, 0.0f, 1.0f);
				pixel = glm::ivec2(x, y);
				if (uv_slot)
					*uv_slot = glm::vec2(((float)x + 0.5f) * inv_w, ((float)y + 0.5f) * inv_h);
				if (pixel_uv_slot) // ffx-fsr/{upscale,sharpen}.vert: vUV = (0.5 * Position + 0.5) * out_resolution
					*pixel_uv_slot = glm::vec2((float)x + 0.5f, (float)y + 0.5f);
				itf->invoke(sh);
				emit(x, y);
			}
	}
};

// "renderTargetFp16": the HDR image K7 / K11 / K13 sample is R16G16B16A16_SFLOAT instead of B10G11R11 (set by the tests
// around a call; the pointer the entry points declare as uint32_t then addresses 8-byte texels)
static int g_hdr_fp16 = 0;

sampler2D make_sampler(const void *data, int w, int h, int format, bool linear = true)
{
	if (format == spirv_cross::FMT_R11G11B10 && g_hdr_fp16)
		format = spirv_cross::FMT_RGBA16F;
	sampler2D s;
	s.data = data;
	s.w = w;
	s.h = h;
	s.format = format;
	s.linear = linear;
	return s;
}
} // namespace

extern "C" {
void refk_set_hdr_fp16(int enable) { g_hdr_fp16 = enable; }
#if KERNEL == 7
// hdr.cpp:115-144: dispatch ceil(w/8) x ceil(h/8), inv_output_size = 1 / out size
void refk7_bloom_threshold(const uint32_t *hdr, int w_in, int h_in, const float *lum3, uint16_t *out, int w, int h)
{
	sampler2D s = make_sampler(hdr, w_in, h_in, spirv_cross::FMT_R11G11B10);
	image2D o;
	o.data = out;
	o.w = w;
	o.h = h;
	Sh::Resources::LuminanceData lum = { lum3[0], lum3[1], lum3[2] };
	Sh::Resources::Registers reg;
	reg.num_threads = glm::uvec2(w, h);
	reg.inv_output_size = glm::vec2(1.0f / (float)w, 1.0f / (float)h);
	Runner r;
	r.resource(0, 0, &s);
	r.resource(0, 1, &lum);
	r.resource(0, 2, &o);
	r.push(&reg, sizeof(reg));
	r.dispatch((unsigned)(w + 7) / 8, (unsigned)(h + 7) / 8);
}
#elif KERNEL == 8 || KERNEL == 18
// hdr.cpp:146-187
void
#if KERNEL == 8
refk8_bloom_downsample
#else
refk8_bloom_downsample_feedback
#endif
(const uint16_t *in, int w_in, int h_in, const uint16_t *history, float lerp, uint16_t *out, int w, int h)
{
	sampler2D s = make_sampler(in, w_in, h_in, spirv_cross::FMT_RGBA16F);
	image2D o;
	o.data = out;
	o.w = w;
	o.h = h;
	Sh::Resources::Registers reg;
	reg.num_threads = glm::uvec2(w, h);
	reg.inv_output_size = glm::vec2(1.0f / (float)w, 1.0f / (float)h);
	reg.inv_input_size = glm::vec2(1.0f / (float)w_in, 1.0f / (float)h_in);
	Runner r;
	r.resource(0, 0, &s);
	r.resource(0, 1, &o);
#if KERNEL == 18
	sampler2D hs = make_sampler(history, w, h, spirv_cross::FMT_RGBA16F, false); // NearestClamp (hdr.cpp:166-167)
	r.resource(0, 2, &hs);
	reg.lerp = lerp;
#else
	(void)history;
	(void)lerp;
#endif
	r.push(&reg, sizeof(reg));
	r.dispatch((unsigned)(w + 7) / 8, (unsigned)(h + 7) / 8);
}
#elif KERNEL == 9
// hdr.cpp:189-216
void refk9_bloom_upsample(const uint16_t *in, int w_in, int h_in, uint16_t *out, int w, int h)
{
	sampler2D s = make_sampler(in, w_in, h_in, spirv_cross::FMT_RGBA16F);
	image2D o;
	o.data = out;
	o.w = w;
	o.h = h;
	Sh::Resources::Registers reg;
	reg.num_threads = glm::uvec2(w, h);
	reg.inv_output_size = glm::vec2(1.0f / (float)w, 1.0f / (float)h);
	reg.inv_input_size = glm::vec2(1.0f / (float)w_in, 1.0f / (float)h_in);
	Runner r;
	r.resource(0, 0, &s);
	r.resource(0, 1, &o);
	r.push(&reg, sizeof(reg));
	r.dispatch((unsigned)(w + 7) / 8, (unsigned)(h + 7) / 8);
}
#elif KERNEL == 10
// hdr.cpp:68-98: one workgroup; size = d3 / 2 (integer), lerp = 1 - 0.5^frame_time, clamp [min, max]
void refk10_luminance(const uint16_t *d3, int w, int h, float lerp, float min_loglum, float max_loglum, float *lum3)
{
	sampler2D s = make_sampler(d3, w, h, spirv_cross::FMT_RGBA16F);
	Sh::Resources::Registers reg;
	reg.size = glm::ivec2(w / 2, h / 2);
	reg.lerp = lerp;
	reg.min_loglum = min_loglum;
	reg.max_loglum = max_loglum;
	Runner r;
	r.resource(0, 0, lum3);
	r.resource(0, 1, &s);
	r.push(&reg, sizeof(reg));
	r.dispatch(1, 1);
}
#elif KERNEL == 11
// hdr.cpp:283-306; attachment R8G8B8A8_SRGB: the store applies the OETF (oracle helper)
void refk11_tonemap(const uint32_t *hdr, int w, int h, const uint16_t *bloom, int bw, int bh, const float *lum3, float exposure, uint32_t *out,
                    int y0, int y1)
{
	sampler2D s_hdr = make_sampler(hdr, w, h, spirv_cross::FMT_R11G11B10);
	sampler2D s_bloom = make_sampler(bloom, bw, bh, spirv_cross::FMT_RGBA16F);
	Sh::Resources::LuminanceData lum = { lum3[0], lum3[1], lum3[2] };
	Sh::Resources::Registers reg;
	reg.dynamic_exposure = exposure;
	glm::vec2 uv(0.0f);
	glm::vec3 color(0.0f);
	Runner r;
	s_hdr.frag_uv = &uv;
	s_hdr.frag_px = &r.pixel;
	r.resource(0, 0, &s_hdr);
	r.resource(0, 1, &s_bloom);
	r.resource(0, 2, &lum);
	r.push(&reg, sizeof(reg));
	spirv_cross_set_stage_input(r.sh, 0, &uv, sizeof(uv));
	spirv_cross_set_stage_output(r.sh, 0, &color, sizeof(color));
	r.raster(w, h, y0, y1, &uv, [&](int x, int y) {
		out[(size_t)y * w + x] = orc_linear_to_srgb8(color.x) | (orc_linear_to_srgb8(color.y) << 8) | (orc_linear_to_srgb8(color.z) << 16) | 0xff000000u;
	});
}
#elif KERNEL == 12 || KERNEL == 22
// fxaa.cpp:28-56: the input is read through a UNORM view; KERNEL 22 = sRGB target (FXAA_TARGET_SRGB=1)
void
#if KERNEL == 12
refk12_fxaa_unorm
#else
refk12_fxaa_srgb
#endif
(const uint32_t *in, int w, int h, uint32_t *out, int y0, int y1)
{
	sampler2D s = make_sampler(in, w, h, spirv_cross::FMT_RGBA8_UNORM);
	Sh::Resources::Registers reg;
	reg.inv_resolution = glm::vec2(1.0f / (float)w, 1.0f / (float)h); // fxaa.cpp:45-46
	glm::vec2 uv(0.0f);
	glm::vec3 color(0.0f);
	Runner r;
	s.frag_uv = &uv;
	s.frag_px = &r.pixel;
	r.resource(0, 0, &s);
	r.push(&reg, sizeof(reg));
	spirv_cross_set_stage_input(r.sh, 0, &uv, sizeof(uv));
	spirv_cross_set_stage_output(r.sh, 0, &color, sizeof(color));
	r.raster(w, h, y0, y1, &uv, [&](int x, int y) {
		auto q = [](float c) -> uint32_t {
#if KERNEL == 22
			return orc_linear_to_srgb8(c);
#else
			c = c > 0.0f ? (c < 1.0f ? c : 1.0f) : 0.0f;
			return (uint32_t)std::floor(c * 255.0f + 0.5f);
#endif
		};
		out[(size_t)y * w + x] = q(color.x) | (q(color.y) << 8) | (q(color.z) << 16) | 0xff000000u;
	});
}
#elif KERNEL == 13 || KERNEL == 23 || KERNEL == 33 || KERNEL == 43
// temporal.cpp:199-266
void
#if KERNEL == 13
refk13_taa_q2
#elif KERNEL == 23
refk13_taa_q0
#elif KERNEL == 33
refk13_taa_q1
#else
refk13_taa_nohistory
#endif
(const uint32_t *hdr, const float *depth, const uint16_t *mv, const uint16_t *history, int w, int h, const float *reproj16, uint32_t *out_color,
 uint16_t *out_history, int y0, int y1)
{
	sampler2D s_cur = make_sampler(hdr, w, h, spirv_cross::FMT_R11G11B10, false);
	Runner r;
	Sh::Resources::Registers reg;
	std::memset(&reg, 0, sizeof(reg));
#if KERNEL != 43
	sampler2D s_depth = make_sampler(depth, w, h, spirv_cross::FMT_D32F, false);
	sampler2D s_mv = make_sampler(mv, w, h, spirv_cross::FMT_RG16F, false);
	sampler2D s_hist = make_sampler(history, w, h, spirv_cross::FMT_RGBA16F, true);
#if KERNEL == 13
	s_hist.snap_centres = true; // quality 2 = sample_catmull_rom
#endif
	r.resource(0, 1, &s_depth);
	r.resource(0, 2, &s_mv);
	r.resource(0, 3, &s_hist);
	std::memcpy(&reg.reproj, reproj16, 64);
	reg.rt_metrics = glm::vec4(1.0f / (float)w, 1.0f / (float)h, (float)w, (float)h);
#else
	(void)depth; (void)mv; (void)history; (void)reproj16;
#endif
	r.resource(0, 0, &s_cur);
	glm::vec2 uv(0.0f);
	glm::vec3 color(0.0f);
	glm::vec3 hist(0.0f);
	r.push(&reg, sizeof(reg));
	spirv_cross_set_stage_input(r.sh, 0, &uv, sizeof(uv));
	spirv_cross_set_stage_output(r.sh, 0, &color, sizeof(color));
	spirv_cross_set_stage_output(r.sh, 1, &hist, sizeof(hist));
	r.raster(w, h, y0, y1, &uv, [&](int x, int y) {
		const size_t i = (size_t)y * w + x;
		out_color[i] = orc_pack_r11g11b10(color.x, color.y, color.z);
		out_history[4 * i + 0] = orc_f32_to_f16(hist.x);
		out_history[4 * i + 1] = orc_f32_to_f16(hist.y);
		out_history[4 * i + 2] = orc_f32_to_f16(hist.z);
		out_history[4 * i + 3] = orc_f32_to_f16(1.0f); // a vec3 output to an RGBA16F attachment: alpha = 1
	});
}
#elif KERNEL == 14
// hdr.cpp:619-642: texelFetch of the scene colour and the UI layer, Config UBO at (1, 0); the attachment is
// A2B10G10R10_UNORM (HDR10 swapchain): round to nearest, NaN -> 0
void refk14_pq10_encode(const uint32_t *hdr, const uint32_t *ui, int w, int h, const float *primary16, float hdr_pre_exposure, float ui_pre_exposure,
                        float max_light_level, uint32_t *out, int y0, int y1)
{
	sampler2D s_hdr = make_sampler(hdr, w, h, spirv_cross::FMT_R11G11B10, false);
	sampler2D s_ui = make_sampler(ui, w, h, spirv_cross::FMT_RGBA8_UNORM, false);
	Sh::Resources::Config cfg;
	std::memcpy(&cfg.primary_conversion, primary16, 64);
	cfg.hdr_pre_exposure = hdr_pre_exposure;
	cfg.ui_pre_exposure = ui_pre_exposure;
	cfg.max_light_level = max_light_level;
	cfg.inv_max_light_level = 1.0f / max_light_level; // hdr.cpp:637
	glm::vec4 color(0.0f);
	Runner r;
	r.resource(0, 0, &s_hdr);
	r.resource(0, 1, &s_ui);
	r.resource(1, 0, &cfg);
	spirv_cross_set_stage_output(r.sh, 0, &color, sizeof(color));
	r.raster(w, h, y0, y1, nullptr, [&](int x, int y) {
		auto q = [](float c) -> uint32_t {
			if (!(c > 0.0f)) c = 0.0f;
			if (c > 1.0f) c = 1.0f;
			return (uint32_t)std::floor(c * 1023.0f + 0.5f);
		};
		out[(size_t)y * w + x] = q(color.x) | (q(color.y) << 10) | (q(color.z) << 20) | (3u << 30);
	});
}
#elif KERNEL >= 150 && KERNEL < 180
// SMAA (renderer/post/smaa.cpp:32-209).  KERNEL = 150 + q: smaa_edge_detection.frag, 160 + q: smaa_blend_weight.frag
// (SMAA_SUBPIXEL_MODE = 0), 170 + q: smaa_neighbor_blend.frag (SMAA_TARGET_SRGB = 1), q = SMAA_QUALITY 0..3.
// The vertex stage (smaa_*.vert) only adds constant multiples of rt_metrics to the texture coordinate; the
// interpolated values are those functions of the fragment's own coordinate, evaluated here with the same fma.
static inline glm::vec4 mad4(const glm::vec4 &a, const glm::vec4 &b, const glm::vec4 &c) { return fma(a, b, c); }
#if KERNEL < 160
void refk_smaa_edge(const uint32_t *color_unorm, int w, int h, uint8_t *edges_rg8, int y0, int y1)
{
	sampler2D s = make_sampler(color_unorm, w, h, spirv_cross::FMT_RGBA8_UNORM);
	Sh::Resources::Registers reg;
	reg.rt_metrics = glm::vec4(1.0f / (float)w, 1.0f / (float)h, (float)w, (float)h);
	glm::vec2 uv(0.0f), out(0.0f);
	glm::vec4 o0(0.0f), o1(0.0f), o2(0.0f);
	Runner r;
	s.frag_uv = &uv;
	s.frag_px = &r.pixel;
	r.resource(0, 0, &s);
	r.push(&reg, sizeof(reg));
	spirv_cross_set_stage_input(r.sh, 0, &uv, sizeof(uv));
	spirv_cross_set_stage_input(r.sh, 1, &o0, sizeof(o0));
	spirv_cross_set_stage_input(r.sh, 2, &o1, sizeof(o1));
	spirv_cross_set_stage_input(r.sh, 3, &o2, sizeof(o2));
	spirv_cross_set_stage_output(r.sh, 0, &out, sizeof(out));
	const glm::vec4 m = glm::vec4(reg.rt_metrics.x, reg.rt_metrics.y, reg.rt_metrics.x, reg.rt_metrics.y);
	const float inv_w = 1.0f / (float)w, inv_h = 1.0f / (float)h;
	for (int y = y0; y < y1; y++)
		for (int x = 0; x < w; x++)
		{
			r.pixel = glm::ivec2(x, y);
			uv = glm::vec2(((float)x + 0.5f) * inv_w, ((float)y + 0.5f) * inv_h);
			const glm::vec4 t(uv.x, uv.y, uv.x, uv.y);
			o0 = mad4(m, glm::vec4(-1.0f, 0.0f, 0.0f, -1.0f), t); // SMAAEdgeDetectionVS
			o1 = mad4(m, glm::vec4(1.0f, 0.0f, 0.0f, 1.0f), t);
			o2 = mad4(m, glm::vec4(-2.0f, 0.0f, 0.0f, -2.0f), t);
			out = glm::vec2(0.0f);
			r.itf->invoke(r.sh);
			auto q = [](float c) -> uint8_t { c = c > 0.0f ? (c < 1.0f ? c : 1.0f) : 0.0f; return (uint8_t)std::floor(c * 255.0f + 0.5f); };
			edges_rg8[2 * ((size_t)y * w + x) + 0] = q(out.x);
			edges_rg8[2 * ((size_t)y * w + x) + 1] = q(out.y);
		}
}
#elif KERNEL < 170
void refk_smaa_weights(const uint8_t *edges_rg8, int w, int h, const uint8_t *area_rg8, const uint8_t *search_r8, int max_search_steps, uint32_t *weights_rgba8,
                       int y0, int y1)
{
	sampler2D se = make_sampler(edges_rg8, w, h, spirv_cross::FMT_RG8_UNORM);
	sampler2D sa = make_sampler(area_rg8, 160, 560, spirv_cross::FMT_RG8_UNORM);
	sampler2D ss = make_sampler(search_r8, 64, 16, spirv_cross::FMT_R8_UNORM);
	Sh::Resources::Registers reg;
	reg.rt_metrics = glm::vec4(1.0f / (float)w, 1.0f / (float)h, (float)w, (float)h);
	glm::vec2 uv(0.0f), pix(0.0f);
	glm::vec4 o0(0.0f), o1(0.0f), o2(0.0f), out(0.0f);
	Runner r;
	se.frag_uv = &uv;
	se.frag_px = &r.pixel;
	r.resource(0, 0, &se);
	r.resource(0, 1, &sa);
	r.resource(0, 2, &ss);
	r.push(&reg, sizeof(reg));
	spirv_cross_set_stage_input(r.sh, 0, &uv, sizeof(uv));
	spirv_cross_set_stage_input(r.sh, 1, &pix, sizeof(pix));
	spirv_cross_set_stage_input(r.sh, 2, &o0, sizeof(o0));
	spirv_cross_set_stage_input(r.sh, 3, &o1, sizeof(o1));
	spirv_cross_set_stage_input(r.sh, 4, &o2, sizeof(o2));
	spirv_cross_set_stage_output(r.sh, 0, &out, sizeof(out));
	const glm::vec4 m = glm::vec4(reg.rt_metrics.x, reg.rt_metrics.y, reg.rt_metrics.x, reg.rt_metrics.y);
	const glm::vec4 mxxyy = glm::vec4(reg.rt_metrics.x, reg.rt_metrics.x, reg.rt_metrics.y, reg.rt_metrics.y);
	const float steps = (float)max_search_steps;
	const float inv_w = 1.0f / (float)w, inv_h = 1.0f / (float)h;
	for (int y = y0; y < y1; y++)
		for (int x = 0; x < w; x++)
		{
			r.pixel = glm::ivec2(x, y);
			uv = glm::vec2(((float)x + 0.5f) * inv_w, ((float)y + 0.5f) * inv_h);
			const glm::vec4 t(uv.x, uv.y, uv.x, uv.y);
			pix = uv * glm::vec2(reg.rt_metrics.z, reg.rt_metrics.w); // SMAABlendingWeightCalculationVS
			o0 = mad4(m, glm::vec4(-0.25f, -0.125f, 1.25f, -0.125f), t);
			o1 = mad4(m, glm::vec4(-0.125f, -0.25f, -0.125f, 1.25f), t);
			o2 = mad4(mxxyy, glm::vec4(-2.0f, 2.0f, -2.0f, 2.0f) * steps, glm::vec4(o0.x, o0.z, o1.y, o1.w));
			out = glm::vec4(0.0f);
			r.itf->invoke(r.sh);
			auto q = [](float c) -> uint32_t { c = c > 0.0f ? (c < 1.0f ? c : 1.0f) : 0.0f; return (uint32_t)std::floor(c * 255.0f + 0.5f); };
			weights_rgba8[(size_t)y * w + x] = q(out.x) | (q(out.y) << 8) | (q(out.z) << 16) | (q(out.w) << 24);
		}
}
#else
void refk_smaa_blend(const uint32_t *color_unorm, const uint32_t *weights_rgba8, int w, int h, uint32_t *out_srgb8, int y0, int y1)
{
	sampler2D sc = make_sampler(color_unorm, w, h, spirv_cross::FMT_RGBA8_UNORM);
	sampler2D sb = make_sampler(weights_rgba8, w, h, spirv_cross::FMT_RGBA8_UNORM);
	Sh::Resources::Registers reg;
	reg.rt_metrics = glm::vec4(1.0f / (float)w, 1.0f / (float)h, (float)w, (float)h);
	glm::vec2 uv(0.0f);
	glm::vec4 off(0.0f), out(0.0f);
	Runner r;
	sc.frag_uv = &uv;
	sc.frag_px = &r.pixel;
	sb.frag_uv = &uv;
	sb.frag_px = &r.pixel;
	r.resource(0, 0, &sc);
	r.resource(0, 1, &sb);
	r.push(&reg, sizeof(reg));
	spirv_cross_set_stage_input(r.sh, 0, &uv, sizeof(uv));
	spirv_cross_set_stage_input(r.sh, 1, &off, sizeof(off));
	spirv_cross_set_stage_output(r.sh, 0, &out, sizeof(out));
	const glm::vec4 m = glm::vec4(reg.rt_metrics.x, reg.rt_metrics.y, reg.rt_metrics.x, reg.rt_metrics.y);
	const float inv_w = 1.0f / (float)w, inv_h = 1.0f / (float)h;
	for (int y = y0; y < y1; y++)
		for (int x = 0; x < w; x++)
		{
			r.pixel = glm::ivec2(x, y);
			uv = glm::vec2(((float)x + 0.5f) * inv_w, ((float)y + 0.5f) * inv_h);
			off = mad4(m, glm::vec4(1.0f, 0.0f, 0.0f, 1.0f), glm::vec4(uv.x, uv.y, uv.x, uv.y)); // SMAANeighborhoodBlendingVS
			r.itf->invoke(r.sh);
			// SMAA_TARGET_SRGB: the shader decoded to linear, the sRGB attachment encodes on store
			auto a8 = [](float c) -> uint32_t { c = c > 0.0f ? (c < 1.0f ? c : 1.0f) : 0.0f; return (uint32_t)std::floor(c * 255.0f + 0.5f); };
			out_srgb8[(size_t)y * w + x] = orc_linear_to_srgb8(out.x) | (orc_linear_to_srgb8(out.y) << 8) | (orc_linear_to_srgb8(out.z) << 16) | (a8(out.w) << 24);
		}
}
#endif
#elif KERNEL == 24 || KERNEL == 25
// aa.cpp:75-118 "<output>-scale": upscale.frag, FP16 = 0; KERNEL 25 = sRGB target (TARGET_SRGB = 1).  vUV is the
// output pixel position (upscale.vert:19: (0.5 * Position + 0.5) * out_resolution), con16 = FsrEasuCon (aa.cpp:33-61).
void
#if KERNEL == 24
refk24_fsr_upscale_unorm
#else
refk25_fsr_upscale_srgb
#endif
(const uint32_t *in, int w_in, int h_in, const float *con16, uint32_t *out, int w, int h, int y0, int y1)
{
	sampler2D s = make_sampler(in, w_in, h_in, spirv_cross::FMT_RGBA8_UNORM, false); // set_unorm_texture + NearestClamp
	Sh::Resources::Params params;
	std::memcpy(&params, con16, sizeof(params));
	glm::vec2 uv(0.0f);
	glm::vec4 color(0.0f);
	Runner r;
	r.resource(0, 0, &s);
	r.resource(1, 0, &params);
	spirv_cross_set_stage_input(r.sh, 0, &uv, sizeof(uv));
	spirv_cross_set_stage_output(r.sh, 0, &color, sizeof(color));
	r.raster(w, h, y0, y1, nullptr, [&](int x, int y) {
		auto q = [](float c) -> uint32_t {
#if KERNEL == 25
			return orc_linear_to_srgb8(c);
#else
			c = c > 0.0f ? (c < 1.0f ? c : 1.0f) : 0.0f;
			return (uint32_t)std::floor(c * 255.0f + 0.5f);
#endif
		};
		out[(size_t)y * w + x] = q(color.x) | (q(color.y) << 8) | (q(color.z) << 16) | 0xff000000u;
	}, &uv);
}
#elif KERNEL == 26
// aa.cpp:120-171 "<output>-sharpen": sharpen.frag.  srgb != 0: sRGB backbuffer, the input is bound through an sRGB view.
void refk26_fsr_sharpen(const uint32_t *in, int w, int h, const float *con4, int srgb, uint32_t *out, int y0, int y1)
{
	sampler2D s = make_sampler(in, w, h, srgb ? spirv_cross::FMT_RGBA8_SRGB : spirv_cross::FMT_RGBA8_UNORM, false);
	Sh::Resources::UBO ubo;
	std::memcpy(&ubo.param0, con4, 16);
	ubo.range = glm::ivec4(0, 0, w - 1, h - 1); // aa.cpp:154-157
	glm::vec2 uv(0.0f);
	glm::vec4 color(0.0f);
	Runner r;
	r.resource(0, 0, &s);
	r.resource(1, 0, &ubo);
	spirv_cross_set_stage_input(r.sh, 0, &uv, sizeof(uv));
	spirv_cross_set_stage_output(r.sh, 0, &color, sizeof(color));
	r.raster(w, h, y0, y1, nullptr, [&](int x, int y) {
		auto q = [srgb](float c) -> uint32_t {
			if (srgb)
				return orc_linear_to_srgb8(c);
			c = c > 0.0f ? (c < 1.0f ? c : 1.0f) : 0.0f;
			return (uint32_t)std::floor(c * 255.0f + 0.5f);
		};
		out[(size_t)y * w + x] = q(color.x) | (q(color.y) << 8) | (q(color.z) << 16) | 0xff000000u;
	}, &uv);
}
#elif KERNEL == 27
// volumetric_fog.cpp:236-254: fog_accumulate.comp, dispatch ceil(w / 8) x ceil(h / 8) x 1, the shader loops over the slices
void refk27_fog_accumulate(const uint16_t *light, int w, int h, int d, uint16_t *fog)
{
	spirv_cross::sampler3D s;
	s.data = light;
	s.w = w;
	s.h = h;
	s.d = d;
	spirv_cross::image3D o;
	o.data = fog;
	o.w = w;
	o.h = h;
	o.d = d;
	Sh::Resources::Registers reg;
	reg.inv_resolution = glm::vec3(1.0f / (float)w, 1.0f / (float)h, 1.0f / (float)d);
	reg.count = glm::uvec3((unsigned)w, (unsigned)h, (unsigned)d);
	Runner r;
	r.resource(0, 1, &s);
	r.resource(0, 0, &o);
	r.push(&reg, sizeof(reg));
	r.dispatch((unsigned)((w + 7) / 8), (unsigned)((h + 7) / 8));
}
#endif
}
