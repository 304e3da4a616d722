// oracle/ref_light_shim.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Host driver for the reference's own deferred-lighting fragment shaders, executed on the CPU:
//   KERNEL=5  assets/shaders/lights/clustering.frag  (K5: clustered point / spot lights)
//   KERNEL=6  assets/shaders/lights/directional.frag (K6: directional light, VOLUMETRIC_DIFFUSE_FALLBACK,
//             no shadows -- the variant renderer.cpp:1018-1105 selects for the viewer without shadow maps)
//   KERNEL=7  clustering.frag with POSITIONAL_LIGHTS_SHADOW=1 (renderer.cpp:369,1126): shadowed point / spot lights
//             through point.h:45-74, spot.h:51-77, pcf.h:98-99.  The shader's statements (clip transform, reference
//             depth, the products around shadow_falloff) are the reference's; the two comparison samplers
//             (sampler2DShadow / samplerCubeShadow behind StockSampler::LinearShadow) are a texture unit's job and
//             are supplied here by the oracle's restatement of the Vulkan filtering rules.
// GLSL where it lies under /root/reference -> SPIR-V (vendored glslang) -> C++ (vendored spirv-cross
// `--cpp --vulkan-semantics`) -> #included here (GEN_CPP); see oracle/Makefile `ref-shaders`.  The
// statements that run are the reference's: compute_cluster_light (clusterer_bindless.h:29-84),
// cluster_mask_range, compute_point_light / compute_spot_light (point.h, spot.h), the BRDF (pbr.h),
// compute_lighting (lighting.h:6-82).
//
// What the shim supplies (nothing of it is shader arithmetic):
//   * subpassLoad: the G-buffer texel of the fragment, decoded by the attachment formats
//     (R8G8B8A8_SRGB through the oracle's EOTF table, A2B10G10R10 / R8G8 as n / (2^k - 1), D32);
//   * vClip: clustering.vert:10-14 outputs invVP * (ndc.xy, 0, 1) at the three vertices of the
//     full-screen triangle and the rasteriser interpolates it; per pixel that is
//     (col0 * ndc.x + col1 * ndc.y) + col3 with ndc = 2 (frag + 0.5) / size - 1;
//   * a subgroup of ONE invocation: subgroupMin / Max / Or are the identity, i.e. every pixel walks
//     exactly its own (tile, Z-slice) mask.  (On a GPU the OR over a subgroup only ADDS lights whose
//     range falloff is exactly 0 at the pixels that did not have them: tests/test_oracle_cpu.py's
//     brute-force test shows the sum does not change by a bit.)
//   * the SSBOs: spirv-cross's deprecated C++ backend does not declare StorageBuffer-class blocks, so the
//     three block names the generated code uses are rewritten by the Makefile's sed to the *_BLOCK macros
//     below, which are plain pointers.
#include <cmath>
#include <cstdint>
#include <cstring>

#define GLM_FORCE_PURE
#ifndef GLM_SWIZZLE
#define GLM_SWIZZLE
#endif
#ifndef GLM_FORCE_RADIANS
#define GLM_FORCE_RADIANS
#endif
#include <glm/glm.hpp>

extern "C" float orc_srgb8_to_linear(uint32_t v);

#define SPIRV_CROSS_SAMPLER_HPP
#define SPIRV_CROSS_IMAGE_HPP
namespace spirv_cross
{
struct subpassInput
{
	glm::vec4 value;
};
inline glm::vec4 subpassLoad(const subpassInput &s) { return s.value; }
struct sampler2D
{
	int unused;
};
#if KERNEL == 7 || KERNEL == 8
// bindless shadow maps: one D16 image per light; a null map samples as 1.0 ("casts no shadow")
struct texture2D
{
	const uint16_t *map;
	int res;
};
struct textureCube
{
	const uint16_t *map;
	int res;
};
struct sampler
{
	int unused;
};
struct sampler2DShadow
{
	texture2D t;
	sampler2DShadow(const texture2D &t_, const sampler &) : t(t_) {}
};
struct samplerCubeShadow
{
	textureCube t;
	samplerCubeShadow(const textureCube &t_, const sampler &) : t(t_) {}
};
template <typename T> inline T nonuniformEXT(const T &v) { return v; }
extern "C" float orc_shadow_sample_2d(const uint16_t *map, int res, float clip_x, float clip_y, float clip_z, float clip_w);
extern "C" float orc_shadow_sample_cube(const uint16_t *map, int res, float dx, float dy, float dz, float ref);
// textureProjLod(sampler2DShadow, vec4(s, t, D_ref, q), lod): (s, t, D_ref) / q
inline float textureProjLod(const sampler2DShadow &s, const glm::vec4 &p, float) { return s.t.map ? orc_shadow_sample_2d(s.t.map, s.t.res, p.x, p.y, p.z, p.w) : 1.0f; }
// texture(samplerCubeShadow, vec4(direction, D_ref))
inline float texture(const samplerCubeShadow &s, const glm::vec4 &p) { return s.t.map ? orc_shadow_sample_cube(s.t.map, s.t.res, p.x, p.y, p.z, p.w) : 1.0f; }
// KERNEL 8 (SHADOW_MAP_PCF_KERNEL_WIDE, pcf.h:7-80): textureSize and the comparison gathers -- four texels of the bilinear
// footprint at uv (+ an integer offset), each compared GREATER_OR_EQUAL with the clamped reference; components
// x = (i0, j1), y = (i1, j1), z = (i1, j0), w = (i0, j0); clamp to edge.  A null map compares as lit.
inline glm::ivec2 textureSize(const texture2D &t, int) { return glm::ivec2(t.res, t.res); }
inline glm::vec4 textureGatherOffset(const sampler2DShadow &s, const glm::vec2 &uv, float ref, const glm::ivec2 &off)
{
	if (!s.t.map)
		return glm::vec4(1.0f);
	const int res = s.t.res;
	ref = ref > 0.0f ? (ref < 1.0f ? ref : 1.0f) : 0.0f;
	const float fx = uv.x * (float)res - 0.5f, fy = uv.y * (float)res - 0.5f;
	const int x0 = (int)std::floor(fx) + off.x, y0 = (int)std::floor(fy) + off.y;
	auto cmp = [&](int x, int y) -> float {
		x = x < 0 ? 0 : (x > res - 1 ? res - 1 : x);
		y = y < 0 ? 0 : (y > res - 1 ? res - 1 : y);
		return ref >= (float)s.t.map[(size_t)y * res + x] / 65535.0f ? 1.0f : 0.0f;
	};
	return glm::vec4(cmp(x0, y0 + 1), cmp(x0 + 1, y0 + 1), cmp(x0 + 1, y0), cmp(x0, y0));
}
inline glm::vec4 textureGather(const sampler2DShadow &s, const glm::vec2 &uv, float ref) { return textureGatherOffset(s, uv, ref, glm::ivec2(0)); }
#endif
#if KERNEL == 9
// fog_light_density.comp: a texel buffer of floats, an array texture of RGBA8 texels, an RGBA16F storage volume
extern "C" uint16_t orc_f32_to_f16(float f);
struct samplerBuffer
{
	const float *data;
};
inline glm::vec4 texelFetch(const samplerBuffer &b, int i) { return glm::vec4(b.data[i], 0.0f, 0.0f, 1.0f); }
struct texture2DArray
{
	const uint32_t *data; // layers x 128 x 128, R8G8B8A8_UNORM
	int w, h;
};
inline glm::vec4 texelFetch(const texture2DArray &t, const glm::ivec3 &p, int)
{
	const uint32_t v = t.data[((size_t)p.z * t.h + p.y) * t.w + p.x];
	return glm::vec4((float)(v & 255u) / 255.0f, (float)((v >> 8) & 255u) / 255.0f, (float)((v >> 16) & 255u) / 255.0f, (float)(v >> 24) / 255.0f);
}
struct image3D
{
	uint16_t *data;
	int w, h, d;
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
#endif
template <typename T> inline T subgroupMin(T v) { return v; }
template <typename T> inline T subgroupMax(T v) { return v; }
template <typename T> inline T subgroupOr(T v) { return v; }
} // namespace spirv_cross

#include "spirv_cross/internal_interface.hpp"

// GLSL mix(): the Vulkan specification's form x * (1 - a) + y * a (see ref_shader_shim.cpp)
inline float mix(const float &x, const float &y, const float &a) { return x * (1.0f - a) + y * a; }
inline glm::vec3 mix(const glm::vec3 &x, const glm::vec3 &y, const glm::vec3 &a) { return x * (glm::vec3(1.0f) - a) + y * a; }

namespace shim
{
static const void *g_transforms = nullptr;
static const void *g_bitmask = nullptr;
static const void *g_range = nullptr;
#if KERNEL == 7 || KERNEL == 8
static const spirv_cross::texture2D *g_spot_atlas = nullptr;   // indexed by light (uSpotShadowAtlas[index])
static const spirv_cross::textureCube *g_point_atlas = nullptr; // the same descriptors seen as cubes (uPointShadowAtlas[index])
#endif
struct BitmaskBlock
{
	uint32_t cluster_bitmask[1];
};
struct RangeBlock
{
	glm::uvec2 cluster_range[1];
};
template <typename S>
struct TransformsBlock
{
	typename S::ClustererBindlessTransforms cluster_transforms;
};
} // namespace shim
#define BITMASK_BLOCK (*static_cast<const shim::BitmaskBlock *>(shim::g_bitmask))
#define RANGE_BLOCK (*static_cast<const shim::RangeBlock *>(shim::g_range))
#define TRANSFORMS_BLOCK (*static_cast<const shim::TransformsBlock<Shader> *>(shim::g_transforms))

#include GEN_CPP

#if KERNEL != 9
namespace
{
using Sh = Impl::Shader;
using spirv_cross::subpassInput;

struct GBufferIn
{
	int w, h;
	const uint32_t *albedo, *normal;
	const uint16_t *pbr;
	const float *depth;
};

struct Frag
{
	subpassInput base, normal, pbr, depth;
	glm::vec4 vclip, frag_coord;
	glm::vec3 color;
};

// returns false for sky pixels (depth == 0: never shaded, renderer.cpp:1056-1057)
bool load_fragment(const GBufferIn &g, const float *ivp, int x, int y, Frag &f)
{
	const size_t i = (size_t)y * g.w + x;
	const float depth = g.depth[i];
	if (depth == 0.0f)
		return false;
	const uint32_t a = g.albedo[i], n = g.normal[i];
	const uint32_t mr = g.pbr[i];
	f.base.value = glm::vec4(orc_srgb8_to_linear(a & 255u), orc_srgb8_to_linear((a >> 8) & 255u), orc_srgb8_to_linear((a >> 16) & 255u), (float)(a >> 24) / 255.0f);
	f.normal.value = glm::vec4((float)(n & 1023u) / 1023.0f, (float)((n >> 10) & 1023u) / 1023.0f, (float)((n >> 20) & 1023u) / 1023.0f, (float)(n >> 30) / 3.0f);
	f.pbr.value = glm::vec4((float)(mr & 255u) / 255.0f, (float)(mr >> 8) / 255.0f, 0.0f, 1.0f);
	f.depth.value = glm::vec4(depth, 0.0f, 0.0f, 1.0f);
	const float inv_x = 1.0f / (float)g.w, inv_y = 1.0f / (float)g.h;
	const float ndc_x = 2.0f * ((float)x + 0.5f) * inv_x - 1.0f, ndc_y = 2.0f * ((float)y + 0.5f) * inv_y - 1.0f;
	f.vclip = glm::vec4(ivp[0] * ndc_x + ivp[4] * ndc_y + ivp[12], ivp[1] * ndc_x + ivp[5] * ndc_y + ivp[13], ivp[2] * ndc_x + ivp[6] * ndc_y + ivp[14],
	                    ivp[3] * ndc_x + ivp[7] * ndc_y + ivp[15]);
	f.frag_coord = glm::vec4((float)x + 0.5f, (float)y + 0.5f, depth, 1.0f);
	return true;
}

struct Runner
{
	spirv_cross_shader_t *sh;
	const spirv_cross_interface *itf;
	Frag f;
	Runner() : sh(nullptr), itf(spirv_cross_get_interface())
	{
		sh = itf->construct();
		resource(3, 0, &f.base);
		resource(3, 1, &f.normal);
		resource(3, 2, &f.pbr);
		resource(3, 3, &f.depth);
		spirv_cross_set_stage_input(sh, 0, &f.vclip, sizeof(f.vclip));
		spirv_cross_set_stage_output(sh, 0, &f.color, sizeof(f.color));
		spirv_cross_set_builtin(sh, SPIRV_CROSS_BUILTIN_FRAG_COORD, &f.frag_coord, sizeof(f.frag_coord));
#if KERNEL == 7 || KERNEL == 8
		static spirv_cross::sampler linear_shadow_sampler = { 0 }; // LinearShadowSampler: behaviour lives in the sample functions
		spirv_cross_set_uniform_constant(sh, 0, &linear_shadow_sampler, sizeof(linear_shadow_sampler));
#endif
	}
	~Runner() { itf->destruct(sh); }
	void resource(unsigned set, unsigned binding, void *ptr)
	{
		void *p = ptr;
		spirv_cross_set_resource(sh, set, binding, &p, sizeof(p));
	}
	// out_rgb: h x w x 3 floats, the fragment colour of this draw (0 where nothing was drawn)
	void draw(const GBufferIn &g, const float *ivp, int y0, int y1, float *out_rgb)
	{
		for (int y = y0; y < y1; y++)
			for (int x = 0; x < g.w; x++)
			{
				float *o = out_rgb + ((size_t)y * g.w + x) * 3;
				o[0] = o[1] = o[2] = 0.0f;
				if (!load_fragment(g, ivp, x, y, f))
					continue;
				itf->invoke(sh);
				o[0] = f.color.x;
				o[1] = f.color.y;
				o[2] = f.color.z;
			}
	}
};
} // namespace
#endif

extern "C" {
#if KERNEL == 5 || KERNEL == 7 || KERNEL == 8
// renderer.cpp:1107-1156.  The cluster parameters are the oracle's orc_cluster_params_t fields.
#if KERNEL == 7 || KERNEL == 8
// transforms16: num_lights x mat4 (ClustererBindlessTransforms::shadow); maps: num_lights pointers (null = no shadow)
#if KERNEL == 8
void refk8_clustering_shadowed_pcf_wide(
#else
void refk7_clustering_shadowed(
#endif
                      const float *transforms16, const uint16_t *const *maps, int shadow_res, int w, int h,
#else
void refk5_clustering(int w, int h,
#endif
                      const uint32_t *albedo, const uint32_t *normal, const uint16_t *pbr, const float *depth, const float *inv_view_projection16,
                      const float *camera_pos3, const float *camera_base3, const float *camera_front3, const float *xy_scale2, const int32_t *resolution_xy2,
                      int num_lights, int num_lights_32, int z_max_index, float z_scale, const void *lights48, const uint32_t *type_mask, const uint32_t *bitmask,
                      const uint32_t *cluster_range, int y0, int y1, float *out_rgb)
{
	static_assert(sizeof(Sh::PositionalLightInfo) == 48, "light record layout");
	auto *blob = new shim::TransformsBlock<Sh>();
	std::memset(static_cast<void *>(blob), 0, sizeof(*blob));
	std::memcpy(blob->cluster_transforms.lights.data(), lights48, (size_t)num_lights * 48);
	std::memcpy(blob->cluster_transforms.type_mask.data(), type_mask, (size_t)num_lights_32 * 4);
#if KERNEL == 7 || KERNEL == 8
	std::memcpy(static_cast<void *>(blob->cluster_transforms.shadow.data()), transforms16, (size_t)num_lights * 64);
	auto *spot_atlas = new spirv_cross::texture2D[num_lights > 0 ? num_lights : 1];
	auto *point_atlas = new spirv_cross::textureCube[num_lights > 0 ? num_lights : 1];
	for (int i = 0; i < num_lights; i++)
	{
		spot_atlas[i].map = point_atlas[i].map = maps[i];
		spot_atlas[i].res = point_atlas[i].res = shadow_res;
	}
	shim::g_spot_atlas = spot_atlas;
	shim::g_point_atlas = point_atlas;
#endif
	shim::g_transforms = blob;
	shim::g_bitmask = bitmask;
	shim::g_range = cluster_range;
	Sh::Resources::ClusterParameters ubo;
	std::memset(static_cast<void *>(&ubo), 0, sizeof(ubo));
	ubo.cluster.camera_base = glm::vec3(camera_base3[0], camera_base3[1], camera_base3[2]);
	ubo.cluster.camera_front = glm::vec3(camera_front3[0], camera_front3[1], camera_front3[2]);
	ubo.cluster.xy_scale = glm::vec2(xy_scale2[0], xy_scale2[1]);
	ubo.cluster.resolution_xy = glm::ivec2(resolution_xy2[0], resolution_xy2[1]);
	ubo.cluster.num_lights = num_lights;
	ubo.cluster.num_lights_32 = num_lights_32;
	ubo.cluster.z_max_index = z_max_index;
	ubo.cluster.z_scale = z_scale;
	Sh::Resources::Registers reg;
	std::memset(static_cast<void *>(&reg), 0, sizeof(reg));
	reg.inverse_view_projection_col2 = glm::vec4(inv_view_projection16[8], inv_view_projection16[9], inv_view_projection16[10], inv_view_projection16[11]);
	reg.camera_pos = glm::vec3(camera_pos3[0], camera_pos3[1], camera_pos3[2]);
	reg.inv_resolution = glm::vec2(1.0f / (float)w, 1.0f / (float)h); // renderer.cpp:1101-1102,1120
	GBufferIn g = { w, h, albedo, normal, pbr, depth };
	{
		Runner r;
		r.resource(0, 8, &ubo);
		spirv_cross_set_push_constant(r.sh, &reg, sizeof(reg));
		r.draw(g, inv_view_projection16, y0, y1, out_rgb);
	}
	delete blob;
#if KERNEL == 7 || KERNEL == 8
	delete[] spot_atlas;
	delete[] point_atlas;
#endif
}
#elif KERNEL == 6
// renderer.cpp:1018-1105
void refk6_directional(int w, int h, const uint32_t *albedo, const uint32_t *normal, const uint16_t *pbr, const float *depth, const float *inv_view_projection16,
                       const float *camera_pos3, const float *camera_front3, const float *dir_color3, const float *dir_direction3, int y0, int y1, float *out_rgb)
{
	Sh::Resources::Registers reg;
	std::memset(static_cast<void *>(&reg), 0, sizeof(reg));
	reg.inverse_view_projection_col2 = glm::vec4(inv_view_projection16[8], inv_view_projection16[9], inv_view_projection16[10], inv_view_projection16[11]);
	reg.color = glm::vec3(dir_color3[0], dir_color3[1], dir_color3[2]);
	reg.camera_pos = glm::vec3(camera_pos3[0], camera_pos3[1], camera_pos3[2]);
	reg.direction = glm::vec3(dir_direction3[0], dir_direction3[1], dir_direction3[2]);
	reg.camera_front = glm::vec3(camera_front3[0], camera_front3[1], camera_front3[2]);
	reg.inv_resolution = glm::vec2(1.0f / (float)w, 1.0f / (float)h);
	GBufferIn g = { w, h, albedo, normal, pbr, depth };
	Runner r;
	spirv_cross_set_push_constant(r.sh, &reg, sizeof(reg));
	r.draw(g, inv_view_projection16, y0, y1, out_rgb);
}
#elif KERNEL == 9
// volumetric_fog.cpp:142-228: fog_light_density.comp (base variant), dispatch ceil(w / 4) x ceil(h / 4) x ceil(d / 4) groups of 64
void refk9_fog_light_density(int w, int h, int d, int dither_offset, float slice_z_log2_scale, float density_mod, float in_scatter_strength,
                             const float *inv_view_projection16, const float *projection16, const float *inv_projection16, const float *camera_pos3,
                             const float *camera_front3, const float *dir_color3, const float *dir_direction3, const float *cluster_transform16,
                             const float *camera_base3, const float *cluster_front3, const float *xy_scale2, const int32_t *resolution_xy2, int num_lights,
                             int num_lights_32, int z_max_index, float z_scale, const void *lights48, const uint32_t *type_mask, const uint32_t *bitmask,
                             const uint32_t *cluster_range, const float *slice_extents, const uint32_t *dither_lut, uint16_t *out)
{
	using Sh = Impl::Shader;
	auto *blob = new shim::TransformsBlock<Sh>();
	std::memset(static_cast<void *>(blob), 0, sizeof(*blob));
	std::memcpy(blob->cluster_transforms.lights.data(), lights48, (size_t)num_lights * 48);
	std::memcpy(blob->cluster_transforms.type_mask.data(), type_mask, (size_t)num_lights_32 * 4);
	shim::g_transforms = blob;
	shim::g_bitmask = bitmask;
	shim::g_range = cluster_range;
	Sh::Resources::ClusterParameters ubo;
	std::memset(static_cast<void *>(&ubo), 0, sizeof(ubo));
	std::memcpy(&ubo.cluster.transform, cluster_transform16, 64);
	ubo.cluster.camera_base = glm::vec3(camera_base3[0], camera_base3[1], camera_base3[2]);
	ubo.cluster.camera_front = glm::vec3(cluster_front3[0], cluster_front3[1], cluster_front3[2]);
	ubo.cluster.xy_scale = glm::vec2(xy_scale2[0], xy_scale2[1]);
	ubo.cluster.resolution_xy = glm::ivec2(resolution_xy2[0], resolution_xy2[1]);
	ubo.cluster.num_lights = num_lights;
	ubo.cluster.num_lights_32 = num_lights_32;
	ubo.cluster.z_max_index = z_max_index;
	ubo.cluster.z_scale = z_scale;
	Sh::Resources::Registers reg;
	std::memset(static_cast<void *>(&reg), 0, sizeof(reg));
	std::memcpy(&reg.inv_view_projection, inv_view_projection16, 64);
	reg.z_transform = glm::vec4(projection16[10], projection16[11], projection16[14], projection16[15]); // :161-162
	reg.count = glm::uvec3((unsigned)w, (unsigned)h, (unsigned)d);
	reg.dither_offset = dither_offset;
	reg.inv_resolution = glm::vec3(1.0f / (float)w, 1.0f / (float)h, 1.0f / (float)d);
	reg.in_scatter_strength = in_scatter_strength;
	reg.xy_scale = glm::vec2(inv_projection16[0], inv_projection16[5]); // :167-168
	reg.slice_z_log2_scale = slice_z_log2_scale;
	reg.density_mod = density_mod;
	Sh::Resources::RenderParameters render_ubo; // ("global" and "registers" are macros of the generated code)
	std::memset(static_cast<void *>(&render_ubo), 0, sizeof(render_ubo));
	render_ubo.camera_position = glm::vec3(camera_pos3[0], camera_pos3[1], camera_pos3[2]);
	render_ubo.camera_front = glm::vec3(camera_front3[0], camera_front3[1], camera_front3[2]);
	Sh::Resources::LightingParameters lighting;
	std::memset(static_cast<void *>(&lighting), 0, sizeof(lighting));
	lighting.directional.color = glm::vec3(dir_color3[0], dir_color3[1], dir_color3[2]);
	lighting.directional.direction = glm::vec3(dir_direction3[0], dir_direction3[1], dir_direction3[2]);
	spirv_cross::samplerBuffer extents = { slice_extents };
	spirv_cross::texture2DArray lut = { dither_lut, 128, 128 };
	spirv_cross::image3D image = { out, w, h, d };
	spirv_cross::sampler2D brdf = { 0 };
	const spirv_cross_interface *itf = spirv_cross_get_interface();
	spirv_cross_shader_t *sh = itf->construct();
	auto bind = [&](unsigned set, unsigned binding, void *ptr) {
		void *p = ptr;
		spirv_cross_set_resource(sh, set, binding, &p, sizeof(p));
	};
	bind(0, 8, &ubo);
	bind(3, 0, &reg);
	bind(0, 0, &render_ubo);
	bind(0, 1, &lighting);
	bind(2, 1, &extents);
	bind(2, 2, &lut);
	bind(2, 0, &image);
	bind(0, 4, &brdf);
	glm::uvec3 num((unsigned)((w + 3) / 4), (unsigned)((h + 3) / 4), (unsigned)((d + 3) / 4)), id(0);
	spirv_cross_set_builtin(sh, SPIRV_CROSS_BUILTIN_NUM_WORK_GROUPS, &num, sizeof(num));
	spirv_cross_set_builtin(sh, SPIRV_CROSS_BUILTIN_WORK_GROUP_ID, &id, sizeof(id));
	for (unsigned z = 0; z < num.z; z++)
		for (unsigned y = 0; y < num.y; y++)
			for (unsigned x = 0; x < num.x; x++)
			{
				id = glm::uvec3(x, y, z);
				itf->invoke(sh);
			}
	itf->destruct(sh);
	delete blob;
}
#endif
}
