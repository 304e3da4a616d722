// oracle/ref_shader_shim.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Host driver for ONE of the reference's own clusterer compute shaders, executed on the CPU:
//   GLSL (/root/reference/assets/shaders/lights/*.comp, untouched)
//     -> SPIR-V   by the reference's vendored glslang      (third_party/glslang)
//     -> C++      by the reference's vendored spirv-cross  (third_party/spirv-cross, `--cpp`)
//     -> this translation unit #includes that generated C++ (GEN_CPP) and links nothing else.
// The arithmetic that runs is therefore the reference shader's, statement for statement, on
// GLM vector types with plain IEEE fp32 (-ffp-contract=off, no FMA).  `oracle/Makefile ref-shaders`
// is the recipe; outputs live in oracle/_ref/ (git-ignored).  Nothing of the reference is copied
// into the repository; the generated C++ is a build product under oracle/_ref/gen/.
//
// One shared library per kernel (the generated files all define the same C entry points):
//   KERNEL=1  clusterer_bindless_spot_transform.comp   -> refk1_spot_transform
//   KERNEL=2  clusterer_bindless_setup.comp            -> refk2_cull_setup
//   KERNEL=3  clusterer_bindless_binning.comp (SUBGROUPS=0: one workgroup per (chunk, tile))
//                                                        -> refk3_binning
//   KERNEL=4  clusterer_bindless_z_range.comp          -> refk4_z_range
//
// Shim-provided pieces (not reference arithmetic): atomicOr (the deprecated C++ backend's runtime
// only ships atomicAdd), mix() in the Vulkan specification's form (see below) and the buffer plumbing.
#include <atomic>
#include <cstdint>
#include <cstring>
#include <vector>

#define GLM_FORCE_PURE
#include "spirv_cross/internal_interface.hpp"

namespace spirv_cross
{
template <typename T>
inline T atomicOr(T &v, T a)
{
	static_assert(sizeof(std::atomic<T>) == sizeof(T), "atomic cast");
	return std::atomic_fetch_or_explicit(reinterpret_cast<std::atomic<T> *>(&v), a, std::memory_order_relaxed);
}
} // namespace spirv_cross

// GLSL mix().  GLM evaluates x + a * (y - x); the Vulkan specification defines the operation (and
// its precision) as x * (1 - a) + y * a ("Precision and Operation of SPIR-V Instructions", FMix), so
// these non-template overloads -- preferred over GLM's templates by overload resolution -- supply
// that form.  It is reached only by the near-plane triangle clipping of K2
// (clusterer_bindless_setup.comp:71-111); everything else is GLM's / the compiler's IEEE arithmetic.
inline float mix(const float &x, const float &y, const float &a) { return x * (1.0f - a) + y * a; }
inline glm::vec2 mix(const glm::vec2 &x, const glm::vec2 &y, const glm::vec2 &a) { return x * (glm::vec2(1.0f) - a) + y * a; }
inline glm::vec3 mix(const glm::vec3 &x, const glm::vec3 &y, const glm::vec3 &a) { return x * (glm::vec3(1.0f) - a) + y * a; }
inline glm::vec4 mix(const glm::vec4 &x, const glm::vec4 &y, const glm::vec4 &a) { return x * (glm::vec4(1.0f) - a) + y * a; }

#include GEN_CPP

namespace
{
using Sh = Impl::Shader;

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
	void dispatch(unsigned gx, unsigned gy, unsigned gz)
	{
		glm::uvec3 num(gx, gy, gz), id(0);
		spirv_cross_set_builtin(sh, SPIRV_CROSS_BUILTIN_NUM_WORK_GROUPS, &num, sizeof(num));
		spirv_cross_set_builtin(sh, SPIRV_CROSS_BUILTIN_WORK_GROUP_ID, &id, sizeof(id));
		for (unsigned z = 0; z < gz; z++)
			for (unsigned y = 0; y < gy; y++)
				for (unsigned x = 0; x < gx; x++)
				{
					id = glm::uvec3(x, y, z);
					itf->invoke(sh);
				}
	}
};

#if KERNEL != 4 && KERNEL != 5
// ClustererBindlessTransforms (assets/shaders/lights/clusterer_data.h:46-53) filled from the packed
// host arrays: 48-byte light records, mat_affine rows, type mask.
std::vector<unsigned char> make_transforms(const void *lights48, const float *model_rows12, const uint32_t *type_mask, int n)
{
	std::vector<unsigned char> blob(sizeof(Sh::ClustererBindlessTransforms), 0);
	auto *t = reinterpret_cast<Sh::ClustererBindlessTransforms *>(blob.data());
	static_assert(sizeof(Sh::PositionalLightInfo) == 48, "light record layout");
	static_assert(sizeof(Sh::mat_affine) == 48, "mat_affine layout");
	if (lights48)
		std::memcpy(t->lights.data(), lights48, (size_t)n * 48);
	if (model_rows12)
		std::memcpy(t->model.data(), model_rows12, (size_t)n * 48);
	if (type_mask)
		std::memcpy(t->type_mask.data(), type_mask, (size_t)((n + 31) / 32) * 4);
	return blob;
}
#endif

#if KERNEL == 2 || KERNEL == 3
// ClustererParametersBindless (clusterer_data.h / math/render_parameters.hpp:90-108), by field.
void fill_params(Sh::ClustererParametersBindless &q, const float *transform16, const float *clip_scale4, const float *camera_base3,
                 const float *camera_front3, const float *xy_scale2, const int32_t *resolution_xy2, const float *inv_resolution_xy2,
                 int num_lights, int num_lights_32, int z_max_index, float z_scale)
{
	std::memset(&q, 0, sizeof(q));
	std::memcpy(&q.transform, transform16, 64);
	q.clip_scale = glm::vec4(clip_scale4[0], clip_scale4[1], clip_scale4[2], clip_scale4[3]);
	q.camera_base = glm::vec3(camera_base3[0], camera_base3[1], camera_base3[2]);
	q.camera_front = glm::vec3(camera_front3[0], camera_front3[1], camera_front3[2]);
	q.xy_scale = glm::vec2(xy_scale2[0], xy_scale2[1]);
	q.resolution_xy = glm::ivec2(resolution_xy2[0], resolution_xy2[1]);
	q.inv_resolution_xy = glm::vec2(inv_resolution_xy2[0], inv_resolution_xy2[1]);
	q.num_lights = num_lights;
	q.num_lights_32 = num_lights_32;
	q.z_max_index = z_max_index;
	q.z_scale = z_scale;
}
#endif
} // namespace

extern "C" {
#if KERNEL == 1
// dispatch (num_lights + 63) / 64 (renderer/lights/clusterer.cpp:1475-1495)
void refk1_spot_transform(const float *vp16, const float *camera_pos3, const float *camera_front3, float z_near, float z_far,
                          const float *model_rows12, int n, float *out_spots24)
{
	auto blob = make_transforms(nullptr, model_rows12, nullptr, n);
	Sh::Resources::Registers reg;
	std::memcpy(&reg.vp, vp16, 64);
	reg.camera_pos = glm::vec3(camera_pos3[0], camera_pos3[1], camera_pos3[2]);
	reg.num_lights = (uint32_t)n;
	reg.camera_front = glm::vec3(camera_front3[0], camera_front3[1], camera_front3[2]);
	reg.z_near = z_near;
	reg.z_far = z_far;
	static_assert(sizeof(Sh::TransformedSpot) == 96, "TransformedSpot layout");
	Runner r;
	r.resource(0, 0, blob.data());
	r.resource(0, 1, out_spots24);
	r.push(&reg, sizeof(reg));
	r.dispatch((unsigned)(n + 63) / 64, 1, 1);
}
#elif KERNEL == 2
// dispatch (num_lights + 63) / 64 (clusterer.cpp:1500-1511)
void refk2_cull_setup(const float *view16, const float *transform16, const float *clip_scale4, const float *camera_base3,
                      const float *camera_front3, const float *xy_scale2, const int32_t *resolution_xy2, const float *inv_resolution_xy2,
                      int num_lights_32, int z_max_index, float z_scale, const void *lights48, const uint32_t *type_mask,
                      const float *spots24, int n, float *out_cull128)
{
	auto blob = make_transforms(lights48, nullptr, type_mask, n);
	Sh::Resources::ClustererParameters ubo;
	fill_params(ubo.parameters, transform16, clip_scale4, camera_base3, camera_front3, xy_scale2, resolution_xy2, inv_resolution_xy2, n,
	            num_lights_32, z_max_index, z_scale);
	Sh::Resources::Registers reg;
	std::memcpy(&reg.view, view16, 64);
	reg.num_lights = (uint32_t)n;
	static_assert(sizeof(Sh::CullSetup) == 512, "CullSetup layout");
	Runner r;
	r.resource(0, 2, out_cull128);
	r.resource(0, 0, blob.data());
	r.resource(1, 0, &ubo);
	r.resource(0, 1, const_cast<float *>(spots24));
	r.push(&reg, sizeof(reg));
	r.dispatch((unsigned)(n + 63) / 64, 1, 1);
}
#elif KERNEL == 3
// SUBGROUPS=0 path: dispatch (num_lights_32, resolution_x, resolution_y), one 32-thread workgroup per
// (chunk, tile) (clusterer_bindless_binning.comp:177-215).  tile_x0/x1, tile_y0/y1 restrict the
// dispatch to a window of tiles (the others keep whatever out_bitmask holds) so tests stay fast.
void refk3_binning(const float *clip_scale4, const int32_t *resolution_xy2, const float *inv_resolution_xy2, int num_lights_32,
                   const uint32_t *type_mask, const float *cull128, int n, int tile_x0, int tile_x1, int tile_y0, int tile_y1,
                   uint32_t *out_bitmask)
{
	auto blob = make_transforms(nullptr, nullptr, type_mask, n);
	Sh::Resources::ClustererParameters ubo;
	const float zero16[16] = {}, zero3[3] = {}, one2[2] = { 1.0f, 1.0f };
	fill_params(ubo.parameters, zero16, clip_scale4, zero3, zero3, one2, resolution_xy2, inv_resolution_xy2, n, num_lights_32, 0, 0.0f);
	Runner r;
	r.resource(0, 2, const_cast<float *>(cull128));
	r.resource(1, 0, &ubo);
	r.resource(0, 0, blob.data());
	r.resource(0, 3, out_bitmask);
	glm::uvec3 num((unsigned)num_lights_32, (unsigned)resolution_xy2[0], (unsigned)resolution_xy2[1]), id(0);
	spirv_cross_set_builtin(r.sh, SPIRV_CROSS_BUILTIN_NUM_WORK_GROUPS, &num, sizeof(num));
	spirv_cross_set_builtin(r.sh, SPIRV_CROSS_BUILTIN_WORK_GROUP_ID, &id, sizeof(id));
	for (int ty = tile_y0; ty < tile_y1; ty++)
		for (int tx = tile_x0; tx < tile_x1; tx++)
			for (int c = 0; c < num_lights_32; c++)
			{
				id = glm::uvec3((unsigned)c, (unsigned)tx, (unsigned)ty);
				r.itf->invoke(r.sh);
			}
}
#elif KERNEL == 5
// clusterer_bindless_binning_decal.comp, SUBGROUPS=0 path: dispatch (num_decals_32, resolution_x, resolution_y), one
// 32-thread workgroup per (chunk, tile) (clusterer.cpp:1454-1457).  mvps16: num_decals column-major mat4.
void refk5_decal_binning(const int32_t *resolution_xy2, const float *inv_resolution_xy2, int num_decals, const float *mvps16, uint32_t *out_bitmask)
{
	Sh::Resources::ClustererParameters ubo;
	std::memset(&ubo, 0, sizeof(ubo));
	ubo.parameters.resolution_xy = glm::ivec2(resolution_xy2[0], resolution_xy2[1]);
	ubo.parameters.inv_resolution_xy = glm::vec2(inv_resolution_xy2[0], inv_resolution_xy2[1]);
	ubo.parameters.num_decals = num_decals;
	ubo.parameters.num_decals_32 = (num_decals + 31) / 32;
	Runner r;
	r.resource(2, 0, const_cast<float *>(mvps16));
	r.resource(1, 0, &ubo);
	r.resource(0, 0, out_bitmask);
	r.dispatch((unsigned)((num_decals + 31) / 32), (unsigned)resolution_xy2[0], (unsigned)resolution_xy2[1]);
}
#elif KERNEL == 4
// naive form = the specification; dispatch res_z / 64 (clusterer.cpp:1286-1300)
void refk4_z_range(const uint32_t *z_ranges, int num_ranges, int res_z, uint32_t *out_cluster_range)
{
	Sh::Resources::Registers reg;
	reg.num_lights = (uint32_t)num_ranges;
	Runner r;
	r.resource(0, 0, const_cast<uint32_t *>(z_ranges));
	r.resource(0, 1, out_cluster_range);
	r.push(&reg, sizeof(reg));
	r.dispatch((unsigned)(res_z + 63) / 64, 1, 1);
}
#endif
}
