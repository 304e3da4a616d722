// ref_shim.cpp -- TEST INFRASTRUCTURE ONLY.
// Thin extern "C" exports over the REFERENCE's own host math (math/muglm, math/transforms),
// compiled from the sources where they lie under /root/reference into oracle/_ref/
// (see oracle/Makefile target `ref`).  Used only to pin oracle_host.c's restatements
// (perspective, inverse, mat4 multiply, floatToHalf, camera look_at, frustum planes, AABB
// transform and the frustum/AABB test of the light gather) bit-for-bit.
// No reference source is copied into this repository.
#include "muglm/muglm_impl.hpp"
#include "muglm/matrix_helper.hpp"
#include "transforms.hpp"
#include "aabb.hpp"
#include "frustum.hpp"
#include "simd.hpp"
#include <cstring>

using namespace muglm;

extern "C" {
void ref_perspective(float fovy, float aspect, float z_near, float z_far, float *out16)
{
	mat4 m = perspective(fovy, aspect, z_near, z_far);
	memcpy(out16, &m, 64);
}

void ref_mat4_inverse(const float *in16, float *out16)
{
	mat4 m;
	memcpy(&m, in16, 64);
	mat4 r = inverse(m);
	memcpy(out16, &r, 64);
}

void ref_mat4_mul(const float *a16, const float *b16, float *out16)
{
	mat4 a, b;
	memcpy(&a, a16, 64);
	memcpy(&b, b16, 64);
	mat4 r = a * b;
	memcpy(out16, &r, 64);
}

uint16_t ref_float_to_half(float v)
{
	return floatToHalf(v);
}

// renderer/camera.cpp:61-64,76-80: view = mat4_cast(look_at(at - eye, up)) * translate(-eye)
void ref_camera_view(const float *eye, const float *at, const float *up, float *out16)
{
	vec3 e(eye[0], eye[1], eye[2]), a(at[0], at[1], at[2]), u(up[0], up[1], up[2]);
	quat rot = Granite::look_at(a - e, u);
	mat4 view = mat4_cast(rot) * translate(-e);
	memcpy(out16, &view, 64);
}

float ref_infinite_far_plane(void)
{
	return InfiniteFarPlane;
}

// math/frustum.cpp:109-156
void ref_frustum_planes(const float *inv_vp16, float *planes24)
{
	mat4 m;
	memcpy(&m, inv_vp16, 64);
	Granite::Frustum f;
	f.build_planes(m);
	memcpy(planes24, f.get_planes(), 6 * 16);
}

// math/simd.hpp:386-419 (mat_affine overload): what Scene::update_cached_transforms_range applies
void ref_transform_aabb(const float *rows12, const float *lo3, const float *hi3, float *out_lo3, float *out_hi3)
{
	mat_affine m;
	memcpy(&m, rows12, 48);
	Granite::AABB in(vec3(lo3[0], lo3[1], lo3[2]), vec3(hi3[0], hi3[1], hi3[2])), out;
	Granite::SIMD::transform_aabb(out, in, m);
	memcpy(out_lo3, &out.get_minimum(), 12);
	memcpy(out_hi3, &out.get_maximum(), 12);
}

// math/simd.hpp:34-60: 1 = visible
int ref_frustum_cull(const float *lo3, const float *hi3, const float *planes24)
{
	Granite::AABB box(vec3(lo3[0], lo3[1], lo3[2]), vec3(hi3[0], hi3[1], hi3[2]));
	vec4 planes[6];
	memcpy(planes, planes24, sizeof(planes));
	return Granite::SIMD::frustum_cull(box, planes) ? 1 : 0;
}

// renderer/post/hdr.cpp:580-593 compute_rec709_to_st2020, with the reference's compute_xyz_matrix
// (math/transforms.cpp:352-370) and muglm's mat3 inverse / product: column-major mat3 out.
void ref_rec709_to_display_primaries(const float *primaries8, float *out9)
{
	Granite::Primaries rec709 = { vec2(0.640f, 0.330f), vec2(0.3f, 0.6f), vec2(0.150f, 0.060f), vec2(0.3127f, 0.3290f) };
	Granite::Primaries disp = { vec2(primaries8[0], primaries8[1]), vec2(primaries8[2], primaries8[3]), vec2(primaries8[4], primaries8[5]),
	                            vec2(primaries8[6], primaries8[7]) };
	const mat3 srgb_to_xyz = Granite::compute_xyz_matrix(rec709);
	const mat3 xyz_to_display = inverse(Granite::compute_xyz_matrix(disp));
	const mat3 m = xyz_to_display * srgb_to_xyz;
	memcpy(out9, &m, 36);
}
// renderer/lights/clusterer.cpp:467-474: the statements of gather_bindless_spot_shadow_renderables, on the reference's math
void ref_spot_shadow_transform(const float *direction3, const float *position3, float inv_radius, float xy_range, float *out16)
{
	vec3 direction(direction3[0], direction3[1], direction3[2]), position(position3[0], position3[1], position3[2]);
	float range = tan(xy_range);
	mat4 view = mat4_cast(Granite::look_at_arbitrary_up(direction)) * translate(-position);
	mat4 proj = Granite::projection(range * 2.0f, 1.0f, 0.005f / inv_radius, 1.0f / inv_radius);
	mat4 shadow = translate(vec3(0.5f, 0.5f, 0.0f)) * scale(vec3(0.5f, 0.5f, 1.0f)) * proj * view;
	memcpy(out16, &shadow, 64);
}

// renderer/lights/clusterer.cpp:518-521 (gather_bindless_point_shadow_renderables)
void ref_point_shadow_transform(const float *position3, float inv_radius, float *out4)
{
	mat4 view, proj;
	Granite::compute_cube_render_transform(vec3(position3[0], position3[1], position3[2]), 0, proj, view, 0.005f / inv_radius, 1.0f / inv_radius);
	vec4 r(proj[2].zw(), proj[3].zw());
	memcpy(out4, &r, 16);
}
}
