/*
 * oracle_host.c -- TEST INFRASTRUCTURE ONLY.  Host-side math and light preparation of the
 * reference, restated in plain C.  See oracle_math.h for the arithmetic contract.
 */
#include "oracle.h"
#include "oracle_math.h"

#include <float.h>

/* math/muglm/muglm.cpp:319-345 perspective(): reverse-Z, Y-flipped, infinite far when
 * far == InfiniteFarPlane (= FLT_MAX, math/muglm/matrix_helper.hpp:44). */
void orc_perspective(float fovy, float aspect, float z_near, float z_far, float *m)
{
	float tan_half = tanf(fovy / 2.0f);
	for (int i = 0; i < 16; i++)
		m[i] = 0.0f;
	m[0] = 1.0f / (aspect * tan_half);
	m[5] = 1.0f / (tan_half);
	if (z_far == FLT_MAX)
		m[14] = z_near; /* result[3][2] */
	else
	{
		m[10] = -1.0f - z_far / (z_near - z_far);
		m[14] = -(z_far * z_near) / (z_near - z_far);
	}
	m[11] = -1.0f; /* result[2][3] */
	m[1] *= -1.0f;
	m[5] *= -1.0f;
	m[9] *= -1.0f;
	m[13] *= -1.0f;
}

/* math/muglm/muglm_impl.hpp:609-627: a*b = (a*b[0], a*b[1], a*b[2], a*b[3]),
 * a*v = ((a0*v.x + a1*v.y) + a2*v.z) + a3*v.w */
void orc_mat4_mul(const float *a, const float *b, float *out)
{
	float r[16];
	for (int c = 0; c < 4; c++)
	{
		vec4 v = m4_mul_v4(a, v4(b[c * 4 + 0], b[c * 4 + 1], b[c * 4 + 2], b[c * 4 + 3]));
		r[c * 4 + 0] = v.x; r[c * 4 + 1] = v.y; r[c * 4 + 2] = v.z; r[c * 4 + 3] = v.w;
	}
	memcpy(out, r, sizeof(r));
}

/* math/muglm/muglm.cpp:146-200 inverse(mat4): cofactor expansion, same association order. */
#define M(c, r) m[(c) * 4 + (r)]
void orc_mat4_inverse(const float *m, float *out)
{
	float c00 = M(2, 2) * M(3, 3) - M(3, 2) * M(2, 3);
	float c02 = M(1, 2) * M(3, 3) - M(3, 2) * M(1, 3);
	float c03 = M(1, 2) * M(2, 3) - M(2, 2) * M(1, 3);
	float c04 = M(2, 1) * M(3, 3) - M(3, 1) * M(2, 3);
	float c06 = M(1, 1) * M(3, 3) - M(3, 1) * M(1, 3);
	float c07 = M(1, 1) * M(2, 3) - M(2, 1) * M(1, 3);
	float c08 = M(2, 1) * M(3, 2) - M(3, 1) * M(2, 2);
	float c10 = M(1, 1) * M(3, 2) - M(3, 1) * M(1, 2);
	float c11 = M(1, 1) * M(2, 2) - M(2, 1) * M(1, 2);
	float c12 = M(2, 0) * M(3, 3) - M(3, 0) * M(2, 3);
	float c14 = M(1, 0) * M(3, 3) - M(3, 0) * M(1, 3);
	float c15 = M(1, 0) * M(2, 3) - M(2, 0) * M(1, 3);
	float c16 = M(2, 0) * M(3, 2) - M(3, 0) * M(2, 2);
	float c18 = M(1, 0) * M(3, 2) - M(3, 0) * M(1, 2);
	float c19 = M(1, 0) * M(2, 2) - M(2, 0) * M(1, 2);
	float c20 = M(2, 0) * M(3, 1) - M(3, 0) * M(2, 1);
	float c22 = M(1, 0) * M(3, 1) - M(3, 0) * M(1, 1);
	float c23 = M(1, 0) * M(2, 1) - M(2, 0) * M(1, 1);

	float fac0[4] = { c00, c00, c02, c03 };
	float fac1[4] = { c04, c04, c06, c07 };
	float fac2[4] = { c08, c08, c10, c11 };
	float fac3[4] = { c12, c12, c14, c15 };
	float fac4[4] = { c16, c16, c18, c19 };
	float fac5[4] = { c20, c20, c22, c23 };
	float vec0[4] = { M(1, 0), M(0, 0), M(0, 0), M(0, 0) };
	float vec1[4] = { M(1, 1), M(0, 1), M(0, 1), M(0, 1) };
	float vec2_[4] = { M(1, 2), M(0, 2), M(0, 2), M(0, 2) };
	float vec3_[4] = { M(1, 3), M(0, 3), M(0, 3), M(0, 3) };
	static const float sign_a[4] = { +1, -1, +1, -1 };
	static const float sign_b[4] = { -1, +1, -1, +1 };
	float inv[16];
	for (int i = 0; i < 4; i++)
	{
		float i0 = vec1[i] * fac0[i] - vec2_[i] * fac1[i] + vec3_[i] * fac2[i];
		float i1 = vec0[i] * fac0[i] - vec2_[i] * fac3[i] + vec3_[i] * fac4[i];
		float i2 = vec0[i] * fac1[i] - vec1[i] * fac3[i] + vec3_[i] * fac5[i];
		float i3 = vec0[i] * fac2[i] - vec1[i] * fac4[i] + vec2_[i] * fac5[i];
		inv[0 * 4 + i] = i0 * sign_a[i];
		inv[1 * 4 + i] = i1 * sign_b[i];
		inv[2 * 4 + i] = i2 * sign_a[i];
		inv[3 * 4 + i] = i3 * sign_b[i];
	}
	float d0x = M(0, 0) * inv[0 * 4 + 0];
	float d0y = M(0, 1) * inv[1 * 4 + 0];
	float d0z = M(0, 2) * inv[2 * 4 + 0];
	float d0w = M(0, 3) * inv[3 * 4 + 0];
	float dot1 = (d0x + d0y) + (d0z + d0w);
	float ood = 1.0f / dot1;
	for (int i = 0; i < 16; i++)
		out[i] = inv[i] * ood;
}
#undef M

/* math/muglm/muglm_impl.hpp:860-907 floatToHalf(float): rounds half UP on magnitude (not RNE). */
uint16_t orc_float_to_half(float v)
{
	int i = (int)f_bits(v);
	int s = (i >> 16) & 0x00008000;
	int e = ((i >> 23) & 0x000000ff) - (127 - 15);
	int m = i & 0x007fffff;

	if (e <= 0)
	{
		if (e < -10)
			return (uint16_t)s;
		m = (m | 0x00800000) >> (1 - e);
		if (m & 0x00001000)
			m += 0x00002000;
		return (uint16_t)(s | (m >> 13));
	}
	else if (e == 0xff - (127 - 15))
	{
		if (m == 0)
			return (uint16_t)(s | 0x7c00);
		m >>= 13;
		return (uint16_t)(s | 0x7c00 | m | (m == 0));
	}
	else
	{
		if (m & 0x00001000)
		{
			m += 0x00002000;
			if (m & 0x00800000)
			{
				m = 0;
				e += 1;
			}
		}
		if (e > 30)
			return (uint16_t)(s | 0x7c00);
		return (uint16_t)(s | (e << 10) | (m >> 13));
	}
}

/* renderer/render_context.cpp:54-87 RenderContext::set_camera */
void orc_camera_setup(const float *projection, const float *view, orc_camera_t *c)
{
	memcpy(c->projection, projection, 64);
	memcpy(c->view, view, 64);
	orc_mat4_mul(projection, view, c->view_projection);
	orc_mat4_inverse(projection, c->inv_projection);
	orc_mat4_inverse(view, c->inv_view);
	orc_mat4_inverse(c->view_projection, c->inv_view_projection);
	for (int i = 0; i < 3; i++)
	{
		c->camera_position[i] = c->inv_view[12 + i];
		c->camera_front[i] = -c->inv_view[8 + i];
	}
	/* mat2 inv_zw(inv_projection[2].zw, inv_projection[3].zw); project(zw) = -zw.x / zw.y */
	const float *ip = c->inv_projection;
	float a = ip[2 * 4 + 2], b = ip[2 * 4 + 3]; /* column 0 of inv_zw */
	float cc = ip[3 * 4 + 2], d = ip[3 * 4 + 3]; /* column 1 */
	int infinite_z = c->inv_view_projection[15] == 0.0f;
	{
		float zx = a * 1.0f + cc * 1.0f, zy = b * 1.0f + d * 1.0f;
		c->z_near = -zx / zy;
	}
	{
		float f = infinite_z ? 1e-10f : 0.0f;
		float zx = a * f + cc * 1.0f, zy = b * f + d * 1.0f;
		c->z_far = -zx / zy;
	}
}

/* renderer/lights/lights.cpp:63-70 recompute_range + :203-220 PointLight::get_shader_info
 * with a unit-scale node transform (scale_factor == 1). */
void orc_point_light_info(const float *color, const float *position, float cutoff_range, orc_light_t *out)
{
	const float target_atten = 0.1f;
	float max_color = f_max(f_max(color[0], color[1]), color[2]);
	float falloff_range = sqrtf(max_color / target_atten);
	float max_range = f_min(falloff_range, cutoff_range) * 1.0f;
	memset(out, 0, sizeof(*out));
	for (int i = 0; i < 3; i++)
	{
		out->color[i] = color[i] * (1.0f * 1.0f);
		out->position[i] = position[i];
	}
	out->spot_scale_bias[0] = 0;
	out->spot_scale_bias[1] = 0;
	out->offset_radius[0] = orc_float_to_half(0.0f);
	out->offset_radius[1] = orc_float_to_half(max_range);
	/* transform.get_forward() of an identity-rotation node: (0,0,-1). "This shouldn't matter". */
	out->direction[0] = -0.0f; out->direction[1] = -0.0f; out->direction[2] = -1.0f; /* -(row.z) of identity */
	out->inv_radius = 1.0f / max_range;
}

/* renderer/lights/lights.cpp:71-146: set_spot_parameters, set_range, build_model_matrix,
 * SpotLight::get_shader_info.  rot_cols9 = orthonormal node rotation, column-major 3x3. */
void orc_spot_light_info(const float *color, const float *position, const float *rot,
                         float inner_cone_, float outer_cone_, float cutoff_range,
                         orc_light_t *out, float *model_rows)
{
	float inner_cone = f_clamp(inner_cone_, 0.001f, 1.0f);
	float outer_cone = f_clamp(outer_cone_, 0.001f, 1.0f);
	const float target_atten = 0.1f;
	float max_color = f_max(f_max(color[0], color[1]), color[2]);
	float falloff_range = sqrtf(max_color / target_atten);
	float max_range0 = f_min(falloff_range, cutoff_range);
	float xy_range = sqrtf(1.0f - outer_cone * outer_cone) / outer_cone;

	/* build_model_matrix: transform * scale_affine(xy_range*R, xy_range*R, R) */
	float sx = xy_range * max_range0, sy = xy_range * max_range0, sz = max_range0;
	for (int r = 0; r < 3; r++)
	{
		model_rows[r * 4 + 0] = rot[0 * 3 + r] * sx;
		model_rows[r * 4 + 1] = rot[1 * 3 + r] * sy;
		model_rows[r * 4 + 2] = rot[2 * 3 + r] * sz;
		model_rows[r * 4 + 3] = position[r];
	}

	/* transform.get_uniform_scale() == length(vec[0].xyz()): ROW 0 of the node transform
	 * (math/muglm/muglm.cpp:434-437); ~1 for a rotation, but not bit-exactly 1. */
	float scale_factor = sqrtf(rot[0] * rot[0] + rot[3] * rot[3] + rot[6] * rot[6]);
	float max_range = max_range0 * scale_factor;
	float spot_scale = 1.0f / f_max(0.001f, inner_cone - outer_cone);
	float spot_bias = -outer_cone * spot_scale;
	float tan2 = (1.0f - outer_cone * outer_cone) / (outer_cone * outer_cone);
	float center_distance = ((tan2 + 1.0f) * max_range) * 0.5f;
	float spot_offset, spot_radius;
	if (center_distance < max_range)
	{
		spot_offset = center_distance;
		spot_radius = center_distance;
	}
	else
	{
		spot_offset = max_range;
		spot_radius = sqrtf(tan2) * max_range;
	}
	memset(out, 0, sizeof(*out));
	for (int i = 0; i < 3; i++)
	{
		out->color[i] = color[i] * (scale_factor * scale_factor);
		out->position[i] = position[i];
	}
	out->spot_scale_bias[0] = orc_float_to_half(spot_scale);
	out->spot_scale_bias[1] = orc_float_to_half(spot_bias);
	out->offset_radius[0] = orc_float_to_half(spot_offset);
	out->offset_radius[1] = orc_float_to_half(spot_radius);
	/* normalize(transform.get_forward()), forward = -column 2 */
	vec3 fwd = v3_normalize(v3(-rot[6], -rot[7], -rot[8]));
	out->direction[0] = fwd.x; out->direction[1] = fwd.y; out->direction[2] = fwd.z;
	out->inv_radius = 1.0f / max_range;
}

/* renderer/lights/clusterer.cpp:700-703 get_z_slice_extent */
static float z_slice_extent(const orc_camera_t *cam, int res_z)
{
	return f_min(0.5f, cam->z_far / (float)res_z);
}

/* renderer/lights/clusterer.cpp:803-826 */
void orc_cluster_params(const orc_camera_t *cam, int num_lights, int res_x, int res_y, int res_z,
                        orc_cluster_params_t *p)
{
	memset(p, 0, sizeof(*p));
	p->num_lights = num_lights;
	p->num_lights_32 = (num_lights + 31) / 32;
	p->clip_scale[0] = cam->projection[0];
	p->clip_scale[1] = -cam->projection[5];
	p->clip_scale[2] = cam->inv_projection[0];
	p->clip_scale[3] = -cam->inv_projection[5];
	/* translate(.5,.5,0) * scale(.5,.5,1) * view_projection */
	float t[16] = { 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0.5f, 0.5f, 0.0f, 1 };
	float s[16] = { 0.5f, 0, 0, 0, 0, 0.5f, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1 };
	float ts[16];
	orc_mat4_mul(t, s, ts);
	orc_mat4_mul(ts, cam->view_projection, p->transform);
	for (int i = 0; i < 3; i++)
	{
		p->camera_front[i] = cam->camera_front[i];
		p->camera_base[i] = cam->camera_position[i];
	}
	p->xy_scale[0] = (float)res_x;
	p->xy_scale[1] = (float)res_y;
	p->resolution_xy[0] = res_x;
	p->resolution_xy[1] = res_y;
	p->inv_resolution_xy[0] = 1.0f / (float)res_x;
	p->inv_resolution_xy[1] = 1.0f / (float)res_y;
	p->z_scale = 1.0f / z_slice_extent(cam, res_z);
	p->z_max_index = res_z - 1;
}

/* renderer/lights/clusterer.cpp:1265-1275 compute_uint_range */
static void compute_uint_range(float lo, float hi, float extent, int res_z, uint32_t *out)
{
	lo = lo / extent;
	hi = hi / extent;
	if (hi < 0.0f)
	{
		out[0] = 0xffffffffu;
		out[1] = 0u;
		return;
	}
	lo = f_max(lo, 0.0f);
	/* uvec2(vec2): C++ float -> uint32 conversion (truncation). Values here are < 2^32. */
	uint32_t ux = (uint32_t)lo;
	uint32_t uy = (uint32_t)hi;
	if (uy > (uint32_t)(res_z - 1))
		uy = (uint32_t)(res_z - 1);
	out[0] = ux;
	out[1] = uy;
}

/* renderer/lights/clusterer.cpp:1322-1346 + renderer/lights/lights.cpp:330-370 */
void orc_light_z_ranges(const orc_camera_t *cam, const orc_light_t *lights, const float *model_rows,
                        const uint32_t *type_mask, int num_lights, int res_z, uint32_t *z_ranges)
{
	float extent = z_slice_extent(cam, res_z);
	vec3 pos = v3(cam->camera_position[0], cam->camera_position[1], cam->camera_position[2]);
	vec3 front = v3(cam->camera_front[0], cam->camera_front[1], cam->camera_front[2]);
	for (int i = 0; i < num_lights; i++)
	{
		float lo, hi;
		if (type_mask[i >> 5] & (1u << (i & 31)))
		{
			/* point_light_z_range(center, 1/inv_radius) */
			vec3 c = v3(lights[i].position[0], lights[i].position[1], lights[i].position[2]);
			float radius = 1.0f / lights[i].inv_radius;
			float z = v3_dot(v3_sub(c, pos), front);
			lo = z - radius;
			hi = z + radius;
		}
		else
		{
			const float *m = model_rows + (size_t)i * 12;
			vec3 base = v3(m[3], m[7], m[11]);
			vec3 x_off = v3(m[0], m[4], m[8]);
			vec3 y_off = v3(m[1], m[5], m[9]);
			vec3 z_off = v3(-m[2], -m[6], -m[10]);
			vec3 z_base = v3_add(base, z_off);
			vec3 wp[5];
			wp[0] = base;
			wp[1] = v3_add(v3_add(z_base, x_off), y_off);
			wp[2] = v3_add(v3_sub(z_base, x_off), y_off);
			wp[3] = v3_sub(v3_add(z_base, x_off), y_off);
			wp[4] = v3_sub(v3_sub(z_base, x_off), y_off);
			lo = INFINITY;
			hi = -INFINITY;
			for (int k = 0; k < 5; k++)
			{
				float z = v3_dot(v3_sub(wp[k], pos), front);
				lo = f_min(z, lo);
				hi = f_max(z, hi);
			}
		}
		compute_uint_range(lo, hi, extent, res_z, z_ranges + 2 * (size_t)i);
	}
	if (num_lights == 0)
	{
		z_ranges[0] = 0xffffffffu;
		z_ranges[1] = 0u;
	}
}

uint32_t orc_pack_r11g11b10(float r, float g, float b) { return pack_r11g11b10(v3(r, g, b)); }
void orc_unpack_r11g11b10(uint32_t p, float *rgb) { vec3 c = unpack_r11g11b10(p); rgb[0] = c.x; rgb[1] = c.y; rgb[2] = c.z; }
uint16_t orc_f32_to_f16(float f) { return f32_to_f16_rne(f); }
float orc_f16_to_f32(uint16_t h) { return f16_to_f32(h); }
uint32_t orc_linear_to_srgb8(float c) { return linear_to_srgb8(c); }
float orc_srgb8_to_linear(uint32_t v) { return srgb8_to_linear(v); }

/* ---- visibility of positional lights (renderer/scene.cpp:333-358) -------------------------------
 * A light is gathered when the world-space AABB of its static AABB passes the visibility frustum.
 * math/frustum.cpp:109-156 Frustum::build_planes(inv_view_projection): six planes (left, right, near,
 * far, top, bottom) through the unprojected corners, far plane dropped for an infinite projection,
 * each flipped to face the frustum centre.  planes24 = 6 x vec4. */
static void unproject3(const float *m, float x, float y, float z, float *out3)
{
	/* inv_view_projection * vec4(x, y, z, 1): muglm mat4 * vec4 = ((c0*x + c1*y) + c2*z) + c3*w  (math/muglm/muglm_impl.hpp) */
	float r[4];
	for (int i = 0; i < 4; i++)
		r[i] = m[i] * x + m[4 + i] * y + m[8 + i] * z + m[12 + i] * 1.0f;
	out3[0] = r[0] / r[3];
	out3[1] = r[1] / r[3];
	out3[2] = r[2] / r[3];
}

static void cross3(const float *a, const float *b, float *o)
{
	o[0] = a[1] * b[2] - a[2] * b[1];
	o[1] = a[2] * b[0] - a[0] * b[2];
	o[2] = a[0] * b[1] - a[1] * b[0];
}

static void plane_normal(const float *p0, const float *p1, const float *q0, const float *q1, float *n)
{
	/* normalize(cross(p0 - p1, q0 - q1)), normalize(v) = v * (1 / sqrt(dot(v, v))) */
	float a[3] = { p0[0] - p1[0], p0[1] - p1[1], p0[2] - p1[2] };
	float b[3] = { q0[0] - q1[0], q0[1] - q1[1], q0[2] - q1[2] };
	float c[3];
	cross3(a, b, c);
	float inv = 1.0f / sqrtf(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
	n[0] = c[0] * inv;
	n[1] = c[1] * inv;
	n[2] = c[2] * inv;
}

void orc_frustum_planes(const float *inv_vp, float *planes24)
{
	const int infinite_z = inv_vp[15] == 0.0f;
	const float far_z = infinite_z ? 1e-10f : 0.0f; /* FarClipInfiniteClamp, frustum.cpp:62 */
	float TLN[3], BLN[3], BLF[3], TRN[3], TRF[3], BRN[3], BRF[3];
	unproject3(inv_vp, -1.0f, -1.0f, 1.0f, TLN);
	unproject3(inv_vp, -1.0f, +1.0f, 1.0f, BLN);
	unproject3(inv_vp, -1.0f, +1.0f, far_z, BLF);
	unproject3(inv_vp, +1.0f, -1.0f, 1.0f, TRN);
	unproject3(inv_vp, +1.0f, -1.0f, far_z, TRF);
	unproject3(inv_vp, +1.0f, +1.0f, 1.0f, BRN);
	unproject3(inv_vp, +1.0f, +1.0f, far_z, BRF);
	float center[4];
	for (int i = 0; i < 4; i++)
		center[i] = inv_vp[i] * 0.0f + inv_vp[4 + i] * 0.0f + inv_vp[8 + i] * 0.5f + inv_vp[12 + i] * 1.0f;
	float l[3], r[3], n[3], f[3], t[3], b[3];
	plane_normal(BLF, BLN, TLN, BLN, l);
	plane_normal(TRF, TRN, BRN, TRN, r);
	plane_normal(BLN, BRN, TRN, BRN, n);
	plane_normal(TRF, BRF, BLF, BRF, f);
	plane_normal(TLN, TRN, TRF, TRN, t);
	plane_normal(BRF, BRN, BLN, BRN, b);
	const float *normals[6] = { l, r, n, f, t, b };
	const float *points[6] = { BLN, TRN, BRN, BRF, TRN, BRN };
	for (int i = 0; i < 6; i++)
	{
		float *p = planes24 + 4 * i;
		const float *nn = normals[i], *pt = points[i];
		p[0] = nn[0];
		p[1] = nn[1];
		p[2] = nn[2];
		p[3] = -(nn[0] * pt[0] + nn[1] * pt[1] + nn[2] * pt[2]);
		if (i == 3 && infinite_z)
			p[0] = p[1] = p[2] = p[3] = 0.0f;
		/* winding: dot(center, p) < 0 => p = -p  (vec4 dot: (x + y) ... muglm dot(vec4) = a.x*b.x + a.y*b.y + a.z*b.z + a.w*b.w) */
		float d = center[0] * p[0] + center[1] * p[1] + center[2] * p[2] + center[3] * p[3];
		if (d < 0.0f)
		{
			p[0] = -p[0];
			p[1] = -p[1];
			p[2] = -p[2];
			p[3] = -p[3];
		}
	}
}

/* math/simd.hpp:386-419 SIMD::transform_aabb(AABB&, const AABB&, const mat_affine&): per output
 * component c: t[c] + M[c][0] * pick(x) + M[c][1] * pick(y) + M[c][2] * pick(z), added in that order;
 * pick = hi where the matrix element is > 0 (for the maximum), lo otherwise (and vice versa). */
void orc_transform_aabb(const float *rows12, const float *lo3, const float *hi3, float *out_lo3, float *out_hi3)
{
	for (int c = 0; c < 3; c++)
	{
		const float *row = rows12 + 4 * c;
		float hi = row[3], lo = row[3];
		for (int k = 0; k < 3; k++)
		{
			const int pos = row[k] > 0.0f;
			hi = hi + row[k] * (pos ? hi3[k] : lo3[k]);
			lo = lo + row[k] * (pos ? lo3[k] : hi3[k]);
		}
		out_hi3[c] = hi;
		out_lo3[c] = lo;
	}
}

/* math/simd.hpp:34-60 SIMD::frustum_cull: per plane, the corner furthest along the plane normal
 * (hi where the plane component is > 0, w = 1); the products are summed pairwise ((x + y) + (z + w),
 * two horizontal adds) and the box is visible when no sum has its sign bit set. */
int orc_frustum_cull(const float *lo3, const float *hi3, const float *planes24)
{
	for (int i = 0; i < 6; i++)
	{
		const float *p = planes24 + 4 * i;
		float d[4];
		for (int k = 0; k < 3; k++)
			d[k] = p[k] * (p[k] > 0.0f ? hi3[k] : lo3[k]);
		d[3] = p[3] * 1.0f;
		float s = (d[0] + d[1]) + (d[2] + d[3]);
		if (signbit(s))
			return 0;
	}
	return 1;
}

/* Static AABB of a light (renderer/lights/lights.cpp:77-89 SpotLight::set_range, :196-201
 * PointLight::set_range), moved to world space by the node transform and tested: 1 = gathered. */
int orc_light_visible(const float *planes24, int is_point, const float *color3, float cutoff_range, float outer_cone, const float *rows12)
{
	const float target_atten = 0.1f;
	float max_color = f_max(f_max(color3[0], color3[1]), color3[2]);
	float falloff_range = sqrtf(max_color / target_atten);
	float max_range = f_min(falloff_range, cutoff_range);
	float lo[3], hi[3];
	if (is_point)
	{
		lo[0] = lo[1] = lo[2] = -max_range;
		hi[0] = hi[1] = hi[2] = max_range;
	}
	else
	{
		float oc = f_clamp(outer_cone, 0.001f, 1.0f);
		float xy = sqrtf(1.0f - oc * oc) / oc;
		xy *= max_range;
		lo[0] = lo[1] = -xy;
		lo[2] = -max_range;
		hi[0] = hi[1] = xy;
		hi[2] = 0.0f;
	}
	float wlo[3], whi[3];
	orc_transform_aabb(rows12, lo, hi, wlo, whi);
	return orc_frustum_cull(wlo, whi, planes24);
}

/* ---- HDR10 output: Rec.709 -> display primaries (hdr.cpp:580-593, math/transforms.cpp:352-370) ----
 * Evaluated in double precision and rounded once: the matrix is a per-swapchain constant, and the
 * reference's fp32 chain (two 3x3 inverses) agrees with it to a few ulps (tests/test_oracle_cpu.py). */
static void xyz_matrix_d(const float *p8, double m[9])
{
	double prim[4][3];
	for (int i = 0; i < 4; i++)
	{
		double x = p8[2 * i], y = p8[2 * i + 1];
		prim[i][0] = x / y; /* convert_primary */
		prim[i][1] = 1.0;
		prim[i][2] = (1.0 - x - y) / y;
	}
	/* component_scale = inverse([r g b]) * white  (Cramer) */
	const double *r = prim[0], *g = prim[1], *b = prim[2], *wt = prim[3];
	double det = r[0] * (g[1] * b[2] - b[1] * g[2]) - g[0] * (r[1] * b[2] - b[1] * r[2]) + b[0] * (r[1] * g[2] - g[1] * r[2]);
	double sx = (wt[0] * (g[1] * b[2] - b[1] * g[2]) - g[0] * (wt[1] * b[2] - b[1] * wt[2]) + b[0] * (wt[1] * g[2] - g[1] * wt[2])) / det;
	double sy = (r[0] * (wt[1] * b[2] - b[1] * wt[2]) - wt[0] * (r[1] * b[2] - b[1] * r[2]) + b[0] * (r[1] * wt[2] - wt[1] * r[2])) / det;
	double sz = (r[0] * (g[1] * wt[2] - wt[1] * g[2]) - g[0] * (r[1] * wt[2] - wt[1] * r[2]) + wt[0] * (r[1] * g[2] - g[1] * r[2])) / det;
	for (int i = 0; i < 3; i++)
	{
		m[0 + i] = r[i] * sx; /* column 0 */
		m[3 + i] = g[i] * sy;
		m[6 + i] = b[i] * sz;
	}
}

static void inverse3_d(const double a[9], double o[9])
{
	/* column-major a[c*3+r] */
	double c00 = a[4] * a[8] - a[7] * a[5], c01 = a[7] * a[2] - a[1] * a[8], c02 = a[1] * a[5] - a[4] * a[2];
	double det = a[0] * c00 + a[3] * c01 + a[6] * c02;
	o[0] = c00 / det; o[1] = c01 / det; o[2] = c02 / det;
	o[3] = (a[6] * a[5] - a[3] * a[8]) / det; o[4] = (a[0] * a[8] - a[6] * a[2]) / det; o[5] = (a[3] * a[2] - a[0] * a[5]) / det;
	o[6] = (a[3] * a[7] - a[6] * a[4]) / det; o[7] = (a[6] * a[1] - a[0] * a[7]) / det; o[8] = (a[0] * a[4] - a[3] * a[1]) / det;
}

void orc_rec709_to_display_primaries(const float *primaries8, float *out9)
{
	static const float rec709[8] = { 0.640f, 0.330f, 0.3f, 0.6f, 0.150f, 0.060f, 0.3127f, 0.3290f }; /* hdr.cpp:585-589 */
	double src[9], dst[9], inv[9];
	xyz_matrix_d(rec709, src);
	xyz_matrix_d(primaries8, dst);
	inverse3_d(dst, inv);
	for (int c = 0; c < 3; c++)
		for (int r = 0; r < 3; r++)
			out9[c * 3 + r] = (float)(inv[0 * 3 + r] * src[c * 3 + 0] + inv[1 * 3 + r] * src[c * 3 + 1] + inv[2 * 3 + r] * src[c * 3 + 2]);
}

/* ---- shadow transforms of the positional lights: ClustererBindlessTransforms::shadow[index] ---- */
/* math/transforms.cpp:122-148 rotate_vector (muglm normalize = v * (1 / sqrt(dot))), :180-183 look_at_arbitrary_up;
 * quaternion as (w, x, y, z) */
static vec3 v3_cross(vec3 a, vec3 b)
{
	return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}

static vec3 normalize_muglm(vec3 v)
{
	float inv = 1.0f / sqrtf(v3_dot(v, v));
	return v3(v.x * inv, v.y * inv, v.z * inv);
}

static void rotate_vector(vec3 from, vec3 to, float *q)
{
	from = normalize_muglm(from);
	to = normalize_muglm(to);
	float cos_angle = v3_dot(from, to);
	if (fabsf(cos_angle) > 0.9999f)
	{
		if (cos_angle > 0.9999f)
		{
			q[0] = 1.0f; q[1] = q[2] = q[3] = 0.0f;
			return;
		}
		vec3 rotation = v3_cross(v3(1.0f, 0.0f, 0.0f), from);
		if (v3_dot(rotation, rotation) > 0.001f)
			rotation = normalize_muglm(rotation);
		else
			rotation = normalize_muglm(v3_cross(v3(0.0f, 1.0f, 0.0f), from));
		q[0] = 0.0f; q[1] = rotation.x; q[2] = rotation.y; q[3] = rotation.z;
		return;
	}
	vec3 rotation = normalize_muglm(v3_cross(from, to));
	vec3 half_vector = normalize_muglm(v3_add(from, to));
	float cos_half_range = f_clamp(v3_dot(half_vector, from), 0.0f, 1.0f);
	float sin_half_angle = sqrtf(1.0f - cos_half_range * cos_half_range);
	q[0] = cos_half_range; q[1] = rotation.x * sin_half_angle; q[2] = rotation.y * sin_half_angle; q[3] = rotation.z * sin_half_angle;
}

/* math/muglm/muglm.cpp:29-62 mat3_cast / mat4_cast, column-major 4x4 */
static void mat4_cast(const float *q, float *m)
{
	float w = q[0], x = q[1], y = q[2], z = q[3];
	float qxx = x * x, qyy = y * y, qzz = z * z, qxz = x * z, qxy = x * y, qyz = y * z, qwx = w * x, qwy = w * y, qwz = w * z;
	for (int i = 0; i < 16; i++)
		m[i] = 0.0f;
	m[0] = 1.0f - 2.0f * (qyy + qzz); m[1] = 2.0f * (qxy + qwz); m[2] = 2.0f * (qxz - qwy);
	m[4] = 2.0f * (qxy - qwz); m[5] = 1.0f - 2.0f * (qxx + qzz); m[6] = 2.0f * (qyz + qwx);
	m[8] = 2.0f * (qxz + qwy); m[9] = 2.0f * (qyz - qwx); m[10] = 1.0f - 2.0f * (qxx + qyy);
	m[15] = 1.0f;
}

static void mat4_translate(float x, float y, float z, float *m)
{
	for (int i = 0; i < 16; i++)
		m[i] = (i % 5) == 0 ? 1.0f : 0.0f;
	m[12] = x; m[13] = y; m[14] = z;
}

static void mat4_scale(float x, float y, float z, float *m)
{
	for (int i = 0; i < 16; i++)
		m[i] = 0.0f;
	m[0] = x; m[5] = y; m[10] = z; m[15] = 1.0f;
}

/* renderer/lights/clusterer.cpp:467-474 (gather_bindless_spot_shadow_renderables) */
void orc_spot_shadow_transform(const orc_light_t *light, float xy_range, float *out16)
{
	float range = tanf(xy_range);
	float q[4], rot[16], tr[16], view[16], proj[16], bias_t[16], bias_s[16], a[16], b[16];
	vec3 dir = normalize_muglm(v3(light->direction[0], light->direction[1], light->direction[2]));
	rotate_vector(dir, v3(0.0f, 0.0f, -1.0f), q);
	mat4_cast(q, rot);
	mat4_translate(-light->position[0], -light->position[1], -light->position[2], tr);
	orc_mat4_mul(rot, tr, view);
	orc_perspective(range * 2.0f, 1.0f, 0.005f / light->inv_radius, 1.0f / light->inv_radius, proj);
	mat4_translate(0.5f, 0.5f, 0.0f, bias_t);
	mat4_scale(0.5f, 0.5f, 1.0f, bias_s);
	orc_mat4_mul(bias_t, bias_s, a); /* left to right: ((T * S) * proj) * view */
	orc_mat4_mul(a, proj, b);
	orc_mat4_mul(b, view, out16);
}

/* renderer/lights/clusterer.cpp:518-521 with math/transforms.cpp:223-224: column 0 = (proj[2].zw, proj[3].zw) */
void orc_point_shadow_transform(const orc_light_t *light, float *out16)
{
	float flip[16], p[16], proj[16];
	mat4_scale(-1.0f, 1.0f, 1.0f, flip);
	orc_perspective(0.5f * 3.1415926535897932384626433832795f, 1.0f, 0.005f / light->inv_radius, 1.0f / light->inv_radius, p);
	orc_mat4_mul(flip, p, proj);
	for (int i = 0; i < 16; i++)
		out16[i] = 0.0f;
	out16[0] = proj[10]; out16[1] = proj[11]; out16[2] = proj[14]; out16[3] = proj[15];
}

/* ---- volumetric decals: host side of the decal binning (clusterer.cpp:1348-1412) ---- */
/* mvp = view_projection * to_mat4(world) (clusterer.cpp:1408-1409); world: 3 rows of 4 (mat_affine) */
void orc_decal_mvp(const float *view_projection16, const float *world_rows12, float *out16)
{
	float w[16];
	for (int c = 0; c < 4; c++)
	{
		w[4 * c + 0] = world_rows12[c];
		w[4 * c + 1] = world_rows12[4 + c];
		w[4 * c + 2] = world_rows12[8 + c];
		w[4 * c + 3] = c == 3 ? 1.0f : 0.0f;
	}
	orc_mat4_mul(view_projection16, w, out16);
}

/* decal_z_range (clusterer.cpp:1348-1369): view-depth range of the unit cube's eight corners, corner i of an AABB =
 * (i & 1 ? hi : lo, i & 2 ? hi : lo, i & 4 ? hi : lo) (math/aabb.cpp get_corner) */
void orc_decal_z_range(const orc_camera_t *cam, const float *world_rows12, float *lo_hi2)
{
	float lo = INFINITY, hi = -INFINITY;
	for (int i = 0; i < 8; i++)
	{
		const float t[4] = { (i & 1) ? 0.5f : -0.5f, (i & 2) ? 0.5f : -0.5f, (i & 4) ? 0.5f : -0.5f, 1.0f };
		float wpos[3];
		for (int r = 0; r < 3; r++)
		{
			const float *row = world_rows12 + 4 * r;
			/* SIMD::mul(vec4, mat_affine, vec4) (math/simd.hpp:107-119): one DPPS per row = (p0 + p1) + (p2 + p3) */
			wpos[r] = (row[0] * t[0] + row[1] * t[1]) + (row[2] * t[2] + row[3] * t[3]);
		}
		const float z = (wpos[0] - cam->camera_position[0]) * cam->camera_front[0] + (wpos[1] - cam->camera_position[1]) * cam->camera_front[1] +
		                (wpos[2] - cam->camera_position[2]) * cam->camera_front[2];
		lo = z < lo ? z : lo;
		hi = z > hi ? z : hi;
	}
	lo_hi2[0] = lo;
	lo_hi2[1] = hi;
}
