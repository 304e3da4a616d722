/*
 * oracle_cluster.c -- TEST INFRASTRUCTURE ONLY.  Clusterer kernels K1..K4 restated from the
 * reference GLSL (assets/shaders/lights/clusterer_bindless_*.comp).  Integer outputs
 * (bitmask, ranges) are the bit-exact contract.  See oracle_math.h for arithmetic rules.
 */
#include "oracle.h"

#include <stdlib.h>
#include "oracle_math.h"

/* ---- K1: clusterer_bindless_spot_transform.comp:33-73 ---- */
void orc_spot_transform(const orc_camera_t *cam, const float *model_rows, int num_lights, float *out)
{
	vec3 cpos = v3(cam->camera_position[0], cam->camera_position[1], cam->camera_position[2]);
	vec3 cfront = v3(cam->camera_front[0], cam->camera_front[1], cam->camera_front[2]);
	for (int index = 0; index < num_lights; index++)
	{
		const float *m = model_rows + (size_t)index * 12;
		vec3 p[5];
		p[0] = v3(m[3], m[7], m[11]);                       /* mat_affine_get_translation */
		vec3 pz = v3_add(p[0], v3(-m[2], -m[6], -m[10]));   /* + get_forward */
		vec3 right = v3(m[0], m[4], m[8]);
		vec3 up = v3(m[1], m[5], m[9]);
		p[1] = v3_add(v3_add(pz, right), up);
		p[2] = v3_add(v3_sub(pz, right), up);
		p[3] = v3_sub(v3_sub(pz, right), up);
		p[4] = v3_sub(v3_add(pz, right), up);

		float z[5];
		for (int i = 0; i < 5; i++)
			z[i] = v3_dot(v3_sub(p[i], cpos), cfront);
		float z_lo = z[0], z_hi = z[0];
		for (int i = 1; i < 5; i++)
		{
			z_lo = f_min(z_lo, z[i]);
			z_hi = f_max(z_hi, z[i]);
		}
		float cull;
		if (z_lo <= cam->z_near && z_hi >= cam->z_far)
			cull = 0.0f;
		else if (z_lo <= cam->z_near)
			cull = -1.0f;
		else
			cull = 1.0f;

		float *o = out + (size_t)index * 24;
		for (int i = 0; i < 5; i++)
		{
			vec4 c = m4_mul_v4(cam->view_projection, v4(p[i].x, p[i].y, p[i].z, 1.0f));
			o[i * 4 + 0] = c.x; o[i * 4 + 1] = c.y; o[i * 4 + 2] = c.z; o[i * 4 + 3] = c.w;
		}
		o[20] = cull; o[21] = z_lo; o[22] = z_hi; o[23] = 0.0f;
	}
}

/* ---- K2: clusterer_bindless_setup.comp ---- */

/* :45-69 project_sphere_flat */
static vec2 project_sphere_flat(float view_xy, float view_z, float radius)
{
	float len = v2_length(v2(view_xy, view_z));
	float sin_xy = radius / len;
	vec2 result;
	if (sin_xy < 0.999f)
	{
		float cos_xy = sqrtf(1.0f - sin_xy * sin_xy);
		/* mat2(cos, sin, -sin, cos) * (xy, z) = (cos*xy + (-sin)*z, sin*xy + cos*z) */
		vec2 rot_lo = v2(cos_xy * view_xy + (-sin_xy) * view_z, sin_xy * view_xy + cos_xy * view_z);
		/* mat2(cos, -sin, +sin, cos) * (xy, z) */
		vec2 rot_hi = v2(cos_xy * view_xy + sin_xy * view_z, (-sin_xy) * view_xy + cos_xy * view_z);
		if (rot_lo.y <= 0.0f)
			rot_lo = v2(-1.0f, 0.0f);
		if (rot_hi.y <= 0.0f)
			rot_hi = v2(+1.0f, 0.0f);
		result = v2(rot_lo.x / rot_lo.y, rot_hi.x / rot_hi.y);
	}
	else
		result = v2(-INFINITY, +INFINITY);
	return result;
}

typedef struct { vec2 c[3]; } tri2;
typedef struct { vec3 c[3]; } tri3;

static vec3 v3_mix3(vec3 a, vec3 b, float t) { return v3_mixf(a, b, t); }

/* :71-90 clip_*_output (vec3 / mat3x2 overloads, clip on .z) */
static void clip_single_z(tri2 *out, vec3 c0, vec3 c1, vec3 c2, float target)
{
	float la = (target - c0.z) / (c2.z - c0.z);
	float lb = (target - c1.z) / (c2.z - c1.z);
	c0 = v3_mix3(c0, c2, la);
	c1 = v3_mix3(c1, c2, lb);
	out->c[0] = v2(c0.x, c0.y); out->c[1] = v2(c1.x, c1.y); out->c[2] = v2(c2.x, c2.y);
}

static void clip_dual_z(tri2 *o0, tri2 *o1, vec3 c0, vec3 c1, vec3 c2, float target)
{
	float l_ab = (target - c0.z) / (c1.z - c0.z);
	float l_ac = (target - c0.z) / (c2.z - c0.z);
	vec3 ab = v3_mix3(c0, c1, l_ab);
	vec3 ac = v3_mix3(c0, c2, l_ac);
	o0->c[0] = v2(ab.x, ab.y); o0->c[1] = v2(c1.x, c1.y); o0->c[2] = v2(ac.x, ac.y);
	o1->c[0] = v2(ac.x, ac.y); o1->c[1] = v2(c1.x, c1.y); o1->c[2] = v2(c2.x, c2.y);
}

static vec3 v4_xyz_div(vec4 c, float d) { return v3(c.x / d, c.y / d, c.z / d); }

/* :92-111 vec4 / mat3 overloads, clip on .w */
static void clip_single_w(tri3 *out, vec4 c0, vec4 c1, vec4 c2, float target)
{
	float la = (target - c0.w) / (c2.w - c0.w);
	float lb = (target - c1.w) / (c2.w - c1.w);
	c0 = v4_mixf(c0, c2, la);
	c1 = v4_mixf(c1, c2, lb);
	out->c[0] = v4_xyz_div(c0, target);
	out->c[1] = v4_xyz_div(c1, target);
	out->c[2] = v4_xyz_div(c2, c2.w);
}

static void clip_dual_w(tri3 *o0, tri3 *o1, vec4 c0, vec4 c1, vec4 c2, float target)
{
	float l_ab = (target - c0.w) / (c1.w - c0.w);
	float l_ac = (target - c0.w) / (c2.w - c0.w);
	vec4 ab = v4_mixf(c0, c1, l_ab);
	vec4 ac = v4_mixf(c0, c2, l_ac);
	o0->c[0] = v4_xyz_div(ab, target); o0->c[1] = v4_xyz_div(c1, c1.w); o0->c[2] = v4_xyz_div(ac, target);
	o1->c[0] = v4_xyz_div(ac, target); o1->c[1] = v4_xyz_div(c1, c1.w); o1->c[2] = v4_xyz_div(c2, c2.w);
}

static float cross_2d(vec2 a, vec2 b) { return a.x * b.y - a.y * b.x; }
static vec2 v2_sub(vec2 a, vec2 b) { return v2(a.x - b.x, a.y - b.y); }
static vec2 v2_neg(vec2 a) { return v2(-a.x, -a.y); }

#define MAX_TRIANGLES 8u

/* :118-145 setup_triangle(mat3x2) */
static void setup_triangle2(uint32_t *num_triangles, tri2 t, float cull, float *data)
{
	vec2 c0 = t.c[0], c1 = t.c[1], c2 = t.c[2];
	vec2 ab = v2_sub(c1, c0);
	vec2 bc = v2_sub(c2, c1);
	vec2 ca = v2_sub(c0, c2);
	float z = cross_2d(ab, v2_neg(ca));
	if (fabsf(z) < 0.000001f || f_sign(cull) == f_sign(z))
		return;
	float inv_z = 1.0f / z;
	vec3 base = v3(inv_z * cross_2d(ab, v2_neg(c0)), inv_z * cross_2d(bc, v2_neg(c1)), inv_z * cross_2d(ca, v2_neg(c2)));
	vec3 dx = v3(inv_z * -ab.y, inv_z * -bc.y, inv_z * -ca.y);
	vec3 dy = v3(inv_z * ab.x, inv_z * bc.x, inv_z * ca.x);
	if (*num_triangles < MAX_TRIANGLES)
	{
		float *d = data + 16u * *num_triangles;
		d[0] = base.x; d[1] = base.y; d[2] = base.z; d[3] = 0.0f;
		d[4] = dx.x; d[5] = dx.y; d[6] = dx.z; d[7] = z;
		d[8] = dy.x; d[9] = dy.y; d[10] = dy.z; d[11] = inv_z;
		d[12] = f_min(f_min(c0.x, c1.x), c2.x);
		d[13] = f_min(f_min(c0.y, c1.y), c2.y);
		d[14] = f_max(f_max(c0.x, c1.x), c2.x);
		d[15] = f_max(f_max(c0.y, c1.y), c2.y);
	}
	(*num_triangles)++;
}

/* :147-198 setup_triangle(mat3): clip against z < 0 */
static void setup_triangle3(uint32_t *num_triangles, tri3 t, float cull, float *data)
{
	vec3 c0 = t.c[0], c1 = t.c[1], c2 = t.c[2];
	uint32_t clip_code = (uint32_t)(c0.z < 0.0f) + (uint32_t)(c1.z < 0.0f) * 2u + (uint32_t)(c2.z < 0.0f) * 4u;
	tri2 k0, k1;
	int dual = 0;
	switch (clip_code)
	{
	case 0:
		k0.c[0] = v2(c0.x, c0.y); k0.c[1] = v2(c1.x, c1.y); k0.c[2] = v2(c2.x, c2.y);
		break;
	case 1: clip_dual_z(&k0, &k1, c0, c1, c2, 0.0f); dual = 1; break;
	case 2: clip_dual_z(&k0, &k1, c1, c2, c0, 0.0f); dual = 1; break;
	case 4: clip_dual_z(&k0, &k1, c2, c0, c1, 0.0f); dual = 1; break;
	case 3: clip_single_z(&k0, c0, c1, c2, 0.0f); break;
	case 5: clip_single_z(&k0, c2, c0, c1, 0.0f); break;
	case 6: clip_single_z(&k0, c1, c2, c0, 0.0f); break;
	default: return;
	}
	setup_triangle2(num_triangles, k0, cull, data);
	if (dual)
		setup_triangle2(num_triangles, k1, cull, data);
}

/* :200-250 setup_triangle(vec4 x3): clip against w < MIN_W */
static void setup_triangle4(uint32_t *num_triangles, vec4 c0, vec4 c1, vec4 c2, float cull, float *data)
{
	const float MIN_W = 1.0f / 1024.0f;
	uint32_t clip_code = (uint32_t)(c0.w < MIN_W) + (uint32_t)(c1.w < MIN_W) * 2u + (uint32_t)(c2.w < MIN_W) * 4u;
	tri3 k0, k1;
	int dual = 0;
	switch (clip_code)
	{
	case 0:
		k0.c[0] = v4_xyz_div(c0, c0.w); k0.c[1] = v4_xyz_div(c1, c1.w); k0.c[2] = v4_xyz_div(c2, c2.w);
		break;
	case 1: clip_dual_w(&k0, &k1, c0, c1, c2, MIN_W); dual = 1; break;
	case 2: clip_dual_w(&k0, &k1, c1, c2, c0, MIN_W); dual = 1; break;
	case 4: clip_dual_w(&k0, &k1, c2, c0, c1, MIN_W); dual = 1; break;
	case 3: clip_single_w(&k0, c0, c1, c2, MIN_W); break;
	case 5: clip_single_w(&k0, c2, c0, c1, MIN_W); break;
	case 6: clip_single_w(&k0, c1, c2, c0, MIN_W); break;
	default: return;
	}
	setup_triangle3(num_triangles, k0, cull, data);
	if (dual)
		setup_triangle3(num_triangles, k1, cull, data);
}

static vec4 ld4(const float *p) { return v4(p[0], p[1], p[2], p[3]); }

/* :252-322 main */
void orc_cull_setup(const orc_camera_t *cam, const orc_cluster_params_t *p, const orc_light_t *lights,
                    const uint32_t *type_mask, const float *spots, float *cull_setup)
{
	for (int index = 0; index < p->num_lights; index++)
	{
		float *data = cull_setup + (size_t)index * 128;
		int point = (type_mask[index >> 5] & (1u << (index & 31))) != 0u;
		if (point)
		{
			vec4 pos = v4(lights[index].position[0], lights[index].position[1], lights[index].position[2], 1.0f);
			float radius = 1.0f / lights[index].inv_radius;
			vec4 vw = m4_mul_v4(cam->view, pos);
			vec3 view = v3(vw.x, -vw.y, -vw.z);

			vec2 rx = project_sphere_flat(view.x, view.z, radius);
			vec2 ry = project_sphere_flat(view.y, view.z, radius);
			vec4 ranges = v4(rx.x, rx.y, ry.x, ry.y);

			float xy_length = v2_length(v2(view.x, view.y));
			float ct[4]; /* mat2 columns: (ct0,ct1), (ct2,ct3) */
			if (xy_length < 0.00001f)
			{
				ct[0] = 1.0f; ct[1] = 0.0f; ct[2] = 0.0f; ct[3] = 1.0f;
			}
			else
			{
				float inv_xy_length = 1.0f / xy_length;
				ct[0] = view.x * inv_xy_length;
				ct[1] = -view.y * inv_xy_length;
				ct[2] = view.y * inv_xy_length;
				ct[3] = view.x * inv_xy_length;
			}
			vec2 txy = v2(ct[0] * view.x + ct[2] * view.y, ct[1] * view.x + ct[3] * view.y);
			vec2 tx = project_sphere_flat(txy.x, view.z, radius);
			vec2 ty = project_sphere_flat(txy.y, view.z, radius);
			vec4 tr = v4(tx.x, tx.y, ty.x, ty.y);
			int ellipsis = !isinf(tr.x) && !isinf(tr.y) && !isinf(tr.z) && !isinf(tr.w);
			vec2 center = v2((tr.x + tr.y) * 0.5f, (tr.z + tr.w) * 0.5f);
			vec2 ellipse_radius = v2(tr.y - center.x, tr.w - center.y);

			/* ranges * clip_scale.xxyy, stored as ranges.xzyw */
			ranges = v4(ranges.x * p->clip_scale[0], ranges.y * p->clip_scale[0],
			            ranges.z * p->clip_scale[1], ranges.w * p->clip_scale[1]);
			data[0] = ranges.x; data[1] = ranges.z; data[2] = ranges.y; data[3] = ranges.w;
			data[4] = tr.x; data[5] = tr.y; data[6] = tr.z; data[7] = tr.w;
			data[8] = ct[0]; data[9] = ct[1]; data[10] = ct[2]; data[11] = ct[3];
			data[12] = ellipsis ? 1.0f : 0.0f;
			data[13] = 1.0f / ellipse_radius.x;
			data[14] = 1.0f / ellipse_radius.y;
			data[15] = 0.0f;
		}
		else
		{
			const float *s = spots + (size_t)index * 24;
			float cull = s[20];
			if (cull != 0.0f)
			{
				uint32_t num_triangles = 0u;
				vec4 c0 = ld4(s), c1 = ld4(s + 4), c2 = ld4(s + 8), c3 = ld4(s + 12), c4 = ld4(s + 16);
				setup_triangle4(&num_triangles, c0, c1, c2, cull, data);
				setup_triangle4(&num_triangles, c0, c2, c3, cull, data);
				setup_triangle4(&num_triangles, c0, c3, c4, cull, data);
				setup_triangle4(&num_triangles, c0, c4, c1, cull, data);
				setup_triangle4(&num_triangles, c2, c1, c3, cull, data);
				setup_triangle4(&num_triangles, c4, c3, c1, cull, data);
				data[3] = bits_f(num_triangles);
			}
			else
				data[3] = bits_f(0xffffffffu);
		}
	}
}

/* ---- K3: clusterer_bindless_binning.comp ---- */

/* :38-85 test_point_light */
static int test_point_light(const orc_cluster_params_t *p, vec2 uv, vec2 uv_stride, const float *d)
{
	if (d[12] != 0.0f) /* ellipsis_inv_radius.x */
	{
		vec2 ic = v2(0.5f * (d[4] + d[5]), 0.5f * (d[6] + d[7]));
		vec2 clip_lo = uv;
		vec2 clip_hi = v2(uv.x + uv_stride.x, uv.y + uv_stride.y);
		clip_lo = v2(clip_lo.x * p->clip_scale[2], clip_lo.y * p->clip_scale[3]);
		clip_hi = v2(clip_hi.x * p->clip_scale[2], clip_hi.y * p->clip_scale[3]);
		float m0 = d[8], m1 = d[9], m2 = d[10], m3 = d[11];
		float irx = d[13], iry = d[14];
#define XF(px, py, ox, oy) \
		do { ox = (m0 * (px) + m2 * (py)) - ic.x; oy = (m1 * (px) + m3 * (py)) - ic.y; ox *= irx; oy *= iry; } while (0)
		float d00x, d00y, d01x, d01y, d10x, d10y, d11x, d11y;
		XF(clip_lo.x, clip_lo.y, d00x, d00y);
		XF(clip_lo.x, clip_hi.y, d01x, d01y);
		XF(clip_hi.x, clip_lo.y, d10x, d10y);
		XF(clip_hi.x, clip_hi.y, d11x, d11y);
#undef XF
		float max_diag = f_max(v2_length(v2(d00x - d11x, d00y - d11y)), v2_length(v2(d01x - d10x, d01y - d10y)));
		float min_sq_dist = 1.0f + max_diag;
		min_sq_dist *= min_sq_dist;
		float q0 = d00x * d00x + d00y * d00y;
		float q1 = d01x * d01x + d01y * d01y;
		float q2 = d10x * d10x + d10y * d10y;
		float q3 = d11x * d11x + d11y * d11y;
		return q0 < min_sq_dist && q1 < min_sq_dist && q2 < min_sq_dist && q3 < min_sq_dist;
	}
	return (uv.x + uv_stride.x > d[0]) && (uv.y + uv_stride.y > d[1]) && (uv.x < d[2]) && (uv.y < d[3]);
}

/* :87-119 test_spot_light */
static int test_spot_light(vec2 uv, vec2 uv_stride, const float *d)
{
	uint32_t num_triangles = f_bits(d[3]);
	if (num_triangles > MAX_TRIANGLES)
		return 1;
	for (uint32_t i = 0; i < num_triangles; i++)
	{
		const float *t = d + 16u * i;
		if ((uv.x + uv_stride.x > t[12]) && (uv.y + uv_stride.y > t[13]) && (uv.x < t[14]) && (uv.y < t[15]))
		{
			float base[3], ok = 1;
			for (int k = 0; k < 3; k++)
			{
				float dxk = t[4 + k], dyk = t[8 + k];
				float b = t[k];
				b += dxk * uv.x;
				b += dyk * uv.y;
				/* mix(vec3(0), stride*dx, greaterThan(dx, 0)) : bvec mix selects */
				b += dxk > 0.0f ? uv_stride.x * dxk : 0.0f;
				b += dyk > 0.0f ? uv_stride.y * dyk : 0.0f;
				base[k] = b;
				if (!(b > 0.0f))
					ok = 0;
			}
			(void)base;
			if (ok)
				return 1;
		}
	}
	return 0;
}

/* :125-179 main, SUBGROUPS=1 with gl_SubgroupSize == 32: a coarse conservative test over the
 * 8x4-tile block decides which of the chunk's 32 lights get the per-tile fine test.  Bits for
 * light indices >= num_lights read uninitialised cull data in the reference; defined 0 here. */
void orc_binning(const orc_cluster_params_t *p, const uint32_t *type_mask, const float *cull_setup, uint32_t *bitmask)
{
	const int res_x = p->resolution_xy[0], res_y = p->resolution_xy[1];
	const int n32 = p->num_lights_32;
#pragma omp parallel for schedule(dynamic, 1)
	for (int by = 0; by < res_y / 4; by++)
	{
		for (int bx = 0; bx < res_x / 8; bx++)
		{
			for (int chunk = 0; chunk < n32; chunk++)
			{
				/* tile_uv = 2.0 * vec2(tile * TILE_SIZE) * inv_resolution_xy - 1.0 */
				vec2 tile_uv = v2(2.0f * (float)(bx * 8) * p->inv_resolution_xy[0] - 1.0f,
				                  2.0f * (float)(by * 4) * p->inv_resolution_xy[1] - 1.0f);
				vec2 tile_stride = v2((2.0f * 8.0f) * p->inv_resolution_xy[0], (2.0f * 4.0f) * p->inv_resolution_xy[1]);
				uint32_t tm = type_mask[chunk];
				uint32_t ballot = 0;
				for (int lane = 0; lane < 32; lane++)
				{
					int light = chunk * 32 + lane;
					if (light >= p->num_lights)
						continue;
					const float *d = cull_setup + (size_t)light * 128;
					int passed = (tm & (1u << lane)) ? test_point_light(p, tile_uv, tile_stride, d)
					                                 : test_spot_light(tile_uv, tile_stride, d);
					if (passed)
						ballot |= 1u << lane;
				}
				for (int ty = 0; ty < 4; ty++)
				{
					for (int tx = 0; tx < 8; tx++)
					{
						int px = bx * 8 + tx, py = by * 4 + ty;
						vec2 uv = v2(2.0f * (float)px * p->inv_resolution_xy[0] - 1.0f,
						             2.0f * (float)py * p->inv_resolution_xy[1] - 1.0f);
						vec2 stride = v2(2.0f * p->inv_resolution_xy[0], 2.0f * p->inv_resolution_xy[1]);
						uint32_t mask = 0, b = ballot;
						while (b)
						{
							int lsb = __builtin_ctz(b);
							b &= ~(1u << lsb);
							const float *d = cull_setup + (size_t)(chunk * 32 + lsb) * 128;
							int passed = (tm & (1u << lsb)) ? test_point_light(p, uv, stride, d)
							                                : test_spot_light(uv, stride, d);
							if (passed)
								mask |= 1u << lsb;
						}
						bitmask[((size_t)py * res_x + px) * n32 + chunk] = mask;
					}
				}
			}
		}
	}
}

/* ---- K4: clusterer_bindless_z_range.comp:20-51 (naive form = the specification; the _opt
 * variant computes the same function) ---- */
void orc_z_range(const uint32_t *z_ranges, int num_ranges, int res_z, uint32_t *cluster_range)
{
#pragma omp parallel for
	for (int zi = 0; zi < res_z; zi++)
	{
		uint32_t z = (uint32_t)zi;
		uint32_t z_lo = 0xffffffffu, z_hi = 0u;
		for (uint32_t i = 0; i < (uint32_t)num_ranges; i++)
		{
			if (z >= z_ranges[2 * i] && z <= z_ranges[2 * i + 1])
			{
				z_lo = i;
				break;
			}
		}
		int z_lo_int = (int)z_lo > 0 ? (int)z_lo : 0;
		for (int i = num_ranges - 1; i >= z_lo_int; i--)
		{
			if (z >= z_ranges[2 * i] && z <= z_ranges[2 * i + 1])
			{
				z_hi = (uint32_t)i;
				break;
			}
		}
		cluster_range[2 * zi] = z_lo;
		cluster_range[2 * zi + 1] = z_hi;
	}
}

/* ---------------------------------------------------------------------------------------------
 * Volumetric-decal binning: clusterer_bindless_binning_decal.comp (SUBGROUPS = 0 path, :118-141) over the same
 * (resolution_x x resolution_y) tile grid as the lights, one bit per decal and tile.  Host side
 * LightClusterer::update_bindless_mask_buffer_decal_gpu (clusterer.cpp:1391-1461): mvp[i] = view_projection * world.
 * --------------------------------------------------------------------------------------------- */
/* compute_decal_screen_bb, :39-70.  mat4 * vec4 as the generated reference code evaluates it (GLM: pairwise);
 * min / max keep the accumulated value when the projected coordinate is NaN (GLM's (y < x) ? y : x). */
void orc_decal_screen_bb(const float *m, float *bb4)
{
	float bb[4] = { 1.0f, 1.0f, -1.0f, -1.0f };
	float lo_w = 1.0f, hi_w = -1.0f;
	float corner[8][4];
	for (int r = 0; r < 4; r++)
		corner[0][r] = (m[r] * -0.5f + m[4 + r] * -0.5f) + (m[8 + r] * -0.5f + m[12 + r] * 1.0f);
	for (int r = 0; r < 4; r++)
	{
		corner[1][r] = corner[0][r] + m[r];
		corner[2][r] = corner[0][r] + m[4 + r];
		corner[3][r] = corner[1][r] + m[4 + r];
		corner[4][r] = corner[0][r] + m[8 + r];
		corner[5][r] = corner[1][r] + m[8 + r];
		corner[6][r] = corner[2][r] + m[8 + r];
		corner[7][r] = corner[3][r] + m[8 + r];
	}
	for (int i = 0; i < 8; i++)
	{
		const float w = corner[i][3];
		lo_w = w < lo_w ? w : lo_w;
		hi_w = hi_w < w ? w : hi_w;
		const float px = corner[i][0] / w, py = corner[i][1] / w;
		bb[0] = px < bb[0] ? px : bb[0];
		bb[1] = py < bb[1] ? py : bb[1];
		bb[2] = bb[2] < px ? px : bb[2];
		bb[3] = bb[3] < py ? py : bb[3];
	}
	if (hi_w <= 0.0f)
		bb[0] = bb[1] = bb[2] = bb[3] = -10.0f;
	else if (lo_w <= 0.0f)
	{
		bb[0] = bb[1] = -1.0f;
		bb[2] = bb[3] = 1.0f;
	}
	for (int i = 0; i < 4; i++)
		bb4[i] = bb[i];
}

/* main(), SUBGROUPS = 0: bitmask[(y * res_x + x) * num_decals_32 + chunk], bit = decal & 31 */
void orc_decal_binning(int res_x, int res_y, const float *inv_resolution_xy2, int num_decals, const float *mvps16, uint32_t *bitmask)
{
	const int n32 = (num_decals + 31) / 32;
	const float stride_x = 2.0f * inv_resolution_xy2[0], stride_y = 2.0f * inv_resolution_xy2[1];
	float *bbs = (float *)malloc(sizeof(float) * 4 * (size_t)(num_decals > 0 ? num_decals : 1));
	for (int i = 0; i < num_decals; i++)
		orc_decal_screen_bb(mvps16 + 16 * (size_t)i, bbs + 4 * (size_t)i);
#pragma omp parallel for schedule(static)
	for (int y = 0; y < res_y; y++)
		for (int x = 0; x < res_x; x++)
		{
			const float u = 2.0f * (float)x * inv_resolution_xy2[0] - 1.0f, v = 2.0f * (float)y * inv_resolution_xy2[1] - 1.0f;
			for (int c = 0; c < n32; c++)
			{
				uint32_t mask = 0u;
				for (int b = 0; b < 32 && 32 * c + b < num_decals; b++)
				{
					const float *bb = bbs + 4 * (size_t)(32 * c + b);
					if (u + stride_x > bb[0] && v + stride_y > bb[1] && u < bb[2] && v < bb[3]) /* test_decal, :28-31 */
						mask |= 1u << b;
				}
				bitmask[((size_t)y * res_x + x) * n32 + c] = mask;
			}
		}
	free(bbs);
}
