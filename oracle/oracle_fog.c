/*
 * oracle_fog.c -- TEST INFRASTRUCTURE ONLY.  Volumetric fog, second pass: fog_accumulate.comp (VolumetricFog::build_fog,
 * renderer/lights/volumetric_fog.cpp:236-254) integrates the froxel grid of in-scattered light and extinction front to
 * back along every view ray.  light: R16G16B16A16_SFLOAT, w x h x d (rgb = in-scattered light, a = optical depth of the
 * froxel); fog: same size and format (rgb = light accumulated up to and including the slice, a = transmittance).
 * The first pass (fog_light_density.comp) is not restated: DESIGN.md section 0, row f4.
 *
 * Per slice a 17-tap blur of the density grid (the centre, the four edge and four corner neighbours in the slice, and the
 * same eight neighbours -- not the centre -- in the slice in front), read through a NearestClamp sampler at texel centres
 * with integer offsets = exact fetches with the coordinate clamped to the edge, then
 *   light += back.rgb * (exp2(-depth) * back.a);  depth += back.a;  store (light, exp2(-depth)).
 * fp32, the shader's order of operations, weights as glslang folds them (double, rounded once).
 */
#include "oracle.h"
#include "oracle_math.h"

typedef struct
{
	const uint16_t *p;
	int w, h, d;
} vol16f;

static inline vec4 vol_texel(vol16f v, int x, int y, int z)
{
	x = x < 0 ? 0 : (x > v.w - 1 ? v.w - 1 : x);
	y = y < 0 ? 0 : (y > v.h - 1 ? v.h - 1 : y);
	z = z < 0 ? 0 : (z > v.d - 1 ? v.d - 1 : z);
	const uint16_t *t = v.p + 4 * (((size_t)z * v.h + y) * v.w + x);
	return v4(f16_to_f32(t[0]), f16_to_f32(t[1]), f16_to_f32(t[2]), f16_to_f32(t[3]));
}

/* fog_accumulate.comp:27-63 */
void orc_fog_accumulate(const uint16_t *light_rgba16f, int w, int h, int d, uint16_t *fog_rgba16f)
{
	const vol16f light = { light_rgba16f, w, h, d };
	const float inv_x = 1.0f / (float)w, inv_y = 1.0f / (float)h, inv_z = 1.0f / (float)d; /* volumetric_fog.cpp:245-247 */
	const float w3 = (float)(1.0 / (1.375 * 32.0)), w2 = (float)(1.0 / (1.375 * 16.0)), w1 = (float)(1.0 / (1.375 * 8.0)), w0 = (float)(1.0 / (1.375 * 4.0));
	static const int taps[17][3] = { { 0, 0, 0 },   { 0, -1, -1 }, { -1, 0, -1 }, { 1, 0, -1 }, { 0, 1, -1 }, { -1, -1, -1 }, { 1, -1, -1 }, { -1, 1, -1 }, { 1, 1, -1 },
		                         { 0, -1, 0 },  { -1, 0, 0 },  { 1, 0, 0 },   { 0, 1, 0 },  { 1, -1, 0 }, { -1, -1, 0 },  { -1, 1, 0 },  { 1, 1, 0 } };
	const float weights[17] = { w0, w2, w2, w2, w2, w3, w3, w3, w3, w1, w1, w1, w1, w2, w2, w2, w2 };
#pragma omp parallel for schedule(static)
	for (int y = 0; y < h; y++)
		for (int x = 0; x < w; x++)
		{
			/* the sampler's nearest texel of ((x + 0.5) / w, (y + 0.5) / h, (z + 0.5) / d) */
			const int sx = (int)floorf(((float)x + 0.5f) * inv_x * (float)w), sy = (int)floorf(((float)y + 0.5f) * inv_y * (float)h);
			vec4 front = v4(0.0f, 0.0f, 0.0f, 0.0f);
			for (int z = 0; z < d; z++)
			{
				const int sz = (int)floorf(((float)z + 0.5f) * inv_z * (float)d);
				vec4 back = v4(0.0f, 0.0f, 0.0f, 0.0f);
				for (int k = 0; k < 17; k++)
				{
					const vec4 t = vol_texel(light, sx + taps[k][0], sy + taps[k][1], sz + taps[k][2]);
					back.x += weights[k] * t.x;
					back.y += weights[k] * t.y;
					back.z += weights[k] * t.z;
					back.w += weights[k] * t.w;
				}
				/* accumulate_scattering, :17-22 */
				const float s = exp2f(-front.w) * back.w;
				front = v4(front.x + back.x * s, front.y + back.y * s, front.z + back.z * s, front.w + back.w);
				uint16_t *o = fog_rgba16f + 4 * (((size_t)z * h + y) * w + x);
				o[0] = f32_to_f16_rne(front.x);
				o[1] = f32_to_f16_rne(front.y);
				o[2] = f32_to_f16_rne(front.z);
				o[3] = f32_to_f16_rne(exp2f(-front.w));
			}
		}
}
