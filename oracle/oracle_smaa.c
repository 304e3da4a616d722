/* oracle_smaa.c -- CPU restatement of the reference's SMAA passes.  TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Follows assets/shaders/post/SMAA.hlsl as the reference compiles it (smaa_common.h: SMAA_GLSL_4, presets by
 * SMAA_QUALITY) through smaa_edge_detection.{vert,frag}, smaa_blend_weight.{vert,frag} (SMAA_SUBPIXEL_MODE = 0) and
 * smaa_neighbor_blend.{vert,frag} (SMAA_TARGET_SRGB = 1), as wired by renderer/post/smaa.cpp:32-209.
 * Pinned bit for bit to those shaders executed on the CPU (oracle/ref_post_shim.cpp, KERNEL 150-173;
 * tests/test_oracle_ref_smaa.py).
 *
 * Decisions the reference leaves to the implementation, the same as everywhere else:
 *   - LinearClamp = bilinear with exact fp32 weights, clamp to edge; a sample at the fragment's own normalised
 *     coordinate is a texel fetch (integer offsets added to the texel);
 *   - mad() is GLSL fma(): fused, one rounding (SMAA.hlsl:573);
 *   - round() rounds half away from zero (no tie occurs: the bilinear values it is applied to sit at multiples of
 *     1/4 by construction of the fetch offsets);
 *   - the varyings the vertex stage computes (offsets = texcoord + constants * rt_metrics) are those functions
 *     of the fragment's own coordinate;
 *   - constant sub-expressions are folded in double by glslang and rounded once (the literals below are the ones
 *     in the SPIR-V);
 *   - UNORM8 stores round to nearest, sRGB8 stores use the exact OETF (oracle_math.h). */
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "oracle.h"
#include "oracle_math.h"

typedef struct
{
	const uint8_t *data;
	int w, h, ch; /* channels per texel: 1 (search), 2 (edges, area), 4 (colour, weights) */
} tex8;

static inline vec4 tex8_texel(tex8 t, int x, int y)
{
	x = x < 0 ? 0 : (x > t.w - 1 ? t.w - 1 : x);
	y = y < 0 ? 0 : (y > t.h - 1 ? t.h - 1 : y);
	const uint8_t *p = t.data + ((size_t)y * t.w + x) * t.ch;
	vec4 r = v4(0.0f, 0.0f, 0.0f, 1.0f);
	r.x = (float)p[0] / 255.0f;
	if (t.ch > 1)
		r.y = (float)p[1] / 255.0f;
	if (t.ch > 2)
	{
		r.z = (float)p[2] / 255.0f;
		r.w = (float)p[3] / 255.0f;
	}
	return r;
}

/* The fragment a sample belongs to: its normalised coordinate and its texel. */
typedef struct
{
	float u, v;
	int x, y;
} frag_t;

/* textureLod / texture / textureLodOffset on a LinearClamp sampler.  `same_size` textures (edges, colour, weights)
 * are fetched when sampled at the fragment's own coordinate. */
static inline vec4 tex8_sample(tex8 t, float u, float v, int ox, int oy, const frag_t *f)
{
	if (f && u == f->u && v == f->v)
		return tex8_texel(t, f->x + ox, f->y + oy);
	bilin_t s = bilin_setup(u, v, t.w, t.h);
	vec4 t00 = tex8_texel(t, s.x0 + ox, s.y0 + oy), t10 = tex8_texel(t, s.x1 + ox, s.y0 + oy);
	vec4 t01 = tex8_texel(t, s.x0 + ox, s.y1 + oy), t11 = tex8_texel(t, s.x1 + ox, s.y1 + oy);
	return v4(bilin_mix(t00.x, t10.x, t01.x, t11.x, s.a, s.b), bilin_mix(t00.y, t10.y, t01.y, t11.y, s.a, s.b),
	          bilin_mix(t00.z, t10.z, t01.z, t11.z, s.a, s.b), bilin_mix(t00.w, t10.w, t01.w, t11.w, s.a, s.b));
}

static inline float step_f(float edge, float x) { return x < edge ? 0.0f : 1.0f; }
static inline float round_f(float x) { return roundf(x); } /* half away from zero, like the shim's GLM */
static inline uint32_t unorm8(float c)
{
	c = c > 0.0f ? (c < 1.0f ? c : 1.0f) : 0.0f; /* NaN -> 0 */
	return (uint32_t)floorf(c * 255.0f + 0.5f);
}

/* SMAA.hlsl:304-324 */
typedef struct
{
	float threshold;
	float max_search_steps;
	float max_search_steps_diag; /* 0: SMAA_DISABLE_DIAG_DETECTION */
	int corner_detection;        /* 0: SMAA_DISABLE_CORNER_DETECTION; rounding is 25 % when enabled */
} smaa_preset;

static smaa_preset preset_of(int quality)
{
	static const smaa_preset p[4] = {
		{ 0.15f, 4.0f, 0.0f, 0 },
		{ 0.1f, 8.0f, 0.0f, 0 },
		{ 0.1f, 16.0f, 8.0f, 1 },
		{ 0.05f, 32.0f, 16.0f, 1 },
	};
	return p[quality < 0 ? 0 : (quality > 3 ? 3 : quality)];
}

/* ------------------------------------------------------------------------------------------------ edges */
/* SMAALumaEdgeDetectionPS (SMAA.hlsl:689-746) with SMAAEdgeDetectionVS (:645-650); colour = the sRGB image viewed as
 * UNORM (smaa.cpp:124); output R8G8_UNORM, cleared to 0 where the shader discards. */
void orc_smaa_edge_detection(const uint32_t *color_unorm, int w, int h, int quality, uint8_t *edges_rg8, int y0, int y1)
{
	const smaa_preset P = preset_of(quality);
	const tex8 col = { (const uint8_t *)color_unorm, w, h, 4 };
	const float mx = 1.0f / (float)w, my = 1.0f / (float)h;
#pragma omp parallel for
	for (int y = y0; y < y1; y++)
		for (int x = 0; x < w; x++)
		{
			frag_t f = { ((float)x + 0.5f) * mx, ((float)y + 0.5f) * my, x, y };
			uint8_t *out = edges_rg8 + 2 * ((size_t)y * w + x);
			out[0] = out[1] = 0;
			const vec3 wts = v3(0.2126f, 0.7152f, 0.0722f);
#define LUMA(U, V) ({ vec4 c_ = tex8_sample(col, (U), (V), 0, 0, &f); v3_dot(v3(c_.x, c_.y, c_.z), wts); })
			/* offset[0] = mad(rt.xyxy, (-1, 0, 0, -1), tc.xyxy) etc. */
			const float L = LUMA(f.u, f.v);
			const float Lleft = LUMA(fmaf(mx, -1.0f, f.u), fmaf(my, 0.0f, f.v));
			const float Ltop = LUMA(fmaf(mx, 0.0f, f.u), fmaf(my, -1.0f, f.v));
			float dx = fabsf(L - Lleft), dy = fabsf(L - Ltop);
			float ex = step_f(P.threshold, dx), ey = step_f(P.threshold, dy);
			if (ex * 1.0f + ey * 1.0f == 0.0f)
				continue; /* discard */
			const float Lright = LUMA(fmaf(mx, 1.0f, f.u), fmaf(my, 0.0f, f.v));
			const float Lbottom = LUMA(fmaf(mx, 0.0f, f.u), fmaf(my, 1.0f, f.v));
			float dz = fabsf(L - Lright), dw = fabsf(L - Lbottom);
			float maxx = f_max(dx, dz), maxy = f_max(dy, dw);
			const float Lleftleft = LUMA(fmaf(mx, -2.0f, f.u), fmaf(my, 0.0f, f.v));
			const float Ltoptop = LUMA(fmaf(mx, 0.0f, f.u), fmaf(my, -2.0f, f.v));
			dz = fabsf(Lleft - Lleftleft);
			dw = fabsf(Ltop - Ltoptop);
			maxx = f_max(maxx, dz);
			maxy = f_max(maxy, dw);
			const float final_delta = f_max(maxx, maxy);
			ex *= step_f(final_delta, dx * 2.0f); /* SMAA_LOCAL_CONTRAST_ADAPTATION_FACTOR */
			ey *= step_f(final_delta, dy * 2.0f);
			out[0] = (uint8_t)unorm8(ex);
			out[1] = (uint8_t)unorm8(ey);
#undef LUMA
		}
}

/* ------------------------------------------------------------------------------------------------ weights */
typedef struct
{
	tex8 edges, area, search;
	float mx, my, mz, mw; /* rt_metrics = (1/w, 1/h, w, h) */
	smaa_preset P;
	const frag_t *f;
} wctx;

static inline vec2 decode_diag2(vec2 e)
{
	e.x = e.x * fabsf(5.0f * e.x - 3.75f);
	return v2(round_f(e.x), round_f(e.y));
}

/* SMAASearchDiag1 / 2 (SMAA.hlsl:862-895): returns coord.zw, *e = the last edges sample */
static vec2 search_diag(const wctx *c, float tu, float tv, float dirx, float diry, vec2 *e, int second)
{
	float cx = tu, cy = tv, cz = -1.0f, cw = 1.0f;
	if (second)
		cx += 0.25f * c->mx;
	while (cz < c->P.max_search_steps_diag - 1.0f && cw > 0.9f)
	{
		cx = fmaf(c->mx, dirx, cx);
		cy = fmaf(c->my, diry, cy);
		cz = fmaf(1.0f, 1.0f, cz);
		vec4 s = tex8_sample(c->edges, cx, cy, 0, 0, c->f);
		*e = v2(s.x, s.y);
		if (second)
			*e = decode_diag2(*e);
		cw = e->x * 0.5f + e->y * 0.5f;
	}
	return v2(cz, cw);
}

/* SMAAAreaDiag (SMAA.hlsl:900-914) */
static vec2 area_diag(const wctx *c, vec2 dist, vec2 e, float offset)
{
	float tx = fmaf(20.0f, e.x, dist.x), ty = fmaf(20.0f, e.y, dist.y);
	tx = fmaf(0.0062500000931322574615478515625f, tx, 0.00312500004656612873077392578125f);
	ty = fmaf(0.001785714295692741870880126953125f, ty, 0.0008928571478463709354400634765625f);
	tx += 0.5f;
	ty += 0.14285714924335479736328125f * offset;
	vec4 s = tex8_sample(c->area, tx, ty, 0, 0, NULL);
	return v2(s.x, s.y);
}

/* SMAACalculateDiagWeights (SMAA.hlsl:919-985); texcoord is the fragment's own coordinate */
static vec2 diag_weights(const wctx *c, vec2 e_in, float sub_z, float sub_w)
{
	const float tu = c->f->u, tv = c->f->v;
	vec2 weights = v2(0.0f, 0.0f), end = v2(0.0f, 0.0f);
	float d_x, d_y, d_z, d_w;
	if (e_in.x > 0.0f)
	{
		vec2 r = search_diag(c, tu, tv, -1.0f, 1.0f, &end, 0);
		d_x = r.x;
		d_z = r.y;
		d_x += end.y > 0.9f ? 1.0f : 0.0f;
	}
	else
		d_x = d_z = 0.0f;
	{
		vec2 r = search_diag(c, tu, tv, 1.0f, -1.0f, &end, 0);
		d_y = r.x;
		d_w = r.y;
	}
	if (d_x + d_y > 2.0f)
	{
		const float c0x = fmaf(-d_x + 0.25f, c->mx, tu), c0y = fmaf(d_x, c->my, tv);
		const float c1x = fmaf(d_y, c->mx, tu), c1y = fmaf(-d_y - 0.25f, c->my, tv);
		vec4 a = tex8_sample(c->edges, c0x, c0y, -1, 0, c->f), b = tex8_sample(c->edges, c1x, c1y, 1, 0, c->f);
		/* c.yxwz = SMAADecodeDiagBilinearAccess(c.xyzw): the .x and .z components get the |5x - 3.75| treatment */
		float qx = a.x * fabsf(a.x * 5.0f - 3.75f), qz = b.x * fabsf(b.x * 5.0f - 3.75f);
		float rx = round_f(qx), ry = round_f(a.y), rz = round_f(qz), rw = round_f(b.y);
		/* c = (r.y, r.x, r.w, r.z); cc = mad(2, c.xz, c.yw) */
		float ccx = fmaf(2.0f, ry, rx), ccy = fmaf(2.0f, rw, rz);
		if (step_f(0.9f, d_z) != 0.0f)
			ccx = 0.0f;
		if (step_f(0.9f, d_w) != 0.0f)
			ccy = 0.0f;
		vec2 ar = area_diag(c, v2(d_x, d_y), v2(ccx, ccy), sub_z);
		weights.x += ar.x;
		weights.y += ar.y;
	}
	{
		vec2 r = search_diag(c, tu, tv, -1.0f, -1.0f, &end, 1);
		d_x = r.x;
		d_z = r.y;
	}
	if (tex8_sample(c->edges, tu, tv, 1, 0, c->f).x > 0.0f)
	{
		vec2 r = search_diag(c, tu, tv, 1.0f, 1.0f, &end, 1);
		d_y = r.x;
		d_w = r.y;
		d_y += end.y > 0.9f ? 1.0f : 0.0f;
	}
	else
		d_y = d_w = 0.0f;
	if (d_x + d_y > 2.0f)
	{
		const float c0x = fmaf(-d_x, c->mx, tu), c0y = fmaf(-d_x, c->my, tv);
		const float c1x = fmaf(d_y, c->mx, tu), c1y = fmaf(d_y, c->my, tv);
		float c_x = tex8_sample(c->edges, c0x, c0y, -1, 0, c->f).y;
		float c_y = tex8_sample(c->edges, c0x, c0y, 0, -1, c->f).x;
		vec4 s = tex8_sample(c->edges, c1x, c1y, 1, 0, c->f);
		float c_z = s.y, c_w = s.x;
		float ccx = fmaf(2.0f, c_x, c_y), ccy = fmaf(2.0f, c_z, c_w);
		if (step_f(0.9f, d_z) != 0.0f)
			ccx = 0.0f;
		if (step_f(0.9f, d_w) != 0.0f)
			ccy = 0.0f;
		vec2 ar = area_diag(c, v2(d_x, d_y), v2(ccx, ccy), sub_w);
		weights.x += ar.y;
		weights.y += ar.x;
	}
	return weights;
}

/* SMAASearchLength (SMAA.hlsl:997-1014) */
static float search_length(const wctx *c, float ex, float ey, float offset)
{
	float sx = 33.0f, sy = -33.0f;
	float bx = 66.0f * offset, by = 33.0f * 1.0f;
	sx += -1.0f;
	sy += 1.0f;
	bx += 0.5f;
	by += -0.5f;
	sx *= 0.015625f;
	sy *= 0.0625f;
	bx *= 0.015625f;
	by *= 0.0625f;
	return tex8_sample(c->search, fmaf(sx, ex, bx), fmaf(sy, ey, by), 0, 0, NULL).x;
}

/* SMAASearchXLeft / XRight / YUp / YDown (SMAA.hlsl:1019-1086).  axis 0: x, 1: y; sign -1: towards smaller. */
static float search_axis(const wctx *c, float tu, float tv, float end, int axis, float sign)
{
	float ex = axis ? 1.0f : 0.0f, ey = axis ? 0.0f : 1.0f;
	for (;;)
	{
		const float pos = axis ? tv : tu;
		const int inside = sign < 0.0f ? pos > end : pos < end;
		const float along = axis ? ex : ey, cross = axis ? ey : ex;
		if (!(inside && along > 0.828100025653839111328125f && cross == 0.0f))
			break;
		vec4 s = tex8_sample(c->edges, tu, tv, 0, 0, c->f);
		ex = s.x;
		ey = s.y;
		/* texcoord = mad(+-(2, 0) or +-(0, 2), rt.xy, texcoord) */
		tu = fmaf(axis ? (sign < 0.0f ? -0.0f : 0.0f) : sign * 2.0f, c->mx, tu);
		tv = fmaf(axis ? sign * 2.0f : (sign < 0.0f ? -0.0f : 0.0f), c->my, tv);
	}
	const float len = axis ? search_length(c, ey, ex, sign < 0.0f ? 0.0f : 0.5f) : search_length(c, ex, ey, sign < 0.0f ? 0.0f : 0.5f);
	const float offset = fmaf(-2.007874011993408203125f, len, 3.25f);
	if (axis)
		return fmaf(sign < 0.0f ? c->my : -c->my, offset, tv);
	return fmaf(sign < 0.0f ? c->mx : -c->mx, offset, tu);
}

/* SMAAArea (SMAA.hlsl:1091-1103) */
static vec2 area_ortho(const wctx *c, float dx, float dy, float e1, float e2, float offset)
{
	float tx = fmaf(16.0f, round_f(e1 * 4.0f), dx), ty = fmaf(16.0f, round_f(e2 * 4.0f), dy);
	tx = fmaf(0.0062500000931322574615478515625f, tx, 0.00312500004656612873077392578125f);
	ty = fmaf(0.001785714295692741870880126953125f, ty, 0.0008928571478463709354400634765625f);
	ty = fmaf(0.14285714924335479736328125f, offset, ty);
	vec4 s = tex8_sample(c->area, tx, ty, 0, 0, NULL);
	return v2(s.x, s.y);
}

/* SMAADetectHorizontal / VerticalCornerPattern (SMAA.hlsl:1108-1140); (ax, ay) and (bx, by) are texcoord.xy / .zw */
static void corner_pattern(const wctx *c, float *w0, float *w1, float ax, float ay, float bx, float by, float dx, float dy, int vertical)
{
	if (!c->P.corner_detection)
		return;
	const float lx = step_f(dx, dy), ly = step_f(dy, dx); /* leftRight = step(d.xy, d.yx) */
	float rx = lx * 0.75f, ry = ly * 0.75f;
	const float sum = lx + ly;
	rx /= sum;
	ry /= sum;
	float fx = 1.0f, fy = 1.0f;
	if (!vertical)
	{
		fx -= rx * tex8_sample(c->edges, ax, ay, 0, 1, c->f).x;
		fx -= ry * tex8_sample(c->edges, bx, by, 1, 1, c->f).x;
		fy -= rx * tex8_sample(c->edges, ax, ay, 0, -2, c->f).x;
		fy -= ry * tex8_sample(c->edges, bx, by, 1, -2, c->f).x;
	}
	else
	{
		fx -= rx * tex8_sample(c->edges, ax, ay, 1, 0, c->f).y;
		fx -= ry * tex8_sample(c->edges, bx, by, 1, 1, c->f).y;
		fy -= rx * tex8_sample(c->edges, ax, ay, -2, 0, c->f).y;
		fy -= ry * tex8_sample(c->edges, bx, by, -2, 1, c->f).y;
	}
	*w0 *= f_clamp(fx, 0.0f, 1.0f);
	*w1 *= f_clamp(fy, 0.0f, 1.0f);
}

/* SMAABlendingWeightCalculationPS (SMAA.hlsl:1145-1247) with SMAABlendingWeightCalculationVS (:655-668),
 * subsampleIndices = 0 (SMAA_SUBPIXEL_MODE 0); output R8G8B8A8_UNORM */
void orc_smaa_blend_weights(const uint8_t *edges_rg8, int w, int h, const uint8_t *area_rg8, const uint8_t *search_r8, int quality, uint32_t *weights_rgba8,
                            int y0, int y1)
{
	const smaa_preset P = preset_of(quality);
#pragma omp parallel for
	for (int y = y0; y < y1; y++)
		for (int x = 0; x < w; x++)
		{
			wctx c;
			c.edges = (tex8){ edges_rg8, w, h, 2 };
			c.area = (tex8){ area_rg8, 160, 560, 2 };
			c.search = (tex8){ search_r8, 64, 16, 1 };
			c.mx = 1.0f / (float)w;
			c.my = 1.0f / (float)h;
			c.mz = (float)w;
			c.mw = (float)h;
			c.P = P;
			frag_t f = { ((float)x + 0.5f) * c.mx, ((float)y + 0.5f) * c.my, x, y };
			c.f = &f;
			const float pixx = f.u * c.mz, pixy = f.v * c.mw;
			/* offset[0] = mad(rt.xyxy, (-0.25, -0.125, 1.25, -0.125), tc.xyxy); offset[1] = mad(rt.xyxy, (-0.125, -0.25, -0.125, 1.25), tc.xyxy) */
			const float o0x = fmaf(c.mx, -0.25f, f.u), o0y = fmaf(c.my, -0.125f, f.v), o0z = fmaf(c.mx, 1.25f, f.u), o0w = fmaf(c.my, -0.125f, f.v);
			const float o1x = fmaf(c.mx, -0.125f, f.u), o1y = fmaf(c.my, -0.25f, f.v), o1z = fmaf(c.mx, -0.125f, f.u), o1w = fmaf(c.my, 1.25f, f.v);
			/* offset[2] = mad(rt.xxyy, (-2, 2, -2, 2) * steps, (offset[0].xz, offset[1].yw)) */
			const float o2x = fmaf(c.mx, -2.0f * P.max_search_steps, o0x), o2y = fmaf(c.mx, 2.0f * P.max_search_steps, o0z);
			const float o2z = fmaf(c.my, -2.0f * P.max_search_steps, o1y), o2w = fmaf(c.my, 2.0f * P.max_search_steps, o1w);

			float wx = 0.0f, wy = 0.0f, wz = 0.0f, ww = 0.0f;
			vec4 e4 = tex8_sample(c.edges, f.u, f.v, 0, 0, &f);
			float ex = e4.x, ey = e4.y;
			if (ey > 0.0f)
			{
				int ortho = 1;
				if (P.max_search_steps_diag > 0.0f)
				{
					vec2 dw = diag_weights(&c, v2(ex, ey), 0.0f, 0.0f);
					wx = dw.x;
					wy = dw.y;
					ortho = wx == -wy;
				}
				if (ortho)
				{
					float cx = search_axis(&c, o0x, o0y, o2x, 0, -1.0f);
					float cy = o1y;
					float d_x = cx;
					const float e1 = tex8_sample(c.edges, cx, cy, 0, 0, &f).x;
					const float cz = search_axis(&c, o0z, o0w, o2y, 0, 1.0f);
					float d_y = cz;
					d_x = fabsf(round_f(fmaf(c.mz, d_x, -pixx)));
					d_y = fabsf(round_f(fmaf(c.mz, d_y, -pixx)));
					const float sx = sqrtf(d_x), sy = sqrtf(d_y);
					const float e2 = tex8_sample(c.edges, cz, cy, 1, 0, &f).x;
					vec2 a = area_ortho(&c, sx, sy, e1, e2, 0.0f);
					wx = a.x;
					wy = a.y;
					cy = f.v;
					corner_pattern(&c, &wx, &wy, cx, cy, cz, cy, d_x, d_y, 0);
				}
				else
					ex = 0.0f;
			}
			if (ex > 0.0f)
			{
				const float cy = search_axis(&c, o1x, o1y, o2z, 1, -1.0f);
				float cx = o0x;
				float d_x = cy;
				const float e1 = tex8_sample(c.edges, cx, cy, 0, 0, &f).y;
				const float cz = search_axis(&c, o1z, o1w, o2w, 1, 1.0f);
				float d_y = cz;
				d_x = fabsf(round_f(fmaf(c.mw, d_x, -pixy)));
				d_y = fabsf(round_f(fmaf(c.mw, d_y, -pixy)));
				const float sx = sqrtf(d_x), sy = sqrtf(d_y);
				const float e2 = tex8_sample(c.edges, cx, cz, 0, 1, &f).y;
				vec2 a = area_ortho(&c, sx, sy, e1, e2, 0.0f);
				wz = a.x;
				ww = a.y;
				cx = f.u;
				corner_pattern(&c, &wz, &ww, cx, cy, cx, cz, d_x, d_y, 1);
			}
			weights_rgba8[(size_t)y * w + x] = unorm8(wx) | (unorm8(wy) << 8) | (unorm8(wz) << 16) | (unorm8(ww) << 24);
		}
}

/* ------------------------------------------------------------------------------------------------ blend */
/* inc/srgb.h:4-10 (the literals as glslang folds them) */
static float smaa_decode_srgb(float c)
{
	const float small_side = c / 12.9200000762939453125f;
	const float pow_side = powf((c + 0.054999999701976776123046875f) / 1.05499994754791259765625f, 2.400000095367431640625f);
	return f_clamp(c <= 0.0404482372105121612548828125f ? small_side : pow_side, 0.0f, 1.0f);
}

/* SMAANeighborhoodBlendingPS (SMAA.hlsl:1252-1307) with SMAANeighborhoodBlendingVS (:673-676), SMAA_TARGET_SRGB = 1:
 * colour = the sRGB image viewed as UNORM, the result is decoded to linear and stored into an sRGB attachment. */
void orc_smaa_neighborhood_blend(const uint32_t *color_unorm, const uint32_t *weights_rgba8, int w, int h, uint32_t *out_srgb8, int y0, int y1)
{
	const tex8 col = { (const uint8_t *)color_unorm, w, h, 4 }, bl = { (const uint8_t *)weights_rgba8, w, h, 4 };
	const float mx = 1.0f / (float)w, my = 1.0f / (float)h;
#pragma omp parallel for
	for (int y = y0; y < y1; y++)
		for (int x = 0; x < w; x++)
		{
			frag_t f = { ((float)x + 0.5f) * mx, ((float)y + 0.5f) * my, x, y };
			/* offset = mad(rt.xyxy, (1, 0, 0, 1), tc.xyxy) */
			const float ox = fmaf(mx, 1.0f, f.u), oy = fmaf(my, 0.0f, f.v), oz = fmaf(mx, 0.0f, f.u), ow = fmaf(my, 1.0f, f.v);
			float ax = tex8_sample(bl, ox, oy, 0, 0, &f).w;
			float ay = tex8_sample(bl, oz, ow, 0, 0, &f).y;
			vec4 here = tex8_sample(bl, f.u, f.v, 0, 0, &f);
			float aw = here.x, az = here.z;
			vec4 color;
			/* dot(a, vec4(1)): the pairwise order of the SPIR-V consumer the pin runs on, (x + y) + (z + w) */
			if ((ax * 1.0f + ay * 1.0f) + (az * 1.0f + aw * 1.0f) < 9.9999997473787516355514526367188e-06f)
				color = tex8_sample(col, f.u, f.v, 0, 0, &f);
			else
			{
				const int hz = f_max(ax, az) > f_max(ay, aw);
				float box = 0.0f, boy = ay, boz = 0.0f, bow = aw;
				float bwx = ay, bwy = aw;
				if (hz)
				{
					box = ax;
					boy = 0.0f;
					boz = az;
					bow = 0.0f;
					bwx = ax;
					bwy = az;
				}
				const float sum = bwx * 1.0f + bwy * 1.0f;
				bwx /= sum;
				bwy /= sum;
				/* blendingCoord = mad(blendingOffset, (rt.xy, -rt.xy), tc.xyxy) */
				const float cx = fmaf(box, mx, f.u), cy = fmaf(boy, my, f.v), cz = fmaf(boz, -mx, f.u), cw = fmaf(bow, -my, f.v);
				vec4 c0 = tex8_sample(col, cx, cy, 0, 0, &f), c1 = tex8_sample(col, cz, cw, 0, 0, &f);
				color = v4(c0.x * bwx, c0.y * bwx, c0.z * bwx, c0.w * bwx);
				color.x += c1.x * bwy;
				color.y += c1.y * bwy;
				color.z += c1.z * bwy;
				color.w += c1.w * bwy;
			}
			out_srgb8[(size_t)y * w + x] = linear_to_srgb8(smaa_decode_srgb(color.x)) | (linear_to_srgb8(smaa_decode_srgb(color.y)) << 8) |
			                              (linear_to_srgb8(smaa_decode_srgb(color.z)) << 16) | (unorm8(color.w) << 24);
		}
}
