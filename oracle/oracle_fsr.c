/*
 * oracle_fsr.c -- TEST INFRASTRUCTURE ONLY.  FidelityFX FSR 1 as the reference runs it after the post chain
 * (renderer/post/aa.cpp:34-174 setup_after_post_chain_upscaling): the edge-adaptive upscale pass
 * (assets/shaders/post/ffx-fsr/upscale.frag -> FsrEasuF, ffx_fsr1.h:232-436) and the contrast-adaptive sharpen
 * pass (sharpen.frag -> FsrRcasF, ffx_fsr1.h:679-775), both in their 32-bit paths (FP16 = 0).
 *
 * The algorithm is a fixed sequence of fp32 multiplies / adds, min / max, three integer-trick reciprocal
 * approximations (ffx_a.h:1843-1845) and one true division per pixel; it is restated here operation by operation in
 * the order the shader writes them (no fused multiply-add: the build uses -ffp-contract=off), so the stored 8-bit
 * codes can be compared exactly.  Pinned to the reference's own two shaders run on the CPU
 * (tests/test_oracle_ref_fsr.py, oracle/ref_post_shim.cpp).
 *
 * Inputs and stores (aa.cpp:84-86,96,112-114,141-146):
 *   upscale: reads the sRGB image as UNORM (gamma-space values) with textureGather through a NearestClamp sampler;
 *            writes R8G8B8A8_UNORM when a sharpen pass follows (TARGET_SRGB = 0: the gamma-space colour as it is),
 *            or the sRGB backbuffer when it is the last pass (TARGET_SRGB = 1: decode_srgb, then the attachment
 *            store encodes again);
 *   sharpen: reads the upscaled image through an sRGB view when the backbuffer is sRGB (texel fetch decodes to
 *            linear), sharpens, and the store encodes; UNORM in and out otherwise.
 */
#include "oracle.h"
#include "oracle_math.h"

#include <string.h>

static inline float as_f32(uint32_t u)
{
	float f;
	memcpy(&f, &u, 4);
	return f;
}

static inline uint32_t as_u32(float f)
{
	uint32_t u;
	memcpy(&u, &f, 4);
	return u;
}

/* ffx_a.h:1843-1845 */
static inline float prx_lo_rcp(float a) { return as_f32(0x7ef07ebbu - as_u32(a)); }
static inline float prx_lo_rsq(float a) { return as_f32(0x5f347d74u - (as_u32(a) >> 1)); }
static inline float prx_med_rcp(float a)
{
	float b = as_f32(0x7ef19fffu - as_u32(a));
	return b * (-b * a + 2.0f);
}
/* min / max return the non-NaN operand (IEEE minNum / maxNum, what a GPU's FMNMX does).  RCAS depends on it: where a
 * channel is 0 over the whole ring, 0 * (1 / 0) is NaN and max(-NaN, hitMax) must fall through to hitMax, or a pure-red
 * region would lose all three channels. */
#define f_min fminf
#define f_max fmaxf
static inline float sat(float x) { return f_min(f_max(x, 0.0f), 1.0f); }
static inline float min3(float x, float y, float z) { return f_min(x, f_min(y, z)); }
static inline float max3(float x, float y, float z) { return f_max(x, f_max(y, z)); }

/* aa.cpp:33-61 (FsrEasuCon with viewport == image): con0 .. con3 as 16 floats */
void orc_fsr_easu_constants(int w_in, int h_in, int w_out, int h_out, float *con16)
{
	const float ix = (float)w_in, iy = (float)h_in, ox = (float)w_out, oy = (float)h_out;
	con16[0] = ix / ox;
	con16[1] = iy / oy;
	con16[2] = 0.5f * ix / ox - 0.5f;
	con16[3] = 0.5f * iy / oy - 0.5f;
	con16[4] = 1.0f / ix;
	con16[5] = 1.0f / iy;
	con16[6] = 1.0f / ix;
	con16[7] = -1.0f / iy;
	con16[8] = -1.0f / ix;
	con16[9] = 2.0f / iy;
	con16[10] = 1.0f / ix;
	con16[11] = 2.0f / iy;
	con16[12] = 0.0f / ix;
	con16[13] = 4.0f / iy;
	con16[14] = con16[15] = 0.0f;
}

/* aa.cpp:63-73 (FsrRcasCon): con[0] = 2^-sharpness, con[1] = that value as two packed halves (read by the 16-bit
 * path only), con[2..3] = 0 */
void orc_fsr_rcas_constants(float sharpness, float *con4)
{
	sharpness = exp2f(-sharpness);
	uint32_t half = orc_float_to_half(sharpness);
	con4[0] = sharpness;
	con4[1] = as_f32(half | (half << 16));
	con4[2] = 0.0f;
	con4[3] = 0.0f;
}

typedef struct
{
	const uint32_t *p;
	int w, h;
} img8;

/* one texel of the UNORM view, clamp to edge */
static inline vec3 unorm_texel(img8 im, int x, int y)
{
	x = x < 0 ? 0 : (x > im.w - 1 ? im.w - 1 : x);
	y = y < 0 ? 0 : (y > im.h - 1 ? im.h - 1 : y);
	uint32_t t = im.p[(size_t)y * im.w + x];
	return v3((float)(t & 255u) / 255.0f, (float)((t >> 8) & 255u) / 255.0f, (float)((t >> 16) & 255u) / 255.0f);
}

/* textureGather at normalised (u, v): the 2 x 2 footprint of a bilinear sample, components
 * x = (i0, j1), y = (i1, j1), z = (i1, j0), w = (i0, j0) */
static inline void gather(img8 im, float u, float v, vec3 out[4])
{
	bilin_t s = bilin_setup(u, v, im.w, im.h);
	out[0] = unorm_texel(im, s.x0, s.y1);
	out[1] = unorm_texel(im, s.x1, s.y1);
	out[2] = unorm_texel(im, s.x1, s.y0);
	out[3] = unorm_texel(im, s.x0, s.y0);
}

static inline float luma2(vec3 c) { return c.z * 0.5f + (c.x * 0.5f + c.y); } /* ffx_fsr1.h:362-365 */

/* FsrEasuSetF, ffx_fsr1.h:275-313: one of the four '+' patterns around the resolve position */
static inline void easu_set(float *dir_x, float *dir_y, float *len, float w, float lA, float lB, float lC, float lD, float lE)
{
	float dc = lD - lC, cb = lC - lB;
	float lenX = f_max(fabsf(dc), fabsf(cb));
	lenX = prx_lo_rcp(lenX);
	float dirX = lD - lB;
	*dir_x += dirX * w;
	lenX = sat(fabsf(dirX) * lenX);
	lenX *= lenX;
	*len += lenX * w;
	float ec = lE - lC, ca = lC - lA;
	float lenY = f_max(fabsf(ec), fabsf(ca));
	lenY = prx_lo_rcp(lenY);
	float dirY = lE - lA;
	*dir_y += dirY * w;
	lenY = sat(fabsf(dirY) * lenY);
	lenY *= lenY;
	*len += lenY * w;
}

/* FsrEasuTapF, ffx_fsr1.h:239-273 */
static inline void easu_tap(vec3 *aC, float *aW, float off_x, float off_y, float dir_x, float dir_y, float len_x, float len_y, float lob, float clp, vec3 c)
{
	float vx = (off_x * dir_x) + (off_y * dir_y);
	float vy = (off_x * (-dir_y)) + (off_y * dir_x);
	vx *= len_x;
	vy *= len_y;
	float d2 = vx * vx + vy * vy;
	d2 = f_min(d2, clp);
	float wB = 0.4f * d2 + -1.0f;
	float wA = lob * d2 + -1.0f;
	wB *= wB;
	wA *= wA;
	wB = 1.5625f * wB + -0.5625f;
	float w = wB * wA;
	aC->x += c.x * w;
	aC->y += c.y * w;
	aC->z += c.z * w;
	*aW += w;
}

/* inc/srgb.h:4-10 with the literals as glslang folds them (upscale.frag:45-47) */
static float fsr_decode_srgb(float c)
{
	const float small_side = c / 12.9200000762939453125f;
	const float pow_side = powf((c + 0.054999999701976776123046875f) / 1.05499994754791259765625f, 2.400000095367431640625f);
	return f_clamp(c <= 0.0404482372105121612548828125f ? small_side : pow_side, 0.0f, 1.0f);
}

/* FsrEasuF, ffx_fsr1.h:315-436, for output pixel (x, y) */
static vec3 easu_pixel(img8 im, const float *con, int x, int y)
{
	/* position of 'f' in the input, and the fraction inside that texel */
	float ppx = (float)x * con[0] + con[2], ppy = (float)y * con[1] + con[3];
	float fpx = floorf(ppx), fpy = floorf(ppy);
	ppx -= fpx;
	ppy -= fpy;
	/* four gathers:   (0) b c / (1) e f i j / (2) g h k l / (3) n o */
	float p0x = fpx * con[4] + con[6], p0y = fpy * con[5] + con[7];
	vec3 g0[4], g1[4], g2[4], g3[4];
	gather(im, p0x, p0y, g0);
	gather(im, p0x + con[8], p0y + con[9], g1);
	gather(im, p0x + con[10], p0y + con[11], g2);
	gather(im, p0x + con[12], p0y + con[13], g3);
	const vec3 b = g0[0], c = g0[1];
	const vec3 i = g1[0], j = g1[1], f = g1[2], e = g1[3];
	const vec3 k = g2[0], l = g2[1], h = g2[2], g = g2[3];
	const vec3 o = g3[2], n = g3[3];
	const float bL = luma2(b), cL = luma2(c), iL = luma2(i), jL = luma2(j), fL = luma2(f), eL = luma2(e);
	const float kL = luma2(k), lL = luma2(l), hL = luma2(h), gL = luma2(g), oL = luma2(o), nL = luma2(n);
	/* direction and length, bilinearly weighted over the four texels around the position */
	float dir_x = 0.0f, dir_y = 0.0f, len = 0.0f;
	easu_set(&dir_x, &dir_y, &len, (1.0f - ppx) * (1.0f - ppy), bL, eL, fL, gL, jL);
	easu_set(&dir_x, &dir_y, &len, ppx * (1.0f - ppy), cL, fL, gL, hL, kL);
	easu_set(&dir_x, &dir_y, &len, (1.0f - ppx) * ppy, fL, iL, jL, kL, nL);
	easu_set(&dir_x, &dir_y, &len, ppx * ppy, gL, jL, kL, lL, oL);
	/* normalise the direction */
	float dirR = dir_x * dir_x + dir_y * dir_y;
	const int zro = dirR < (1.0f / 32768.0f);
	dirR = prx_lo_rsq(dirR);
	dirR = zro ? 1.0f : dirR;
	dir_x = zro ? 1.0f : dir_x;
	dir_x *= dirR;
	dir_y *= dirR;
	len = len * 0.5f;
	len *= len;
	const float stretch = (dir_x * dir_x + dir_y * dir_y) * prx_lo_rcp(f_max(fabsf(dir_x), fabsf(dir_y)));
	const float len_x = 1.0f + (stretch - 1.0f) * len, len_y = 1.0f + -0.5f * len;
	const float lob = 0.5f + -0.29f * len; /* AF1_((1.0/4.0-0.04)-0.5) */
	const float clp = prx_lo_rcp(lob);
	/* ring of the four nearest, for the de-ringing clamp */
	const vec3 mn = v3(f_min(min3(f.x, g.x, j.x), k.x), f_min(min3(f.y, g.y, j.y), k.y), f_min(min3(f.z, g.z, j.z), k.z));
	const vec3 mx = v3(f_max(max3(f.x, g.x, j.x), k.x), f_max(max3(f.y, g.y, j.y), k.y), f_max(max3(f.z, g.z, j.z), k.z));
	vec3 aC = v3(0.0f, 0.0f, 0.0f);
	float aW = 0.0f;
	easu_tap(&aC, &aW, 0.0f - ppx, -1.0f - ppy, dir_x, dir_y, len_x, len_y, lob, clp, b);
	easu_tap(&aC, &aW, 1.0f - ppx, -1.0f - ppy, dir_x, dir_y, len_x, len_y, lob, clp, c);
	easu_tap(&aC, &aW, -1.0f - ppx, 1.0f - ppy, dir_x, dir_y, len_x, len_y, lob, clp, i);
	easu_tap(&aC, &aW, 0.0f - ppx, 1.0f - ppy, dir_x, dir_y, len_x, len_y, lob, clp, j);
	easu_tap(&aC, &aW, 0.0f - ppx, 0.0f - ppy, dir_x, dir_y, len_x, len_y, lob, clp, f);
	easu_tap(&aC, &aW, -1.0f - ppx, 0.0f - ppy, dir_x, dir_y, len_x, len_y, lob, clp, e);
	easu_tap(&aC, &aW, 1.0f - ppx, 1.0f - ppy, dir_x, dir_y, len_x, len_y, lob, clp, k);
	easu_tap(&aC, &aW, 2.0f - ppx, 1.0f - ppy, dir_x, dir_y, len_x, len_y, lob, clp, l);
	easu_tap(&aC, &aW, 2.0f - ppx, 0.0f - ppy, dir_x, dir_y, len_x, len_y, lob, clp, h);
	easu_tap(&aC, &aW, 1.0f - ppx, 0.0f - ppy, dir_x, dir_y, len_x, len_y, lob, clp, g);
	easu_tap(&aC, &aW, 1.0f - ppx, 2.0f - ppy, dir_x, dir_y, len_x, len_y, lob, clp, o);
	easu_tap(&aC, &aW, 0.0f - ppx, 2.0f - ppy, dir_x, dir_y, len_x, len_y, lob, clp, n);
	const float rcp = 1.0f / aW;
	return v3(f_min(mx.x, f_max(mn.x, aC.x * rcp)), f_min(mx.y, f_max(mn.y, aC.y * rcp)), f_min(mx.z, f_max(mn.z, aC.z * rcp)));
}

void orc_fsr_easu(const uint32_t *src_unorm, int w_in, int h_in, const float *con16, uint32_t *dst, int w_out, int h_out, int target_srgb, int y0, int y1)
{
	const img8 im = { src_unorm, w_in, h_in };
	(void)h_out;
#pragma omp parallel for schedule(static)
	for (int y = y0; y < y1; y++)
		for (int x = 0; x < w_out; x++)
		{
			const vec3 c = easu_pixel(im, con16, x, y);
			uint32_t px;
			if (target_srgb)
				px = linear_to_srgb8(fsr_decode_srgb(c.x)) | (linear_to_srgb8(fsr_decode_srgb(c.y)) << 8) | (linear_to_srgb8(fsr_decode_srgb(c.z)) << 16);
			else
				px = float_to_unorm8(c.x) | (float_to_unorm8(c.y) << 8) | (float_to_unorm8(c.z) << 16);
			dst[(size_t)y * w_out + x] = px | 0xff000000u;
		}
}

/* FsrRcasLoadF of sharpen.frag:17: texelFetch with the coordinate clamped to the image */
static inline vec3 rcas_load(img8 im, int x, int y, int srgb)
{
	x = x < 0 ? 0 : (x > im.w - 1 ? im.w - 1 : x);
	y = y < 0 ? 0 : (y > im.h - 1 ? im.h - 1 : y);
	uint32_t t = im.p[(size_t)y * im.w + x];
	if (srgb)
		return v3(srgb8_to_linear(t & 255u), srgb8_to_linear((t >> 8) & 255u), srgb8_to_linear((t >> 16) & 255u));
	return v3((float)(t & 255u) / 255.0f, (float)((t >> 8) & 255u) / 255.0f, (float)((t >> 16) & 255u) / 255.0f);
}

#define RCAS_LIMIT (0.25f - (1.0f / 16.0f)) /* ffx_fsr1.h FSR_RCAS_LIMIT */

/* FsrRcasF, ffx_fsr1.h:684-775 (no FSR_RCAS_DENOISE, no alpha pass-through: sharpen.frag) */
void orc_fsr_rcas(const uint32_t *src, int w, int h, const float *con4, uint32_t *dst, int srgb, int y0, int y1)
{
	const img8 im = { src, w, h };
#pragma omp parallel for schedule(static)
	for (int y = y0; y < y1; y++)
		for (int x = 0; x < w; x++)
		{
			const vec3 b = rcas_load(im, x, y - 1, srgb), d = rcas_load(im, x - 1, y, srgb), e = rcas_load(im, x, y, srgb);
			const vec3 f = rcas_load(im, x + 1, y, srgb), hh = rcas_load(im, x, y + 1, srgb);
			/* ring minimum / maximum per channel */
			const float mnR = f_min(min3(b.x, d.x, f.x), hh.x), mnG = f_min(min3(b.y, d.y, f.y), hh.y), mnB = f_min(min3(b.z, d.z, f.z), hh.z);
			const float mxR = f_max(max3(b.x, d.x, f.x), hh.x), mxG = f_max(max3(b.y, d.y, f.y), hh.y), mxB = f_max(max3(b.z, d.z, f.z), hh.z);
			/* how much negative lobe each channel tolerates before clipping at 0 or 1 */
			const float hitMinR = mnR * (1.0f / (4.0f * mxR)), hitMinG = mnG * (1.0f / (4.0f * mxG)), hitMinB = mnB * (1.0f / (4.0f * mxB));
			const float hitMaxR = (1.0f - mxR) * (1.0f / (4.0f * mnR + -4.0f)), hitMaxG = (1.0f - mxG) * (1.0f / (4.0f * mnG + -4.0f)),
			            hitMaxB = (1.0f - mxB) * (1.0f / (4.0f * mnB + -4.0f));
			const float lobeR = f_max(-hitMinR, hitMaxR), lobeG = f_max(-hitMinG, hitMaxG), lobeB = f_max(-hitMinB, hitMaxB);
			const float lobe = f_max(-RCAS_LIMIT, f_min(max3(lobeR, lobeG, lobeB), 0.0f)) * con4[0];
			const float rcpL = prx_med_rcp(4.0f * lobe + 1.0f);
			const float pr = (lobe * b.x + lobe * d.x + lobe * hh.x + lobe * f.x + e.x) * rcpL;
			const float pg = (lobe * b.y + lobe * d.y + lobe * hh.y + lobe * f.y + e.y) * rcpL;
			const float pb = (lobe * b.z + lobe * d.z + lobe * hh.z + lobe * f.z + e.z) * rcpL;
			uint32_t px;
			if (srgb)
				px = linear_to_srgb8(pr) | (linear_to_srgb8(pg) << 8) | (linear_to_srgb8(pb) << 16);
			else
				px = float_to_unorm8(pr) | (float_to_unorm8(pg) << 8) | (float_to_unorm8(pb) << 16);
			dst[(size_t)y * w + x] = px | 0xff000000u;
		}
}
