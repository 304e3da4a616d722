// grb_fsr.cu -- FidelityFX FSR 1 after the post chain (second half): the reference's
// setup_after_post_chain_upscaling (renderer/post/aa.cpp:75-174) draws two full-screen passes,
//   upscale.frag -> FsrEasuF  (assets/shaders/post/ffx-fsr/ffx_fsr1.h:232-436): edge-adaptive spatial upsampling,
//                              12 taps of an anisotropic Lanczos-like kernel turned along the local gradient;
//   sharpen.frag -> FsrRcasF  (ffx_fsr1.h:679-775): robust contrast-adaptive sharpening, 5 taps;
// here one kernel each.  Both are streaming passes over 8-bit images (HBM-bound: the upscale reads w_in h_in 4 B and
// writes w_out h_out 4 B, its 12 taps overlap between neighbouring threads and come from L1; the sharpen pass moves
// 8 B per pixel), one thread per output pixel in 32 x 8 blocks.
//
// The arithmetic is the shaders' 32-bit path operation by operation (this file is built with -fmad=false: no
// contraction), so the stored 8-bit codes are compared EXACTLY with the oracle (oracle/oracle_fsr.c, itself pinned to
// the two reference shaders run on the CPU); min / max are fminf / fmaxf -- FMNMX returns the non-NaN operand, which
// RCAS relies on where a channel is 0 over the whole ring.  The kernels are also compiled for the CPU and checked
// against the oracle without a GPU (tests/cpp/emulate_fsr.cpp).
#include "grb_common.cuh"

namespace grb
{
namespace
{
// R8G8B8A8_SRGB texel fetch of the sharpen pass's input view (aa.cpp:141-144)
__device__ const float k_srgb8_to_linear[256] = {
#include "grb_srgb_table.inc"
};

// ffx_a.h:1843-1845: reciprocal / reciprocal square root from the exponent trick, with one Newton step for "Med"
GRB_DEV float prx_lo_rcp(float a) { return __uint_as_float(0x7ef07ebbu - __float_as_uint(a)); }
GRB_DEV float prx_lo_rsq(float a) { return __uint_as_float(0x5f347d74u - (__float_as_uint(a) >> 1)); }
GRB_DEV float prx_med_rcp(float a)
{
	const float b = __uint_as_float(0x7ef19fffu - __float_as_uint(a));
	return b * (-b * a + 2.0f);
}
GRB_DEV float sat(float x) { return fminf(fmaxf(x, 0.0f), 1.0f); }
GRB_DEV float min3(float x, float y, float z) { return fminf(x, fminf(y, z)); }
GRB_DEV float max3(float x, float y, float z) { return fmaxf(x, fmaxf(y, z)); }

struct Img8
{
	const uint32_t *p;
	int w, h, pitch; // pitch in texels
};

GRB_DEV float3 unorm_texel(const Img8 &im, int x, int y)
{
	x = iclamp(x, 0, im.w - 1);
	y = iclamp(y, 0, im.h - 1);
	const uint32_t t = __ldg(im.p + (size_t)y * im.pitch + x);
	return make_float3((float)(t & 255u) / 255.0f, (float)((t >> 8) & 255u) / 255.0f, (float)((t >> 16) & 255u) / 255.0f);
}

// top-left texel of the 2 x 2 footprint textureGather reads at normalised (u, v)
GRB_DEV int2 gather_origin(float u, float v, int w, int h)
{
	float fx = floorf(u * (float)w - 0.5f), fy = floorf(v * (float)h - 0.5f);
	fx = fminf(fmaxf(fx, -2.0f), (float)w + 1.0f);
	fy = fminf(fmaxf(fy, -2.0f), (float)h + 1.0f);
	return make_int2((int)fx, (int)fy);
}

GRB_DEV float luma2(float3 c) { return c.z * 0.5f + (c.x * 0.5f + c.y); }

// FsrEasuSetF (ffx_fsr1.h:275-313): gradient direction and edge length of one '+' pattern, bilinear weight w
GRB_DEV void easu_set(float &dir_x, float &dir_y, float &len, float w, float lA, float lB, float lC, float lD, float lE)
{
	const float dc = lD - lC, cb = lC - lB;
	float len_x = prx_lo_rcp(fmaxf(fabsf(dc), fabsf(cb)));
	const float dx = lD - lB;
	dir_x += dx * w;
	len_x = sat(fabsf(dx) * len_x);
	len_x *= len_x;
	len += len_x * w;
	const float ec = lE - lC, ca = lC - lA;
	float len_y = prx_lo_rcp(fmaxf(fabsf(ec), fabsf(ca)));
	const float dy = lE - lA;
	dir_y += dy * w;
	len_y = sat(fabsf(dy) * len_y);
	len_y *= len_y;
	len += len_y * w;
}

struct EasuKernel
{
	float dir_x, dir_y, len_x, len_y, lob, clp;
};

// FsrEasuTapF (ffx_fsr1.h:239-273): rotate the offset into the gradient frame, stretch, window
GRB_DEV void easu_tap(float3 &acc, float &acc_w, float off_x, float off_y, const EasuKernel &k, float3 c)
{
	float vx = (off_x * k.dir_x) + (off_y * k.dir_y);
	float vy = (off_x * (-k.dir_y)) + (off_y * k.dir_x);
	vx *= k.len_x;
	vy *= k.len_y;
	const float d2 = fminf(vx * vx + vy * vy, k.clp);
	float wb = 0.4f * d2 + -1.0f;
	float wa = k.lob * d2 + -1.0f;
	wb *= wb;
	wa *= wa;
	wb = 1.5625f * wb + -0.5625f;
	const float w = wb * wa;
	acc.x += c.x * w;
	acc.y += c.y * w;
	acc.z += c.z * w;
	acc_w += w;
}

// inc/srgb.h:4-10 with the literals glslang folds (upscale.frag:45-47)
GRB_DEV float fsr_decode_srgb(float c)
{
	const float small_side = c / 12.9200000762939453125f;
	const float pow_side = powf((c + 0.054999999701976776123046875f) / 1.05499994754791259765625f, 2.400000095367431640625f);
	return fclamp(c <= 0.0404482372105121612548828125f ? small_side : pow_side, 0.0f, 1.0f);
}

struct EasuConstants
{
	float c[16]; // con0 .. con3 of FsrEasuCon (aa.cpp:33-61)
};

// FsrEasuF (ffx_fsr1.h:315-436) for output pixel (x, y)
template <bool SrgbTarget>
__global__ void __launch_bounds__(256) fsr_easu_kernel(Img8 src, View<uint32_t> dst, EasuConstants con, int y0, int y1)
{
	const int x = blockIdx.x * 32 + threadIdx.x, y = y0 + blockIdx.y * 8 + threadIdx.y;
	if (x >= dst.w || y >= y1)
		return;
	const float *k = con.c;
	float ppx = (float)x * k[0] + k[2], ppy = (float)y * k[1] + k[3];
	const float fpx = floorf(ppx), fpy = floorf(ppy);
	ppx -= fpx;
	ppy -= fpy;
	// the four gathers of the shader, as footprint origins:  (0) b c   (1) e f / i j   (2) g h / k l   (3) n o
	const float p0x = fpx * k[4] + k[6], p0y = fpy * k[5] + k[7];
	const int2 o0 = gather_origin(p0x, p0y, src.w, src.h), o1 = gather_origin(p0x + k[8], p0y + k[9], src.w, src.h);
	const int2 o2 = gather_origin(p0x + k[10], p0y + k[11], src.w, src.h), o3 = gather_origin(p0x + k[12], p0y + k[13], src.w, src.h);
	const float3 b = unorm_texel(src, o0.x, o0.y + 1), c = unorm_texel(src, o0.x + 1, o0.y + 1);
	const float3 i = unorm_texel(src, o1.x, o1.y + 1), j = unorm_texel(src, o1.x + 1, o1.y + 1), f = unorm_texel(src, o1.x + 1, o1.y), e = unorm_texel(src, o1.x, o1.y);
	const float3 kk = unorm_texel(src, o2.x, o2.y + 1), l = unorm_texel(src, o2.x + 1, o2.y + 1), h = unorm_texel(src, o2.x + 1, o2.y), g = unorm_texel(src, o2.x, o2.y);
	const float3 o = unorm_texel(src, o3.x + 1, o3.y), n = unorm_texel(src, o3.x, o3.y);
	const float bL = luma2(b), cL = luma2(c), iL = luma2(i), jL = luma2(j), fL = luma2(f), eL = luma2(e);
	const float kL = luma2(kk), lL = luma2(l), hL = luma2(h), gL = luma2(g), oL = luma2(o), nL = luma2(n);

	float dir_x = 0.0f, dir_y = 0.0f, len = 0.0f;
	easu_set(dir_x, dir_y, len, (1.0f - ppx) * (1.0f - ppy), bL, eL, fL, gL, jL);
	easu_set(dir_x, dir_y, len, ppx * (1.0f - ppy), cL, fL, gL, hL, kL);
	easu_set(dir_x, dir_y, len, (1.0f - ppx) * ppy, fL, iL, jL, kL, nL);
	easu_set(dir_x, dir_y, len, ppx * ppy, gL, jL, kL, lL, oL);

	float dir_r = dir_x * dir_x + dir_y * dir_y;
	const bool zro = dir_r < (1.0f / 32768.0f);
	dir_r = zro ? 1.0f : prx_lo_rsq(dir_r);
	dir_x = zro ? 1.0f : dir_x;
	EasuKernel kern;
	kern.dir_x = dir_x * dir_r;
	kern.dir_y = dir_y * dir_r;
	len = len * 0.5f;
	len *= len;
	const float stretch = (kern.dir_x * kern.dir_x + kern.dir_y * kern.dir_y) * prx_lo_rcp(fmaxf(fabsf(kern.dir_x), fabsf(kern.dir_y)));
	kern.len_x = 1.0f + (stretch - 1.0f) * len;
	kern.len_y = 1.0f + -0.5f * len;
	kern.lob = 0.5f + -0.29f * len;
	kern.clp = prx_lo_rcp(kern.lob);

	const float3 mn = make_float3(fminf(min3(f.x, g.x, j.x), kk.x), fminf(min3(f.y, g.y, j.y), kk.y), fminf(min3(f.z, g.z, j.z), kk.z));
	const float3 mx = make_float3(fmaxf(max3(f.x, g.x, j.x), kk.x), fmaxf(max3(f.y, g.y, j.y), kk.y), fmaxf(max3(f.z, g.z, j.z), kk.z));
	float3 acc = make_float3(0.0f, 0.0f, 0.0f);
	float acc_w = 0.0f;
	easu_tap(acc, acc_w, 0.0f - ppx, -1.0f - ppy, kern, b);
	easu_tap(acc, acc_w, 1.0f - ppx, -1.0f - ppy, kern, c);
	easu_tap(acc, acc_w, -1.0f - ppx, 1.0f - ppy, kern, i);
	easu_tap(acc, acc_w, 0.0f - ppx, 1.0f - ppy, kern, j);
	easu_tap(acc, acc_w, 0.0f - ppx, 0.0f - ppy, kern, f);
	easu_tap(acc, acc_w, -1.0f - ppx, 0.0f - ppy, kern, e);
	easu_tap(acc, acc_w, 1.0f - ppx, 1.0f - ppy, kern, kk);
	easu_tap(acc, acc_w, 2.0f - ppx, 1.0f - ppy, kern, l);
	easu_tap(acc, acc_w, 2.0f - ppx, 0.0f - ppy, kern, h);
	easu_tap(acc, acc_w, 1.0f - ppx, 0.0f - ppy, kern, g);
	easu_tap(acc, acc_w, 1.0f - ppx, 2.0f - ppy, kern, o);
	easu_tap(acc, acc_w, 0.0f - ppx, 2.0f - ppy, kern, n);
	const float rcp = 1.0f / acc_w;
	const float r = fminf(mx.x, fmaxf(mn.x, acc.x * rcp)), gg = fminf(mx.y, fmaxf(mn.y, acc.y * rcp)), bb = fminf(mx.z, fmaxf(mn.z, acc.z * rcp));
	uint32_t px;
	if (SrgbTarget)
		px = linear_to_srgb8(fsr_decode_srgb(r)) | (linear_to_srgb8(fsr_decode_srgb(gg)) << 8) | (linear_to_srgb8(fsr_decode_srgb(bb)) << 16);
	else
		px = float_to_unorm8(r) | (float_to_unorm8(gg) << 8) | (float_to_unorm8(bb) << 16);
	dst.at(x, y) = px | 0xff000000u;
}

template <bool Srgb>
GRB_DEV float3 rcas_load(const Img8 &im, int x, int y)
{
	x = iclamp(x, 0, im.w - 1);
	y = iclamp(y, 0, im.h - 1);
	const uint32_t t = __ldg(im.p + (size_t)y * im.pitch + x);
	if (Srgb)
		return make_float3(k_srgb8_to_linear[t & 255u], k_srgb8_to_linear[(t >> 8) & 255u], k_srgb8_to_linear[(t >> 16) & 255u]);
	return make_float3((float)(t & 255u) / 255.0f, (float)((t >> 8) & 255u) / 255.0f, (float)((t >> 16) & 255u) / 255.0f);
}

// FsrRcasF (ffx_fsr1.h:684-775; sharpen.frag defines neither FSR_RCAS_DENOISE nor the alpha pass-through)
template <bool Srgb>
__global__ void __launch_bounds__(256) fsr_rcas_kernel(Img8 src, View<uint32_t> dst, float sharpness, int y0, int y1)
{
	const int x = blockIdx.x * 32 + threadIdx.x, y = y0 + blockIdx.y * 8 + threadIdx.y;
	if (x >= dst.w || y >= y1)
		return;
	const float3 b = rcas_load<Srgb>(src, x, y - 1), d = rcas_load<Srgb>(src, x - 1, y), e = rcas_load<Srgb>(src, x, y);
	const float3 f = rcas_load<Srgb>(src, x + 1, y), h = rcas_load<Srgb>(src, x, y + 1);
	const float mn_r = fminf(min3(b.x, d.x, f.x), h.x), mn_g = fminf(min3(b.y, d.y, f.y), h.y), mn_b = fminf(min3(b.z, d.z, f.z), h.z);
	const float mx_r = fmaxf(max3(b.x, d.x, f.x), h.x), mx_g = fmaxf(max3(b.y, d.y, f.y), h.y), mx_b = fmaxf(max3(b.z, d.z, f.z), h.z);
	// the negative lobe each channel tolerates before the result clips at 0 or at 1 (true divisions: "need to be high
	// precision RCPs")
	const float hit_min_r = mn_r * (1.0f / (4.0f * mx_r)), hit_min_g = mn_g * (1.0f / (4.0f * mx_g)), hit_min_b = mn_b * (1.0f / (4.0f * mx_b));
	const float hit_max_r = (1.0f - mx_r) * (1.0f / (4.0f * mn_r + -4.0f)), hit_max_g = (1.0f - mx_g) * (1.0f / (4.0f * mn_g + -4.0f)),
	            hit_max_b = (1.0f - mx_b) * (1.0f / (4.0f * mn_b + -4.0f));
	const float lobe_r = fmaxf(-hit_min_r, hit_max_r), lobe_g = fmaxf(-hit_min_g, hit_max_g), lobe_b = fmaxf(-hit_min_b, hit_max_b);
	const float lobe = fmaxf(-(0.25f - (1.0f / 16.0f)), fminf(max3(lobe_r, lobe_g, lobe_b), 0.0f)) * sharpness;
	const float rcp_l = prx_med_rcp(4.0f * lobe + 1.0f);
	const float pr = (lobe * b.x + lobe * d.x + lobe * h.x + lobe * f.x + e.x) * rcp_l;
	const float pg = (lobe * b.y + lobe * d.y + lobe * h.y + lobe * f.y + e.y) * rcp_l;
	const float pb = (lobe * b.z + lobe * d.z + lobe * h.z + lobe * f.z + e.z) * rcp_l;
	uint32_t px;
	if (Srgb)
		px = linear_to_srgb8(pr) | (linear_to_srgb8(pg) << 8) | (linear_to_srgb8(pb) << 16);
	else
		px = float_to_unorm8(pr) | (float_to_unorm8(pg) << 8) | (float_to_unorm8(pb) << 16);
	dst.at(x, y) = px | 0xff000000u;
}

Img8 img8_of(const GrbImage *im)
{
	Img8 t;
	t.p = static_cast<const uint32_t *>(im->data);
	t.w = im->width;
	t.h = im->height;
	t.pitch = im->row_pitch / 4;
	return t;
}

bool fsr_rgba8(const GrbImage *im) { return image_ok(im, GRB_FORMAT_R8G8B8A8_UNORM, 4) || image_ok(im, GRB_FORMAT_R8G8B8A8_SRGB, 4); }
} // namespace
} // namespace grb

#ifndef GRB_HOST_EMULATION // tests/cpp/emulate_fsr.cpp compiles the kernels above for the CPU and supplies its own loops
using namespace grb;

// FsrEasuCon as the reference's builder evaluates it (aa.cpp:33-61, viewport == whole input image)
extern "C" int32_t grb_fsr_easu_constants(int32_t in_width, int32_t in_height, int32_t out_width, int32_t out_height, float *con16)
{
	if (!con16 || in_width <= 0 || in_height <= 0 || out_width <= 0 || out_height <= 0)
	{
		set_last_error("grb_fsr_easu_constants: bad arguments");
		return GRB_ERR_INVALID_ARGUMENT;
	}
	const float ix = (float)in_width, iy = (float)in_height, ox = (float)out_width, oy = (float)out_height;
	const float v[16] = { ix / ox, iy / oy, 0.5f * ix / ox - 0.5f, 0.5f * iy / oy - 0.5f, 1.0f / ix, 1.0f / iy, 1.0f / ix, -1.0f / iy,
		              -1.0f / ix, 2.0f / iy, 1.0f / ix, 2.0f / iy, 0.0f / ix, 4.0f / iy, 0.0f, 0.0f };
	for (int i = 0; i < 16; i++)
		con16[i] = v[i];
	return GRB_OK;
}

extern "C" int32_t grb_fsr_upscale(const GrbImage *color, const GrbImage *out, GrbRows rows, void *stream)
{
	if (!color || !out || !fsr_rgba8(color) || !fsr_rgba8(out) || color->data == out->data)
	{
		set_last_error("grb_fsr_upscale: color R8G8B8A8 (read as UNORM), out R8G8B8A8 (UNORM: gamma-space colour as it is; SRGB: decode + encode), out != color");
		return GRB_ERR_UNSUPPORTED_FORMAT;
	}
	rows = full_rows(rows, out->height);
	if (rows.y1 <= rows.y0)
		return GRB_OK;
	EasuConstants con;
	grb_fsr_easu_constants(color->width, color->height, out->width, out->height, con.c);
	const dim3 grid((unsigned)((out->width + 31) / 32), (unsigned)((rows.y1 - rows.y0 + 7) / 8), 1), block(32, 8);
	if (out->format == GRB_FORMAT_R8G8B8A8_SRGB)
		fsr_easu_kernel<true><<<grid, block, 0, as_stream(stream)>>>(img8_of(color), view_of<uint32_t>(out), con, rows.y0, rows.y1);
	else
		fsr_easu_kernel<false><<<grid, block, 0, as_stream(stream)>>>(img8_of(color), view_of<uint32_t>(out), con, rows.y0, rows.y1);
	return check_launch("grb_fsr_upscale");
}

extern "C" int32_t grb_fsr_sharpen(const GrbImage *color, const GrbImage *out, float sharpness_stops, GrbRows rows, void *stream)
{
	if (!color || !out || !fsr_rgba8(color) || !fsr_rgba8(out) || color->width != out->width || color->height != out->height || color->data == out->data)
	{
		set_last_error("grb_fsr_sharpen: color and out R8G8B8A8 of one size, out != color (out SRGB: color is read through an sRGB view)");
		return GRB_ERR_UNSUPPORTED_FORMAT;
	}
	if (!(sharpness_stops >= 0.0f))
	{
		set_last_error("grb_fsr_sharpen: sharpness is a number of stops >= 0 (0 = maximum)");
		return GRB_ERR_INVALID_ARGUMENT;
	}
	rows = full_rows(rows, out->height);
	if (rows.y1 <= rows.y0)
		return GRB_OK;
	const float sharpness = exp2f(-sharpness_stops); // FsrRcasCon, aa.cpp:63-66
	const dim3 grid((unsigned)((out->width + 31) / 32), (unsigned)((rows.y1 - rows.y0 + 7) / 8), 1), block(32, 8);
	if (out->format == GRB_FORMAT_R8G8B8A8_SRGB)
		fsr_rcas_kernel<true><<<grid, block, 0, as_stream(stream)>>>(img8_of(color), view_of<uint32_t>(out), sharpness, rows.y0, rows.y1);
	else
		fsr_rcas_kernel<false><<<grid, block, 0, as_stream(stream)>>>(img8_of(color), view_of<uint32_t>(out), sharpness, rows.y0, rows.y1);
	return check_launch("grb_fsr_sharpen");
}
#endif // GRB_HOST_EMULATION
