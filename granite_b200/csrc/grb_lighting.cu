// grb_lighting.cu -- clustered deferred lighting as one sm_100a kernel.
//
// Replaces DeferredLightRenderer::render_light (renderer/renderer.cpp:1004-1156), i.e. the two
// full-screen draws directional.frag and clustering.frag that are additively blended into
// "HDR-main".  Here both happen in ONE pass over the G-buffer: each thread owns one pixel,
// reads its 18 bytes of G-buffer + 4 bytes of HDR once, evaluates the directional light, then
// walks the pixel's light cluster, and writes 4 bytes.  The two blends' intermediate
// B10G11R11 quantisation is reproduced in registers, so the HBM traffic is the compulsory
// 22 B/pixel.
//
// Work mapping: a warp is an 8x4 pixel quad-block (the footprint the reference's fragment
// subgroups have), a CTA is 4 warps side by side (32x4 pixels, so every G-buffer row segment a
// CTA touches is a full 128-byte line).  The light loop is warp-uniform like the reference's
// subgroup-scalarised loop (clusterer_bindless.h:49-81): the warp walks the union of its lanes'
// cluster masks, every lane evaluates the same light (its record is a broadcast load), and a
// lane only ACCUMULATES a light that is in its own (tile, z-slice) mask -- which makes the
// result exactly the per-pixel function, in ascending light order.
//
// Numerics: the cluster indices (tile, z slice) are part of the bit-exact contract, so the
// position reconstruction up to those indices uses non-contracted IEEE ops (fmul/fadd/...).
// The BRDF itself is plain fp32 with FMA and fast reciprocal-sqrt: the result is stored as
// B10G11R11 (6/5 mantissa bits), five orders of magnitude coarser than those rounding
// differences.
#include "grb_common.cuh"

#include <cstdlib>

namespace grb
{
__device__ float g_srgb8_to_linear[256];

namespace
{
constexpr float kPi = 3.1415628f; // assets/shaders/lights/pbr.h:5 (sic)
constexpr float kInvPi = 1.0f / kPi;

struct LightingParams
{
	View<const uint32_t> albedo, normal;
	View<const uint16_t> pbr;
	View<const float> depth;
	View<uint32_t> hdr;
	View<const uint32_t> emissive; // blend destination's initial contents (may be the hdr image itself)
	float ivp[16];
	float3 camera_pos;
	float3 dir_color, dir_dir;
	// cluster
	float3 cbase, cfront;
	float2 xy_scale;
	int res_x, res_y;
	int n32, z_max_index;
	float z_scale;
	float inv_res_x, inv_res_y;
	const GrbPositionalLight *lights;
	const uint32_t *type_mask;
	const uint32_t *bitmask;
	const uint2 *cluster_range;
	int y0, y1;
};

struct Surface
{
	float3 pos, N, V;
	float3 F0, one_minus_F0;
	float3 diffuse_k;       // base_color * (1 - metallic) / PI
	float m2_minus_1;       // roughness'^4 - 1
	float c_gd;             // 0.25 * roughness'^4 / PI  (numerator of G*D)
	float one_minus_k, k, Vk; // Schlick-GGX visibility pieces
	float NoV_raw;          // dot(N, V) before the clamp (half-vector algebra)
};

__device__ __forceinline__ float dot3(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

__device__ __forceinline__ uint32_t cluster_mask_range(uint32_t mask, uint32_t rx, uint32_t ry, uint32_t start)
{
	uint32_t hi = start + 32u;
	rx = min(max(rx, start), hi);
	uint32_t ry1 = min(max(ry + 1u, rx), hi);
	uint32_t num_bits = ry1 - rx;
	uint32_t range_mask = num_bits == 32u ? 0xffffffffu : (((1u << num_bits) - 1u) << (rx - start));
	return mask & range_mask;
}

// Cook-Torrance terms shared by the directional and the positional lights (lighting.h:26-46,
// point.h:121-141, spot.h:124-144), arranged for the fewest issue slots -- the pass is bound by
// instruction issue, not by HBM:
//   specular + diffuse = F*G*D + (1-F)*dk = dk + F*(G*D - dk)
//   D*G = (m2 / (PI d^2)) * (0.25 / max(Vk*Lk, 1e-3)),  d = NoH^2 (m2 - 1) + 1
// Returns that sum per channel and NoL; the caller scales by NoL * colour * attenuation.
__device__ __forceinline__ float3 brdf(const Surface &s, float3 L, float &NoL)
{
	// With unit V and L:  |V+L|^2 = 2 + 2 VoL,  N.(V+L) = NoV + NoL,  V.(V+L) = 1 + VoL -- the half vector
	// itself is never formed.
	float VoL = dot3(s.V, L);
	float NoLr = dot3(s.N, L);
	float inv_h = rsqrt_fast(fmaf(VoL, 2.0f, 2.0f));
	NoL = fminf(fmaxf(NoLr, 0.001f), 1.0f);
	float NoH = fminf(fmaxf((NoLr + s.NoV_raw) * inv_h, 0.0001f), 1.0f);
	// HoV = sqrt((1 + VoL) / 2) cannot exceed 1 by more than rounding, and 1 - HoV only enters as f^5
	float HoV = fmaxf(fmaf(VoL, inv_h, inv_h), 0.001f);
	float f = 1.0f - HoV;
	float f2 = f * f;
	float f5 = f2 * f2 * f;
	float d = fmaf(NoH * NoH, s.m2_minus_1, 1.0f);
	// max(Vk * Lk, 1e-3) of the shader is the identity: k = (r' + 1)^2 / 8 >= 0.195 bounds both factors
	float vl = s.Vk * fmaf(NoL, s.one_minus_k, s.k);
	float GD = s.c_gd * rcp_fast(d * d * vl);
	float Fx = fmaf(s.one_minus_F0.x, f5, s.F0.x), Fy = fmaf(s.one_minus_F0.y, f5, s.F0.y), Fz = fmaf(s.one_minus_F0.z, f5, s.F0.z);
	return make_float3(fmaf(Fx, GD - s.diffuse_k.x, s.diffuse_k.x), fmaf(Fy, GD - s.diffuse_k.y, s.diffuse_k.y),
	                   fmaf(Fz, GD - s.diffuse_k.z, s.diffuse_k.z));
}

// World position of a pixel and its cluster coordinates.  The tile index and Z slice are part
// of the bit-exact contract with the reference (clustering.vert:10-14, clustering.frag:38-39,
// clusterer_bindless.h:39-47), so every op here is a non-contracted IEEE op in a fixed order.
__device__ __forceinline__ float3 reconstruct_position_and_cluster(const LightingParams &p, int x, int y, float depth, int &tile_index, int &z_index)
{
	// vClip = invVP * (ndc.xy, 0, 1) interpolated at the pixel centre, + depth * invVP[2]
	const float ndc_x = fsub(fmul(fmul(2.0f, fadd((float)x, 0.5f)), p.inv_res_x), 1.0f);
	const float ndc_y = fsub(fmul(fmul(2.0f, fadd((float)y, 0.5f)), p.inv_res_y), 1.0f);
	const float *m = p.ivp;
	float cx = fadd(fadd(fadd(fmul(m[0], ndc_x), fmul(m[4], ndc_y)), m[12]), fmul(depth, m[8]));
	float cy = fadd(fadd(fadd(fmul(m[1], ndc_x), fmul(m[5], ndc_y)), m[13]), fmul(depth, m[9]));
	float cz = fadd(fadd(fadd(fmul(m[2], ndc_x), fmul(m[6], ndc_y)), m[14]), fmul(depth, m[10]));
	float cw = fadd(fadd(fadd(fmul(m[3], ndc_x), fmul(m[7], ndc_y)), m[15]), fmul(depth, m[11]));
	float3 pos = make_float3(fdiv(cx, cw), fdiv(cy, cw), fdiv(cz, cw));
	int tx = __float2int_rz(fmul(fmul(fadd((float)x, 0.5f), p.inv_res_x), p.xy_scale.x));
	int ty = __float2int_rz(fmul(fmul(fadd((float)y, 0.5f), p.inv_res_y), p.xy_scale.y));
	tx = iclamp(tx, 0, p.res_x - 1);
	ty = iclamp(ty, 0, p.res_y - 1);
	tile_index = ty * p.res_x + tx;
	float zv = fadd(fadd(fmul(fsub(pos.x, p.cbase.x), p.cfront.x), fmul(fsub(pos.y, p.cbase.y), p.cfront.y)), fmul(fsub(pos.z, p.cbase.z), p.cfront.z));
	z_index = iclamp(__float2int_rz(fmul(zv, p.z_scale)), 0, p.z_max_index);
	return pos;
}

// Diagnostic twin of the lighting kernel's addressing: writes (tile index, z slice) per pixel.
__global__ void __launch_bounds__(128) cluster_indices_kernel(const LightingParams p, int *__restrict__ out_tile, int *__restrict__ out_z)
{
	int x = blockIdx.x * 32 + (threadIdx.x & 31);
	int y = p.y0 + blockIdx.y * 4 + (threadIdx.x >> 5);
	if (x >= p.depth.w || y >= p.y1)
		return;
	float depth = __ldg(&p.depth.at(x, y));
	int tile = -1, z = -1;
	if (depth != 0.0f)
		reconstruct_position_and_cluster(p, x, y, depth, tile, z);
	out_tile[(size_t)y * p.depth.w + x] = tile;
	out_z[(size_t)y * p.depth.w + x] = z;
}

constexpr int kWarpsPerCta = 4;

__global__ void __launch_bounds__(32 * kWarpsPerCta) deferred_lighting_kernel(const LightingParams p)
{
	__shared__ float s_srgb[256];
	for (int i = threadIdx.x; i < 256; i += blockDim.x)
		s_srgb[i] = g_srgb8_to_linear[i];
	__syncthreads();

	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int x = (blockIdx.x * kWarpsPerCta + warp) * 8 + (lane & 7);
	const int y = p.y0 + blockIdx.y * 4 + (lane >> 3);
	const bool inside = x < p.hdr.w && y < p.y1;

	float depth = 0.0f;
	if (inside)
		depth = __ldg(&p.depth.at(x, y));
	// depth test NOT_EQUAL against the quad's z = 0: sky pixels keep the attachment value
	const bool lit = inside && depth != 0.0f;

	Surface s;
	uint32_t dst = 0u;
	uint32_t rx = 0xffffffffu, ry = 0u;
	int cluster_base = 0;
	float3 base_color = make_float3(0.f, 0.f, 0.f);
	if (lit)
	{
		const uint32_t a8 = __ldg(&p.albedo.at(x, y));
		const uint32_t n10 = __ldg(&p.normal.at(x, y));
		const uint32_t mr = __ldg(&p.pbr.at(x, y));
		dst = __ldg(&p.emissive.at(x, y));

		base_color = make_float3(s_srgb[a8 & 0xffu], s_srgb[(a8 >> 8) & 0xffu], s_srgb[(a8 >> 16) & 0xffu]);
		// UNORM decode: these feed only the BRDF (not the bit-exact indices), a multiply by the
		// reciprocal is within half an ulp of the division
		s.N = make_float3(fmaf((float)(n10 & 0x3ffu), 2.0f / 1023.0f, -1.0f), fmaf((float)((n10 >> 10) & 0x3ffu), 2.0f / 1023.0f, -1.0f),
		                  fmaf((float)((n10 >> 20) & 0x3ffu), 2.0f / 1023.0f, -1.0f));
		const float metallic = (float)(mr & 0xffu) * (1.0f / 255.0f);
		const float roughness_in = (float)(mr >> 8) * (1.0f / 255.0f);

		int tile_index, z_index;
		s.pos = reconstruct_position_and_cluster(p, x, y, depth, tile_index, z_index);
		cluster_base = tile_index * p.n32;
		uint2 zr = __ldg(&p.cluster_range[z_index]);
		rx = zr.x;
		ry = zr.y;

		// per-pixel BRDF invariants
		float3 v = make_float3(p.camera_pos.x - s.pos.x, p.camera_pos.y - s.pos.y, p.camera_pos.z - s.pos.z);
		float inv_v = rsqrt_fast(dot3(v, v));
		s.V = make_float3(v.x * inv_v, v.y * inv_v, v.z * inv_v);
		float rough = roughness_in * 0.75f + 0.25f;
		float mm = rough * rough;
		float m2 = mm * mm;
		s.m2_minus_1 = m2 - 1.0f;
		s.c_gd = m2 * (0.25f * kInvPi);
		float r1 = rough + 1.0f;
		s.k = r1 * r1 * 0.125f;
		s.one_minus_k = 1.0f - s.k;
		s.NoV_raw = dot3(s.N, s.V);
		float NoV = fminf(fmaxf(s.NoV_raw, 0.001f), 1.0f);
		s.Vk = NoV * s.one_minus_k + s.k;
		s.F0 = make_float3(0.04f * (1.0f - metallic) + base_color.x * metallic, 0.04f * (1.0f - metallic) + base_color.y * metallic,
		                   0.04f * (1.0f - metallic) + base_color.z * metallic);
		s.one_minus_F0 = make_float3(1.0f - s.F0.x, 1.0f - s.F0.y, 1.0f - s.F0.z);
		float dk = (1.0f - metallic) * kInvPi;
		s.diffuse_k = make_float3(base_color.x * dk, base_color.y * dk, base_color.z * dk);

		// ---- draw 1: directional.frag (LIGHTING_NO_AMBIENT, no shadows, VOLUMETRIC_DIFFUSE_FALLBACK) ----
		float NoL;
		float3 b = brdf(s, p.dir_dir, NoL);
		float3 e = unpack_r11g11b10(dst);
		dst = pack_r11g11b10(e.x + p.dir_color.x * NoL * b.x + base_color.x * 0.05f, e.y + p.dir_color.y * NoL * b.y + base_color.y * 0.05f,
		                     e.z + p.dir_color.z * NoL * b.z + base_color.z * 0.05f);
	}

	// ---- draw 2: clustering.frag, warp-uniform walk over the union of the lanes' masks ----
	const uint32_t lo_word = rx >> 5, hi_word = ry >> 5; // inactive lanes: (0x7ffffff, 0) => empty
	int z_start = (int)__reduce_min_sync(0xffffffffu, lo_word);
	int z_end = (int)__reduce_max_sync(0xffffffffu, lit ? hi_word : 0u);
	z_end = min(z_end, p.n32 - 1);
	float3 acc = make_float3(0.f, 0.f, 0.f);
	for (int i = z_start; i <= z_end; i++)
	{
		uint32_t own = 0u;
		if (lit && (uint32_t)i >= lo_word && (uint32_t)i <= hi_word)
			own = cluster_mask_range(__ldg(&p.bitmask[cluster_base + i]), rx, ry, 32u * (uint32_t)i);
		uint32_t wmask = __reduce_or_sync(0xffffffffu, own);
		const uint32_t tm = __ldg(&p.type_mask[i]);
		while (wmask)
		{
			const int bit = __ffs(wmask) - 1;
			wmask &= wmask - 1u;
			const float4 *lp = reinterpret_cast<const float4 *>(p.lights + (i * 32 + bit));
			const float4 l1 = __ldg(lp + 1), l2 = __ldg(lp + 2); // position|offset_radius, direction|inv_radius
			float3 l = make_float3(l1.x - s.pos.x, l1.y - s.pos.y, l1.z - s.pos.z);
			float d2 = dot3(l, l);
			// quick reject: beyond the light's radius the falloff is exactly 0 (point.h:41-43); most lights
			// of a 30x34-pixel cluster tile do not reach this 8x4 block, so the warp usually leaves here
			const bool near = ((own >> bit) & 1u) && (d2 * l2.w * l2.w < 1.0f);
			if (!__any_sync(0xffffffffu, near))
				continue;
			const float4 l0 = __ldg(lp); // color|spot scale_bias
			float inv_d = rsqrt_fast(d2);
			float inv_ld = fminf(inv_d, 10.0f); // 1 / max(0.1, dist)
			float xr = fmaxf(0.1f, d2 * inv_d) * l2.w;
			float t = __saturatef(fmaf(xr, 1.0f / (1.0f - 0.9f), -0.9f / (1.0f - 0.9f)));
			float falloff = fmaf(-t * t, fmaf(-2.0f, t, 3.0f), 1.0f);
			float3 L = make_float3(l.x * inv_d, l.y * inv_d, l.z * inv_d);
			if (!((tm >> bit) & 1u))
			{
				// spot.h:34-84: cone term from the packed fp16 scale/bias
				float2 sb = __half22float2(*reinterpret_cast<const __half2 *>(&l0.w));
				float cone_angle = -(L.x * l2.x + L.y * l2.y + L.z * l2.z);
				float cone = __saturatef(fmaf(cone_angle, sb.x, sb.y));
				falloff *= cone * cone;
			}
			float NoL;
			float3 b = brdf(s, L, NoL);
			float w = NoL * falloff * inv_ld * inv_ld;
			if (near && falloff > 0.0f)
			{
				acc.x = fmaf(l0.x * w, b.x, acc.x);
				acc.y = fmaf(l0.y * w, b.y, acc.y);
				acc.z = fmaf(l0.z * w, b.z, acc.z);
			}
		}
	}

	if (lit)
	{
		float3 e = unpack_r11g11b10(dst);
		p.hdr.at(x, y) = pack_r11g11b10(e.x + acc.x, e.y + acc.y, e.z + acc.z);
	}
	else if (inside && p.emissive.p != p.hdr.p)
		p.hdr.at(x, y) = __ldg(&p.emissive.at(x, y)); // sky keeps the attachment value
}
// ---------------------------------------------------------------------------------------------
// Two pixels per thread, packed fp32 (FFMA2 / FMUL2 / FADD2 of sm_100).
//
// The pass is bound by instruction issue (ncu: ~80 % issue utilisation, 3 % DRAM), and ~60 % of
// the issued instructions are fp32 multiply/add.  Blackwell issues TWO fp32 operations per
// FFMA2-class instruction, so the kernel below carries two horizontally adjacent pixels per
// thread in float2 lanes: the per-light vector math (light vector, distances, half vector, the
// three dot products, Fresnel, the GGX terms) is issued once for both, and the per-light
// control overhead (mask walk, record loads, votes) is amortised over twice the pixels.  A warp
// covers a 16x4 pixel block, a CTA 64x4.  Results are the same function as the 1-pixel kernel;
// lanes differ only in fp32 rounding of reassociated terms, far below the B10G11R11 step.
using f2 = float2;
__device__ __forceinline__ f2 mk2(float a) { return make_float2(a, a); }
__device__ __forceinline__ f2 add2(f2 a, f2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ f2 rsqrt2(f2 a) { return make_float2(rsqrt_fast(a.x), rsqrt_fast(a.y)); }
__device__ __forceinline__ f2 clamp2(f2 a, float lo, float hi) { return make_float2(fminf(fmaxf(a.x, lo), hi), fminf(fmaxf(a.y, lo), hi)); }
__device__ __forceinline__ f2 dot3_2(f2 ax, f2 ay, f2 az, f2 bx, f2 by, f2 bz) { return fma2(az, bz, fma2(ay, by, mul2(ax, bx))); }

struct Surface2
{
	f2 npx, npy, npz; // -position
	f2 Nx, Ny, Nz, Vx, Vy, Vz;
	f2 F0x, F0y, F0z, oFx, oFy, oFz; // F0, 1 - F0
	f2 dkx, dky, dkz, ndkx, ndky, ndkz; // diffuse_k and its negation
	f2 m2m1, cgd, omk, k, Vk, NoVr;
};

// dk + F * (G*D - dk) per channel for both pixels; NoL returned for the caller's weight.
__device__ __forceinline__ void brdf2(const Surface2 &s, f2 NoLr, f2 VoL, f2 &NoL, f2 &tx, f2 &ty, f2 &tz)
{
	// NoLr = N.L and VoL = V.L for unit L; the half vector is never formed:
	// |V+L|^2 = 2 + 2 VoL,  N.(V+L) = NoV + NoL,  V.(V+L) = 1 + VoL
	f2 inv_h = rsqrt2(fma2(VoL, mk2(2.0f), mk2(2.0f)));
	NoL = clamp2(NoLr, 0.001f, 1.0f);
	f2 NoH = clamp2(mul2(add2(NoLr, s.NoVr), inv_h), 0.0001f, 1.0f);
	f2 HoV = fma2(VoL, inv_h, inv_h); // <= 1 up to rounding, see brdf()
	HoV = make_float2(fmaxf(HoV.x, 0.001f), fmaxf(HoV.y, 0.001f));
	f2 f = fma2(HoV, mk2(-1.0f), mk2(1.0f));
	f2 fsq = mul2(f, f);
	f2 f5 = mul2(mul2(fsq, fsq), f);
	f2 d = fma2(mul2(NoH, NoH), s.m2m1, mk2(1.0f));
	f2 vl = mul2(s.Vk, fma2(NoL, s.omk, s.k)); // >= 0.038, the shader's max(.., 1e-3) is the identity
	f2 den = mul2(mul2(d, d), vl);
	f2 GD = mul2(s.cgd, make_float2(rcp_fast(den.x), rcp_fast(den.y)));
	f2 Fx = fma2(s.oFx, f5, s.F0x), Fy = fma2(s.oFy, f5, s.F0y), Fz = fma2(s.oFz, f5, s.F0z);
	tx = fma2(Fx, add2(GD, s.ndkx), s.dkx);
	ty = fma2(Fy, add2(GD, s.ndky), s.dky);
	tz = fma2(Fz, add2(GD, s.ndkz), s.dkz);
}

struct PixelSetup
{
	bool lit;
	uint32_t dst, rx, ry;
	int cluster_base;
	float3 pos, N, V, F0, dk, base_color;
	float m2m1, cgd, omk, k, Vk, NoVr;
};

// Everything the 1-pixel kernel does before its light loop, for one pixel.
__device__ __forceinline__ PixelSetup setup_pixel(const LightingParams &p, const float *s_srgb, int x, int y, bool inside, float depth, uint32_t a8,
                                                 uint32_t n10, uint32_t mr, uint32_t emissive)
{
	PixelSetup q;
	q.lit = inside && depth != 0.0f;
	q.dst = emissive;
	q.rx = 0xffffffffu;
	q.ry = 0u;
	q.cluster_base = 0;
	q.pos = q.N = q.V = q.F0 = q.dk = q.base_color = make_float3(0.f, 0.f, 0.f);
	q.m2m1 = q.cgd = q.omk = q.k = q.Vk = q.NoVr = 0.0f;
	if (!q.lit)
		return q;
	q.base_color = make_float3(s_srgb[a8 & 0xffu], s_srgb[(a8 >> 8) & 0xffu], s_srgb[(a8 >> 16) & 0xffu]);
	q.N = make_float3(fmaf((float)(n10 & 0x3ffu), 2.0f / 1023.0f, -1.0f), fmaf((float)((n10 >> 10) & 0x3ffu), 2.0f / 1023.0f, -1.0f),
	                  fmaf((float)((n10 >> 20) & 0x3ffu), 2.0f / 1023.0f, -1.0f));
	const float metallic = (float)(mr & 0xffu) * (1.0f / 255.0f);
	const float roughness_in = (float)(mr >> 8) * (1.0f / 255.0f);
	int tile_index, z_index;
	q.pos = reconstruct_position_and_cluster(p, x, y, depth, tile_index, z_index);
	q.cluster_base = tile_index * p.n32;
	uint2 zr = __ldg(&p.cluster_range[z_index]);
	q.rx = zr.x;
	q.ry = zr.y;
	float3 v = make_float3(p.camera_pos.x - q.pos.x, p.camera_pos.y - q.pos.y, p.camera_pos.z - q.pos.z);
	float inv_v = rsqrt_fast(dot3(v, v));
	q.V = make_float3(v.x * inv_v, v.y * inv_v, v.z * inv_v);
	float rough = roughness_in * 0.75f + 0.25f;
	float mm = rough * rough;
	float m2 = mm * mm;
	q.m2m1 = m2 - 1.0f;
	q.cgd = m2 * (0.25f * kInvPi);
	float r1 = rough + 1.0f;
	q.k = r1 * r1 * 0.125f;
	q.omk = 1.0f - q.k;
	q.NoVr = dot3(q.N, q.V);
	float NoV = fminf(fmaxf(q.NoVr, 0.001f), 1.0f);
	q.Vk = NoV * q.omk + q.k;
	q.F0 = make_float3(0.04f * (1.0f - metallic) + q.base_color.x * metallic, 0.04f * (1.0f - metallic) + q.base_color.y * metallic,
	                   0.04f * (1.0f - metallic) + q.base_color.z * metallic);
	float dk = (1.0f - metallic) * kInvPi;
	q.dk = make_float3(q.base_color.x * dk, q.base_color.y * dk, q.base_color.z * dk);
	return q;
}

// G-buffer fetch + per-pixel invariants for the pixel pair (x, x + 1) of row y.
struct PairSetup
{
	PixelSetup A, B;
	Surface2 s;
	int z_start, z_end; // the warp's range of 32-light words (empty when z_end < z_start)
};

__device__ __forceinline__ void setup_pair(const LightingParams &p, const float *s_srgb, int x, int y, bool inside, PairSetup &q)
{
	float2 depth = make_float2(0.f, 0.f);
	uint2 a8 = make_uint2(0u, 0u), n10 = make_uint2(0u, 0u), em = make_uint2(0u, 0u);
	uint32_t mr2 = 0u;
	if (inside)
	{
		depth = __ldg(reinterpret_cast<const float2 *>(&p.depth.at(x, y)));
		// not gated on depth != 0: a dependent second round trip to HBM costs more than the sky's bytes
		a8 = __ldg(reinterpret_cast<const uint2 *>(&p.albedo.at(x, y)));
		n10 = __ldg(reinterpret_cast<const uint2 *>(&p.normal.at(x, y)));
		mr2 = __ldg(reinterpret_cast<const uint32_t *>(&p.pbr.at(x, y)));
		em = __ldg(reinterpret_cast<const uint2 *>(&p.emissive.at(x, y)));
	}
	q.A = setup_pixel(p, s_srgb, x, y, inside, depth.x, a8.x, n10.x, mr2 & 0xffffu, em.x);
	q.B = setup_pixel(p, s_srgb, x + 1, y, inside, depth.y, a8.y, n10.y, mr2 >> 16, em.y);
	const PixelSetup &A = q.A, &B = q.B;
	Surface2 &s = q.s;
	s.npx = make_float2(-A.pos.x, -B.pos.x); s.npy = make_float2(-A.pos.y, -B.pos.y); s.npz = make_float2(-A.pos.z, -B.pos.z);
	s.Nx = make_float2(A.N.x, B.N.x); s.Ny = make_float2(A.N.y, B.N.y); s.Nz = make_float2(A.N.z, B.N.z);
	s.Vx = make_float2(A.V.x, B.V.x); s.Vy = make_float2(A.V.y, B.V.y); s.Vz = make_float2(A.V.z, B.V.z);
	s.F0x = make_float2(A.F0.x, B.F0.x); s.F0y = make_float2(A.F0.y, B.F0.y); s.F0z = make_float2(A.F0.z, B.F0.z);
	s.oFx = make_float2(1.0f - A.F0.x, 1.0f - B.F0.x); s.oFy = make_float2(1.0f - A.F0.y, 1.0f - B.F0.y); s.oFz = make_float2(1.0f - A.F0.z, 1.0f - B.F0.z);
	s.dkx = make_float2(A.dk.x, B.dk.x); s.dky = make_float2(A.dk.y, B.dk.y); s.dkz = make_float2(A.dk.z, B.dk.z);
	s.ndkx = make_float2(-A.dk.x, -B.dk.x); s.ndky = make_float2(-A.dk.y, -B.dk.y); s.ndkz = make_float2(-A.dk.z, -B.dk.z);
	s.m2m1 = make_float2(A.m2m1, B.m2m1); s.cgd = make_float2(A.cgd, B.cgd);
	s.omk = make_float2(A.omk, B.omk); s.k = make_float2(A.k, B.k); s.Vk = make_float2(A.Vk, B.Vk);
	s.NoVr = make_float2(A.NoVr, B.NoVr);
	const uint32_t loA = A.rx >> 5, hiA = A.ry >> 5, loB = B.rx >> 5, hiB = B.ry >> 5; // unlit: (0x7ffffff, 0) => empty
	q.z_start = (int)__reduce_min_sync(0xffffffffu, min(loA, loB));
	q.z_end = min((int)__reduce_max_sync(0xffffffffu, max(A.lit ? hiA : 0u, B.lit ? hiB : 0u)), p.n32 - 1);
}

// draw 2 (clustering.frag): warp-uniform walk over the union of all 64 pixels' masks, words
// first, first + step, ... <= q.z_end; adds the lights' contribution to (accx, accy, accz).
__device__ __forceinline__ void walk_lights(const LightingParams &p, const PairSetup &q, int first, int step, float4 (*staged)[32], f2 &accx, f2 &accy,
                                            f2 &accz)
{
	const int lane = threadIdx.x & 31;
	const PixelSetup &A = q.A, &B = q.B;
	const Surface2 &s = q.s;
	const uint32_t loA = A.rx >> 5, hiA = A.ry >> 5, loB = B.rx >> 5, hiB = B.ry >> 5;
	for (int i = first; i <= q.z_end; i += step)
	{
		uint32_t ownA = 0u, ownB = 0u;
		if (A.lit && (uint32_t)i >= loA && (uint32_t)i <= hiA)
			ownA = cluster_mask_range(__ldg(&p.bitmask[A.cluster_base + i]), A.rx, A.ry, 32u * (uint32_t)i);
		if (B.lit && (uint32_t)i >= loB && (uint32_t)i <= hiB)
			ownB = cluster_mask_range(__ldg(&p.bitmask[B.cluster_base + i]), B.rx, B.ry, 32u * (uint32_t)i);
		uint32_t wmask = __reduce_or_sync(0xffffffffu, ownA | ownB);
		if (!wmask)
			continue;
		const uint32_t tm = __ldg(&p.type_mask[i]);
		// Stage the word's records in shared memory, lane b fetching light 32 i + b: one parallel trip
		// to L1/L2 per word.  Reading each record with a broadcast load right before its first use
		// (one dependent trip per light) was the largest stall in light-dense regions.
		__syncwarp();
		if ((wmask >> lane) & 1u)
		{
			const float4 *mine = reinterpret_cast<const float4 *>(p.lights + (i * 32 + lane));
			staged[0][lane] = __ldg(mine);
			staged[1][lane] = __ldg(mine + 1);
			staged[2][lane] = __ldg(mine + 2);
		}
		__syncwarp();
		while (wmask)
		{
			const int bit = __ffs(wmask) - 1;
			wmask &= wmask - 1u;
			const float4 l1 = staged[1][bit], l2 = staged[2][bit]; // position|offset_radius, direction|inv_radius
			f2 lx = add2(mk2(l1.x), s.npx), ly = add2(mk2(l1.y), s.npy), lz = add2(mk2(l1.z), s.npz);
			f2 d2 = dot3_2(lx, ly, lz, lx, ly, lz);
			const float inv_r2 = l2.w * l2.w;
			const bool nearA = ((ownA >> bit) & 1u) && (d2.x * inv_r2 < 1.0f);
			const bool nearB = ((ownB >> bit) & 1u) && (d2.y * inv_r2 < 1.0f);
			// quick reject: beyond the light's radius the falloff is exactly 0 (point.h:41-43)
			if (!__any_sync(0xffffffffu, nearA || nearB))
				continue;
			const float4 l0 = staged[0][bit]; // color|spot scale_bias
			f2 inv_d = rsqrt2(d2);
			f2 inv_ld = make_float2(fminf(inv_d.x, 10.0f), fminf(inv_d.y, 10.0f));
			f2 dist = mul2(d2, inv_d);
			f2 xr = mul2(make_float2(fmaxf(dist.x, 0.1f), fmaxf(dist.y, 0.1f)), mk2(l2.w));
			f2 t = fma2(xr, mk2(1.0f / (1.0f - 0.9f)), mk2(-0.9f / (1.0f - 0.9f)));
			t = make_float2(__saturatef(t.x), __saturatef(t.y));
			f2 falloff = fma2(mul2(mul2(t, t), fma2(mk2(-2.0f), t, mk2(3.0f))), mk2(-1.0f), mk2(1.0f));
			// L = l * inv_d is never formed either: every use of it is a dot product
			if (!((tm >> bit) & 1u))
			{
				float2 sb = __half22float2(*reinterpret_cast<const __half2 *>(&l0.w));
				f2 cone_angle = mul2(dot3_2(lx, ly, lz, mk2(l2.x), mk2(l2.y), mk2(l2.z)), inv_d);
				f2 cone = fma2(cone_angle, mk2(-sb.x), mk2(sb.y));
				cone = make_float2(__saturatef(cone.x), __saturatef(cone.y));
				falloff = mul2(falloff, mul2(cone, cone));
			}
			f2 NoLr = mul2(dot3_2(s.Nx, s.Ny, s.Nz, lx, ly, lz), inv_d);
			f2 VoL = mul2(dot3_2(s.Vx, s.Vy, s.Vz, lx, ly, lz), inv_d);
			f2 NoL, tx, ty, tz;
			brdf2(s, NoLr, VoL, NoL, tx, ty, tz);
			f2 w = mul2(mul2(NoL, falloff), mul2(inv_ld, inv_ld));
			// a lane outside the light's mask or radius adds exactly 0 (falloff is 0 beyond the radius)
			w = make_float2(nearA ? w.x : 0.0f, nearB ? w.y : 0.0f);
			accx = fma2(mul2(mk2(l0.x), w), tx, accx);
			accy = fma2(mul2(mk2(l0.y), w), ty, accy);
			accz = fma2(mul2(mk2(l0.z), w), tz, accz);
		}
	}
}

__global__ void __launch_bounds__(32 * kWarpsPerCta, 5) deferred_lighting2_kernel(const LightingParams p)
{
	__shared__ float s_srgb[256];
	for (int i = threadIdx.x; i < 256; i += blockDim.x)
		s_srgb[i] = g_srgb8_to_linear[i];
	__syncthreads();

	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int x = ((blockIdx.x * kWarpsPerCta + warp) * 8 + (lane & 7)) * 2; // pixels x and x + 1 (width is even on this path)
	const int y = p.y0 + blockIdx.y * 4 + (lane >> 3);
	const bool inside = x < p.hdr.w && y < p.y1;

	__shared__ float4 s_lights[kWarpsPerCta][3][32];
	PairSetup q;
	setup_pair(p, s_srgb, x, y, inside, q);
	const PixelSetup &A = q.A, &B = q.B;

	// ---- draw 1: directional light for both pixels ----
	uint32_t dstA = A.dst, dstB = B.dst;
	{
		const Surface2 &s = q.s;
		f2 NoL, tx, ty, tz;
		const f2 dx = mk2(p.dir_dir.x), dy = mk2(p.dir_dir.y), dz = mk2(p.dir_dir.z);
		brdf2(s, dot3_2(s.Nx, s.Ny, s.Nz, dx, dy, dz), dot3_2(s.Vx, s.Vy, s.Vz, dx, dy, dz), NoL, tx, ty, tz);
		if (A.lit)
		{
			float3 e = unpack_r11g11b10(dstA);
			dstA = pack_r11g11b10(e.x + p.dir_color.x * NoL.x * tx.x + A.base_color.x * 0.05f, e.y + p.dir_color.y * NoL.x * ty.x + A.base_color.y * 0.05f,
			                      e.z + p.dir_color.z * NoL.x * tz.x + A.base_color.z * 0.05f);
		}
		if (B.lit)
		{
			float3 e = unpack_r11g11b10(dstB);
			dstB = pack_r11g11b10(e.x + p.dir_color.x * NoL.y * tx.y + B.base_color.x * 0.05f, e.y + p.dir_color.y * NoL.y * ty.y + B.base_color.y * 0.05f,
			                      e.z + p.dir_color.z * NoL.y * tz.y + B.base_color.z * 0.05f);
		}
	}

	// ---- draw 2 ----
	// (Handing the light-dense blocks -- hundreds of lights per pixel, one warp busy for ~100 us --
	// to a second kernel that spreads a block's lights over four warps was tried: bit-compatible,
	// but the frame got 9 % slower, because the dense blocks then no longer overlap the cheap ones.)
	f2 accx = mk2(0.0f), accy = mk2(0.0f), accz = mk2(0.0f);
	walk_lights(p, q, q.z_start, 1, s_lights[warp], accx, accy, accz);
	const float3 accA = make_float3(accx.x, accy.x, accz.x), accB = make_float3(accx.y, accy.y, accz.y);

	if (inside)
	{
		uint2 out = make_uint2(dstA, dstB);
		if (A.lit)
		{
			float3 e = unpack_r11g11b10(dstA);
			out.x = pack_r11g11b10(e.x + accA.x, e.y + accA.y, e.z + accA.z);
		}
		if (B.lit)
		{
			float3 e = unpack_r11g11b10(dstB);
			out.y = pack_r11g11b10(e.x + accB.x, e.y + accB.y, e.z + accB.z);
		}
		// sky pixels carry the emissive value through (identical bits when blending in place)
		if (A.lit || B.lit || p.emissive.p != p.hdr.p)
			*reinterpret_cast<uint2 *>(&p.hdr.at(x, y)) = out;
	}
}

// ---------------------------------------------------------------------------------------------
// Work estimate of the lighting pass per group of 4 pixel rows (one row of CTAs), in issued warp
// instructions.  The light density of a frame is far from uniform (in the bench scene 3 % of the
// rows hold over half of the light evaluations), so equal-height row bands do not split the pass
// equally between GPUs.  This kernel repeats the lighting kernel's cluster walk for the same
// 16x4 pixel blocks -- position, (tile, Z slice), word range, union of the lanes' masks, radius
// test -- without shading, and charges each step what the shading kernel's SASS spends on it.
constexpr uint32_t kCostSetup = 700u;   // per warp: G-buffer decode, position, BRDF invariants, directional light, stores
constexpr uint32_t kCostWord = 45u;     // per 32-light word of the Z range: two mask loads, range masks, warp OR
constexpr uint32_t kCostUnion = 29u;    // per light in the warp's union: record load, distance, radius vote
constexpr uint32_t kCostEvaluate = 89u; // per light that reaches a pixel of the block: falloff, cone, BRDF, accumulate

__global__ void __launch_bounds__(32 * kWarpsPerCta) lighting_cost_kernel(const LightingParams p, uint32_t *__restrict__ cost)
{
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int x = ((blockIdx.x * kWarpsPerCta + warp) * 8 + (lane & 7)) * 2;
	const int y = p.y0 + blockIdx.y * 4 + (lane >> 3);
	if (__all_sync(0xffffffffu, x >= p.depth.w || y >= p.y1))
		return;
	float3 pos[2];
	uint32_t rx[2] = { 0xffffffffu, 0xffffffffu }, ry[2] = { 0u, 0u };
	int base[2] = { 0, 0 };
	bool lit[2] = { false, false };
#pragma unroll
	for (int k = 0; k < 2; k++)
	{
		pos[k] = make_float3(0.f, 0.f, 0.f);
		if (x + k < p.depth.w && y < p.y1)
		{
			const float depth = __ldg(&p.depth.at(x + k, y));
			if (depth != 0.0f)
			{
				int tile_index, z_index;
				pos[k] = reconstruct_position_and_cluster(p, x + k, y, depth, tile_index, z_index);
				base[k] = tile_index * p.n32;
				const uint2 zr = __ldg(&p.cluster_range[z_index]);
				rx[k] = zr.x;
				ry[k] = zr.y;
				lit[k] = true;
			}
		}
	}
	const uint32_t lo0 = rx[0] >> 5, hi0 = ry[0] >> 5, lo1 = rx[1] >> 5, hi1 = ry[1] >> 5;
	int z_start = (int)__reduce_min_sync(0xffffffffu, min(lo0, lo1));
	int z_end = (int)__reduce_max_sync(0xffffffffu, max(lit[0] ? hi0 : 0u, lit[1] ? hi1 : 0u));
	z_end = min(z_end, p.n32 - 1);
	uint32_t total = kCostSetup;
	for (int i = z_start; i <= z_end; i++)
	{
		uint32_t own0 = 0u, own1 = 0u;
		if (lit[0] && (uint32_t)i >= lo0 && (uint32_t)i <= hi0)
			own0 = cluster_mask_range(__ldg(&p.bitmask[base[0] + i]), rx[0], ry[0], 32u * (uint32_t)i);
		if (lit[1] && (uint32_t)i >= lo1 && (uint32_t)i <= hi1)
			own1 = cluster_mask_range(__ldg(&p.bitmask[base[1] + i]), rx[1], ry[1], 32u * (uint32_t)i);
		uint32_t wmask = __reduce_or_sync(0xffffffffu, own0 | own1);
		total += kCostWord;
		while (wmask)
		{
			const int bit = __ffs(wmask) - 1;
			wmask &= wmask - 1u;
			const float4 *lp = reinterpret_cast<const float4 *>(p.lights + (i * 32 + bit));
			const float4 l1 = __ldg(lp + 1), l2 = __ldg(lp + 2);
			const float inv_r2 = l2.w * l2.w;
			const float3 a = make_float3(l1.x - pos[0].x, l1.y - pos[0].y, l1.z - pos[0].z);
			const float3 b = make_float3(l1.x - pos[1].x, l1.y - pos[1].y, l1.z - pos[1].z);
			const bool near = (((own0 >> bit) & 1u) && dot3(a, a) * inv_r2 < 1.0f) || (((own1 >> bit) & 1u) && dot3(b, b) * inv_r2 < 1.0f);
			total += __any_sync(0xffffffffu, near) ? (kCostUnion + kCostEvaluate) : kCostUnion;
		}
	}
	if (lane == 0)
		atomicAdd(&cost[blockIdx.y], total);
}
} // namespace

int32_t upload_srgb_lut(const float *lut256)
{
	cudaError_t err = cudaMemcpyToSymbol(g_srgb8_to_linear, lut256, 256 * sizeof(float));
	if (err != cudaSuccess)
	{
		set_last_error(cudaGetErrorString(err));
		return GRB_ERR_CUDA;
	}
	return GRB_OK;
}
} // namespace grb

using namespace grb;

extern "C" int32_t grb_deferred_lighting(const GrbGBuffer *g, const GrbCamera *cam, const GrbClusterParameters *params, const GrbClusterBuffers *buf,
                                         const GrbImage *hdr, GrbRows rows, void *stream)
{
	if (!g || !cam || !params || !buf || !hdr)
	{
		set_last_error("grb_deferred_lighting: null argument");
		return GRB_ERR_INVALID_ARGUMENT;
	}
	if (!image_ok(&g->albedo, GRB_FORMAT_R8G8B8A8_SRGB, 4) || !image_ok(&g->normal, GRB_FORMAT_A2B10G10R10_UNORM_PACK32, 4) ||
	    !image_ok(&g->pbr, GRB_FORMAT_R8G8_UNORM, 2) || !image_ok(&g->depth, GRB_FORMAT_D32_SFLOAT, 4) ||
	    !image_ok(hdr, GRB_FORMAT_B10G11R11_UFLOAT_PACK32, 4))
	{
		set_last_error("grb_deferred_lighting: G-buffer must be R8G8B8A8_SRGB / A2B10G10R10_UNORM / R8G8_UNORM / D32_SFLOAT, hdr B10G11R11_UFLOAT");
		return GRB_ERR_UNSUPPORTED_FORMAT;
	}
	const int w = hdr->width, h = hdr->height;
	if (g->albedo.width != w || g->albedo.height != h || g->normal.width != w || g->normal.height != h || g->pbr.width != w || g->pbr.height != h ||
	    g->depth.width != w || g->depth.height != h)
	{
		set_last_error("grb_deferred_lighting: attachment sizes differ");
		return GRB_ERR_INVALID_ARGUMENT;
	}
	if (params->num_lights > 0 && (!buf->lights || !buf->type_mask || !buf->bitmask))
	{
		set_last_error("grb_deferred_lighting: null cluster buffer");
		return GRB_ERR_INVALID_ARGUMENT;
	}
	if (!buf->cluster_range)
	{
		set_last_error("grb_deferred_lighting: null cluster_range");
		return GRB_ERR_INVALID_ARGUMENT;
	}
	rows = full_rows(rows, h);
	if (rows.y1 <= rows.y0)
		return GRB_OK;

	LightingParams p;
	p.albedo = view_of<const uint32_t>(&g->albedo);
	p.normal = view_of<const uint32_t>(&g->normal);
	p.pbr = view_of<const uint16_t>(&g->pbr);
	p.depth = view_of<const float>(&g->depth);
	p.hdr = view_of<uint32_t>(hdr);
	if (g->emissive.data)
	{
		if (!image_ok(&g->emissive, GRB_FORMAT_B10G11R11_UFLOAT_PACK32, 4) || g->emissive.width != w || g->emissive.height != h)
		{
			set_last_error("grb_deferred_lighting: emissive must be B10G11R11_UFLOAT of the G-buffer's size");
			return GRB_ERR_UNSUPPORTED_FORMAT;
		}
		p.emissive = view_of<const uint32_t>(&g->emissive);
	}
	else
		p.emissive = view_of<const uint32_t>(hdr);
	for (int i = 0; i < 16; i++)
		p.ivp[i] = cam->inv_view_projection[i];
	p.camera_pos = make_float3(cam->camera_position[0], cam->camera_position[1], cam->camera_position[2]);
	p.dir_color = make_float3(g->directional_color[0], g->directional_color[1], g->directional_color[2]);
	p.dir_dir = make_float3(g->directional_direction[0], g->directional_direction[1], g->directional_direction[2]);
	p.cbase = make_float3(params->camera_base[0], params->camera_base[1], params->camera_base[2]);
	p.cfront = make_float3(params->camera_front[0], params->camera_front[1], params->camera_front[2]);
	p.xy_scale = make_float2(params->xy_scale[0], params->xy_scale[1]);
	p.res_x = params->resolution_xy[0];
	p.res_y = params->resolution_xy[1];
	p.n32 = params->num_lights_32;
	p.z_max_index = params->z_max_index;
	p.z_scale = params->z_scale;
	p.inv_res_x = 1.0f / (float)w; // renderer.cpp:1101-1102,1120
	p.inv_res_y = 1.0f / (float)h;
	p.lights = buf->lights;
	p.type_mask = buf->type_mask;
	p.bitmask = buf->bitmask;
	p.cluster_range = reinterpret_cast<const uint2 *>(buf->cluster_range);
	p.y0 = rows.y0;
	p.y1 = rows.y1;

	// two pixels per thread (packed fp32) whenever rows can be addressed as aligned pixel pairs
	static const bool force_1px = getenv("GRB_LIGHTING_1PX") != nullptr;
	auto aligned8 = [](const void *ptr, int pitch_bytes) { return (reinterpret_cast<uintptr_t>(ptr) % 8) == 0 && (pitch_bytes % 8) == 0; };
	const bool pairs = !force_1px && (w % 2) == 0 && aligned8(g->albedo.data, g->albedo.row_pitch) && aligned8(g->normal.data, g->normal.row_pitch) &&
	                   aligned8(g->depth.data, g->depth.row_pitch) && (reinterpret_cast<uintptr_t>(g->pbr.data) % 4) == 0 && (g->pbr.row_pitch % 4) == 0 &&
	                   aligned8(hdr->data, hdr->row_pitch) && (!g->emissive.data || aligned8(g->emissive.data, g->emissive.row_pitch));
	if (pairs)
	{
		dim3 grid2((w / 2 + 8 * kWarpsPerCta - 1) / (8 * kWarpsPerCta), (rows.y1 - rows.y0 + 3) / 4, 1);
		deferred_lighting2_kernel<<<grid2, 32 * kWarpsPerCta, 0, as_stream(stream)>>>(p);
		return check_launch("grb_deferred_lighting");
	}
	dim3 grid((w + 8 * kWarpsPerCta - 1) / (8 * kWarpsPerCta), (rows.y1 - rows.y0 + 3) / 4, 1);
	deferred_lighting_kernel<<<grid, 32 * kWarpsPerCta, 0, as_stream(stream)>>>(p);
	return check_launch("grb_deferred_lighting");
}

extern "C" int32_t grb_debug_cluster_indices(const GrbImage *depth, const GrbCamera *cam, const GrbClusterParameters *params, int32_t *out_tile,
                                             int32_t *out_z, GrbRows rows, void *stream)
{
	if (!image_ok(depth, GRB_FORMAT_D32_SFLOAT, 4) || !cam || !params || !out_tile || !out_z)
	{
		set_last_error("grb_debug_cluster_indices: bad arguments");
		return GRB_ERR_INVALID_ARGUMENT;
	}
	rows = full_rows(rows, depth->height);
	if (rows.y1 <= rows.y0)
		return GRB_OK;
	LightingParams p{};
	p.depth = view_of<const float>(depth);
	for (int i = 0; i < 16; i++)
		p.ivp[i] = cam->inv_view_projection[i];
	p.cbase = make_float3(params->camera_base[0], params->camera_base[1], params->camera_base[2]);
	p.cfront = make_float3(params->camera_front[0], params->camera_front[1], params->camera_front[2]);
	p.xy_scale = make_float2(params->xy_scale[0], params->xy_scale[1]);
	p.res_x = params->resolution_xy[0];
	p.res_y = params->resolution_xy[1];
	p.n32 = params->num_lights_32;
	p.z_max_index = params->z_max_index;
	p.z_scale = params->z_scale;
	p.inv_res_x = 1.0f / (float)depth->width;
	p.inv_res_y = 1.0f / (float)depth->height;
	p.y0 = rows.y0;
	p.y1 = rows.y1;
	dim3 grid((depth->width + 31) / 32, (rows.y1 - rows.y0 + 3) / 4, 1);
	cluster_indices_kernel<<<grid, 128, 0, as_stream(stream)>>>(p, out_tile, out_z);
	return check_launch("grb_debug_cluster_indices");
}

extern "C" int32_t grb_lighting_row_cost(const GrbImage *depth, const GrbCamera *cam, const GrbClusterParameters *params, const GrbClusterBuffers *buf,
                                         GrbRows rows, uint32_t *cost_per_4_rows, void *stream)
{
	if (!image_ok(depth, GRB_FORMAT_D32_SFLOAT, 4) || !cam || !params || !buf || !cost_per_4_rows)
	{
		set_last_error("grb_lighting_row_cost: bad arguments");
		return GRB_ERR_INVALID_ARGUMENT;
	}
	if (!buf->cluster_range || (params->num_lights > 0 && (!buf->lights || !buf->bitmask)))
	{
		set_last_error("grb_lighting_row_cost: null cluster buffer");
		return GRB_ERR_INVALID_ARGUMENT;
	}
	rows = full_rows(rows, depth->height);
	if (rows.y1 <= rows.y0)
		return GRB_OK;
	const int groups = (rows.y1 - rows.y0 + 3) / 4;
	cudaError_t err = cudaMemsetAsync(cost_per_4_rows, 0, sizeof(uint32_t) * (size_t)groups, as_stream(stream));
	if (err != cudaSuccess)
	{
		set_last_error(cudaGetErrorString(err));
		return GRB_ERR_CUDA;
	}
	LightingParams p{};
	p.depth = view_of<const float>(depth);
	for (int i = 0; i < 16; i++)
		p.ivp[i] = cam->inv_view_projection[i];
	p.cbase = make_float3(params->camera_base[0], params->camera_base[1], params->camera_base[2]);
	p.cfront = make_float3(params->camera_front[0], params->camera_front[1], params->camera_front[2]);
	p.xy_scale = make_float2(params->xy_scale[0], params->xy_scale[1]);
	p.res_x = params->resolution_xy[0];
	p.res_y = params->resolution_xy[1];
	p.n32 = params->num_lights_32;
	p.z_max_index = params->z_max_index;
	p.z_scale = params->z_scale;
	p.inv_res_x = 1.0f / (float)depth->width;
	p.inv_res_y = 1.0f / (float)depth->height;
	p.lights = buf->lights;
	p.bitmask = buf->bitmask;
	p.cluster_range = reinterpret_cast<const uint2 *>(buf->cluster_range);
	p.y0 = rows.y0;
	p.y1 = rows.y1;
	const int pairs = (depth->width + 1) / 2;
	dim3 grid((pairs + 8 * kWarpsPerCta - 1) / (8 * kWarpsPerCta), groups, 1);
	lighting_cost_kernel<<<grid, 32 * kWarpsPerCta, 0, as_stream(stream)>>>(p, cost_per_4_rows);
	return check_launch("grb_lighting_row_cost");
}
