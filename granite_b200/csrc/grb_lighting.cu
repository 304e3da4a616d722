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
	float3 diffuse_k;   // base_color * (1 - metallic) / PI
	float m2;           // roughness'^4
	float one_minus_k, k, Vk; // Schlick-GGX visibility pieces
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
// point.h:121-141, spot.h:124-144).  Returns (specular + diffuse) * NoL, to be scaled by the
// light's colour/attenuation.
__device__ __forceinline__ float3 brdf(const Surface &s, float3 L)
{
	float3 h = make_float3(s.V.x + L.x, s.V.y + L.y, s.V.z + L.z);
	float inv_h = rsqrtf(dot3(h, h));
	float NoL = fminf(fmaxf(dot3(s.N, L), 0.001f), 1.0f);
	float NoH = fminf(fmaxf(dot3(s.N, h) * inv_h, 0.0001f), 1.0f);
	float HoV = fminf(fmaxf(dot3(h, s.V) * inv_h, 0.001f), 1.0f);
	float f = 1.0f - HoV;
	float f2 = f * f;
	float f5 = f2 * f2 * f;
	float3 F = make_float3(s.F0.x + s.one_minus_F0.x * f5, s.F0.y + s.one_minus_F0.y * f5, s.F0.z + s.one_minus_F0.z * f5);
	float d = (NoH * s.m2 - NoH) * NoH + 1.0f;
	float D = __fdividef(s.m2, kPi * d * d);
	float G = __fdividef(0.25f, fmaxf(s.Vk * (NoL * s.one_minus_k + s.k), 0.001f));
	float GD = G * D;
	return make_float3(NoL * (F.x * GD + (1.0f - F.x) * s.diffuse_k.x), NoL * (F.y * GD + (1.0f - F.y) * s.diffuse_k.y),
	                   NoL * (F.z * GD + (1.0f - F.z) * s.diffuse_k.z));
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
		s.N = make_float3(fsub(fmul(fdiv((float)(n10 & 0x3ffu), 1023.0f), 2.0f), 1.0f), fsub(fmul(fdiv((float)((n10 >> 10) & 0x3ffu), 1023.0f), 2.0f), 1.0f),
		                  fsub(fmul(fdiv((float)((n10 >> 20) & 0x3ffu), 1023.0f), 2.0f), 1.0f));
		const float metallic = fdiv((float)(mr & 0xffu), 255.0f);
		const float roughness_in = fdiv((float)(mr >> 8), 255.0f);

		int tile_index, z_index;
		s.pos = reconstruct_position_and_cluster(p, x, y, depth, tile_index, z_index);
		cluster_base = tile_index * p.n32;
		uint2 zr = __ldg(&p.cluster_range[z_index]);
		rx = zr.x;
		ry = zr.y;

		// per-pixel BRDF invariants
		float3 v = make_float3(p.camera_pos.x - s.pos.x, p.camera_pos.y - s.pos.y, p.camera_pos.z - s.pos.z);
		float inv_v = rsqrtf(dot3(v, v));
		s.V = make_float3(v.x * inv_v, v.y * inv_v, v.z * inv_v);
		float rough = roughness_in * 0.75f + 0.25f;
		float mm = rough * rough;
		s.m2 = mm * mm;
		float r1 = rough + 1.0f;
		s.k = r1 * r1 * 0.125f;
		s.one_minus_k = 1.0f - s.k;
		float NoV = fminf(fmaxf(dot3(s.N, s.V), 0.001f), 1.0f);
		s.Vk = NoV * s.one_minus_k + s.k;
		s.F0 = make_float3(0.04f * (1.0f - metallic) + base_color.x * metallic, 0.04f * (1.0f - metallic) + base_color.y * metallic,
		                   0.04f * (1.0f - metallic) + base_color.z * metallic);
		s.one_minus_F0 = make_float3(1.0f - s.F0.x, 1.0f - s.F0.y, 1.0f - s.F0.z);
		float dk = (1.0f - metallic) * kInvPi;
		s.diffuse_k = make_float3(base_color.x * dk, base_color.y * dk, base_color.z * dk);

		// ---- draw 1: directional.frag (LIGHTING_NO_AMBIENT, no shadows, VOLUMETRIC_DIFFUSE_FALLBACK) ----
		float3 b = brdf(s, p.dir_dir);
		float3 e = unpack_r11g11b10(dst);
		dst = pack_r11g11b10(e.x + p.dir_color.x * b.x + base_color.x * 0.05f, e.y + p.dir_color.y * b.y + base_color.y * 0.05f,
		                     e.z + p.dir_color.z * b.z + base_color.z * 0.05f);
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
			const bool mine = (own >> bit) & 1u;
			const float4 *lp = reinterpret_cast<const float4 *>(p.lights + (i * 32 + bit));
			const float4 l0 = __ldg(lp), l1 = __ldg(lp + 1), l2 = __ldg(lp + 2); // color|scale_bias, position|offset_radius, direction|inv_radius
			float3 l = make_float3(l1.x - s.pos.x, l1.y - s.pos.y, l1.z - s.pos.z);
			float d2 = dot3(l, l);
			float inv_d = rsqrtf(d2);
			float light_dist = fmaxf(0.1f, d2 * inv_d);
			float t = __saturatef((light_dist * l2.w - 0.9f) * (1.0f / (1.0f - 0.9f)));
			float falloff = 1.0f - t * t * (3.0f - 2.0f * t);
			float3 L = make_float3(l.x * inv_d, l.y * inv_d, l.z * inv_d);
			if (!((tm >> bit) & 1u))
			{
				// spot.h:34-84: cone term from the packed fp16 scale/bias
				float2 sb = __half22float2(*reinterpret_cast<const __half2 *>(&l0.w));
				float cone_angle = -(L.x * l2.x + L.y * l2.y + L.z * l2.z);
				float cone = __saturatef(cone_angle * sb.x + sb.y);
				falloff *= cone * cone;
			}
			const bool contributes = mine && falloff > 0.0f;
			if (__any_sync(0xffffffffu, contributes))
			{
				float atten = __fdividef(falloff, light_dist * light_dist);
				float3 b = brdf(s, L);
				if (contributes)
				{
					acc.x += l0.x * atten * b.x;
					acc.y += l0.y * atten * b.y;
					acc.z += l0.z * atten * b.z;
				}
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
