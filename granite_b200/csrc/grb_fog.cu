// grb_fog.cu -- volumetric fog, accumulation pass: VolumetricFog::build_fog
// (renderer/lights/volumetric_fog.cpp:236-254) dispatching assets/shaders/lights/fog_accumulate.comp -- the froxel grid
// of in-scattered light and optical depth (R16G16B16A16_SFLOAT, w x h x d) is integrated front to back along every
// view ray: per slice a 17-tap blur (the 3 x 3 neighbourhood of the slice and the 8 non-centre neighbours of the slice
// in front), then light += back.rgb * (exp2(-depth) * back.a), depth += back.a, store (light, exp2(-depth)).
//
// One thread per (x, y) column, walking z -- the recurrence is sequential in z and the columns are independent.  A thread
// keeps the 3 x 3 neighbourhood of the previous slice in registers: a slice's nine texels are loaded once and serve two
// steps (as the "slice" taps of step z and as the "slice in front" taps of step z + 1), 9 instead of 17 eight-byte loads
// per step, all but the column's own out of L1 (neighbouring threads read the same texels).  Compulsory traffic is
// 16 B per froxel (read + write): 15 MB at the reference's 160 x 92 x 64 grid -- a few microseconds of HBM time on a
// grid of only 14 720 threads, so the pass is latency-bound by the z walk, not bandwidth-bound; a scan along z would
// parallelise it and is not worth its weight here.
//
// fp32 in the shader's order (built with -fmad=false), weights as glslang folds them; exp2f is CUDA's (<= 2 ulp from the
// oracle's glibc value: the stored fp16 results are compared to 1 fp16 ulp on hardware, exactly under CPU emulation).
#include "grb_common.cuh"

namespace grb
{
namespace
{
struct Vol16
{
	const uint2 *p;
	int w, h, d;
};

GRB_DEV float4 vol_texel(const Vol16 &v, int x, int y, int z)
{
	x = iclamp(x, 0, v.w - 1);
	y = iclamp(y, 0, v.h - 1);
	z = iclamp(z, 0, v.d - 1);
	return unpack_rgba16f(__ldg(v.p + ((size_t)z * v.h + y) * v.w + x));
}

GRB_DEV void tap(float4 &acc, float w, const float4 &t)
{
	acc.x += w * t.x;
	acc.y += w * t.y;
	acc.z += w * t.z;
	acc.w += w * t.w;
}

// fog_accumulate.comp:27-63
__global__ void __launch_bounds__(256) fog_accumulate_kernel(Vol16 light, uint2 *__restrict__ fog)
{
	const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
	if (x >= light.w || y >= light.h)
		return;
	const float w3 = (float)(1.0 / (1.375 * 32.0)), w2 = (float)(1.0 / (1.375 * 16.0)), w1 = (float)(1.0 / (1.375 * 8.0)), w0 = (float)(1.0 / (1.375 * 4.0));
	// neighbourhood index: n[j][i] = texel (x + i - 1, y + j - 1) of a slice
	float4 prev[3][3], cur[3][3];
#pragma unroll
	for (int j = 0; j < 3; j++)
#pragma unroll
		for (int i = 0; i < 3; i++)
			prev[j][i] = vol_texel(light, x + i - 1, y + j - 1, -1); // slice "-1" clamps to slice 0
	float4 front = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
	for (int z = 0; z < light.d; z++)
	{
#pragma unroll
		for (int j = 0; j < 3; j++)
#pragma unroll
			for (int i = 0; i < 3; i++)
				cur[j][i] = vol_texel(light, x + i - 1, y + j - 1, z);
		float4 back = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
		tap(back, w0, cur[1][1]);
		tap(back, w2, prev[0][1]); // (0, -1, -1)
		tap(back, w2, prev[1][0]); // (-1, 0, -1)
		tap(back, w2, prev[1][2]); // (1, 0, -1)
		tap(back, w2, prev[2][1]); // (0, +1, -1)
		tap(back, w3, prev[0][0]); // (-1, -1, -1)
		tap(back, w3, prev[0][2]); // (1, -1, -1)
		tap(back, w3, prev[2][0]); // (-1, +1, -1)
		tap(back, w3, prev[2][2]); // (1, +1, -1)
		tap(back, w1, cur[0][1]);  // (0, -1, 0)
		tap(back, w1, cur[1][0]);  // (-1, 0, 0)
		tap(back, w1, cur[1][2]);  // (1, 0, 0)
		tap(back, w1, cur[2][1]);  // (0, +1, 0)
		tap(back, w2, cur[0][2]);  // (1, -1, 0)
		tap(back, w2, cur[0][0]);  // (-1, -1, 0)
		tap(back, w2, cur[2][0]);  // (-1, +1, 0)
		tap(back, w2, cur[2][2]);  // (1, +1, 0)
		const float s = exp2f(-front.w) * back.w; // accumulate_scattering, .comp:17-22
		front = make_float4(front.x + back.x * s, front.y + back.y * s, front.z + back.z * s, front.w + back.w);
		fog[((size_t)z * light.h + y) * light.w + x] = pack_rgba16f(make_float4(front.x, front.y, front.z, exp2f(-front.w)));
#pragma unroll
		for (int j = 0; j < 3; j++)
#pragma unroll
			for (int i = 0; i < 3; i++)
				prev[j][i] = cur[j][i];
	}
}

// ------------------------------------------------------------------------------------------------
// First pass: fog_light_density.comp in its base variant (VolumetricFog::build_light_density, volumetric_fog.cpp:142-228:
// no fog regions -> density 0.1, no temporal reprojection = the first frame, no floor lighting, unshadowed lights).  One
// thread per froxel (x fastest: contiguous 8-byte stores); the clustered in-scatter walks the froxel's own (tile, slice)
// mask like the lighting pass does per pixel (clusterer_bindless.h:158-203).  942 080 froxels at 160 x 92 x 64, 8 bytes
// written each: the pass is bound by its arithmetic (two exp2 / sqrt / divisions per froxel plus the light walk), not by
// the 7.5 MB it writes.  Exact fp32 in the shader's order (no fast intrinsics: the grid is 1/9 of a 4K frame's pixels).
struct FogDensityArgs
{
	int w, h, d, dither_offset;
	float slice_z_log2_scale, density_mod, in_scatter_strength;
	float ivp[16];        // inv_view_projection
	float zt[4];          // (projection[2].zw, projection[3].zw)
	float xy_scale[2];    // (inv_projection[0].x, inv_projection[1].y)
	float camera_pos[3], dir_color[3], dir_direction[3];
	// cluster
	float ctransform[16], cbase[3], cfront[3], cxy_scale[2];
	int res_x, res_y, n32, z_max_index;
	float z_scale;
	const GrbPositionalLight *lights;
	const uint32_t *type_mask, *bitmask;
	const uint2 *cluster_range;
	const float *slice_extents;
	const uint32_t *dither_lut; // N x 128 x 128 R8G8B8A8_UNORM
};

GRB_DEV float3 sub3(float3 a, float3 b) { return make_float3(a.x - b.x, a.y - b.y, a.z - b.z); }
GRB_DEV float dot3f(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
GRB_DEV float3 normalize3(float3 a)
{
	const float inv = 1.0f / sqrtf(dot3f(a, a));
	return make_float3(a.x * inv, a.y * inv, a.z * inv);
}
GRB_DEV float smoothstep_f(float e0, float e1, float x)
{
	const float t = fclamp((x - e0) / (e1 - e0), 0.0f, 1.0f);
	return t * t * (3.0f - 2.0f * t);
}
GRB_DEV int f2i_sat(float v) { return v >= 2147483520.0f ? 2147483647 : (v <= -2147483648.0f ? (-2147483647 - 1) : (v != v ? 0 : (int)v)); }
GRB_DEV float4 mul_pairwise(const float *m, float x, float y, float z, float w)
{
	return make_float4((m[0] * x + m[4] * y) + (m[8] * z + m[12] * w), (m[1] * x + m[5] * y) + (m[9] * z + m[13] * w),
	                   (m[2] * x + m[6] * y) + (m[10] * z + m[14] * w), (m[3] * x + m[7] * y) + (m[11] * z + m[15] * w));
}

// point.h:33-81 / spot.h:34-84 without shadows: the light's colour at world_pos
GRB_DEV float3 positional_color(const GrbPositionalLight &l, bool is_point, float3 world_pos)
{
	const float3 lpos = make_float3(l.position[0], l.position[1], l.position[2]);
	if (is_point)
	{
		const float3 full = sub3(world_pos, lpos);
		const float dist = fmaxf(0.1f, sqrtf(dot3f(full, full)));
		const float falloff = 1.0f - smoothstep_f(0.9f, 1.0f, dist * l.inv_radius);
		if (!(falloff > 0.0f))
			return make_float3(0.0f, 0.0f, 0.0f);
		const float s = 1.0f * falloff, d2 = dist * dist;
		return make_float3(l.color[0] * s / d2, l.color[1] * s / d2, l.color[2] * s / d2);
	}
	const float3 full = sub3(lpos, world_pos);
	const float dist = fmaxf(0.1f, sqrtf(dot3f(full, full)));
	const float cone_angle = dot3f(normalize3(sub3(world_pos, lpos)), make_float3(l.direction[0], l.direction[1], l.direction[2]));
	const float scale = h2f(l.spot_scale_bias[0]), bias = h2f(l.spot_scale_bias[1]);
	float cone = fclamp(cone_angle * scale + bias, 0.0f, 1.0f);
	cone *= cone;
	cone *= 1.0f - smoothstep_f(0.9f, 1.0f, dist * l.inv_radius);
	if (!(cone > 0.0f))
		return make_float3(0.0f, 0.0f, 0.0f);
	const float k = (cone * 1.0f) / (dist * dist);
	return make_float3(l.color[0] * k, l.color[1] * k, l.color[2] * k);
}

__global__ void __launch_bounds__(128) fog_light_density_kernel(FogDensityArgs a, uint2 *__restrict__ out)
{
	const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 4 + threadIdx.y, z = blockIdx.z;
	if (x >= a.w || y >= a.h)
		return;
	const float inv_x = 1.0f / (float)a.w, inv_y = 1.0f / (float)a.h, inv_z = 1.0f / (float)a.d;
	float u = ((float)x + 0.5f) * inv_x, v = ((float)y + 0.5f) * inv_y, w = ((float)z + 0.5f) * inv_z;
	const uint32_t dl = __ldg(a.dither_lut + ((size_t)a.dither_offset * 128 + (size_t)(y & 127)) * 128 + (size_t)(x & 127));
	float dx = (float)(dl & 255u) / 255.0f, dy = (float)((dl >> 8) & 255u) / 255.0f, dz = (float)((dl >> 16) & 255u) / 255.0f;
	dx -= 0.5f;
	dy -= 0.5f;
	dz = -dz;
	u = fclamp(u + dx * inv_x, 0.0f, 1.0f);
	v = fclamp(v + dy * inv_y, 0.0f, 1.0f);
	w = fclamp(w + dz * inv_z, 0.001f, 1.0f);
	// get_world_position
	const float world_z = exp2f(w / a.slice_z_log2_scale) - 1.0f;
	const float clip_z = (a.zt[2] - a.zt[0] * world_z) / (a.zt[3] - a.zt[1] * world_z);
	const float4 clip = mul_pairwise(a.ivp, u * 2.0f - 1.0f, v * 2.0f - 1.0f, clip_z, 1.0f);
	const float3 pos = make_float3(clip.x / clip.w, clip.y / clip.w, clip.z / clip.w);
	// fog albedo
	const float lx = (u * 2.0f - 1.0f) * a.xy_scale[0], ly = (v * 2.0f - 1.0f) * a.xy_scale[1];
	const float length_mod = sqrtf(1.0f * 1.0f + lx * lx + ly * ly);
	// the shader indexes the extents with gl_GlobalInvocationID.z, which for its 64 x 1 x 1 workgroup (remapped to 4 x 4 x 4
	// froxels) is the workgroup's z = z / 4: four slices share one extent.  Reproduced as the reference behaves.
	float albedo = a.density_mod * __ldg(a.slice_extents + (z >> 2)) * length_mod;
	albedo = albedo * 0.1f;
	// directional in-scatter
	const float3 cam = make_float3(a.camera_pos[0], a.camera_pos[1], a.camera_pos[2]);
	const float3 to_cam = normalize3(sub3(cam, pos));
	const float phase_d = (0.55f - 0.45f * dot3f(to_cam, make_float3(a.dir_direction[0], a.dir_direction[1], a.dir_direction[2]))) * 1.0f;
	float3 s = make_float3(a.dir_color[0] * phase_d, a.dir_color[1] * phase_d, a.dir_color[2] * phase_d);
	// compute_cluster_scatter_light
	float3 r = make_float3(0.0f, 0.0f, 0.0f);
	const float4 cc = mul_pairwise(a.ctransform, pos.x, pos.y, pos.z, 1.0f);
	if (cc.w > 0.0f)
	{
		const int cx = iclamp(f2i_sat((cc.x * a.cxy_scale[0]) / cc.w), 0, a.res_x - 1), cy = iclamp(f2i_sat((cc.y * a.cxy_scale[1]) / cc.w), 0, a.res_y - 1);
		const int cluster_base = (cy * a.res_x + cx) * a.n32;
		const float zs = dot3f(sub3(pos, make_float3(a.cbase[0], a.cbase[1], a.cbase[2])), make_float3(a.cfront[0], a.cfront[1], a.cfront[2])) * a.z_scale;
		const int z_index = iclamp(f2i_sat(zs), 0, a.z_max_index);
		const uint2 zr = __ldg(a.cluster_range + z_index);
		for (int i = (int)(zr.x >> 5); i <= (int)(zr.y >> 5) && i < a.n32; i++)
		{
			// cluster_mask_range (clusterer_bindless_buffers.h:17-27)
			const uint32_t start = 32u * (uint32_t)i, hi = start + 32u;
			const uint32_t rx = min(max(zr.x, start), hi), ry1 = min(max(zr.y + 1u, rx), hi);
			const uint32_t bits = ry1 - rx;
			uint32_t mask = __ldg(a.bitmask + cluster_base + i) & (bits == 32u ? 0xffffffffu : (((1u << bits) - 1u) << (rx - start)));
			const uint32_t tm = __ldg(a.type_mask + i);
			while (mask)
			{
				const int bit = __ffs(mask) - 1;
				mask &= mask - 1u;
				const GrbPositionalLight l = a.lights[32 * i + bit];
				const float3 color = positional_color(l, ((tm >> bit) & 1u) != 0u, pos);
				const float vol = dot3f(to_cam, normalize3(sub3(make_float3(l.position[0], l.position[1], l.position[2]), pos)));
				const float phase = 0.55f - 0.45f * vol;
				r = make_float3(r.x + color.x * phase, r.y + color.y * phase, r.z + color.z * phase);
			}
		}
	}
	s = make_float3(s.x + r.x, s.y + r.y, s.z + r.z);
	out[((size_t)z * a.h + y) * a.w + x] = pack_rgba16f(make_float4(a.in_scatter_strength * s.x, a.in_scatter_strength * s.y, a.in_scatter_strength * s.z, albedo));
}
} // namespace
} // namespace grb

#ifndef GRB_HOST_EMULATION // tests/cpp/emulate_fog.cpp compiles the kernel above for the CPU and supplies its own loops
using namespace grb;

extern "C" int32_t grb_fog_accumulate(const void *light_density, int32_t width, int32_t height, int32_t depth, void *fog, void *stream)
{
	if (!light_density || !fog || width <= 0 || height <= 0 || depth <= 0 || light_density == fog || (reinterpret_cast<uintptr_t>(light_density) % 8) != 0 ||
	    (reinterpret_cast<uintptr_t>(fog) % 8) != 0)
	{
		set_last_error("grb_fog_accumulate: light_density and fog are distinct, 8-byte aligned R16G16B16A16_SFLOAT volumes of width x height x depth texels");
		return GRB_ERR_INVALID_ARGUMENT;
	}
	Vol16 v;
	v.p = static_cast<const uint2 *>(light_density);
	v.w = width;
	v.h = height;
	v.d = depth;
	const dim3 grid((unsigned)((width + 31) / 32), (unsigned)((height + 7) / 8), 1), block(32, 8);
	fog_accumulate_kernel<<<grid, block, 0, as_stream(stream)>>>(v, static_cast<uint2 *>(fog));
	return check_launch("grb_fog_accumulate");
}
extern "C" int32_t grb_fog_light_density(const GrbFogParameters *fog, const GrbCamera *cam, const float *projection16, const float *inv_projection16,
                                         const GrbClusterParameters *params, const GrbClusterBuffers *buf, const float *directional_color3,
                                         const float *directional_direction3, const float *slice_extents, const void *dither_lut, void *light_density,
                                         void *stream)
{
	if (!fog || !cam || !projection16 || !inv_projection16 || !params || !buf || !directional_color3 || !directional_direction3 || !slice_extents || !dither_lut ||
	    !light_density || fog->width <= 0 || fog->height <= 0 || fog->depth <= 0 || fog->depth > 65535 || fog->dither_offset < 0 ||
	    !(fog->slice_z_log2_scale > 0.0f) || (reinterpret_cast<uintptr_t>(light_density) % 8) != 0)
	{
		set_last_error("grb_fog_light_density: null argument, empty grid (depth <= 65535), negative dither layer, slice_z_log2_scale <= 0 or misaligned output");
		return GRB_ERR_INVALID_ARGUMENT;
	}
	if (!buf->cluster_range || (params->num_lights > 0 && (!buf->lights || !buf->type_mask || !buf->bitmask)))
	{
		set_last_error("grb_fog_light_density: null cluster buffer");
		return GRB_ERR_INVALID_ARGUMENT;
	}
	FogDensityArgs a;
	a.w = fog->width;
	a.h = fog->height;
	a.d = fog->depth;
	a.dither_offset = fog->dither_offset;
	a.slice_z_log2_scale = fog->slice_z_log2_scale;
	a.density_mod = fog->density_mod;
	a.in_scatter_strength = fog->in_scatter_strength;
	for (int i = 0; i < 16; i++)
	{
		a.ivp[i] = cam->inv_view_projection[i];
		a.ctransform[i] = params->transform[i];
	}
	a.zt[0] = projection16[10]; // volumetric_fog.cpp:161-162: vec4(projection[2].zw, projection[3].zw)
	a.zt[1] = projection16[11];
	a.zt[2] = projection16[14];
	a.zt[3] = projection16[15];
	a.xy_scale[0] = inv_projection16[0]; // :167-168
	a.xy_scale[1] = inv_projection16[5];
	for (int i = 0; i < 3; i++)
	{
		a.camera_pos[i] = cam->camera_position[i];
		a.dir_color[i] = directional_color3[i];
		a.dir_direction[i] = directional_direction3[i];
		a.cbase[i] = params->camera_base[i];
		a.cfront[i] = params->camera_front[i];
	}
	a.cxy_scale[0] = params->xy_scale[0];
	a.cxy_scale[1] = params->xy_scale[1];
	a.res_x = params->resolution_xy[0];
	a.res_y = params->resolution_xy[1];
	a.n32 = params->num_lights_32;
	a.z_max_index = params->z_max_index;
	a.z_scale = params->z_scale;
	a.lights = buf->lights;
	a.type_mask = buf->type_mask;
	a.bitmask = buf->bitmask;
	a.cluster_range = reinterpret_cast<const uint2 *>(buf->cluster_range);
	a.slice_extents = slice_extents;
	a.dither_lut = static_cast<const uint32_t *>(dither_lut);
	const dim3 grid((unsigned)((a.w + 31) / 32), (unsigned)((a.h + 3) / 4), (unsigned)a.d), block(32, 4);
	fog_light_density_kernel<<<grid, block, 0, as_stream(stream)>>>(a, static_cast<uint2 *>(light_density));
	return check_launch("grb_fog_light_density");
}
#endif // GRB_HOST_EMULATION
