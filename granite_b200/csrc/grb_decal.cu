// grb_decal.cu -- volumetric-decal binning: the consumer next to the light clusterer that
// shares its tile grid.  Reference: LightClusterer::update_bindless_mask_buffer_decal_gpu (renderer/lights/
// clusterer.cpp:1391-1461) dispatching assets/shaders/lights/clusterer_bindless_binning_decal.comp -- per decal the
// screen-space bounding box of its unit cube under mvp = view_projection * world, per (tile, decal) one overlap test,
// one bit in bitmask[(tile_y * res_x + tile_x) * num_decals_32 + decal / 32].
//
// Two launches instead of the shader's one: the bounding box is a property of the decal, not of the (decal, tile)
// pair the shader recomputes it for (8192 tiles at the viewer's 128 x 64 grid), so a first kernel computes the boxes
// (one thread per decal, 16 B each), and the binning kernel is one thread per OUTPUT WORD -- thread t owns
// bitmask[t], i.e. (tile, chunk of 32 decals) with the chunk index fastest, so a warp's stores are one contiguous
// 128-byte line and its 32 box reads per iteration are a broadcast-friendly strided walk through L1.  The pass is
// bound by the bitmask write (res_x res_y num_decals / 8 bytes: 4 MiB at 4096 decals).
//
// Arithmetic as the shader's statements, one IEEE operation each (built with -fmad=false); mat4 * vec4 pairwise, as the
// generated reference code evaluates it.  Bit-exact against the oracle, which is pinned to the shader run on the CPU
// (its SUBGROUPS = 0 path: every pixel tests every decal of its chunk -- the subgroup path's coarse 8 x 4 pre-test only
// prunes work).  Also compiled for the CPU and checked without a GPU (tests/cpp/emulate_decal.cpp).
#include "grb_common.cuh"

namespace grb
{
namespace
{
// compute_decal_screen_bb (.comp:39-70)
GRB_DEV float4 decal_screen_bb(const float *__restrict__ m)
{
	float bx0 = 1.0f, by0 = 1.0f, bx1 = -1.0f, by1 = -1.0f, lo_w = 1.0f, hi_w = -1.0f;
	float c0[4], col0[4], col1[4], col2[4];
#pragma unroll
	for (int r = 0; r < 4; r++)
	{
		col0[r] = __ldg(m + r);
		col1[r] = __ldg(m + 4 + r);
		col2[r] = __ldg(m + 8 + r);
		c0[r] = (col0[r] * -0.5f + col1[r] * -0.5f) + (col2[r] * -0.5f + __ldg(m + 12 + r) * 1.0f);
	}
#pragma unroll
	for (int i = 0; i < 8; i++)
	{
		// corner1 = corner0 + c[0]; corner2 = corner0 + c[1]; corner3 = corner1 + c[1]; corners 4..7 = corners 0..3 + c[2]
		float v[4];
#pragma unroll
		for (int r = 0; r < 4; r++)
		{
			float t = c0[r];
			if (i & 1)
				t = t + col0[r];
			if (i & 2)
				t = t + col1[r];
			if (i & 4)
				t = t + col2[r];
			v[r] = t;
		}
		lo_w = fminf(lo_w, v[3]);
		hi_w = fmaxf(hi_w, v[3]);
		const float px = v[0] / v[3], py = v[1] / v[3];
		bx0 = fminf(bx0, px); // fminf / fmaxf keep the accumulated value when the projection is NaN, as GLSL's (y < x) ? y : x does
		by0 = fminf(by0, py);
		bx1 = fmaxf(bx1, px);
		by1 = fmaxf(by1, py);
	}
	if (hi_w <= 0.0f)
		return make_float4(-10.0f, -10.0f, -10.0f, -10.0f);
	if (lo_w <= 0.0f)
		return make_float4(-1.0f, -1.0f, 1.0f, 1.0f);
	return make_float4(bx0, by0, bx1, by1);
}

__global__ void __launch_bounds__(128) decal_setup_kernel(const float *__restrict__ mvps, int num_decals, float4 *__restrict__ boxes)
{
	const int i = blockIdx.x * 128 + threadIdx.x;
	if (i < num_decals)
		boxes[i] = decal_screen_bb(mvps + 16 * (size_t)i);
}

// main() of the shader (SUBGROUPS = 0, .comp:118-141), one thread per bitmask word
__global__ void __launch_bounds__(256) decal_binning_kernel(const float4 *__restrict__ boxes, int num_decals, int num_decals_32, int res_x, int total_words,
                                                           float inv_x, float inv_y, uint32_t *__restrict__ bitmask)
{
	const int t = blockIdx.x * 256 + threadIdx.x;
	if (t >= total_words)
		return;
	const int chunk = t % num_decals_32, tile = t / num_decals_32;
	const int x = tile % res_x, y = tile / res_x;
	const float u = 2.0f * (float)x * inv_x - 1.0f, v = 2.0f * (float)y * inv_y - 1.0f;
	const float u1 = u + 2.0f * inv_x, v1 = v + 2.0f * inv_y;
	uint32_t mask = 0u;
	const int first = 32 * chunk, count = min(32, num_decals - first);
	for (int b = 0; b < count; b++)
	{
		const float4 bb = __ldg(boxes + first + b);
		if (u1 > bb.x && v1 > bb.y && u < bb.z && v < bb.w) // test_decal (.comp:28-31)
			mask |= 1u << b;
	}
	bitmask[t] = mask;
}
} // namespace
} // namespace grb

#ifndef GRB_HOST_EMULATION // tests/cpp/emulate_decal.cpp compiles the kernels above for the CPU and supplies its own loops
using namespace grb;

extern "C" int32_t grb_cluster_decal_binning(const GrbClusterParameters *params, const float *mvps, int32_t num_decals, float *boxes, uint32_t *bitmask,
                                             void *stream)
{
	if (!params || num_decals < 0 || num_decals > 4096 || params->resolution_xy[0] <= 0 || params->resolution_xy[1] <= 0)
	{
		set_last_error("grb_cluster_decal_binning: bad arguments (at most CLUSTERER_MAX_DECALS_BINDLESS = 4096 decals)");
		return GRB_ERR_INVALID_ARGUMENT;
	}
	if (num_decals == 0)
		return GRB_OK; // clusterer.cpp:1394-1395
	if (!mvps || !boxes || !bitmask)
	{
		set_last_error("grb_cluster_decal_binning: null buffer");
		return GRB_ERR_INVALID_ARGUMENT;
	}
	const int n32 = (num_decals + 31) / 32;
	const long long words = (long long)params->resolution_xy[0] * params->resolution_xy[1] * n32;
	if (words > 0x7fffffffLL)
	{
		set_last_error("grb_cluster_decal_binning: grid too large");
		return GRB_ERR_INVALID_ARGUMENT;
	}
	decal_setup_kernel<<<(num_decals + 127) / 128, 128, 0, as_stream(stream)>>>(mvps, num_decals, reinterpret_cast<float4 *>(boxes));
	decal_binning_kernel<<<(unsigned)((words + 255) / 256), 256, 0, as_stream(stream)>>>(reinterpret_cast<const float4 *>(boxes), num_decals, n32,
	                                                                                    params->resolution_xy[0], (int)words, params->inv_resolution_xy[0],
	                                                                                    params->inv_resolution_xy[1], bitmask);
	return check_launch("grb_cluster_decal_binning");
}
#endif // GRB_HOST_EMULATION
