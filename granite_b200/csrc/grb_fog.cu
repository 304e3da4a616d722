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
#endif // GRB_HOST_EMULATION
