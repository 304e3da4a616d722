// emulate_decal.cpp -- the decal-binning kernels of granite_b200/csrc/grb_decal.cu compiled for the CPU (cuda_host_emul.h)
// and driven thread by thread, exported with a C ABI for tests/test_decal_cpu.py.
#include "cuda_host_emul.h"

#define GRB_HOST_EMULATION 1
#include "../../granite_b200/csrc/grb_decal.cu"

extern "C" void emu_decal_binning(const float *mvps, int num_decals, int res_x, int res_y, float inv_x, float inv_y, float *boxes, uint32_t *bitmask)
{
	for (unsigned b = 0; b < (unsigned)((num_decals + 127) / 128); b++)
		for (unsigned t = 0; t < 128; t++)
		{
			emu_blockIdx.x = b;
			emu_threadIdx.x = t;
			grb::decal_setup_kernel(mvps, num_decals, reinterpret_cast<float4 *>(boxes));
		}
	const int n32 = (num_decals + 31) / 32, words = res_x * res_y * n32;
	for (unsigned b = 0; b < (unsigned)((words + 255) / 256); b++)
		for (unsigned t = 0; t < 256; t++)
		{
			emu_blockIdx.x = b;
			emu_threadIdx.x = t;
			grb::decal_binning_kernel(reinterpret_cast<const float4 *>(boxes), num_decals, n32, res_x, words, inv_x, inv_y, bitmask);
		}
}
