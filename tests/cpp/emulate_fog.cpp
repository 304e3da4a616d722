// emulate_fog.cpp -- the fog accumulation kernel of granite_b200/csrc/grb_fog.cu compiled for the CPU (cuda_host_emul.h)
// and driven column by column, exported with a C ABI for tests/test_fog_cpu.py.
#include "cuda_host_emul.h"

#define GRB_HOST_EMULATION 1
#include "../../granite_b200/csrc/grb_fog.cu"

extern "C" void emu_fog_accumulate(const uint16_t *light, int w, int h, int d, uint16_t *fog)
{
	grb::Vol16 v;
	v.p = reinterpret_cast<const uint2 *>(light);
	v.w = w;
	v.h = h;
	v.d = d;
	for (unsigned by = 0; by < (unsigned)((h + 7) / 8); by++)
		for (unsigned bx = 0; bx < (unsigned)((w + 31) / 32); bx++)
			for (unsigned ty = 0; ty < 8; ty++)
				for (unsigned tx = 0; tx < 32; tx++)
				{
					emu_blockIdx.x = bx;
					emu_blockIdx.y = by;
					emu_threadIdx.x = tx;
					emu_threadIdx.y = ty;
					grb::fog_accumulate_kernel(v, reinterpret_cast<uint2 *>(fog));
				}
}
