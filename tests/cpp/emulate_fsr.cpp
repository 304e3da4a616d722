// emulate_fsr.cpp -- the FSR 1 kernels of granite_b200/csrc/grb_fsr.cu compiled for the CPU (cuda_host_emul.h) and
// driven pixel by pixel, exported with a C ABI for tests/test_fsr_kernel_source_cpu.py.
#include "cuda_host_emul.h"

#define GRB_HOST_EMULATION 1
#include "../../granite_b200/csrc/grb_fsr.cu"

namespace
{
template <typename F>
void for_each_thread(int w, int rows, F &&f)
{
	const unsigned gx = (unsigned)((w + 31) / 32), gy = (unsigned)((rows + 7) / 8);
	for (unsigned by = 0; by < gy; by++)
		for (unsigned bx = 0; bx < gx; bx++)
			for (unsigned ty = 0; ty < 8; ty++)
				for (unsigned tx = 0; tx < 32; tx++)
				{
					emu_blockIdx.x = bx;
					emu_blockIdx.y = by;
					emu_threadIdx.x = tx;
					emu_threadIdx.y = ty;
					f();
				}
}

GrbImage image(const void *data, int w, int h, int format)
{
	GrbImage im = {};
	im.data = const_cast<void *>(data);
	im.width = w;
	im.height = h;
	im.row_pitch = w * 4;
	im.format = format;
	return im;
}
} // namespace

extern "C" void emu_fsr_easu(const uint32_t *color, int w_in, int h_in, const float *con16, uint32_t *out, int w_out, int h_out, int target_srgb, int y0, int y1)
{
	const GrbImage c = image(color, w_in, h_in, GRB_FORMAT_R8G8B8A8_UNORM), o = image(out, w_out, h_out, GRB_FORMAT_R8G8B8A8_UNORM);
	grb::EasuConstants con;
	for (int i = 0; i < 16; i++)
		con.c[i] = con16[i];
	for_each_thread(w_out, y1 - y0, [&] {
		if (target_srgb)
			grb::fsr_easu_kernel<true>(grb::img8_of(&c), grb::view_of<uint32_t>(&o), con, y0, y1);
		else
			grb::fsr_easu_kernel<false>(grb::img8_of(&c), grb::view_of<uint32_t>(&o), con, y0, y1);
	});
}

extern "C" void emu_fsr_rcas(const uint32_t *color, int w, int h, float sharpness, uint32_t *out, int srgb, int y0, int y1)
{
	const GrbImage c = image(color, w, h, GRB_FORMAT_R8G8B8A8_UNORM), o = image(out, w, h, GRB_FORMAT_R8G8B8A8_UNORM);
	for_each_thread(w, y1 - y0, [&] {
		if (srgb)
			grb::fsr_rcas_kernel<true>(grb::img8_of(&c), grb::view_of<uint32_t>(&o), sharpness, y0, y1);
		else
			grb::fsr_rcas_kernel<false>(grb::img8_of(&c), grb::view_of<uint32_t>(&o), sharpness, y0, y1);
	});
}

extern "C" float emu_srgb8_to_linear(int v) { return grb::k_srgb8_to_linear[v & 255]; }
