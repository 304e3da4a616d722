// emulate_smaa.cpp -- the SMAA kernels of granite_b200/csrc/grb_smaa.cu compiled for the CPU (cuda_host_emul.h) and
// driven pixel by pixel, exported with a C ABI for tests/test_smaa_kernel_source_cpu.py.
#include "cuda_host_emul.h"

#define GRB_HOST_EMULATION 1
#include "../../granite_b200/csrc/grb_smaa.cu"

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

GrbImage image(const void *data, int w, int h, int format, int texel)
{
	GrbImage im = {};
	im.data = const_cast<void *>(data);
	im.width = w;
	im.height = h;
	im.row_pitch = w * texel;
	im.format = format;
	return im;
}
} // namespace

extern "C" void emu_smaa_edge(const uint32_t *color, int w, int h, int quality, uint8_t *edges, int y0, int y1)
{
	const GrbImage c = image(color, w, h, GRB_FORMAT_R8G8B8A8_UNORM, 4), e = image(edges, w, h, GRB_FORMAT_R8G8_UNORM, 2);
	for_each_thread(w, y1 - y0, [&] { grb::smaa_edge_kernel(grb::tex_of<4>(&c), grb::view_of<uchar2>(&e), grb::preset_of(quality), y0, y1); });
}

extern "C" void emu_smaa_weights(const uint8_t *edges, int w, int h, const uint8_t *area, const uint8_t *search, int quality, uint32_t *weights, int y0, int y1)
{
	const GrbImage e = image(edges, w, h, GRB_FORMAT_R8G8_UNORM, 2), a = image(area, 160, 560, GRB_FORMAT_R8G8_UNORM, 2),
	               s = image(search, 64, 16, GRB_FORMAT_R8_UNORM, 1), o = image(weights, w, h, GRB_FORMAT_R8G8B8A8_UNORM, 4);
	for_each_thread(w, y1 - y0, [&] {
		grb::smaa_weights_kernel(grb::tex_of<2>(&e), grb::tex_of<2>(&a), grb::tex_of<1>(&s), grb::view_of<uint32_t>(&o), grb::preset_of(quality), y0, y1);
	});
}

extern "C" void emu_smaa_blend(const uint32_t *color, const uint32_t *weights, int w, int h, int srgb, uint32_t *out, int y0, int y1)
{
	const GrbImage c = image(color, w, h, GRB_FORMAT_R8G8B8A8_UNORM, 4), b = image(weights, w, h, GRB_FORMAT_R8G8B8A8_UNORM, 4),
	               o = image(out, w, h, srgb ? GRB_FORMAT_R8G8B8A8_SRGB : GRB_FORMAT_R8G8B8A8_UNORM, 4);
	for_each_thread(w, y1 - y0, [&] {
		if (srgb)
			grb::smaa_blend_kernel<true>(grb::tex_of<4>(&c), grb::tex_of<4>(&b), grb::view_of<uint32_t>(&o), y0, y1);
		else
			grb::smaa_blend_kernel<false>(grb::tex_of<4>(&c), grb::tex_of<4>(&b), grb::view_of<uint32_t>(&o), y0, y1);
	});
}
