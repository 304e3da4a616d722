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

extern "C" void emu_fog_light_density(const GrbFogParameters *fog, const GrbCamera *cam, const float *projection16, const float *inv_projection16,
                                      const GrbClusterParameters *params, const GrbClusterBuffers *buf, const float *dir_color3, const float *dir_direction3,
                                      const float *slice_extents, const uint32_t *dither_lut, uint16_t *out)
{
	grb::FogDensityArgs a;
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
	a.zt[0] = projection16[10];
	a.zt[1] = projection16[11];
	a.zt[2] = projection16[14];
	a.zt[3] = projection16[15];
	a.xy_scale[0] = inv_projection16[0];
	a.xy_scale[1] = inv_projection16[5];
	for (int i = 0; i < 3; i++)
	{
		a.camera_pos[i] = cam->camera_position[i];
		a.dir_color[i] = dir_color3[i];
		a.dir_direction[i] = dir_direction3[i];
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
	a.dither_lut = dither_lut;
	for (unsigned z = 0; z < (unsigned)a.d; z++)
		for (unsigned by = 0; by < (unsigned)((a.h + 3) / 4); by++)
			for (unsigned bx = 0; bx < (unsigned)((a.w + 31) / 32); bx++)
				for (unsigned ty = 0; ty < 4; ty++)
					for (unsigned tx = 0; tx < 32; tx++)
					{
						emu_blockIdx.x = bx;
						emu_blockIdx.y = by;
						emu_blockIdx.z = z;
						emu_threadIdx.x = tx;
						emu_threadIdx.y = ty;
						grb::fog_light_density_kernel(a, reinterpret_cast<uint2 *>(out));
					}
}
