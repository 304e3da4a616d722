// emulate_shadow.cpp -- the shadow-map comparison sampling of granite_b200/csrc/grb_shadow.cuh compiled for the CPU
// (cuda_host_emul.h), exported with a C ABI for tests/test_shadow_source_cpu.py.
#include "cuda_host_emul.h"

#include "../../granite_b200/csrc/grb_shadow.cuh"

extern "C" void emu_shadow_2d(const uint16_t *map, int res, const float *clip4, int n, float *out)
{
	for (int i = 0; i < n; i++)
		out[i] = grb::shadow_sample_2d(map, res, clip4[4 * i], clip4[4 * i + 1], clip4[4 * i + 2], clip4[4 * i + 3]);
}

extern "C" void emu_shadow_cube(const uint16_t *map, int res, const float *dir_ref4, int n, float *out)
{
	for (int i = 0; i < n; i++)
		out[i] = grb::shadow_sample_cube(map, res, dir_ref4[4 * i], dir_ref4[4 * i + 1], dir_ref4[4 * i + 2], dir_ref4[4 * i + 3]);
}

extern "C" void emu_spot_shadow_falloff(const float *transform16, const float *pos3, int n, const uint16_t *map, int res, float *out)
{
	for (int i = 0; i < n; i++)
		out[i] = grb::spot_shadow_falloff(transform16, pos3[3 * i], pos3[3 * i + 1], pos3[3 * i + 2], map, res);
}

extern "C" void emu_point_shadow_falloff(const float *transform16, const float *full3, int n, const uint16_t *map, int res, float *out)
{
	for (int i = 0; i < n; i++)
		out[i] = grb::point_shadow_falloff(transform16, full3[3 * i], full3[3 * i + 1], full3[3 * i + 2], map, res);
}

extern "C" void emu_shadow_2d_wide(const uint16_t *map, int res, const float *clip4, int n, float *out)
{
	for (int i = 0; i < n; i++)
		out[i] = grb::shadow_sample_2d_wide(map, res, clip4[4 * i], clip4[4 * i + 1], clip4[4 * i + 2], clip4[4 * i + 3]);
}
