// grb_api.cu -- library-level entry points of libgranite_b200: ABI version, per-device
// initialisation of constant tables, and the thread-local error string behind the C ABI's
// "int return code + grb_last_error_string()" convention (SURVEY.md §8b error conventions:
// no exceptions and no aborts cross the C boundary).
#include "grb_common.cuh"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>

namespace grb
{
static thread_local char t_last_error[512] = "";

void set_last_error(const char *msg)
{
	std::snprintf(t_last_error, sizeof(t_last_error), "%s", msg ? msg : "");
}

int32_t check_launch(const char *what)
{
	cudaError_t err = cudaGetLastError();
	if (err == cudaSuccess)
		return GRB_OK;
	std::snprintf(t_last_error, sizeof(t_last_error), "%s: %s", what, cudaGetErrorString(err));
	return GRB_ERR_CUDA;
}

int32_t upload_srgb_lut(const float *lut256); // grb_lighting.cu
} // namespace grb

extern "C" int32_t grb_abi_version(void)
{
	return GRB_ABI_VERSION;
}

extern "C" const char *grb_last_error_string(void)
{
	return grb::t_last_error;
}

extern "C" int32_t grb_init(void)
{
	int device = -1;
	cudaError_t err = cudaGetDevice(&device);
	if (err != cudaSuccess)
	{
		grb::set_last_error(cudaGetErrorString(err));
		return GRB_ERR_CUDA;
	}
	cudaDeviceProp prop;
	err = cudaGetDeviceProperties(&prop, device);
	if (err != cudaSuccess)
	{
		grb::set_last_error(cudaGetErrorString(err));
		return GRB_ERR_CUDA;
	}
	if (prop.major != 10)
	{
		char msg[256];
		std::snprintf(msg, sizeof(msg), "grb_init: libgranite_b200 is built for sm_100a only; device %d is sm_%d%d", device, prop.major, prop.minor);
		grb::set_last_error(msg);
		return GRB_ERR_NOT_INITIALIZED;
	}
	// R8G8B8A8_SRGB texel fetch: exact EOTF evaluated in double, rounded once to fp32.
	float lut[256];
	for (int v = 0; v < 256; v++)
	{
		double c = (double)v / 255.0;
		double l = c <= 0.04045 ? c / 12.92 : std::pow((c + 0.055) / 1.055, 2.4);
		lut[v] = (float)l;
	}
	static std::mutex lock;
	std::lock_guard<std::mutex> hold(lock);
	return grb::upload_srgb_lut(lut);
}
