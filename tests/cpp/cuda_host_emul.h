// cuda_host_emul.h -- just enough of the CUDA device environment to compile a kernel source file with g++ and run
// its kernels on the CPU, one "thread" at a time (tests only: this checks the transliteration of a kernel -- indexing,
// control flow, arithmetic order -- on machines without a GPU; it says nothing about launch configuration, memory
// spaces or alignment).  Build with -ffp-contract=off: the *_rn intrinsics below must stay single operations.
#pragma once

#include <cmath>
#include <cstdint>
#include <cstring>

#include <cuda_runtime.h>

// per-thread built-ins: the real ones are `extern const`, so the kernel source sees these names instead
struct EmuIdx
{
	unsigned x = 0, y = 0, z = 0;
};
static thread_local EmuIdx emu_threadIdx, emu_blockIdx;
static thread_local EmuIdx emu_blockDim, emu_gridDim;
#define threadIdx emu_threadIdx
#define blockIdx emu_blockIdx
#define blockDim emu_blockDim
#define gridDim emu_gridDim

#define __launch_bounds__(...)
static inline uint32_t __float_as_uint(float f)
{
	uint32_t u;
	std::memcpy(&u, &f, 4);
	return u;
}
static inline float __uint_as_float(uint32_t u)
{
	float f;
	std::memcpy(&f, &u, 4);
	return f;
}
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
template <typename T>
static inline T __ldg(const T *p) { return *p; }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
