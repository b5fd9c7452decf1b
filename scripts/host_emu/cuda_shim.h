// DEVELOPER DEBUGGING AID -- NOT PART OF THE PRODUCT, NOT BUILT BY build(), NOT
// IMPORTED BY tests/, bench.py OR pyhelp_b200/.  Lets the CUDA kernels of
// pyhelp_b200/csrc be compiled by g++ and stepped through one "thread" at a
// time under gdb/printf when a GPU-vs-oracle difference has to be localised.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#define __host__
#define __device__
#define __global__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __constant__ static const
#define __launch_bounds__(x)
struct dim3s { unsigned x, y, z; };
static thread_local dim3s blockIdx{0, 0, 0}, threadIdx{0, 0, 0}, blockDim{1, 1, 1}, gridDim{1, 1, 1};
template <typename T> static inline T __ldg(const T* p) { return *p; }
static inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
#include <math.h>

#ifdef HELP_EMU_GLIBC_FLOAT
#define HELP_R4_OVERRIDE
namespace helpb200 {
static inline float r4_exp(float x) { return expf(x); }
static inline float r4_log(float x) { return logf(x); }
static inline float r4_log10(float x) { return log10f(x); }
static inline float r4_sin(float x) { return sinf(x); }
static inline float r4_cos(float x) { return cosf(x); }
static inline float r4_tan(float x) { return tanf(x); }
static inline float r4_acos(float x) { return acosf(x); }
static inline float r4_pow(float x, float y) { return powf(x, y); }
static inline float r4_sqrt(float x) { return powf(x, 0.5f); }
static inline float r4_pow8(float x) { return powf(x, 8.f); }
static inline float r4_pow7(float x) { return powf(x, 7.f); }
static inline float r4_pow4(float x) { return powf(x, 4.f); }
}
#endif
