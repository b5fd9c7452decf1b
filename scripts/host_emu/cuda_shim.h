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
#define __launch_bounds__(...)
#define __maxnreg__(...)
struct dim3s { unsigned x, y, z; };
static thread_local dim3s blockIdx{0, 0, 0}, threadIdx{0, 0, 0}, blockDim{1, 1, 1}, gridDim{1, 1, 1};
template <typename T> static inline T __ldg(const T* p) { return *p; }
static inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
static inline int __float_as_int(float f) { int i; std::memcpy(&i, &f, 4); return i; }
struct float4 { float x, y, z, w; }; struct double2 { double x, y; }; struct float2 { float x, y; };
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline double2 make_double2(double x, double y) { return double2{x, y}; }
static inline unsigned __ballot_sync(unsigned, int p) { return p ? 1u : 0u; }
static inline int __any_sync(unsigned, int p) { return p; }
static inline unsigned __activemask() { return 1u; }
static inline void __syncthreads() {}
static inline int __syncthreads_or(int p) { return p; }
static inline int atomicAdd(int* p, int v) { int o = *p; *p += v; return o; }
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
static inline double r8_pow(double x, double y) { return pow(x, y); }
static inline void r8_pow2(double x0, double y0, double x1, double y1, double& r0, double& r1) { r0 = pow(x0, y0); r1 = pow(x1, y1); }
static inline void r8_pow_sb(double x, double y0, double y1, double& r0, double& r1) { r0 = pow(x, y0); r1 = pow(x, y1); }
static inline double r8_log(double x) { return log(x); }
static inline void r8_log_pair(double, double& hi, double& lo) { hi = lo = 0; }
static inline double r8_pow_logbase(double x, double, double, double y) { return pow(x, y); }
static inline double r8_atan(double x) { return atan(x); }
static inline double r8_sin(double x) { return sin(x); }
static inline double r8_cos(double x) { return cos(x); }
static inline double r8_sqrt(double x) { return sqrt(x); }
}
#endif

#ifdef HELP_EMU_COUNT   // count r8_pow calls per source line (where do the pows go?)
#define HELP_R4_OVERRIDE
#include "../../pyhelp_b200/csrc/help_math.h"
static long long g_pow_count[4096];
namespace helpb200 {
using helpmath::r4_exp; using helpmath::r4_log; using helpmath::r4_log10; using helpmath::r4_sin;
using helpmath::r4_cos; using helpmath::r4_tan; using helpmath::r4_acos; using helpmath::r4_pow;
using helpmath::r4_sqrt; using helpmath::r4_pow8; using helpmath::r4_pow7; using helpmath::r4_pow4;
static inline double r8_log(double x) { return helpmath::det_log(x); }
static inline void r8_log_pair(double x, double& hi, double& lo) { helpmath::det_log_pair(x, hi, lo); }
static inline double r8_pow_logbase(double x, double hi, double lo, double y) { return helpmath::det_pow_logbase(x, hi, lo, y); }
static inline double r8_atan(double x) { return helpmath::g_atan(x); }
static inline double r8_sin(double x) { return helpmath::g_sin(x); }
static inline double r8_cos(double x) { return helpmath::g_cos(x); }
static inline double r8_sqrt(double x) { return helpmath::hm_sqrt(x); }
}
static long long g_cap[8]; static float g_prev_relsw[128];
#define HELP_DBG_CAP { g_cap[0]++; if (RELSW == RELMIN) g_cap[1]++; if ((double)PSIJP1 == c.bub) g_cap[2]++; if ((double)SWLIM == c.wp) g_cap[3]++; if ((double)SWLIM == c.fc) g_cap[4]++; if (RELSW == g_prev_relsw[J]) g_cap[5]++; g_prev_relsw[J] = RELSW; }
#define r8_pow(x, y) (g_pow_count[__LINE__ & 4095]++, helpmath::det_pow(x, y))
static inline void emu_pow2(double x0, double y0, double x1, double y1, double& r0, double& r1) { auto R = helpmath::det_pow2(x0, y0, x1, y1); r0 = R.a; r1 = R.b; }
static inline void emu_pow_sb(double x, double y0, double y1, double& r0, double& r1) { auto R = helpmath::det_pow_samebase(x, y0, y1); r0 = R.a; r1 = R.b; }
#define r8_pow2(x0, y0, x1, y1, r0, r1) (g_pow_count[__LINE__ & 4095] += 2, emu_pow2(x0, y0, x1, y1, r0, r1))
#define r8_pow_sb(x, y0, y1, r0, r1) (g_pow_count[__LINE__ & 4095] += 2, emu_pow_sb(x, y0, y1, r0, r1))
#endif

#ifdef HELP_EMU_MEMO   // per-cell event streams: would a memo of the previous evaluation have hit?
#include <vector>
struct MemoStat { std::vector<unsigned char> pre, swp; float prelsw[128]; double pb[128], pd[128], pe[128], ps[128]; };
static MemoStat g_memo[4096];
#define HELP_DBG_PRE2 { MemoStat& M = g_memo[s]; M.pre.push_back(ra == M.prelsw[J]); M.pre.push_back(rb == M.prelsw[J + 1]); M.prelsw[J] = ra; M.prelsw[J + 1] = rb; }
#define HELP_DBG_SWEEP { MemoStat& M = g_memo[s]; bool hit = (base == M.pb[J] && drin == M.pd[J] && Ej == M.pe[J] && swl == M.ps[J]); M.swp.push_back(hit); M.pb[J] = base; M.pd[J] = drin; M.pe[J] = Ej; M.ps[J] = swl; }
#endif

#ifdef HELP_EMU_STAGE   // per-cell stream of (capillary limit?, single powers, shared-base pairs) per segment iteration
#define HELP_R4_OVERRIDE
#include <vector>
#include "../../pyhelp_b200/csrc/help_math.h"
static int g_np = 0, g_nsb = 0;
static std::vector<unsigned char> g_stage[8192];
namespace helpb200 {
using helpmath::r4_exp; using helpmath::r4_log; using helpmath::r4_log10; using helpmath::r4_sin;
using helpmath::r4_cos; using helpmath::r4_tan; using helpmath::r4_acos; using helpmath::r4_pow;
using helpmath::r4_sqrt; using helpmath::r4_pow8; using helpmath::r4_pow7; using helpmath::r4_pow4;
static inline double r8_log(double x) { return helpmath::det_log(x); }
static inline void r8_log_pair(double x, double& hi, double& lo) { helpmath::det_log_pair(x, hi, lo); }
static inline double r8_pow_logbase(double x, double hi, double lo, double y) { return helpmath::det_pow_logbase(x, hi, lo, y); }
static inline double r8_atan(double x) { return helpmath::g_atan(x); }
static inline double r8_sin(double x) { return helpmath::g_sin(x); }
static inline double r8_cos(double x) { return helpmath::g_cos(x); }
static inline double r8_sqrt(double x) { return helpmath::hm_sqrt(x); }
static inline double r8_pow(double x, double y) { g_np++; return helpmath::det_pow(x, y); }
static inline void r8_pow2(double x0, double y0, double x1, double y1, double& r0, double& r1) { g_np += 2; auto R = helpmath::det_pow2(x0, y0, x1, y1); r0 = R.a; r1 = R.b; }
static inline void r8_pow_sb(double x, double y0, double y1, double& r0, double& r1) { g_nsb++; auto R = helpmath::det_pow_samebase(x, y0, y1); r0 = R.a; r1 = R.b; }
}
// at the hook of iteration i: the powers since the previous hook are [HIGH, LOW, Newton of i-1] + [capillary limit of i]
#define HELP_DBG_SWEEP { auto& V = g_stage[s]; const int cap = (kind == 0) ? 3 : 0; const bool dryj = (kind == 0) && ((base + drin - Ej - wp) <= 0.0); V.push_back((unsigned char)(g_np - cap)); V.push_back((unsigned char)g_nsb); V.push_back((unsigned char)(cap + (dryj ? 1 : 0))); V.push_back((unsigned char)J); g_np = 0; g_nsb = 0; }
#endif
