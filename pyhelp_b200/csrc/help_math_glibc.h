// help_math_glibc.h -- the transcendental functions of the HELP engine, BIT-IDENTICAL to the libm the
// reference links in this image (glibc 2.39, x86-64, the ifunc variants a CPU with FMA + AVX2 selects).
//
// Why.  HELP's water balance is chaotic in the last bit of its transcendentals:
// one differently rounded pow()/expf() moves yearly totals of a cell by per cents.  The reference's
// Fortran gets `**`, EXP, ALOG, SIN ... from the libm its compiler links, so "the reference's result" is
// defined only together with that libm.  These functions therefore do not approximate the mathematical
// functions in some way of their own: each one executes the operation sequence of the corresponding glibc
// routine -- same tables, same polynomial forms, and the same fused multiply-adds the library's FMA build
// contains (read off the disassembly of __pow_fma, __expf_fma, __logf_fma, __powf_fma, __sinf_fma,
// __cosf_fma; tanf, acosf and log10f are glibc's baseline SSE2 code, which has no contraction).  The same
// source compiled by g++ (-ffp-contract=off) for the CPU oracle and by nvcc (-fmad=false) for sm_100a
// returns the same bits as libm.so.6: tests/test_help_math.py checks 10^7 arguments per function against
// the library, tests/test_gpu_math.py the device against the host.  Consequence: the GPU engine and the
// oracle build that calls the system libm directly agree bit for bit (tests/test_parity_envelope.py).
//
// Provenance of the algorithms: pow/expf/logf/powf/sinf/cosf are Arm Optimized Routines (Szabolcs Nagy,
// Wilco Dijkstra; MIT OR Apache-2.0 WITH LLVM-exception), shipped by glibc >= 2.27/2.28 as
// sysdeps/ieee754/dbl-64/e_pow.c and sysdeps/ieee754/flt-32/{e_expf,e_logf,e_powf,s_sinf,s_cosf}.c;
// tanf/acosf/log10f are the FreeBSD msun float routines (k_tanf.c, e_acosf.c, e_log10f.c; Sun
// Microsystems permissive notice) as glibc 2.39 still ships them.  This is third-party arithmetic, not
// the reference's; the table data is extracted from the library (scripts/extract_glibc_tables.py).
//
// Domain notes.  Every IEEE special case of the library routines is reproduced (zeros, infinities, NaNs,
// negative bases, subnormals); errno and floating-point exception flags are not.
#pragma once
#include <stdint.h>

#include "help_math_glibc_tables.h"

namespace helpmath {

// ---- tables ---------------------------------------------------------------------------------------
struct GExpEntBits { uint64_t tail, sbits; };
struct GF2 { double invc, logc; };
static const LogEnt kGPowLogHost[128] = {HG_POW_LOG_TABLE};
static const GExpEntBits kGExpHost[128] = {HG_EXP_TABLE};
static const uint64_t kGExp2fHost[32] = HG_EXP2F_TABLE;
static const GF2 kGLogfHost[16] = HG_LOGF_TABLE;
static const GF2 kGPowfLog2Host[16] = HG_POWF_LOG2_TABLE;
static const double kGSinCosfHost[2][14] = {HG_SINCOSF_0, HG_SINCOSF_1};
static const uint32_t kGInvPio4Host[24] = HG_INV_PIO4;
static const GF2 kGLogHost[128] = {HG_LOG_TABLE};
static const double kGSinCosHost[440] = HG_SINCOS_TABLE;
static const double kGAtanCijHost[241 * 7] = HG_ATAN_CIJ;
#if defined(__CUDACC__)
static __device__ const LogEnt kGPowLogDev[128] = {HG_POW_LOG_TABLE};
static __device__ const GExpEntBits kGExpDev[128] = {HG_EXP_TABLE};
static __device__ const uint64_t kGExp2fDev[32] = HG_EXP2F_TABLE;
static __device__ const GF2 kGLogfDev[16] = HG_LOGF_TABLE;
static __device__ const GF2 kGPowfLog2Dev[16] = HG_POWF_LOG2_TABLE;
static __device__ const double kGSinCosfDev[2][14] = {HG_SINCOSF_0, HG_SINCOSF_1};
static __device__ const uint32_t kGInvPio4Dev[24] = HG_INV_PIO4;
static __device__ const GF2 kGLogDev[128] = {HG_LOG_TABLE};
static __device__ const double kGSinCosDev[440] = HG_SINCOS_TABLE;
static __device__ const double kGAtanCijDev[241 * 7] = HG_ATAN_CIJ;
#endif
#if defined(__CUDACC__) && !defined(HM_GLOBAL_TABLES)
// The pow tables (5 KB) live in shared memory on the device, one copy per block, filled by
// hm_load_tables() at the top of every kernel that evaluates a power: a table entry is one LDS at an
// immediate base instead of a global load with its descriptor set-up (profiles/r1_notes.md).
__shared__ LogEnt s_hmLogTab[128];
__shared__ GExpEntBits s_hmExpTab[128];
__shared__ GF2 s_hmLog2Tab[128];                           // log()'s own table (2 KB): geomembrane leakage takes logarithms every sub-step
static __device__ __forceinline__ void hm_load_tables() {
  for (int i = threadIdx.x; i < 128; i += blockDim.x) { s_hmLogTab[i] = kGPowLogDev[i]; s_hmExpTab[i] = kGExpDev[i]; s_hmLog2Tab[i] = kGLogDev[i]; }
  __syncthreads();
}
#define HG_SHARED 1
#elif defined(__CUDACC__)
static __device__ __forceinline__ void hm_load_tables() {}
#else
static inline void hm_load_tables() {}
#endif
#if defined(__CUDA_ARCH__) && defined(HG_SHARED)
#define HG_POWLOG(i) s_hmLogTab[i]
#define HG_EXPT(i) s_hmExpTab[i]
#define HG_LOGT(i) s_hmLog2Tab[i]
#elif defined(__CUDA_ARCH__)
#define HG_POWLOG(i) kGPowLogDev[i]
#define HG_EXPT(i) kGExpDev[i]
#define HG_LOGT(i) kGLogDev[i]
#else
#define HG_POWLOG(i) kGPowLogHost[i]
#define HG_EXPT(i) kGExpHost[i]
#define HG_LOGT(i) kGLogHost[i]
#endif
#if defined(__CUDA_ARCH__)
#define HG_EXP2F(i) kGExp2fDev[i]
#define HG_LOGF(i) kGLogfDev[i]
#define HG_POWFLOG2(i) kGPowfLog2Dev[i]
#define HG_SINCOSF(k, j) kGSinCosfDev[k][j]
#define HG_INVPIO4(i) kGInvPio4Dev[i]
#define HG_SCT(i) kGSinCosDev[i]
#define HG_CIJ(i, j) kGAtanCijDev[(i) * 7 + (j)]
#else
#define HG_EXP2F(i) kGExp2fHost[i]
#define HG_LOGF(i) kGLogfHost[i]
#define HG_POWFLOG2(i) kGPowfLog2Host[i]
#define HG_SINCOSF(k, j) kGSinCosfHost[k][j]
#define HG_INVPIO4(i) kGInvPio4Host[i]
#define HG_SCT(i) kGSinCosHost[i]
#define HG_CIJ(i, j) kGAtanCijHost[(i) * 7 + (j)]
#endif

HM_FN uint32_t as_u32(float f) {
#if defined(__CUDA_ARCH__)
  return (uint32_t)__float_as_int(f);
#else
  uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
HM_FN float as_f32(uint32_t u) {
#if defined(__CUDA_ARCH__)
  return __int_as_float((int)u);
#else
  float f; memcpy(&f, &u, 4); return f;
#endif
}
HM_FN float hm_sqrtf(float a) {
#if defined(__CUDA_ARCH__)
  return __fsqrt_rn(a);
#else
  return __builtin_sqrtf(a);
#endif
}
HM_FN float hm_nanf() { return as_f32(0x7fc00000u); }
HM_FN float hm_inff() { return as_f32(0x7f800000u); }

// ==== pow (sysdeps/ieee754/dbl-64/e_pow.c, FMA build) ==================================================
// log(x) for the bit pattern (ixh:ixl) of a positive normal x (or a normalised subnormal whose exponent
// field went negative): returns hi, sets tail; hi + tail = log x to ~2^-68 relative (log_inline).
HM_FN double g_log_inline(uint32_t ixh, uint32_t ixl, double& tail) {
  const double A[7] = HG_POW_A;
  const uint32_t tmp = ixh - 0x3fe69555u;                 // OFF = 0x3fe6955500000000: its low word is zero
  const int i = (int)((tmp >> 13) & 127u);
  const int k = (int32_t)tmp >> 20;
  const double z = from_words(ixh - (tmp & 0xfff00000u), ixl);
  // (double)k without the integer->double conversion unit: 2^52 + 2^31 + k is exact, so is the difference
  const double kd = hm_add(from_words(0x43300000u, (uint32_t)k ^ 0x80000000u), -4503601774854144.0);
  const LogEnt e = HG_POWLOG(i);
  const double r = hm_fma(z, e.invc, -1.0);
  const double t1 = hm_fma(kd, HG_POW_LN2HI, e.logc);
  const double t2 = hm_add(t1, r);
  const double lo1 = hm_fma(kd, HG_POW_LN2LO, e.logctail);
  const double lo2 = hm_add(hm_add(t1, -t2), r);
  const double ar = hm_mul(A[0], r);
  const double ar2 = hm_mul(r, ar);
  const double ar3 = hm_mul(r, ar2);
  const double hi = hm_add(t2, ar2);
  const double lo3 = hm_fma(ar, r, -ar2);
  const double lo4 = hm_add(hm_add(t2, -hi), ar2);
  const double q12 = hm_fma(r, A[2], A[1]), q34 = hm_fma(r, A[4], A[3]), q56 = hm_fma(r, A[6], A[5]);
  const double p = hm_fma(ar2, hm_fma(q56, ar2, q34), q12);
  const double lo = hm_fma(ar3, p, hm_add(hm_add(hm_add(lo1, lo2), lo3), lo4));
  const double y = hm_add(hi, lo);
  tail = hm_add(hm_add(hi, -y), lo);
  return y;
}

// the result is subnormal, or scale overflowed the exponent field (specialcase of e_pow.c)
HM_NOINLINE double g_exp_special(double tmp, uint64_t sbits, uint32_t ki_lo) {
  if ((ki_lo & 0x80000000u) == 0) {                       // k > 0
    sbits -= 1009ull << 52;
    const double scale = as_f64(sbits);
    return hm_mul(0x1p1009, hm_fma(scale, tmp, scale));
  }
  sbits += 1022ull << 52;
  const double scale = as_f64(sbits);
  const double st = hm_mul(scale, tmp);
  double y = hm_add(scale, st);
  if ((y < 0.0 ? -y : y) < 1.0) {
    const double one = (y < 0.0) ? -1.0 : 1.0;
    double lo = hm_add(hm_add(scale, -y), st);
    const double hi = hm_add(one, y);
    lo = hm_add(hm_add(hm_add(one, -hi), y), lo);
    y = hm_add(hm_add(hi, lo), -one);
    if (y == 0.0) y = as_f64(sbits & 0x8000000000000000ull);
  }
  return hm_mul(0x1p-1022, y);
}

// exp(x + xtail) * (sign_bias ? -1 : 1) for 2^-54 <= |x| < 512: the straight-line part of exp_inline
HM_FN double g_exp_core(double x, double xtail, uint32_t sign_bias, bool special) {
  const double C[4] = HG_EXP_C;                            // C2..C5
  const double kds = hm_fma(x, HG_EXP_INVLN2N, HG_EXP_SHIFT);
  const uint32_t ki = lo32(kds);
  const double kd = hm_add(kds, -HG_EXP_SHIFT);
  double r = hm_fma(kd, HG_EXP_NEGLN2HIN, x);
  r = hm_fma(kd, HG_EXP_NEGLN2LON, r);
  r = hm_add(r, xtail);
  const GExpEntBits e = HG_EXPT(ki & 127u);
  const uint32_t sh = (uint32_t)(e.sbits >> 32) + ((ki + sign_bias) << 13), sl = (uint32_t)e.sbits;   // sbits + ((ki + sign_bias) << 45)
  const double tail = as_f64(e.tail);
  const double r2 = hm_mul(r, r);
  const double a = hm_fma(hm_fma(r, C[1], C[0]), r2, hm_add(tail, r));
  const double tmp = hm_fma(hm_fma(r, C[3], C[2]), hm_mul(r2, r2), a);
  if (special) return g_exp_special(tmp, ((uint64_t)sh << 32) | sl, ki);
  const double scale = from_words(sh, sl);
  return hm_fma(scale, tmp, scale);
}

// everything of exp_inline outside 2^-54 <= |x| < 512
HM_NOINLINE double g_exp_slow(double x, double xtail, uint32_t sign_bias) {
  const uint32_t abstop = (hi32(x) >> 20) & 0x7ffu;
  if ((int32_t)(abstop - 0x3c9u) < 0) { const double one = hm_add(1.0, x); return sign_bias ? -one : one; }
  if (abstop >= 0x409u) {                                  // |x| >= 1024: under/overflow
    if (hi32(x) >> 31) return sign_bias ? -0.0 : 0.0;
    return sign_bias ? -hm_inf() : hm_inf();
  }
  return g_exp_core(x, xtail, sign_bias, true);            // 512 <= |x| < 1024
}

HM_FN double g_exp_inline(double x, double xtail, uint32_t sign_bias) {
  const uint32_t abstop = (hi32(x) >> 20) & 0x7ffu;
  if (abstop - 0x3c9u < 0x3fu) return g_exp_core(x, xtail, sign_bias, false);
  if (abstop < 0x3c9u && sign_bias == 0) return hm_add(1.0, x);          // (0 is a common input: x == 1 or y log x tiny)
  return g_exp_slow(x, xtail, sign_bias);
}

HM_FN int g_checkint(uint64_t iy) {                        // 0: not an integer, 1: odd, 2: even
  const int e = (int)(iy >> 52) & 0x7ff;
  if (e < 0x3ff) return 0;
  if (e > 0x3ff + 52) return 2;
  if (iy & ((1ull << (0x3ff + 52 - e)) - 1)) return 0;
  if (iy & (1ull << (0x3ff + 52 - e))) return 1;
  return 2;
}

// pow outside the straight-line case: x not a positive normal, or |y| outside [2^-65, 2^63)
HM_NOINLINE double g_pow_special(double x, double y) {
  // NaN operands first, with floating-point tests: 1 for y == 0 or x == 1, else the library's NaN (x + y; x * x with the
  // sign of x**y for a finite y) -- what the integer tests below return for them.  A cell with inconsistent soil constants
  // carries NaN through every power of every sub-step, and a launch waits for its slowest warp (profiles/r2_notes.md).
  if (x != x || y != y) {
    if (y == 0.0) return 1.0;
    if (x == 1.0) return 1.0;
    if (y != y || (hi32(y) & 0x7fffffffu) == 0x7ff00000u) return hm_add(x, y);
    double x2 = hm_mul(x, x);
    if ((hi32(x) >> 31) && g_checkint(as_u64(y)) == 1) x2 = -x2;
    return x2;
  }
  uint64_t ix = as_u64(x);
  const uint64_t iy = as_u64(y);
  uint32_t topx = (uint32_t)(ix >> 52);
  const uint32_t topy = (uint32_t)(iy >> 52);
  uint32_t sign_bias = 0;
  const double inf = hm_inf();
  if (2 * iy - 1 >= 2 * as_u64(inf) - 1) {                 // y is 0, inf or nan
    if (2 * iy == 0) return 1.0;
    if (ix == as_u64(1.0)) return 1.0;
    if (2 * ix > 2 * as_u64(inf) || 2 * iy > 2 * as_u64(inf)) return hm_add(x, y);
    if (2 * ix == 2 * as_u64(1.0)) return 1.0;
    if ((2 * ix < 2 * as_u64(1.0)) == !(iy >> 63)) return 0.0;
    return hm_mul(y, y);
  }
  if (2 * ix - 1 >= 2 * as_u64(inf) - 1) {                 // x is 0, inf or nan
    double x2 = hm_mul(x, x);
    if ((ix >> 63) && g_checkint(iy) == 1) x2 = -x2;
    return (iy >> 63) ? 1.0 / x2 : x2;
  }
  if (ix >> 63) {                                           // finite x < 0
    const int yint = g_checkint(iy);
    if (yint == 0) return hm_nan();
    if (yint == 1) sign_bias = 0x40000u;                   // SIGN_BIAS = 0x800 << EXP_TABLE_BITS
    ix &= 0x7fffffffffffffffull;
    topx &= 0x7ffu;
  }
  if ((topy & 0x7ffu) - 0x3beu >= 0x43eu - 0x3beu) {
    if (ix == as_u64(1.0)) return 1.0;
    if ((topy & 0x7ffu) < 0x3beu) return ix > as_u64(1.0) ? hm_add(1.0, y) : hm_add(1.0, -y);     // |y| < 2^-65
    return ((ix > as_u64(1.0)) == (topy < 0x800u)) ? inf : 0.0;
  }
  if (topx == 0) {                                          // subnormal x: normalise, the exponent field goes negative
    ix = as_u64(hm_mul(as_f64(ix), 0x1p52));
    ix &= 0x7fffffffffffffffull;
    ix -= 52ull << 52;
  }
  double lo;
  const double hi = g_log_inline((uint32_t)(ix >> 32), (uint32_t)ix, lo);
  const double ehi = hm_mul(y, hi);
  const double elo = hm_fma(y, lo, hm_fma(hi, y, -ehi));
  return g_exp_inline(ehi, elo, sign_bias);
}

// x positive and normal, 2^-65 <= |y| < 2^63: the case pow() runs straight through
HM_FN bool pow_hot(double x, double y) {
  return ((hi32(x) >> 20) - 1u) < 0x7feu && (((hi32(y) >> 20) & 0x7ffu) - 0x3beu) < 0x80u;
}

// x**y, the bits of glibc's pow().  Not inlined on the device: one copy of the code keeps the
// simulation kernel inside the instruction cache.
HM_NOINLINE double det_pow(double x, double y) {
  if (!pow_hot(x, y)) return g_pow_special(x, y);
  double lo;
  const double hi = g_log_inline(hi32(x), lo32(x), lo);
  const double ehi = hm_mul(y, hi);
  const double elo = hm_fma(y, lo, hm_fma(hi, y, -ehi));
  return g_exp_inline(ehi, elo, 0);
}

// x**y with the logarithm pair (hi, lo) = log_inline(x) supplied by the caller, for a base that is a
// constant of the caller (x positive and normal): the same bits as det_pow(x, y).
HM_FN void det_log_pair(double x, double& hi, double& lo) { hi = g_log_inline(hi32(x), lo32(x), lo); }
HM_NOINLINE double det_pow_logbase(double x, double hi, double lo, double y) {
  if (!pow_hot(x, y)) return g_pow_special(x, y);
  const double ehi = hm_mul(y, hi);
  const double elo = hm_fma(y, lo, hm_fma(hi, y, -ehi));
  return g_exp_inline(ehi, elo, 0);
}

// Two independent powers in one call, each the SAME bits det_pow returns: the two dependency chains are
// written side by side in one basic block so that the FP64 pipe has a second, independent instruction
// to issue.
struct Pair { double a, b; };
HM_NOINLINE Pair det_pow2(double x0, double y0, double x1, double y1) {
  Pair R;
  if (!(pow_hot(x0, y0) && pow_hot(x1, y1))) { R.a = det_pow(x0, y0); R.b = det_pow(x1, y1); return R; }
  double l0, l1;
  const double h0 = g_log_inline(hi32(x0), lo32(x0), l0), h1 = g_log_inline(hi32(x1), lo32(x1), l1);
  const double e0 = hm_mul(y0, h0), e1 = hm_mul(y1, h1);
  const double t0 = hm_fma(y0, l0, hm_fma(h0, y0, -e0)), t1 = hm_fma(y1, l1, hm_fma(h1, y1, -e1));
  R.a = g_exp_inline(e0, t0, 0);
  R.b = g_exp_inline(e1, t1, 0);
  return R;
}

// Two powers of ONE base: r0 = x**y0, r1 = x**y1, each the SAME bits det_pow returns.  The logarithm
// (more than half of a power) is evaluated once.  (The Newton iteration of the drainage solve needs
// x**C and x**(C-1), HELP3O.FOR:4470-4473.)
HM_NOINLINE Pair det_pow_samebase(double x, double y0, double y1) {
  Pair R;
  if (!(pow_hot(x, y0) && pow_hot(x, y1))) { R.a = det_pow(x, y0); R.b = det_pow(x, y1); return R; }
  double l;
  const double h = g_log_inline(hi32(x), lo32(x), l);
  const double e0 = hm_mul(y0, h), e1 = hm_mul(y1, h);
  const double t0 = hm_fma(y0, l, hm_fma(h, y0, -e0)), t1 = hm_fma(y1, l, hm_fma(h, y1, -e1));
  R.a = g_exp_inline(e0, t0, 0);
  R.b = g_exp_inline(e1, t1, 0);
  return R;
}

// ==== log (sysdeps/ieee754/dbl-64/e_log.c, FMA build) ==================================================
HM_NOINLINE double det_log(double x) {
  const double A[5] = HG_LOG_A;
  const double B[11] = HG_LOG_B;
  uint32_t ixh = hi32(x), ixl = lo32(x);                  // (32-bit words: the GPU's registers)
  if (ixh - 0x3fee0000u < 0x30900u) {                      // 1 - 2^-4 <= x < 1 + 0x1.09p-4
    if (ixh == 0x3ff00000u && ixl == 0) return 0.0;
    const double r = hm_add(x, -1.0);
    const double r2 = hm_mul(r, r);
    const double r3 = hm_mul(r, r2);
    const double q1 = hm_fma(r2, B[3], hm_fma(B[2], r, B[1]));
    const double q2 = hm_fma(r2, B[6], hm_fma(B[5], r, B[4]));
    double q3 = hm_fma(r2, B[9], hm_fma(B[8], r, B[7]));
    q3 = hm_fma(r3, B[10], q3);
    const double q = hm_fma(hm_fma(q3, r3, q2), r3, q1);
    const double rw = hm_fma(r, 0x1p27, r);
    const double rhi = hm_fma(-0x1p27, r, rw);
    const double rhi2 = hm_mul(rhi, rhi);
    const double rlo = hm_add(r, -rhi);
    const double hi = hm_fma(rhi2, B[0], r);
    double lo = hm_fma(rhi2, B[0], hm_add(r, -hi));
    lo = hm_fma(hm_mul(B[0], rlo), hm_add(r, rhi), lo);
    const double y = hm_fma(q, r3, lo);
    return hm_add(hi, y);
  }
  const uint32_t top = ixh >> 16;
  if (top - 0x0010u >= 0x7ff0u - 0x0010u) {                 // x < 2^-1022, inf or nan
    if (((ixh << 1) | ixl) == 0) return -hm_inf();
    if (ixh == 0x7ff00000u && ixl == 0) return x;
    if ((top & 0x8000u) || (top & 0x7ff0u) == 0x7ff0u) return hm_nan();
    const double xs = hm_mul(x, 0x1p52);                    // subnormal: normalise, the exponent field goes negative
    ixh = hi32(xs) - (52u << 20); ixl = lo32(xs);
  }
  const uint32_t tmp = ixh - 0x3fe60000u;                   // OFF = 0x3fe6000000000000
  const int i = (int)((tmp >> 13) & 127u);
  const int k = (int32_t)tmp >> 20;
  const double z = from_words(ixh - (tmp & 0xfff00000u), ixl);
  const GF2 e = HG_LOGT(i);
  const double r = hm_fma(z, e.invc, -1.0);
  const double kd = hm_add(from_words(0x43300000u, (uint32_t)k ^ 0x80000000u), -4503601774854144.0);   // (double)k
  const double w = hm_fma(kd, HG_LOG_LN2HI, e.logc);
  const double hi = hm_add(r, w);
  const double lo = hm_fma(kd, HG_LOG_LN2LO, hm_add(hm_add(w, -hi), r));
  const double r2 = hm_mul(r, r);
  const double r3 = hm_mul(r, r2);
  const double p = hm_fma(hm_fma(r, A[4], A[3]), r2, hm_fma(A[2], r, A[1]));
  const double y = hm_fma(r3, p, hm_fma(r2, A[0], lo));
  return hm_add(y, hi);
}

// ==== sin / cos / atan (IBM Accurate Mathematical Library: sysdeps/ieee754/dbl-64/s_sin.c, s_atan.c; FMA
// build) on the domain the engine uses them on: the angle of a drain layer's slope (LATFLO,
// HELP3O.FOR:4686-4712), alpha = atan(slope) in [0, pi/2) and its sine and cosine.  |x| < 2.426265 for
// sin/cos and |x| < 16 for atan run the library's operation sequence; beyond that the older
// implementations of help_math.h answer (not libm-identical; the engine never gets there).
HM_FN double g_copysign(double mag, double sgn) { return from_words((hi32(mag) & 0x7fffffffu) | (hi32(sgn) & 0x80000000u), lo32(mag)); }
HM_FN double g_abs(double x) { return from_words(hi32(x) & 0x7fffffffu, lo32(x)); }
HM_FN double g_do_sin(double x, double dx) {
  const double ax = g_abs(x);
  if (ax < 0.126) {                                         // TAYLOR_SIN
    const double S[5] = HG_SIN_S;
    const double xx = hm_mul(x, x);
    const double p = hm_fma(hm_fma(hm_fma(hm_fma(S[4], xx, S[3]), xx, S[2]), xx, S[1]), xx, S[0]);
    const double t = hm_fma(xx, hm_fma(p, x, -hm_mul(0.5, dx)), dx);
    return hm_add(x, t);
  }
  if (x <= 0.0) dx = -dx;
  const double u = hm_add(0x1.8p45, ax);
  const int k = (int)(lo32(u) << 2);
  const double xr = hm_add(ax, -hm_add(u, -0x1.8p45));
  const double xx = hm_mul(xr, xr);
  const double s = hm_add(xr, hm_fma(hm_mul(xr, xx), hm_fma(HG_SIN_SN5, xx, HG_SIN_SN3), dx));
  const double c = hm_fma(xr, dx, hm_mul(xx, hm_fma(hm_fma(HG_SIN_CS6, xx, HG_SIN_CS4), xx, HG_SIN_CS2)));
  const double sn = HG_SCT(k), ssn = HG_SCT(k + 1), cs = HG_SCT(k + 2), ccs = HG_SCT(k + 3);
  const double cor = hm_fma(s, cs, hm_fma(-c, sn, hm_fma(ccs, s, ssn)));
  return g_copysign(hm_add(sn, cor), x);
}
HM_FN double g_do_cos(double x, double dx) {
  if (x < 0.0) dx = -dx;
  const double ax = g_abs(x);
  const double u = hm_add(0x1.8p45, ax);
  const int k = (int)(lo32(u) << 2);
  const double xr = hm_add(hm_add(ax, -hm_add(u, -0x1.8p45)), dx);
  const double xx = hm_mul(xr, xr);
  const double s = hm_fma(hm_mul(xr, xx), hm_fma(HG_SIN_SN5, xx, HG_SIN_SN3), xr);
  const double c = hm_mul(xx, hm_fma(hm_fma(HG_SIN_CS6, xx, HG_SIN_CS4), xx, HG_SIN_CS2));
  const double sn = HG_SCT(k), ssn = HG_SCT(k + 1), cs = HG_SCT(k + 2), ccs = HG_SCT(k + 3);
  const double cor = hm_fma(-s, sn, hm_fma(-c, cs, hm_fma(-s, ssn, ccs)));
  return hm_add(cs, cor);
}
HM_NOINLINE double g_sin(double x) {
  const uint32_t k = hi32(x) & 0x7fffffffu;
  if (k < 0x3e500000u) return x;                            // |x| < 2^-26
  if (k < 0x3feb6000u) return g_do_sin(x, 0.0);             // |x| < 0.855469
  if (k < 0x400368fdu) return g_copysign(g_do_cos(hm_add(HG_HP0, -g_abs(x)), HG_HP1), x);
  return det_sin(x);
}
HM_NOINLINE double g_cos(double x) {
  const uint32_t k = hi32(x) & 0x7fffffffu;
  if (k < 0x3e400000u) return 1.0;                          // |x| < 2^-27
  if (k < 0x3feb6000u) return g_do_cos(x, 0.0);
  if (k < 0x400368fdu) {
    const double y = hm_add(HG_HP0, -g_abs(x));
    const double a = hm_add(y, HG_HP1);
    const double da = hm_add(hm_add(y, -a), HG_HP1);
    return g_do_sin(a, da);
  }
  return det_cos(x);
}
HM_NOINLINE double g_atan(double x) {
  const double D[6] = HG_ATAN_D;
  if (x != x) return hm_add(x, x);
  const double u = g_abs(x);
  if (u < 0.0625) {
    if (u < HG_ATAN_A) return x;
    const double v = hm_mul(x, x);
    const double t1 = hm_fma(hm_fma(hm_fma(hm_fma(hm_fma(D[5], v, D[4]), v, D[3]), v, D[2]), v, D[1]), v, D[0]);
    return hm_fma(hm_mul(x, v), t1, x);
  }
  if (u < 1.0) {
    const int i = (int)hm_add(hm_fma(u, 256.0, 0x1p52), -0x1p52) - 16;
    const double z = hm_add(u, -HG_CIJ(i, 0));
    const double yy = hm_fma(hm_fma(hm_fma(hm_fma(HG_CIJ(i, 6), z, HG_CIJ(i, 5)), z, HG_CIJ(i, 4)), z, HG_CIJ(i, 3)), z, HG_CIJ(i, 2));
    return g_copysign(hm_fma(yy, z, HG_CIJ(i, 1)), x);
  }
  if (u < 16.0) {
    const double w = 1.0 / u;
    const double t1 = hm_mul(w, u);
    const double t2 = hm_fma(u, w, -t1);
    const double e = hm_add(hm_add(1.0, -t1), -t2);
    const int i = (int)hm_add(hm_fma(w, 256.0, 0x1p52), -0x1p52) - 16;
    const double z = hm_fma(e, w, hm_add(w, -HG_CIJ(i, 0)));
    const double yy = hm_fma(hm_fma(hm_fma(hm_fma(HG_CIJ(i, 6), z, HG_CIJ(i, 5)), z, HG_CIJ(i, 4)), z, HG_CIJ(i, 3)), z, HG_CIJ(i, 2));
    const double b2 = hm_fma(-yy, z, HG_HP1);
    const double b1 = hm_add(HG_HP0, -HG_CIJ(i, 1));
    return g_copysign(hm_add(b1, b2), x);
  }
  return det_atan(x);
}

// ==== expf (sysdeps/ieee754/flt-32/e_expf.c, FMA build) ================================================
HM_NOINLINE float r4_exp(float x) {
  const double C[3] = HG_EXP2F_POLY_SCALED;
  const uint32_t ux = as_u32(x);
  const uint32_t abstop = (ux >> 20) & 0x7ffu;
  const double xd = (double)x;
  if (abstop > 0x42au) {                                   // |x| >= 88 or nan
    if (ux == 0xff800000u) return 0.0f;
    if (abstop >= 0x7f8u) return x + x;
    if (x > 0x1.62e42ep6f) return hm_inff();
    if (x < -0x1.9fe368p6f) return 0.0f;
    if (x < -0x1.9d1d9ep6f) return 0x1p-149f;              // __math_may_uflowf: 0x1.4p-75f squared, rounded
  }
  const double zs = hm_fma(HG_EXP2F_INVLN2_SCALED, xd, HG_EXP2F_SHIFT);
  const uint32_t ki = lo32(zs);
  const double kd = hm_add(zs, -HG_EXP2F_SHIFT);
  const double r = hm_fma(HG_EXP2F_INVLN2_SCALED, xd, -kd);
  const uint64_t t = HG_EXP2F(ki & 31u);
  const double s = from_words((uint32_t)(t >> 32) + (ki << 15), (uint32_t)t);          // t + (ki << 47)
  const double z = hm_fma(C[0], r, C[1]);
  const double r2 = hm_mul(r, r);
  double y = hm_fma(C[2], r, 1.0);
  y = hm_fma(z, r2, y);
  y = hm_mul(y, s);
  return (float)y;
}

// ==== logf (e_logf.c, FMA build) =======================================================================
HM_NOINLINE float r4_log(float x) {
  const double A[3] = HG_LOGF_A;
  uint32_t ix = as_u32(x);
  if (ix == 0x3f800000u) return 0.0f;
  if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u) {     // x < 2^-126, inf or nan
    if (ix * 2 == 0) return -hm_inff();
    if (ix == 0x7f800000u) return x;
    if ((ix & 0x80000000u) || ix * 2 >= 0xff000000u) return hm_nanf();
    ix = as_u32(x * 0x1p23f);                              // subnormal: normalise
    ix -= 23u << 23;
  }
  const uint32_t tmp = ix - 0x3f330000u;
  const int i = (int)((tmp >> 19) & 15u);
  const int k = (int32_t)tmp >> 23;
  const uint32_t iz = ix - (tmp & 0xff800000u);
  const GF2 e = HG_LOGF(i);
  const double z = (double)as_f32(iz);
  const double r = hm_fma(z, e.invc, -1.0);
  const double y0 = hm_fma((double)k, HG_LOGF_LN2, e.logc);
  const double r2 = hm_mul(r, r);
  double y = hm_fma(A[1], r, A[2]);
  y = hm_fma(A[0], r2, y);
  y = hm_fma(y, r2, hm_add(y0, r));
  return (float)y;
}

// ==== powf (e_powf.c, FMA build) =======================================================================
HM_FN int g_checkintf(uint32_t iy) {
  const int e = (int)(iy >> 23) & 0xff;
  if (e < 0x7f) return 0;
  if (e > 0x7f + 23) return 2;
  if (iy & ((1u << (0x7f + 23 - e)) - 1)) return 0;
  if (iy & (1u << (0x7f + 23 - e))) return 1;
  return 2;
}
HM_NOINLINE float r4_pow(float x, float y) {
  const double A[5] = HG_POWF_A;
  const double C[3] = HG_EXP2F_POLY;
  uint32_t sign_bias = 0;
  uint32_t ix = as_u32(x);
  const uint32_t iy = as_u32(y);
  const bool zin_y = (2 * iy - 1 >= 2u * 0x7f800000u - 1);
  if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u || zin_y) {
    if (zin_y) {
      if (2 * iy == 0) return 1.0f;
      if (ix == 0x3f800000u) return 1.0f;
      if (2 * ix > 2u * 0x7f800000u || 2 * iy > 2u * 0x7f800000u) return x + y;
      if (2 * ix == 2 * 0x3f800000u) return 1.0f;
      if ((2 * ix < 2 * 0x3f800000u) == !(iy & 0x80000000u)) return 0.0f;
      return y * y;
    }
    if (2 * ix - 1 >= 2u * 0x7f800000u - 1) {              // x is 0, inf or nan
      float x2 = x * x;
      if ((ix & 0x80000000u) && g_checkintf(iy) == 1) x2 = -x2;
      return (iy & 0x80000000u) ? 1.0f / x2 : x2;
    }
    if (ix & 0x80000000u) {                                 // finite x < 0
      const int yint = g_checkintf(iy);
      if (yint == 0) return hm_nanf();
      if (yint == 1) sign_bias = 0x10000u;                 // 1 << (EXP2F_TABLE_BITS + 11)
      ix &= 0x7fffffffu;
    }
    if (ix < 0x00800000u) {                                 // subnormal x
      ix = as_u32(x * 0x1p23f);
      ix &= 0x7fffffffu;
      ix -= 23u << 23;
    }
  }
  // log2_inline
  const uint32_t tmp = ix - 0x3f330000u;
  const int i = (int)((tmp >> 19) & 15u);
  const uint32_t top = tmp & 0xff800000u;
  const uint32_t iz = ix - top;
  const int k = (int32_t)top >> 23;
  const GF2 e = HG_POWFLOG2(i);
  const double z = (double)as_f32(iz);
  const double r = hm_fma(z, e.invc, -1.0);
  const double y0 = hm_add((double)k, e.logc);
  const double ya = hm_fma(A[0], r, A[1]);
  const double p = hm_fma(A[2], r, A[3]);
  const double r2 = hm_mul(r, r);
  double q = hm_fma(A[4], r, y0);
  const double r4 = hm_mul(r2, r2);
  q = hm_fma(p, r2, q);
  const double logx = hm_fma(ya, r4, q);
  const double ylogx = hm_mul((double)y, logx);
  if (((hi32(ylogx) >> 15) & 0xffffu) > 0x80beu) {         // |y log2 x| >= 126
    if (ylogx > 0x1.fffffffd1d571p+6) return sign_bias ? -hm_inff() : hm_inff();
    if (ylogx <= -150.0) return sign_bias ? -0.0f : 0.0f;
    if (ylogx < -149.0) return sign_bias ? -0x1p-149f : 0x1p-149f;
  }
  // exp2_inline
  const double kds = hm_add(ylogx, HG_EXP2F_SHIFT_SCALED);
  const uint32_t ki = lo32(kds);
  const double kd = hm_add(kds, -HG_EXP2F_SHIFT_SCALED);
  const double rr = hm_add(ylogx, -kd);
  const uint64_t t = HG_EXP2F(ki & 31u);
  const double s = from_words((uint32_t)(t >> 32) + ((ki + sign_bias) << 15), (uint32_t)t);
  const double zz = hm_fma(C[0], rr, C[1]);
  const double rr2 = hm_mul(rr, rr);
  double yy = hm_fma(C[2], rr, 1.0);
  yy = hm_fma(zz, rr2, yy);
  yy = hm_mul(yy, s);
  return (float)yy;
}

// ==== sinf / cosf (s_sinf.c, s_cosf.c, sincosf.h; FMA build) ==========================================
// table p (14 doubles): sign[4] 0..3, hpi_inv 4, hpi 5, c0 6, c1 7, s1 8, c2 9, s2 10, c3 11, s3 12, c4 13
HM_FN float g_sinf_poly(double x, double x2, int tb, int n) {
  if ((n & 1) == 0) {
    const double s1 = hm_fma(HG_SINCOSF(tb, 12), x2, HG_SINCOSF(tb, 10));
    const double x3 = hm_mul(x, x2);
    const double x7 = hm_mul(x3, x2);
    const double s = hm_fma(x3, HG_SINCOSF(tb, 8), x);
    return (float)hm_fma(s1, x7, s);
  }
  const double x4 = hm_mul(x2, x2);
  const double c1 = hm_fma(HG_SINCOSF(tb, 7), x2, HG_SINCOSF(tb, 6));
  const double c2 = hm_fma(HG_SINCOSF(tb, 13), x2, HG_SINCOSF(tb, 11));
  const double x6 = hm_mul(x4, x2);
  const double c = hm_fma(x4, HG_SINCOSF(tb, 9), c1);
  return (float)hm_fma(c2, x6, c);
}
// |x| >= 120: 192-bit 4/pi multiplication (reduce_large)
HM_FN double g_reduce_large(uint32_t xi, int& n) {
  const int a = (int)((xi >> 26) & 15u);
  const int shift = (int)((xi >> 23) & 7u);
  xi = (xi & 0xffffffu) | 0x800000u;
  xi <<= shift;
  uint64_t res0 = (uint64_t)(xi * HG_INVPIO4(a));          // 32-bit product
  const uint64_t res1 = (uint64_t)xi * HG_INVPIO4(a + 4);
  const uint64_t res2 = (uint64_t)xi * HG_INVPIO4(a + 8);
  res0 = (res2 >> 32) | (res0 << 32);
  res0 += res1;
  const uint64_t nn = (res0 + (1ull << 61)) >> 62;
  res0 -= nn << 62;
  n = (int)nn;
  return hm_mul((double)(int64_t)res0, HG_PIO2_2M64);
}
HM_FN float g_sincosf(float y, bool want_cos) {
  const uint32_t uy = as_u32(y);
  const uint32_t abstop = (uy >> 20) & 0x7ffu;
  const double x = (double)y;
  if (abstop <= 0x3f3u) {                                   // |y| < pi/4
    const double x2 = hm_mul(x, x);
    if (abstop <= 0x397u) return want_cos ? 1.0f : y;       // |y| < 2^-12
    return g_sinf_poly(x, x2, 0, want_cos ? 1 : 0);
  }
  int n, tb;
  double xr, s;
  if (abstop <= 0x42eu) {                                   // |y| < 120: reduce_fast
    const double r = hm_mul(x, HG_SINCOSF(0, 4));
    n = ((int32_t)r + 0x800000) >> 24;
    xr = hm_fma(-(double)n, HG_SINCOSF(0, 5), x);
    s = HG_SINCOSF(0, n & 3);
    tb = (n & 2) ? 1 : 0;
  } else if (abstop <= 0x7f7u) {
    const int sign = (int)(uy >> 31);
    xr = g_reduce_large(uy, n);
    s = HG_SINCOSF(0, (n + sign) & 3);
    tb = ((n + sign) & 2) ? 1 : 0;
  } else {
    return hm_nanf();                                       // inf or nan
  }
  return g_sinf_poly(hm_mul(xr, s), hm_mul(xr, xr), tb, want_cos ? (n ^ 1) : n);
}
HM_NOINLINE float r4_sin(float x) { return g_sincosf(x, false); }
HM_NOINLINE float r4_cos(float x) { return g_sincosf(x, true); }

// ==== tanf (s_tanf.c + k_tanf.c, REAL*4 arithmetic, baseline build: no contraction) ===================
HM_FN float g_kernel_tanf(float x, float y, int iy) {
  const float T[13] = HG_TANF_T;
  const uint32_t hx = as_u32(x);
  const uint32_t ix = hx & 0x7fffffffu;
  if (ix < 0x39000000u) {                                   // |x| < 2^-13
    if ((int)x == 0) {
      if ((ix | (uint32_t)(iy + 1)) == 0) return 1.0f / (x < 0.0f ? -x : x);
      if (iy == 1) return x;
      return -1.0f / x;
    }
  }
  if (ix >= 0x3f2ca140u) {                                  // |x| >= 0.6744
    if ((int32_t)hx < 0) { x = -x; y = -y; }
    const float z = HG_TANF_PIO4 - x;
    const float w = HG_TANF_PIO4LO - y;
    x = z + w; y = 0.0f;
    if ((x < 0.0f ? -x : x) < 0x1p-13f) return (float)((1 - (int)((hx >> 30) & 2u)) * iy) * (1.0f - (float)(2 * iy) * x);
  }
  float z = x * x;
  float w = z * z;
  float r = T[1] + w * (T[3] + w * (T[5] + w * (T[7] + w * (T[9] + w * T[11]))));
  float v = z * (T[2] + w * (T[4] + w * (T[6] + w * (T[8] + w * (T[10] + w * T[12])))));
  float s = z * x;
  r = y + z * (s * (v + r) + y);
  r = T[0] * s + r;
  w = x + r;
  if (ix >= 0x3f2ca140u) {
    v = (float)iy;
    return (float)(1 - (int)((hx >> 30) & 2u)) * (v - 2.0f * (x - (w * w / (w + v) - r)));
  }
  if (iy == 1) return w;
  // -1/(x + r), accurately
  z = as_f32(as_u32(w) & 0xfffff000u);
  v = r - (z - x);
  const float a = -1.0f / w;
  const float t = as_f32(as_u32(a) & 0xfffff000u);
  s = 1.0f + t * z;
  return t + a * (s + t * v);
}
HM_NOINLINE float r4_tan(float x) {
  const uint32_t ux = as_u32(x);
  const uint32_t ix = ux & 0x7fffffffu;
  if (ix <= 0x3f490fdau) return g_kernel_tanf(x, 0.0f, 1);          // |x| <= pi/4
  if (ix > 0x7f7fffffu) return hm_nanf();                            // inf or nan: x - x
  const uint32_t abstop = (ux >> 20) & 0x7ffu;
  double xd = (double)x;
  int n;
  if (abstop <= 0x42eu) {                                   // |x| < 120 (__ieee754_rem_pio2f, double arithmetic, no fma)
    const double r = hm_mul(xd, HG_SINCOSF(0, 4));
    n = ((int32_t)r + 0x800000) >> 24;
    xd = hm_add(xd, -hm_mul((double)n, HG_SINCOSF(0, 5)));
  } else {
    const int sign = (int)(ux >> 31);
    xd = g_reduce_large(ux, n);
    if (sign) { xd = -xd; n = -n; }
  }
  const float y0 = (float)xd;
  const float y1 = (float)hm_add(xd, -(double)y0);
  return g_kernel_tanf(y0, y1, 1 - ((n & 1) << 1));
}

// ==== acosf (e_acosf.c, REAL*4 arithmetic, baseline build) ============================================
HM_NOINLINE float r4_acos(float x) {
  const float pS[6] = HG_ACOSF_PS, qS[4] = HG_ACOSF_QS;
  const uint32_t hx = as_u32(x);
  const uint32_t ix = hx & 0x7fffffffu;
  if (ix == 0x3f800000u) return ((int32_t)hx > 0) ? 0.0f : HG_ACOSF_PI + HG_ACOSF_2PIO2LO;
  if (ix > 0x3f800000u) return hm_nanf();
  if (ix < 0x3f000000u) {                                   // |x| < 0.5
    if (ix <= 0x32800000u) return HG_ACOSF_PIO2HI + HG_ACOSF_PIO2LO;
    const float z = x * x;
    const float p = z * (pS[0] + z * (pS[1] + z * (pS[2] + z * (pS[3] + z * (pS[4] + z * pS[5])))));
    const float q = 1.0f + z * (qS[0] + z * (qS[1] + z * (qS[2] + z * qS[3])));
    const float r = p / q;
    return HG_ACOSF_PIO2HI - (x - (HG_ACOSF_PIO2LO - r * x));
  }
  if ((int32_t)hx < 0) {                                    // x < -0.5
    const float z = (1.0f + x) * 0.5f;
    const float p = z * (pS[0] + z * (pS[1] + z * (pS[2] + z * (pS[3] + z * (pS[4] + z * pS[5])))));
    const float q = 1.0f + z * (qS[0] + z * (qS[1] + z * (qS[2] + z * qS[3])));
    const float s = hm_sqrtf(z);
    const float r = p / q;
    const float w = r * s - HG_ACOSF_PIO2LO;
    return HG_ACOSF_PI - 2.0f * (s + w);
  }
  const float z = (1.0f - x) * 0.5f;                        // x > 0.5
  const float s = hm_sqrtf(z);
  const float df = as_f32(as_u32(s) & 0xfffff000u);
  const float c = (z - df * df) / (s + df);
  const float p = z * (pS[0] + z * (pS[1] + z * (pS[2] + z * (pS[3] + z * (pS[4] + z * pS[5])))));
  const float q = 1.0f + z * (qS[0] + z * (qS[1] + z * (qS[2] + z * qS[3])));
  const float r = p / q;
  const float w = r * s + c;
  return 2.0f * (df + w);
}

// ==== log10f (e_log10f.c, baseline build; calls logf) =================================================
HM_NOINLINE float r4_log10(float x) {
  int32_t hx = (int32_t)as_u32(x);
  int32_t k = 0;
  if (hx < 0x00800000) {
    if ((hx & 0x7fffffff) == 0) return -hm_inff();
    if (hx < 0) return hm_nanf();
    k -= 25; x = x * HG_LOG10F_TWO25;
    hx = (int32_t)as_u32(x);
  }
  if (hx >= 0x7f800000) return x + x;
  k += (hx >> 23) - 127;
  const int32_t i = (int32_t)(((uint32_t)k & 0x80000000u) >> 31);
  hx = (hx & 0x007fffff) | ((0x7f - i) << 23);
  const float y = (float)(k + i);
  const float xm = as_f32((uint32_t)hx);
  const float z = y * HG_LOG10F_2LO + HG_LOG10F_IVLN10 * r4_log(xm);
  return z + y * HG_LOG10F_2HI;
}

// x**0.5, x**8., x**7., x**4. as a Fortran compiler without -ffast-math evaluates them: a call of powf
HM_FN float r4_sqrt(float x) { return r4_pow(x, 0.5f); }
HM_FN float r4_pow8(float x) { return r4_pow(x, 8.0f); }
HM_FN float r4_pow7(float x) { return r4_pow(x, 7.0f); }
HM_FN float r4_pow4(float x) { return r4_pow(x, 4.0f); }

}  // namespace helpmath
