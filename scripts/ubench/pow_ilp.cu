// Micro-benchmark: does a warp get instruction-level parallelism out of two independent powers?
// Chains of dependent powers (the routing loop's situation: every power's argument comes from the one before) timed on
// ONE SM with 1 / 2 / 4 / 6 warps per scheduler:
//   seq    x <- pow(f(x)), y <- pow(f(y)) one after the other (two det_pow calls)
//   pow2   det_pow2: both in one call, chains written one after the other (what the compiler schedules sequentially)
//   pow2i  det_pow2i: the same operations written statement by statement for both chains (interleaved source)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -I../../pyhelp_b200/csrc -o pow_ilp pow_ilp.cu && ./pow_ilp
#include <cstdio>
#include <cuda_runtime.h>
#include "help_math.h"
using namespace helpmath;

// two powers, statement-interleaved; hot path only (x positive normal, moderate y), else the plain calls
__device__ __noinline__ Pair det_pow2i(double x0, double y0, double x1, double y1) {
  Pair R;
  if (!(pow_hot(x0, y0) && pow_hot(x1, y1))) { R.a = det_pow(x0, y0); R.b = det_pow(x1, y1); return R; }
  const double A[7] = HG_POW_A;
  const uint32_t h0 = hi32(x0), h1 = hi32(x1);
  const uint32_t t0 = h0 - 0x3fe69555u, t1 = h1 - 0x3fe69555u;
  const int i0 = (int)((t0 >> 13) & 127u), i1 = (int)((t1 >> 13) & 127u);
  const int k0 = (int32_t)t0 >> 20, k1 = (int32_t)t1 >> 20;
  const double z0 = from_words(h0 - (t0 & 0xfff00000u), lo32(x0)), z1 = from_words(h1 - (t1 & 0xfff00000u), lo32(x1));
  const double kd0 = hm_add(from_words(0x43300000u, (uint32_t)k0 ^ 0x80000000u), -4503601774854144.0);
  const double kd1 = hm_add(from_words(0x43300000u, (uint32_t)k1 ^ 0x80000000u), -4503601774854144.0);
  const LogEnt e0 = HG_POWLOG(i0), e1 = HG_POWLOG(i1);
  const double r0 = hm_fma(z0, e0.invc, -1.0), r1 = hm_fma(z1, e1.invc, -1.0);
  const double u0 = hm_fma(kd0, HG_POW_LN2HI, e0.logc), u1 = hm_fma(kd1, HG_POW_LN2HI, e1.logc);
  const double v0 = hm_add(u0, r0), v1 = hm_add(u1, r1);
  const double lo10 = hm_fma(kd0, HG_POW_LN2LO, e0.logctail), lo11 = hm_fma(kd1, HG_POW_LN2LO, e1.logctail);
  const double lo20 = hm_add(hm_add(u0, -v0), r0), lo21 = hm_add(hm_add(u1, -v1), r1);
  const double ar0 = hm_mul(A[0], r0), ar1 = hm_mul(A[0], r1);
  const double ar20 = hm_mul(r0, ar0), ar21 = hm_mul(r1, ar1);
  const double ar30 = hm_mul(r0, ar20), ar31 = hm_mul(r1, ar21);
  const double hi0 = hm_add(v0, ar20), hi1 = hm_add(v1, ar21);
  const double lo30 = hm_fma(ar0, r0, -ar20), lo31 = hm_fma(ar1, r1, -ar21);
  const double lo40 = hm_add(hm_add(v0, -hi0), ar20), lo41 = hm_add(hm_add(v1, -hi1), ar21);
  const double q120 = hm_fma(r0, A[2], A[1]), q121 = hm_fma(r1, A[2], A[1]);
  const double q340 = hm_fma(r0, A[4], A[3]), q341 = hm_fma(r1, A[4], A[3]);
  const double q560 = hm_fma(r0, A[6], A[5]), q561 = hm_fma(r1, A[6], A[5]);
  const double p0 = hm_fma(ar20, hm_fma(q560, ar20, q340), q120), p1 = hm_fma(ar21, hm_fma(q561, ar21, q341), q121);
  const double lo0 = hm_fma(ar30, p0, hm_add(hm_add(hm_add(lo10, lo20), lo30), lo40));
  const double lo1 = hm_fma(ar31, p1, hm_add(hm_add(hm_add(lo11, lo21), lo31), lo41));
  const double ly0 = hm_add(hi0, lo0), ly1 = hm_add(hi1, lo1);
  const double lt0 = hm_add(hm_add(hi0, -ly0), lo0), lt1 = hm_add(hm_add(hi1, -ly1), lo1);
  const double eh0 = hm_mul(y0, ly0), eh1 = hm_mul(y1, ly1);
  const double el0 = hm_fma(y0, lt0, hm_fma(ly0, y0, -eh0)), el1 = hm_fma(y1, lt1, hm_fma(ly1, y1, -eh1));
  // exp, straight-line case for both (2^-54 <= |x| < 512), else the library-order code
  const uint32_t at0 = (hi32(eh0) >> 20) & 0x7ffu, at1 = (hi32(eh1) >> 20) & 0x7ffu;
  if (!((at0 - 0x3c9u < 0x3fu) && (at1 - 0x3c9u < 0x3fu))) { R.a = g_exp_inline(eh0, el0, 0); R.b = g_exp_inline(eh1, el1, 0); return R; }
  const double C[4] = HG_EXP_C;
  const double ks0 = hm_fma(eh0, HG_EXP_INVLN2N, HG_EXP_SHIFT), ks1 = hm_fma(eh1, HG_EXP_INVLN2N, HG_EXP_SHIFT);
  const uint32_t ki0 = lo32(ks0), ki1 = lo32(ks1);
  const double kk0 = hm_add(ks0, -HG_EXP_SHIFT), kk1 = hm_add(ks1, -HG_EXP_SHIFT);
  double s0 = hm_fma(kk0, HG_EXP_NEGLN2HIN, eh0), s1 = hm_fma(kk1, HG_EXP_NEGLN2HIN, eh1);
  s0 = hm_fma(kk0, HG_EXP_NEGLN2LON, s0); s1 = hm_fma(kk1, HG_EXP_NEGLN2LON, s1);
  s0 = hm_add(s0, el0); s1 = hm_add(s1, el1);
  const GExpEntBits x0e = HG_EXPT(ki0 & 127u), x1e = HG_EXPT(ki1 & 127u);
  const uint32_t sh0 = (uint32_t)(x0e.sbits >> 32) + (ki0 << 13), sh1 = (uint32_t)(x1e.sbits >> 32) + (ki1 << 13);
  const double tl0 = as_f64(x0e.tail), tl1 = as_f64(x1e.tail);
  const double s20 = hm_mul(s0, s0), s21 = hm_mul(s1, s1);
  const double a0 = hm_fma(hm_fma(s0, C[1], C[0]), s20, hm_add(tl0, s0)), a1 = hm_fma(hm_fma(s1, C[1], C[0]), s21, hm_add(tl1, s1));
  const double m0 = hm_fma(hm_fma(s0, C[3], C[2]), hm_mul(s20, s20), a0), m1 = hm_fma(hm_fma(s1, C[3], C[2]), hm_mul(s21, s21), a1);
  const double sc0 = from_words(sh0, (uint32_t)x0e.sbits), sc1 = from_words(sh1, (uint32_t)x1e.sbits);
  R.a = hm_fma(sc0, m0, sc0); R.b = hm_fma(sc1, m1, sc1);
  return R;
}

template <int MODE>
__global__ void k(double* out, long long* cyc, int iters, double seed) {
  hm_load_tables();
  double x = seed + threadIdx.x * 1e-4, y = seed * 0.7 + threadIdx.x * 2e-4;
  const double ex = 3.7, ey = 4.3;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const double ax = 0.2 + 0.6 * x, ay = 0.25 + 0.5 * y;           // arguments in (0, 1), like relative saturations
    if (MODE == 0) { x = det_pow(ax, ex); y = det_pow(ay, ey); }
    if (MODE == 1) { const Pair R = det_pow2(ax, ex, ay, ey); x = R.a; y = R.b; }
    if (MODE == 2) { const Pair R = det_pow2i(ax, ex, ay, ey); x = R.a; y = R.b; }
    if (MODE == 3) { x = det_pow(ax, ex); }                          // one chain only
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = x + y;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 8 * 4096); cudaMalloc(&cyc, 8 * 16);
  const int iters = 20000;
  const char* names[4] = {"seq (2 det_pow)", "det_pow2", "det_pow2i", "one det_pow"};
  double ref[3];
  for (int threads : {32, 128, 256, 512, 768}) {
    for (int mode = 0; mode < 4; ++mode) {
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) k<0><<<1, threads>>>(out, cyc, iters, 0.4);
        if (mode == 1) k<1><<<1, threads>>>(out, cyc, iters, 0.4);
        if (mode == 2) k<2><<<1, threads>>>(out, cyc, iters, 0.4);
        if (mode == 3) k<3><<<1, threads>>>(out, cyc, iters, 0.4);
        cudaDeviceSynchronize();
      }
      long long c; double v;
      cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost); cudaMemcpy(&v, out, 8, cudaMemcpyDeviceToHost);
      if (mode < 3) ref[mode] = v;
      printf("threads %4d  %-16s %8.1f clk per iteration   (result %.17g)\n", threads, names[mode], (double)c / iters, v);
    }
    printf("   bits equal: %s\n", (ref[0] == ref[1] && ref[0] == ref[2]) ? "yes" : "NO");
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
