// Micro-benchmark: clocks per call of the engine's transcendentals on special arguments (NaN, negative, zero, infinity) next
// to an ordinary one -- a cell HELP returns NaN for must not be slower than a healthy one (the launch waits for it).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -I../../pyhelp_b200/csrc -o special_args special_args.cu
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
#include "help_math.h"
using namespace helpmath;

template <int FN>
__global__ void k(double* out, long long* cyc, int iters, double x0, double y0) {
  hm_load_tables();
  double x = x0, acc = 0.0;
  const double y = y0;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    double r;
    if (FN == 0) r = det_pow(x, y);
    if (FN == 1) { const Pair R = det_pow_samebase(x, y, y - 1.0); r = R.a + R.b; }
    if (FN == 2) r = (double)r4_pow((float)x, (float)y);
    if (FN == 3) r = (double)r4_exp((float)x);
    if (FN == 4) r = (double)r4_sin((float)x);
    if (FN == 5) r = (double)r4_cos((float)x);
    if (FN == 6) r = (double)r4_tan((float)x);
    if (FN == 7) r = (double)r4_acos((float)x);
    if (FN == 8) r = div_by_const(x, y, 1.0 / y);
    if (FN == 9) r = det_log(x);
    acc += r;
    asm volatile("" : "+d"(x));          // keep the call in the loop, same argument every time
  }
  long long t1 = clock64();
  out[threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 8 * 64); cudaMalloc(&cyc, 8);
  const int iters = 20000;
  const char* names[10] = {"det_pow", "det_pow_samebase", "r4_pow", "r4_exp", "r4_sin", "r4_cos", "r4_tan", "r4_acos", "div_by_const", "det_log"};
  const double qn = NAN, inf = INFINITY;
  const double xs[] = {0.37, qn, -0.37, 0.0, inf, 1e-310, -inf};
  const char* xn[] = {"0.37", "nan", "-0.37", "0", "inf", "1e-310", "-inf"};
  for (int fn = 0; fn < 10; ++fn) {
    printf("%-18s", names[fn]);
    for (int a = 0; a < 7; ++a) {
      for (int rep = 0; rep < 2; ++rep) {
        switch (fn) {
          case 0: k<0><<<1, 32>>>(out, cyc, iters, xs[a], 3.7); break;
          case 1: k<1><<<1, 32>>>(out, cyc, iters, xs[a], 3.7); break;
          case 2: k<2><<<1, 32>>>(out, cyc, iters, xs[a], 0.3); break;
          case 3: k<3><<<1, 32>>>(out, cyc, iters, xs[a], 0.0); break;
          case 4: k<4><<<1, 32>>>(out, cyc, iters, xs[a], 0.0); break;
          case 5: k<5><<<1, 32>>>(out, cyc, iters, xs[a], 0.0); break;
          case 6: k<6><<<1, 32>>>(out, cyc, iters, xs[a], 0.0); break;
          case 7: k<7><<<1, 32>>>(out, cyc, iters, xs[a], 0.0); break;
          case 8: k<8><<<1, 32>>>(out, cyc, iters, xs[a], 0.73); break;
          case 9: k<9><<<1, 32>>>(out, cyc, iters, xs[a], 0.0); break;
        }
        cudaDeviceSynchronize();
      }
      long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
      printf("  %s: %6.0f", xn[a], (double)c / iters);
    }
    printf("\n");
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
