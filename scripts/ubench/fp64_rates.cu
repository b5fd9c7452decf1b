// Micro-benchmark: issue rates of the operations the HELP routing loop is made of, on one
// B200 SM (FP64 add/mul/fma, float<->double conversions, FP64 compare/select, reciprocal
// seed, full IEEE division, 64-bit loads from shared/L1).  Secondary roofline denominator
// ("measure the FP64 DFMA peak on the box").
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_rates fp64_rates.cu && ./fp64_rates
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
enum { OP_DFMA, OP_DADD, OP_DMUL, OP_CVT_F2D, OP_CVT_D2F, OP_CVT_PAIR, OP_DSETP_SEL, OP_RCP64H, OP_DDIV, OP_IMAD, OP_FFMA,
       OP_LDS64, OP_LDG64, OP_DMNMX, OP_N };
static const char* kNames[] = {"dfma", "dadd", "dmul", "cvt.f64.f32", "cvt.rn.f32.f64", "cvt pair (round to f32)", "dsetp+selp.f64",
                               "rcp.approx.ftz.f64", "div.rn.f64", "imad", "ffma", "ld.shared.f64", "ld.global.f64 (L1 hit)", "min.f64"};
template <int OP, int CH>
__global__ void k(double* out, const double* in, long long* cyc, double seed) {
  __shared__ double sm[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = in[i];
  __syncthreads();
  double a[CH]; float f[CH]; int n[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c) { a[c] = seed + c * 0.125 + threadIdx.x * 1e-3; f[c] = (float)a[c]; n[c] = c + threadIdx.x; }
  const double b = 1.0000001, d = 1e-9;
  const int idx = threadIdx.x & 1023;
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      if (OP == OP_DFMA) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(a[c]) : "d"(b), "d"(d));
      if (OP == OP_DADD) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(a[c]) : "d"(d));
      if (OP == OP_DMUL) asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(a[c]) : "d"(b));
      if (OP == OP_CVT_F2D) { asm volatile("cvt.f64.f32 %0, %1;" : "=d"(a[c]) : "f"(f[c])); asm volatile("" : "+f"(f[c]) : "d"(a[c])); }
      if (OP == OP_CVT_D2F) { asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(f[c]) : "d"(a[c])); asm volatile("" : "+d"(a[c]) : "f"(f[c])); }
      if (OP == OP_CVT_PAIR) { asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(f[c]) : "d"(a[c])); asm volatile("cvt.f64.f32 %0, %1;" : "=d"(a[c]) : "f"(f[c])); }
      if (OP == OP_DSETP_SEL) asm volatile("{.reg .pred p; setp.lt.f64 p, %0, %1; selp.f64 %0, %1, %0, p;}" : "+d"(a[c]) : "d"(b));
      if (OP == OP_RCP64H) asm volatile("rcp.approx.ftz.f64 %0, %0;" : "+d"(a[c]));
      if (OP == OP_DDIV) asm volatile("div.rn.f64 %0, %1, %0;" : "+d"(a[c]) : "d"(b));
      if (OP == OP_IMAD) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(n[c]) : "r"(n[(c + 1) % CH]), "r"(it));
      if (OP == OP_FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[c]) : "f"(1.0000001f), "f"(1e-9f));
      if (OP == OP_LDS64) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"((unsigned)__cvta_generic_to_shared(&sm[(idx + c * 32 + (it & 7) * 64) & 1023]))); a[c] += 0; asm volatile("" : "+d"(a[c]) : "d"(v)); }
      if (OP == OP_LDG64) { double v; asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(in + ((idx + c * 32 + (it & 7) * 64) & 1023))); asm volatile("" : "+d"(a[c]) : "d"(v)); }
      if (OP == OP_DMNMX) asm volatile("min.f64 %0, %0, %1;" : "+d"(a[c]) : "d"(b));
    }
  }
  long long t1 = clock64();
  double s = 0; for (int c = 0; c < CH; ++c) s += a[c] + f[c] + n[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int OP, int CH> void run(int threads, double* out, const double* in, long long* cyc) {
  k<OP, CH><<<1, threads>>>(out, in, cyc, 1.5);
  cudaDeviceSynchronize();
  k<OP, CH><<<1, threads>>>(out, in, cyc, 1.5);
  cudaDeviceSynchronize();
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  const double ops = (double)ITERS * CH * threads;
  printf("  %-26s threads %4d chains %d: %8.2f thread-ops/clk/SM  (%.2f clk per warp-instr per SMSP, latency-bound if chains=1: %.1f clk/op)\n",
         kNames[OP], threads, CH, ops / c, (double)c / (ITERS * CH * (threads / 128.0)), (double)c / (ITERS * CH));
}
template <int OP> void all(double* out, const double* in, long long* cyc) {
  run<OP, 1>(32, out, in, cyc);        // latency
  run<OP, 8>(128, out, in, cyc);
  run<OP, 8>(512, out, in, cyc);       // throughput
  run<OP, 4>(1024, out, in, cyc);
}
int main() {
  double *out, *in; long long* cyc;
  cudaMalloc(&out, 1024 * 8 * 16); cudaMalloc(&in, 1024 * 8); cudaMalloc(&cyc, 8 * 16);
  double h[1024]; for (int i = 0; i < 1024; ++i) h[i] = 1.0 + i * 1e-3;
  cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  printf("%s, %d SMs, clock %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
  all<OP_DFMA>(out, in, cyc); all<OP_DADD>(out, in, cyc); all<OP_DMUL>(out, in, cyc); all<OP_CVT_F2D>(out, in, cyc);
  all<OP_CVT_D2F>(out, in, cyc); all<OP_CVT_PAIR>(out, in, cyc); all<OP_DSETP_SEL>(out, in, cyc); all<OP_RCP64H>(out, in, cyc);
  all<OP_DDIV>(out, in, cyc); all<OP_IMAD>(out, in, cyc); all<OP_FFMA>(out, in, cyc); all<OP_LDS64>(out, in, cyc);
  all<OP_LDG64>(out, in, cyc); all<OP_DMNMX>(out, in, cyc);
  printf("status %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
