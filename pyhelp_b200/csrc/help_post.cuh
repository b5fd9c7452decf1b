// help_post.cuh -- the reductions PyHELP runs on the result of the hot path, on the device
// (SURVEY.md 8f row 3), straight from the engine's monthly[7][year][12][cell] buffer:
//   HelpManager._post_process_output      pyhelp/managers.py:448-506   NaN cells, recharge of
//                                          cells next to a stream (context 2) turned into
//                                          subsurface runoff
//   HelpOutput.calc_area_monthly_avg      pyhelp/output.py:127-152     nansum over cells / Np
//   HelpOutput.calc_cells_yearly_avg      pyhelp/output.py:171-203     per cell: mean over the
//                                          selected years of the yearly sums
// so that a 5 M-cell run never materialises 7 x (Np, Ny, 12) float64 arrays on the host.
// Rows are the engine's (help_b200.h HR_*): precip, runoff, evapo, subrun1, subrun2, perco, rechg.
// Values are REAL*4 multiples of 0.001 mm; every sum is accumulated in double, in a fixed
// order (deterministic; numpy's pairwise order differs in the last bits).
#pragma once
#include <stdint.h>

namespace helpb200 {

// per cell: mode 0 as is, 1 NaN cell (rechg has a NaN), 2 rechg -> subrun1, 3 rechg -> subrun2
__global__ void post_cell_kernel(const float* __restrict__ monthly, const int32_t* __restrict__ context, int64_t n, int ny,
                                 int y0, int y1, int8_t* __restrict__ mode, double* __restrict__ cells_yearly) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  const int64_t rs = (int64_t)ny * 12 * n;                  // row stride
  bool nan = false, sub2_zero = true;
  for (int ym = 0; ym < ny * 12; ++ym) {
    const float r = monthly[6 * rs + (int64_t)ym * n + c], s2 = monthly[4 * rs + (int64_t)ym * n + c];
    nan = nan || (r != r);
    sub2_zero = sub2_zero && (s2 == 0.0f);
  }
  int m = 0;
  if (nan) m = 1;
  else if (context && context[c] == 2) m = sub2_zero ? 2 : 3;
  mode[c] = (int8_t)m;
  if (!cells_yearly) return;
  double acc[7] = {0, 0, 0, 0, 0, 0, 0};
  int nsel = 0;
  for (int y = y0; y < y1; ++y) {
    ++nsel;
    double ysum[7] = {0, 0, 0, 0, 0, 0, 0};
    for (int mo = 0; mo < 12; ++mo) {
      const int64_t o = ((int64_t)y * 12 + mo) * n + c;
      double v[7];
#pragma unroll
      for (int r = 0; r < 7; ++r) v[r] = (double)monthly[r * rs + o];
      if (m == 2) { v[3] += v[6]; v[6] = 0.0; }
      if (m == 3) { v[4] += v[6]; v[6] = 0.0; }
#pragma unroll
      for (int r = 0; r < 7; ++r) ysum[r] += v[r];
    }
#pragma unroll
    for (int r = 0; r < 7; ++r) acc[r] += ysum[r];
  }
#pragma unroll
  for (int r = 0; r < 7; ++r) cells_yearly[(int64_t)r * n + c] = nsel ? acc[r] / nsel : __longlong_as_double(0x7ff8000000000000LL);
}

// one block per (year, month): nansum over the cells of the 7 (adjusted) rows -> area[7][ny*12]
__global__ void __launch_bounds__(256) post_area_kernel(const float* __restrict__ monthly, const int8_t* __restrict__ mode,
                                                        int64_t n, int ny, double* __restrict__ area, int32_t* __restrict__ n_nan) {
  __shared__ double sh[7][256];
  __shared__ int shn[256];
  const int ym = blockIdx.x;
  const int64_t rs = (int64_t)ny * 12 * n;
  double acc[7] = {0, 0, 0, 0, 0, 0, 0};
  int nn = 0;
  for (int64_t c = threadIdx.x; c < n; c += blockDim.x) {
    const int m = mode[c];
    nn += (m == 1);
    double v[7];
#pragma unroll
    for (int r = 0; r < 7; ++r) v[r] = (double)monthly[r * rs + (int64_t)ym * n + c];
    if (m == 2) { v[3] += v[6]; v[6] = 0.0; }
    if (m == 3) { v[4] += v[6]; v[6] = 0.0; }
#pragma unroll
    for (int r = 0; r < 7; ++r) acc[r] += (v[r] == v[r]) ? v[r] : 0.0;       // nansum
  }
#pragma unroll
  for (int r = 0; r < 7; ++r) sh[r][threadIdx.x] = acc[r];
  shn[threadIdx.x] = nn;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) {
#pragma unroll
      for (int r = 0; r < 7; ++r) sh[r][threadIdx.x] += sh[r][threadIdx.x + s];
      shn[threadIdx.x] += shn[threadIdx.x + s];
    }
    __syncthreads();
  }
  if (threadIdx.x < 7) area[(int64_t)threadIdx.x * ny * 12 + ym] = sh[threadIdx.x][0];
  if (threadIdx.x == 0 && ym == 0) *n_nan = shn[0];
}

// ---- [row*year*month][cell] -> [cell][row*year*month] ------------------------------------------------------
// The reference returns one dict per cell (pyhelp/processing.py:139-146): the per-cell view of the engine's
// cell-minor monthly buffer is a transposition of a (7*Ny*12) x n matrix -- 1 GB for 100 000 cells x 30 yr, a second
// on one host core, a millisecond here.  32 x 32 tiles through shared memory, both sides coalesced.
__global__ void transpose_to_cell_major_kernel(const float* __restrict__ in, int64_t rows, int64_t n, float* __restrict__ out) {
  __shared__ float tile[32][33];
  const int64_t c0 = (int64_t)blockIdx.x * 32, r0 = (int64_t)blockIdx.y * 32;
  for (int k = threadIdx.y; k < 32; k += 8) {
    const int64_t r = r0 + k, c = c0 + threadIdx.x;
    tile[k][threadIdx.x] = (r < rows && c < n) ? in[r * n + c] : 0.0f;
  }
  __syncthreads();
  for (int k = threadIdx.y; k < 32; k += 8) {
    const int64_t c = c0 + k, r = r0 + threadIdx.x;
    if (c < n && r < rows) out[c * rows + r] = tile[threadIdx.x][k];
  }
}

}  // namespace helpb200
