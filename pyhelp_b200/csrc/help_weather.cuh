// help_weather.cuh -- the reference's weather-file round trip on the device.
//
// PyHELP hands weather to HELP3O as text: save_precip_to_HELP / save_airtemp_to_HELP /
// save_solrad_to_HELP (pyhelp/weather_reader.py:22-104) print every daily value of a node with
// '{:>5.1f}' (mm), '{:>6.1f}' (deg C) or '{:>6.2f}' (MJ/m2), 37 records of 10 values per year,
// and READCD (pyhelp/HELP3O.FOR:1100-1129) reads them back into REAL*4.  The numbers the
// simulation sees are therefore float32(float(format(x, '.Nf'))) -- the quantisation is part
// of the function.  pack_weather_kernel computes exactly that for a whole
// [day][node] table and writes the engine's [node][year][kDays] layout, so that the host never
// formats, parses or transposes 10^8 values (numpy: 0.4 s per variable per 1000 nodes x 30 yr).
//
// HBM-bound: 8 B read + 4 B written per node-day; a 32 x 32 tile goes through shared memory so
// that both the read (node-minor) and the write (day-minor) are coalesced.
#pragma once
#include <stdint.h>

#include "help_image.h"

namespace helpb200 {

// float(format(x, '.{d}f')) for S = 10^d: the exact product x*S rounded to the nearest integer,
// ties to even -- Python formats the exact binary value -- then the correctly rounded quotient
// (== strtod of the printed decimal: r < 2^53 and S <= 10^22 are exact doubles).
__device__ __forceinline__ double quantize_decimals(double x, double S) {
  const double p = __dmul_rn(x, S);
  const double e = __fma_rn(x, S, -p);          // x*S == p + e exactly
  double r = rint(p);
  const double diff = __dadd_rn(p, -r);         // exact
  if (diff == 0.5 && e > 0.0) r = r + 1.0;      // the rounded product sits on a tie the exact one is past
  else if (diff == -0.5 && e < 0.0) r = r - 1.0;
  return __dadd_rn(__ddiv_rn(r, S), 0.0);       // (-0.0 -> 0.0)
}

// table[day][node] (float64) -> out[node][year][kDays] (float32, zero beyond the year's days).
// grid: (ceil(kDays/32), n_years, ceil(n_nodes/32)); block (32, 8).
__global__ void pack_weather_kernel(const double* __restrict__ table, int64_t n_nodes, const int32_t* __restrict__ year_day0,
                                    const int32_t* __restrict__ year_ndays, int n_years, double S, float* __restrict__ out) {
  __shared__ float tile[32][33];
  const int y = blockIdx.y;
  const int d0 = year_day0[y], nd = year_ndays[y];
  const int dd0 = blockIdx.x * 32;
  const int64_t node0 = (int64_t)blockIdx.z * 32;
  for (int k = threadIdx.y; k < 32; k += 8) {           // rows = days of the tile, lanes = nodes
    const int dd = dd0 + k;
    const int64_t node = node0 + threadIdx.x;
    float v = 0.0f;
    if (dd < nd && node < n_nodes) v = (float)quantize_decimals(table[(int64_t)(d0 + dd) * n_nodes + node], S);
    tile[k][threadIdx.x] = v;
  }
  __syncthreads();
  for (int k = threadIdx.y; k < 32; k += 8) {           // rows = nodes, lanes = days
    const int64_t node = node0 + k;
    const int dd = dd0 + threadIdx.x;
    if (node < n_nodes && dd < kDays) out[(node * n_years + y) * kDays + dd] = tile[threadIdx.x][k];
  }
}

// READIN's pre-scan of the radiation file for the soil-thaw day count IDFS (pyhelp/HELP3O.FOR:1348-1375; host twin:
// pyhelp_b200/grid.py idfs_prescan_nodes): REAL*4 sum, in file order, of the December (days 335-365, northern sites) and June
// (days 152-182, southern sites) values of the first <= 100 years of rad[node][year][kDays]; one thread per node and
// hemisphere (the sum is sequential by definition).  out[node][2].
__global__ void idfs_prescan_kernel(const float* __restrict__ rad, int n_nodes, int n_years, int iu13, int32_t* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * n_nodes) return;
  const int node = i >> 1, h = i & 1;
  const int lo = h ? 151 : 334;
  const int nyrs = n_years < 100 ? n_years : 100;
  const float* r = rad + (int64_t)node * n_years * kDays;
  float radtot = 0.0f;
  for (int y = 0; y < nyrs; ++y)
    for (int d = lo; d < lo + 31; ++d) radtot = radtot + r[(int64_t)y * kDays + d];
  float radavg = (nyrs > 1) ? (radtot / (float)nyrs) / 31.0f : radtot / 31.0f;
  if (iu13 != 1) radavg = radavg * 23.89f;
  int idfs = (int)(35.4f - (0.154f * radavg));
  out[i] = idfs < 1 ? 1 : idfs;
}

}  // namespace helpb200
