// help_nodes.cuh -- nearest weather-grid node of every cell.
//
// Reference: HelpManager._generate_d4d7d13_input_files, pyhelp/managers.py:307-313: for every
// cell, dist = calc_dist_from_coord(lat, lon, node_lats, node_lons) (haversine, r = 6373 km,
// pyhelp/utils.py:81-96) and argmin = int(np.argmin(dist)) -- O(cells x nodes) Python/numpy work,
// the first minimum wins.
//
// Here: one thread = one cell; the nodes stream through shared memory as unit vectors.  The
// haversine distance is a monotone function of the dot product of the two unit vectors, so the
// search is 3 FMAs per (cell, node) pair; the reference's own formula is evaluated only for the
// nodes whose dot product is within 1e-13 of the best one, and among those the smallest distance
// -- lowest index on equality -- is returned, which is np.argmin's rule.
#pragma once
#include <stdint.h>
#include "help_math.h"

namespace helpb200 {

struct NodeVec { double x, y, z, lat, lon; };

__global__ void node_vectors_kernel(const double* __restrict__ lat, const double* __restrict__ lon, int32_t n,
                                    NodeVec* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double d2r = 0.017453292519943295;
  const double la = lat[i] * d2r, lo = lon[i] * d2r;
  const double cl = helpmath::det_cos(la);
  NodeVec v;
  v.x = cl * helpmath::det_cos(lo); v.y = cl * helpmath::det_sin(lo); v.z = helpmath::det_sin(la);
  v.lat = la; v.lon = lo;
  out[i] = v;
}

// calc_dist_from_coord (pyhelp/utils.py:81-96), radians in
__device__ __forceinline__ double haversine_km(double lat1, double lon1, double lat2, double lon2) {
  const double dlon = lon2 - lon1, dlat = lat2 - lat1;
  const double s1 = helpmath::det_sin(dlat / 2), s2 = helpmath::det_sin(dlon / 2);
  const double a = s1 * s1 + helpmath::det_cos(lat1) * helpmath::det_cos(lat2) * (s2 * s2);
  const double y = helpmath::hm_sqrt(a), x = helpmath::hm_sqrt(1 - a);
  const double c = 2 * ((x > 0.0) ? helpmath::det_atan(y / x) : 1.5707963267948966);
  return 6373 * c;
}

constexpr int kNodeTile = 1024;

__global__ void __launch_bounds__(256) nearest_node_kernel(const double* __restrict__ clat, const double* __restrict__ clon,
                                                           int64_t n_cells, const NodeVec* __restrict__ nodes, int32_t n_nodes,
                                                           int32_t* __restrict__ out) {
  __shared__ double sx[kNodeTile], sy[kNodeTile], sz[kNodeTile];
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = c < n_cells;
  const double d2r = 0.017453292519943295;
  const double la = live ? clat[c] * d2r : 0.0, lo = live ? clon[c] * d2r : 0.0;
  const double cl = helpmath::det_cos(la);
  const double ux = cl * helpmath::det_cos(lo), uy = cl * helpmath::det_sin(lo), uz = helpmath::det_sin(la);
  // pass 1: largest dot product
  double best = -2.0;
  for (int t0 = 0; t0 < n_nodes; t0 += kNodeTile) {
    const int m = min(kNodeTile, n_nodes - t0);
    __syncthreads();
    for (int k = threadIdx.x; k < m; k += blockDim.x) { const NodeVec v = nodes[t0 + k]; sx[k] = v.x; sy[k] = v.y; sz[k] = v.z; }
    __syncthreads();
    for (int k = 0; k < m; ++k) {
      const double d = fma(ux, sx[k], fma(uy, sy[k], uz * sz[k]));
      best = d > best ? d : best;
    }
  }
  // pass 2: the reference's distance for every node within rounding of the best
  const double cut = best - 1e-13;
  double dmin = 1e300;
  int32_t arg = 0;
  for (int t0 = 0; t0 < n_nodes; t0 += kNodeTile) {
    const int m = min(kNodeTile, n_nodes - t0);
    __syncthreads();
    for (int k = threadIdx.x; k < m; k += blockDim.x) { const NodeVec v = nodes[t0 + k]; sx[k] = v.x; sy[k] = v.y; sz[k] = v.z; }
    __syncthreads();
    for (int k = 0; k < m; ++k) {
      const double d = fma(ux, sx[k], fma(uy, sy[k], uz * sz[k]));
      if (d >= cut) {
        const NodeVec v = nodes[t0 + k];
        const double km = haversine_km(la, lo, v.lat, v.lon);
        if (km < dmin) { dmin = km; arg = t0 + k; }      // strict: the first minimum stays
      }
    }
  }
  if (live) out[c] = arg;
}

}  // namespace helpb200
