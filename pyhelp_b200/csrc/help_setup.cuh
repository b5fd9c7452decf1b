// help_setup.cuh -- per-cell set-up kernel: one thread turns one cell's D10/D11
// numbers into the cell image (help_image.h).
//
// What it computes, with the reference lines it must agree with
// (pyhelp/HELP3O.FOR, "F:"):
//   soil_params()   Brooks-Corey residual/lambda/bubbling pressure, evaporation
//                   coefficient per layer                            F:1307-1340
//   liner_systems() LP/LD/LPQ: which layers close a subprofile       F:2316-2387
//   discretise()    SI->IP conversion, 7 evaporative segments + 1-3 segments per
//                   deeper layer, liner segment typing, FML design case,
//                   thickness-weighted segment properties, SEDEP     F:1406-2311
//   substeps()      sub-daily step count per subprofile              F:2577-2672
//   surface_consts() curve-number retention, CONA/STAGE1, vegetation and
//                   soil-temperature constants                       F:169-185, 2394-2531, 3023-3046
// All REAL*4 arithmetic is kept REAL*4 and in the reference's operation order
// (the file is compiled with -fmad=false) because integer decisions (segment
// counts, NSTEP, design case) hang off it and define the cell's class.
#pragma once
#include "help_image.h"
#include "../../include/help_b200.h"

namespace helpb200 {

struct BatchDev {            // device copy of help_b200_batch
  int n_cells, n_years, max_layers, n_nodes_p, n_nodes_t, n_nodes_r;
  const int32_t* years;
  float *pre, *tmp, *rad;    // converted in place to in / deg F / langley by convert_weather
  const int32_t *iu4, *iu7, *iu13;
  const float* tm_normals;
  const int32_t* idfs;
  const float* cell_f32;
  const int32_t* cell_i32;
  const float* layer_f32;
  const int32_t* layer_i32;
  const double* layer_rc;
};

// ---- weather unit conversion (READCD, F:1138-1152) ---------------------------
// Elementwise over [node][year][day]; HBM-bound, runs once per upload.
__global__ void convert_weather_kernel(float* __restrict__ w, const int32_t* __restrict__ units,
                                       int64_t per_node, int64_t total, int which) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  for (; i < total; i += step) {
    const int node = (int)(i / per_node);
    if (units[node] == 1) continue;              // already inches / deg F / langley
    float v = w[i];
    if (which == 0) v = v / 25.4f;               // PRE(N) / 25.4
    else if (which == 1) v = __fadd_rn(__fmul_rn(v, 1.8f), 32.0f);   // (TMPF*1.8)+32.0
    else v = v * 23.89f;                         // RAD * 23.89
    w[i] = v;
  }
}

// ---- working set of one cell ----------------------------------------------------
struct Layers {   // 1-based like the Fortran; index 0 and LAY+1 read as zero
  float PORO[22], FC[22], WP[22], SW[22], RS[22], XLAMBD[22], BUB[22], THICK[22], SLOPE[22],
      XLENG[22], CON[22], SUBIN[22], RECIR[22], PHOLE[22], DEFEC[22], TRANS[22];
  double RC[22];
  int LAYER[23], LAYR[22], IPQ[22], ISOIL[22];
  int LAY;
};
struct Segs {
  float STHICK[70], UL[70], FCUL[70], WPUL[70], SWUL[70], RCUL[70], RSUL[70], XLAMBU[70], CONUL[70],
      BUBUL[70], SUBINS[70], RCULS[70];
  int LAYSEG1[70], LAYSEG2[70], LSEG[70];
  int NSEGn[7], NSEG;
  int LS1[23], LS2[23];     // LSEGS(K,1..2): first and last segment of layer K (F:1938-2022)
};
struct Topo {
  int LP[8], LD[8], LPQ[8], LCASE[8], NSTEP[8];
};

__device__ inline void soil_params(Layers& L) {                 // F:1307-1340
  for (int J = 1; J <= L.LAY; ++J) {
    if (L.LAYER[J] == 4) { L.PORO[J] = .04f; L.FC[J] = .03f; L.WP[J] = .02f; L.SW[J] = .04f; }
    float wp = L.WP[J], rs;
    if (wp < 0.04f) rs = 0.6f * wp; else rs = 0.014f + 0.253f * wp;
    if (wp < 0.02f) rs = 0.75f * wp;
    L.RS[J] = rs;
    if (L.WP[J] <= 0.0f) L.WP[J] = 0.001f;
    if (L.FC[J] <= L.WP[J]) L.FC[J] = L.WP[J] + 0.001f;
    if (L.PORO[J] <= L.FC[J]) L.PORO[J] = L.FC[J] + 0.001f;
    L.XLAMBD[J] = r4_log((L.FC[J] - rs) / (L.WP[J] - rs)) / r4_log(45.f);
    const float relfc = (L.FC[J] - rs) / (L.PORO[J] - rs);
    float bub = (r4_pow(relfc, 1.f / L.XLAMBD[J])) * 1033.4f / 3.f;
    if (bub > 103.34f) bub = 103.34f;
    L.BUB[J] = bub;
    const float pw = r4_pow(bub / 103.34f, (3.f * L.XLAMBD[J]) + 2.f);
    float con = (float)((double)2.44f + (((double)1485216.f * L.RC[J]) * (double)pw));
    if (relfc < 0.20f) con = 3.30f;
    if (L.RC[J] < (double)0.000005f) con = 3.30f;
    if (con < 3.30f) con = 3.30f;
    if (con > 5.50f) con = 5.50f;
    L.CON[J] = con;
    if (L.LAYER[J] == 4) { L.PORO[J] = .0f; L.FC[J] = .0f; L.WP[J] = .0f; L.SW[J] = .0f; }
  }
}

__device__ inline void liner_systems(Layers& L, Topo& T, int* LRIN) {   // F:2316-2387
  int NL = 0;
  for (int i = 0; i < 8; ++i) { T.LP[i] = 0; T.LD[i] = 0; T.LPQ[i] = 0; }
  for (int I = 1; I <= L.LAY; ++I) {
    LRIN[I] = 0;
    const int t = L.LAYER[I];
    if (t == 3 || t == 4) {
      const int up = L.LAYER[I - 1];
      if (up != 3 && up != 4) NL = NL + 1;
      if (NL >= 1 && NL <= 6) {
        T.LP[NL] = I;
        if (t == 4) {
          if (L.IPQ[I] < 1 || L.IPQ[I] > 6) L.IPQ[I] = 4;
          if (NL <= 5) T.LPQ[NL] = L.IPQ[I];
        }
      }
    }
  }
  for (int I = 1; I <= 5; ++I) {
    const int lp = T.LP[I];
    if (lp >= 2) {
      if (L.LAYER[lp - 1] == 2) T.LD[I] = lp - 1;
      else if (lp > 2 && L.LAYER[lp - 2] == 2) T.LD[I] = lp - 2;
    }
  }
  const int tb = L.LAYER[L.LAY];
  if (tb != 3 && tb != 4) {
    if (NL == 0) T.LP[1] = L.LAY;
    for (int k = 1; k <= 5; ++k)
      if (NL == k && L.LAY > T.LP[k]) T.LP[k + 1] = L.LAY;
  }
  for (int N = 1; N <= 5; ++N) {
    const int ld = T.LD[N];
    if (ld != 0 && L.LAYR[ld] != 0) {
      const int to = L.LAYR[ld];
      if (L.LAYER[to] > 2) { LRIN[to] = 0; L.RECIR[to] = 0.0f; }
      else if (L.RECIR[ld] > 0.0f) LRIN[to] = 1;
    }
  }
}

// One liner / geomembrane layer -> at most one segment (F:1554-1585 and its
// four copies).  Returns false when layer J is an ordinary soil layer.
__device__ inline bool add_liner_segment(const Layers& L, const Topo& T, Segs& S, int J, int n,
                                         int& NS, float liner_thick) {
  const int t = L.LAYER[J];
  if (t == 3) {
    S.STHICK[NS + 1] = liner_thick;
    S.LAYSEG1[NS + 1] = J;
    int kind = 3;
    if (J != T.LP[n]) kind = 5;
    if (L.LAYER[J - 1] == 4) kind = 4;
    S.LSEG[NS + 1] = kind;
    NS = NS + 1;
    return true;
  }
  if (t == 4) {
    if (J == T.LP[n] && L.LAYER[J - 1] != 3) {
      bool below = false;
      if ((J + 1) <= L.LAY) below = (L.RC[J + 1] < L.RC[J - 1]);
      if (below) {
        S.STHICK[NS + 1] = (L.THICK[J + 1] > 18.f) ? 12.f : L.THICK[J + 1];
        S.LSEG[NS + 1] = 6;
        S.LAYSEG1[NS + 1] = J + 1;
      } else {
        S.STHICK[NS + 1] = S.STHICK[NS];
        S.LSEG[NS + 1] = 7;
        S.LAYSEG1[NS + 1] = J - 1;
      }
      NS = NS + 1;
    }
    return true;
  }
  return false;
}
__device__ inline void add_soil_segments(const Layers& L, Segs& S, int J, int& NS, float T) {
  int cnt;
  if (T > 36.f) { S.STHICK[NS + 1] = 12.f; S.STHICK[NS + 2] = T - 12.f - 12.f; S.STHICK[NS + 3] = 12.f; cnt = 3; }
  else if (T > 18.f) { S.STHICK[NS + 1] = 12.f; S.STHICK[NS + 2] = T - 12.f; cnt = 2; }
  else { S.STHICK[NS + 1] = T; cnt = 1; }
  for (int k = 1; k <= cnt; ++k) { S.LAYSEG1[NS + k] = J; S.LAYSEG2[NS + k] = J; S.LSEG[NS + k] = L.LAYER[J]; }
  NS = NS + cnt;
}

struct CellIn {
  float ULAT, ULAI, EDEPTH, WIND, HUM[5], OSNO, FRUNOF, CN2, TFSOIL;
  int IU11, IPL, IHV, IU10, IPRE;
};

// Returns false for a design this engine refuses (flagged, outputs NaN).
__device__ inline bool discretise(Layers& L, Topo& T, Segs& S, CellIn& C, float& SUBINF, float& SEDEP,
                                  float& OLDSNO, double* DSUBINd, int segcap) {   // F:1406-2311
  const float WF[8] = {0, 0.085f, 0.334f, 0.252f, 0.151f, 0.091f, 0.054f, 0.033f};
  for (int n = 0; n < 7; ++n) S.NSEGn[n] = 0;
  for (int n = 0; n < 8; ++n) T.LCASE[n] = 0;
  float CORECT = 1.0f;
  if (L.ISOIL[1] <= 34 && L.ISOIL[1] >= 1) {
    const float u = C.ULAI, u2 = u * u, u3 = u * u2, u4 = u2 * u2, u5 = u * u4;
    CORECT = 1.0f + 0.5966f * u + 0.132659f * u2 + 0.1123454f * u3 - 0.04777627f * u4 + 0.004325035f * u5;
    if (CORECT > 5.0f) CORECT = 5.0f;
  }
  for (int J = 0; J < 70; ++J) { S.STHICK[J] = 0.0f; S.LAYSEG1[J] = 0; S.LAYSEG2[J] = 0; S.LSEG[J] = 0; }
  SUBINF = 0.0f;
  for (int K = 1; K <= L.LAY; ++K) {
    L.SUBIN[K] = L.SUBIN[K] / 365.0f;
    SUBINF = SUBINF + L.SUBIN[K];
    if (C.IPRE == 0 || C.IPRE == 2) L.SW[K] = L.FC[K];
    if (L.LAYER[K] == 4) L.SW[K] = 0.0f;
    if (L.LAYER[K] == 3) L.SW[K] = L.PORO[K];
  }
  C.FRUNOF = C.FRUNOF / 100.0f;
  if (C.IU10 == 2) C.OSNO = C.OSNO / 25.4f;
  OLDSNO = C.OSNO;
  if (C.IU10 == 2) {
    SUBINF = SUBINF / 25.4f;
    for (int J = 1; J <= L.LAY; ++J) {
      L.DEFEC[J] = L.DEFEC[J] / 2.471f; L.PHOLE[J] = L.PHOLE[J] / 2.471f;
      L.THICK[J] = L.THICK[J] / 2.54f;  L.XLENG[J] = L.XLENG[J] * 3.281f;
      L.SUBIN[J] = L.SUBIN[J] / 25.4f;
    }
  }
  float EDEPTH = C.EDEPTH;
  float above = 0.0f;                                   // soil above the first liner
  for (int J = 1; J <= T.LP[1]; ++J) {
    if (L.LAYER[J] == 3 || L.LAYER[J] == 4) break;
    above = above + L.THICK[J];
  }
  if (above < EDEPTH) EDEPTH = above;
  if (!(EDEPTH > 0.0f)) return false;
  // the seven evaporative segments: EDEPTH * {1,5,6,6,6,6,rest}/36 (F:1503-1510)
  S.STHICK[1] = EDEPTH / 36.0f;
  S.STHICK[2] = 5.0f * EDEPTH / 36.0f;
  float tt = S.STHICK[1] + S.STHICK[2];
  for (int J = 3; J <= 6; ++J) { S.STHICK[J] = EDEPTH / 6.0f; tt = tt + S.STHICK[J]; }
  S.STHICK[7] = EDEPTH - tt;
  {
    int J = 1;
    float t1 = L.THICK[1], t2 = 0.0f;
    for (int K = 1; K <= 7; ++K) {
      t2 = t2 + S.STHICK[K];
      S.LAYSEG1[K] = J;
      while (t2 > t1 && J < 20) { J = J + 1; t1 = t1 + L.THICK[J]; }
      if (t2 <= t1) S.LAYSEG2[K] = J;
    }
  }
  int NS = 7;
  {
    float t1 = 0.0f, t2 = EDEPTH + 0.0002f;
    for (int J = 1; J <= T.LP[1]; ++J) {
      S.LAYSEG2[NS + 1] = J;
      t1 = t1 + L.THICK[J];
      if (t1 > (EDEPTH + 0.0001f)) {
        const float rest = t1 - t2;
        if (!add_liner_segment(L, T, S, J, 1, NS, rest)) add_soil_segments(L, S, J, NS, rest);
        t2 = t1;
      }
    }
  }
  S.NSEGn[1] = NS;
  int used = NS;
  bool done = (L.LAY <= T.LP[1]);
  for (int n = 2; n <= 5 && !done; ++n) {
    for (int J = T.LP[n - 1] + 1; J <= T.LP[n]; ++J) {
      S.LAYSEG2[NS + 1] = J;
      if (!add_liner_segment(L, T, S, J, n, NS, L.THICK[J])) add_soil_segments(L, S, J, NS, L.THICK[J]);
    }
    S.NSEGn[n] = NS - used;
    used = NS;
    if (L.LAY <= T.LP[n]) done = true;
  }
  if (!done) {
    for (int J = T.LP[5] + 1; J <= T.LP[6]; ++J) add_soil_segments(L, S, J, NS, L.THICK[J]);
  }
  // 7 + 3 segments per layer <= 67 always fits S; the image may be narrower
  if (NS > segcap) return false;
  S.NSEGn[6] = NS - S.NSEGn[1] - S.NSEGn[2] - S.NSEGn[3] - S.NSEGn[4] - S.NSEGn[5];
  S.NSEG = NS;
  // thickness-weighted properties of the 7 evaporative segments (F:1937-1996);
  // root channels raise K of the top layer for segments 1-4 (F:1943,1955)
  {
    float depseg = 0.0f, depths = 0.0f, depthl = L.THICK[1];
    int K = 1;
    for (int k = 0; k < 23; ++k) { S.LS1[k] = 0; S.LS2[k] = 0; }
    S.LS1[1] = 1;
    const double rc1 = L.RC[1];
    L.RC[1] = L.RC[1] * (double)CORECT;
    for (int J = 1; J <= 7; ++J) {
      float ul = 0.f, wp = 0.f, fc = 0.f, con = 0.f, rs = 0.f, lam = 0.f, sw = 0.f, rcs = 0.f, bub = 0.f, sub = 0.f;
      if (J == 5) L.RC[1] = rc1;
      if (depseg >= depthl) { K = K + 1; S.LS1[K] = J; depthl = depthl + L.THICK[K]; }
      depths = depths + S.STHICK[J];
      for (;;) {
        S.LSEG[J] = L.LAYER[K];
        const bool last = (depths <= depthl) || (K >= 20);
        const float dseg = last ? (depths - depseg) : (depthl - depseg);
        if (dseg > 0.0001f) S.LS2[K] = J;
        ul = L.PORO[K] * dseg + ul;
        fc = L.FC[K] * dseg + fc;
        wp = L.WP[K] * dseg + wp;
        con = L.CON[K] * dseg + con;
        bub = L.BUB[K] * dseg + bub;
        rcs = (float)(L.RC[K] * (double)dseg + (double)rcs);
        rs = L.RS[K] * dseg + rs;
        lam = L.XLAMBD[K] * dseg + lam;
        sub = L.SUBIN[K] * dseg / L.THICK[K] + sub;
        sw = L.SW[K] * dseg + sw;
        if (last) { depseg = depths; break; }
        depseg = depthl;
        K = K + 1;
        S.LS1[K] = J;
        depthl = depthl + L.THICK[K];
      }
      S.UL[J] = ul; S.FCUL[J] = fc; S.WPUL[J] = wp; S.CONUL[J] = con; S.BUBUL[J] = bub; S.RCULS[J] = rcs;
      S.RSUL[J] = rs; S.XLAMBU[J] = lam; S.SUBINS[J] = sub; S.SWUL[J] = sw;
    }
    L.RC[1] = rc1;
  }
  {
    int K = S.LAYSEG2[7];
    for (int J = 8; J <= NS; ++J) {
      if (S.LAYSEG2[J] != K) { K = S.LAYSEG2[J]; S.LS1[K] = J; }
      S.LS2[K] = J;
    }
  }
  for (int J = 8; J <= NS; ++J) {                                      // F:1997-2022
    const int a = S.LAYSEG1[J], b = S.LAYSEG2[J];
    const float th = S.STHICK[J];
    S.UL[J] = L.PORO[a] * th; S.FCUL[J] = L.FC[a] * th; S.WPUL[J] = L.WP[a] * th;
    S.CONUL[J] = L.CON[a] * th; S.BUBUL[J] = L.BUB[a] * th;
    S.RCULS[J] = (float)(L.RC[a] * (double)th);
    S.RSUL[J] = L.RS[a] * th; S.XLAMBU[J] = L.XLAMBD[a] * th;
    if (S.LSEG[J] < 3) S.SUBINS[J] = L.SUBIN[b] * th / L.THICK[b];
    else S.SUBINS[J] = L.SUBIN[b];
    if (S.LSEG[J] == 4) S.SUBINS[J] = S.SUBINS[J] + L.SUBIN[b - 1];
    S.SWUL[J] = L.SW[a] * th;
  }
  for (int J = 1; J <= NS; ++J) {                                      // F:2027-2042
    S.RCUL[J] = (S.RCULS[J] * 24.f * 1417.3f) / S.STHICK[J];
    S.XLAMBU[J] = S.XLAMBU[J] / S.STHICK[J];
    S.BUBUL[J] = S.BUBUL[J] / S.STHICK[J];
    S.CONUL[J] = S.CONUL[J] / S.STHICK[J];
    DSUBINd[J] = S.SUBINS[J];
  }
  // subsurface-inflow placement at subprofile bottoms and the geomembrane design case (F:2057-2294)
  int NSEGE = 0;
  for (int N = 1; N <= 5; ++N) {
    if (N == 1) NSEGE = S.NSEGn[1];
    else { if (!(T.LP[N] > 0)) continue; NSEGE = NSEGE + S.NSEGn[N]; }
    const int lp = T.LP[N];
    if (lp == L.LAY) {
      if (NSEGE >= 2) DSUBINd[NSEGE - 1] = DSUBINd[NSEGE - 1] + L.SUBIN[lp];
      DSUBINd[NSEGE] = 0.0f;
      if ((lp - 1) == 3 || (lp - 1) == 4) { if (NSEGE >= 2) DSUBINd[NSEGE - 1] = DSUBINd[NSEGE - 1] + L.SUBIN[lp - 1]; }
    } else if (S.LSEG[NSEGE] == 5) {
      DSUBINd[NSEGE - 1] = DSUBINd[NSEGE - 1] + L.SUBIN[lp - 1];
      DSUBINd[NSEGE] = L.SUBIN[lp];
    }
    const int kind = S.LSEG[NSEGE];
    if (kind > 3) {
      const double k = L.RC[S.LAYSEG1[NSEGE]];
      int lc = 0;
      if (kind == 4) lc = 3;
      if (kind == 5) lc = 4;
      if (kind == 6) lc = (k > (double)0.0999f) ? 1 : ((k <= (double)0.0000999f) ? 3 : 2);
      if (kind == 7) {
        const bool hi = (N == 1) ? (k > (double)0.0999f) : (k > (double)0.099f);
        lc = hi ? 1 : ((k <= (double)0.0000999f) ? 4 : 2);
      }
      if (T.LPQ[N] == 6) { if (lc == 2 || lc == 3) lc = 5; else if (lc == 4) lc = 6; }
      T.LCASE[N] = lc;
    }
  }
  float rces = 0.0f;                                                   // F:2296-2309
  for (int J = 1; J <= 7; ++J) rces = rces + WF[J] * S.RCUL[J];
  rces = rces / 34016.f;
  rces = -1.0f * r4_log10(rces);
  float sedmx = 4.606745f * r4_pow(1.595155f, rces);
  if (sedmx < 18.f) sedmx = 18.f;
  if (sedmx > 48.f) sedmx = 48.f;
  SEDEP = (EDEPTH > sedmx) ? sedmx : EDEPTH;
  C.EDEPTH = EDEPTH;
  return true;
}

__device__ inline int trunc_to_int(float x) {
  if (!(x < 1.0e8f)) return 100000000;
  if (x < -1.0e8f) return -100000000;
  return (int)x;
}

__device__ inline void substeps(Layers& L, Topo& T) {                   // F:2577-2672
  for (int N = 1; N <= 6; ++N) T.NSTEP[N] = 4;
  float RTOP = (float)L.RC[1];
  for (int I = 2; I <= L.LAY; ++I) {
    if (L.LAYER[I] > 2) break;
    if (L.RC[I] < (double)RTOP) RTOP = (float)L.RC[I];
  }
  RTOP = 2.0f * RTOP;
  if (RTOP > 0.0005f) RTOP = 0.0005f;
  auto steps_for = [&](int ld, float rtop) {
    if (L.XLENG[ld] == 0.0f) L.XLENG[ld] = 1.f;
    if (L.SLOPE[ld] == 0.0f) L.SLOPE[ld] = 0.00001f;
    const float rdr = (float)(L.RC[ld] * (double)L.THICK[ld] * (double)L.SLOPE[ld] / (double)(1200.f * L.XLENG[ld]));
    const float cap = (L.PORO[ld] - L.FC[ld]) * L.THICK[ld] * 2.54f;
    const float dt = (rdr > rtop) ? cap / (rtop * 86400.0f) : cap / (rdr * 86400.0f);
    return (trunc_to_int(1.0f / dt) + 1) * 2;
  };
  if (T.LD[1] > 1) {
    T.NSTEP[1] = steps_for(T.LD[1], RTOP);
    if (T.NSTEP[1] < 4) T.NSTEP[1] = 4;
  }
  float RTOP2 = 0.0f;
  for (int N = 2; N <= 5; ++N) {
    if (T.LP[N] > 1) {
      const int lpp = T.LP[N - 1], LPP1 = lpp + 1, LPM1 = lpp - 1;
      const int tm1 = (LPM1 >= 0) ? L.LAYER[LPM1] : 0;
      if (tm1 == 3) RTOP2 = (float)L.RC[LPM1];
      if (L.LAYER[lpp] == 3) RTOP2 = (float)L.RC[T.LP[1]];
      if (tm1 == 4) RTOP2 = RTOP2 * 0.001f;
      if (L.LAYER[lpp] == 4) {
        if (tm1 == 3) RTOP2 = RTOP2 * 0.004f;
        else {
          const double k = (L.RC[LPM1] < L.RC[LPP1]) ? L.RC[LPM1] : L.RC[LPP1];
          RTOP2 = (float)((double)0.004f * k);
          if (RTOP2 > 0.000005f) RTOP2 = 0.000005f;
        }
      }
      RTOP2 = RTOP2 * 2.0f;
      if (RTOP2 < RTOP) RTOP = RTOP2;
      const int hi = (T.LD[N] > 1) ? T.LD[N] : T.LP[N];
      for (int I = lpp + 1; I <= hi; ++I) if (L.RC[I] < (double)RTOP) RTOP = (float)L.RC[I];
      if (T.LD[N] > 1) T.NSTEP[N] = steps_for(T.LD[N], RTOP);
    }
  }
  for (int N = 1; N <= 6; ++N) {
    if (T.NSTEP[N] < 4) T.NSTEP[N] = 4;
    else if (T.NSTEP[N] > 48 && N > 1) T.NSTEP[N] = 48;
    else if (T.NSTEP[N] > 288) T.NSTEP[N] = 288;
  }
}

// harmonic fit of the monthly normals and potential heat units (F:2484-2531)
__device__ inline float heat_units(int M, int K, float A1, float A2, float A3) {
  float acc = 0.f;
  for (int Lq = M; Lq <= K; ++Lq) {
    const float ai = (float)Lq - 0.5f;
    const float ang = 6.283185f * ai / 366.f;
    const float ta = (A1 + A2 * r4_cos(ang) + A3 * r4_sin(ang)) - 5.f;
    if (ta > 0.0f) acc = acc + ta;
  }
  return acc;
}

__global__ void __launch_bounds__(64)
help_setup_kernel(BatchDev B, Image img, const int32_t* __restrict__ perm, uint32_t* __restrict__ flags,
                  int32_t* __restrict__ topo, uint64_t* __restrict__ sortkey, int32_t* __restrict__ n_narrow) {
  helpmath::hm_load_tables();
  const int64_t slot = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // position in the (sorted) image
  if (slot >= B.n_cells) return;
  const int64_t c = perm ? perm[slot] : slot;                            // caller's cell index
  const int64_t n = B.n_cells;
  Layers L; Segs S; Topo T; CellIn C;
  int LRIN[22];
  double DSUBINd[70];
  // ---- gather this cell's records (READIN, F:1221-1300) ------------------------------
  C.ULAT = B.cell_f32[HB_ULAT * n + c]; C.ULAI = B.cell_f32[HB_ULAI * n + c];
  C.EDEPTH = B.cell_f32[HB_EDEPTH * n + c]; C.WIND = B.cell_f32[HB_WIND * n + c];
  for (int q = 0; q < 4; ++q) C.HUM[q + 1] = B.cell_f32[(HB_HUM1 + q) * n + c];
  C.OSNO = B.cell_f32[HB_OSNO * n + c]; C.FRUNOF = B.cell_f32[HB_FRUNOF * n + c];
  C.CN2 = B.cell_f32[HB_CN2 * n + c]; C.TFSOIL = B.cell_f32[HB_TFSOIL * n + c];
  C.IU11 = B.cell_i32[HB_IU11 * n + c]; C.IPL = B.cell_i32[HB_IPL * n + c]; C.IHV = B.cell_i32[HB_IHV * n + c];
  C.IU10 = B.cell_i32[HB_IU10 * n + c]; C.IPRE = B.cell_i32[HB_IPRE * n + c];
  const int node_p = B.cell_i32[HB_NODE_P * n + c], node_t = B.cell_i32[HB_NODE_T * n + c],
            node_r = B.cell_i32[HB_NODE_R * n + c];
  int nlay = B.cell_i32[HB_NLAY * n + c];
  uint32_t flag = 0;
  if (nlay < 1 || nlay > B.max_layers || nlay > kMaxLay) { flag |= kFlagBadCell; nlay = 0; }
  if (node_p < 0 || node_p >= B.n_nodes_p || node_t < 0 || node_t >= B.n_nodes_t || node_r < 0 ||
      node_r >= B.n_nodes_r) flag |= kFlagBadCell;
  L.LAY = nlay;
  for (int J = 0; J < 22; ++J) {
    L.PORO[J] = L.FC[J] = L.WP[J] = L.SW[J] = L.RS[J] = L.XLAMBD[J] = L.BUB[J] = L.THICK[J] = L.SLOPE[J] = 0.f;
    L.XLENG[J] = L.CON[J] = L.SUBIN[J] = L.RECIR[J] = L.PHOLE[J] = L.DEFEC[J] = L.TRANS[J] = 0.f;
    L.RC[J] = 0.0; L.LAYER[J] = 0; L.LAYR[J] = 0; L.IPQ[J] = 0; L.ISOIL[J] = 0;
  }
  L.LAYER[22] = 0;
  const int64_t ml = B.max_layers;
  for (int J = 1; J <= nlay; ++J) {
    const int64_t o = (int64_t)(J - 1) * n + c;
    L.THICK[J] = B.layer_f32[(HL_THICK * ml) * n + o]; L.PORO[J] = B.layer_f32[(HL_PORO * ml) * n + o];
    L.FC[J] = B.layer_f32[(HL_FC * ml) * n + o];       L.WP[J] = B.layer_f32[(HL_WP * ml) * n + o];
    L.SW[J] = B.layer_f32[(HL_SW * ml) * n + o];       L.XLENG[J] = B.layer_f32[(HL_XLENG * ml) * n + o];
    L.SLOPE[J] = B.layer_f32[(HL_SLOPE * ml) * n + o]; L.RECIR[J] = B.layer_f32[(HL_RECIR * ml) * n + o];
    L.SUBIN[J] = B.layer_f32[(HL_SUBIN * ml) * n + o]; L.PHOLE[J] = B.layer_f32[(HL_PHOLE * ml) * n + o];
    L.DEFEC[J] = B.layer_f32[(HL_DEFEC * ml) * n + o]; L.TRANS[J] = B.layer_f32[(HL_TRANS * ml) * n + o];
    L.LAYER[J] = B.layer_i32[(HL_TYPE * ml) * n + o];  L.ISOIL[J] = B.layer_i32[(HL_ISOIL * ml) * n + o];
    L.LAYR[J] = B.layer_i32[(HL_LAYR * ml) * n + o];   L.IPQ[J] = B.layer_i32[(HL_IPQ * ml) * n + o];
    L.RC[J] = B.layer_rc[o];
  }
  // unit conversions of the D11 record (F:1227-1233)
  if (C.IU11 != 1) { C.EDEPTH = C.EDEPTH / 2.54f; C.WIND = C.WIND / 1.609f; }
  float SUBINF = 0.f, SEDEP = 0.f, OLDSNO = 0.f;
  bool ok = !(flag & kFlagBadCell);
  if (ok) {
    soil_params(L);
    liner_systems(L, T, LRIN);
    ok = discretise(L, T, S, C, SUBINF, SEDEP, OLDSNO, DSUBINd, img.segcap);
    if (!ok) flag |= kFlagBadCell;
  }
  if (!ok) {
    img.I(I_NSEG, slot) = 0; img.I(I_FLAGS, slot) = (int32_t)flag; img.I(I_NPROF, slot) = 0;
    if (flags) flags[c] = flag;
    if (sortkey) sortkey[slot] = ~0ull;
    if (topo) for (int k = 0; k < 16; ++k) topo[(int64_t)k * n + c] = 0;
    return;
  }
  substeps(L, T);
  for (int K = 1; K <= L.LAY; ++K) if (LRIN[K] == 1) flag |= kFlagRecirc;
  // ---- curve-number retention (F:169-185) ------------------------------------------------
  float SMXN, SMXF;
  {
    const float cn = C.CN2, c2 = cn * cn, c3 = cn * c2, c4 = cn * c3;
    const float cn1 = 0.3750701f * cn + 2.756779E-03f * c2 - 1.638951E-05f * c3 + 5.142644E-07f * c4;
    SMXN = (1000.f / cn1) - 10.0f;
    const float f = (cn >= 80.f) ? 98.0f : 95.0f;
    const float f2 = f * f, f3 = f * f2, f4 = f * f3;
    const float cn11 = 0.3750701f * f + 2.756779E-03f * f2 - 1.638951E-05f * f3 + 5.142644E-07f * f4;
    SMXF = (1000.f / cn11) - 10.f;
  }
  // ---- evaporation coefficient (F:3035-3044) ---------------------------------------------
  const float WF[8] = {0, 0.085f, 0.334f, 0.252f, 0.151f, 0.091f, 0.054f, 0.033f};
  float CONA = 0.0f;
  for (int J = 1; J <= 7; ++J) CONA = CONA + (WF[J] * S.CONUL[J]);
  if (CONA < 3.3f) CONA = 3.3f;
  if (CONA > 5.1f) CONA = 5.1f;
  const float STAGE1 = (23.f * r4_pow(CONA - 3.0f, 0.42f)) / 25.4f;
  // ---- vegetation / soil temperature constants (INITIA, F:2394-2479) ------------------------
  float TM[13];
  for (int I = 1; I <= 12; ++I) {
    float t = B.tm_normals[(int64_t)node_t * 12 + (I - 1)];
    if (B.iu7[node_t] == 1) t = (t - 32.0f) / 1.8f;                     // F:1227-1231
    TM[I] = t;
  }
  float AVT = 0.f, TBB = 0.f, TS = 200.f;
  for (int I = 1; I <= 12; ++I) { if (TM[I] > TBB) TBB = TM[I]; if (TM[I] < TS) TS = TM[I]; AVT = AVT + TM[I]; }
  AVT = AVT / 12.f;
  const float AMP = (TBB - TS) / 2.0f;
  float A1, A2, A3;
  {                                                                     // DATFIT (F:2501-2524)
    const float PI = 3.14159f, AM = 12.f;
    float sumd = 0.f, suma = 0.f, sumb = 0.f;
    for (int I = 1; I <= 12; ++I) sumd = sumd + TM[I];
    A1 = sumd / AM;
    for (int I = 1; I <= 12; ++I) {
      const float ti = (float)I - 0.5f;
      const float th = 2.0f * PI * 1.0f * ti;
      suma = suma + TM[I] * r4_cos(th / AM);
      sumb = sumb + TM[I] * r4_sin(th / AM);
    }
    A2 = 2.0f / AM * suma;
    A3 = 2.0f / AM * sumb;
  }
  float PHU, PHUX = 0.f;
  if (C.IPL >= C.IHV) { PHU = heat_units(C.IPL, 365, A1, A2, A3); PHUX = PHU; PHU = PHU + heat_units(1, C.IHV, A1, A2, A3); }
  else PHU = heat_units(C.IPL, C.IHV, A1, A2, A3);
  if (PHU < 30.f) PHU = 30.f;
  const float XLAI0 = (PHUX / PHU) * C.ULAI;
  const float GS0 = PHUX / PHU;
  const float RSD0 = 8000.f * C.ULAI / 5.f;
  const float DM0 = XLAI0 * 100.f;
  float ABD = 0.f, SWS = 0.f, Z7 = 0.f, FCS7 = 0.f, ST7 = 0.f;
  {
    float z = 0.f;
    for (int I = 1; I <= 7; ++I) {
      z = z + S.STHICK[I] * 25.4f;
      ABD = ABD + S.UL[I] * 25.4f * 10.f;
      const float fcs = (S.FCUL[I] - S.WPUL[I]) * 25.4f;
      float st = (S.SWUL[I] - S.WPUL[I]) * 25.4f;
      SWS = SWS + st;
      st = st / 25.4f;
      if (I == 7) { Z7 = z; FCS7 = fcs; ST7 = st; }
    }
  }
  ABD = ABD / (10.f * Z7);
  float Fq = ABD / (ABD + 686.f * r4_exp(-5.63f * ABD));
  const float DP = 1000.f + 2500.f * Fq;
  const float WW = .356f - .144f * ABD;
  const float Bq = r4_log(500.f / DP);
  const float WC = SWS / (WW * Z7);
  const float q = (1.f - WC) / (1.f + WC);
  Fq = r4_exp(Bq * (q * q));
  const float DD = Fq * DP;
  // ---- write the image --------------------------------------------------------------------
  img.F(F_EDEPTH, slot) = C.EDEPTH; img.F(F_SEDEP, slot) = SEDEP; img.F(F_SMXN, slot) = SMXN;
  img.F(F_SMXF, slot) = SMXF; img.F(F_FRUNOF, slot) = C.FRUNOF; img.F(F_CONA, slot) = CONA;
  img.F(F_STAGE1, slot) = STAGE1; img.F(F_ULAT, slot) = C.ULAT; img.F(F_ULAI, slot) = C.ULAI;
  img.F(F_WIND, slot) = C.WIND;
  for (int qd = 0; qd < 4; ++qd) img.F(F_RH1 + qd, slot) = C.HUM[qd + 1];
  img.F(F_PHU, slot) = PHU; img.F(F_AVT, slot) = AVT; img.F(F_AMP, slot) = AMP; img.F(F_DD, slot) = DD;
  img.F(F_Z7, slot) = Z7; img.F(F_FCS7, slot) = FCS7; img.F(F_ST7, slot) = ST7;
  img.F(F_TS0, slot) = TS; img.F(F_XLAI0, slot) = XLAI0; img.F(F_GS0, slot) = GS0; img.F(F_RSD0, slot) = RSD0;
  img.F(F_DM0, slot) = DM0; img.F(F_OSNO, slot) = C.OSNO; img.F(F_TFSOIL, slot) = C.TFSOIL;
  img.F(F_SUBINF, slot) = SUBINF;
  int nprof = 0, e = 0;
  for (int N = 1; N <= 6; ++N) {
    const int ns = S.NSEGn[N];
    const bool active = (N == 1) || (ns > 0);
    int ibar = 0, lat = 0, NSEGE = e + ns, NSEGL = NSEGE;
    float drculs = 0.f, dthics = 0.f, slope = 0.f, xleng = 0.f, thklin = 0.f, phole = 0.f, defec = 0.f,
          trans = 0.f, recir = 0.f, sublin = 0.f;
    double rclin = 0.0, rcs = 0.0;
    int ltype = 0;
    if (active && ns > 0) {
      nprof = N;
      const int layb = (N == 1) ? 1 : T.LP[N - 1] + 1, laye = T.LP[N];
      if (S.LSEG[NSEGE] == 3) ibar = 1;                                // DRAIN, F:4174-4181
      if (S.LSEG[NSEGE] >= 4) ibar = 2;
      if (layb == laye) ibar = 0;
      if (N == 6) ibar = 0;
      if (ibar >= 1) {
        NSEGL = NSEGE - 1;
        if (N <= 5 && T.LD[N] > 0) lat = 1;
        if (lat == 1) { const int layd = S.LAYSEG2[NSEGL]; slope = L.SLOPE[layd]; xleng = L.XLENG[layd]; }
      }
      drculs = S.RCUL[NSEGE]; dthics = S.STHICK[NSEGE]; sublin = (float)DSUBINd[NSEGE];
      if (S.NSEG == NSEGE) {                                            // F:4201-4210
        const int lp = T.LP[N];
        if (ibar == 1) { if (L.SUBIN[lp] > 0.0f) ibar += 3; }
        else if (ibar == 2) { if (L.SUBIN[lp] > 0.0f || L.SUBIN[lp - 1] > 0.0f) ibar += 3; }
        else if (DSUBINd[NSEGE] > 0.0f) ibar += 3;
      }
      if (ibar == 2 || ibar == 5) {                                      // FML inputs, F:4902-4905
        const int lp = T.LP[N];
        int liner = 0;
        if (L.LAYER[lp] == 4) liner = lp;
        if (lp >= 1 && L.LAYER[lp - 1] == 4) liner = lp - 1;
        thklin = L.THICK[liner]; phole = L.PHOLE[liner]; defec = L.DEFEC[liner]; trans = L.TRANS[liner];
        rclin = L.RC[liner]; rcs = L.RC[S.LAYSEG1[NSEGE]]; ltype = L.LAYER[lp];
      }
      if (N <= 5 && T.LD[N] > 0) recir = L.RECIR[T.LD[N]];
    }
    img.PF(P_DRCULS, N, slot) = drculs; img.PF(P_DTHICS, N, slot) = dthics; img.PF(P_SLOPE, N, slot) = slope;
    img.PF(P_XLENG, N, slot) = xleng; img.PF(P_THKLIN, N, slot) = thklin; img.PF(P_PHOLE, N, slot) = phole;
    img.PF(P_DEFEC, N, slot) = defec; img.PF(P_TRANS, N, slot) = trans; img.PF(P_RECIR, N, slot) = recir;
    img.PF(P_SUBLIN, N, slot) = sublin;
    img.PD(D_RCLIN, N, slot) = rclin; img.PD(D_RCS, N, slot) = rcs;
    img.PI(Q_NSEGN, N, slot) = ns; img.PI(Q_NSTEP, N, slot) = T.NSTEP[N]; img.PI(Q_IBAR, N, slot) = ibar;
    img.PI(Q_LAT, N, slot) = lat; img.PI(Q_LCASE, N, slot) = (N <= 5) ? T.LCASE[N] : 0;
    img.PI(Q_LPQ, N, slot) = (N <= 5) ? T.LPQ[N] : 0; img.PI(Q_LTYPE, N, slot) = ltype;
    img.PI(Q_LD, N, slot) = (N <= 5) ? T.LD[N] : 0;
    // per-segment "above the drain layer" flag (F:4494,4499)
    for (int J = e + 1; J <= e + ns; ++J) {
      const int ld = (N <= 5) ? T.LD[N] : 0;
      img.SI(J, slot) = (S.LAYSEG2[J] < ld) ? 1 : 0;
    }
    e += ns;
  }
  for (int J = 1; J <= img.segcap; ++J) {
    const bool in = (J <= S.NSEG);
    img.SF(S_THICK, J, slot) = in ? S.STHICK[J] : 0.f; img.SF(S_UL, J, slot) = in ? S.UL[J] : 0.f;
    img.SF(S_FC, J, slot) = in ? S.FCUL[J] : 0.f;      img.SF(S_WP, J, slot) = in ? S.WPUL[J] : 0.f;
    img.SF(S_RS, J, slot) = in ? S.RSUL[J] : 0.f;      img.SF(S_LAM, J, slot) = in ? S.XLAMBU[J] : 1.f;
    img.SF(S_BUB, J, slot) = in ? S.BUBUL[J] : 0.f;    img.SF(S_RC, J, slot) = in ? S.RCUL[J] : 0.f;
    img.SF(S_SUBIN, J, slot) = in ? (float)DSUBINd[J] : 0.f;  img.SF(S_SW0, J, slot) = in ? S.SWUL[J] : 0.f;
    img.SD(SD_SUBIN, J, slot) = in ? DSUBINd[J] : 0.0;
    if (!in) img.SI(J, slot) = 0;
  }
  // ---- routing-loop constants per segment (help_image.h ImgSD; F:4329, 4420, 4465-4466) -------------
  {
    int N = 1, eseg = S.NSEGn[1];
    for (int J = 1; J <= img.segcap; ++J) {
      while (N < 6 && J > eseg) { ++N; eseg += S.NSEGn[N]; }
      const bool in = (J <= S.NSEG);
      double C = 0.0, nil = 0.0, bc = 0.0, lowwp = 0.0, lowfc = 0.0;
      float relmin = 0.f;
      if (in) {
        const double DT = 1.0 / (double)T.NSTEP[N];
        const double ul = (double)S.UL[J], rs = (double)S.RSUL[J], lam = (double)S.XLAMBU[J];
        const double wp = (double)S.WPUL[J], fc = (double)S.FCUL[J];
        const double span = ul - rs;
        const double kdt = (double)S.RCUL[J] * DT;
        C = 3.0 + (2.0 / lam);
        nil = -1.0 / lam;
        relmin = (float)((wp - rs) / span);
        bc = kdt * r8_pow(1.0 / span, C);
        lowwp = kdt * r8_pow((wp - rs) / span, C);
        lowfc = kdt * r8_pow((fc - rs) / span, C);
      }
      img.SF(S_RELMIN, J, slot) = relmin;
      img.SD(SD_C, J, slot) = C; img.SD(SD_NIL, J, slot) = nil; img.SD(SD_BC, J, slot) = bc;
      img.SD(SD_LOWWP, J, slot) = lowwp; img.SD(SD_LOWFC, J, slot) = lowfc;
    }
    for (int J = 1; J < img.segcap; ++J) img.SI(J, slot) = (img.SI(J, slot) & 1) | ((img.SI(J + 1, slot) & 1) << 1);
  }
  // ---- routing records (help_image.h RecF / RecD) ---------------------------------------------------
  {
    char* p = img.rec_p(slot);
    for (int J = 1; J <= img.segcap; ++J, p += kRecBytes) {
      const bool in = (J <= S.NSEG);
      const float ul = img.SF(S_UL, J, slot), rs = img.SF(S_RS, J, slot);
      const double span = (double)ul - (double)rs;
      int N = 1, eseg = S.NSEGn[1];
      while (N < 6 && J > eseg) { ++N; eseg += S.NSEGn[N]; }
      const double DT = 1.0 / (double)T.NSTEP[N];
      *(float4*)(p + RV_A * 512) = make_float4(ul, rs, img.SF(S_WP, J, slot), img.SF(S_FC, J, slot));
      *(float4*)(p + RV_B * 512) = make_float4(img.SF(S_LAM, J, slot), img.SF(S_BUB, J, slot), img.SF(S_RELMIN, J, slot),
                                               __int_as_float(img.SI(J, slot)));
      *(double2*)(p + RV_C * 512) = make_double2(in ? 1.0 / span : 0.0, in ? (double)img.SF(S_RC, J, slot) * DT : 0.0);
      *(double2*)(p + RV_D * 512) = make_double2(img.SD(SD_C, J, slot), img.SD(SD_NIL, J, slot));
      *(double2*)(p + RV_E * 512) = make_double2(img.SD(SD_BC, J, slot), img.SD(SD_LOWWP, J, slot));
      RecTail t; t.lowfc = img.SD(SD_LOWFC, J, slot); t.thick = img.SF(S_THICK, J, slot); t.subin = img.SF(S_SUBIN, J, slot);
      *(RecTail*)(p + RV_F * 512) = t;
    }
    int e2 = 0;
    for (int N = 1; N <= 6; ++N) {
      int any = 0, flg = 0;
      for (int J = e2 + 1; J <= e2 + S.NSEGn[N]; ++J) {
        if (img.SF(S_SUBIN, J, slot) != 0.0f) any = 1;
        flg |= img.SI(J, slot);
      }
      const int plain = (img.PI(Q_IBAR, N, slot) == 0 && !any && flg == 0 && img.PF(P_SUBLIN, N, slot) == 0.0f) ? 2 : 0;
      img.PI(Q_SUBIN, N, slot) = any | plain;
      e2 += S.NSEGn[N];
    }
  }
  // ---- recirculation plan (RECIRC, F:5311-5440): which segments receive the recirculated share
  // of which layer, and with what weight -- geometry only, so it is laid out once ---------------
  {
    for (int J = 1; J <= img.segcap; ++J) { img.SI2(J, slot) = 0; img.SF(S_RCW, J, slot) = 0.f; }
    for (int N = 1; N <= 6; ++N) {
      int t = 0;
      if (N <= 5 && T.LD[N] != 0 && L.RECIR[T.LD[N]] > 0.0f) t = L.LAYR[T.LD[N]];
      img.PI(Q_RCRT, N, slot) = t;
    }
    for (int K = 1; K <= L.LAY; ++K) {
      if (LRIN[K] != 1) continue;
      const int a = S.LS1[K], b = S.LS2[K];
      if (a < 1 || b < a || b > S.NSEG) continue;
      auto put = [&](int J, int mode, float w) { img.SI2(J, slot) = mode | (K << 8); img.SF(S_RCW, J, slot) = w; };
      if (a > 7) { for (int J = a; J <= b; ++J) put(J, 1, L.THICK[K]); }
      else if (K == 1) {
        float RCRF = 0.0f;
        for (int Lx = a; Lx <= b; ++Lx) {
          if (Lx < b) { put(Lx, 1, L.THICK[K]); RCRF = (float)((double)RCRF + ((double)S.STHICK[Lx] / (double)L.THICK[K])); }
          else put(Lx, 2, 1.0f - RCRF);
        }
      } else if (a == b) put(a, 3, 0.f);
      else {
        float TKM1 = 0.0f;
        for (int M = 1; M <= K - 1; ++M) TKM1 = TKM1 + L.THICK[M];
        float TJM1 = 0.0f;
        if (a > 1) for (int M = 1; M <= a - 1; ++M) TJM1 = (float)((double)TJM1 + (double)S.STHICK[M]);
        const float DSEG = TKM1 - TJM1;
        float RCRF = (float)(((double)S.STHICK[a] - (double)DSEG) / (double)L.THICK[K]);
        put(a, 2, RCRF);
        for (int Lx = a + 1; Lx <= b; ++Lx) {
          if (Lx < b) { put(Lx, 1, L.THICK[K]); RCRF = (float)((double)RCRF + ((double)S.STHICK[Lx] / (double)L.THICK[K])); }
          else put(Lx, 2, 1.0f - RCRF);
        }
      }
    }
  }
  const int hemi = (C.ULAT >= 0.0f) ? 0 : 1;
  const int IDFS = B.idfs[(int64_t)node_r * 2 + hemi];
  img.I(I_NSEG, slot) = S.NSEG; img.I(I_IPL, slot) = C.IPL; img.I(I_IHV, slot) = C.IHV; img.I(I_IDFS, slot) = IDFS;
  img.I(I_NODE_P, slot) = node_p; img.I(I_NODE_T, slot) = node_t; img.I(I_NODE_R, slot) = node_r;
  img.I(I_FLAGS, slot) = (int32_t)flag; img.I(I_NPROF, slot) = nprof; img.I(I_IPRE, slot) = C.IPRE;
  if (flags) flags[c] = flag;
  if (topo) {
    topo[0 * n + c] = S.NSEG;
    for (int N = 1; N <= 6; ++N) { topo[(int64_t)N * n + c] = S.NSEGn[N]; topo[(int64_t)(6 + N) * n + c] = T.NSTEP[N]; }
    topo[13 * n + c] = IDFS; topo[14 * n + c] = T.LP[1]; topo[15 * n + c] = T.LD[1];
  }
  if (sortkey) {
    // Warps should be uniform in (1) code path: the liner structure of the profile (IBAR of every
    // subprofile), (2) trip counts: cost = sum of segments x sub-steps, (3) weather: the node,
    // which decides rain, snow and frozen-soil days for all lanes at once, (4) soil: conductivity
    // of the top segment.  Most significant first (measured orders: profiles/r1_notes.md).
    uint64_t cost = 0, cls = 0;
    for (int N = 1; N <= 6; ++N) {
      cost += (uint64_t)S.NSEGn[N] * (uint64_t)T.NSTEP[N];
      cls = cls * 7u + (uint64_t)(S.NSEGn[N] > 0 ? 1 + img.PI(Q_IBAR, N, slot) : 0);
    }
    if (cost > 0xFFFFull) cost = 0xFFFFull;
    const uint64_t soil = ((uint64_t)(uint32_t)__float_as_int(S.RCUL[1]) >> 21) & 0x3FFull;
    // Most significant bit: the launch tier.  0 = one plain subprofile (no liner, drain layer or inflow) of at most 16
    // segments -- every natural-soil PyHELP cell: these go to the narrow kernel instantiation (16-segment state, the
    // register budget that fills one wave); 1 = everything else (help_capi.cu launch plan).
    const bool narrow = (nprof == 1) && (img.PI(Q_SUBIN, 1, slot) & 2) && S.NSEG <= 16 && !(flag & kFlagRecirc);
    if (narrow && n_narrow) atomicAdd(n_narrow, 1);
    sortkey[slot] = ((uint64_t)(narrow ? 0 : 1) << 63) | ((cls & 0x7FFFull) << 48) | (cost << 32) |
                    (((uint64_t)node_t & 0x3FFFFFull) << 10) | soil;
  }
}

}  // namespace helpb200
