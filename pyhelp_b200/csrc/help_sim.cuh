// help_sim.cuh -- the daily water-balance kernel: ONE THREAD = ONE CELL, the full
// (spin-up + n_years) daily time loop over that cell's segment profile.
//
// Physics and the reference lines each routine must agree with
// (pyhelp/HELP3O.FOR, "F:"):
//   frozen_soil()   30-day mean air temperature state machine            F:3590-3628
//   snow()          NWS HYDRO-17 accumulation / melt / lag / attenuation F:3756-4122
//   scs_runoff()    SCS curve-number runoff                              F:2752-2796
//   soil_temp7()    soil temperature at the bottom of the evaporative zone F:3516-3560
//   penman()        Penman potential ET with WGEN clear-sky radiation    F:3327-3436
//   evaporation()   surface/soil evaporation and potential transpiration F:3054-3320
//   et_distribute() ET demand over the 7 evaporative segments            F:2803-3015
//   vegetation()    residue decay, biomass, leaf area index              F:3440-3512, 3568-3582
//   route_subprofile() sub-daily routing of one subprofile: head on the liner,
//                   liner leakage (Darcy or geomembrane), lateral drainage,
//                   unsaturated drainage segment by segment (Newton), upward
//                   redistribution of super-saturation                   F:4133-4734, 4870-5304
//   recirculation of lateral drainage into a layer above (RECIRC)      F:5311-5440
//   the day/year driver, monthly accounting, spin-up restart            F:490-997
//   optional daily channel (OUTDAY columns) for selected cells          F:5984-6029
//   monthly report value semantics (x25.4, F7.3, row->key mapping)      F:6415-6600,
//                                                    pyhelp/processing.py:64-147
//
// Arithmetic types follow the reference: surface processes in REAL*4 (float),
// subsurface routing in REAL*8 (double) with the handful of implicit REAL*4
// temporaries the Fortran has (SWLIM, RELSW, PSIJP1, ... F:4291-4297;
// QSTARN F:4723; QPHOLE/QDEFEC/RSQR/DENOM F:4896-4897).  The file is compiled
// with -fmad=false: every operation rounds once, like gfortran's SSE2 code.
//
// State held per thread between days (SegState):
//   SB[j].x  segment storage (DSWUL == SWULY between sub-steps)
//   SB[j].y  storage change of the last sub-step (BALT == BALY == DCHG at day end)
//   DRCRS[j] recirculated drainage entering the segment (cells that recirculate only)
// plus ~45 float/int scalars (snow pack, frozen-soil window, vegetation, ...).
//
// The per-segment constants of the routing loop come from the tile-major routing records
// (help_image.h); what shaped the loop -- and what was tried and rejected -- is in DESIGN.md
// section 3 and profiles/r1_notes.md.
#pragma once
#include "help_image.h"

namespace helpb200 {

constexpr int kDailyCols = 16;   // pre, run, ett, swe, (hed, drn, prc) x 3 subprofiles, air-freeze, soil-freeze, sno

struct SimArgs {
  Image img;
  const float* pre;       // [node][year][kDays] inches
  const float* tmp;       // deg F
  const float* rad;       // langley/day
  const int32_t* years;   // [n_years]
  int n_years;
  int64_t n_cells;        // cells of the batch (stride of the result arrays)
  int64_t first, count;   // image slots [first, first + count) are simulated by this launch
  const int32_t* perm;    // image slot -> caller's cell index (NULL = identity)
  float* monthly;         // [7][ny][12][n] or NULL
  float* yearly;          // [5][ny][n] or NULL
  float* raw;             // [14][ny][12][n] or NULL
  uint32_t* flags;        // [n]
  // optional daily channel (OUTDAY columns, F:5984-5996) for selected cells
  double* daily;          // [n_selected][daily_days][kDailyCols] or NULL
  const int32_t* daily_row;   // [n] caller cell -> row of `daily`, -1 = not selected
  int64_t daily_days;     // days of the n_years output years
};

// value an F7.3 field takes after the reference's print + float32 re-parse
// (F:6032-6072 "12(1X,F7.3)"; processing.py:105-122).  v*1000 is exact in
// double (24+10 bits), so rint() is printf's round-half-even on the exact value.
__device__ __forceinline__ float f73(float v, uint32_t& flag) {
  const double k = rint((double)v * 1000.0);
  if (!(k < 1000000.0 && k > -100000.0)) {          // "*******" (or NaN)
    flag |= (k != k) ? kFlagNaN : kFlagOverflow;
    return __int_as_float(0x7fc00000);
  }
  return (float)(k / 1000.0);
}

// last julian day of each month in a non-leap year (MONTH, F:1073)
__constant__ int kICAL[13] = {0, 31, 59, 90, 120, 151, 181, 212, 243, 273, 304, 334, 365};

// ---- geomembrane leakage (FML, F:4870-5304) ------------------------------------
struct FmlC {
  int lcase, lpq, ltype; float thklin, phole, defec, trans; double rclin, rcs, dth;
  double rcs_ek;   // RCS**ek of the Giroud formulas (F:5010-5160): a constant of the liner, evaluated once per cell
};

// contact-quality constants of the Giroud wetted-radius formulas; form 3 only
__device__ __forceinline__ bool fml_giroud(int LC, int PQ, float& cp, float& cd, float& eh, float& ek) {
  if (LC == 2 && PQ > 1 && PQ < 5) { cp = 0.00363f; cd = 0.0229f; eh = 0.38f; ek = -0.25f; return true; }
  if (LC == 3 || LC == 4) {
    if (PQ == 2) { cp = 0.052f; cd = 0.0663f; eh = 0.5f; ek = -0.06f; return true; }
    if (PQ == 3) { cp = 0.0449f; cd = 0.0572f; eh = 0.45f; ek = -0.13f; return true; }
    if (PQ == 4) { cp = 0.1053f; cd = 0.134f; eh = 0.45f; ek = -0.13f; return true; }
  }
  return false;
}

__device__ __noinline__ double fml_leakage(const FmlC& P, double TH) {
  double QVAPOR = 0.0, RPHOLE = 0.0, RDEFEC = 0.0, HPHOLE = 0.0, HDEFEC = 0.0;
  float QPHOLE = 0.0f, QDEFEC = 0.0f, RSQR, DENOM;
  const int LC = P.lcase, PQ = P.lpq;
  const double RCS = P.rcs, DTH = P.dth;
  if (TH > (double)P.thklin) QVAPOR = 34016.0 * P.rclin * TH / (double)P.thklin;
  else if (LC == 3 || LC == 5) QVAPOR = (TH > 0.0) ? 34016.0 * P.rclin : 0.0;
  else QVAPOR = 34016.0 * P.rclin;
  if (TH > 0.0) {
    const double vap1_p = (double)(P.phole * 0.000177f), vap1_d = (double)(P.defec * 0.0356f);
    int form = 0;          // 1: free flow through holes, 2: perfect contact, 3: Giroud wetted radius
    float cp = 0.f, cd = 0.f, eh = 0.f, ek = 0.f;
    double head = TH;
    if (LC == 1) form = 1;
    else if (LC == 2) {
      if (PQ == 1) form = 2;
      else if (PQ > 1 && PQ < 5) { form = 3; cp = 0.00363f; cd = 0.0229f; eh = 0.38f; ek = -0.25f; }
      else if (PQ == 5) form = 1;
    } else if (LC == 3 || LC == 4) {
      if (LC == 4) head = TH + DTH;
      if (PQ == 1) form = 2;
      else if (PQ == 2) { form = 3; cp = 0.052f; cd = 0.0663f; eh = 0.5f; ek = -0.06f; }
      else if (PQ == 3) { form = 3; cp = 0.0449f; cd = 0.0572f; eh = 0.45f; ek = -0.13f; }
      else if (PQ == 4) { form = 3; cp = 0.1053f; cd = 0.134f; eh = 0.45f; ek = -0.13f; }
      else if (PQ == 5) form = 1;
    }
    if (form == 1) {
      if (LC == 4) {     // sic: F:5167-5168 divides only DTHICK by THICK(LINER)
        QPHOLE = (float)(vap1_p * (TH + DTH / (double)P.thklin));
        QDEFEC = (float)(vap1_d * r8_sqrt(TH + DTH));
      } else {
        QPHOLE = (float)(vap1_p * TH / (double)P.thklin);
        QDEFEC = (float)(vap1_d * r8_sqrt(TH));
      }
    } else if (form == 2) {
      QPHOLE = (float)((double)(P.phole * 0.000671f) * head * RCS);
      QDEFEC = (float)((double)(P.defec * 0.00758f) * head * RCS);
    } else if (form == 3) {
      // head**eh is shared by the two formulas, RCS**ek is a constant of the liner (same bits as
      // evaluating each power where the reference writes it)
      const double ph = r8_pow(head, (double)eh), pr = P.rcs_ek;
      if (P.phole > 0.0f) {
        RPHOLE = (double)cp * ph * pr;
        HPHOLE = 1.0 + (head / (2.0 * DTH * r8_log(RPHOLE / (double)0.0005f)));
        QPHOLE = (float)((double)(P.phole * 26.4f) * RCS * HPHOLE * (RPHOLE * RPHOLE));
      }
      if (P.defec > 0.0f) {
        RDEFEC = (double)cd * ph * pr;
        HDEFEC = 1.0 + (head / (2.0 * DTH * r8_log(RDEFEC / (double)0.00564f)));
        QDEFEC = (float)((double)(P.defec * 26.4f) * RCS * HDEFEC * (RDEFEC * RDEFEC));
      }
    }
    if (LC == 5 || LC == 6) {            // geotextile cushion: fixed-point wetted radius (F:5177-5285)
      head = (LC == 5) ? TH : (TH + DTH);
      RPHOLE = (double)0.001f;
      RDEFEC = (double)0.0113f;
      const double rhs = (double)(4.0f * P.trans) * head / (double)10000.f / (double)39.37f;
      if (P.phole > 0.0f) {
        if (((RCS / (double)100.f) * (double)0.000001f) < rhs) {
          for (int I = 1; I <= 10; ++I) {
            RSQR = (float)rhs;
            DENOM = (float)(2.0 * r8_log(RPHOLE / (double)0.0005f));
            const double t = (double)0.0005f / RPHOLE;
            DENOM = (float)((double)DENOM + (t * t) - 1.0);
            RSQR = (float)((double)(RSQR * 100.f) / RCS / (double)DENOM);
            RPHOLE = (double)r4_sqrt(RSQR);
          }
        }
        HPHOLE = 1.0 + (head / (2.0 * DTH * r8_log(RPHOLE / (double)0.0005f)));
        QPHOLE = (float)((double)(P.phole * 26.4f) * RCS * HPHOLE * (RPHOLE * RPHOLE));
      }
      if (P.defec > 0.0f) {
        if (((RCS / (double)100.f) * (double)0.000127f) < rhs) {
          const double cdef = (LC == 5) ? (double)0.00564f : (double)0.00565f;   // sic F:5214 / F:5271
          for (int I = 1; I <= 10; ++I) {
            RSQR = (float)rhs;
            DENOM = (float)(2.0 * r8_log(RDEFEC / (double)0.00564f));
            const double t = cdef / RDEFEC;
            DENOM = (float)((double)DENOM + (t * t) - 1.0);
            RSQR = (float)((double)(RSQR * 100.f) / RCS / (double)DENOM);
            RDEFEC = (double)r4_sqrt(RSQR);
          }
        }
        HDEFEC = 1.0 + (head / (2.0 * DTH * r8_log(RDEFEC / (double)0.00564f)));
        QDEFEC = (float)((double)(P.defec * 26.4f) * RCS * HDEFEC * (RDEFEC * RDEFEC));
      }
    }
  }
  double Q = QVAPOR + (double)QPHOLE + (double)QDEFEC;
  if ((LC == 3 && P.ltype == 3) || LC == 4) {
    const double cap = 34016.0 * RCS * (TH + DTH) / DTH;
    if (Q > cap) Q = cap;
  } else {
    if (Q > (34016.0 * RCS)) Q = 34016.0 * RCS;
  }
  return Q;
}

struct LatC { double dslope, dxleng, alpha, sa, ca, A, B, C, D, lbh, lbl; };   // constants of LATFLO (F:4686-4712) for one drain layer
// LATFLO's slope/length terms (F:4686-4712): they depend on the drain layer only
__device__ __forceinline__ LatC lateral_consts(float slope_pct, float xleng_ft) {
  LatC L;
  const double PI = 3.141592654;
  L.dslope = (double)slope_pct / 100.0;
  L.dxleng = (double)xleng_ft * 12.0;
  if (L.dxleng == 0.0) L.dxleng = 1.0;
  L.alpha = r8_atan(L.dslope);
  L.sa = r8_sin(L.alpha); L.ca = r8_cos(L.alpha);
  L.A = PI / 4.0 / L.ca;
  L.B = 2.0 * r8_sqrt(0.4) / PI;
  L.C = 0.4 * (L.sa * L.sa);
  L.D = 0.5 / r8_log(L.B);
  r8_log_pair(L.B, L.lbh, L.lbl);            // B is the base of a power in every LATFLO iteration
  return L;
}

// ---- lateral drainage rate for a head TH (LATKS + LATFLO, F:4626-4734) --------------
__device__ __noinline__ double lateral_rate(const Image& img, int64_t s, int NSEGB, int NSEGL, double TH, const LatC& L) {
  double ELKS = 0.0;
  {
    double H = TH;
    for (int K = NSEGL; K >= NSEGB; --K) {
      if (H <= 0.0) break;
      const double th = (double)__ldg(&img.SF(S_THICK, K, s)), rc = (double)__ldg(&img.SF(S_RC, K, s));
      if (!(H > th)) { ELKS = ELKS + rc * H / TH; break; }
      ELKS = ELKS + rc * th / TH;
      H = H - th;
    }
  }
  const double EPS = 0.030, PI = 3.141592654;
  const double YSTAR = TH / L.dxleng;
  double QSTAR = 2.0 * YSTAR * L.sa * L.ca;
  if (L.alpha <= 0.0) { const double t = (4.0 / PI) * YSTAR; return (t * t) * ELKS; }
  if (YSTAR < ((double)0.2f * L.dslope)) return QSTAR * ELKS;
  float QSTARN;
  int ITER = 1;
  for (;;) {
    const double Fq = L.A * r8_pow_logbase(L.B, L.lbh, L.lbl, r8_pow(QSTAR / L.C, L.D));
    QSTARN = (float)((YSTAR * YSTAR) / (Fq * Fq));
    const double TOLER = 2.0 * fabs((double)QSTARN - QSTAR) / ((double)QSTARN + QSTAR);
    if (EPS < TOLER && ITER < 10) { ITER = ITER + 1; QSTAR = (QSTAR + (double)QSTARN) / 2.0; continue; }
    break;
  }
  return (QSTAR + (double)QSTARN) * ELKS / 2.0;
}

// ---- one subprofile, one day (DRAIN/AVROUT/PROFIL/HEAD, F:4133-4617) ------------------
// SW/BAL: per-segment state (1-based).  DRIN: scratch, DRIN[j] = inflow into segment j.
// ET7[1..7]: today's evapotranspiration demand per evaporative segment (inches).
// Per-segment constants come from the routing records (help_image.h): one pointer pair walks
// down the profile, every constant is a load at a compile-time offset.
struct ProfC {
  int n, NSEGB, NSEGE, NSEG, nstep, ibar, lat, subin, rcr;   // rcr: the cell recirculates drainage (DRCRS may be non-zero)
  float drculs, dthics, slope, xleng, sublin;
  FmlC fml;
  LatC latc;
};


// Per-cell segment state, one object so that the routing function addresses all of it from one
// pointer: SB[j] = {SW, BAL} of segment j (always read and written together: one 16-byte
// access), DRIN[j] = inflow into segment j during the current sub-step.
template <int MAXSEG>
struct SegState {
  double2 SB[MAXSEG + 2];
  double DRIN[MAXSEG + 3];
  double DRCRS[MAXSEG + 2];   // recirculated drainage entering segment j, in/day (RECIRC); touched only by cells that recirculate
};
struct RouteOut { double DRN, PRC, HED; float EXTRA; uint32_t flag; };
#define SWa(j) SS.SB[(j)].x
#define BALa(j) SS.SB[(j)].y
#define DRINa(j) SS.DRIN[(j)]

// PLAIN = the subprofile has no liner (IBAR = 0), no subsurface inflow and no lateral-drain
// layer below it (every natural-soil profile of a PyHELP grid): the liner, lateral-drainage
// and inflow terms fold away at compile time.  Adding the inflow term 0.0 is the identity
// there (none of the partial sums can be -0.0), so it is dropped too.
template <int MAXSEG, bool PLAIN>
__device__ __noinline__ RouteOut route_subprofile(const Image& img, int64_t s, const ProfC& P, float FIN,
                                                  const float* __restrict__ ET7, int IFREZ,
                                                  SegState<MAXSEG>& SS) {
  uint32_t flag = 0;
  float EXTRA = 0.0f;
  const int NSEGB = P.NSEGB, NSEGE = P.NSEGE, NSEG = P.NSEG;
  const int IBAR = PLAIN ? 0 : P.ibar, LAT = PLAIN ? 0 : P.lat;
  const int NSEGL = (IBAR == 0 || IBAR == 3) ? NSEGE : NSEGE - 1;
  const double DT = 1.0 / (double)P.nstep;
  const double F = (double)FIN * DT;
  const double DSUBE = PLAIN ? 0.0 : (double)P.sublin * DT;            // DSUB(NSEGE)
  const bool bottom = (NSEGE == NSEG);
  const bool liner = !(IBAR == 0 || IBAR == 3);
  const bool has_sub = PLAIN ? false : ((P.subin & 1) != 0);
  const bool has_rcr = PLAIN ? false : (P.rcr != 0);
  const char* const p0 = img.rec_p(s) + (int64_t)(NSEGB - 1) * kRecBytes;
  const char* const pL = p0 + (int64_t)(NSEGL - NSEGB) * kRecBytes;
  const double LOWFC_L = rec_tail(pL).lowfc;            // SWLIM = FC occurs at segment NSEGL only
  double Ed[8];                                          // E(J) = ET(J) dt of the evaporative segments
#pragma unroll
  for (int J = 1; J <= 7; ++J) Ed[J] = (double)ET7[J] * DT;
  double EXCESS = 0.0, PRC = 0.0, DRN = 0.0, HED = 0.0;
  for (int K = 1; K <= P.nstep; ++K) {
    DRINa(NSEGB) = F + EXCESS;
    EXCESS = 0.0;
    double QPERCY = 0.0, QLATY = 0.0;
    if (liner) {
      // HEAD (F:4582-4617): saturated depth above the liner, bottom -> top
      double THY = 0.0;
      for (int Kk = NSEGL; Kk >= NSEGB; --Kk) {
        const double fc = (double)__ldg(&img.SF(S_FC, Kk, s));
        const double cur = SWa(Kk) + (BALa(Kk) / 2.0);
        if (cur <= fc) break;
        const double th = (double)__ldg(&img.SF(S_THICK, Kk, s)), ul = (double)__ldg(&img.SF(S_UL, Kk, s));
        double XHEAD = th * (cur - fc) / (ul - fc);
        if (XHEAD > th) XHEAD = th + SWa(Kk) + (BALa(Kk) / 2.0) - ul;
        THY = THY + XHEAD;
        XHEAD = XHEAD - th;
        if ((XHEAD + 0.00005) < 0.0) break;
      }
      HED = HED + THY * DT;                    // mean head on the liner over the day (F:4253)
      if (THY > 0.0) {
        // (IBAR 4/5 = bottom liner with subsurface inflow: the Fortran computes no leakage, F:4241-4242)
        if (IBAR == 1) QPERCY = (double)P.drculs * (THY + (double)P.dthics) / (double)P.dthics;
        if (IBAR == 2) QPERCY = fml_leakage(P.fml, THY);
        if (LAT == 1) QLATY = lateral_rate(img, s, NSEGB, NSEGL, THY, P.latc);
      }
    }
    // ---- AVROUT (F:4287-4527): top -> bottom ---------------------------------------------------
    // Nothing is carried from one segment to the next across the powers (register pressure):
    // segment J's own vectors are loaded at the top of its iteration -- L1 hits, they were
    // prefetched one iteration earlier -- segment J+1's are prefetched now and loaded where the
    // capillary limit of J needs them.  (Carrying J+1's vectors in registers into the next
    // iteration measured 0..5 % slower at every register budget.)
    const char* p = p0;
    double drin = DRINa(NSEGB);
    for (int J = NSEGB; J <= NSEGL; ++J, p += kRecBytes) {
      const char* p1 = (J < NSEGL) ? p + kRecBytes : p;        // always a valid record
      const float4 A = rec_f4<RV_A>(p), B = rec_f4<RV_B>(p);
      const double2 Cv = rec_d2<RV_C>(p), Dv = rec_d2<RV_D>(p);
      const double2 sbj = SS.SB[J];
      const double swj = sbj.x, balj = sbj.y;
      rec_prefetch(p1);
      const double2 Ev = rec_d2<RV_E>(p);                 // {B of the Newton solve, LOW at WP}
      const double ul = (double)A.x, rs = (double)A.y, wp = (double)A.z, fc = (double)A.w;
      const double rspan = Cv.x, kdt = Cv.y;
      const double span = ul - rs;
      const double Ej = (J <= 7) ? Ed[J] : 0.0;
      const double DSUBj = has_sub ? img.SD(SD_SUBIN, J, s) * DT : 0.0;      // (REAL*8: not in the routing records; rare)
      const double DRCRj = has_rcr ? SS.DRCRS[J] * DT : 0.0;            // DRCR(J) (F:4195)
      // lower storage limit SWLIM and the drainage LOW at that limit; kind: 0 capillary
      // equilibrium with the segment below, 1 wilting point, 2 field capacity, 3 porosity
      // (frozen soil)
      int kind;
      if (J <= 7) {
        kind = (drin == 0.0 && J < NSEGL) ? 0 : 1;
        if (J == NSEGL) kind = 2;
        if (J == NSEG) kind = 2;
        if (J == NSEG && drin > 0.0) kind = 1;
        if (IFREZ == 1) kind = 3;
      } else if (J == NSEGL && J != NSEG) kind = 2;
      else if (J == NSEG && drin > 0.0) kind = 1;
      else if (J == NSEG && drin <= 0.0) kind = 2;
      else kind = 0;
      double swl, LOW;
      if (kind == 0) {
        // F:4330-4340 == F:4356-4366: matric potential of J+1 (Brooks-Corey) carried to J,
        // clipped to [WP, FC]; then LOW = K dt Theta(SWLIM)^C (F:4420-4421)
        const float4 A1 = rec_f4<RV_A>(p1), B1 = rec_f4<RV_B>(p1);
        const double2 C1 = rec_d2<RV_C>(p1), D1 = rec_d2<RV_D>(p1);
        const double2 sb1 = SS.SB[J + 1];              // (SB has MAXSEG + 2 entries)
        const double sw1 = sb1.x, bal1 = sb1.y;
        const double ul1 = (double)A1.x, rs1 = (double)A1.y, bub1 = (double)B1.y;
        const double bub = (double)B.y, lam = (double)B.x;
        const double base1 = sw1 + (bal1 / 2.0);
        float RELSW = (float)fdiv_by(base1 - rs1, ul1 - rs1, C1.x);
        if (RELSW < B1.z) RELSW = B1.z;
#ifdef HELP_DBG_CAPMEMO
        HELP_DBG_CAPMEMO
#endif
        float PSIJP1 = (float)(bub1 * r8_pow((double)RELSW, D1.y));
        if ((double)PSIJP1 < bub) PSIJP1 = (float)bub;
        const float RELSWL = (float)r8_pow(bub / (double)PSIJP1, lam);
        float SWLIM = (float)(((double)RELSWL * span) + rs);
        if ((double)SWLIM < wp) SWLIM = (float)wp;
        if ((double)SWLIM > fc) SWLIM = (float)fc;
        swl = (double)SWLIM;
        LOW = kdt * r8_pow(fdiv_by(swl - rs, span, rspan), Dv.x);
      }
      else if (kind == 1) { swl = wp; LOW = Ev.y; }
      else if (kind == 2) { swl = fc; LOW = LOWFC_L; }
      else { swl = ul; LOW = kdt; }            // ((ul - rs)/(ul - rs))**C == 1
      const double base = swj + (balj / 2.0);                     // storage at the start of the step
#ifdef HELP_DBG_SWEEP
      HELP_DBG_SWEEP
#endif
      double DRNMAX = PLAIN ? (base + drin - Ej - swl) : (base + drin + DSUBj + DRCRj - Ej - swl);
      if (J == NSEGL) {
        if (bottom) {
          if (IBAR == 3) DRNMAX = 0.0;
          if (IBAR > 3) DRNMAX = DRNMAX + DSUBE;
          if (IBAR >= 3) {
            DRINa(NSEGE + 1) = 0.0;
            if (LAT == 0) { DRINa(NSEGL + 1) = 0.0; break; }
          }
        }
        if (liner) {    // split the gravity water between leakage and lateral flow (F:4384-4413)
          double GRVWAT = DRNMAX + swl - fc;
          if (GRVWAT <= 0.0) {
            DRINa(NSEGE + 1) = 0.0; DRINa(NSEGL + 1) = 0.0;
            if (IBAR < 3) DRINa(NSEGE + 1) = DRINa(NSEGE + 1) + DSUBE;
          } else {
            if (GRVWAT >= ((QPERCY + QLATY) * DT)) {
              DRINa(NSEGL + 1) = (QPERCY + QLATY) * DT;
              DRINa(NSEGE + 1) = QPERCY * DT;
            } else if (LAT == 0) {
              DRINa(NSEGL + 1) = GRVWAT; DRINa(NSEGE + 1) = GRVWAT;
            } else {
              DRINa(NSEGL + 1) = GRVWAT;
              const double QP = (GRVWAT / DT) * QPERCY / (QPERCY + QLATY);
              DRINa(NSEGE + 1) = QP * DT;
            }
            if (IBAR < 3) DRINa(NSEGE + 1) = DRINa(NSEGE + 1) + DSUBE;
          }
          break;
        }
      }
      if (DRNMAX < 0.0) DRNMAX = 0.0;
      const bool add_sub = (J == NSEGL && bottom && IBAR > 3);
      double out;
      if (LOW > DRNMAX) out = DRNMAX;
      else {
        double SWMAX = PLAIN ? (base + drin - Ej) : (base + DSUBj + DRCRj + drin - Ej);
        if (add_sub) SWMAX = SWMAX + DSUBE;
        if (SWMAX > ul) SWMAX = ul;
        if (SWMAX <= swl) out = 0.0;
        else {
          const double C = Dv.x;
          double HIGH = kdt * r8_pow(fdiv_by(SWMAX - rs, span, rspan), C);
          if (HIGH > DRNMAX) HIGH = DRNMAX;
          double SWMIN = SWMAX - HIGH;
          if (SWMIN < swl) SWMIN = swl;
          LOW = kdt * r8_pow(fdiv_by(SWMIN - rs, span, rspan), C);
          if (LOW > DRNMAX) LOW = DRNMAX;
          SWMAX = PLAIN ? (base + drin - Ej - LOW) : (base + DSUBj + DRCRj + drin - Ej - LOW);
          if (add_sub) SWMAX = SWMAX + DSUBE;
          if (SWMAX > ul) SWMAX = ul;
          if (SWMAX <= SWMIN) out = DRNMAX;
          else {
            // Newton on f(x) = A - B x^C - 2x (F:4461-4488); B = K dt (1/(ul-rs))^C is a
            // constant of the segment (set-up kernel).  x^C and x^(C-1) share one logarithm.
            double An = PLAIN ? (2.0 * (swj - rs) + balj + drin - Ej) : (2.0 * (swj - rs) + balj + drin + DSUBj + DRCRj - Ej);
            if (add_sub) An = An + (2.0 * DSUBE);
            const double Bc = Ev.x;
            const double TOLER = 0.003 * span;
            const double ftol = (double)0.3f * TOLER;
            const double Cm1 = C - 1.0;
            double X0 = (SWMAX + SWMIN) / 2.0;
            int ITER;
#pragma unroll 1
            for (ITER = 1; ITER <= 100; ++ITER) {
              double pc, pc1;
              r8_pow_sb(X0, C, Cm1, pc, pc1);
              const double FX0 = An - (Bc * pc) - (2.0 * X0);
              if (FX0 <= ftol) break;
              const double FX0DER = -(Bc * C * pc1) - 2.0;
              if (fabs(FX0DER) < 1.0e-03) break;
              const double DX = FX0 / FX0DER;
              X0 = X0 - DX;
              // (written as the negated test so that a NaN step leaves the loop too: from there on the
              // reference only repeats NaN - NaN until its 100th iteration and reports non-convergence,
              // F:4481-4483 -- same X0, same flag, without the 99 idle iterations that would make one bad
              // cell hold back its whole batch)
              if (!(fabs(DX) >= TOLER)) { if (DX != DX) ITER = 101; break; }
            }
            if (ITER > 100) flag |= kFlagNonConv;
            if (X0 > SWMAX) X0 = SWMAX;
            if (X0 < SWMIN) X0 = SWMIN;
            out = PLAIN ? (2.0 * (swj - X0) + balj + drin - Ej) : (2.0 * (swj - X0) + balj + drin + DRCRj + DSUBj - Ej);
            if (bottom && IBAR > 3) out = out + DSUBE;
          }
        }
      }
      const int above = PLAIN ? 0 : __float_as_int(B.w);        // bit0: J above the drain layer, bit1: J+1 is
      if (above & 1) { if (out > kdt) out = kdt; }
      if (out > DRNMAX) out = DRNMAX;
      if (out < 0.0) out = 0.0;
      if (J < NSEGL && (above & 2)) {
        const float4 A1 = rec_f4<RV_A>(p1);
        const double2 C1 = rec_d2<RV_C>(p1), sb1 = SS.SB[J + 1];
        const double sw1 = sb1.x, bal1 = sb1.y;
        const double kdt1 = C1.y;
        if (out > kdt1) {
          const RecTail t0 = rec_tail(p), t1 = rec_tail(p + kRecBytes);
          const double th = (double)t0.thick, th1 = (double)t1.thick;
          const double ul1 = (double)A1.x, sub1 = has_sub ? img.SD(SD_SUBIN, J + 1, s) : 0.0;
          const float sw = (float)fdiv_by(base - rs, span, rspan);
          const float TRATIO = (float)(th / th1);
          const double En = (J + 1 <= 7) ? Ed[J + 1] : 0.0;
          double DRMX = ul1 - sw1 - (bal1 / 2.0) + kdt1 + En - sub1 * DT - (has_rcr ? SS.DRCRS[J + 1] * DT : 0.0);
          if (DRMX > DRNMAX) DRMX = DRNMAX;
          double DRNMX = kdt1 * (double)(1.0f + (sw * TRATIO));
          if (DRNMX > DRMX) DRNMX = DRMX;
          if (DRNMX > DRNMAX) DRNMX = DRNMAX;
          if (out > DRNMX) out = DRNMX;
          if (out < 0.0) out = 0.0;
        }
      }
      DRINa(J + 1) = out;
      drin = out;
    }
    // F:4518-4525 and PROFIL (F:4535-4574) fused: new balance bottom -> top, pushing
    // water above saturation into the segment above
    if (IBAR == 3) DRINa(NSEGE + 1) = 0.0;
    const double prc_step = DRINa(NSEGE + 1);
    double EXC = 0.0;
    {
      const char* pk = pL;
      double dnext = DRINa(NSEGL + 1);
      double din = DRINa(NSEGL), swk, balk;
      { const double2 sb = SS.SB[NSEGL]; swk = sb.x; balk = sb.y; }
      float ulf = rec_f4<RV_A>(pk).x;
      for (int Kk = NSEGL; Kk >= NSEGB; --Kk) {
        const double sub = has_sub ? img.SD(SD_SUBIN, Kk, s) * DT : 0.0;
        // the loads of the next segment up, ahead of their use; unconditional (Kk-1 >= 0 is a
        // valid index of SB and DRIN, the record pointer stays on the top segment)
        if (Kk > NSEGB) pk -= kRecBytes;
        const double din_n = DRINa(Kk - 1);
        const float ulf_n = rec_f4<RV_A>(pk).x;
        const double2 sbn = SS.SB[Kk - 1];
        const double swk_n = sbn.x, balk_n = sbn.y;
        const double ul = (double)ulf;
        const double Ek = (Kk <= 7) ? Ed[Kk] : 0.0;
        double BALT = PLAIN ? (din - dnext - Ek) : (din + sub + (has_rcr ? SS.DRCRS[Kk] * DT : 0.0) - dnext - Ek);
        if (Kk == NSEGL && IBAR > 3) BALT = BALT + DSUBE;
        const double DSW = swk + ((balk + BALT) / 2.0);
        const double S = DSW + EXC + (BALT / 2.0);
        double EXCE = EXC;
        if (EXCE < 0.0) EXCE = 0.0;
        EXC = S - ul;
        double newsw;
        if (EXC > 0.0) {
          BALT = ul - swk - (balk / 2.0);
          newsw = ul - (BALT / 2.0);
        } else {
          EXC = 0.0;
          BALT = BALT + EXCE;
          newsw = swk + ((balk + BALT) / 2.0);
        }
        SS.SB[Kk] = make_double2(newsw, BALT);
        dnext = din; din = din_n; swk = swk_n; balk = balk_n; ulf = ulf_n;
      }
    }
    EXCESS = EXC;
    EXTRA = (float)EXCESS;
    PRC = PRC + prc_step;
    // left to right, as the reference sums it (F:4250): the daily channel exposes the last bit
    DRN = liner ? (DRN + DRINa(NSEGL + 1) - DRINa(NSEGE + 1) + DSUBE) : 0.0;
  }
  RouteOut R;
  R.DRN = DRN; R.PRC = PRC; R.HED = HED; R.EXTRA = EXTRA; R.flag = flag;
  return R;
}

// ---- evapotranspiration demand over the 7 evaporative segments (ETCHK/WFS) --------------
__device__ __forceinline__ float awf(float D) {
  return 4.759177E-4f + 3.14529f * D - 4.319952f * D * D + 3.089057f * D * D * D - 0.9150572f * D * D * D * D;
}
__device__ __forceinline__ float wfs(float D1, float D2, float D3) {
  const float R1 = (D1 < D3) ? D1 / D3 : 1.0f;
  const float R2 = (D2 < D3) ? D2 / D3 : 1.0f;
  return (R1 < R2) ? (awf(R2) - awf(R1)) : 0.0f;
}

struct Surf {      // REAL*4 surface-process state of one cell
  // snow pack (F:123 /BLK28/, /BLK9/)
  float WE, XNEGHS, XLIQW, TINDEX, STORGE, EXLAG[5], TWE, SNO, OLDSNO, OSNO;
  // frozen soil (F:122, F:1190)
  float TMPSUM, TAVG[30]; int ring, KCNT, MXKCNT, IFREZ, IDFS;
  // evaporation / vegetation
  float ES1T, TS2, XLAI, DLAI, GS, RSD, CV, DM, TS, ST7, TSEG7, EAJ;
  // today's fluxes
  float RAIN, SMELT, GMRO, ADDRUN, RUN, PINF, ETT, ESS, EP, ES, ETO, ABST, ET[8];
};

// Block shape.  Every warp of a block is kept in the same day of the simulation by one
// barrier per day (HELP_SIM_DAYSYNC): the routing code is larger than the 32 KB L1.5
// instruction cache, and warps that run the same region together share its fetches
// (measured on B200: 100 000 natural cells 716 ms with the barrier, 838 ms without; profiles/r2_notes.md).
// The kernel is compiled for several register budgets (CAP = threads per SM a budget keeps resident =
// the largest block of the instantiation: 128 registers -> 512, 96 -> 640, 80 -> 768, 64 -> 1024).
// All cells of a launch take about the same time, so the launcher (help_capi.cu) chooses the budget
// with the most registers that still fits the batch into the fewest, evenly filled waves (pick_shape),
// and the block size by profile class (pick_block): 128 threads for plain profiles, one block per SM
// for batches with liner profiles, whose routing code does not fit the instruction cache.
#ifndef HELP_SIM_NO_DAYSYNC
#define HELP_SIM_DAYSYNC 1
#endif
template <int MAXSEG, int CAP, int REGS>
__global__ void __launch_bounds__(CAP) __maxnreg__(REGS) help_sim_kernel(SimArgs A) {
  helpmath::hm_load_tables();
  // No thread leaves before the end: the day barrier below must be reached by every thread of the
  // block the same number of times.  A thread past the end of the batch, or one whose cell cannot
  // be simulated, reads the last cell's constants (valid addresses) and sits out every day
  // (`live == false`); it only keeps the barrier count.
#ifndef HELP_FWD_BLOCKS
  // The sort key ascends with the cost of a cell and the hardware hands out blocks in index order: the last blocks -- the ones
  // that land on SMs which already hold a full share, and the tail of a launch with more blocks than the GPU keeps resident --
  // must be the cheapest.  Measured on B200: 100 000 natural cells 5.2 % faster, 125 000 mixed cells 8.6 % (profiles/r2_notes.md).
  const int64_t s_raw = (int64_t)(gridDim.x - 1 - blockIdx.x) * blockDim.x + threadIdx.x;
#else
  const int64_t s_raw = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
#endif
  bool live = s_raw < A.count;
  const int64_t s = A.first + (live ? s_raw : A.count - 1);
  const Image& img = A.img;
  const int64_t n = A.n_cells;
  const int64_t cell = A.perm ? A.perm[s] : s;
  const int NY = A.n_years;
  uint32_t flag = (uint32_t)img.I(I_FLAGS, s);
  const int NSEG_img = img.I(I_NSEG, s);
  const bool badcell = (NSEG_img <= 0 || NSEG_img > MAXSEG);
  const int NSEG = badcell ? 0 : NSEG_img;
  if (badcell && live) {
    flag |= kFlagBadCell;
    const float qnan = __int_as_float(0x7fc00000);
    for (int y = 0; y < NY; ++y) {
      for (int m = 0; m < 12; ++m) {
        if (A.monthly) for (int r = 0; r < 7; ++r) A.monthly[(((int64_t)r * NY + y) * 12 + m) * n + cell] = qnan;
        if (A.raw) for (int r = 0; r < 14; ++r) A.raw[(((int64_t)r * NY + y) * 12 + m) * n + cell] = qnan;
      }
      if (A.yearly) for (int r = 0; r < 5; ++r) A.yearly[((int64_t)r * NY + y) * n + cell] = qnan;
    }
    A.flags[cell] = flag;
  }
  if (badcell) live = false;
  // ---- constants of this cell --------------------------------------------------------------
  const float EDEPTH = img.F(F_EDEPTH, s), SEDEP = img.F(F_SEDEP, s), SMXN = img.F(F_SMXN, s),
              SMXF = img.F(F_SMXF, s), FRUNOF = img.F(F_FRUNOF, s), CONA = img.F(F_CONA, s),
              STAGE1 = img.F(F_STAGE1, s), ULAT = img.F(F_ULAT, s), ULAI = img.F(F_ULAI, s),
              WIND = img.F(F_WIND, s), PHU = img.F(F_PHU, s), AVT = img.F(F_AVT, s), AMP = img.F(F_AMP, s),
              DD = img.F(F_DD, s), Z7 = img.F(F_Z7, s), FCS7 = img.F(F_FCS7, s), TFSOIL = img.F(F_TFSOIL, s);
  const float RHq[4] = {img.F(F_RH1, s), img.F(F_RH2, s), img.F(F_RH3, s), img.F(F_RH4, s)};
  const int IPL = img.I(I_IPL, s), IHV = img.I(I_IHV, s);
  const int nprof = img.I(I_NPROF, s);
  int IPRE = img.I(I_IPRE, s);
  const float* __restrict__ wpre = A.pre + (int64_t)img.I(I_NODE_P, s) * NY * kDays;
  const float* __restrict__ wtmp = A.tmp + (int64_t)img.I(I_NODE_T, s) * NY * kDays;
  const float* __restrict__ wrad = A.rad + (int64_t)img.I(I_NODE_R, s) * NY * kDays;
  const float WF[8] = {0, 0.085f, 0.334f, 0.252f, 0.151f, 0.091f, 0.054f, 0.033f};
  const float TB = 5.f, TO = 20.f;
  const float WFT5 = (0.2f / (1.f - 0.4f + 0.2f)) * 5.f;            // WFT(MO)*5. (F:2419, 3545)
  // evaporative-zone segment constants (REAL*4 mirrors, /BLK7/)
  float STH[8], ULf[8], FCf[8], WPf[8], RCf[8];
#pragma unroll
  for (int J = 1; J <= 7; ++J) {
    STH[J] = img.SF(S_THICK, J, s); ULf[J] = img.SF(S_UL, J, s); FCf[J] = img.SF(S_FC, J, s);
    WPf[J] = img.SF(S_WP, J, s);    RCf[J] = img.SF(S_RC, J, s);
  }
  // subprofile descriptors
  ProfC PR[7];
  int LDn[7] = {0, 0, 0, 0, 0, 0, 0};
  float RECIR[7];
  int RCRT[7] = {0, -1, -1, -1, -1, -1, -1};      // layer the recirculated drainage of subprofile n goes to (0 = none)
  const bool has_rcr = (flag & kFlagRecirc) != 0;
  {
    int e = 0;
    for (int N = 1; N <= 6; ++N) {
      ProfC& p = PR[N];
      const int ns = img.PI(Q_NSEGN, N, s);
      p.n = N; p.NSEGB = e + 1; p.NSEGE = e + ns; p.NSEG = NSEG;
      p.nstep = img.PI(Q_NSTEP, N, s); p.ibar = img.PI(Q_IBAR, N, s); p.lat = img.PI(Q_LAT, N, s); p.subin = img.PI(Q_SUBIN, N, s);
      p.rcr = (flag & kFlagRecirc) ? 1 : 0;
      p.drculs = img.PF(P_DRCULS, N, s); p.dthics = img.PF(P_DTHICS, N, s);
      p.slope = img.PF(P_SLOPE, N, s); p.xleng = img.PF(P_XLENG, N, s); p.sublin = img.PF(P_SUBLIN, N, s);
      p.fml.lcase = img.PI(Q_LCASE, N, s); p.fml.lpq = img.PI(Q_LPQ, N, s); p.fml.ltype = img.PI(Q_LTYPE, N, s);
      p.fml.thklin = img.PF(P_THKLIN, N, s); p.fml.phole = img.PF(P_PHOLE, N, s);
      p.fml.defec = img.PF(P_DEFEC, N, s); p.fml.trans = img.PF(P_TRANS, N, s);
      p.fml.rclin = img.PD(D_RCLIN, N, s); p.fml.rcs = img.PD(D_RCS, N, s); p.fml.dth = (double)p.dthics;
      {
        float cp, cd, eh, ek = 0.f;
        p.fml.rcs_ek = ((p.ibar == 2 || p.ibar == 5) && fml_giroud(p.fml.lcase, p.fml.lpq, cp, cd, eh, ek)) ? r8_pow(p.fml.rcs, (double)ek) : 0.0;
        if (p.lat == 1) p.latc = lateral_consts(p.slope, p.xleng);
      }
      RECIR[N] = img.PF(P_RECIR, N, s);
      RCRT[N] = img.PI(Q_RCRT, N, s);
      LDn[N] = (N <= 5 && ns > 0) ? img.PI(Q_LD, N, s) : 0;
      e += ns;
    }
  }
  // ---- state (run_simulation F:200-317; READIN F:1344-1351; INITIA F:2437-2442) ---------------
  // Segment state: thread-local arrays (L1-cached).  Keeping them in shared memory instead was
  // measured slower on B200: the carve-out shrinks the L1 the routing records stream through
  // (profiles/r1_notes.md).
  SegState<MAXSEG> SS;
  for (int J = 1; J <= NSEG; ++J) { SWa(J) = (double)img.SF(S_SW0, J, s); BALa(J) = 0.0; }
  if (has_rcr) for (int J = 0; J < MAXSEG + 2; ++J) SS.DRCRS[J] = 0.0;
  SWa(0) = 0.0; BALa(0) = 0.0; SWa(NSEG + 1) = 0.0; BALa(NSEG + 1) = 0.0;
  Surf U;
  U.RUN = 0.f; U.ES1T = 0.f; U.TS2 = 0.f; U.ADDRUN = 0.f;
  for (int J = 0; J < 8; ++J) U.ET[J] = 0.f;
  float EXWAT[7] = {0, 0, 0, 0, 0, 0, 0};
  U.XLAI = img.F(F_XLAI0, s); U.DLAI = U.XLAI; U.GS = img.F(F_GS0, s); U.RSD = img.F(F_RSD0, s);
  U.CV = U.RSD; U.DM = img.F(F_DM0, s); U.TS = img.F(F_TS0, s); U.ST7 = img.F(F_ST7, s); U.TSEG7 = 0.f;
  U.IFREZ = 0; U.TMPSUM = 990.0f; U.ring = 0; U.KCNT = 0; U.IDFS = img.I(I_IDFS, s);
  U.MXKCNT = (U.IDFS + 2) / 3;
  for (int k = 0; k < 30; ++k) U.TAVG[k] = 33.0f;
  U.OSNO = img.F(F_OSNO, s); U.OLDSNO = U.OSNO; U.SNO = 0.f;
  U.WE = U.XNEGHS = U.XLIQW = U.TINDEX = U.STORGE = U.TWE = 0.f;
  for (int k = 0; k < 5; ++k) U.EXLAG[k] = 0.f;
  U.EAJ = 0.f; U.ETO = 0.f; U.ABST = 0.f; U.ETT = U.ESS = U.EP = U.ES = 0.f; U.PINF = 0.f;
  U.RAIN = U.SMELT = U.GMRO = 0.f;
  // HYDRO-17 constants (SNOCOEF, F:3659-3680)
  const float XMFMAX = 5.2f, XMFMIN = 2.0f, XNMF = 0.6f, TIPM = 0.3f, TBASE = 0.0f, PXTEMP = 0.0f,
              PLWHC = 0.01f, GM = 0.5f, PLQW = 0.01f;
  const float UADJ = 0.262f * (1.0f + 0.24f * WIND);
  const float SFNEW = 1.5f * 24.f, RFMIN = 0.25f * 24.f;

  // label 1060: the spin-up pass (IPRE < 2: one year, F:986-996), then the reported pass.  The pass
  // and year loops are block-uniform -- a block runs the spin-up year iff one of its cells needs it --
  // and a thread that is not part of a pass (`act == false`) only keeps the day barrier.
  const bool my_spin = live && IPRE < 2;
  const bool blk_spin = __syncthreads_or(my_spin ? 1 : 0) != 0;
  for (int pass = blk_spin ? 0 : 1; pass < 2; ++pass) {
    const bool act = live && (pass == 1 || my_spin);
    // SNOCOEF (F:3636-3751)
    if (!act) {
    } else if (IPRE < 2) {
      U.TWE = U.WE = U.XNEGHS = U.XLIQW = U.TINDEX = U.STORGE = 0.f;
      for (int k = 1; k <= 4; ++k) U.EXLAG[k] = 0.f;
      if (U.OSNO > 0.f && IPRE == 1) { U.XLIQW = (U.OSNO * 25.4f) * PLQW; U.WE = (U.OSNO * 25.4f) - U.XLIQW; U.SNO = U.OSNO; }
      else U.SNO = 0.0f;
      U.OLDSNO = U.OSNO;
    } else if (IPRE == 3) {
      // (IPRE=1 input: user-specified moisture; snow pack rescaled to OSNO, F:3705-3744)
      if (U.OSNO > 0.0f && U.OLDSNO > 0.0f) {
        U.WE = U.WE * U.OSNO / U.OLDSNO; U.XNEGHS = U.XNEGHS * U.OSNO / U.OLDSNO;
        U.XLIQW = U.XLIQW * U.OSNO / U.OLDSNO; U.STORGE = U.STORGE * U.OSNO / U.OLDSNO;
        for (int k = 1; k <= 4; ++k) U.EXLAG[k] = U.EXLAG[k] * U.OSNO / U.OLDSNO;
      } else {
        U.TWE = U.WE = U.XNEGHS = U.XLIQW = U.TINDEX = U.STORGE = 0.f;
        for (int k = 1; k <= 4; ++k) U.EXLAG[k] = 0.f;
        if (U.OSNO > 0.0f) { U.XLIQW = (U.OSNO * 25.4f) * PLQW; U.WE = (U.OSNO * 25.4f) - U.XLIQW; }
      }
      U.SNO = U.OSNO; U.OLDSNO = U.OSNO;
    } else {
      U.OSNO = U.OLDSNO; U.SNO = U.OLDSNO;
    }
    // SETUPS second pass (F:2673-2741)
    if (act && IPRE >= 2) {
      U.ADDRUN = 0.0f;
      for (int k = 2; k <= 6; ++k) EXWAT[k] = 0.0f;
      if (IPRE == 3) for (int J = 1; J <= NSEG; ++J) SWa(J) = (double)img.SF(S_SW0, J, s);
      for (int J = 1; J <= NSEG; ++J) BALa(J) = 0.0;
    }
    const bool report = (pass == 1);                    // == (IPRE >= 2) for the threads of this pass
    const int dsel = (A.daily && report && act) ? A.daily_row[cell] : -1;
    int64_t drow = 0;
    bool it_snow = false;
    const int npass_years = report ? NY : 1;
    for (int iy = 0; iy < npass_years; ++iy) {           // DO 1090
      const int year = A.years[iy];
      const int NT = (year % 4 != 0) ? 0 : 1;             // LEAP (F:1048)
      const int ND = 365 + NT;
      float PREM = 0.f, RUNM = 0.f, ETM = 0.f, DRNM[6] = {0, 0, 0, 0, 0, 0}, PRCM[7] = {0, 0, 0, 0, 0, 0, 0};
      float ysum[5] = {0, 0, 0, 0, 0};
      int MO = 1, mend = 31;
      const float* __restrict__ ypre = wpre + (int64_t)iy * kDays;
      const float* __restrict__ ytmp = wtmp + (int64_t)iy * kDays;
      const float* __restrict__ yrad = wrad + (int64_t)iy * kDays;
      for (int IDA = 1; IDA <= ND; ++IDA) {               // DO 1100
#ifdef HELP_SIM_DAYSYNC
        __syncthreads();      // keep the block's warps in the same code region (instruction cache); reached by every thread of the block
#endif
        if (!act) continue;
        const float PRE = __ldg(ypre + IDA - 1), TMPF = __ldg(ytmp + IDA - 1), RAD = __ldg(yrad + IDA - 1);
        const float RH = RHq[(IDA <= 91) ? 0 : (IDA <= 182) ? 1 : (IDA <= 274) ? 2 : 3];
        PREM = PREM + PRE;
        // ---------------- FRZCHK (F:3590-3628) ----------------
        {
          U.TMPSUM = U.TMPSUM - U.TAVG[U.ring] + TMPF;
          const float TMPAVG = U.TMPSUM / 30;
          U.TAVG[U.ring] = TMPF;
          U.ring = (U.ring == 29) ? 0 : U.ring + 1;
          if (TMPF < TFSOIL && U.KCNT > 0) U.KCNT = U.KCNT - 1;
          if (TMPF >= TFSOIL && U.KCNT < U.MXKCNT) U.KCNT = U.KCNT + 1;
          if (U.IFREZ == 0) {
            if (TMPAVG < TFSOIL && U.KCNT == 0) { U.IFREZ = 1; U.MXKCNT = U.IDFS; }
          } else if (U.KCNT >= U.IDFS && U.SNO <= 0.0f) {
            U.IFREZ = 0; U.MXKCNT = (U.IDFS + 2) / 3; U.KCNT = U.MXKCNT;
          }
        }
        // ---------------- SNOW (F:3756-4122) ----------------
        {
          const float TA = (TMPF - 32.f) * 0.55556f;
          const float PXI = PRE * 25.4f;
          float SFALL = 0.f, CNHSPX = 0.f, CNHS = 0.f, ROS = 0.f, RAINM = 0.f, XMELT = 0.f, GMWLOS = 0.f, GMSLOS = 0.f;
          float RAIN = 0.f, GMRO = 0.f, SMELT = 0.f;
          bool pack = true, gone = false;
          it_snow = (TA < PXTEMP);
          if (TA < PXTEMP) {
            U.XLIQW = U.XLIQW + U.ADDRUN * 25.4f;
            U.SNO = U.SNO + U.ADDRUN;
            U.TWE = U.TWE + U.ADDRUN * 25.4f;
            U.ADDRUN = 0.0f;
            float ts = TA;
            if (ts > 0.0f) ts = 0.0f;
            SFALL = PXI;
            U.WE = U.WE + SFALL;
            if (U.WE == 0.0f) pack = false;
            else { CNHSPX = -ts * SFALL / 160.0f; if (SFALL > SFNEW) U.TINDEX = ts; }
          } else {
            RAIN = PXI;
            if (U.WE == 0.0f) pack = false;
          }
          if (pack) {
            float TR = TA;
            if (TR < 0.0f) TR = 0.0f;
            RAINM = 0.0125f * RAIN * TR;
            float GMC = GM;
            if (U.IFREZ == 1) GMC = 0.0f;
            if (U.WE <= GMC) { GMRO = U.WE + U.XLIQW; gone = true; }
            else {
              GMWLOS = (GMC / U.WE) * U.XLIQW;
              GMSLOS = GMC;
              int IDANG;
              if (ULAT >= 0.0f) IDANG = (ND == 366) ? 81 : 80;
              else IDANG = (ND == 366) ? 264 : 263;
              int IDN = IDA - IDANG;
              if (IDN <= 0) IDN = IDN + ND;
              const float DIFF = (XMFMAX - XMFMIN);
              const float DAYN = (float)IDN;
              float X;
              if (IDN >= 275) X = (DAYN - 275.0f) / (458.0f - 275.0f);
              else if (IDN >= 92) X = (275.0f - DAYN) / (275.0f - 92.0f);
              else X = (91.0f + DAYN) / 183.0f;
              const float sn = r4_sin(DAYN * 2.0f * 3.1416f / 366.0f);
              const float XX = (sn * 0.5f) + 0.5f;
              float ADJMF;
              if (X <= 0.48f) ADJMF = 0.0f;
              else if (X >= 0.70f) ADJMF = 1.0f;
              else ADJMF = (X - 0.48f) / (0.70f - 0.48f);
              const float RMF60 = (XX * ADJMF * DIFF) + XMFMIN;
              float XMF = RMF60;
              if (fabsf(ULAT) < 60.0f) {
                const float RMF50 = (sn * DIFF * 0.5f) + (XMFMAX + XMFMIN) * 0.5f;
                XMF = RMF50;
                if (fabsf(ULAT) > 50.0f) { const float RINT = (fabsf(ULAT) - 50.f) / 10.f; XMF = RMF50 + RINT * (RMF60 - RMF50); }
              }
              const float RATIO = XMF / XMFMAX;
              float TMX = TA - TBASE;
              if (TMX < 0.0f) TMX = 0.0f;
              float TSUR = TA;
              if (TSUR > 0.0f) TSUR = 0.0f;
              const float TNMX = U.TINDEX - TSUR;
              const float XNMRATE = RATIO * XNMF;
              CNHS = XNMRATE * TNMX;
              U.TINDEX = U.TINDEX + TIPM * (TA - U.TINDEX);
              if (U.TINDEX > 0.0f) U.TINDEX = 0.0f;
              if (TMX > 0.0f) XMELT = XMF * TMX;
              if (RAIN <= RFMIN) XMELT = XMELT + RAINM;
              else {
                const float SVP1 = r4_pow8(0.00738f * TA + 0.8072f);
                const float SVP2 = 0.000019f * fabsf(1.8f * TA + 48.f);
                const float SVP = 33.8639f * (SVP1 - SVP2 + 0.001316f);
                const float AVP = 0.9f * SVP;
                const float TK = (TA + 273.f);
                const float tk2 = TK * TK;
                const float QN = (11.71E-8f * (tk2 * tk2)) / 7.97f - 81.61f;
                const float QE = 8.5f * (AVP - 6.11f) * UADJ;
                const float BR = 0.68f * TA / (AVP - 6.11f);
                const float QH = BR * QE;
                XMELT = QN + QE + QH + RAINM;
                if (XMELT < 0.0f) XMELT = 0.0f;
              }
              ROS = RAIN;
              RAIN = 0.f;
              if ((CNHS + U.XNEGHS) < 0.0f) CNHS = -1.0f * U.XNEGHS;
              U.WE = U.WE - GMSLOS;
              U.XLIQW = U.XLIQW - GMWLOS;
              GMRO = GMSLOS + GMWLOS;
              if (XMELT > 0.0f) {
                if (XMELT >= U.WE) { XMELT = U.WE + U.XLIQW; gone = true; }
                else U.WE = U.WE - XMELT;
              }
              if (!gone) {
                const float WATER = XMELT + ROS;
                const float HEAT = CNHS + CNHSPX;
                const float XLIQWMX = PLWHC * U.WE;
                float EXCESS;
                U.XNEGHS = U.XNEGHS + HEAT;
                if (U.XNEGHS < 0.0f) U.XNEGHS = 0.0f;
                if (U.XNEGHS > 0.33f * U.WE) U.XNEGHS = 0.33f * U.WE;
                if ((WATER + U.XLIQW) >= (XLIQWMX + U.XNEGHS)) {
                  EXCESS = WATER + U.XLIQW - XLIQWMX - U.XNEGHS;
                  U.XLIQW = XLIQWMX; U.WE = U.WE + U.XNEGHS; U.XNEGHS = 0.0f;
                } else if (WATER >= U.XNEGHS) {
                  U.XLIQW = U.XLIQW + WATER - U.XNEGHS; U.WE = U.WE + U.XNEGHS; U.XNEGHS = 0.0f; EXCESS = 0.0f;
                } else {
                  U.WE = U.WE + WATER; U.XNEGHS = U.XNEGHS - WATER; EXCESS = 0.0f;
                }
                if (U.XNEGHS == 0.0f) U.TINDEX = 0.0f;
                float PACKRO = 0.0f;
                const float FIT = 24.f;
                if (EXCESS != 0.0f) {
                  if (EXCESS >= 0.1f && U.WE >= 1.0f) {
                    int Nl = (int)((r4_pow(EXCESS * 4.0f, 0.3f)) + 0.5f);
                    if (Nl == 0) Nl = 1;
                    const float FN = (float)Nl;
                    for (int I = 1; I <= Nl; ++I) {
                      const float FI = (float)I;
                      float TERM = 0.03f * U.WE * FN / (EXCESS * (FI - 0.5f));
                      if (TERM > 50.0f) TERM = 50.0f;
                      const float FLAG = 5.33f * (1.0f - r4_exp(-TERM));
                      int L2 = (int)((FLAG + FIT) / FIT + 1.0f);
                      if (L2 > 4) L2 = 4;          // cannot exceed 2 (FLAG <= 5.33); guards the array
                      const int L1 = L2 - 1;
                      const float ENDL1 = (float)(L1 * 24);
                      const float POR2 = (FLAG + FIT - ENDL1) / FIT;
                      const float POR1 = 1.0f - POR2;
                      U.EXLAG[L2] = U.EXLAG[L2] + POR2 * EXCESS / FN;
                      U.EXLAG[L1] = U.EXLAG[L1] + POR1 * EXCESS / FN;
                    }
                  } else U.EXLAG[1] = U.EXLAG[1] + EXCESS;
                }
                if ((U.STORGE + U.EXLAG[1]) != 0.0f) {
                  if ((U.STORGE + U.EXLAG[1]) < 0.1f) { PACKRO = U.STORGE + U.EXLAG[1]; U.STORGE = 0.0f; }
                  else {
                    const float EL = U.EXLAG[1] / FIT;
                    const float ELS = EL / 25.4f;
                    const float WES = U.WE / 25.4f;
                    float TERM = 500.0f * ELS / (r4_pow(WES, 1.3f));
                    if (TERM > 50.0f) TERM = 50.0f;
                    const float R1 = 1.0f / (5.0f * r4_exp(-TERM) + 1.0f);
                    for (int I = 1; I <= 24; ++I) {
                      const float OS = (U.STORGE + EL) * R1;
                      PACKRO = PACKRO + OS;
                      U.STORGE = U.STORGE + EL - OS;
                    }
                    if (U.STORGE <= 0.001f) { PACKRO = PACKRO + U.STORGE; U.STORGE = 0.0f; }
                  }
                }
                for (int I = 2; I <= 4; ++I) U.EXLAG[I - 1] = U.EXLAG[I];
                U.EXLAG[4] = 0.0f;
                SMELT = PACKRO;
              }
            }
            if (gone) {                                     // label 350: the pack is gone
              float TEX = 0.0f;
              for (int k = 1; k <= 4; ++k) TEX = TEX + U.EXLAG[k];
              SMELT = XMELT + TEX + U.STORGE + ROS;
              U.WE = 0.0f; U.XNEGHS = 0.0f; U.XLIQW = 0.0f; U.TINDEX = 0.0f; U.STORGE = 0.0f;
              for (int k = 1; k <= 4; ++k) U.EXLAG[k] = 0.0f;
            }
          }
          float TEX = 0.0f;                                   // label 380
          for (int k = 1; k <= 4; ++k) TEX = TEX + U.EXLAG[k];
          U.TWE = U.WE + U.XLIQW + U.STORGE + TEX;
          U.GMRO = GMRO / 25.4f;
          U.SMELT = SMELT / 25.4f;
          U.RAIN = RAIN / 25.4f + U.SMELT;
          U.SNO = U.TWE / 25.4f;
        }
        // ---------------- RUNOFF (F:2752-2796) ----------------
        {
          const float SMX = (U.IFREZ > 0) ? SMXF : SMXN;
          if (U.RAIN > 0.0f) {
            float WTSM = 0.0f;
#pragma unroll
            for (int I = 1; I <= 7; ++I) {
              const float swul = (float)SWa(I);
              float sw = swul;
              const float SWAMCI = (FCf[I] + WPf[I]) / 2.0f;
              if (swul > ULf[I]) sw = ULf[I];
              if (swul < SWAMCI) sw = SWAMCI;
              WTSM = WTSM + (WF[I] * (sw - SWAMCI) / (ULf[I] - SWAMCI));
            }
            float S = SMX * (1.0f - WTSM);
            if (S < 0.0f) S = 0.0f;
            const float AB = 0.2f * S;
            if (U.RAIN > AB) { const float d = U.RAIN - AB; U.RUN = (d * d) / (U.RAIN + 0.8f * S); U.RUN = U.RUN * FRUNOF; }
            else U.RUN = 0.0f;
          } else U.RUN = 0.0f;
        }
        // ---------------- SOLT, segment 7 only (F:3516-3560) ----------------
        {
          if (U.CV < 0.0f) U.CV = 0.0f;
          if (U.CV > 20000.f) U.CV = 20000.f;
          float BCV = U.CV / (U.CV + r4_exp(7.563f - (0.0001297f * U.CV)));
          if ((U.SNO * 25.4f) > 120.f) BCV = fmaxf(1.f, BCV);
          else if (U.SNO > 0.0f) {
            const float XX = (U.SNO * 25.4f) / (U.SNO * 25.4f + r4_exp(6.055f - (.3002f * U.SNO * 25.4f)));
            BCV = fmaxf(XX, BCV);
          }
          const float XI = (float)IDA;
          const float ALX = (XI - 200.f) / 58.13f;
          const float TMPC = (TMPF - 32.f) / 1.8f;
          const float ST0 = (RAD == 0.f) ? (WFT5 + TMPC) : (WFT5 + TMPC - 5.f);
          const float XX = BCV * TO + (1.f - BCV) * ST0;
          const float DTs = XX - (AVT + AMP * r4_cos(ALX));
          const float ZD = -Z7 / DD;
          if (fabsf(ZD) < 37.f) U.TSEG7 = AVT + (AMP * r4_cos(ALX + ZD) + DTs) * r4_exp(ZD);
          else U.TSEG7 = AVT;
        }
        // ---------------- EVAPOT + POTET (F:3054-3436) ----------------
        {
          U.RAIN = U.RAIN - U.SMELT;
          float ESNO = 0.f, EABS = 0.f, ESM = 0.f, EGM = 0.f, EMELT = 0.f, EADD = 0.f;
          U.ESS = 0.f; U.EP = 0.f; U.ES = 0.f; U.ABST = 0.f;
          const float TA = (TMPF - 32.f) * 0.55556f;
          // POTET
          float ETO1, ETO2;
          {
            if (U.CV > 40000.f) U.CV = 40000.f;
            U.EAJ = r4_exp(-0.000029f * (U.CV + .1f));
            float ALB;
            if (U.SNO <= 5.f / 25.4f) { ALB = 0.23f; if (U.XLAI > 0.0f) ALB = .23f * (1.f - U.EAJ) + 0.23f * U.EAJ; }
            else ALB = .6f;
            const float TK = ((TMPF - 32.0f) * 5.0f / 9.0f) + 273.0f;
            const float TC = TK - 273.f;
            const float base = 0.00738f * TC + 0.8072f;
            const float SVP = 33.8639f * (r4_pow8(base) - 0.000019f * fabsf(1.8f * TC + 48.f) + 0.001316f);
            const float AVP = (RH / 100.f) * SVP;
            const float DELTA = 33.8639f * (0.05904f * r4_pow7(base) - 0.0000342f);
            const float RBO = 11.71E-8f * (0.39f - 0.05f * r4_sqrt(AVP)) * r4_pow4(TK);
            const float XLAT = ULAT * 6.2832f / 360.f;
            const float XI = (float)IDA;
            const float SD = 0.4102f * r4_sin(0.0172f * (XI - 80.25f));
            const float CH = -r4_tan(XLAT) * r4_tan(SD);
            float H = 0.0f;
            if (CH <= 1.0f && CH >= -1.0f) H = r4_acos(CH);
            else if (CH > 1.0f) H = 0.0f;
            else if (CH < -1.0f) H = 3.1416f;
            const float DDD = 1.0f + 0.0335f * r4_sin(0.0172f * (XI + 88.2f));
            float RSO = 889.2305f * DDD * ((H * r4_sin(XLAT) * r4_sin(SD)) + (r4_cos(SD) * r4_cos(XLAT) * r4_sin(H)));
            RSO = RSO * 0.8f;
            float ACOEF, BCOEF;
            if (RH < 50.f) { ACOEF = 1.2f; BCOEF = -0.2f; }
            else if (RH < 75.f) { ACOEF = 1.1f; BCOEF = -0.1f; }
            else { ACOEF = 1.0f; BCOEF = 0.0f; }
            const float RB = (ACOEF * RAD / RSO + BCOEF) * RBO;
            float RN = (1.f - ALB) * RAD - RB;
            if (RN < 0.f) RN = 0.f;
            const float WINDF = 1.0f + 0.2394f * WIND;
            const float TERM1 = (DELTA / (DELTA + 0.68f)) * RN;
            const float TERM2 = (0.68f / (0.68f + DELTA)) * 15.36f * WINDF * (SVP - AVP);
            U.ETO = (TERM1 + TERM2) / (58.3f * 25.4f);
            ETO1 = TERM1 / (58.3f * 25.4f);
            ETO2 = TERM2 / (58.3f * 25.4f);
            if (U.ETO < 0.0f) U.ETO = 0.0f;
          }
          float TDEST = TA;
          if (PRE == 0.0f) TDEST = (r4_pow(RH / 100.f, 1.f / 8.f)) * (112.f + (0.9f * TA)) - 112.f + (0.1f * TA);
          bool done = false;
          if (U.SNO > 0.0f && TDEST >= U.TINDEX) {
            U.PINF = U.RAIN + U.SMELT + U.ADDRUN + U.GMRO - U.RUN;
            if (U.PINF < 0.0f) { U.RUN = U.RUN + U.PINF; if (U.RUN < 0.0f) U.RUN = 0.0f; }
            U.ES1T = U.ES1T - U.PINF;
            if (U.ES1T <= 0.0f) U.ES1T = 0.0f;
            done = true;
          }
          if (!done) {
            float eta = U.ETO;
            if (U.SNO > 0.0f || U.SMELT > 0.0f) {
              if (U.SMELT > 0.0f) {
                ESM = ESM + (0.1367f * U.SMELT) + (0.0240f * U.SMELT);
                eta = U.ETO - ESM;
                if (ESM > U.ETO) { eta = 0.0f; ESM = U.ETO; }
                EMELT = eta;
                if (EMELT > U.SMELT) { EMELT = U.SMELT; eta = eta - U.SMELT; U.SMELT = 0.0f; }
                else { eta = 0.0f; U.SMELT = U.SMELT - EMELT; }
              } else { EMELT = 0.0f; eta = U.ETO; }
              ESNO = 0.8615f * eta;
              if (ESNO >= U.SNO) {
                ESNO = U.SNO; U.SNO = 0.0f;
                ESM = (ESNO * (0.1367f + 0.0240f)) + ESM;
                eta = eta - (ESNO / 0.8615f);
                U.WE = 0.f; U.XNEGHS = 0.0f; U.TINDEX = 0.0f;
                for (int k = 1; k <= 4; ++k) U.EXLAG[k] = 0.0f;
                U.XLIQW = 0.f; U.STORGE = 0.f;
              } else {
                const float PRSNO = U.SNO;
                U.SNO = U.SNO - ESNO;
                ESM = (ESNO * (0.1367f + 0.0240f)) + ESM;
                eta = 0.0f;
                U.WE = U.SNO / PRSNO * U.WE;
                U.XNEGHS = U.SNO / PRSNO * U.XNEGHS;
                U.XLIQW = U.SNO / PRSNO * U.XLIQW;
                U.STORGE = U.SNO / PRSNO * U.STORGE;
                float TEX = 0.0f;
                for (int k = 1; k <= 4; ++k) { U.EXLAG[k] = U.SNO / PRSNO * U.EXLAG[k]; TEX = TEX + U.EXLAG[k]; }
                U.TWE = U.WE + U.XLIQW + U.STORGE + TEX;
              }
            } else {
              if (U.RAIN > 0.0f) {
                float ABSMAX = 0.05f;
                if (U.CV < 14000.f) ABSMAX = 0.05f * (U.CV + .1f) / 14000.f;
                if ((U.RAIN / ABSMAX) < 88.f) U.ABST = ABSMAX * (1.f - r4_exp(-U.RAIN / ABSMAX));
                else U.ABST = ABSMAX;
              }
              EABS = eta;
              if (EABS >= U.ABST) { EABS = U.ABST; eta = eta - U.ABST; }
              else { U.ABST = EABS; eta = 0.0f; }
            }
            if (U.ADDRUN > 0.0f) {
              EADD = eta;
              if (EADD > U.ADDRUN) { EADD = U.ADDRUN; eta = eta - EADD; U.ADDRUN = 0.0f; }
              else { U.ADDRUN = U.ADDRUN - EADD; eta = 0.0f; }
            }
            if (U.GMRO > 0.0f) {
              EGM = eta;
              if (EGM > U.GMRO) { EGM = U.GMRO; eta = eta - EGM; U.GMRO = 0.0f; }
              else { U.GMRO = U.GMRO - EGM; eta = 0.0f; }
            }
            U.PINF = U.RAIN + U.SMELT + U.ADDRUN + U.GMRO - U.RUN - U.ABST;
            if (U.PINF < 0.0f) {
              U.RUN = U.RUN + U.PINF;
              U.PINF = 0.0f;
              if (U.RUN < 0.0f) { U.ABST = U.ABST + U.RUN; U.RUN = 0.0f; EABS = U.ABST; }
            }
            U.ESS = EMELT + ESNO + EABS + EADD + EGM;
            if (eta <= 0.0f || (U.ETO - U.ESS - ESM) <= 0.0f) {
              U.ES1T = U.ES1T - U.PINF;
              if (U.ES1T <= 0.0f) U.ES1T = 0.0f;
            } else if (TMPF >= 23.0f && U.IFREZ != 1) {
              float ETO2X = (1.0f - (7.143E-05f * U.CV)) * ETO2;
              if (ETO2X < 0.0f) ETO2X = 0.0f;
              float ESO = U.EAJ * (ETO1 + ETO2X);
              if (ESO > eta) ESO = eta;
              if (ESO < 0.0f) ESO = 0.0f;
              if (U.ES1T <= STAGE1) {
                float ES1 = eta;
                if (ES1 >= ESO) { ES1 = ESO; eta = eta - ESO; }
                else if (ES1 <= 0.0f) { ES1 = 0.0f; eta = 0.0f; }
                else eta = eta - ES1;
                U.ES1T = U.ES1T + ES1 - U.PINF;
                if (U.ES1T <= 0.0f) U.ES1T = 0.0f;
                if (U.ES1T <= STAGE1) U.TS2 = 0.0f;
                U.ES = ES1;
              } else {
                U.TS2 = U.TS2 + 1.0f;
                float ES2 = CONA * (r4_sqrt(U.TS2) - r4_sqrt(U.TS2 - 1.0f)) / 25.4f;
                if (ES2 >= ESO) { ES2 = ESO; eta = 0.0f; }
                else if (ES2 <= 0.0f) ES2 = 0.0f;
                else eta = eta - ES2;
                U.ES1T = U.ES1T + ES2 - U.PINF;
                if (U.ES1T <= 0.0f) U.ES1T = 0.0f;
                U.ES = ES2;
              }
              U.EP = U.ETO * U.XLAI / 3.0f;
              if (U.EP > eta) { U.EP = eta; eta = 0.0f; }
              else if (U.EP < 0.0f) U.EP = 0.0f;
              else eta = eta - U.EP;
              if (U.EP >= (U.ETO - ESM - U.ESS - U.ES)) U.EP = U.ETO - ESM - U.ES - U.ESS;
              if (U.EP < 0.0f) { U.ES = U.ES + U.EP; U.EP = 0.0f; }
            } else {
              U.ES1T = U.ES1T - U.PINF;
              if (U.ES1T <= 0.0f) U.ES1T = 0.0f;
            }
          }
        }
        // ---------------- CRPMOD / ETCHK / TSTR (F:3440-3582, 2803-2976) ----------------
        {
          const float SUT = U.ST7 * 25.4f / FCS7;
          const float CDG = (.9f * U.TSEG7 / (U.TSEG7 + r4_exp(9.93f - (.312f * U.TSEG7)))) + .1f;
          const float DECR = .05f * fminf(CDG, SUT);
          U.RSD = U.RSD * (1.f - DECR);
          if (IPL == 0 && IDA == 30) U.GS = 0.f;
          if (IPL == 367 && IDA == 210) U.GS = 0.f;
          if (IPL == 0 && IHV == 367) U.DM = U.DM * (1.f - DECR);
          if (IPL == 367 && IHV == 0) U.DM = U.DM * (1.f - DECR);
          U.CV = .8f * U.DM + U.RSD;
          // growing-season state machine (the arithmetic IFs of F:3467-3479)
          int mode;          // 0: dormant, 1: harvest day, 2: growing
          {
            const int dP = IDA - IPL, dH = IDA - IHV;
            if (IPL < IHV) {
              if (dP < 0) mode = 0;
              else if (dP == 0) { U.GS = 0.f; U.DM = 0.f; mode = 2; }          // 1000 -> 1030 -> 1040
              else mode = (dH < 0) ? 2 : (dH == 0 ? 1 : 0);                    // 1020
            } else {
              if (dP < 0) mode = (dH < 0) ? 2 : (dH == 0 ? 1 : 0);             // 1000 -> 1020
              else if (dP == 0) { U.GS = 0.f; U.DM = 0.f; mode = 2; }
              else mode = 2;                                                   // 1040
            }
          }
          if (mode == 1) { U.RSD = (0.53f * U.DM + U.RSD); U.DM = 0.f; U.XLAI = 0.f; }
          // ETCHK
          float WS = 0.f;
          {
            const float TET = U.ES + U.EP;
            if (TET <= 0.0f) {
              U.ETT = U.ESS; WS = 0.f;
#pragma unroll
              for (int J = 1; J <= 7; ++J) U.ET[J] = 0.0f;
            } else {
              float ESJ[8], ULL[8], EPJ[8], AW[8], XP[8], ST[8];
              float ESX = U.ES, XES = 0.0f, THCK = 0.0f, OTHCK = 0.0f, DRINF = U.PINF, XPINF;
#pragma unroll
              for (int J = 1; J <= 7; ++J) {
                XP[J] = 0.0f; ESJ[J] = 0.0f;
                ULL[J] = FCf[J] - WPf[J];
                // SWUL(J) is the REAL*4 mirror of DSWUL(J) (F:666-668); DCHG(J) == BALa(J)
                ST[J] = (float)((double)(float)SWa(J) + (BALa(J) / 2.0) - (double)WPf[J] + (double)DRINF);
                if (ST[J] < 0.0f) ST[J] = 0.0f;
                if (ST[J] > ULf[J]) { XPINF = ST[J] - ULf[J]; ST[J] = ULf[J]; }
                else XPINF = 0.0f;
                if (OTHCK < SEDEP) {
                  if (XPINF > RCf[J]) {
                    XPINF = XPINF - RCf[J];
                    if (XPINF < ESX) { ESX = ESX - XPINF; ESJ[J] = XPINF; XP[J] = XPINF; }
                    else { ESJ[J] = ESX; XP[J] = ESX; ESX = 0.0f; }
                    DRINF = RCf[J];
                  } else DRINF = XPINF;
                  if (ESX > 0.0f) {
                    THCK = THCK + STH[J];
                    float ESJO = wfs(OTHCK, THCK, SEDEP) * U.ES + XES;
                    if (ESJO > ESX) ESJO = ESX;
                    if (THCK <= SEDEP) {
                      if (ST[J] < ESJO) { XES = ESJO - ST[J]; ESX = ESX - ST[J]; ESJ[J] = ST[J] + ESJ[J]; }
                      else { XES = 0.0f; ESX = ESX - ESJO; ESJ[J] = ESJO + ESJ[J]; }
                    } else {
                      const float RTO = ST[J] * (SEDEP - OTHCK) / (STH[J]);
                      if (RTO < ESJO) { XES = ESJO - RTO; ESX = ESX - RTO; ESJ[J] = RTO + ESJ[J]; }
                      else { XES = 0.0f; ESX = ESX - ESJO; ESJ[J] = ESJO + ESJ[J]; }
                      U.ES = U.ES - ESX;
                      U.ES1T = U.ES1T - ESX;
                      if (U.ES1T <= 0.0f) U.ES1T = 0.0f;
                      ESX = 0.0f;
                    }
                  }
                  OTHCK = THCK;
                }
              }
#pragma unroll
              for (int J = 1; J <= 7; ++J) {
                EPJ[J] = 0.0f;
                AW[J] = ST[J] - ESJ[J] + XP[J];
                if (AW[J] < 0.0f) AW[J] = 0.0f;
              }
              if (U.EP > 0.0f) {
                float SUM = 0.0f, EPX = 0.0f, EPXX, UL4;
#pragma unroll
                for (int J = 1; J <= 7; ++J) {
                  EPXX = U.EP * WF[J];
                  EPXX = SUM + EPXX;
                  if (AW[J] > ULL[J]) UL4 = (AW[J] - ULL[J]) + ULL[J] / 4.f;
                  else UL4 = ULL[J] / 4.f;
                  float e;
                  if (UL4 < EPXX) e = (UL4 <= AW[J]) ? UL4 : AW[J];
                  else e = (EPXX < AW[J]) ? EPXX : AW[J];
                  EPJ[J] = e;
                  EPX = EPX + e;
                  SUM = EPXX - e;
                }
                U.EP = EPX;
              }
              U.ETT = U.ESS;
#pragma unroll
              for (int J = 1; J <= 7; ++J) { U.ET[J] = ESJ[J] + EPJ[J]; U.ETT = U.ETT + U.ET[J]; }
              WS = (U.ETT - U.ESS) / TET;
              U.ST7 = ST[7];
            }
          }
          if (mode == 2) {
            const float TMPC = (TMPF - 32.f) / 1.8f;
            float TGX = TMPC - TB;
            if (TGX > 0.f) {                                // TSTR (F:3568-3582)
              if (TMPC > TO) TGX = 2.f * TO - TB - TMPC;
              const float r = (TO - TMPC) / TGX;
              const float RTO = r * r;
              U.TS = (RTO > 200.f) ? 0.f : r4_exp(-0.1054f * RTO);
              if ((TMPC - 5.f) <= (AVT - 15.f)) U.TS = 0.f;
            } else U.TS = 0.f;
            const float DDM = .5f * RAD * (1.f - r4_exp((-.65f) * (U.XLAI + .05f)));
            const float REG = fminf(WS, U.TS);
            U.DM = U.DM + DDM * REG;
            if (U.DM <= 0.0f) U.DM = 0.0f;
            if (TMPC > TB) U.GS = U.GS + (TMPC - TB) / PHU;
            if (U.GS > .75f) U.XLAI = 16.f * U.DLAI * (1.f - U.GS) * (1.f - U.GS);
            else {
              const float WLV = .8f * U.DM;
              const float Fv = WLV / (WLV + 5512.f * r4_exp((-0.000608f) * WLV));
              U.XLAI = ULAI * Fv;
              U.DLAI = U.XLAI;
            }
          }
          if (mode != 0 && U.XLAI < 0.0f) U.XLAI = 0.0f;
        }
        // ---------------- subsurface routing, subprofile 1 (F:631-693) ----------------
        double DRN[7], PRC[7];
        double HEDd[4] = {0.0, 0.0, 0.0, 0.0}, DRNd[4] = {0.0, 0.0, 0.0, 0.0};   // daily channel only
        double RCRn[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};                          // RCR(n): today's recirculated drainage
        U.ADDRUN = 0.0f;
#ifdef HELP_DBG_PRE
        HELP_DBG_PRE
#endif
        {
          const RouteOut R = (PR[1].subin & 2) ? route_subprofile<MAXSEG, true>(img, s, PR[1], U.PINF, U.ET, U.IFREZ, SS)
                                               : route_subprofile<MAXSEG, false>(img, s, PR[1], U.PINF, U.ET, U.IFREZ, SS);
          DRN[1] = R.DRN; PRC[1] = R.PRC; U.ADDRUN = R.EXTRA; flag |= R.flag; HEDd[1] = R.HED; DRNd[1] = R.DRN;
        }
        {
          float EXET = 0.0f;
#pragma unroll
          for (int J = 1; J <= 7; ++J) {
            const double wp = (double)WPf[J];
            if ((SWa(J) + (BALa(J) / 2.0)) < wp) {
              const float EXET2 = (float)((double)EXET + wp - SWa(J) - (BALa(J) / 2.0));
              if (EXET <= U.ETT) { SWa(J) = wp - (BALa(J) / 2.0); EXET = EXET2; }
            }
          }
          U.RUN = U.RUN + U.ADDRUN * FRUNOF;
          U.ADDRUN = U.ADDRUN * (1.0f - FRUNOF);
          if (U.SNO > 0.0f) {
            U.SNO = U.SNO + U.ADDRUN;
            U.TWE = U.TWE + U.ADDRUN * 25.4f;
            U.XLIQW = U.XLIQW + U.ADDRUN * 25.4f;
            U.ADDRUN = 0.0f;
          }
          RUNM = RUNM + U.RUN;
          U.ETT = U.ETT - EXET;
          ETM = ETM + U.ETT;
        }
        if (LDn[1] > 0) {
          const double RCR = (double)RECIR[1] * DRN[1] / (double)100.0f;
          DRN[1] = DRN[1] - RCR;
          DRNd[1] = DRN[1] + RCR;
          RCRn[1] = RCR;
        }
        DRNM[1] = (float)((double)DRNM[1] + DRN[1]);
        PRCM[1] = (float)((double)PRCM[1] + PRC[1]);
        for (int N = 2; N <= 6; ++N) {                          // F:694-884
          const float FINn = (float)(PRC[N - 1] + (double)EXWAT[N]);
          if (N > nprof) break;
          EXWAT[N] = 0.0f;
          {
            const RouteOut R = (PR[N].subin & 2) ? route_subprofile<MAXSEG, true>(img, s, PR[N], FINn, U.ET, U.IFREZ, SS)
                                                 : route_subprofile<MAXSEG, false>(img, s, PR[N], FINn, U.ET, U.IFREZ, SS);
            DRN[N] = R.DRN; PRC[N] = R.PRC; EXWAT[N] = R.EXTRA; flag |= R.flag;
            if (N <= 3) { HEDd[N] = R.HED; DRNd[N] = R.DRN; }
          }
          if (N <= 5) {
            if (LDn[N] > 0) {
              const double RCR = (double)RECIR[N] * DRN[N] / (double)100.0f;
              DRN[N] = DRN[N] - RCR;
              if (N <= 3) DRNd[N] = DRN[N] + RCR;
              RCRn[N] = RCR;
            }
            DRNM[N] = (float)((double)DRNM[N] + DRN[N]);
          }
          PRCM[N] = (float)((double)PRCM[N] + PRC[N]);
        }
        // ---------------- RECIRC (F:5311-5440): today's recirculated drainage becomes tomorrow's
        // source term DRCRS(J) of the segments of the receiving layers ----------------
        if (has_rcr) {
          for (int J = 1; J <= NSEG; ++J) {
            const int plan = img.SI2(J, s);
            const int mode = plan & 0xff, K = plan >> 8;
            if (mode == 0) continue;
            double RCRI = 0.0;
            for (int N = 1; N <= 5; ++N) if (RCRT[N] == K) RCRI = RCRI + RCRn[N];
            const float w = img.SF(S_RCW, J, s);
            SS.DRCRS[J] = (mode == 1) ? RCRI * (double)img.SF(S_THICK, J, s) / (double)w
                        : (mode == 2) ? RCRI * (double)w : RCRI;
          }
        }
#ifdef HELP_DBG_DAY
        HELP_DBG_DAY
#endif
        // ---------------- daily channel (OUTDAY columns, F:5984-5996; inches) ----------------
        if (dsel >= 0) {
          float SWULE = 0.0f;
          for (int J = 1; J <= 7; ++J) SWULE = (float)((double)SWULE + SWa(J) + BALa(J) / 2.0);      // F:652
          double* d = A.daily + ((int64_t)dsel * A.daily_days + drow) * kDailyCols;
          d[0] = (double)PRE; d[1] = (double)U.RUN; d[2] = (double)U.ETT; d[3] = (double)(SWULE / EDEPTH);
          for (int N = 1; N <= 3; ++N) {
            const bool on = (N <= nprof);
            d[4 + 3 * (N - 1)] = on ? HEDd[N] : 0.0;
            d[5 + 3 * (N - 1)] = on ? DRNd[N] : 0.0;       // before the recirculated share is taken out (DRN + RCR)
            d[6 + 3 * (N - 1)] = on ? PRC[N] : 0.0;
          }
          d[13] = it_snow ? 1.0 : 0.0; d[14] = (double)U.IFREZ; d[15] = (double)U.SNO;
          ++drow;
        }
        // ---------------- month end: report (OUTMO2 semantics) ----------------
        if (IDA == mend + ((MO >= 2) ? NT : 0) || IDA == ND) {
          if (report) {
            const int m = MO - 1;
            const int64_t o = ((int64_t)iy * 12 + m) * n + cell;
            const int64_t rs = (int64_t)NY * 12 * n;
            if (A.raw) {
              A.raw[0 * rs + o] = PREM; A.raw[1 * rs + o] = RUNM; A.raw[2 * rs + o] = ETM;
              for (int N = 1; N <= 5; ++N) A.raw[(2 + N) * rs + o] = DRNM[N];
              for (int N = 1; N <= 6; ++N) A.raw[(7 + N) * rs + o] = PRCM[N];
            }
            const float precip = f73(PREM * 25.4f, flag), runoff = f73(RUNM * 25.4f, flag),
                        evapo = f73(ETM * 25.4f, flag);
            float sub1 = 0.f, sub2 = 0.f, perco = 0.f, rechg = 0.f;
            bool have1 = false;
            for (int N = 1; N <= nprof; ++N) {
              if (N <= 5 && LDn[N] > 0) {
                const float v = f73(DRNM[N] * 25.4f, flag);
                if (!have1) { sub1 = v; have1 = true; } else sub2 = sub2 + v;
              }
              const float v = (N == 6 && MO >= 7) ? f73(PRCM[N], flag) : f73(PRCM[N] * 25.4f, flag);   // sic F:6593
              if (N == 1) perco = v;
              rechg = v;
            }
            if (A.monthly) {
              A.monthly[0 * rs + o] = precip; A.monthly[1 * rs + o] = runoff; A.monthly[2 * rs + o] = evapo;
              A.monthly[3 * rs + o] = sub1;   A.monthly[4 * rs + o] = sub2;   A.monthly[5 * rs + o] = perco;
              A.monthly[6 * rs + o] = rechg;
            }
            ysum[0] += precip; ysum[1] += runoff; ysum[2] += evapo; ysum[3] += sub1 + sub2; ysum[4] += rechg;
          }
          PREM = RUNM = ETM = 0.f;
          for (int N = 0; N < 6; ++N) DRNM[N] = 0.f;
          for (int N = 0; N < 7; ++N) PRCM[N] = 0.f;
          if (MO < 12) { MO = MO + 1; mend = kICAL[MO]; }
        }
      }   // day
      if (!act) continue;
      if (report && A.yearly) {
        for (int r = 0; r < 5; ++r) A.yearly[((int64_t)r * NY + iy) * n + cell] = ysum[r];
      }
      // year end (F:938-996)
      U.OLDSNO = U.ADDRUN + U.SNO;
      if (IPRE < 2) {               // the spin-up pass is this one year
        IPRE = IPRE + 2;
        for (int J = 1; J <= NSEG; ++J) { SWa(J) = SWa(J) + BALa(J) / 2.0; BALa(J) = 0.0; }
      }
    }   // year
  }
  if (live) A.flags[cell] = flag;
}

}  // namespace helpb200
