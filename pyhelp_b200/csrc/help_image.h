// help_image.h -- device-resident "cell image": everything the daily kernel needs
// about one cell, produced once by the set-up kernel (help_setup.cuh) and read
// by the simulation kernel (help_sim.cuh).
//
// Layout in HBM: structure of arrays, field-major, cell-minor
//     f32[field][cell]   i32[field][cell]   f64[field][cell]
// with `stride` (= n_cells rounded up to 32) between fields, so that the 32
// threads of a warp (32 consecutive cells) read 128 contiguous bytes per
// field.  Per-segment fields are laid out [field][segment][cell] with
// `segcap` segments (the batch's upper bound, <= 67, HELP3O.FOR:70).
//
// Field meaning follows the reference's COMMON blocks (pyhelp/HELP3O.FOR:59-131);
// the name of the Fortran variable is quoted next to each field.
#pragma once
#include <stdint.h>

#ifndef HELP_R4_OVERRIDE
#include "help_math.h"
#endif

namespace helpb200 {

constexpr int kMaxLay = 20;
constexpr int kMaxSeg = 67;
constexpr int kNProf = 6;      // subprofiles (HELP3O.FOR:96-98)
constexpr int kDays = 368;     // row stride of a weather year, floats (16 B aligned rows)

// ---- f32 scalars ------------------------------------------------------------
enum ImgF : int {
  F_EDEPTH = 0,  // EDEPTH after unit conversion and liner clipping (F:1232,1498), in
  F_SEDEP,       // SEDEP  max soil-evaporation depth (F:2296-2309)
  F_SMXN,        // SCS retention parameter, unfrozen (F:169-174)
  F_SMXF,        // ... frozen soil (F:178-185)
  F_FRUNOF,      // FRUNOF/100 (F:1475)
  F_CONA,        // CONA (F:3035-3040)
  F_STAGE1,      // STAGE1 (F:3044)
  F_ULAT, F_ULAI, F_WIND,            // after unit conversion (F:1233)
  F_RH1, F_RH2, F_RH3, F_RH4,        // quarterly RH (F:1234-1245)
  F_PHU, F_AVT, F_AMP, F_DD,         // INITIA (F:2411-2477)
  F_Z7, F_FCS7, F_ST7,               // Z(7), FCS(7), initial ST(7) (F:2458-2465)
  F_TS0, F_XLAI0, F_GS0, F_RSD0, F_DM0,  // vegetation start state (F:2413-2442)
  F_OSNO,        // OSNO in inches (F:1476)
  F_TFSOIL,      // TFSOIL, deg F
  F_SUBINF,      // SUBINF (F:1465-1480)
  F_NSCALAR
};
// ---- f32 per subprofile (6 each) ---------------------------------------------
enum ImgPF : int {
  P_DRCULS = 0,  // DRCUL(NSEGE) liner conductivity, in/day (F:4165)
  P_DTHICS,      // DTHICK(NSEGE) liner (pseudo-)segment thickness (F:4164)
  P_SLOPE,       // SLOPE(LAYD), percent, after SETUPS' zero fix (F:2602,4686)
  P_XLENG,       // XLENG(LAYD), ft (F:2601,4687)
  P_THKLIN,      // THICK(LINER) geomembrane thickness, in (F:4904)
  P_PHOLE, P_DEFEC, P_TRANS,   // PHOLE/DEFEC/TRANS(LINER) (F:1482-1483)
  P_RECIR,       // RECIR(LD(n)), percent of drainage recirculated (F:682)
  P_SUBLIN,      // DSUB source term of the liner segment, DSUBIN(NSEGE) (F:4193)
  P_NF
};
// ---- f32 per segment ----------------------------------------------------------
enum ImgSF : int {
  S_THICK = 0,   // STHICK   in
  S_UL,          // UL       porosity * thickness, in
  S_FC,          // FCUL
  S_WP,          // WPUL
  S_RS,          // RSUL     residual storage
  S_LAM,         // XLAMBU   Brooks-Corey lambda
  S_BUB,         // BUBUL    bubbling pressure
  S_RC,          // RCUL     saturated K, in/day
  S_SUBIN,       // SUBINS   subsurface inflow, in/day
  S_SW0,         // SWUL     initial storage
  S_RELMIN,      // REAL*4 (WPUL - RSUL)/(UL - RSUL): RELMIN of F:4329, a constant of the segment
  S_RCW,         // recirculation: THICK(K) of the receiving layer (mode 1) or the REAL*4 share (mode 2), see SI2
  S_NF
};
// ---- f64 per segment: constants of the routing loop that cost a division or a pow ---------------
// Each is the value the reference's expression takes (same operations, same order), hoisted
// out of the sub-step loop because it depends only on the segment and on DT = 1/NSTEP(n).
enum ImgSD : int {
  SD_C = 0,      // 3 + 2/lambda                                   (F:4420, 4465)
  SD_NIL,        // -1/lambda                                      (F:4334, 4360)
  SD_BC,         // B = K dt (1/(UL-RS))**C of the Newton solve     (F:4466)
  SD_LOWWP,      // LOW = K dt ((WP-RS)/(UL-RS))**C, i.e. SWLIM = WP (F:4420)
  SD_LOWFC,      // LOW at SWLIM = FC
  SD_SUBIN,      // DSUBIN   subsurface inflow, in/day, REAL*8: the inflow of a liner is ADDED to the segment above it in REAL*8 (F:2060-2071)
  SD_NF
};
// ---- f64 per subprofile ---------------------------------------------------------
enum ImgPD : int {
  D_RCLIN = 0,   // RC(LINER), cm/s (F:4905)
  D_RCS,         // RC(LAYSEG(NSEGE,1)) controlling-soil K, cm/s (F:4951)
  D_NF
};
// ---- i32 scalars -------------------------------------------------------------------
enum ImgI : int {
  I_NSEG = 0, I_IPL, I_IHV, I_IDFS, I_NODE_P, I_NODE_T, I_NODE_R, I_FLAGS, I_NPROF, I_IPRE,
  I_NSCALAR
};
// ---- i32 per subprofile --------------------------------------------------------------
enum ImgPI : int {
  Q_NSEGN = 0,   // NSEGn
  Q_NSTEP,       // NSTEP(n) (F:2577-2672)
  Q_IBAR,        // IBAR incl. the +3 subsurface-inflow cases (F:4166-4210)
  Q_LAT,         // LAT (F:4179)
  Q_LCASE,       // LCASE(n) (F:2074-2102)
  Q_LPQ,         // LPQ(n)
  Q_LTYPE,       // LAYER(LP(n)) (F:5292)
  Q_LD,          // LD(n)
  Q_RCRT,        // LAYR(LD(n)): layer the recirculated drainage of subprofile n returns to, 0 = none (F:5325-5340)
  Q_SUBIN,       // bit0: a segment of the subprofile has subsurface inflow (SUBINS > 0);
                 // bit1: "plain" subprofile: IBAR = 0, no inflow, not above a drain layer
  Q_NF
};
// i32 per segment: bit0 = LAYSEG(J,2) < LD(NPROF) (F:4494), bit1 = the same for J+1 (F:4499)

// ---- routing records: what the sub-daily routing loop reads every sub-step ---------------------
// The routing loop (help_sim.cuh route_subprofile) touches ~14 constants of every segment of
// every resident cell 4..288 times per simulated day; this stream, not the weather or the
// results, is the kernel's memory traffic.  Its layout is therefore tile-major and vectorised:
//     rec[tile = cell/32][segment][vector][lane = cell%32]      (16 bytes per lane and vector)
// so that (1) a warp reads one contiguous 512-byte row per vector, (2) every vector of a segment
// sits at a compile-time offset from ONE per-thread pointer that advances by kRecBytes per
// segment, and (3) all loads of a segment can be issued together, one segment ahead of their use.
// REAL*4 members hold the reference's REAL*4 values; the REAL*8 members hold values the reference
// recomputes from them every sub-step (same operations, same order; F:4329-4466), which depend
// only on the segment and on DT = 1/NSTEP(n).
//   RV_A float4  {UL, RSUL, WPUL, FCUL}                       inches of water
//   RV_B float4  {XLAMBU, BUBUL, RELMIN, flags}               RELMIN = REAL*4 (WPUL-RSUL)/(UL-RSUL), F:4329;
//                                                             flags: bit0 = above the drain layer, bit1 = so is J+1 (F:4494,4499)
//   RV_C double2 {1/(UL-RS) correctly rounded, K dt}          (fdiv_by below)
//   RV_D double2 {C = 3 + 2/lambda, -1/lambda}                (F:4420, 4465; F:4334, 4360)
//   RV_E double2 {B = K dt (1/(UL-RS))**C, LOW at SWLIM = WP} (F:4466; F:4420)
//   RV_F double2 {LOW at SWLIM = FC, (STHICK, SUBINS) as two REAL*4}
enum RecV : int { RV_A = 0, RV_B, RV_C, RV_D, RV_E, RV_F, kRecV };
constexpr int kRecBytes = kRecV * 512;
struct alignas(16) RecTail { double lowfc; float thick, subin; };   // RV_F

struct Image {
  float* f32;
  int32_t* i32;
  double* f64;
  int64_t stride;   // cells per field row
  int segcap;
  char* rec;        // routing records, [stride/32][segcap][kRecBytes]

  // per-thread base pointer of cell c's records (vector RV_A of segment 1)
  __host__ __device__ __forceinline__ char* rec_p(int64_t c) const {
    return rec + (c >> 5) * ((int64_t)segcap * kRecBytes) + (c & 31) * 16;
  }
  __host__ __device__ static int64_t rec_bytes(int64_t stride, int segcap) { return (stride / 32) * (int64_t)segcap * kRecBytes; }

  __host__ __device__ static int64_t n_f32(int segcap) { return F_NSCALAR + P_NF * kNProf + (int64_t)S_NF * segcap; }
  __host__ __device__ static int64_t n_i32(int segcap) { return I_NSCALAR + Q_NF * kNProf + 2 * (int64_t)segcap; }
  __host__ __device__ static int64_t n_f64(int segcap) { return D_NF * kNProf + (int64_t)SD_NF * segcap; }

  __device__ __forceinline__ float& F(int f, int64_t c) const { return f32[(int64_t)f * stride + c]; }
  __device__ __forceinline__ float& PF(int f, int n, int64_t c) const {   // n = 1..6
    return f32[(int64_t)(F_NSCALAR + f * kNProf + (n - 1)) * stride + c];
  }
  __device__ __forceinline__ float& SF(int f, int j, int64_t c) const {   // j = 1..segcap
    return f32[(int64_t)(F_NSCALAR + P_NF * kNProf + f * segcap + (j - 1)) * stride + c];
  }
  __device__ __forceinline__ int32_t& I(int f, int64_t c) const { return i32[(int64_t)f * stride + c]; }
  __device__ __forceinline__ int32_t& PI(int f, int n, int64_t c) const {
    return i32[(int64_t)(I_NSCALAR + f * kNProf + (n - 1)) * stride + c];
  }
  // recirculation plan of segment j (RECIRC, F:5392-5436): low byte = mode (0 none, 1 RCRI*DTHICK(j)/THICK(K),
  // 2 RCRI*share, 3 RCRI), next byte = receiving layer K
  __device__ __forceinline__ int32_t& SI2(int j, int64_t c) const {
    return i32[(int64_t)(I_NSCALAR + Q_NF * kNProf + segcap + (j - 1)) * stride + c];
  }
  __device__ __forceinline__ int32_t& SI(int j, int64_t c) const {
    return i32[(int64_t)(I_NSCALAR + Q_NF * kNProf + (j - 1)) * stride + c];
  }
  __device__ __forceinline__ double& PD(int f, int n, int64_t c) const {
    return f64[(int64_t)(f * kNProf + (n - 1)) * stride + c];
  }
  __device__ __forceinline__ double& SD(int f, int j, int64_t c) const {   // j = 1..segcap
    return f64[(int64_t)(D_NF * kNProf + f * segcap + (j - 1)) * stride + c];
  }
};

// vector accessors: p = rec_p(c) advanced to the segment ((J-1) * kRecBytes)
#if defined(__CUDA_ARCH__)
template <int V> __device__ __forceinline__ float4 rec_f4(const char* p) { return __ldg((const float4*)(p + V * 512)); }
template <int V> __device__ __forceinline__ double2 rec_d2(const char* p) { return __ldg((const double2*)(p + V * 512)); }
// bring the vectors of a segment into L1 without tying up registers
__device__ __forceinline__ void rec_prefetch(const char* p) {
#pragma unroll
  for (int v = 0; v <= RV_E; ++v) asm volatile("prefetch.global.L1 [%0];" ::"l"(p + v * 512));
}
#else
template <int V> inline float4 rec_f4(const char* p) { return *(const float4*)(p + V * 512); }
inline void rec_prefetch(const char*) {}
template <int V> inline double2 rec_d2(const char* p) { return *(const double2*)(p + V * 512); }
#endif
__host__ __device__ __forceinline__ RecTail rec_tail(const char* p) { return *(const RecTail*)(p + RV_F * 512); }

// cell status bits (mirrors include/help_b200.h)
constexpr uint32_t kFlagNonConv = 0x1u, kFlagNaN = 0x2u, kFlagOverflow = 0x4u, kFlagRecirc = 0x8u,
                   kFlagBadCell = 0x10u;

// ---- transcendental functions ------------------------------------------------------------
// The reference's surface-process code is REAL*4 and its routing code REAL*8; both get
// their transcendentals from libm.  HELP is chaotic in the last bit of those functions
// (DESIGN.md section 4), so the engine does not use CUDA's libdevice: it runs the operation
// sequences of the libm the reference links in this image (help_math_glibc.h: glibc 2.39,
// x86-64, FMA variants), the same source the CPU oracle is built from:
//   r4_*(x): REAL*4 intrinsic = glibc's expf/logf/log10f/sinf/cosf/tanf/acosf/powf, bit for bit
//   r8_*(x): REAL*8 intrinsic = glibc's pow/log/sin/cos/atan, bit for bit (sqrt: IEEE)
#ifndef HELP_R4_OVERRIDE   // (scripts/host_emu swaps in glibc's libm to separate logic from libm differences)
using helpmath::r4_exp; using helpmath::r4_log; using helpmath::r4_log10; using helpmath::r4_sin;
using helpmath::r4_cos; using helpmath::r4_tan; using helpmath::r4_acos; using helpmath::r4_pow;
using helpmath::r4_sqrt; using helpmath::r4_pow8; using helpmath::r4_pow7; using helpmath::r4_pow4;
__host__ __device__ __forceinline__ double r8_pow(double x, double y) { return helpmath::det_pow(x, y); }
__host__ __device__ __forceinline__ void r8_pow2(double x0, double y0, double x1, double y1, double& r0, double& r1) {
  const helpmath::Pair R = helpmath::det_pow2(x0, y0, x1, y1);   // two independent powers, same bits as r8_pow
  r0 = R.a; r1 = R.b;
}
__host__ __device__ __forceinline__ void r8_pow_sb(double x, double y0, double y1, double& r0, double& r1) {
  const helpmath::Pair R = helpmath::det_pow_samebase(x, y0, y1);   // x**y0 and x**y1, same bits as two r8_pow calls
  r0 = R.a; r1 = R.b;
}
__host__ __device__ __forceinline__ void r8_log_pair(double x, double& hi, double& lo) { helpmath::det_log_pair(x, hi, lo); }
__host__ __device__ __forceinline__ double r8_pow_logbase(double x, double hi, double lo, double y) {
  return helpmath::det_pow_logbase(x, hi, lo, y);   // x**y, same bits as r8_pow, logarithm of x supplied
}
__host__ __device__ __forceinline__ double r8_log(double x) { return helpmath::det_log(x); }
__host__ __device__ __forceinline__ double r8_atan(double x) { return helpmath::g_atan(x); }
__host__ __device__ __forceinline__ double r8_sin(double x) { return helpmath::g_sin(x); }
__host__ __device__ __forceinline__ double r8_cos(double x) { return helpmath::g_cos(x); }
__host__ __device__ __forceinline__ double r8_sqrt(double x) { return helpmath::hm_sqrt(x); }
#endif


// a / b through the correctly rounded reciprocal of b (help_math.h div_by_const): the bits of
// the IEEE division, 5 FP64 operations.  The glibc-libm emulator build divides directly.
__host__ __device__ __forceinline__ double fdiv_by(double a, double b, double rb) {
#ifdef HELP_R4_OVERRIDE
  (void)rb; return a / b;
#else
  return helpmath::div_by_const(a, b, rb);
#endif
}

}  // namespace helpb200
