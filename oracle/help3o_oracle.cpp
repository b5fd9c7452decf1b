// =============================================================================
// help3o_oracle.cpp  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement (C++17, scalar, one cell per call) of the reference's HELP
// water-balance engine, /root/reference/pyhelp/HELP3O.FOR lines 17-5440 (driver,
// set-up and science routines) plus the value semantics of the monthly report
// writer (OUTMO2, F:6415-6698) and of the Python parser that turns that report
// into the result dict (pyhelp/processing.py:64-147).  "F:" below means
// pyhelp/HELP3O.FOR.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load this library.  The product (pyhelp_b200/)
// never links, imports or calls it.
//
// Parity pin: the reference Fortran cannot be compiled in the build container
// (no gfortran/flang/f2c), so this restatement is pinned against the
// reference's own golden test case pyhelp/tests/rca_original_testcase_1997
// (RCRA.D4/D7/D13/D10/D11 -> RCRA.OUT: daily table, monthly and annual totals,
// post-spin-up moistures, final storages) and the known-answer numbers in
// pyhelp/tests/test_help3o.py:87-118.  See tests/test_oracle_rcra.py.
//
// Typing rules kept from the Fortran: implicit REAL*4 (r4) unless the routine
// declares DOUBLE PRECISION (r8); constants without a D exponent are r4 and are
// promoted per binary operation exactly as Fortran (and C++) promote; integer
// conversions truncate.  Arrays are 1-based (index 0 unused).  Compile with
// -O2 -ffp-contract=off and without -ffast-math so that every operation rounds
// once in its own precision, as gfortran's x86-64 SSE code does.
//
// Documented deviations (SURVEY.md "Hazards"):
//  * KCNT (frozen-soil thaw counter) is never initialised by the reference
//    (F:1344-1345 vs F:3612-3624); a fresh process has 0, which is used here.
//  * Out-of-bounds reads LAYER(0) (F:1559, F:2343) are defined as 0.
//  * Locals the Fortran leaves undefined on paths that valid designs do not
//    take (RTOP2, LINER, H) start at 0.
// =============================================================================
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

// -----------------------------------------------------------------------------
// Transcendental functions.  Two builds of this file (oracle/Makefile):
//   libhelp3o_oracle.so        (default) the deterministic functions of
//       pyhelp_b200/csrc/help_math.h -- REAL*4 intrinsics correctly rounded,
//       REAL*8 intrinsics < 1 ulp, identical bits on CPU and GPU.  This is the
//       build the GPU engine is compared with bit for bit.
//   libhelp3o_oracle_glibc.so  (-DHELP_ORACLE_GLIBC) the host's libm, i.e. what a
//       gfortran build of the reference links (powf/expf/pow ...).  Used to
//       measure how far HELP's results move when only the last bit of libm
//       moves (tests/test_libm_sensitivity.py, DESIGN.md "Numerics").
// Both builds are pinned against the RCRA golden case.
// -----------------------------------------------------------------------------
#if defined(HELP_ORACLE_GLIBC) && defined(HELP_ORACLE_ULP_NOISE)
// Third build (libhelp3o_oracle_glibc_noise.so): the glibc build with ONE unit in the last place
// added to or taken from the result of 1 in HELP_ORACLE_ULP_NOISE calls of pow, log, powf, expf and
// logf (which calls: a hash of the argument bits, so a run is reproducible).  That is the
// difference between two correct libm versions -- each is within 1 ulp of the exact value --
// and it gives the envelope inside which "the reference's results" are defined at all
// (tests/test_parity_envelope.py).
#include <cmath>
#include <cstdint>
#include <cstring>
static inline uint64_t noise_hash(uint64_t b) {
  b ^= b >> 30; b *= 0xBF58476D1CE4E5B9ull; b ^= b >> 27; b *= 0x94D049BB133111EBull; b ^= b >> 31;
  return b;
}
static inline double noisy8(double r, double x, double y) {
  uint64_t bx, by; std::memcpy(&bx, &x, 8); std::memcpy(&by, &y, 8);
  const uint64_t h = noise_hash(bx * 0x9E3779B97F4A7C15ull + by);
  if (h % HELP_ORACLE_ULP_NOISE != 0 || !(r == r) || std::isinf(r) || r == 0.0) return r;
  return std::nextafter(r, ((h >> 40) & 1) ? INFINITY : -INFINITY);
}
static inline float noisy4(float r, float x, float y) {
  uint32_t bx, by; std::memcpy(&bx, &x, 4); std::memcpy(&by, &y, 4);
  const uint64_t h = noise_hash(((uint64_t)bx << 32 | by) * 0x9E3779B97F4A7C15ull + 1);
  if (h % HELP_ORACLE_ULP_NOISE != 0 || !(r == r) || std::isinf(r) || r == 0.0f) return r;
  return std::nextafterf(r, ((h >> 40) & 1) ? INFINITY : -INFINITY);
}
#define R4_EXP(x) noisy4(expf(x), (x), 0.f)
#define R4_LOG(x) noisy4(logf(x), (x), 1.f)
#define R4_LOG10(x) log10f(x)
#define R4_SIN(x) sinf(x)
#define R4_COS(x) cosf(x)
#define R4_TAN(x) tanf(x)
#define R4_ACOS(x) acosf(x)
#define R4_POW(x, y) noisy4(powf(x, y), (x), (y))
#define R4_SQRT(x) powf(x, 0.5f)
#define R4_SQRTF(x) sqrtf(x)
#define R4_POW8(x) powf(x, 8.f)
#define R4_POW7(x) powf(x, 7.f)
#define R4_POW4(x) powf(x, 4.f)
#define R8_POW(x, y) noisy8(pow(x, y), (x), (y))
#define R8_LOG(x) noisy8(log(x), (x), 1.0)
#define R8_ATAN(x) atan(x)
#define R8_SIN(x) sin(x)
#define R8_COS(x) cos(x)
#define R8_SQRT(x) sqrt(x)
#elif defined(HELP_ORACLE_GLIBC)
#define R4_EXP(x) expf(x)
#define R4_LOG(x) logf(x)
#define R4_LOG10(x) log10f(x)
#define R4_SIN(x) sinf(x)
#define R4_COS(x) cosf(x)
#define R4_TAN(x) tanf(x)
#define R4_ACOS(x) acosf(x)
#define R4_POW(x, y) powf(x, y)
#define R4_SQRT(x) powf(x, 0.5f)
#define R4_SQRTF(x) sqrtf(x)
#define R4_POW8(x) powf(x, 8.f)
#define R4_POW7(x) powf(x, 7.f)
#define R4_POW4(x) powf(x, 4.f)
#define R8_POW(x, y) pow(x, y)
#define R8_LOG(x) log(x)
#define R8_ATAN(x) atan(x)
#define R8_SIN(x) sin(x)
#define R8_COS(x) cos(x)
#define R8_SQRT(x) sqrt(x)
#else
#include "../pyhelp_b200/csrc/help_math.h"
#define R4_EXP(x) helpmath::r4_exp(x)
#define R4_LOG(x) helpmath::r4_log(x)
#define R4_LOG10(x) helpmath::r4_log10(x)
#define R4_SIN(x) helpmath::r4_sin(x)
#define R4_COS(x) helpmath::r4_cos(x)
#define R4_TAN(x) helpmath::r4_tan(x)
#define R4_ACOS(x) helpmath::r4_acos(x)
#define R4_POW(x, y) helpmath::r4_pow(x, y)
#define R4_SQRT(x) helpmath::r4_sqrt(x)
#define R4_SQRTF(x) helpmath::r4_sqrt(x)
#define R4_POW8(x) helpmath::r4_pow8(x)
#define R4_POW7(x) helpmath::r4_pow7(x)
#define R4_POW4(x) helpmath::r4_pow4(x)
#define R8_POW(x, y) helpmath::det_pow(x, y)
#define R8_LOG(x) helpmath::det_log(x)
#define R8_ATAN(x) helpmath::g_atan(x)
#define R8_SIN(x) helpmath::g_sin(x)
#define R8_COS(x) helpmath::g_cos(x)
#define R8_SQRT(x) helpmath::hm_sqrt(x)
#endif

namespace {

typedef float r4;
typedef double r8;

// -----------------------------------------------------------------------------
// Fortran formatted-record helpers (blank = null, explicit '.' overrides Fw.d).
// -----------------------------------------------------------------------------
struct Unit {                       // a sequential formatted file
  const std::vector<std::string>* lines = nullptr;
  size_t pos = 0;
  bool eof() const { return !lines || pos >= lines->size(); }
  const std::string* next() {       // one READ record; nullptr at end of file
    if (eof()) return nullptr;
    return &(*lines)[pos++];
  }
  void rewind() { pos = 0; }
};

static std::string field(const std::string& rec, int col0, int w) {
  std::string out;
  for (int i = col0; i < col0 + w; ++i) {
    char c = (i < (int)rec.size()) ? rec[i] : ' ';
    if (c != ' ' && c != '\t' && c != '\r' && c != '\n') out.push_back(c);
  }
  return out;
}
static bool has_point_or_exp(const std::string& s) {
  for (size_t i = 0; i < s.size(); ++i) {
    char c = s[i];
    if (c == '.' || c == 'e' || c == 'E' || c == 'd' || c == 'D') return true;
  }
  return false;
}
static std::string dexp_to_e(std::string s) {
  for (auto& c : s) if (c == 'd' || c == 'D') c = 'E';
  return s;
}
static r4 rd_r4(const std::string& rec, int col0, int w, int d) {
  std::string s = field(rec, col0, w);
  if (s.empty()) return 0.0f;
  s = dexp_to_e(s);
  if (has_point_or_exp(s)) return strtof(s.c_str(), nullptr);
  // no decimal point: the rightmost d digits are the fraction
  r8 v = strtod(s.c_str(), nullptr);
  for (int i = 0; i < d; ++i) v /= 10.0;
  return (r4)v;
}
static r8 rd_r8(const std::string& rec, int col0, int w, int d) {
  std::string s = field(rec, col0, w);
  if (s.empty()) return 0.0;
  s = dexp_to_e(s);
  r8 v = strtod(s.c_str(), nullptr);
  if (!has_point_or_exp(s)) for (int i = 0; i < d; ++i) v /= 10.0;
  return v;
}
static int rd_i(const std::string& rec, int col0, int w) {
  std::string s = field(rec, col0, w);
  if (s.empty()) return 0;
  return atoi(s.c_str());
}

// integer powers as gfortran expands x**n (binary exponentiation)
static inline r4 pw2(r4 x) { return x * x; }
static inline r4 pw3(r4 x) { return x * (x * x); }
static inline r4 pw4(r4 x) { r4 t = x * x; return t * t; }
static inline r4 pw5(r4 x) { r4 t = x * x; return x * (t * t); }
static inline r8 dpw2(r8 x) { return x * x; }
// INT() with the (undefined in Fortran) overflow case pinned to a large value
static inline int trunc_i(r4 x) { if (!(x < 1.0e8f)) return 100000000; if (x < -1.0e8f) return -100000000; return (int)x; }

static std::map<std::string, std::vector<std::string>> g_files;
static const std::vector<std::string>* load_file(const char* path) {
  auto it = g_files.find(path);
  if (it != g_files.end()) return &it->second;
  FILE* f = fopen(path, "rb");
  if (!f) return nullptr;
  std::vector<std::string> lines;
  std::string cur;
  int c;
  while ((c = fgetc(f)) != EOF) {
    if (c == '\n') { lines.push_back(cur); cur.clear(); }
    else if (c != '\r') cur.push_back((char)c);
  }
  if (!cur.empty()) lines.push_back(cur);
  fclose(f);
  auto res = g_files.emplace(path, std::move(lines));
  return &res.first->second;
}

// -----------------------------------------------------------------------------
// Result layout shared with oracle/oracle.py (ctypes).
// -----------------------------------------------------------------------------
enum { MXYR = 100, NROW = 14, NDAILY = 16 };
// monthly rows (inches, REAL*4 accumulators exactly as the Fortran holds them):
//   0 PREM 1 RUNM 2 ETM 3..7 DRN1M..DRN5M 8..13 PRC1M..PRC6M
// parsed rows (mm, 3 decimals, as processing.py:139-146 returns):
//   0 precip 1 runoff 2 evapo 3 subrun1 4 subrun2 5 perco 6 rechg

struct Help {
  // ---- COMMON blocks (F:59-131) ------------------------------------------
  int IT4 = 0, IT7 = 0, IT13 = 0, IU4 = 0, IU7 = 0, IU13 = 0, IU11 = 0, IU10 = 0;
  r4 PORO[22] = {0}, FC[22] = {0}, WP[22] = {0}, SW[22] = {0}, RS[22] = {0},
     XLAMBD[22] = {0}, BUB[22] = {0}, THICK[22] = {0}, SLOPE[22] = {0},
     XLENG[22] = {0}, CON[22] = {0}, SUBIN[22] = {0}, RECIR[22] = {0},
     PHOLE[22] = {0}, DEFEC[22] = {0}, TRANS[22] = {0};
  r8 RC[22] = {0};
  r4 EDEPTH = 0, SUBINF = 0, SEDEP = 0, CORECT = 0;
  int LAYER[23] = {0}, LAYR[22] = {0}, IPQ[22] = {0}, ISOIL[22] = {0}, LAY = 0;
  int LAYSEG[72][3] = {{0}}, LSEG[72] = {0}, LSEGS[23][3] = {{0}}, LRIN[22] = {0},
      LSIN[22] = {0};
  r4 AREA = 0, FRUNOF = 0, CN2 = 0, OCN2 = 0, SSLOPE = 0, SLENG = 0, SMX = 0;
  r4 ULAI = 0, WIND = 0, RH[368] = {0}, OSNO = 0, ULAT = 0;
  r4 STHICK[72] = {0}, UL[72] = {0}, FCUL[72] = {0}, WPUL[72] = {0}, SWUL[72] = {0},
     RCUL[72] = {0}, RSUL[72] = {0}, XLAMBU[72] = {0}, CONUL[72] = {0},
     BUBUL[72] = {0}, SUBINS[72] = {0}, SWULI[72] = {0};
  r4 PRE[372] = {0}, TMPF[368] = {0}, RAD[368] = {0};
  r4 PINF = 0, ES1T = 0, SMELT = 0, GMRO = 0, ADDRUN = 0, OLDSNO = 0, SNO = 0,
     WF[8] = {0}, SSNO = 0;
  r4 ETT = 0, ESS = 0, EP = 0, ES = 0, ET[72] = {0};
  r4 ETO = 0, XLAI = 0, STAGE1 = 0, CONA = 0, RAIN = 0, RUN = 0, ABST = 0, EAJ = 0,
     TS2 = 0;
  // monthly (index [n][month]) and annual accumulators, REAL*4 (F:78-88)
  r4 PRCM[7][13] = {{0}}, DRNM[6][13] = {{0}}, RCRM[6][13] = {{0}}, HEDM[6][13] = {{0}},
     PREM[13] = {0}, RUNM[13] = {0}, ETM[13] = {0}, RCRIM[22][13] = {{0}};
  r4 PRCA[7] = {0}, DRNA[6] = {0}, RCRA_[6] = {0}, BAL = 0, PREA = 0, RUNA = 0, ETA = 0,
     STOR = 0, RCRIA[22] = {0}, OSWULE = 0, PSWULE = 0;
  r4 PRCRI[22] = {0};
  // daily fluxes, DOUBLE PRECISION (F:93-95)
  r8 PRC[7] = {0}, DRN[7] = {0}, HED[7] = {0}, RCR[6] = {0}, RCRI[22] = {0};
  int LD[8] = {0}, LP[8] = {0}, LPQ[8] = {0}, LCASE[8] = {0}, IYR = 0, IDA = 0, MO = 0,
      NSTEP[7] = {0}, ND = 0;
  int NSEGn[7] = {0}, NSEG = 0;          // NSEG1..NSEG6, NSEG
  r4 RM[13] = {0}, TM[13] = {0}, RLAT = 0;
  int IPL = 0, IHV = 0, IPRE = 0, IRUN = 0, ITSOIL = 0, IVEG = 0;
  r4 TSEG[8] = {0}, Z[8] = {0}, FCS[8] = {0}, ST[8] = {0};
  r4 TB = 0, TO = 0, TS = 0, AVT = 0, AMP = 0, ABD = 0, PHU = 0, GS = 0, CV = 0, RSD = 0,
     DM = 0, WFT[13] = {0}, DD = 0, DLAI = 0;
  r4 TMPSUM = 0, TMPAVG = 0, TAVG[31] = {0};
  int IFREZ = 0, IDFS = 0, IFCNT = 0, KCNT = 0, MXKCNT = 0;
  r4 WE = 0, XNEGHS = 0, XLIQW = 0, TINDEX = 0, STORGE = 0, EXLAG[5] = {0}, TWE = 0;
  r4 XMFMAX = 0, XMFMIN = 0, XNMF = 0, TIPM = 0, UADJ = 0, TBASE = 0, PXTEMP = 0,
     PLWHC = 0, GM = 0, PLQW = 0, SFNEW = 0, RFMIN = 0;
  r8 DRCUL[72] = {0}, DBUBUL[72] = {0}, DRSUL[72] = {0}, DLAMUL[72] = {0}, DUL[72] = {0},
     DFCUL[72] = {0}, DWPUL[72] = {0}, DSWUL[72] = {0}, DTHICK[72] = {0},
     DSUBIN[72] = {0}, DCHG[72] = {0}, DRCRS[72] = {0};
  int JYEAR[MXYR + 2] = {0};
  int NSEGB[7] = {0}, NSEGE_[7] = {0}, NSEGD[7] = {0}, LAYB[7] = {0}, LAYE[7] = {0};
  // run_simulation locals that are passed around by reference (F:133-134)
  r8 BALY[72] = {0}, BALT[72] = {0}, DRIN[73] = {0}, SWULY[72] = {0};
  r4 SMXN = 0, SMXF = 0;
  r4 EXWAT[7] = {0};
  int nonconv = 0;                       // count of F:4481 warnings
  // ---- I/O units -----------------------------------------------------------
  Unit u4, u7, u13;
  std::vector<std::string> d10, d11;
  int err = 0;

  // ---- routines --------------------------------------------------------------
  static int LEAP(int NYEAR, int& NT) {                       // F:1043-1056
    if (NYEAR % 4 != 0) { NT = 0; return 365; }
    NT = 1; return 366;
  }
  static int MONTH(int JDAY, int NT) {                        // F:1067-1093
    static const int ICAL[13] = {0, 31, 59, 90, 120, 151, 181, 212, 243, 273, 304, 334, 365};
    int JTMP = JDAY;
    if (JDAY > ICAL[2]) JTMP = JDAY - NT;
    int I;
    for (I = 1; I <= 12; ++I) if (JTMP <= ICAL[I]) return I;
    return 1;
  }

  bool skip_header(Unit& u) {                                 // FORMAT(I2/I2/A40/12F6.2): 4 records
    for (int i = 0; i < 4; ++i) if (!u.next()) return false;
    return true;
  }

  void READCD(int& NYEAR, int& NDv, int& NT) {                // F:1100-1154
    int MYEAR = 0, LYEAR = 0;
    for (int K = 1; K <= 37; ++K) {
      int J = 10 * K, I = J - 9;
      const std::string* r = u4.next();
      if (!r) { err = 3; return; }
      NYEAR = rd_i(*r, 0, 10);
      for (int N = I; N <= J; ++N) PRE[N] = rd_r4(*r, 10 + 5 * (N - I), 5, 2);
    }
    NDv = LEAP(NYEAR, NT);
    for (int K = 1; K <= 37; ++K) {
      int J = 10 * K, I = J - 9;
      if (J > 366) J = 366;
      const std::string* r7 = u7.next();
      const std::string* r13 = u13.next();
      if (!r7 || !r13) { err = 3; return; }
      MYEAR = rd_i(*r7, 0, 5);
      for (int N = I; N <= J; ++N) TMPF[N] = rd_r4(*r7, 5 + 6 * (N - I), 6, 1);
      LYEAR = rd_i(*r13, 0, 5);
      for (int N = I; N <= J; ++N) RAD[N] = rd_r4(*r13, 5 + 6 * (N - I), 6, IU13 == 1 ? 1 : 0);
    }
    (void)MYEAR; (void)LYEAR;   // year-mismatch warnings (F:1130-1137) are text only
    if (IU4 != 1) for (int N = 1; N <= 366; ++N) PRE[N] = PRE[N] / 25.4f;
    if (IU7 != 1) for (int N = 1; N <= 366; ++N) TMPF[N] = (TMPF[N] * 1.8f) + 32.0f;
    if (IU13 != 1) for (int N = 1; N <= 366; ++N) RAD[N] = RAD[N] * 23.89f;
  }

  void READIN() {                                             // F:1161-1397
    u4.rewind(); u7.rewind(); u13.rewind();
    // READ (4,5040) IT4, IU4, CITY4, (RM(I),I=1,12)   FORMAT (I2/I2/A40/12F6.2)
    {
      const std::string *a = u4.next(), *b = u4.next(), *c = u4.next(), *d = u4.next();
      if (!a || !b || !c || !d) { err = 2; return; }
      IT4 = rd_i(*a, 0, 2); IU4 = rd_i(*b, 0, 2);
      for (int I = 1; I <= 12; ++I) RM[I] = rd_r4(*d, 6 * (I - 1), 6, 2);
    }
    {
      const std::string *a = u7.next(), *b = u7.next(), *c = u7.next(), *d = u7.next();
      if (!a || !b || !c || !d) { err = 2; return; }
      IT7 = rd_i(*a, 0, 2); IU7 = rd_i(*b, 0, 2);
      for (int I = 1; I <= 12; ++I) TM[I] = rd_r4(*d, 6 * (I - 1), 6, 2);
    }
    {
      const std::string *a = u13.next(), *b = u13.next(), *c = u13.next(), *d = u13.next();
      if (!a || !b || !c || !d) { err = 2; return; }
      IT13 = rd_i(*a, 0, 2); IU13 = rd_i(*b, 0, 2);
      RLAT = rd_r4(*d, 0, 6, 2);
    }
    // D11 (F:1221-1225): FORMAT (I2) ; (F10.2,I4,I4,F7.0,F8.0,5F5.0)
    IU11 = rd_i(d11[0], 0, 2);
    r4 HUM1, HUM2, HUM3, HUM4;
    {
      const std::string& r = d11[2];
      ULAT = rd_r4(r, 0, 10, 2);
      IPL = rd_i(r, 10, 4); IHV = rd_i(r, 14, 4);
      ULAI = rd_r4(r, 18, 7, 0); EDEPTH = rd_r4(r, 25, 8, 0);
      WIND = rd_r4(r, 33, 5, 0);
      HUM1 = rd_r4(r, 38, 5, 0); HUM2 = rd_r4(r, 43, 5, 0);
      HUM3 = rd_r4(r, 48, 5, 0); HUM4 = rd_r4(r, 53, 5, 0);
    }
    if (IU7 == 1) for (int I = 1; I <= 12; ++I) TM[I] = (TM[I] - 32.0f) / 1.8f;
    if (IU11 != 1) EDEPTH = EDEPTH / 2.54f;
    if (IU11 != 1) WIND = WIND / 1.609f;
    for (int I = 1; I <= 91; ++I) RH[I] = HUM1;
    for (int I = 92; I <= 182; ++I) RH[I] = HUM2;
    for (int I = 183; I <= 274; ++I) RH[I] = HUM3;
    for (int I = 275; I <= 366; ++I) RH[I] = HUM4;
    // D10 (F:1248-1300)
    OCN2 = 0.f; SSLOPE = 0.f; SLENG = 0.f; IVEG = 0; ITSOIL = 0; LAY = 0;
    for (int I = 1; I <= 20; ++I) LAYR[I] = 0;
    int nl = (int)d10.size();
    int IL = 0;                      // 0-based record cursor (ILINE-1)
    if (nl < 3) { err = 4; return; }
    IL = 1;                          // TITLE consumed
    {
      const std::string& r = d10[IL++];         // FORMAT(I2,I2,2F10.0,F6.0,I2)
      IU10 = rd_i(r, 0, 2); IPRE = rd_i(r, 2, 2);
      OSNO = rd_r4(r, 4, 10, 0); AREA = rd_r4(r, 14, 10, 0);
      FRUNOF = rd_r4(r, 24, 6, 0); IRUN = rd_i(r, 30, 2);
    }
    if (IRUN == 1) { const std::string& r = d10[IL++]; CN2 = rd_r4(r, 0, 7, 0); }
    else if (IRUN == 2) {
      const std::string& r = d10[IL++];
      OCN2 = rd_r4(r, 0, 7, 0); SSLOPE = rd_r4(r, 7, 7, 0);
      SLENG = rd_r4(r, 14, 7, 0); CN2 = rd_r4(r, 21, 7, 0);
    } else if (IRUN == 3) {
      const std::string& r = d10[IL++];
      SSLOPE = rd_r4(r, 0, 7, 0); SLENG = rd_r4(r, 7, 7, 0);
      ITSOIL = rd_i(r, 14, 3); IVEG = rd_i(r, 17, 3); CN2 = rd_r4(r, 20, 7, 0);
    }
    for (int J = 1; J <= 20; ++J) {
      if (IL >= nl) break;
      {
        const std::string& r = d10[IL++];       // FORMAT(I2,F7.0,I4,4F6.0,F16.0)
        LAYER[J] = rd_i(r, 0, 2); THICK[J] = rd_r4(r, 2, 7, 0); ISOIL[J] = rd_i(r, 9, 4);
        PORO[J] = rd_r4(r, 13, 6, 0); FC[J] = rd_r4(r, 19, 6, 0); WP[J] = rd_r4(r, 25, 6, 0);
        SW[J] = rd_r4(r, 31, 6, 0); RC[J] = rd_r8(r, 37, 16, 0);
      }
      LAY = LAY + 1;
      if (IL >= nl) break;
      {
        const std::string& r = d10[IL++];       // FORMAT(F7.0,2F6.0,I3,F13.0,2F7.0,I2,G14.6)
        XLENG[J] = rd_r4(r, 0, 7, 0); SLOPE[J] = rd_r4(r, 7, 6, 0); RECIR[J] = rd_r4(r, 13, 6, 0);
        LAYR[J] = rd_i(r, 19, 3); SUBIN[J] = rd_r4(r, 22, 13, 0); PHOLE[J] = rd_r4(r, 35, 7, 0);
        DEFEC[J] = rd_r4(r, 42, 7, 0); IPQ[J] = rd_i(r, 49, 2); TRANS[J] = rd_r4(r, 51, 14, 6);
      }
    }
    // Brooks-Corey parameters and evaporation coefficient (F:1307-1340)
    for (int J = 1; J <= LAY; ++J) {
      if (LAYER[J] == 4) { PORO[J] = .04f; FC[J] = .03f; WP[J] = .02f; SW[J] = .04f; }
      if (WP[J] < 0.04f) RS[J] = 0.6f * WP[J];
      else RS[J] = 0.014f + 0.253f * WP[J];
      if (WP[J] < 0.02f) RS[J] = 0.75f * WP[J];
      if (WP[J] <= 0.0f) WP[J] = 0.001f;
      if (FC[J] <= WP[J]) FC[J] = WP[J] + 0.001f;
      if (PORO[J] <= FC[J]) PORO[J] = FC[J] + 0.001f;
      XLAMBD[J] = R4_LOG((FC[J] - RS[J]) / (WP[J] - RS[J])) / R4_LOG(45.f);
      r4 RELFC = (FC[J] - RS[J]) / (PORO[J] - RS[J]);
      BUB[J] = (R4_POW(RELFC, 1.f / XLAMBD[J])) * 1033.4f / 3.f;
      if (BUB[J] > 103.34f) BUB[J] = 103.34f;
      CON[J] = (r4)(2.44f + (1485216.f * RC[J] * (R4_POW(BUB[J] / 103.34f, (3.f * XLAMBD[J]) + 2.f))));
      if (RELFC < 0.20f) CON[J] = 3.30f;
      if (RC[J] < 0.000005f) CON[J] = 3.30f;
      if (CON[J] < 3.30f) CON[J] = 3.30f;
      if (CON[J] > 5.50f) CON[J] = 5.50f;
      if (LAYER[J] == 4) { PORO[J] = .0f; FC[J] = .0f; WP[J] = .0f; SW[J] = .0f; }
    }
    // frozen-soil model parameters (F:1344-1375)
    IFREZ = 0; IFCNT = 0; TMPAVG = 33.0f; TMPSUM = 990.0f;
    r4 RADTOT = 0.0f;
    for (int I = 1; I <= 30; ++I) TAVG[I] = 33.0f;
    int NYRS = 0;
    r4 DRAD[32];
    for (int I = 1; I <= 100; ++I) {
      // every READ here consumes the 37 records of one year block
      const std::string* rec[37];
      bool ok = true;
      size_t save = u13.pos;
      for (int k = 0; k < 37; ++k) { rec[k] = u13.next(); if (!rec[k]) { ok = false; break; } }
      if (!ok) {
        // Fortran hits END= part way; a partly read block is discarded.  (For the
        // southern branch the first READ needs only 19 records; honour that.)
        if (ULAT < 0.0f) {
          u13.pos = save;
          bool ok19 = true;
          for (int k = 0; k < 19; ++k) { rec[k] = u13.next(); if (!rec[k]) { ok19 = false; break; } }
          (void)ok19;  // the second READ (5420) then fails -> END=1410 before RADTOT is touched
        }
        break;
      }
      if (ULAT >= 0.0f) {          // FORMAT(33(/),29X,6F6.1/2(5X,10F6.1/),5X,5F6.1)
        int n = 1;
        for (int k = 0; k < 6; ++k) DRAD[n++] = rd_r4(*rec[33], 29 + 6 * k, 6, 1);
        for (int k = 0; k < 10; ++k) DRAD[n++] = rd_r4(*rec[34], 5 + 6 * k, 6, 1);
        for (int k = 0; k < 10; ++k) DRAD[n++] = rd_r4(*rec[35], 5 + 6 * k, 6, 1);
        for (int k = 0; k < 5; ++k) DRAD[n++] = rd_r4(*rec[36], 5 + 6 * k, 6, 1);
      } else {                      // FORMAT(15(/),11X,9F6.1/2(5X,10F6.1/),5X,2F6.1) + 17(/)
        int n = 1;
        for (int k = 0; k < 9; ++k) DRAD[n++] = rd_r4(*rec[15], 11 + 6 * k, 6, 1);
        for (int k = 0; k < 10; ++k) DRAD[n++] = rd_r4(*rec[16], 5 + 6 * k, 6, 1);
        for (int k = 0; k < 10; ++k) DRAD[n++] = rd_r4(*rec[17], 5 + 6 * k, 6, 1);
        for (int k = 0; k < 2; ++k) DRAD[n++] = rd_r4(*rec[18], 5 + 6 * k, 6, 1);
      }
      for (int N = 1; N <= 31; ++N) RADTOT = RADTOT + DRAD[N];
      NYRS = I;
    }
    r4 RADAVG;
    if (NYRS > 1) RADAVG = RADTOT / NYRS / 31;
    else RADAVG = RADTOT / 31;
    if (IU13 != 1) RADAVG = RADAVG * 23.89f;
    IDFS = (int)(35.4f - (0.154f * RADAVG));
    if (IDFS < 1) IDFS = 1;
    MXKCNT = (IDFS + 2) / 3;
    u13.rewind();
    skip_header(u13);
  }

  void CNTRLD() {                                             // F:2316-2387
    int NL = 0;
    for (int I = 1; I <= 5; ++I) { LP[I] = 0; LPQ[I] = 0; LD[I] = 0; }
    LP[6] = 0;
    for (int I = 1; I <= LAY; ++I) {
      LRIN[I] = 0; LSIN[I] = 0;
      if (LAYER[I] == 3 || LAYER[I] == 4) {
        if (LAYER[I - 1] != 3 && LAYER[I - 1] != 4) NL = NL + 1;
        if (NL >= 1) {
          LP[NL] = I;
          if (LAYER[I] == 4) {
            if (IPQ[I] < 1 || IPQ[I] > 6) IPQ[I] = 4;
            LPQ[NL] = IPQ[I];
          }
        }
      }
    }
    for (int I = 1; I <= 5; ++I) {
      if (LP[I] >= 2) {
        if (LAYER[LP[I] - 1] == 2) LD[I] = LP[I] - 1;
        else if (LP[I] > 2) { if (LAYER[LP[I] - 2] == 2) LD[I] = LP[I] - 2; }
      }
    }
    if (LAYER[LAY] != 3 && LAYER[LAY] != 4) {
      if (NL == 0) LP[1] = LAY;
      if (NL == 1 && LAY > LP[1]) LP[2] = LAY;
      if (NL == 2 && LAY > LP[2]) LP[3] = LAY;
      if (NL == 3 && LAY > LP[3]) LP[4] = LAY;
      if (NL == 4 && LAY > LP[4]) LP[5] = LAY;
      if (NL == 5 && LAY > LP[5]) LP[6] = LAY;
    }
    for (int N = 1; N <= 5; ++N) {
      if (LD[N] != 0 && LAYR[LD[N]] != 0) {
        if (LAYER[LAYR[LD[N]]] > 2) { LRIN[LAYR[LD[N]]] = 0; RECIR[LAYR[LD[N]]] = 0.0f; }
        else if (RECIR[LD[N]] > 0.0f) LRIN[LAYR[LD[N]]] = 1;
      }
    }
    for (int N = 1; N <= LAY; ++N) if (SUBIN[N] > 0.0f) LSIN[N] = 1;
  }

  // liner / FML segment rules shared by the five upper subprofile loops
  // (F:1554-1585, 1624-1655, ...).  Returns true when layer J was consumed.
  bool seg_liner(int J, int n, int& NSEGS, r4 thick_liner) {
    if (LAYER[J] == 3) {
      STHICK[NSEGS + 1] = thick_liner;
      LAYSEG[NSEGS + 1][1] = J;
      LSEG[NSEGS + 1] = LAYER[J];
      if (J != LP[n]) LSEG[NSEGS + 1] = 5;
      if (LAYER[J - 1] == 4) LSEG[NSEGS + 1] = 4;
      NSEGS = NSEGS + 1;
      return true;
    }
    if (LAYER[J] == 4) {
      if (J == LP[n]) {
        if (LAYER[J - 1] != 3) {
          if ((J + 1) <= LAY) {
            if (RC[J + 1] < RC[J - 1]) {
              if (THICK[J + 1] > 18.f) STHICK[NSEGS + 1] = 12.f;
              else STHICK[NSEGS + 1] = THICK[J + 1];
              LSEG[NSEGS + 1] = 6;
              LAYSEG[NSEGS + 1][1] = J + 1;
            } else {
              STHICK[NSEGS + 1] = STHICK[NSEGS];
              LSEG[NSEGS + 1] = 7;
              LAYSEG[NSEGS + 1][1] = J - 1;
            }
          } else {
            STHICK[NSEGS + 1] = STHICK[NSEGS];
            LSEG[NSEGS + 1] = 7;
            LAYSEG[NSEGS + 1][1] = J - 1;
          }
          NSEGS = NSEGS + 1;
        }
      }
      return true;
    }
    return false;
  }
  // 1 / 2 / 3 segments for a soil layer of thickness T (F:1586-1616 ...)
  void seg_soil(int J, int& NSEGS, r4 T) {
    if (T > 36.f) {
      STHICK[NSEGS + 1] = 12.f; STHICK[NSEGS + 2] = T - 12.f - 12.f; STHICK[NSEGS + 3] = 12.f;
      for (int k = 1; k <= 3; ++k) { LAYSEG[NSEGS + k][1] = J; LAYSEG[NSEGS + k][2] = J; LSEG[NSEGS + k] = LAYER[J]; }
      NSEGS = NSEGS + 3;
    } else if (T > 18.f) {
      STHICK[NSEGS + 1] = 12.f; STHICK[NSEGS + 2] = T - 12.f;
      for (int k = 1; k <= 2; ++k) { LAYSEG[NSEGS + k][1] = J; LAYSEG[NSEGS + k][2] = J; LSEG[NSEGS + k] = LAYER[J]; }
      NSEGS = NSEGS + 2;
    } else {
      STHICK[NSEGS + 1] = T;
      LAYSEG[NSEGS + 1][1] = J; LAYSEG[NSEGS + 1][2] = J; LSEG[NSEGS + 1] = LAYER[J];
      NSEGS = NSEGS + 1;
    }
  }

  void SGMNT() {                                              // F:1406-2311
    r4 RCULS[72] = {0};
    for (int n = 1; n <= 6; ++n) NSEGn[n] = 0;
    for (int N = 1; N <= 5; ++N) LCASE[N] = 0;
    r4 THICK1 = 0.0f;
    CORECT = 1.0f;
    if (ISOIL[1] <= 34 && ISOIL[1] >= 1) {
      CORECT = 1.0f + 0.5966f * ULAI + 0.132659f * pw2(ULAI) + 0.1123454f * pw3(ULAI) -
               0.04777627f * pw4(ULAI) + 0.004325035f * pw5(ULAI);
      if (CORECT > 5.0f) CORECT = 5.0f;
    }
    for (int J = 1; J <= 67; ++J) STHICK[J] = 0.0f;
    SUBINF = 0.0f;
    for (int K = 1; K <= LAY; ++K) {
      LSEGS[K][1] = 0; LSEGS[K][2] = 0;
      SUBIN[K] = SUBIN[K] / 365.0f;
      SUBINF = SUBINF + SUBIN[K];
      if (IPRE == 0 || IPRE == 2) SW[K] = FC[K];
      if (LAYER[K] == 4) SW[K] = 0.0f;
      if (LAYER[K] == 3) SW[K] = PORO[K];
    }
    FRUNOF = FRUNOF / 100.0f;
    if (IU10 == 2) OSNO = OSNO / 25.4f;
    OLDSNO = OSNO;
    if (IU10 == 2) {
      SLENG = SLENG * 3.281f;
      SUBINF = SUBINF / 25.4f;
      for (int J = 1; J <= LAY; ++J) {
        DEFEC[J] = DEFEC[J] / 2.471f; PHOLE[J] = PHOLE[J] / 2.471f;
        THICK[J] = THICK[J] / 2.54f; XLENG[J] = XLENG[J] * 3.281f;
        SUBIN[J] = SUBIN[J] / 25.4f;
      }
    }
    // depth to the top liner limits the evaporative depth (F:1493-1498)
    for (int J = 1; J <= LP[1]; ++J) {
      if (LAYER[J] == 3 || LAYER[J] == 4) break;
      THICK1 = THICK1 + THICK[J];
    }
    if (THICK1 < EDEPTH) EDEPTH = THICK1;
    STHICK[1] = EDEPTH / 36.0f;
    STHICK[2] = 5.0f * EDEPTH / 36.0f;
    r4 TTHICK = STHICK[1] + STHICK[2];
    for (int J = 3; J <= 6; ++J) { STHICK[J] = EDEPTH / 6.0f; TTHICK = TTHICK + STHICK[J]; }
    STHICK[7] = EDEPTH - TTHICK;
    {
      int J = 1;
      THICK1 = THICK[J];
      r4 THICK2 = 0.0f;
      for (int K = 1; K <= 7; ++K) {
        THICK2 = THICK2 + STHICK[K];
        LAYSEG[K][1] = J;
        for (;;) {
          if (THICK2 <= THICK1) LAYSEG[K][2] = J;
          if (THICK2 > THICK1) {
            J = J + 1;
            if (J > 20) break;      // guard: the Fortran would run off THICK()
            THICK1 = THICK1 + THICK[J];
            continue;
          }
          break;
        }
      }
    }
    int NSEGS = 7;
    THICK1 = 0.0f;
    r4 THICK2 = EDEPTH + 0.0002f;
    for (int J = 1; J <= LP[1]; ++J) {
      LAYSEG[NSEGS + 1][2] = J;
      THICK1 = THICK1 + THICK[J];
      if (THICK1 > (EDEPTH + 0.0001f)) {
        r4 THICKL = THICK1 - THICK2;
        if (!seg_liner(J, 1, NSEGS, THICKL)) seg_soil(J, NSEGS, THICKL);
        THICK2 = THICK1;
      }
    }
    NSEGn[1] = NSEGS;
    int used = NSEGn[1];
    bool done = (LAY <= LP[1]);
    for (int n = 2; n <= 5 && !done; ++n) {
      for (int J = LP[n - 1] + 1; J <= LP[n]; ++J) {
        LAYSEG[NSEGS + 1][2] = J;
        if (!seg_liner(J, n, NSEGS, THICK[J])) seg_soil(J, NSEGS, THICK[J]);
      }
      NSEGn[n] = NSEGS - used;
      used = NSEGS;
      if (LAY <= LP[n]) done = true;
    }
    if (!done) {
      for (int J = LP[5] + 1; J <= LP[6]; ++J) seg_soil(J, NSEGS, THICK[J]);
    }
    NSEGn[6] = NSEGS - NSEGn[1] - NSEGn[2] - NSEGn[3] - NSEGn[4] - NSEGn[5];
    NSEG = NSEGS;
    // effective segment properties (F:1937-2042)
    r4 DEPSEG = 0.0f, DEPTHS = 0.0f, DEPTHL = THICK[1], DSEG;
    int K = 1;
    LSEGS[K][1] = 1;
    r8 RC1 = RC[1];
    RC[1] = RC[1] * CORECT;
    for (int J = 1; J <= 7; ++J) {
      UL[J] = 0.f; WPUL[J] = 0.f; FCUL[J] = 0.f; CONUL[J] = 0.f; RSUL[J] = 0.f; XLAMBU[J] = 0.f;
      SWUL[J] = 0.f; RCULS[J] = 0.f; BUBUL[J] = 0.f; SUBINS[J] = 0.f;
      if (J == 5) RC[1] = RC1;
      if (DEPSEG >= DEPTHL) { K = K + 1; LSEGS[K][1] = J; DEPTHL = DEPTHL + THICK[K]; }
      DEPTHS = DEPTHS + STHICK[J];
      for (;;) {
        LSEG[J] = LAYER[K];
        if (DEPTHS <= DEPTHL || K >= 20) break;
        DSEG = DEPTHL - DEPSEG;
        if (DSEG > 0.0001f) LSEGS[K][2] = J;
        DEPSEG = DEPTHL;
        UL[J] = PORO[K] * DSEG + UL[J];
        FCUL[J] = FC[K] * DSEG + FCUL[J];
        WPUL[J] = WP[K] * DSEG + WPUL[J];
        CONUL[J] = CON[K] * DSEG + CONUL[J];
        BUBUL[J] = BUB[K] * DSEG + BUBUL[J];
        RCULS[J] = (r4)(RC[K] * DSEG + RCULS[J]);
        RSUL[J] = RS[K] * DSEG + RSUL[J];
        XLAMBU[J] = XLAMBD[K] * DSEG + XLAMBU[J];
        SUBINS[J] = SUBIN[K] * DSEG / THICK[K] + SUBINS[J];
        SWUL[J] = SW[K] * DSEG + SWUL[J];
        K = K + 1;
        LSEGS[K][1] = J;
        DEPTHL = DEPTHL + THICK[K];
      }
      DSEG = DEPTHS - DEPSEG;
      if (DSEG > 0.0001f) LSEGS[K][2] = J;
      UL[J] = PORO[K] * DSEG + UL[J];
      FCUL[J] = FC[K] * DSEG + FCUL[J];
      WPUL[J] = WP[K] * DSEG + WPUL[J];
      CONUL[J] = CON[K] * DSEG + CONUL[J];
      BUBUL[J] = BUB[K] * DSEG + BUBUL[J];
      RCULS[J] = (r4)(RC[K] * DSEG + RCULS[J]);
      RSUL[J] = RS[K] * DSEG + RSUL[J];
      XLAMBU[J] = XLAMBD[K] * DSEG + XLAMBU[J];
      SUBINS[J] = SUBIN[K] * DSEG / THICK[K] + SUBINS[J];
      SWUL[J] = SW[K] * DSEG + SWUL[J];
      DEPSEG = DEPTHS;
    }
    if (NSEG > 7) {
      K = LAYSEG[7][2];
      for (int J = 8; J <= NSEG; ++J) {
        if (LAYSEG[J][2] != K) { K = LAYSEG[J][2]; LSEGS[K][1] = J; }
        LSEGS[K][2] = J;
        int L1 = LAYSEG[J][1], L2 = LAYSEG[J][2];
        UL[J] = PORO[L1] * STHICK[J];
        FCUL[J] = FC[L1] * STHICK[J];
        WPUL[J] = WP[L1] * STHICK[J];
        CONUL[J] = CON[L1] * STHICK[J];
        BUBUL[J] = BUB[L1] * STHICK[J];
        RCULS[J] = (r4)(RC[L1] * STHICK[J]);
        RSUL[J] = RS[L1] * STHICK[J];
        XLAMBU[J] = XLAMBD[L1] * STHICK[J];
        if (LSEG[J] < 3) SUBINS[J] = SUBIN[L2] * STHICK[J] / THICK[L2];
        else SUBINS[J] = SUBIN[L2];
        if (LSEG[J] == 4) SUBINS[J] = SUBINS[J] + SUBIN[L2 - 1];
        SWUL[J] = SW[L1] * STHICK[J];
      }
    }
    for (int J = 1; J <= NSEG; ++J) {                         // cm/s -> in/day (F:2027-2042)
      RCUL[J] = (RCULS[J] * 24.f * 1417.3f) / STHICK[J];
      XLAMBU[J] = XLAMBU[J] / STHICK[J];
      BUBUL[J] = BUBUL[J] / STHICK[J];
      CONUL[J] = CONUL[J] / STHICK[J];
      DRCUL[J] = RCUL[J]; DUL[J] = UL[J]; DRSUL[J] = RSUL[J]; DLAMUL[J] = XLAMBU[J];
      DSUBIN[J] = SUBINS[J]; DBUBUL[J] = BUBUL[J]; DWPUL[J] = WPUL[J]; DTHICK[J] = STHICK[J];
      DSWUL[J] = SWUL[J]; DFCUL[J] = FCUL[J];
    }
    // (F:2043-2056 computes an unused TTHICK for recirculation targets)
    int NSEGE = 0;
    for (int N = 1; N <= 5; ++N) {                            // F:2057-2294
      if (N == 1) NSEGE = NSEGn[1];
      else { if (!(LP[N] > 0)) continue; NSEGE = NSEGE + NSEGn[N]; }
      if (LP[N] == LAY) {
        DSUBIN[NSEGE - 1] = DSUBIN[NSEGE - 1] + SUBIN[LP[N]];
        SUBINS[NSEGE - 1] = SUBINS[NSEGE - 1] + SUBIN[LP[N]];
        DSUBIN[NSEGE] = 0.0; SUBINS[NSEGE] = 0.0f;
        if ((LP[N] - 1) == 3 || (LP[N] - 1) == 4) {           // sic: compares the layer NUMBER
          DSUBIN[NSEGE - 1] = DSUBIN[NSEGE - 1] + SUBIN[LP[N] - 1];
          SUBINS[NSEGE - 1] = SUBINS[NSEGE - 1] + SUBIN[LP[N] - 1];
        }
      } else if (LSEG[NSEGE] == 5) {
        DSUBIN[NSEGE - 1] = DSUBIN[NSEGE - 1] + SUBIN[LP[N] - 1];
        SUBINS[NSEGE - 1] = SUBINS[NSEGE - 1] + SUBIN[LP[N] - 1];
        DSUBIN[NSEGE] = SUBIN[LP[N]]; SUBINS[NSEGE] = SUBIN[LP[N]];
      }
      if (LSEG[NSEGE] > 3) {
        r8 rck = RC[LAYSEG[NSEGE][1]];
        if (LSEG[NSEGE] == 4) LCASE[N] = 3;
        if (LSEG[NSEGE] == 5) LCASE[N] = 4;
        if (LSEG[NSEGE] == 6) {
          if (rck > 0.0999f) LCASE[N] = 1;
          else if (rck <= 0.0000999f) LCASE[N] = 3;
          else LCASE[N] = 2;
        }
        if (LSEG[NSEGE] == 7) {
          // threshold literal is 0.0999 for N=1 and 0.099 for N=2..5 (F:2087 vs F:2134...)
          bool hi = (N == 1) ? (rck > 0.0999f) : (rck > 0.099f);
          if (hi) LCASE[N] = 1;
          else if (rck <= 0.0000999f) LCASE[N] = 4;
          else LCASE[N] = 2;
        }
        if (LPQ[N] == 6) {
          if (LCASE[N] == 2 || LCASE[N] == 3) LCASE[N] = 5;
          else if (LCASE[N] == 4) LCASE[N] = 6;
        }
      }
    }
    r4 RCES = 0.0f;                                           // F:2296-2309
    for (int J = 1; J <= 7; ++J) RCES = RCES + WF[J] * RCUL[J];
    RCES = RCES / 34016.f;
    RCES = -1.0f * R4_LOG10(RCES);
    r4 SEDMX = 4.606745f * (R4_POW(1.595155f, RCES));
    if (SEDMX < 18.f) SEDMX = 18.f;
    if (SEDMX > 48.f) SEDMX = 48.f;
    if (EDEPTH > SEDMX) SEDEP = SEDMX; else SEDEP = EDEPTH;
  }

  static r4 COMPUT(r4 AC, r4 A, r4 B, int I, int N) {         // F:2525-2531
    r4 AI = (r4)I - 0.5f, AN = (r4)N;
    r4 ANG = 6.283185f * AI / AN;
    return AC + A * R4_COS(ANG) + B * R4_SIN(ANG);
  }
  static r4 AHU(int M, int K, r4 A1, r4 A2, r4 A3) {          // F:2484-2495
    r4 a = 0.f;
    for (int L = M; L <= K; ++L) {
      r4 TA = COMPUT(A1, A2, A3, L, 366) - 5.f;
      if (TA > 0.0f) a = a + TA;
    }
    return a;
  }
  static void DATFIT(int M, const r4* D, r4& AC, r4& A, r4& B) { // F:2501-2524
    r4 PI = 3.14159f, SUMD = 0.0f, AM = (r4)M;
    for (int I = 1; I <= M; ++I) SUMD = SUMD + D[I];
    AC = SUMD / AM;
    r4 AN = 1.0f, SUMA = 0.0f, SUMB = 0.0f;
    for (int I = 1; I <= M; ++I) {
      r4 TI = (r4)I - 0.5f;
      r4 TH = 2.0f * PI * AN * TI;
      r4 FCOS = R4_COS(TH / AM), FSIN = R4_SIN(TH / AM);
      SUMA = SUMA + D[I] * FCOS;
      SUMB = SUMB + D[I] * FSIN;
    }
    A = 2.0f / AM * SUMA;
    B = 2.0f / AM * SUMB;
  }

  void INITIA() {                                             // F:2394-2479
    TB = 5.f; TO = 20.f;
    AVT = 0.f;
    r4 TBB = 0.f;
    TS = 200.f;
    for (int I = 1; I <= 12; ++I) {
      if (TM[I] > TBB) TBB = TM[I];
      if (TM[I] < TS) TS = TM[I];
      AVT = AVT + TM[I];
      WFT[I] = 0.2f / (1.f - 0.4f + 0.2f);
    }
    AVT = AVT / 12.f;
    AMP = (TBB - TS) / 2.0f;
    r4 A1, A2, A3;
    DATFIT(12, TM, A1, A2, A3);
    r4 PHUX = 0.f;
    if (IPL >= IHV) {
      PHU = AHU(IPL, 365, A1, A2, A3);
      PHUX = PHU;
      PHU = PHU + AHU(1, IHV, A1, A2, A3);
    } else {
      PHU = AHU(IPL, IHV, A1, A2, A3);
    }
    if (PHU < 30.f) PHU = 30.f;
    XLAI = (PHUX / PHU) * ULAI;
    DLAI = XLAI;
    GS = PHUX / PHU;
    RSD = 8000.f * ULAI / 5.f;
    CV = RSD;
    DM = XLAI * 100.f;
    ABD = 0.f;
    r4 SWS = 0.f;
    for (int I = 1; I <= 7; ++I) Z[I] = 0.0f;
    for (int I = 1; I <= 7; ++I) {
      r4 ZSEG = (I == 1) ? 0.f : Z[I - 1];
      Z[I] = ZSEG + STHICK[I] * 25.4f;
      r4 XZ = UL[I] * 25.4f * 10.f;
      ABD = ABD + XZ;
      FCS[I] = (FCUL[I] - WPUL[I]) * 25.4f;
      ST[I] = (SWUL[I] - WPUL[I]) * 25.4f;
      SWS = SWS + ST[I];
      ST[I] = ST[I] / 25.4f;
    }
    ABD = ABD / (10.f * Z[7]);
    r4 F = ABD / (ABD + 686.f * R4_EXP(-5.63f * ABD));
    r4 DP = 1000.f + 2500.f * F;
    r4 WW = .356f - .144f * ABD;
    r4 B = R4_LOG(500.f / DP);
    r4 WC = SWS / (WW * Z[7]);
    r4 q = (1.f - WC) / (1.f + WC);
    F = R4_EXP(B * pw2(q));
    DD = F * DP;
  }

  // per-layer water (inches) from segment storages: SETUPS F:2694-2735 and
  // OUTSW F:8493-8531 (report only; does not feed the simulation)
  void layer_water(r4* SWULL) {
    for (int K = 1; K <= 20; ++K) SWULL[K] = 0.0f;
    int J = 1;
    r4 OTHCK1 = 0.0f, THICK1 = STHICK[J], THICK2 = 0.0f;
    r4 SWULJ = (r4)(DSWUL[J] + (DCHG[J] / 2.0));
    for (int K = 1; K <= LP[1]; ++K) {
      if (LAYER[K] > 2) goto L1150;
      THICK2 = THICK2 + THICK[K];
    L1130:
      if (LAYSEG[J][2] < K) goto L1140;
      if (THICK1 > THICK2) {
        r4 DSEG = THICK2 - OTHCK1;
        OTHCK1 = THICK2;
        SWULL[K] = SWULL[K] + ((DSEG * SWULJ) / STHICK[J]);
        continue;
      } else {
        SWULL[K] = SWULL[K] + (SWULJ * (THICK1 - OTHCK1) / STHICK[J]);
      }
    L1140:
      J = J + 1;
      if (J > NSEGn[1]) goto L1150;
      if (LSEG[J] > 2) goto L1150;
      SWULJ = (r4)(DSWUL[J] + (DCHG[J] / 2.0));
      OTHCK1 = THICK1;
      THICK1 = (r4)(THICK1 + DTHICK[J]);
      if (LAYSEG[J][1] > K) continue;
      goto L1130;
    }
  L1150:
    if (NSEG > NSEGn[1]) {
      for (int Jx = NSEGn[1] + 1; Jx <= NSEG; ++Jx)
        if (LSEG[Jx] <= 2)
          SWULL[LAYSEG[Jx][2]] = (r4)(SWULL[LAYSEG[Jx][2]] + DSWUL[Jx] + (DCHG[Jx] / 2.0));
    }
    for (int K = 1; K <= LAY; ++K) if (LAYER[K] > 2) SWULL[K] = SW[K] * THICK[K];
  }

  void SETUPS(r4* SWULL) {                                    // F:2540-2745
    if (IPRE <= 1) {
      for (int N = 1; N <= 6; ++N) NSTEP[N] = 4;
      r4 RTOP = (r4)RC[1];
      for (int I = 2; I <= LAY; ++I) {
        if (LAYER[I] > 2) break;
        if (RC[I] < RTOP) RTOP = (r4)RC[I];
      }
      RTOP = 2.0f * RTOP;
      if (RTOP > 0.0005f) RTOP = 0.0005f;
      r4 RDR, DT;
      if (LD[1] > 1) {
        int L = LD[1];
        if (XLENG[L] == 0.0f) XLENG[L] = 1.f;
        if (SLOPE[L] == 0.0f) SLOPE[L] = 0.00001f;
        RDR = (r4)(RC[L] * THICK[L] * SLOPE[L] / (1200.f * XLENG[L]));
        if (RDR > RTOP) DT = (PORO[L] - FC[L]) * THICK[L] * 2.54f / (RTOP * 86400.0f);
        else DT = (PORO[L] - FC[L]) * THICK[L] * 2.54f / (RDR * 86400.0f);
        NSTEP[1] = (trunc_i(1.0f / DT) + 1) * 2;
        if (NSTEP[1] < 4) NSTEP[1] = 4;
      }
      r4 RTOP2 = 0.0f;
      for (int N = 2; N <= 5; ++N) {
        if (LP[N] > 1) {
          int LPP1 = LP[N - 1] + 1, LPM1 = LP[N - 1] - 1;
          int lm1 = (LPM1 >= 0) ? LAYER[LPM1] : 0;
          if (lm1 == 3) RTOP2 = (r4)RC[LPM1];
          if (LAYER[LP[N - 1]] == 3) RTOP2 = (r4)RC[LP[1]];
          if (lm1 == 4) RTOP2 = RTOP2 * 0.001f;
          if (LAYER[LP[N - 1]] == 4) {
            if (lm1 == 3) RTOP2 = RTOP2 * 0.004f;
            else if (RC[LPM1] < RC[LPP1]) {
              RTOP2 = (r4)(0.004f * RC[LPM1]);
              if (RTOP2 > 0.000005f) RTOP2 = 0.000005f;
            } else {
              RTOP2 = (r4)(0.004f * RC[LPP1]);
              if (RTOP2 > 0.000005f) RTOP2 = 0.000005f;
            }
          }
          RTOP2 = RTOP2 * 2.0f;
          if (RTOP2 < RTOP) RTOP = RTOP2;
          if (LD[N] > 1) { for (int I = LP[N - 1] + 1; I <= LD[N]; ++I) if (RC[I] < RTOP) RTOP = (r4)RC[I]; }
          else { for (int I = LP[N - 1] + 1; I <= LP[N]; ++I) if (RC[I] < RTOP) RTOP = (r4)RC[I]; }
          if (LD[N] > 1) {
            int L = LD[N];
            if (XLENG[L] == 0.0f) XLENG[L] = 1.f;
            if (SLOPE[L] == 0.0f) SLOPE[L] = 0.00001f;
            RDR = (r4)(RC[L] * THICK[L] * SLOPE[L] / (1200.f * XLENG[L]));
            if (RDR > RTOP) DT = (PORO[L] - FC[L]) * THICK[L] * 2.54f / (RTOP * 86400.0f);
            else DT = (PORO[L] - FC[L]) * THICK[L] * 2.54f / (RDR * 86400.0f);
            NSTEP[N] = (trunc_i(1.0f / DT) + 1) * 2;
          }
        }
      }
      for (int N = 1; N <= 6; ++N) {
        if (NSTEP[N] < 4) NSTEP[N] = 4;
        else if (NSTEP[N] > 48 && N > 1) NSTEP[N] = 48;
        else if (NSTEP[N] > 288) NSTEP[N] = 288;
      }
    } else if (IPRE == 3 || IPRE == 2) {
      ADDRUN = 0.0f;
      for (int n = 2; n <= 6; ++n) EXWAT[n] = 0.0f;
      if (IPRE == 3) {
        for (int J = 1; J <= NSEG; ++J) {
          SWUL[J] = SWULI[J]; DSWUL[J] = SWULI[J]; DCHG[J] = 0.0; BALT[J] = 0.0; BALY[J] = 0.0;
        }
      } else {
        layer_water(SWULL);
        for (int J = 1; J <= NSEG; ++J) { DCHG[J] = 0.0; BALT[J] = 0.0; BALY[J] = 0.0; }
      }
    }
  }

  void RUNOFF() {                                             // F:2752-2796
    if (RAIN > 0.0f) {
      r4 WTSM = 0.0f;
      for (int I = 1; I <= 7; ++I) {
        r4 sw = SWUL[I];
        r4 SWAMCI = (FCUL[I] + WPUL[I]) / 2.0f;
        if (SWUL[I] > UL[I]) sw = UL[I];
        if (SWUL[I] < SWAMCI) sw = SWAMCI;
        WTSM = WTSM + (WF[I] * (sw - SWAMCI) / (UL[I] - SWAMCI));
      }
      r4 S = SMX * (1.0f - WTSM);
      if (S < 0.0f) S = 0.0f;
      r4 AB = 0.2f * S;
      if (RAIN > AB) {
        RUN = (pw2(RAIN - AB)) / (RAIN + 0.8f * S);
        RUN = RUN * FRUNOF;
      } else RUN = 0.0f;
    } else RUN = 0.0f;
  }

  static r4 AWF(r4 D) {                                       // F:2997-2998
    return 4.759177E-4f + 3.14529f * D - 4.319952f * D * D + 3.089057f * D * D * D -
           0.9150572f * D * D * D * D;
  }
  static r4 WFS(r4 D1, r4 D2, r4 D3) {                        // F:2981-3015
    r4 R1 = (D1 < D3) ? D1 / D3 : 1.0f;
    r4 R2 = (D2 < D3) ? D2 / D3 : 1.0f;
    if (R1 < R2) return AWF(R2) - AWF(R1);
    return 0.0f;
  }

  void ETCHK(r4& WS) {                                        // F:2803-2976
    r4 ESJ[8], ULL[8], EPJ[8], AW[8], XP[8];
    r4 TET = ES + EP;
    if (TET <= 0.0f) {
      ETT = ESS; WS = 0.f;
      for (int J = 1; J <= 7; ++J) ET[J] = 0.0f;
      return;
    }
    r4 ESX = ES, XES = 0.0f, THCK = 0.0f, OTHCK = 0.0f;
    r4 DRINF = PINF, XPINF;
    for (int J = 1; J <= 7; ++J) {
      XP[J] = 0.0f; ESJ[J] = 0.0f;
      ULL[J] = FCUL[J] - WPUL[J];
      ST[J] = (r4)(SWUL[J] + (DCHG[J] / 2.f) - WPUL[J] + DRINF);
      if (ST[J] < 0.0f) ST[J] = 0.0f;
      if (ST[J] > UL[J]) { XPINF = ST[J] - UL[J]; ST[J] = UL[J]; }
      else XPINF = 0.0f;
      if (OTHCK < SEDEP) {
        if (XPINF > RCUL[J]) {
          XPINF = XPINF - RCUL[J];
          if (XPINF < ESX) { ESX = ESX - XPINF; ESJ[J] = XPINF; XP[J] = XPINF; }
          else { ESJ[J] = ESX; XP[J] = ESX; ESX = 0.0f; }
          DRINF = RCUL[J];
        } else DRINF = XPINF;
        if (ESX > 0.0f) {
          THCK = THCK + STHICK[J];
          r4 ESJO = WFS(OTHCK, THCK, SEDEP) * ES + XES;
          if (ESJO > ESX) ESJO = ESX;
          if (THCK <= SEDEP) {
            if (ST[J] < ESJO) { XES = ESJO - ST[J]; ESX = ESX - ST[J]; ESJ[J] = ST[J] + ESJ[J]; }
            else { XES = 0.0f; ESX = ESX - ESJO; ESJ[J] = ESJO + ESJ[J]; }
          } else {
            r4 RTO = ST[J] * (SEDEP - OTHCK) / (STHICK[J]);
            if (RTO < ESJO) { XES = ESJO - RTO; ESX = ESX - RTO; ESJ[J] = RTO + ESJ[J]; }
            else { XES = 0.0f; ESX = ESX - ESJO; ESJ[J] = ESJO + ESJ[J]; }
            ES = ES - ESX;
            ES1T = ES1T - ESX;
            if (ES1T <= 0.0f) ES1T = 0.0f;
            ESX = 0.0f;
          }
        }
        OTHCK = THCK;
      }
    }
    for (int J = 1; J <= 7; ++J) {
      EPJ[J] = 0.0f;
      AW[J] = ST[J] - ESJ[J] + XP[J];
      if (AW[J] < 0.0f) AW[J] = 0.0f;
    }
    if (EP > 0.0f) {
      r4 SUM = 0.0f, EPX = 0.0f, EPXX = 0.0f, UL4;
      for (int J = 1; J <= 7; ++J) {
        EPXX = EP * WF[J];
        EPXX = SUM + EPXX;
        EPJ[J] = EPXX;
        if (AW[J] > ULL[J]) UL4 = (AW[J] - ULL[J]) + ULL[J] / 4.f;
        else UL4 = ULL[J] / 4.f;
        if (UL4 < EPXX) { if (UL4 <= AW[J]) EPJ[J] = UL4; else EPJ[J] = AW[J]; }
        else { if (EPXX < AW[J]) EPJ[J] = EPXX; else EPJ[J] = AW[J]; }
        EPX = EPX + EPJ[J];
        SUM = EPXX - EPJ[J];
      }
      EP = EPX;
    }
    ETT = ESS;
    for (int J = 1; J <= 7; ++J) { ET[J] = ESJ[J] + EPJ[J]; ETT = ETT + ET[J]; }
    WS = (ETT - ESS) / TET;
  }

  void ETCOEF() {                                             // F:3023-3046
    CONA = 0.0f;
    for (int J = 1; J <= 7; ++J) CONA = CONA + (WF[J] * CONUL[J]);
    if (CONA < 3.3f) CONA = 3.3f;
    if (CONA > 5.1f) CONA = 5.1f;
    STAGE1 = (23.f * R4_POW(CONA - 3.0f, 0.42f)) / 25.4f;
  }

  void POTET(r4& ETO1, r4& ETO2) {                            // F:3327-3436
    const r4 SALB = 0.23f;
    if (CV > 40000.f) CV = 40000.f;
    EAJ = R4_EXP(-0.000029f * (CV + .1f));
    r4 ALB;
    if (SNO <= 5.f / 25.4f) {
      ALB = SALB;
      if (XLAI > 0.0f) ALB = .23f * (1.f - EAJ) + SALB * EAJ;
    } else ALB = .6f;
    r4 TK = ((TMPF[IDA] - 32.0f) * 5.0f / 9.0f) + 273.0f;
    r4 TC = TK - 273.f;
    r4 SVP = 33.8639f * (R4_POW8(0.00738f * TC + 0.8072f) - 0.000019f * fabsf(1.8f * TC + 48.f) + 0.001316f);
    r4 AVP = (RH[IDA] / 100.f) * SVP;
    r4 DELTA = 33.8639f * (0.05904f * R4_POW7(0.00738f * TC + 0.8072f) - 0.0000342f);
    r4 RBO = 11.71E-8f * (0.39f - 0.05f * R4_SQRT(AVP)) * R4_POW4(TK);
    r4 XLAT = ULAT * 6.2832f / 360.f;
    r4 XI = (r4)IDA;
    r4 SD = 0.4102f * R4_SIN(0.0172f * (XI - 80.25f));
    r4 CH = -R4_TAN(XLAT) * R4_TAN(SD);
    r4 H = 0.0f;
    if (CH <= 1.0f && CH >= -1.0f) H = R4_ACOS(CH);
    else if (CH > 1.0f) H = 0.0f;
    else if (CH < -1.0f) H = 3.1416f;
    r4 DDD = 1.0f + 0.0335f * R4_SIN(0.0172f * (XI + 88.2f));
    r4 RSO = 889.2305f * DDD * ((H * R4_SIN(XLAT) * R4_SIN(SD)) + (R4_COS(SD) * R4_COS(XLAT) * R4_SIN(H)));
    RSO = RSO * 0.8f;
    r4 ACOEF, BCOEF;
    if (RH[IDA] < 50) { ACOEF = 1.2f; BCOEF = -0.2f; }
    else if (RH[IDA] < 75) { ACOEF = 1.1f; BCOEF = -0.1f; }
    else { ACOEF = 1.0f; BCOEF = 0.0f; }
    r4 RB = (ACOEF * RAD[IDA] / RSO + BCOEF) * RBO;
    r4 RN = (1 - ALB) * RAD[IDA] - RB;
    if (RN < 0.f) RN = 0.f;
    r4 WINDF = 1.0f + 0.2394f * WIND;
    r4 TERM1 = (DELTA / (DELTA + 0.68f)) * RN;
    r4 TERM2 = (0.68f / (0.68f + DELTA)) * 15.36f * WINDF * (SVP - AVP);
    ETO = (TERM1 + TERM2) / (58.3f * 25.4f);
    ETO1 = TERM1 / (58.3f * 25.4f);
    ETO2 = TERM2 / (58.3f * 25.4f);
    if (ETO < 0.0f) ETO = 0.0f;
  }

  void EVAPOT() {                                             // F:3054-3320
    RAIN = RAIN - SMELT;
    r4 ESNO = 0.0f, EABS = 0.0f, ESM = 0.0f, EGM = 0.0f, EMELT = 0.0f, EADD = 0.0f, ES1 = 0.0f, ES2 = 0.0f;
    ESS = 0.0f; EP = 0.0f; ES = 0.0f; ABST = 0.0f;
    r4 TA = (TMPF[IDA] - 32.f) * 0.55556f;
    r4 ETO1, ETO2;
    POTET(ETO1, ETO2);
    r4 TDEST = TA;
    if (PRE[IDA] == 0.0f)
      TDEST = (R4_POW(RH[IDA] / 100.f, 1.f / 8.f)) * (112.f + (0.9f * TA)) - 112.f + (0.1f * TA);
    if (SNO > 0.0f && TDEST >= TINDEX) {
      PINF = RAIN + SMELT + ADDRUN + GMRO - RUN;
      if (PINF < 0.0f) { RUN = RUN + PINF; if (RUN < 0.0f) RUN = 0.0f; }
      ES1T = ES1T - PINF;
      if (ES1T <= 0.0f) ES1T = 0.0f;
      return;
    }
    r4 eta = ETO;                    // local ETA (not the /BLK13/ annual accumulator)
    if (SNO > 0.0f || SMELT > 0.0f) {
      if (SMELT > 0.0f) {
        ESM = ESM + (0.1367f * SMELT) + (0.0240f * SMELT);
        eta = ETO - ESM;
        if (ESM > ETO) { eta = 0.0f; ESM = ETO; }
        EMELT = eta;
        if (EMELT > SMELT) { EMELT = SMELT; eta = eta - SMELT; SMELT = 0.0f; }
        else { eta = 0.0f; SMELT = SMELT - EMELT; }
      } else { EMELT = 0.0f; eta = ETO; }
      ESNO = 0.8615f * eta;
      if (ESNO >= SNO) {
        ESNO = SNO; SNO = 0.0f;
        ESM = (ESNO * (0.1367f + 0.0240f)) + ESM;
        eta = eta - (ESNO / 0.8615f);
        WE = 0.f; XNEGHS = 0.0f; TINDEX = 0.0f;
        for (int N = 1; N <= 4; ++N) EXLAG[N] = 0.0f;
        XLIQW = 0.f; STORGE = 0.f;
      } else {
        r4 PRSNO = SNO;
        SNO = SNO - ESNO;
        ESM = (ESNO * (0.1367f + 0.0240f)) + ESM;
        eta = 0.0f;
        WE = SNO / PRSNO * WE;
        XNEGHS = SNO / PRSNO * XNEGHS;
        XLIQW = SNO / PRSNO * XLIQW;
        STORGE = SNO / PRSNO * STORGE;
        r4 TEX = 0.0f;
        for (int J = 1; J <= 4; ++J) { EXLAG[J] = SNO / PRSNO * EXLAG[J]; TEX = TEX + EXLAG[J]; }
        TWE = WE + XLIQW + STORGE + TEX;
      }
    } else {
      if (RAIN > 0.0f) {
        r4 ABSMAX = 0.05f;
        if (CV < 14000.f) ABSMAX = 0.05f * (CV + .1f) / 14000.f;
        if ((RAIN / ABSMAX) < 88.f) ABST = ABSMAX * (1 - R4_EXP(-RAIN / ABSMAX));
        else ABST = ABSMAX;
      }
      EABS = eta;
      if (EABS >= ABST) { EABS = ABST; eta = eta - ABST; }
      else { ABST = EABS; eta = 0.0f; }
    }
    if (ADDRUN > 0.0f) {
      EADD = eta;
      if (EADD > ADDRUN) { EADD = ADDRUN; eta = eta - EADD; ADDRUN = 0.0f; }
      else { ADDRUN = ADDRUN - EADD; eta = 0.0f; }
    }
    if (GMRO > 0.0f) {
      EGM = eta;
      if (EGM > GMRO) { EGM = GMRO; eta = eta - EGM; GMRO = 0.0f; }
      else { GMRO = GMRO - EGM; eta = 0.0f; }
    }
    PINF = RAIN + SMELT + ADDRUN + GMRO - RUN - ABST;
    if (PINF < 0.0f) {
      RUN = RUN + PINF;
      PINF = 0.0f;
      if (RUN < 0.0f) { ABST = ABST + RUN; RUN = 0.0f; EABS = ABST; }
    }
    ESS = EMELT + ESNO + EABS + EADD + EGM;
    if (eta <= 0.0f || (ETO - ESS - ESM) <= 0.0f) {
      ES1T = ES1T - PINF;
      if (ES1T <= 0.0f) ES1T = 0.0f;
      return;
    }
    if (TMPF[IDA] >= 23.0f && IFREZ != 1) {
      r4 ETO2X = (1.0f - (7.143E-05f * CV)) * ETO2;
      if (ETO2X < 0.0f) ETO2X = 0.0f;
      r4 ESO = EAJ * (ETO1 + ETO2X);
      if (ESO > eta) ESO = eta;
      if (ESO < 0.0f) ESO = 0.0f;
      if (ES1T <= STAGE1) {
        ES1 = eta;
        if (ES1 >= ESO) { ES1 = ESO; eta = eta - ESO; }
        else if (ES1 <= 0.0f) { ES1 = 0.0f; eta = 0.0f; }
        else eta = eta - ES1;
        ES1T = ES1T + ES1 - PINF;
        if (ES1T <= 0.0f) ES1T = 0.0f;
        if (ES1T <= STAGE1) TS2 = 0.0f;
        ES = ES1;
      } else {
        TS2 = TS2 + 1.0f;
        ES2 = CONA * (R4_SQRT(TS2) - R4_SQRT(TS2 - 1.0f)) / 25.4f;
        if (ES2 >= ESO) { ES2 = ESO; eta = 0.0f; }
        else if (ES2 <= 0.0f) ES2 = 0.0f;
        else eta = eta - ES2;
        ES1T = ES1T + ES2 - PINF;
        if (ES1T <= 0.0f) ES1T = 0.0f;
        ES = ES2;
      }
      EP = ETO * XLAI / 3.0f;
      if (EP > eta) { EP = eta; eta = 0.0f; }
      else if (EP < 0.0f) EP = 0.0f;
      else eta = eta - EP;
      if (EP >= (ETO - ESM - ESS - ES)) EP = ETO - ESM - ES - ESS;
      if (EP < 0.0f) { ES = ES + EP; EP = 0.0f; }
    } else {
      ES1T = ES1T - PINF;
      if (ES1T <= 0.0f) ES1T = 0.0f;
    }
  }

  void TSTR(r4& TGX, r4 TMPC) {                               // F:3568-3582
    if (TMPC > TO) TGX = 2.f * TO - TB - TMPC;
    r4 RTO = pw2((TO - TMPC) / TGX);
    if (RTO > 200.f) TS = 0.f;
    else TS = R4_EXP(-0.1054f * RTO);
    if ((TMPC - 5.f) <= (AVT - 15.f)) TS = 0.f;
  }

  void CRPMOD() {                                             // F:3440-3512
    r4 WS = 0.f;
    r4 SUT = ST[7] * 25.4f / FCS[7];
    r4 CDG = (.9f * TSEG[7] / (TSEG[7] + R4_EXP(9.93f - (.312f * TSEG[7])))) + .1f;
    r4 DECR = .05f * fminf(CDG, SUT);
    RSD = RSD * (1.f - DECR);
    if (IPL == 0 && IDA == 30) GS = 0.f;
    if (IPL == 367 && IDA == 210) GS = 0.f;
    if (IPL == 0 && IHV == 367) DM = DM * (1.f - DECR);
    if (IPL == 367 && IHV == 0) DM = DM * (1.f - DECR);
    CV = .8f * DM + RSD;
    int d;
    if (IPL >= IHV) goto L1000;
    d = IDA - IPL;
    if (d < 0) goto L1010; else if (d == 0) goto L1000; else goto L1020;
  L1010:
    ETCHK(WS);
    return;
  L1000:
    d = IDA - IPL;
    if (d < 0) goto L1020; else if (d == 0) goto L1030; else goto L1040;
  L1030:
    GS = 0.f; DM = 0.f;
    goto L1040;
  L1020:
    d = IDA - IHV;
    if (d < 0) goto L1040; else if (d == 0) goto L1050; else goto L1010;
  L1050:
    RSD = (0.53f * DM + RSD);
    DM = 0.f; XLAI = 0.f;
    ETCHK(WS);
    goto L1060;
  L1040: {
      ETCHK(WS);
      r4 TMPC = (TMPF[IDA] - 32.f) / 1.8f;
      r4 TGX = TMPC - TB;
      if (TGX > 0.f) TSTR(TGX, TMPC);
      else TS = 0.f;
      r4 DDM = .5f * RAD[IDA] * (1.f - R4_EXP((-.65f) * (XLAI + .05f)));
      r4 REG = fminf(WS, TS);
      DM = DM + DDM * REG;
      if (DM <= 0.0f) DM = 0.0f;
      if (TMPC > TB) GS = GS + (TMPC - TB) / PHU;
      if (GS > .75f) XLAI = 16.f * DLAI * (1.f - GS) * (1.f - GS);
      else {
        r4 WLV = .8f * DM;
        r4 F = WLV / (WLV + 5512 * R4_EXP((-0.000608f) * WLV));
        XLAI = ULAI * F;
        DLAI = XLAI;
      }
    }
  L1060:
    if (XLAI < 0.0f) XLAI = 0.0f;
  }

  void SOLT() {                                               // F:3516-3560
    if (CV < 0.0f) CV = 0.0f;
    if (CV > 20000.f) CV = 20000.f;
    r4 BCV = CV / (CV + R4_EXP(7.563f - (0.0001297f * CV)));
    if ((SNO * 25.4f) > 120.f) BCV = fmaxf(1.f, BCV);
    else if (SNO > 0.0f) {
      r4 XX = (SNO * 25.4f) / (SNO * 25.4f + R4_EXP(6.055f - (.3002f * SNO * 25.4f)));
      BCV = fmaxf(XX, BCV);
    }
    r4 XI = (r4)IDA;
    r4 ALX = (XI - 200.f) / 58.13f;
    r4 TMPC = (TMPF[IDA] - 32.f) / 1.8f;
    r4 ST0;
    if (RAD[IDA] == 0.f) ST0 = WFT[MO] * 5.f + TMPC;
    else ST0 = WFT[MO] * 5.f + TMPC - 5.f;
    r4 XX = BCV * TO + (1.f - BCV) * ST0;
    r4 DT = XX - (AVT + AMP * R4_COS(ALX));
    for (int K = 1; K <= 7; ++K) {
      r4 ZD = -Z[K] / DD;
      if (fabsf(ZD) < 37.f) TSEG[K] = AVT + (AMP * R4_COS(ALX + ZD) + DT) * R4_EXP(ZD);
      else TSEG[K] = AVT;
    }
  }

  void FRZCHK(r4 TFSOIL) {                                    // F:3590-3628
    TMPSUM = TMPSUM - TAVG[30] + TMPF[IDA];
    TMPAVG = TMPSUM / 30;
    for (int I = 1; I <= 29; ++I) { int J = 30 - I; TAVG[J + 1] = TAVG[J]; }
    TAVG[1] = TMPF[IDA];
    if (TMPF[IDA] < TFSOIL && KCNT > 0) KCNT = KCNT - 1;
    if (TMPF[IDA] >= TFSOIL && KCNT < MXKCNT) KCNT = KCNT + 1;
    if (IFREZ == 0) {
      if (TMPAVG < TFSOIL && KCNT == 0) { IFREZ = 1; MXKCNT = IDFS; }
    } else {
      if (KCNT >= IDFS && SNO <= 0.0f) { IFREZ = 0; MXKCNT = (IDFS + 2) / 3; KCNT = MXKCNT; }
    }
  }

  void snow_reset() {
    TWE = 0.0f; WE = 0.0f; XNEGHS = 0.0f; XLIQW = 0.0f; TINDEX = 0.0f; STORGE = 0.0f;
    for (int N = 1; N <= 4; ++N) EXLAG[N] = 0.0f;
  }
  void SNOCOEF() {                                            // F:3636-3751
    if (IPRE < 2) {
      XMFMAX = 5.2f; XMFMIN = 2.0f; XNMF = 0.6f; TIPM = 0.3f; TBASE = 0.0f; PXTEMP = 0.0f;
      PLWHC = 0.01f; GM = 0.5f; PLQW = 0.01f;
      UADJ = 0.262f * (1.0f + 0.24f * WIND);
      r4 SNEW = 1.5f, RMIN = 0.25f;
      SFNEW = SNEW * 24.f; RFMIN = RMIN * 24.f;
      snow_reset();
      if (OSNO > 0 && IPRE == 1) {
        XLIQW = (OSNO * 25.4f) * PLQW;
        WE = (OSNO * 25.4f) - XLIQW;
        SNO = OSNO;
      } else SNO = 0.0f;
      OLDSNO = OSNO;
    } else if (IPRE == 3) {
      if (OSNO > 0.0f) {
        if (OLDSNO > 0.0f) {
          WE = WE * OSNO / OLDSNO; XNEGHS = XNEGHS * OSNO / OLDSNO;
          XLIQW = XLIQW * OSNO / OLDSNO; STORGE = STORGE * OSNO / OLDSNO;
          for (int I = 1; I <= 4; ++I) EXLAG[I] = EXLAG[I] * OSNO / OLDSNO;
        } else {
          snow_reset();
          XLIQW = (OSNO * 25.4f) * PLQW;
          WE = (OSNO * 25.4f) - XLIQW;
        }
      } else snow_reset();
      SNO = OSNO; OLDSNO = OSNO;
    } else if (IPRE == 2) {
      OSNO = OLDSNO; SNO = OLDSNO;
    }
  }

  void SNOW(int& IT) {                                        // F:3756-4122
    r4 TA = (TMPF[IDA] - 32.f) * 0.55556f;
    r4 PXI = PRE[IDA] * 25.4f;
    r4 SFALL = 0.0f, CNHSPX = 0.0f, CNHS = 0.0f, ROS = 0.0f, RAINM = 0.0f, XMELT = 0.f,
       GMWLOS = 0.f, GMSLOS = 0.f;
    RAIN = 0.0f; GMRO = 0.f; SMELT = 0.f;
    r4 TEX, PACKRO, EXCESS, FIT;
    int NT;
    if (TA < PXTEMP) {
      XLIQW = XLIQW + ADDRUN * 25.4f;
      SNO = SNO + ADDRUN;
      TWE = TWE + ADDRUN * 25.4f;
      ADDRUN = 0.0f;
      r4 ts = TA;                    // local TS (not the /BLK21/ crop stress factor)
      if (ts > 0.0f) ts = 0.0f;
      SFALL = PXI;
      IT = 1;
      WE = WE + SFALL;
      if (WE == 0.0f) goto L380;
      CNHSPX = -ts * SFALL / 160.0f;
      if (SFALL > SFNEW) TINDEX = ts;
    } else {
      RAIN = PXI;
      IT = 2;
      if (WE == 0.0f) goto L380;
    }
    {
      r4 TR = TA;
      if (TR < 0.0f) TR = 0.0f;
      RAINM = 0.0125f * RAIN * TR;
      r4 GMC = GM;
      if (IFREZ == 1) GMC = 0.0f;
      if (WE <= GMC) { GMRO = WE + XLIQW; goto L350; }
      GMWLOS = (GMC / WE) * XLIQW;
      GMSLOS = GMC;
      {  // scope so the jumps to L350 do not cross initialisations
      ND = LEAP(JYEAR[IYR], NT);
      int IDANG;
      if (ULAT >= 0.0f) { IDANG = 80; if (ND == 366) IDANG = 81; }
      else { IDANG = 263; if (ND == 366) IDANG = 264; }
      int IDN = IDA - IDANG;
      if (IDN <= 0) IDN = IDN + ND;
      r4 DIFF = (XMFMAX - XMFMIN);
      r4 DAYN = (r4)IDN;
      r4 X;
      if (IDN >= 275) X = (DAYN - 275.0f) / (458.0f - 275.0f);
      else if (IDN >= 92) X = (275.0f - DAYN) / (275.0f - 92.0f);
      else X = (91.0f + DAYN) / 183.0f;
      r4 XX = (R4_SIN(DAYN * 2.0f * 3.1416f / 366.0f) * 0.5f) + 0.5f;
      r4 ADJMF;
      if (X <= 0.48f) ADJMF = 0.0f;
      else if (X >= 0.70f) ADJMF = 1.0f;
      else ADJMF = (X - 0.48f) / (0.70f - 0.48f);
      r4 RMF60 = (XX * ADJMF * DIFF) + XMFMIN;
      r4 XMF = RMF60;
      if (fabsf(ULAT) < 60.0f) {
        r4 RMF50 = (R4_SIN(DAYN * 2.0f * 3.1416f / 366.0f) * DIFF * 0.5f) + (XMFMAX + XMFMIN) * 0.5f;
        XMF = RMF50;
        if (fabsf(ULAT) > 50.0f) {
          r4 RINT = (fabsf(ULAT) - 50.f) / 10.f;
          XMF = RMF50 + RINT * (RMF60 - RMF50);
        }
      }
      r4 RATIO = XMF / XMFMAX;
      r4 TMX = TA - TBASE;
      if (TMX < 0.0f) TMX = 0.0f;
      r4 TSUR = TA;
      if (TSUR > 0.0f) TSUR = 0.0f;
      r4 TNMX = TINDEX - TSUR;
      r4 XNMRATE = RATIO * XNMF;
      CNHS = XNMRATE * TNMX;
      TINDEX = TINDEX + TIPM * (TA - TINDEX);
      if (TINDEX > 0.0f) TINDEX = 0.0f;
      if (TMX > 0.0f) XMELT = XMF * TMX;
      if (RAIN <= RFMIN) XMELT = XMELT + RAINM;
      else {
        r4 SVP1 = R4_POW8(0.00738f * TA + 0.8072f);
        r4 SVP2 = 0.000019f * fabsf(1.8f * TA + 48.f);
        r4 SVP = 33.8639f * (SVP1 - SVP2 + 0.001316f);
        r4 AVP = 0.9f * SVP;
        r4 TK = (TA + 273.f);
        r4 QN = (11.71E-8f * pw4(TK)) / 7.97f - 81.61f;
        r4 QE = 8.5f * (AVP - 6.11f) * UADJ;
        r4 BR = 0.68f * TA / (AVP - 6.11f);
        r4 QH = BR * QE;
        XMELT = QN + QE + QH + RAINM;
        if (XMELT < 0.0f) XMELT = 0.0f;
      }
      ROS = RAIN;
      RAIN = 0.f;
      if ((CNHS + XNEGHS) < 0.0f) CNHS = -1.0f * XNEGHS;
      WE = WE - GMSLOS;
      XLIQW = XLIQW - GMWLOS;
      GMRO = GMSLOS + GMWLOS;
      if (XMELT > 0.0f) {
        if (XMELT >= WE) { XMELT = WE + XLIQW; goto L350; }
        else WE = WE - XMELT;
      }
      r4 WATER = XMELT + ROS;
      r4 HEAT = CNHS + CNHSPX;
      r4 XLIQWMX = PLWHC * WE;
      XNEGHS = XNEGHS + HEAT;
      if (XNEGHS < 0.0f) XNEGHS = 0.0f;
      if (XNEGHS > 0.33f * WE) XNEGHS = 0.33f * WE;
      if ((WATER + XLIQW) >= (XLIQWMX + XNEGHS)) {
        EXCESS = WATER + XLIQW - XLIQWMX - XNEGHS;
        XLIQW = XLIQWMX; WE = WE + XNEGHS; XNEGHS = 0.0f;
      } else if (WATER >= XNEGHS) {
        XLIQW = XLIQW + WATER - XNEGHS; WE = WE + XNEGHS; XNEGHS = 0.0f; EXCESS = 0.0f;
      } else {
        WE = WE + WATER; XNEGHS = XNEGHS - WATER; EXCESS = 0.0f;
      }
      if (XNEGHS == 0.0f) TINDEX = 0.0f;
      PACKRO = 0.0f;
      FIT = 24;
      if (EXCESS != 0.0f) {
        if (EXCESS >= 0.1f && WE >= 1.0f) {
          int N = (int)((R4_POW(EXCESS * 4.0f, 0.3f)) + 0.5f);
          if (N == 0) N = 1;
          r4 FN = (r4)N;
          for (int I = 1; I <= N; ++I) {
            r4 FI = (r4)I;
            r4 TERM = 0.03f * WE * FN / (EXCESS * (FI - 0.5f));
            if (TERM > 50.0f) TERM = 50.0f;
            r4 FLAG = 5.33f * (1.0f - R4_EXP(-TERM));
            int L2 = (int)((FLAG + FIT) / FIT + 1.0f);
            int L1 = L2 - 1;
            r4 ENDL1 = (r4)(L1 * 24);
            r4 POR2 = (FLAG + FIT - ENDL1) / FIT;
            r4 POR1 = 1.0f - POR2;
            EXLAG[L2] = EXLAG[L2] + POR2 * EXCESS / FN;
            EXLAG[L1] = EXLAG[L1] + POR1 * EXCESS / FN;
          }
        } else EXLAG[1] = EXLAG[1] + EXCESS;
      }
      if ((STORGE + EXLAG[1]) != 0.0f) {
        if ((STORGE + EXLAG[1]) < 0.1f) { PACKRO = STORGE + EXLAG[1]; STORGE = 0.0f; }
        else {
          r4 EL = EXLAG[1] / FIT;
          r4 ELS = EL / 25.4f;
          r4 WES = WE / 25.4f;
          r4 TERM = 500.0f * ELS / (R4_POW(WES, 1.3f));
          if (TERM > 50.0f) TERM = 50.0f;
          r4 R1 = 1.0f / (5.0f * R4_EXP(-TERM) + 1.0f);
          for (int I = 1; I <= 24; ++I) {
            r4 OS = (STORGE + EL) * R1;
            PACKRO = PACKRO + OS;
            STORGE = STORGE + EL - OS;
          }
          if (STORGE <= 0.001f) { PACKRO = PACKRO + STORGE; STORGE = 0.0f; }
        }
      }
      for (int I = 2; I <= 4; ++I) EXLAG[I - 1] = EXLAG[I];
      EXLAG[4] = 0.0f;
      SMELT = PACKRO;
      goto L380;
      }
    L350:
      TEX = 0.0f;
      for (int N = 1; N <= 4; ++N) TEX = TEX + EXLAG[N];
      SMELT = XMELT + TEX + STORGE + ROS;
      WE = 0.0f; XNEGHS = 0.0f; XLIQW = 0.0f; TINDEX = 0.0f; STORGE = 0.0f;
      for (int N = 1; N <= 4; ++N) EXLAG[N] = 0.0f;
    }
  L380:
    TEX = 0.0f;
    for (int N = 1; N <= 4; ++N) TEX = TEX + EXLAG[N];
    TWE = WE + XLIQW + STORGE + TEX;
    GMRO = GMRO / 25.4f;
    SMELT = SMELT / 25.4f;
    RAIN = RAIN / 25.4f + SMELT;
    SNO = TWE / 25.4f;
  }

  // ---- subsurface routing (REAL*8) -------------------------------------------
  void HEAD(r8& TH, int NSEGBv, int NSEGL) {                  // F:4582-4617
    r8 XHEAD = 0.0;
    TH = 0.0;
    for (int J = NSEGBv; J <= NSEGL; ++J) {
      int K = NSEGL + NSEGBv - J;
      if ((SWULY[K] + (BALY[K] / 2.0)) <= DFCUL[K]) return;
      XHEAD = DTHICK[K] * (SWULY[K] + (BALY[K] / 2.0) - DFCUL[K]) / (DUL[K] - DFCUL[K]);
      if (XHEAD > DTHICK[K]) XHEAD = DTHICK[K] + SWULY[K] + (BALY[K] / 2.0) - DUL[K];
      TH = TH + XHEAD;
      XHEAD = XHEAD - DTHICK[K];
      if ((XHEAD + 0.00005) < 0.0) return;
    }
  }

  void LATKS(r8 TH, int NSEGBv, int NSEGL, r8& ELKS) {        // F:4626-4651
    ELKS = 0.0;
    if (TH <= 0.0) return;
    r8 H = TH;
    for (int J = NSEGBv; J <= NSEGL; ++J) {
      int K = NSEGL + NSEGBv - J;
      if (H <= 0.0) return;
      if (!(H > DTHICK[K])) { ELKS = ELKS + DRCUL[K] * H / TH; return; }
      ELKS = ELKS + DRCUL[K] * DTHICK[K] / TH;
      H = H - DTHICK[K];
    }
  }

  void LATFLO(int LAYD, r8 TH, r8 ELKS, r8& DRNv) {           // F:4658-4734
    const r8 EPS = 0.030, PI = 3.141592654;
    r8 DSLOPE = SLOPE[LAYD] / 100.0;
    r8 DXLENG = XLENG[LAYD] * 12.0;
    if (DXLENG == 0.0) DXLENG = 1.0;
    if (!(TH > 0.0)) { DRNv = 0.0; return; }
    int ITER = 1;
    r8 ALPHA = R8_ATAN(DSLOPE);
    r8 YSTAR = TH / DXLENG;
    r8 QSTAR = 2.0f * YSTAR * R8_SIN(ALPHA) * R8_COS(ALPHA);
    if (ALPHA <= 0.f) {
      QSTAR = dpw2((4.0 / PI) * YSTAR);
      DRNv = QSTAR * ELKS;
      return;
    }
    r8 A = PI / 4.0 / R8_COS(ALPHA);
    r8 B = 2.0 * R8_SQRT(0.4) / PI;
    r8 C = 0.4 * dpw2(R8_SIN(ALPHA));
    r8 D = 0.5 / R8_LOG(B);
    if (YSTAR < (0.2f * DSLOPE)) { DRNv = QSTAR * ELKS; return; }
    r4 QSTARN;                       // implicit REAL*4 in the Fortran
    for (;;) {
      r8 Fq = (A * (R8_POW(B, R8_POW(QSTAR / C, D))));
      QSTARN = (r4)((YSTAR * YSTAR) / (Fq * Fq));
      r8 TOLER = 2.0f * fabs(QSTARN - QSTAR) / (QSTARN + QSTAR);
      if (EPS < TOLER && ITER < 10) {
        ITER = ITER + 1;
        QSTAR = (QSTAR + QSTARN) / 2.0;
        continue;
      }
      break;
    }
    DRNv = (QSTAR + QSTARN) * ELKS / 2.0;
  }

  void FML(r8 TH, r8& QPERC, int NSEGE, int NPROF) {          // F:4870-5304
    r8 QVAPOR = 0.0, RPHOLE = 0.0, RDEFEC = 0.0, HPHOLE = 0.0, HDEFEC = 0.0;
    r4 QPHOLE = 0.0f, QDEFEC = 0.0f, RSQR, DENOM;
    int LINER = 0;
    if (LAYER[LP[NPROF]] == 4) LINER = LP[NPROF];
    if (LAYER[LP[NPROF] - 1] == 4) LINER = LP[NPROF] - 1;
    const int LC = LCASE[NPROF], PQ = LPQ[NPROF];
    const r8 RCS = RC[LAYSEG[NSEGE][1]];
    const r8 DTH = DTHICK[NSEGE];
    if (TH > THICK[LINER]) QVAPOR = 34016.0 * RC[LINER] * TH / THICK[LINER];
    else if (LC == 3 || LC == 5) { if (TH > 0.0) QVAPOR = 34016.0 * RC[LINER]; else QVAPOR = 0.0; }
    else QVAPOR = 34016.0 * RC[LINER];
    if (TH > 0.0) {
      // Giroud radius-of-wetted-area form shared by the "contact" cases
      auto giroud = [&](r4 cp, r4 cd, r4 ex_h, r4 ex_k, r8 head) {
        if (PHOLE[LINER] > 0.0f) {
          RPHOLE = cp * (R8_POW(head, ex_h)) * (R8_POW(RCS, ex_k));
          HPHOLE = 1 + (head / (2.f * DTH * R8_LOG(RPHOLE / 0.0005f)));
          QPHOLE = (r4)(PHOLE[LINER] * 26.4f * RCS * HPHOLE * (dpw2(RPHOLE)));
        }
        if (DEFEC[LINER] > 0.0f) {
          RDEFEC = cd * (R8_POW(head, ex_h)) * (R8_POW(RCS, ex_k));
          HDEFEC = 1 + (head / (2.f * DTH * R8_LOG(RDEFEC / 0.00564f)));
          QDEFEC = (r4)(DEFEC[LINER] * 26.4f * RCS * HDEFEC * (dpw2(RDEFEC)));
        }
      };
      if (LC == 1) {
        QPHOLE = (r4)(PHOLE[LINER] * 0.000177f * TH / THICK[LINER]);
        QDEFEC = (r4)(DEFEC[LINER] * 0.0356f * R8_SQRT(TH));
      }
      if (LC == 2) {
        if (PQ == 1) {
          QPHOLE = (r4)(PHOLE[LINER] * 0.000671f * TH * RCS);
          QDEFEC = (r4)(DEFEC[LINER] * 0.00758f * TH * RCS);
        }
        if (PQ > 1 && PQ < 5) giroud(0.00363f, 0.0229f, 0.38f, -0.25f, TH);
        if (PQ == 5) {
          QPHOLE = (r4)(PHOLE[LINER] * 0.000177f * TH / THICK[LINER]);
          QDEFEC = (r4)(DEFEC[LINER] * 0.0356f * R8_SQRT(TH));
        }
      }
      if (LC == 3) {
        if (PQ == 1) {
          QPHOLE = (r4)(PHOLE[LINER] * 0.000671f * TH * RCS);
          QDEFEC = (r4)(DEFEC[LINER] * 0.00758f * TH * RCS);
        }
        if (PQ == 2) giroud(0.052f, 0.0663f, 0.5f, -0.06f, TH);
        if (PQ == 3) giroud(0.0449f, 0.0572f, 0.45f, -0.13f, TH);
        if (PQ == 4) giroud(0.1053f, 0.134f, 0.45f, -0.13f, TH);
        if (PQ == 5) {
          QPHOLE = (r4)(PHOLE[LINER] * 0.000177f * TH / THICK[LINER]);
          QDEFEC = (r4)(DEFEC[LINER] * 0.0356f * R8_SQRT(TH));
        }
      }
      if (LC == 4) {
        if (PQ == 1) {
          QPHOLE = (r4)(PHOLE[LINER] * 0.000671f * (TH + DTH) * RCS);
          QDEFEC = (r4)(DEFEC[LINER] * 0.00758f * (TH + DTH) * RCS);
        }
        if (PQ == 2) giroud(0.0520f, 0.0663f, 0.5f, -0.06f, TH + DTH);
        if (PQ == 3) giroud(0.0449f, 0.0572f, 0.45f, -0.13f, TH + DTH);
        if (PQ == 4) giroud(0.1053f, 0.134f, 0.45f, -0.13f, TH + DTH);
        if (PQ == 5) {
          QPHOLE = (r4)(PHOLE[LINER] * 0.000177f * (TH + DTH / THICK[LINER]));   // sic (F:5167-5168)
          QDEFEC = (r4)(DEFEC[LINER] * 0.0356f * R8_SQRT((TH + DTH)));
        }
      }
      if (LC == 5 || LC == 6) {
        const r8 head = (LC == 5) ? TH : (TH + DTH);
        RPHOLE = 0.001f;
        RDEFEC = 0.0113f;
        if (PHOLE[LINER] > 0.0f) {
          if (((RCS / 100.f) * 0.000001f) < (4.0f * TRANS[LINER] * head / 10000.f / 39.37f)) {
            for (int I = 1; I <= 10; ++I) {
              RSQR = (r4)(4.0f * TRANS[LINER] * head / 10000.f / 39.37f);
              DENOM = (r4)(2.f * R8_LOG(RPHOLE / 0.0005f));
              DENOM = (r4)(DENOM + (dpw2(0.0005f / RPHOLE)) - 1.0f);
              RSQR = (r4)((RSQR * 100.f) / RCS / DENOM);
              RPHOLE = R4_SQRTF(RSQR);
            }
          }
          if (LC == 5) HPHOLE = 1 + (head / (2.f * DTH * R8_LOG(RPHOLE / 0.0005f)));
          else HPHOLE = 1 + (head / (2 * DTH * R8_LOG(RPHOLE / 0.0005f)));
          QPHOLE = (r4)(PHOLE[LINER] * 26.4f * RCS * HPHOLE * dpw2(RPHOLE));
        }
        if (DEFEC[LINER] > 0.0f) {
          if (((RCS / 100.f) * 0.000127f) < (4.0f * TRANS[LINER] * head / 10000.f / 39.37f)) {
            const r4 cdef = (LC == 5) ? 0.00564f : 0.00565f;   // sic (F:5214 vs F:5271)
            for (int I = 1; I <= 10; ++I) {
              RSQR = (r4)(4.0f * TRANS[LINER] * head / 10000.f / 39.37f);
              DENOM = (r4)(2.f * R8_LOG(RDEFEC / 0.00564f));
              DENOM = (r4)(DENOM + (dpw2(cdef / RDEFEC)) - 1.0f);
              RSQR = (r4)((RSQR * 100.f) / RCS / DENOM);
              RDEFEC = R4_SQRTF(RSQR);
            }
          }
          HDEFEC = 1 + (head / (2.f * DTH * R8_LOG(RDEFEC / 0.00564f)));
          QDEFEC = (r4)(DEFEC[LINER] * 26.4f * RCS * HDEFEC * dpw2(RDEFEC));
        }
      }
    }
    QPERC = QVAPOR + QPHOLE + QDEFEC;
    if ((LC == 3 && LAYER[LP[NPROF]] == 3) || LC == 4) {
      r8 cap = 34016.0 * RCS * (TH + DTH) / DTH;
      if (QPERC > cap) QPERC = cap;
    } else {
      if (QPERC > (34016.0 * RCS)) QPERC = 34016.0 * RCS;
    }
  }

  void PROFIL(r8& EXCESS, int NSEGBv, int NSEGL) {            // F:4535-4574
    EXCESS = 0.0;
    for (int J = NSEGBv; J <= NSEGL; ++J) {
      int K = NSEGL + NSEGBv - J;
      r8 S = DSWUL[K] + EXCESS + (BALT[K] / 2.0);
      r8 EXCE = EXCESS;
      if (EXCE < 0.0) EXCE = 0.0;
      EXCESS = S - DUL[K];
      if (EXCESS > 0.0) {
        BALT[K] = DUL[K] - SWULY[K] - (BALY[K] / 2.0);
        DSWUL[K] = DUL[K] - (BALT[K] / 2.0);
        DRIN[K] = DRIN[K] - EXCESS;
      } else {
        EXCESS = 0.0;
        BALT[K] = BALT[K] + EXCE;
        DSWUL[K] = SWULY[K] + ((BALY[K] + BALT[K]) / 2.0);
      }
    }
  }

  void AVROUT(const r8* E, const r8* DSUB, const r8* DRCR, r8 DT, int IBAR, int LAT,
              int NSEGBv, int NSEGL, int NSEGE, int NPROF,
              r8 QPERCY, r8 QLATY, r8& QPERC, r8& QLAT) {     // F:4287-4527
    r8 DRNMX, DRMX, GRVWAT, DRNMAX, A, B, C, X0, FX0, FX0DER, DX, LOW, HIGH, SWMAX, SWMIN, TOLER;
    r4 SWLIM, RELMIN, RELSW, PSIJP1, RELSWL, sw, TRATIO;      // implicit REAL*4 (F:4291-4297)
    int IFLAG = 0;
    for (int J = NSEGBv; J <= NSEGL; ++J) {
      bool capillary = false;
      if (J <= 7) {
        SWLIM = (r4)DWPUL[J];
        if (DRIN[J] == 0.0 && J < NSEGL) capillary = true;
      } else if (J == NSEGL && J != NSEG) SWLIM = (r4)DFCUL[J];
      else if (J == NSEG && DRIN[J] > 0.0) SWLIM = (r4)DWPUL[J];
      else if (J == NSEG && DRIN[J] <= 0.0) SWLIM = (r4)DFCUL[J];
      else capillary = true;
      if (capillary) {
        RELMIN = (r4)((DWPUL[J + 1] - DRSUL[J + 1]) / (DUL[J + 1] - DRSUL[J + 1]));
        RELSW = (r4)((SWULY[J + 1] + (BALY[J + 1] / 2.0) - DRSUL[J + 1]) / (DUL[J + 1] - DRSUL[J + 1]));
        if (RELSW < RELMIN) RELSW = RELMIN;
        PSIJP1 = (r4)(DBUBUL[J + 1] * (R8_POW((r8)RELSW, -1.0f / DLAMUL[J + 1])));
        if (PSIJP1 < DBUBUL[J]) PSIJP1 = (r4)DBUBUL[J];
        RELSWL = (r4)R8_POW(DBUBUL[J] / PSIJP1, DLAMUL[J]);
        SWLIM = (r4)((RELSWL * (DUL[J] - DRSUL[J])) + DRSUL[J]);
        if (SWLIM < DWPUL[J]) SWLIM = (r4)DWPUL[J];
        if (SWLIM > DFCUL[J]) SWLIM = (r4)DFCUL[J];
      }
      if (J <= 7) {
        if (J == NSEGL) SWLIM = (r4)DFCUL[J];
        if (J == NSEG) SWLIM = (r4)DFCUL[J];
        if (J == NSEG && DRIN[J] > 0.0) SWLIM = (r4)DWPUL[J];
        if (IFREZ == 1) SWLIM = (r4)DUL[J];
      }
      DRNMAX = SWULY[J] + (BALY[J] / 2.0) + DRIN[J] + DSUB[J] + DRCR[J] - E[J] - SWLIM;
      if (J == NSEGL) {
        if (NSEGE == NSEG) {
          if (IBAR == 3) DRNMAX = 0.0;
          if (IBAR > 3) DRNMAX = DRNMAX + DSUB[NSEGE];
          if (IBAR >= 3) {
            QPERC = 0.0;
            DRIN[NSEGE + 1] = 0.0;
            if (LAT == 0) { QLAT = 0.0; DRIN[NSEGL + 1] = 0.0; goto L1100; }
          }
        }
        if (IBAR != 0 && IBAR != 3) {
          GRVWAT = DRNMAX + SWLIM - DFCUL[J];
          if (GRVWAT <= 0.0) {
            GRVWAT = 0.0; QLAT = 0.0; QPERC = 0.0;
            DRIN[NSEGE + 1] = 0.0; DRIN[NSEGL + 1] = 0.0;
            if (IBAR < 3) DRIN[NSEGE + 1] = DRIN[NSEGE + 1] + DSUB[NSEGE];
            goto L1100;
          }
          if (GRVWAT >= ((QPERCY + QLATY) * DT)) {
            DRIN[NSEGL + 1] = (QPERCY + QLATY) * DT;
            DRIN[NSEGE + 1] = QPERCY * DT;
            QPERC = QPERCY; QLAT = QLATY;
          } else if (LAT == 0) {
            QLAT = 0.0; QPERC = GRVWAT / DT;
            DRIN[NSEGL + 1] = GRVWAT; DRIN[NSEGE + 1] = GRVWAT;
          } else {
            DRIN[NSEGL + 1] = GRVWAT;
            QLAT = (GRVWAT / DT) * QLATY / (QPERCY + QLATY);
            QPERC = (GRVWAT / DT) * QPERCY / (QPERCY + QLATY);
            DRIN[NSEGE + 1] = QPERC * DT;
            QPERC = GRVWAT / DT;
          }
          if (IBAR < 3) DRIN[NSEGE + 1] = DRIN[NSEGE + 1] + DSUB[NSEGE];
          goto L1100;
        }
      }
      if (DRNMAX < 0.0) DRNMAX = 0.0;
      C = (3.0 + (2.0 / DLAMUL[J]));
      LOW = DRCUL[J] * DT * R8_POW((SWLIM - DRSUL[J]) / (DUL[J] - DRSUL[J]), C);
      if (LOW > DRNMAX) { DRIN[J + 1] = DRNMAX; goto L1070; }
      SWMAX = SWULY[J] + (0.5 * BALY[J]) + DSUB[J] + DRCR[J] + DRIN[J] - E[J];
      if (J == NSEGL && NSEGE == NSEG && IBAR > 3) SWMAX = SWMAX + DSUB[NSEGE];
      if (SWMAX > DUL[J]) SWMAX = DUL[J];
      if (SWMAX <= SWLIM) { DRIN[J + 1] = 0.0; goto L1070; }
      HIGH = DRCUL[J] * DT * R8_POW((SWMAX - DRSUL[J]) / (DUL[J] - DRSUL[J]), C);
      if (HIGH > DRNMAX) HIGH = DRNMAX;
      SWMIN = SWMAX - HIGH;
      if (SWMIN < SWLIM) SWMIN = SWLIM;
      LOW = DRCUL[J] * DT * R8_POW((SWMIN - DRSUL[J]) / (DUL[J] - DRSUL[J]), C);
      if (LOW > DRNMAX) LOW = DRNMAX;
      SWMAX = SWULY[J] + (0.5 * BALY[J]) + DSUB[J] + DRCR[J] + DRIN[J] - E[J] - LOW;
      if (J == NSEGL && NSEGE == NSEG && IBAR > 3) SWMAX = SWMAX + DSUB[NSEGE];
      if (SWMAX > DUL[J]) SWMAX = DUL[J];
      if (SWMAX <= SWMIN) { DRIN[J + 1] = DRNMAX; goto L1070; }
      // Newton on f(x) = A - B x^C - 2x  (F:4461-4488)
      A = 2.0 * (SWULY[J] - DRSUL[J]) + BALY[J] + DRIN[J] + DSUB[J] + DRCR[J] - E[J];
      if (J == NSEGL && NSEGE == NSEG && IBAR > 3) A = A + (2.0 * DSUB[NSEGE]);
      B = (DRCUL[J] * DT) * (R8_POW(1.0 / (DUL[J] - DRSUL[J]), C));
      TOLER = 0.003 * (DUL[J] - DRSUL[J]);
      X0 = (SWMAX + SWMIN) / 2.0;
      {
        int ITER;
        for (ITER = 1; ITER <= 100; ++ITER) {
          FX0 = A - (B * (R8_POW(X0, C))) - (2 * X0);
          if (FX0 <= (0.3f * TOLER)) break;
          FX0DER = -(B * C * (R8_POW(X0, C - 1))) - 2;
          if (fabs(FX0DER) < 1.0e-03) break;
          DX = FX0 / FX0DER;
          X0 = X0 - DX;
          if (fabs(DX) < TOLER) break;
        }
        if (ITER > 100) { if (IFLAG == 0) ++nonconv; IFLAG = 1; }
      }
      DSWUL[J] = X0;
      if (DSWUL[J] > SWMAX) DSWUL[J] = SWMAX;
      if (DSWUL[J] < SWMIN) DSWUL[J] = SWMIN;
      DRIN[J + 1] = 2.0 * (SWULY[J] - DSWUL[J]) + BALY[J] + DRIN[J] + DRCR[J] + DSUB[J] - E[J];
      if (NSEGE == NSEG && IBAR > 3) DRIN[J + 1] = DRIN[J + 1] + DSUB[NSEGE];
    L1070:
      if (LAYSEG[J][2] < LD[NPROF]) {
        if (DRIN[J + 1] > (DRCUL[J] * DT)) DRIN[J + 1] = DRCUL[J] * DT;
      }
      if (DRIN[J + 1] > DRNMAX) DRIN[J + 1] = DRNMAX;
      if (DRIN[J + 1] < 0.0) DRIN[J + 1] = 0.0;
      if (J < NSEGL && LAYSEG[J + 1][2] < LD[NPROF]) {
        if (DRIN[J + 1] > (DRCUL[J + 1] * DT)) {
          sw = (r4)((SWULY[J] + (BALY[J] / 2.0) - DRSUL[J]) / (DUL[J] - DRSUL[J]));
          TRATIO = (r4)(DTHICK[J] / DTHICK[J + 1]);
          DRMX = DUL[J + 1] - SWULY[J + 1] - (BALY[J + 1] / 2.0) + (DRCUL[J + 1] * DT) + E[J + 1] -
                 DSUB[J + 1] - DRCR[J + 1];
          if (DRMX > DRNMAX) DRMX = DRNMAX;
          DRNMX = DRCUL[J + 1] * DT * (1 + (sw * TRATIO));
          if (DRNMX > DRMX) DRNMX = DRMX;
          if (DRNMX > DRNMAX) DRNMX = DRNMAX;
          if (DRIN[J + 1] > DRNMX) DRIN[J + 1] = DRNMX;
          if (DRIN[J + 1] < 0.0) DRIN[J + 1] = 0.0;
        }
      }
    }
  L1100:
    for (int J = NSEGBv; J <= NSEGL; ++J) {
      if (J == NSEGE && IBAR == 3) DRIN[J + 1] = 0.0;
      BALT[J] = DRIN[J] + DSUB[J] + DRCR[J] - DRIN[J + 1] - E[J];
      if (J == NSEGL && IBAR > 3) BALT[J] = BALT[J] + DSUB[NSEGE];
      DSWUL[J] = SWULY[J] + ((BALY[J] + BALT[J]) / 2.0);
    }
  }

  void DRAIN(r4 FIN, r8& DRNv, r8& PRCv, r8& HEDv, r4& EXTRA, int NPROF, int NSEGBv, int NSEGE,
             int LAYBv, int LAYEv) {                          // F:4133-4277
    r8 E[72], DSUB[72], DRCR[72];
    r8 DTHICS = DTHICK[NSEGE], DRCULS = DRCUL[NSEGE];
    int IBAR = 0, LAT = 0, NSEGL = NSEGE, LAYD = 0;
    if (LSEG[NSEGE] == 3) IBAR = 1;
    if (LSEG[NSEGE] >= 4) IBAR = 2;
    if (LAYBv == LAYEv) IBAR = 0;
    if (IBAR >= 1) {
      NSEGL = NSEGE - 1;
      if (LD[NPROF] > 0) LAT = 1;
      if (LAT == 1) LAYD = LAYSEG[NSEGL][2];
    }
    r8 DT = 1.0 / NSTEP[NPROF];
    r8 F = FIN * DT;
    for (int J = NSEGBv; J <= NSEGE; ++J) {
      SWULY[J] = DSWUL[J];
      E[J] = ET[J] * DT;
      DSUB[J] = DSUBIN[J] * DT;
      DRCR[J] = DRCRS[J] * DT;
    }
    EXTRA = 0.0f;
    r8 EXCESS = 0.0, THY = 0.0, QPERC = 0.0, QLAT = 0.0, QPERCY = 0.0, QLATY = 0.0, ELKS = 0.0;
    HEDv = 0.0; PRCv = 0.0; DRNv = 0.0;
    if (NSEG == NSEGE) {
      if (IBAR == 1) { if (SUBIN[LP[NPROF]] > 0.0f) IBAR = IBAR + 3; }
      else if (IBAR == 2) { if (SUBIN[LP[NPROF]] > 0.0f || SUBIN[LP[NPROF] - 1] > 0.0f) IBAR = IBAR + 3; }
      else if (DSUBIN[NSEGE] > 0.0) IBAR = IBAR + 3;
    }
    for (int K = 1; K <= NSTEP[NPROF]; ++K) {
      DRIN[NSEGBv] = F + EXCESS;
      EXCESS = 0.0;
      if (IBAR == 0 || IBAR == 3) {
        QLAT = 0.0; QLATY = 0.0;
        AVROUT(E, DSUB, DRCR, DT, IBAR, LAT, NSEGBv, NSEGL, NSEGE, NPROF, QPERCY, QLATY, QPERC, QLAT);
        QPERC = DRIN[NSEGE + 1] / DT;
        PROFIL(EXCESS, NSEGBv, NSEGE);
        EXTRA = (r4)EXCESS;
        PRCv = PRCv + DRIN[NSEGE + 1];
        DRNv = 0.0;
      } else {
        HEAD(THY, NSEGBv, NSEGL);
        QPERCY = 0.0; QLATY = 0.0;
        if (THY > 0.0) {
          if (IBAR == 1) QPERCY = DRCULS * (THY + DTHICS) / DTHICS;
          if (IBAR == 2) FML(THY, QPERCY, NSEGE, NPROF);
          if (LAT == 1) { LATKS(THY, NSEGBv, NSEGL, ELKS); LATFLO(LAYD, THY, ELKS, QLATY); }
        }
        AVROUT(E, DSUB, DRCR, DT, IBAR, LAT, NSEGBv, NSEGL, NSEGE, NPROF, QPERCY, QLATY, QPERC, QLAT);
        PRCv = PRCv + DRIN[NSEGE + 1];
        DRNv = DRNv + DRIN[NSEGL + 1] - DRIN[NSEGE + 1] + DSUB[NSEGE];
        PROFIL(EXCESS, NSEGBv, NSEGL);
        EXTRA = (r4)EXCESS;
      }
      for (int J = NSEGBv; J <= NSEGL; ++J) { BALY[J] = BALT[J]; SWULY[J] = DSWUL[J]; }
      if (IBAR != 0 && IBAR != 3) HEDv = HEDv + THY * DT;
    }
    for (int J = NSEGBv; J <= NSEGL; ++J) DCHG[J] = BALT[J];
  }

  void RECIRC(int IMO) {                                      // F:5311-5440
    for (int K = 1; K <= LAY; ++K) RCRI[K] = 0.0;
    for (int n = 1; n <= 5; ++n) {
      if (LD[n] != 0) {
        if (RECIR[LD[n]] > 0.0f) {
          int t = LAYR[LD[n]];
          RCRI[t] = RCRI[t] + RCR[n];
          RCRIM[t][IMO] = (r4)(RCRIM[t][IMO] + RCR[n]);
          RCRIA[t] = (r4)(RCRIA[t] + RCR[n]);
        }
      }
    }
    for (int J = 1; J <= LAY; ++J) DRCRS[J] = 0.0;            // sic: LAY, not NSEG
    for (int K = 1; K <= LAY; ++K) {
      if (RCRI[K] > PRCRI[K]) PRCRI[K] = (r4)RCRI[K];
      if (LRIN[K] == 1) {
        for (int J = LSEGS[K][1]; J <= LSEGS[K][2]; ++J) {
          if (LSEGS[K][1] > 7) DRCRS[J] = RCRI[K] * DTHICK[J] / THICK[K];
          else if (K == 1) {
            r4 RCRF = 0.0f;
            for (int L = LSEGS[K][1]; L <= LSEGS[K][2]; ++L) {
              if (L < LSEGS[K][2]) { DRCRS[L] = RCRI[K] * DTHICK[L] / THICK[K]; RCRF = (r4)(RCRF + (DTHICK[L] / THICK[K])); }
              else DRCRS[L] = RCRI[K] * (1.0f - RCRF);
            }
          } else if (LSEGS[K][1] == LSEGS[K][2]) DRCRS[J] = RCRI[K];
          else {
            r4 TKM1 = 0.0f;
            for (int M = 1; M <= K - 1; ++M) TKM1 = TKM1 + THICK[M];
            r4 TJM1 = 0.0f;
            if (LSEGS[K][1] > 1) for (int M = 1; M <= LSEGS[K][1] - 1; ++M) TJM1 = (r4)(TJM1 + DTHICK[M]);
            r4 DSEG = TKM1 - TJM1;
            r4 RCRF = (r4)((DTHICK[LSEGS[K][1]] - DSEG) / THICK[K]);
            DRCRS[LSEGS[K][1]] = RCRI[K] * RCRF;
            for (int L = LSEGS[K][1] + 1; L <= LSEGS[K][2]; ++L) {
              if (L < LSEGS[K][2]) { DRCRS[L] = RCRI[K] * DTHICK[L] / THICK[K]; RCRF = (r4)(RCRF + (DTHICK[L] / THICK[K])); }
              else DRCRS[L] = RCRI[K] * (1.0f - RCRF);
            }
          }
        }
      }
    }
  }
};

// value a Fortran F7.3 field round-trips to through the Python parser
// (F:6032-6072 "12(1X,F7.3)", pyhelp/processing.py:105-122 -> float32)
static float f73(float v) {
  char buf[64];
  snprintf(buf, sizeof buf, "%.3f", (double)v);
  if (strlen(buf) > 7) return NAN;       // the Fortran prints ******* (parser would raise)
  return strtof(buf, nullptr);
}

}  // namespace

// =============================================================================
// C entry point (ctypes).  d10 / d11: n records of exactly `reclen` bytes each.
// Returns 0, or: 1 weather file missing, 2 bad header, 3 weather file too short,
// 4 malformed D10, 5 too many years.
// =============================================================================
extern "C" int help_oracle_run(
    const char* path_d4, const char* path_d7, const char* path_d13,
    const char* d11_buf, int d11_n, const char* d10_buf, int d10_n, int reclen,
    int lmyr, float tfsoil,
    int* years,            // [lmyr]
    float* monthly_in,     // [lmyr][14][12]  raw REAL*4 accumulators, inches
    float* annual_in,      // [lmyr][14]      raw REAL*4 annual accumulators, inches
    float* parsed_mm,      // [lmyr][7][12]   what read_monthly_help_output returns
    double* daily,         // [lmyr*366][16] or NULL
    float* layer_sw,       // [2][20] initial (after spin-up) and final vol/vol, or NULL
    int* info,             // [32] or NULL
    float* balance_in)     // [lmyr] yearly budget residual, or NULL
{
  if (lmyr > MXYR) return 5;
  Help* hp = new Help();
  Help& h = *hp;
  const std::vector<std::string>*f4 = load_file(path_d4), *f7 = load_file(path_d7),
                                *f13 = load_file(path_d13);
  if (!f4 || !f7 || !f13) { delete hp; return 1; }
  h.u4.lines = f4; h.u7.lines = f7; h.u13.lines = f13;
  for (int i = 0; i < d11_n; ++i) h.d11.emplace_back(d11_buf + (size_t)i * reclen, reclen);
  for (int i = 0; i < d10_n; ++i) h.d10.emplace_back(d10_buf + (size_t)i * reclen, reclen);
  if (d11_n < 3) { delete hp; return 4; }
  const r4 TFSOIL = tfsoil;
  const int LMYR = lmyr;
  // ---- run_simulation (F:17-1034) -------------------------------------------
  static const r4 WFv[8] = {0, 0.085f, 0.334f, 0.252f, 0.151f, 0.091f, 0.054f, 0.033f};
  for (int j = 1; j <= 7; ++j) h.WF[j] = WFv[j];
  h.READIN();
  if (h.err) { int e = h.err; delete hp; return e; }
  h.CNTRLD();
  h.SGMNT();
  {
    r4 CN2 = h.CN2;
    r4 C2 = CN2 * CN2, C3 = CN2 * C2, C4 = CN2 * C3;
    r4 CN1 = 0.3750701f * CN2 + 2.756779E-03f * C2 - 1.638951E-05f * C3 + 5.142644E-07f * C4;
    h.SMXN = (1000.f / CN1) - 10.0f;
    r4 CN22 = 95.0f;
    if (CN2 >= 80) CN22 = 98.0f;
    r4 C22 = CN22 * CN22, C33 = CN22 * C22, C44 = CN22 * C33;
    r4 CN11 = 0.3750701f * CN22 + 2.756779E-03f * C22 - 1.638951E-05f * C33 + 5.142644E-07f * C44;
    h.SMXF = (1000.f / CN11) - 10.f;
  }
  h.ETCOEF();
  r4 SWULE;
  h.OSWULE = 0.0f;
  for (int J = 1; J <= h.NSEG; ++J) {
    h.OSWULE = h.OSWULE + h.SWUL[J];
    h.DSWUL[J] = h.SWUL[J];
    h.SWULI[J] = h.SWUL[J];
  }
  {
    int e = 0;
    for (int n = 1; n <= 6; ++n) {
      if (n > 1 && h.NSEGn[n] <= 0) continue;
      h.NSEGB[n] = e + 1;
      h.NSEGE_[n] = e + h.NSEGn[n];
      h.NSEGD[n] = h.NSEGE_[n];
      if (n < 6 && h.LSEG[h.NSEGE_[n]] > 2) h.NSEGD[n] = h.NSEGE_[n] - 1;
      h.LAYB[n] = (n == 1) ? 1 : h.LP[n - 1] + 1;
      h.LAYE[n] = h.LP[n];
      e = h.NSEGE_[n];
    }
    if (h.NSEGn[6] > 0) { h.NSEGE_[6] = h.NSEG; h.NSEGD[6] = h.NSEG; }
  }
  for (int J = 1; J <= 67; ++J) { h.BALT[J] = 0.0; h.BALY[J] = 0.0; h.ET[J] = 0.0f; }
  h.RUN = 0.0f; h.ES1T = 0.0f; h.TS2 = 0.0f;
  for (int n = 1; n <= 6; ++n) { h.DRN[n] = 0.0; h.PRC[n] = 0.0; h.HED[n] = 0.0; }
  for (int n = 1; n <= 5; ++n) h.RCR[n] = 0.0;
  h.ADDRUN = 0.0f;
  for (int n = 2; n <= 6; ++n) h.EXWAT[n] = 0.0f;
  h.INITIA();
  h.IDA = 0;
  h.ND = 365;
  r4 SWULL[22] = {0};
  int out_year = 0;
  long day_row = 0;
  for (;;) {   // label 1060
    h.SNOCOEF();
    h.SETUPS(SWULL);
    if (h.IPRE >= 2) {
      // OUTDAT (F:5451-5953) only echoes the design, except that it REWINDs the
      // weather files and skips their 4 header records (F:5938-5948).
      h.u4.rewind(); h.u7.rewind(); h.u13.rewind();
      h.skip_header(h.u4); h.skip_header(h.u7); h.skip_header(h.u13);
      if (layer_sw) for (int K = 1; K <= 20; ++K)
        layer_sw[K - 1] = (K <= h.LAY && h.THICK[K] != 0.f) ? SWULL[K] / h.THICK[K] : 0.f;
    }
    h.OSWULE = 0.0f;
    for (int J = 1; J <= h.NSEG; ++J) h.OSWULE = (r4)(h.OSWULE + h.DSWUL[J]);
    SWULE = 0.0f;
    for (int J = 1; J <= 7; ++J) SWULE = SWULE + h.SWUL[J];
    int IMO = 0;
    h.IDA = 1;
    for (int K = 1; K <= 20; ++K) h.RCRIA[K] = 0.0f;
    for (int I = 1; I <= 12; ++I) for (int K = 1; K <= 20; ++K) h.RCRIM[K][I] = 0.0f;
    bool restart = false;
    for (h.IYR = 1; h.IYR <= LMYR; ++h.IYR) {                 // DO 1090
      for (int n = 1; n <= 6; ++n) h.PRCA[n] = 0.0f;
      for (int n = 1; n <= 5; ++n) { h.DRNA[n] = 0.0f; h.RCRA_[n] = 0.0f; }
      h.RUNA = 0.0f; h.PREA = 0.0f; h.ETA = 0.0f; h.BAL = 0.0f; h.STOR = 0.0f;
      for (int I = 1; I <= 12; ++I) {
        for (int n = 1; n <= 6; ++n) h.PRCM[n][I] = 0.0f;
        for (int n = 1; n <= 5; ++n) { h.DRNM[n][I] = 0.0f; h.RCRM[n][I] = 0.0f; h.HEDM[n][I] = 0.0f; }
        for (int K = 1; K <= 20; ++K) h.RCRIM[K][I] = 0.0f;
        h.RUNM[I] = 0.0f; h.PREM[I] = 0.0f; h.ETM[I] = 0.0f;
      }
      for (int K = 1; K <= 20; ++K) h.RCRIA[K] = 0.0f;
      int MO1 = 1;
      IMO = 1;
      int NYEAR = 0, NT = 0;
      h.READCD(NYEAR, h.ND, NT);
      if (h.err) { int e = h.err; delete hp; return e; }
      h.JYEAR[h.IYR] = NYEAR;
      const int NDloop = h.ND;       // DO-loop bound is fixed at loop entry
      for (h.IDA = 1; h.IDA <= NDloop; ++h.IDA) {             // DO 1100
        const int IDA = h.IDA;
        h.MO = Help::MONTH(IDA, NT);
        if (h.MO != MO1) IMO = IMO + 1;
        MO1 = h.MO;
        h.PREM[IMO] = h.PREM[IMO] + h.PRE[IDA];
        h.PREA = h.PREA + h.PRE[IDA];
        h.FRZCHK(TFSOIL);
        int IT = 2;
        h.SNOW(IT);
        h.SMX = h.SMXN;
        if (h.IFREZ > 0) h.SMX = h.SMXF;
        h.RUNOFF();
        h.SOLT();
        h.EVAPOT();
        for (int J = 1; J <= 19; ++J) if (fabs(h.BALT[J]) < 1.E-08f) h.BALT[J] = 0.0f;
        h.CRPMOD();
        h.ADDRUN = 0.0f;
        h.DRAIN(h.PINF, h.DRN[1], h.PRC[1], h.HED[1], h.ADDRUN, 1, h.NSEGB[1], h.NSEGE_[1], h.LAYB[1], h.LAYE[1]);
        SWULE = 0.0f;
        r4 EXET = 0.0f;
        for (int J = 1; J <= 7; ++J) {
          if ((h.DSWUL[J] + (h.BALT[J] / 2.0)) < h.DWPUL[J]) {
            r4 EXET2 = (r4)(EXET + h.DWPUL[J] - h.DSWUL[J] - (h.BALT[J] / 2.0));
            if (EXET <= h.ETT) {
              h.DSWUL[J] = h.DWPUL[J] - (h.BALT[J] / 2.0);
              h.SWULY[J] = h.DSWUL[J];
              EXET = EXET2;
            }
          }
          SWULE = (r4)(SWULE + h.DSWUL[J] + h.BALT[J] / 2.0);
        }
        r4 SWE = SWULE / h.EDEPTH;
        h.RUN = h.RUN + h.ADDRUN * h.FRUNOF;
        h.ADDRUN = h.ADDRUN * (1.0f - h.FRUNOF);
        if (h.SNO > 0.0f) {
          h.SNO = h.SNO + h.ADDRUN;
          h.TWE = h.TWE + h.ADDRUN * 25.4f;
          h.XLIQW = h.XLIQW + h.ADDRUN * 25.4f;
          h.ADDRUN = 0.0f;
        }
        for (int J = 1; J <= h.NSEGD[1]; ++J) h.SWUL[J] = (r4)h.DSWUL[J];
        h.RUNM[IMO] = h.RUNM[IMO] + h.RUN;
        h.RUNA = h.RUNA + h.RUN;
        h.ETT = h.ETT - EXET;
        h.ETM[IMO] = h.ETM[IMO] + h.ETT;
        h.ETA = h.ETA + h.ETT;
        h.HEDM[1][IMO] = (r4)(h.HEDM[1][IMO] + h.HED[1]);
        if (h.LD[1] > 0) {
          h.RCR[1] = h.RECIR[h.LD[1]] * h.DRN[1] / 100.0f;
          h.DRN[1] = h.DRN[1] - h.RCR[1];
          h.RCRM[1][IMO] = (r4)(h.RCRM[1][IMO] + h.RCR[1]);
          h.RCRA_[1] = (r4)(h.RCRA_[1] + h.RCR[1]);
        }
        h.DRNM[1][IMO] = (r4)(h.DRNM[1][IMO] + h.DRN[1]);
        h.DRNA[1] = (r4)(h.DRNA[1] + h.DRN[1]);
        h.PRCM[1][IMO] = (r4)(h.PRCM[1][IMO] + h.PRC[1]);
        h.PRCA[1] = (r4)(h.PRCA[1] + h.PRC[1]);
        for (int n = 2; n <= 6; ++n) {                        // F:694-884
          r4 FINn = (r4)(h.PRC[n - 1] + h.EXWAT[n]);
          if (h.NSEGn[n] <= 0) break;
          h.EXWAT[n] = 0.0f;
          r8 dummy1 = 0, dummy2 = 0;
          if (n < 6)
            h.DRAIN(FINn, h.DRN[n], h.PRC[n], h.HED[n], h.EXWAT[n], n, h.NSEGB[n], h.NSEGE_[n], h.LAYB[n], h.LAYE[n]);
          else
            h.DRAIN(FINn, dummy1, h.PRC[n], dummy2, h.EXWAT[n], n, h.NSEGB[n], h.NSEGE_[n], h.LAYB[n], h.LAYE[n]);
          for (int J = h.NSEGB[n]; J <= h.NSEGD[n]; ++J) h.SWUL[J] = (r4)h.DSWUL[J];
          if (n < 6) {
            h.HEDM[n][IMO] = (r4)(h.HEDM[n][IMO] + h.HED[n]);
            if (h.LD[n] > 0) {
              h.RCR[n] = h.RECIR[h.LD[n]] * h.DRN[n] / 100.0f;
              h.DRN[n] = h.DRN[n] - h.RCR[n];
              h.RCRM[n][IMO] = (r4)(h.RCRM[n][IMO] + h.RCR[n]);
              h.RCRA_[n] = (r4)(h.RCRA_[n] + h.RCR[n]);
            }
            h.DRNM[n][IMO] = (r4)(h.DRNM[n][IMO] + h.DRN[n]);
            h.DRNA[n] = (r4)(h.DRNA[n] + h.DRN[n]);
          }
          h.PRCM[n][IMO] = (r4)(h.PRCM[n][IMO] + h.PRC[n]);
          h.PRCA[n] = (r4)(h.PRCA[n] + h.PRC[n]);
        }
        h.RECIRC(IMO);
#ifdef HELP_ORACLE_TRACE   // debugging aid: per-day state dump (build with -DHELP_ORACLE_TRACE)
        {
          fprintf(stderr, "D ipre %d day %d pinf %.9g ett %.9g es %.9g ep %.9g run %.9g addrun %.9g ifrez %d ET:", h.IPRE, h.IDA, h.PINF, h.ETT, h.ES, h.EP, h.RUN, h.ADDRUN, h.IFREZ);
          for (int J = 1; J <= 7; ++J) fprintf(stderr, " %.9g", h.ET[J]);
          fprintf(stderr, " SW:"); for (int J = 1; J <= h.NSEG; ++J) fprintf(stderr, " %.17g", h.DSWUL[J]);
          fprintf(stderr, " BAL:"); for (int J = 1; J <= h.NSEG; ++J) fprintf(stderr, " %.17g", h.BALT[J]);
          fprintf(stderr, "\n");
        }
#endif
        if (daily && h.IPRE >= 2) {                           // OUTDAY columns (F:5984-5996), inches
          double* d = daily + day_row * NDAILY;
          d[0] = h.PRE[IDA]; d[1] = h.RUN; d[2] = h.ETT; d[3] = SWE;
          for (int n = 1; n <= 3; ++n) {
            d[4 + 3 * (n - 1)] = h.HED[n];
            d[5 + 3 * (n - 1)] = h.DRN[n] + h.RCR[n];
            d[6 + 3 * (n - 1)] = h.PRC[n];
          }
          d[13] = (IT == 1) ? 1.0 : 0.0;   // air-temperature freeze star
          d[14] = h.IFREZ;                 // soil-freeze star
          d[15] = h.SNO;
          ++day_row;
        }
      }
      // yearly storage and budget check (F:938-968)
      h.PSWULE = h.EXWAT[2] + h.EXWAT[3] + h.EXWAT[4] + h.EXWAT[5] + h.EXWAT[6];
      h.SSNO = h.ADDRUN + h.SNO;
      for (int J = 1; J <= h.NSEG; ++J) h.PSWULE = (r4)(h.PSWULE + h.DSWUL[J] + (h.BALT[J] / 2.0));
      r4 SUBINY = h.SUBINF * (r4)h.ND;
      h.STOR = h.PSWULE - h.OSWULE + h.SSNO - h.OLDSNO;
      h.BAL = h.PREA - h.RUNA - h.ETA - h.STOR - h.DRNA[1] - h.DRNA[2] - h.DRNA[3] - h.DRNA[4] - h.DRNA[5] + SUBINY;
      {
        int last = 1;
        for (int n = 6; n >= 2; --n) if (h.NSEGn[n] > 0) { last = n; break; }
        h.BAL = h.BAL - h.PRCA[last];
      }
      if (h.IPRE >= 2) {
        // OUTMO/OUTMO2 value semantics (F:6415-6600)
        const int y = out_year++;
        years[y] = h.JYEAR[h.IYR];
        float* m = monthly_in + (size_t)y * NROW * 12;
        float* a = annual_in + (size_t)y * NROW;
        for (int j = 1; j <= 12; ++j) {
          m[0 * 12 + j - 1] = h.PREM[j]; m[1 * 12 + j - 1] = h.RUNM[j]; m[2 * 12 + j - 1] = h.ETM[j];
          for (int n = 1; n <= 5; ++n) m[(2 + n) * 12 + j - 1] = h.DRNM[n][j];
          for (int n = 1; n <= 6; ++n) m[(7 + n) * 12 + j - 1] = h.PRCM[n][j];
        }
        a[0] = h.PREA; a[1] = h.RUNA; a[2] = h.ETA;
        for (int n = 1; n <= 5; ++n) a[2 + n] = h.DRNA[n];
        for (int n = 1; n <= 6; ++n) a[7 + n] = h.PRCA[n];
        if (balance_in) balance_in[y] = h.BAL;
        float* p = parsed_mm + (size_t)y * 7 * 12;
        for (int k = 0; k < 7 * 12; ++k) p[k] = 0.0f;
        bool have_sub1 = false, have_perc = false;
        for (int j = 1; j <= 12; ++j) {
          p[0 * 12 + j - 1] = f73(h.PREM[j] * 25.4f);
          p[1 * 12 + j - 1] = f73(h.RUNM[j] * 25.4f);
          p[2 * 12 + j - 1] = f73(h.ETM[j] * 25.4f);
        }
        for (int n = 1; n <= 6; ++n) {
          if (n > 1 && !(h.LP[n] > 0)) continue;
          if (n <= 5 && h.LD[n] > 0) {                         // ' LAT. DRAINAGE IN LAYER'
            for (int j = 1; j <= 12; ++j) {
              float v = f73(h.DRNM[n][j] * 25.4f);
              if (!have_sub1) p[3 * 12 + j - 1] = v;
              else p[4 * 12 + j - 1] = p[4 * 12 + j - 1] + v;
            }
            have_sub1 = true;
          }
          for (int j = 1; j <= 12; ++j) {                      // ' PERCOLATION THROUGH LAYER'
            float v = (n == 6 && j >= 7) ? f73(h.PRCM[n][j])   // sic: F:6593-6594 omits *25.4
                                         : f73(h.PRCM[n][j] * 25.4f);
            if (!have_perc) p[5 * 12 + j - 1] = v;
            p[6 * 12 + j - 1] = v;
          }
          have_perc = true;
        }
      }
      h.OSWULE = h.PSWULE;
      h.OLDSNO = h.SSNO;
      if (h.IPRE < 2) {                                       // spin-up year just ended (F:986-996)
        h.IPRE = h.IPRE + 2;
        for (int J = 1; J <= h.NSEG; ++J) {
          h.DSWUL[J] = h.DSWUL[J] + h.DCHG[J] / 2.0;
          h.SWUL[J] = (r4)h.DSWUL[J];
          h.DCHG[J] = 0.0; h.BALT[J] = 0.0; h.BALY[J] = 0.0;
        }
        restart = true;
        break;
      }
    }
    if (restart) continue;
    break;
  }
  // F:1008-1012 and OUTSW's per-layer final storage
  for (int n = 2; n <= 6; ++n) if (h.NSEGn[n] > 0) h.DSWUL[h.NSEGB[n]] = h.DSWUL[h.NSEGB[n]] + h.EXWAT[n];
  if (layer_sw) {
    h.layer_water(SWULL);
    for (int K = 1; K <= 20; ++K)
      layer_sw[20 + K - 1] = (K <= h.LAY && h.THICK[K] != 0.f) ? SWULL[K] / h.THICK[K] : 0.f;
  }
  if (info) {
    for (int i = 0; i < 32; ++i) info[i] = 0;
    info[0] = h.LAY; info[1] = h.NSEG; info[2] = h.IDFS; info[3] = h.nonconv;
    for (int n = 1; n <= 6; ++n) { info[3 + n] = h.LP[n]; info[9 + n] = h.NSTEP[n]; info[15 + n] = h.NSEGn[n]; }
    for (int n = 1; n <= 5; ++n) { info[21 + n] = h.LD[n]; info[26 + n] = h.LCASE[n]; }
  }
  delete hp;
  return 0;
}

extern "C" void help_oracle_clear_cache() { g_files.clear(); }

// -----------------------------------------------------------------------------
// Evaluate the transcendental functions this build uses (tests/test_help_math.py
// checks the deterministic ones against mpmath and against the GPU's, bit for bit).
// fn: 0 pow 1 exp 2 log 3 sin 4 cos 5 atan 6 acos(r4) 7 tan(r4) 8 r4_pow 9 r4_exp 10 r4_log
//     11 r4_sin 12 r4_cos 13 r4_log10 14 r4_sqrt 15 r4_pow8 16 r4_pow7 17 r4_pow4
// -----------------------------------------------------------------------------
extern "C" int help_oracle_math_eval(int fn, const double* x, const double* y, double* out, long long n) {
  for (long long i = 0; i < n; ++i) {
    const double a = x[i], b = y ? y[i] : 0.0;
    const float af = (float)a, bf = (float)b;
    double r;
    switch (fn) {
      case 0: r = R8_POW(a, b); break;
      case 1:
#ifdef HELP_ORACLE_GLIBC
        r = exp(a);
#else
        r = helpmath::det_exp(a);
#endif
        break;
      case 2: r = R8_LOG(a); break;
      case 3: r = R8_SIN(a); break;
      case 4: r = R8_COS(a); break;
      case 5: r = R8_ATAN(a); break;
      case 6: r = R4_ACOS(af); break;
      case 7: r = R4_TAN(af); break;
      case 8: r = R4_POW(af, bf); break;
      case 9: r = R4_EXP(af); break;
      case 10: r = R4_LOG(af); break;
      case 11: r = R4_SIN(af); break;
      case 12: r = R4_COS(af); break;
      case 13: r = R4_LOG10(af); break;
      case 14: r = R4_SQRT(af); break;
      case 15: r = R4_POW8(af); break;
      case 16: r = R4_POW7(af); break;
      case 17: r = R4_POW4(af); break;
#ifndef HELP_ORACLE_GLIBC
      case 18: { r = helpmath::det_pow2(a, b, b == 0.0 ? 1.5 : 1.0 / b, a).a; break; }   // dual-issue form
      case 19: { r = helpmath::det_pow_samebase(a, b - 1.0, b).b; break; }                // shared-logarithm form
      case 20: { const double q = a / b; r = helpmath::div_by_const(a, b, 1.0 / b); if (q != q) r = q; break; }     // a/b through RN(1/b)
#endif
      default: return -1;
    }
    out[i] = r;
  }
  return 0;
}
