// DEVELOPER DEBUGGING AID (see cuda_shim.h).  Runs help_setup_kernel + help_sim_kernel
// on the CPU, thread by thread.  Build: g++ -O1 -g -ffp-contract=off -shared -fPIC emu.cpp -o libemu.so
#include "cuda_shim.h"
#ifndef EMU_MAXSEG
#define EMU_MAXSEG 67   // template tier of the emulated instantiation (16: the natural-soil tier with its tight arrays)
#endif
#ifndef EMU_REGS
#define EMU_REGS 128     // register budget of the emulated instantiation (>= 96: carried look-ahead; below: no-carry loop)
#endif
#include <cstdio>
#ifdef HELP_DBG_CAPMEMO
extern "C" void memo2_newcell();
#endif
static int g_dbg_cell = -1; static int g_ipre=0, g_ida=0;
#ifndef HELP_DBG_PRE
#define HELP_DBG_PRE g_ipre=IPRE; g_ida=IDA;
#endif
#define HELP_DBG_SEG if ((int)s==g_dbg_cell && g_ipre==0 && g_ida==150) printf("S J %d swlim %.9g drnmax %.17g out %.17g base %.17g\n", J, SWLIM, DRNMAX, out, base);
#define HELP_DBG_DAY if ((int)s == g_dbg_cell) { printf("D ipre %d day %d pinf %.9g ett %.9g es %.9g ep %.9g run %.9g addrun %.9g ifrez %d ET:", IPRE, IDA, U.PINF, U.ETT, U.ES, U.EP, U.RUN, U.ADDRUN, U.IFREZ); for (int J = 1; J <= 7; ++J) printf(" %.9g", U.ET[J]); printf(" SW:"); for (int J = 1; J <= NSEG; ++J) printf(" %.17g", SWa(J)); printf(" BAL:"); for (int J = 1; J <= NSEG; ++J) printf(" %.17g", BALa(J));printf("\n"); }
#include <cstdio>
#include <vector>
#include "../../include/help_b200.h"
#include "../../pyhelp_b200/csrc/help_image.h"
#include "../../pyhelp_b200/csrc/help_setup.cuh"
#include "../../pyhelp_b200/csrc/help_sim.cuh"
using namespace helpb200;

extern "C" void emu_debug_cell(int c) { g_dbg_cell = c; }
static double* g_daily = nullptr; static int64_t g_daily_days = 0;
extern "C" void emu_set_daily(double* d, long long days) { g_daily = d; g_daily_days = days; }   // all cells, [cell][day][16]
extern "C" int emu_run(const help_b200_batch* b, float* monthly, float* yearly, float* raw, uint32_t* flags, int32_t* topo) {
  const int64_t n = b->n_cells; const int ny = b->n_years;
  const size_t wrow = (size_t)ny * kDays;
  std::vector<float> pre(b->pre, b->pre + b->n_nodes_p * wrow), tmp(b->tmp, b->tmp + b->n_nodes_t * wrow), rad(b->rad, b->rad + b->n_nodes_r * wrow);
  blockDim.x = 1; gridDim.x = 1; blockIdx.x = 0; threadIdx.x = 0;
  convert_weather_kernel(pre.data(), b->iu4, wrow, pre.size(), 0);
  convert_weather_kernel(tmp.data(), b->iu7, wrow, tmp.size(), 1);
  convert_weather_kernel(rad.data(), b->iu13, wrow, rad.size(), 2);
  BatchDev B{};
  B.n_cells = n; B.n_years = ny; B.max_layers = b->max_layers; B.n_nodes_p = b->n_nodes_p; B.n_nodes_t = b->n_nodes_t; B.n_nodes_r = b->n_nodes_r;
  B.years = b->years; B.pre = pre.data(); B.tmp = tmp.data(); B.rad = rad.data(); B.iu4 = b->iu4; B.iu7 = b->iu7; B.iu13 = b->iu13;
  B.tm_normals = b->tm_normals; B.idfs = b->idfs; B.cell_f32 = b->cell_f32; B.cell_i32 = b->cell_i32; B.layer_f32 = b->layer_f32; B.layer_i32 = b->layer_i32; B.layer_rc = b->layer_rc;
  int segcap = 7 + 3 * b->max_layers; if (segcap > 67) segcap = 67;
  int64_t stride = ((n + 31) / 32) * 32;
  std::vector<float> f32(Image::n_f32(segcap) * stride); std::vector<int32_t> i32(Image::n_i32(segcap) * stride); std::vector<double> f64(Image::n_f64(segcap) * stride);
  std::vector<char> rec((size_t)Image::rec_bytes(stride, segcap));
  Image img{f32.data(), i32.data(), f64.data(), stride, segcap, rec.data()};
  for (int64_t c = 0; c < n; ++c) { blockIdx.x = (unsigned)c; help_setup_kernel(B, img, nullptr, flags, topo, nullptr, nullptr); }
  SimArgs A; A.img = img; A.pre = pre.data(); A.tmp = tmp.data(); A.rad = rad.data(); A.years = b->years; A.n_years = ny; A.n_cells = n; A.first = 0; A.count = n; A.perm = nullptr;
  A.monthly = monthly; A.yearly = yearly; A.raw = raw; A.flags = flags; std::vector<int32_t> drow(n); for (int64_t c = 0; c < n; ++c) drow[c] = (int32_t)c;
  A.daily = g_daily; A.daily_row = g_daily ? drow.data() : nullptr; A.daily_days = g_daily_days;
#ifdef HELP_DBG_CAPMEMO
  for (int64_t c = 0; c < n; ++c) { blockIdx.x = (unsigned)c; memo2_newcell(); help_sim_kernel<EMU_MAXSEG, 128, EMU_REGS>(A); }
#else
  for (int64_t c = 0; c < n; ++c) { blockIdx.x = (unsigned)c; help_sim_kernel<EMU_MAXSEG, 128, EMU_REGS>(A); }
#endif
  return 0;
}
#ifdef HELP_EMU_COUNT
extern "C" void emu_cap_counts(long long* out) { for (int i = 0; i < 8; ++i) out[i] = g_cap[i]; }
extern "C" void emu_pow_counts(long long* out) { for (int i = 0; i < 4096; ++i) { out[i] = g_pow_count[i]; g_pow_count[i] = 0; } }
#endif
#ifdef HELP_EMU_MEMO
extern "C" long long emu_memo_len(int cell, int which) { return which ? g_memo[cell].swp.size() : g_memo[cell].pre.size(); }
extern "C" void emu_memo_get(int cell, int which, unsigned char* out) { auto& v = which ? g_memo[cell].swp : g_memo[cell].pre; for (size_t i = 0; i < v.size(); ++i) out[i] = v[i]; }
#endif
