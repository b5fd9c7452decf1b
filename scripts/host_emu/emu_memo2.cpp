#include <vector>
#include <cstdint>
static int g_day = 0;            // running day counter of the current cell (both passes)
static std::vector<uint8_t> g_ev;   // [cell][day][K-1][J]: 0 not evaluated, 1 miss, 2 hit
static int g_nd = 0, g_ncell = 0;
static float g_prev[70];
#define HELP_DBG_PRE { g_day++; }
#define HELP_DBG_CAPMEMO { size_t ix = (((size_t)s * g_nd + (g_day - 1)) * 4 + (K - 1)) * 20 + J; if (g_day - 1 < g_nd && K <= 4 && J < 20) g_ev[ix] = (RELSW == g_prev[J]) ? 2 : 1; g_prev[J] = RELSW; }
#include "emu.cpp"
extern "C" void memo2_init(int ncell, int nd) { g_ncell = ncell; g_nd = nd; g_ev.assign((size_t)ncell * nd * 4 * 20, 0); }
extern "C" void memo2_newcell() { g_day = 0; for (auto& v : g_prev) v = -1.f; }
extern "C" const uint8_t* memo2_data() { return g_ev.data(); }
