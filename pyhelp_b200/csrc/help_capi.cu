// help_capi.cu -- the C ABI of include/help_b200.h over the sm_100a kernels.
//
// Launch plan of one simulate():
//   1. help_setup_kernel (identity order)      -> class/cost sort key per cell
//   2. cub::DeviceRadixSort (key, cell index)  -> perm: image slot -> cell
//   3. help_setup_kernel (perm)                -> cell image in class order
//   4. help_sim_kernel<MAXSEG>                 -> monthly / yearly / raw / flags
// Weather is converted to the engine's internal units once at upload
// (convert_weather_kernel).  No CPU fallback exists: every entry point fails
// with HELP_B200_ENODEV when no CUDA device is usable.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cub/device/device_radix_sort.cuh>
#include <new>
#include <string>
#include <vector>

#include "../../include/help_b200.h"
#include "help_image.h"
#include "help_setup.cuh"
#include "help_sim.cuh"
#include "help_nodes.cuh"
#include "help_post.cuh"
#include "help_weather.cuh"
#include "help_parse.cuh"

using namespace helpb200;

namespace {

thread_local std::string g_err;

int fail(int code, const char* what, cudaError_t ce = cudaSuccess) {
  char buf[512];
  if (ce != cudaSuccess) snprintf(buf, sizeof buf, "%s: %s", what, cudaGetErrorString(ce));
  else snprintf(buf, sizeof buf, "%s", what);
  g_err = buf;
  return code;
}

#define CU(call)                                                            \
  do {                                                                      \
    cudaError_t _e = (call);                                                \
    if (_e != cudaSuccess) return fail(_e == cudaErrorMemoryAllocation ? HELP_B200_ENOMEM : HELP_B200_ECUDA, #call, _e); \
  } while (0)

// Device memory comes from the device's default stream-ordered pool, which is told to keep what
// is freed (release threshold = max): a caller that runs batch after batch through
// help_b200_run() -- create, simulate, fetch, destroy -- re-uses the same ~1.5 GB instead of
// paying cudaMalloc/cudaFree (tens of ms, sporadically most of a second) every call.
// help_b200_release_memory() hands the cached memory back to the driver.
static void pool_setup() {
  static int done_for = -1;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev == done_for) return;
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
    uint64_t keep = UINT64_MAX;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
  }
  done_for = dev;
}

template <typename T>
int dev_alloc(T** p, size_t count) {
  *p = nullptr;
  if (count == 0) count = 1;
  pool_setup();
  CU(cudaMallocAsync((void**)p, count * sizeof(T), (cudaStream_t)0));
  CU(cudaStreamSynchronize((cudaStream_t)0));        // usable from any stream from here on
  return 0;
}
static void dev_free(void* p) { if (p) cudaFreeAsync(p, (cudaStream_t)0); }

__global__ void iota_kernel(int32_t* p, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = (int32_t)i;
}

}  // namespace

struct help_b200_engine {
  int device = 0;
  BatchDev B{};
  int64_t n = 0, stride = 0;
  int ny = 0, segcap = 0, maxseg_t = 0;
  bool want_monthly = false, want_raw = false, sort = true;
  // device buffers owned by the engine
  int32_t* d_years = nullptr;
  float *d_pre = nullptr, *d_tmp = nullptr, *d_rad = nullptr, *d_tm = nullptr;
  int32_t *d_iu4 = nullptr, *d_iu7 = nullptr, *d_iu13 = nullptr, *d_idfs = nullptr;
  float *d_cell_f32 = nullptr, *d_layer_f32 = nullptr;
  int32_t *d_cell_i32 = nullptr, *d_layer_i32 = nullptr;
  double* d_layer_rc = nullptr;
  float* d_img_f32 = nullptr; int32_t* d_img_i32 = nullptr; double* d_img_f64 = nullptr;
  char* d_img_rec = nullptr;
  float *d_monthly = nullptr, *d_yearly = nullptr, *d_raw = nullptr;
  uint32_t* d_flags = nullptr; int32_t* d_topo = nullptr;
  uint64_t *d_key = nullptr, *d_key2 = nullptr; int32_t *d_perm = nullptr, *d_iota = nullptr;
  void* d_cub = nullptr; size_t cub_bytes = 0;
  double* d_daily = nullptr; int32_t* d_daily_row = nullptr; int32_t n_daily = 0; int64_t daily_days = 0;
  int32_t* d_nnarrow = nullptr;
  cudaStream_t stream = nullptr, stream2 = nullptr;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  int64_t n_narrow = 0;
  float ms_h2d = 0, ms_setup = 0, ms_sim = 0;
  int n_launches = 0;
  double alg_bytes = 0;
};

static void engine_free(help_b200_engine* e) {
  if (!e) return;
  cudaSetDevice(e->device);
  void* ptrs[] = {e->d_years, e->d_pre, e->d_tmp, e->d_rad, e->d_tm, e->d_iu4, e->d_iu7, e->d_iu13, e->d_idfs,
                  e->d_cell_f32, e->d_layer_f32, e->d_cell_i32, e->d_layer_i32, e->d_layer_rc, e->d_img_f32,
                  e->d_img_i32, e->d_img_f64, e->d_img_rec, e->d_monthly, e->d_yearly, e->d_raw, e->d_flags, e->d_topo,
                  e->d_key, e->d_key2, e->d_perm, e->d_iota, e->d_cub, e->d_daily, e->d_daily_row, e->d_nnarrow};
  cudaDeviceSynchronize();                    // nothing of this engine is in flight when its memory returns to the pool
  for (void* p : ptrs) dev_free(p);
  for (auto& ev : e->ev) if (ev) cudaEventDestroy(ev);
  if (e->ev_fork) cudaEventDestroy(e->ev_fork);
  if (e->ev_join) cudaEventDestroy(e->ev_join);
  if (e->stream) cudaStreamDestroy(e->stream);
  if (e->stream2) cudaStreamDestroy(e->stream2);
  delete e;
}

static int validate(const help_b200_batch* b) {
  if (!b) return fail(HELP_B200_EINVAL, "batch is NULL");
  if (b->n_cells < 0 || b->n_years < 1 || b->n_years > HELP_B200_MAX_YEARS)
    return fail(HELP_B200_EINVAL, "n_cells must be >= 0 and 1 <= n_years <= 100");
  if (b->max_layers < 1 || b->max_layers > HELP_B200_MAX_LAYERS)
    return fail(HELP_B200_EINVAL, "max_layers must be in 1..20");
  if (b->n_nodes_p < 1 || b->n_nodes_t < 1 || b->n_nodes_r < 1)
    return fail(HELP_B200_EINVAL, "at least one weather node per variable is required");
  if (!b->years || !b->pre || !b->tmp || !b->rad || !b->iu4 || !b->iu7 || !b->iu13 || !b->tm_normals || !b->idfs)
    return fail(HELP_B200_EINVAL, "a weather pointer is NULL");
  if (b->n_cells > 0 && (!b->cell_f32 || !b->cell_i32 || !b->layer_f32 || !b->layer_i32 || !b->layer_rc))
    return fail(HELP_B200_EINVAL, "a cell pointer is NULL");
  return 0;
}

extern "C" double help_b200_algorithmic_bytes(const help_b200_batch* b) {
  if (!b || b->n_cells <= 0) return 0.0;
  const int64_t n = b->n_cells;
  std::vector<uint8_t> sp(b->n_nodes_p, 0), st(b->n_nodes_t, 0), sr(b->n_nodes_r, 0);
  double layers = 0;
  for (int64_t c = 0; c < n; ++c) {
    int p = b->cell_i32[HB_NODE_P * n + c], t = b->cell_i32[HB_NODE_T * n + c], r = b->cell_i32[HB_NODE_R * n + c];
    if (p >= 0 && p < b->n_nodes_p) sp[p] = 1;
    if (t >= 0 && t < b->n_nodes_t) st[t] = 1;
    if (r >= 0 && r < b->n_nodes_r) sr[r] = 1;
    layers += b->cell_i32[HB_NLAY * n + c];
  }
  double wp = 0, wt = 0, wr = 0;
  for (auto v : sp) wp += v;
  for (auto v : st) wt += v;
  for (auto v : sr) wr += v;
  double days = 0;
  for (int y = 0; y < b->n_years; ++y) days += (b->years[y] % 4 == 0) ? 366 : 365;
  // 12*W*D (3 variables x 4 B) + (56 + 36 L) per cell + 336 per cell-year (7 x 12 x 4 B)
  return 4.0 * days * (wp + wt + wr) + 56.0 * n + 36.0 * layers + 336.0 * (double)n * b->n_years;
}

extern "C" int help_b200_abi_version(void) { return HELP_B200_ABI_VERSION; }
extern "C" const char* help_b200_last_error(void) { return g_err.c_str(); }
extern "C" int help_b200_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

extern "C" int help_b200_release_memory(void) {
  if (help_b200_device_count() < 1) return fail(HELP_B200_ENODEV, "no CUDA device");
  int dev = 0; cudaMemPool_t pool;
  CU(cudaGetDevice(&dev));
  CU(cudaDeviceSynchronize());
  CU(cudaDeviceGetDefaultMemPool(&pool, dev));
  CU(cudaMemPoolTrimTo(pool, 0));
  return 0;
}

extern "C" int help_b200_set_device(int device) {
  if (device < 0 || device >= help_b200_device_count()) return fail(HELP_B200_ENODEV, "set_device: no such CUDA device");
  CU(cudaSetDevice(device));
  return 0;
}

extern "C" int help_b200_engine_create(const help_b200_batch* b, int want_monthly, int want_raw,
                                       help_b200_engine** out) {
  if (!out) return fail(HELP_B200_EINVAL, "out is NULL");
  *out = nullptr;
  int rc = validate(b);
  if (rc) return rc;
  if (help_b200_device_count() < 1) return fail(HELP_B200_ENODEV, "no CUDA device (this engine has no CPU path)");
  help_b200_engine* e = new (std::nothrow) help_b200_engine();
  if (!e) return fail(HELP_B200_ENOMEM, "host allocation failed");
  auto bail = [&](int code) { engine_free(e); return code; };
#define TRY(x) do { int _r = (x); if (_r) return bail(_r); } while (0)
#define CUB_(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) return bail(fail(_e == cudaErrorMemoryAllocation ? HELP_B200_ENOMEM : HELP_B200_ECUDA, #call, _e)); } while (0)
  CUB_(cudaGetDevice(&e->device));
  CUB_(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
  for (auto& ev : e->ev) CUB_(cudaEventCreate(&ev));
  CUB_(cudaStreamCreateWithFlags(&e->stream2, cudaStreamNonBlocking));
  CUB_(cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming));
  CUB_(cudaEventCreateWithFlags(&e->ev_join, cudaEventDisableTiming));
  const int64_t n = b->n_cells;
  const int ny = b->n_years, ml = b->max_layers;
  e->n = n; e->ny = ny;
  e->stride = ((n + 31) / 32) * 32;
  if (e->stride == 0) e->stride = 32;
  e->segcap = 7 + 3 * ml;
  if (e->segcap > kMaxSeg) e->segcap = kMaxSeg;
  e->maxseg_t = e->segcap <= 16 ? 16 : e->segcap <= 25 ? 25 : e->segcap <= 40 ? 40 : 67;
  e->want_monthly = want_monthly != 0; e->want_raw = want_raw != 0;
  const char* ns = getenv("HELP_B200_SORT");
  e->sort = !(ns && ns[0] == '0');
  e->alg_bytes = help_b200_algorithmic_bytes(b);
  const size_t wrow = (size_t)ny * kDays;
  TRY(dev_alloc(&e->d_years, ny));
  TRY(dev_alloc(&e->d_pre, (size_t)b->n_nodes_p * wrow));
  TRY(dev_alloc(&e->d_tmp, (size_t)b->n_nodes_t * wrow));
  TRY(dev_alloc(&e->d_rad, (size_t)b->n_nodes_r * wrow));
  TRY(dev_alloc(&e->d_tm, (size_t)b->n_nodes_t * 12));
  TRY(dev_alloc(&e->d_iu4, b->n_nodes_p)); TRY(dev_alloc(&e->d_iu7, b->n_nodes_t)); TRY(dev_alloc(&e->d_iu13, b->n_nodes_r));
  TRY(dev_alloc(&e->d_idfs, (size_t)b->n_nodes_r * 2));
  TRY(dev_alloc(&e->d_cell_f32, (size_t)HB_NCELL_F32 * n)); TRY(dev_alloc(&e->d_cell_i32, (size_t)HB_NCELL_I32 * n));
  TRY(dev_alloc(&e->d_layer_f32, (size_t)HL_NLAYER_F32 * ml * n));
  TRY(dev_alloc(&e->d_layer_i32, (size_t)HL_NLAYER_I32 * ml * n));
  TRY(dev_alloc(&e->d_layer_rc, (size_t)ml * n));
  TRY(dev_alloc(&e->d_img_f32, (size_t)Image::n_f32(e->segcap) * e->stride));
  TRY(dev_alloc(&e->d_img_i32, (size_t)Image::n_i32(e->segcap) * e->stride));
  TRY(dev_alloc(&e->d_img_f64, (size_t)Image::n_f64(e->segcap) * e->stride));
  TRY(dev_alloc(&e->d_img_rec, (size_t)Image::rec_bytes(e->stride, e->segcap)));
  if (e->want_monthly) TRY(dev_alloc(&e->d_monthly, (size_t)HR_NROWS * ny * 12 * n));
  if (e->want_raw) TRY(dev_alloc(&e->d_raw, (size_t)HELP_B200_RAW_ROWS * ny * 12 * n));
  TRY(dev_alloc(&e->d_yearly, (size_t)HY_NROWS * ny * n));
  TRY(dev_alloc(&e->d_flags, n)); TRY(dev_alloc(&e->d_topo, (size_t)16 * n));
  TRY(dev_alloc(&e->d_nnarrow, 1));
  TRY(dev_alloc(&e->d_key, n)); TRY(dev_alloc(&e->d_key2, n)); TRY(dev_alloc(&e->d_perm, n)); TRY(dev_alloc(&e->d_iota, n));
  CUB_(cub::DeviceRadixSort::SortPairs(nullptr, e->cub_bytes, e->d_key, e->d_key2, e->d_iota, e->d_perm, (int)n, 0, 64,
                                       e->stream));
  { uint8_t* q = nullptr; TRY(dev_alloc(&q, e->cub_bytes ? e->cub_bytes : 1)); e->d_cub = q; }
  // ---- upload (timed) -----------------------------------------------------------------
  cudaStream_t st = e->stream;
  CUB_(cudaEventRecord(e->ev[0], st));
#define H2D(dst, src, count) CUB_(cudaMemcpyAsync(dst, src, (size_t)(count) * sizeof(*(dst)), cudaMemcpyHostToDevice, st))
  H2D(e->d_years, b->years, ny);
  // (the weather tables may already be on this device -- help_b200_pack_weather_resident: the copy direction is
  // taken from the pointer; the engine converts its own copy to the reference's internal units in place)
#define ANY2D(dst, src, count) CUB_(cudaMemcpyAsync(dst, src, (size_t)(count) * sizeof(*(dst)), cudaMemcpyDefault, st))
  ANY2D(e->d_pre, b->pre, (size_t)b->n_nodes_p * wrow);
  ANY2D(e->d_tmp, b->tmp, (size_t)b->n_nodes_t * wrow);
  ANY2D(e->d_rad, b->rad, (size_t)b->n_nodes_r * wrow);
#undef ANY2D
  H2D(e->d_tm, b->tm_normals, (size_t)b->n_nodes_t * 12);
  H2D(e->d_iu4, b->iu4, b->n_nodes_p); H2D(e->d_iu7, b->iu7, b->n_nodes_t); H2D(e->d_iu13, b->iu13, b->n_nodes_r);
  H2D(e->d_idfs, b->idfs, (size_t)b->n_nodes_r * 2);
  if (n > 0) {
    H2D(e->d_cell_f32, b->cell_f32, (size_t)HB_NCELL_F32 * n); H2D(e->d_cell_i32, b->cell_i32, (size_t)HB_NCELL_I32 * n);
    H2D(e->d_layer_f32, b->layer_f32, (size_t)HL_NLAYER_F32 * ml * n);
    H2D(e->d_layer_i32, b->layer_i32, (size_t)HL_NLAYER_I32 * ml * n);
    H2D(e->d_layer_rc, b->layer_rc, (size_t)ml * n);
  }
  {
    const int th = 256;
    auto blocks = [&](size_t total) { size_t g = (total + th - 1) / th; return (unsigned)(g > 148 * 16 ? 148 * 16 : (g ? g : 1)); };
    size_t tp = (size_t)b->n_nodes_p * wrow, tt = (size_t)b->n_nodes_t * wrow, tr = (size_t)b->n_nodes_r * wrow;
    convert_weather_kernel<<<blocks(tp), th, 0, st>>>(e->d_pre, e->d_iu4, (int64_t)wrow, (int64_t)tp, 0);
    convert_weather_kernel<<<blocks(tt), th, 0, st>>>(e->d_tmp, e->d_iu7, (int64_t)wrow, (int64_t)tt, 1);
    convert_weather_kernel<<<blocks(tr), th, 0, st>>>(e->d_rad, e->d_iu13, (int64_t)wrow, (int64_t)tr, 2);
    if (n > 0) iota_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(e->d_iota, n);
    e->n_launches += 3 + (n > 0 ? 1 : 0);
  }
  CUB_(cudaEventRecord(e->ev[1], st));
  CUB_(cudaStreamSynchronize(st));
  CUB_(cudaGetLastError());
  CUB_(cudaEventElapsedTime(&e->ms_h2d, e->ev[0], e->ev[1]));
  BatchDev& B = e->B;
  B.n_cells = (int)n; B.n_years = ny; B.max_layers = ml;
  B.n_nodes_p = b->n_nodes_p; B.n_nodes_t = b->n_nodes_t; B.n_nodes_r = b->n_nodes_r;
  B.years = e->d_years; B.pre = e->d_pre; B.tmp = e->d_tmp; B.rad = e->d_rad;
  B.iu4 = e->d_iu4; B.iu7 = e->d_iu7; B.iu13 = e->d_iu13; B.tm_normals = e->d_tm; B.idfs = e->d_idfs;
  B.cell_f32 = e->d_cell_f32; B.cell_i32 = e->d_cell_i32; B.layer_f32 = e->d_layer_f32;
  B.layer_i32 = e->d_layer_i32; B.layer_rc = e->d_layer_rc;
  *out = e;
  return 0;
#undef H2D
#undef TRY
#undef CUB_
}

// Launch shapes of help_sim_kernel<16, ..> (help_sim.cuh): threads per SM they keep resident.
static const int kShapeCap[4] = {512, 640, 768, 1024};
// The shape for a batch of n cells: the fewest waves the largest shape allows, filled evenly,
// with the most registers that still fit.  HELP_B200_SHAPE overrides (developer A/B runs).
static int pick_shape(int64_t n, int device) {
  if (const char* f = getenv("HELP_B200_SHAPE")) { const int v = atoi(f); if (v >= 0 && v <= 3) return v; }
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  const int64_t need = (n + sms - 1) / sms;                     // threads per SM
  const int64_t waves = (need + kShapeCap[3] - 1) / kShapeCap[3];
  const int64_t per = (need + waves - 1) / (waves > 0 ? waves : 1);
  for (int k = 0; k < 4; ++k) if (per <= kShapeCap[k]) return k;
  return 3;
}

// Threads per block = the warps that share the day barrier of help_sim_kernel.  Plain (natural-soil) profiles run best in
// blocks of 128 (larger groups wait longer at the barrier: 100 000 cells 684 / 686 / 699 / 759 ms for blocks of 128 / 256 /
// 352 / 704).  The routing code of a profile with a liner is larger than the instruction cache, so the warps of an SM should
// fetch it together: ONE block per SM, as wide as the batch fills the SMs in evenly filled waves (125 000 mixed cells 1181 /
// 1002 / 966 ms for 128 / 256 / 512; 31 250 landfill covers 810 / 776 ms for 128 / 224; profiles/r2_notes.md).
enum LaunchKind : int { kPlainOnly = 0, kLinerOnly = 1, kMixed = 2 };   // profile classes of a launch
static int pick_block(int64_t n, int cap, int kind, int device) {
  if (const char* f = getenv("HELP_B200_BLOCK")) { const int v = atoi(f); if (v >= 32 && v <= cap) return v; }
  if (kind == kPlainOnly) return 128;
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  const int64_t per = (n + sms - 1) / sms;                      // threads per SM
  // plain and liner profiles in one launch: the liner blocks (launched first, several times the run time of a plain one) are
  // the critical path, plain blocks fill in behind them -- the widest block (448: 1013 ms, 512: 964 ms for 125 000 mixed cells)
  if (kind == kMixed && per > cap) return cap;
  const int64_t waves = (per + cap - 1) / cap;
  const int64_t per_wave = (per + waves - 1) / (waves > 0 ? waves : 1);
  int64_t b = (per_wave + 31) / 32 * 32;
  if (b < 128) b = 128;
  if (b > cap) b = cap;
  return (int)b;
}

// CAP = threads per SM the register budget keeps resident = the largest block the instantiation accepts.
template <int MAXSEG, int CAP, int REGS>
static void launch_sim(const SimArgs& A, cudaStream_t st, int kind, int device) {
  const int block = pick_block(A.count, CAP, kind, device);
  help_sim_kernel<MAXSEG, CAP, REGS><<<(unsigned)((A.count + block - 1) / block), block, 0, st>>>(A);
}

// One launch of image slots [A.first, A.first + A.count) with the instantiation for `maxseg_t` segments;
// the 16-segment one comes in four register budgets (pick_shape).
static int launch_tier(const SimArgs& A, int maxseg_t, int device, cudaStream_t st, int kind) {
  switch (maxseg_t) {
    case 16:
      switch (pick_shape(A.count, device)) {
        case 0: launch_sim<16, 512, 128>(A, st, kind, device); break;
        case 1: launch_sim<16, 640, 96>(A, st, kind, device); break;
        case 2: launch_sim<16, 768, 80>(A, st, kind, device); break;
        default: launch_sim<16, 1024, 64>(A, st, kind, device); break;
      }
      return 0;
#ifndef HELP_DEV_FAST_BUILD
    // (96 / 80 registers for the wide tiers: 125 000 mixed cells 974 -> 1007 / 1108 ms, profiles/r2_notes.md)
    case 25: launch_sim<25, 512, 128>(A, st, kind, device); return 0;
    case 40: launch_sim<40, 512, 128>(A, st, kind, device); return 0;
    default: launch_sim<67, 512, 128>(A, st, kind, device); return 0;
#else
    default: return fail(HELP_B200_EINVAL, "developer build: MAXSEG 16 only");
#endif
  }
}

// Launch plan.  The sort puts the narrow tier -- plain profiles of at most 16 segments, i.e. every natural-soil cell --
// in front.  A batch that holds nothing else runs the 16-segment instantiation whatever its deepest profile would allow
// (a shard of a class-sorted grid, pyhelp_b200.sharding.plan_shards, is such a batch).  A batch that also holds liner
// profiles runs ONE launch of the widest instantiation: launching the tiers separately (HELP_B200_PLAN = 1: narrow first,
// 2: wide first, on two streams) was measured slower in round 1 and again in round 2 (125 000 mixed cells x 6 simulated
// years: 1238 / 1361 / 1618 ms, profiles/r2_plan_ab.log) -- the wide tier alone cannot fill the SMs and becomes the tail.
extern "C" int help_b200_engine_simulate(help_b200_engine* e, void* stream) {
  if (!e) return fail(HELP_B200_EINVAL, "engine is NULL");
  CU(cudaSetDevice(e->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : e->stream;
  const int64_t n = e->n;
  e->n_launches = 0;
  Image img{e->d_img_f32, e->d_img_i32, e->d_img_f64, e->stride, e->segcap, e->d_img_rec};
  CU(cudaEventRecord(e->ev[0], st));
  if (n > 0) {
    const unsigned gs = (unsigned)((n + 63) / 64);
    const int32_t* perm = nullptr;
    int64_t n_narrow = 0;
    if (e->sort) {
      CU(cudaMemsetAsync(e->d_nnarrow, 0, sizeof(int32_t), st));
      help_setup_kernel<<<gs, 64, 0, st>>>(e->B, img, nullptr, nullptr, nullptr, e->d_key, e->d_nnarrow);
      CU(cub::DeviceRadixSort::SortPairs(e->d_cub, e->cub_bytes, e->d_key, e->d_key2, e->d_iota, e->d_perm, (int)n, 0, 64, st));
      perm = e->d_perm;
      e->n_launches += 2;   // + the radix-sort passes, which are library kernels and not counted
      int32_t h = 0;
      CU(cudaMemcpyAsync(&h, e->d_nnarrow, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
      CU(cudaStreamSynchronize(st));
      n_narrow = h;
    }
    e->n_narrow = n_narrow;
    help_setup_kernel<<<gs, 64, 0, st>>>(e->B, img, perm, e->d_flags, e->d_topo, nullptr, nullptr);
    e->n_launches += 1;
    CU(cudaEventRecord(e->ev[1], st));
    SimArgs A;
    A.img = img; A.pre = e->d_pre; A.tmp = e->d_tmp; A.rad = e->d_rad; A.years = e->d_years;
    A.n_years = e->ny; A.n_cells = n; A.first = 0; A.count = n; A.perm = perm;
    A.monthly = e->d_monthly; A.yearly = e->d_yearly; A.raw = e->d_raw; A.flags = e->d_flags;
    A.daily = e->d_daily; A.daily_row = e->d_daily_row; A.daily_days = e->daily_days;
    int plan = 0;
    if (const char* f = getenv("HELP_B200_PLAN")) plan = atoi(f);
    int rc = 0;
    if (n_narrow == n && e->sort) {
      rc = launch_tier(A, 16, e->device, st, kPlainOnly);                     // every cell is narrow, whatever max_layers says
      e->n_launches += 1;
    } else if (plan == 0 || n_narrow == 0 || !e->sort) {
      rc = launch_tier(A, e->maxseg_t, e->device, st, (e->sort && n_narrow > 0) ? kMixed : kLinerOnly);
      e->n_launches += 1;
    } else {
      SimArgs N1 = A, W = A;
      N1.first = 0; N1.count = n_narrow;
      W.first = n_narrow; W.count = n - n_narrow;
      CU(cudaEventRecord(e->ev_fork, st));
      CU(cudaStreamWaitEvent(e->stream2, e->ev_fork, 0));
      if (plan == 2) {
        rc = launch_tier(W, e->maxseg_t, e->device, st, kLinerOnly);
        if (!rc) rc = launch_tier(N1, 16, e->device, e->stream2, kPlainOnly);
      } else {
        rc = launch_tier(N1, 16, e->device, st, kPlainOnly);
        if (!rc) rc = launch_tier(W, e->maxseg_t, e->device, e->stream2, kLinerOnly);
      }
      CU(cudaEventRecord(e->ev_join, e->stream2));
      CU(cudaStreamWaitEvent(st, e->ev_join, 0));
      e->n_launches += 2;
    }
    if (rc) return rc;
  } else {
    CU(cudaEventRecord(e->ev[1], st));
  }
  CU(cudaEventRecord(e->ev[2], st));
  CU(cudaStreamSynchronize(st));
  CU(cudaGetLastError());
  CU(cudaEventElapsedTime(&e->ms_setup, e->ev[0], e->ev[1]));
  CU(cudaEventElapsedTime(&e->ms_sim, e->ev[1], e->ev[2]));
  return 0;
}

extern "C" int help_b200_engine_device_ptrs(help_b200_engine* e, void** monthly, void** yearly, void** raw, void** flags) {
  if (!e) return fail(HELP_B200_EINVAL, "engine is NULL");
  if (monthly) *monthly = e->d_monthly;
  if (yearly) *yearly = e->d_yearly;
  if (raw) *raw = e->d_raw;
  if (flags) *flags = e->d_flags;
  return 0;
}

extern "C" int help_b200_engine_timings(help_b200_engine* e, float* ms_setup, float* ms_sim, int32_t* n_launches) {
  if (!e) return fail(HELP_B200_EINVAL, "engine is NULL");
  if (ms_setup) *ms_setup = e->ms_setup;
  if (ms_sim) *ms_sim = e->ms_sim;
  if (n_launches) *n_launches = e->n_launches;
  return 0;
}

extern "C" int help_b200_engine_fetch(help_b200_engine* e, help_b200_result* r) {
  if (!e || !r) return fail(HELP_B200_EINVAL, "engine or result is NULL");
  CU(cudaSetDevice(e->device));
  cudaStream_t st = e->stream;
  const size_t n = (size_t)e->n, ny = (size_t)e->ny;
  CU(cudaEventRecord(e->ev[0], st));
  if (r->monthly) {
    if (!e->d_monthly) return fail(HELP_B200_EINVAL, "monthly was not requested at engine_create");
    CU(cudaMemcpyAsync(r->monthly, e->d_monthly, sizeof(float) * HR_NROWS * ny * 12 * n, cudaMemcpyDeviceToHost, st));
  }
  if (r->raw_monthly) {
    if (!e->d_raw) return fail(HELP_B200_EINVAL, "raw_monthly was not requested at engine_create");
    CU(cudaMemcpyAsync(r->raw_monthly, e->d_raw, sizeof(float) * HELP_B200_RAW_ROWS * ny * 12 * n, cudaMemcpyDeviceToHost, st));
  }
  if (r->yearly) CU(cudaMemcpyAsync(r->yearly, e->d_yearly, sizeof(float) * HY_NROWS * ny * n, cudaMemcpyDeviceToHost, st));
  if (r->flags) CU(cudaMemcpyAsync(r->flags, e->d_flags, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, st));
  if (r->topo) CU(cudaMemcpyAsync(r->topo, e->d_topo, sizeof(int32_t) * 16 * n, cudaMemcpyDeviceToHost, st));
  CU(cudaEventRecord(e->ev[1], st));
  CU(cudaStreamSynchronize(st));
  CU(cudaEventElapsedTime(&r->ms_d2h, e->ev[0], e->ev[1]));
  r->ms_h2d = e->ms_h2d; r->ms_setup = e->ms_setup; r->ms_simulate = e->ms_sim; r->n_launches = e->n_launches;
  return 0;
}

extern "C" int help_b200_engine_fetch_monthly_by_cell(help_b200_engine* e, float* out) {
  if (!e || !out) return fail(HELP_B200_EINVAL, "engine or out is NULL");
  if (!e->d_monthly) return fail(HELP_B200_EINVAL, "monthly was not requested at engine_create");
  CU(cudaSetDevice(e->device));
  const int64_t n = e->n, rows = (int64_t)HR_NROWS * e->ny * 12;
  if (n == 0) return 0;
  float* d_t = nullptr;
  int rc = dev_alloc(&d_t, (size_t)rows * n);
  if (rc) return rc;
  cudaStream_t st = e->stream;
  const dim3 grid((unsigned)((n + 31) / 32), (unsigned)((rows + 31) / 32));
  transpose_to_cell_major_kernel<<<grid, dim3(32, 8), 0, st>>>(e->d_monthly, rows, n, d_t);
  cudaError_t ce = cudaGetLastError();
  if (ce == cudaSuccess) ce = cudaMemcpyAsync(out, d_t, sizeof(float) * rows * n, cudaMemcpyDeviceToHost, st);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
  dev_free(d_t);
  if (ce != cudaSuccess) return fail(HELP_B200_ECUDA, "fetch_monthly_by_cell", ce);
  return 0;
}

extern "C" int help_b200_engine_select_daily(help_b200_engine* e, const int32_t* cells, int32_t n_selected) {
  if (!e) return fail(HELP_B200_EINVAL, "engine is NULL");
  if (n_selected < 0 || (n_selected > 0 && !cells)) return fail(HELP_B200_EINVAL, "select_daily: bad argument");
  CU(cudaSetDevice(e->device));
  cudaDeviceSynchronize();
  if (e->d_daily) { dev_free(e->d_daily); e->d_daily = nullptr; }
  if (e->d_daily_row) { dev_free(e->d_daily_row); e->d_daily_row = nullptr; }
  e->n_daily = 0;
  if (n_selected == 0) return 0;
  std::vector<int32_t> row((size_t)e->n, -1);
  for (int32_t k = 0; k < n_selected; ++k) {
    if (cells[k] < 0 || cells[k] >= e->n) return fail(HELP_B200_EINVAL, "select_daily: cell index out of range");
    row[cells[k]] = k;
  }
  std::vector<int32_t> years(e->ny);
  CU(cudaMemcpy(years.data(), e->d_years, sizeof(int32_t) * e->ny, cudaMemcpyDeviceToHost));
  e->daily_days = 0;
  for (int y = 0; y < e->ny; ++y) e->daily_days += (years[y] % 4 == 0) ? 366 : 365;      // LEAP, F:1048
  int rc;
  if ((rc = dev_alloc(&e->d_daily_row, (size_t)e->n)) || (rc = dev_alloc(&e->d_daily, (size_t)n_selected * e->daily_days * kDailyCols))) return rc;
  // on the engine's stream and complete on return: the legacy default stream does not order against
  // e->stream (non-blocking) or a caller's stream, where the next simulate() runs
  CU(cudaMemcpyAsync(e->d_daily_row, row.data(), sizeof(int32_t) * e->n, cudaMemcpyHostToDevice, e->stream));
  CU(cudaMemsetAsync(e->d_daily, 0, sizeof(double) * (size_t)n_selected * e->daily_days * kDailyCols, e->stream));
  CU(cudaStreamSynchronize(e->stream));
  e->n_daily = n_selected;
  return 0;
}

extern "C" int help_b200_engine_fetch_daily(help_b200_engine* e, double* out, int64_t* n_days) {
  if (!e) return fail(HELP_B200_EINVAL, "engine is NULL");
  if (n_days) *n_days = e->daily_days;
  if (!e->d_daily || e->n_daily == 0) return fail(HELP_B200_EINVAL, "fetch_daily: no cell was selected (select_daily before simulate)");
  if (!out) return 0;
  CU(cudaSetDevice(e->device));
  CU(cudaMemcpy(out, e->d_daily, sizeof(double) * (size_t)e->n_daily * e->daily_days * kDailyCols, cudaMemcpyDeviceToHost));
  return 0;
}

extern "C" int help_b200_engine_post_process(help_b200_engine* e, const int32_t* context, int32_t year_from, int32_t year_to,
                                             double* area_monthly, double* cells_yearly, uint8_t* nan_mask) {
  if (!e) return fail(HELP_B200_EINVAL, "engine is NULL");
  if (!e->d_monthly) return fail(HELP_B200_EINVAL, "post_process needs an engine created with want_monthly");
  CU(cudaSetDevice(e->device));
  const int64_t n = e->n; const int ny = e->ny;
  if (n == 0) return 0;
  std::vector<int32_t> years(ny);
  CU(cudaMemcpy(years.data(), e->d_years, sizeof(int32_t) * ny, cudaMemcpyDeviceToHost));
  int y0 = ny, y1 = 0;                                   // selected years are a contiguous range (years ascend)
  for (int y = 0; y < ny; ++y) if (years[y] >= year_from && years[y] <= year_to) { if (y < y0) y0 = y; y1 = y + 1; }
  if (y1 < y0) y1 = y0;
  int32_t* d_ctx = nullptr; int8_t* d_mode = nullptr; double *d_cy = nullptr, *d_area = nullptr; int32_t* d_nnan = nullptr;
  int rc = 0;
  auto done = [&](int code) {
    cudaDeviceSynchronize();
    for (void* p : {(void*)d_ctx, (void*)d_mode, (void*)d_cy, (void*)d_area, (void*)d_nnan}) dev_free(p);
    return code;
  };
  if ((rc = dev_alloc(&d_mode, (size_t)n)) || (rc = dev_alloc(&d_area, (size_t)7 * ny * 12)) || (rc = dev_alloc(&d_nnan, 1)) ||
      (context && (rc = dev_alloc(&d_ctx, (size_t)n))) || (cells_yearly && (rc = dev_alloc(&d_cy, (size_t)7 * n))))
    return done(rc);
  cudaError_t ce = cudaSuccess;
  if (context) ce = cudaMemcpy(d_ctx, context, sizeof(int32_t) * n, cudaMemcpyHostToDevice);
  if (ce == cudaSuccess) {
    post_cell_kernel<<<(unsigned)((n + 127) / 128), 128>>>(e->d_monthly, d_ctx, n, ny, y0, y1, d_mode, d_cy);
    post_area_kernel<<<(unsigned)(ny * 12), 256>>>(e->d_monthly, d_mode, n, ny, d_area, d_nnan);
    ce = cudaGetLastError();
  }
  int32_t n_nan = 0;
  if (ce == cudaSuccess) ce = cudaMemcpy(&n_nan, d_nnan, sizeof(int32_t), cudaMemcpyDeviceToHost);
  if (ce == cudaSuccess && area_monthly) {
    ce = cudaMemcpy(area_monthly, d_area, sizeof(double) * 7 * ny * 12, cudaMemcpyDeviceToHost);
    const double np_valid = (double)(n - n_nan);         // Np = cells - NaN cells (output.py:138)
    if (ce == cudaSuccess) for (int64_t i = 0; i < (int64_t)7 * ny * 12; ++i) area_monthly[i] /= np_valid;
  }
  if (ce == cudaSuccess && cells_yearly) ce = cudaMemcpy(cells_yearly, d_cy, sizeof(double) * 7 * n, cudaMemcpyDeviceToHost);
  if (ce == cudaSuccess && nan_mask) {
    std::vector<int8_t> m((size_t)n);
    ce = cudaMemcpy(m.data(), d_mode, (size_t)n, cudaMemcpyDeviceToHost);
    if (ce == cudaSuccess) for (int64_t i = 0; i < n; ++i) nan_mask[i] = (m[i] == 1);
  }
  if (ce != cudaSuccess) return done(fail(HELP_B200_ECUDA, "post_process", ce));
  return done(0);
}

extern "C" void help_b200_engine_destroy(help_b200_engine* e) { engine_free(e); }

// ---- diagnostic: evaluate the engine's transcendental functions on the device -----------------
__global__ void math_eval_kernel(int fn, const double* __restrict__ x, const double* __restrict__ y,
                                 double* __restrict__ out, int64_t n) {
  helpmath::hm_load_tables();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double a = x[i], b = y ? y[i] : 0.0;
  const float af = (float)a, bf = (float)b;
  double r = 0.0;
  switch (fn) {
    case 0: r = r8_pow(a, b); break;
    case 1: r = helpmath::det_exp(a); break;
    case 2: r = r8_log(a); break;
    case 3: r = r8_sin(a); break;
    case 4: r = r8_cos(a); break;
    case 5: r = r8_atan(a); break;
    case 6: r = r4_acos(af); break;
    case 7: r = r4_tan(af); break;
    case 8: r = r4_pow(af, bf); break;
    case 9: r = r4_exp(af); break;
    case 10: r = r4_log(af); break;
    case 11: r = r4_sin(af); break;
    case 12: r = r4_cos(af); break;
    case 13: r = r4_log10(af); break;
    case 14: r = r4_sqrt(af); break;
    case 15: r = r4_pow8(af); break;
    case 16: r = r4_pow7(af); break;
    case 17: r = r4_pow4(af); break;
    case 18: { r = helpmath::det_pow2(a, b, b == 0.0 ? 1.5 : 1.0 / b, a).a; break; }
    case 19: { r = helpmath::det_pow_samebase(a, b - 1.0, b).b; break; }
    case 20: { const double q = a / b; r = helpmath::div_by_const(a, b, 1.0 / b); if (q != q) r = q; break; }
    default: break;
  }
  out[i] = r;
}

extern "C" int help_b200_math_eval(int fn, const double* x, const double* y, double* out, int64_t n) {
  if (fn < 0 || fn > 20 || !x || !out || n < 0) return fail(HELP_B200_EINVAL, "math_eval: bad argument");
  if (help_b200_device_count() < 1) return fail(HELP_B200_ENODEV, "no CUDA device (this engine has no CPU path)");
  if (n == 0) return 0;
  double *dx = nullptr, *dy = nullptr, *d_out = nullptr;
  int rc = 0;
  auto done = [&](int code) { cudaDeviceSynchronize(); dev_free(dx); dev_free(dy); dev_free(d_out); return code; };
  if ((rc = dev_alloc(&dx, (size_t)n)) || (rc = dev_alloc(&d_out, (size_t)n)) || (y && (rc = dev_alloc(&dy, (size_t)n)))) return done(rc);
  cudaError_t ce = cudaMemcpy(dx, x, sizeof(double) * n, cudaMemcpyHostToDevice);
  if (ce == cudaSuccess && y) ce = cudaMemcpy(dy, y, sizeof(double) * n, cudaMemcpyHostToDevice);
  if (ce == cudaSuccess) {
    math_eval_kernel<<<(unsigned)((n + 127) / 128), 128>>>(fn, dx, dy, d_out, n);
    ce = cudaGetLastError();
  }
  if (ce == cudaSuccess) ce = cudaMemcpy(out, d_out, sizeof(double) * n, cudaMemcpyDeviceToHost);
  if (ce != cudaSuccess) return done(fail(HELP_B200_ECUDA, "math_eval", ce));
  return done(0);
}

extern "C" int help_b200_nearest_nodes(int64_t n_cells, const double* cell_lat, const double* cell_lon, int32_t n_nodes,
                                       const double* node_lat, const double* node_lon, int32_t* out_index) {
  if (n_cells < 0 || n_nodes < 1 || !node_lat || !node_lon || (n_cells > 0 && (!cell_lat || !cell_lon || !out_index)))
    return fail(HELP_B200_EINVAL, "nearest_nodes: bad argument");
  if (help_b200_device_count() < 1) return fail(HELP_B200_ENODEV, "no CUDA device (this engine has no CPU path)");
  if (n_cells == 0) return 0;
  double *d_clat = nullptr, *d_clon = nullptr, *d_nlat = nullptr, *d_nlon = nullptr;
  NodeVec* d_vec = nullptr; int32_t* d_out = nullptr;
  int rc = 0;
  auto done = [&](int code) {
    cudaDeviceSynchronize();
    for (void* p : {(void*)d_clat, (void*)d_clon, (void*)d_nlat, (void*)d_nlon, (void*)d_vec, (void*)d_out}) dev_free(p);
    return code;
  };
  if ((rc = dev_alloc(&d_clat, (size_t)n_cells)) || (rc = dev_alloc(&d_clon, (size_t)n_cells)) || (rc = dev_alloc(&d_nlat, (size_t)n_nodes)) ||
      (rc = dev_alloc(&d_nlon, (size_t)n_nodes)) || (rc = dev_alloc(&d_vec, (size_t)n_nodes)) || (rc = dev_alloc(&d_out, (size_t)n_cells)))
    return done(rc);
  cudaError_t ce = cudaMemcpy(d_clat, cell_lat, sizeof(double) * n_cells, cudaMemcpyHostToDevice);
  if (ce == cudaSuccess) ce = cudaMemcpy(d_clon, cell_lon, sizeof(double) * n_cells, cudaMemcpyHostToDevice);
  if (ce == cudaSuccess) ce = cudaMemcpy(d_nlat, node_lat, sizeof(double) * n_nodes, cudaMemcpyHostToDevice);
  if (ce == cudaSuccess) ce = cudaMemcpy(d_nlon, node_lon, sizeof(double) * n_nodes, cudaMemcpyHostToDevice);
  if (ce == cudaSuccess) {
    node_vectors_kernel<<<(unsigned)((n_nodes + 255) / 256), 256>>>(d_nlat, d_nlon, n_nodes, d_vec);
    nearest_node_kernel<<<(unsigned)((n_cells + 255) / 256), 256>>>(d_clat, d_clon, n_cells, d_vec, n_nodes, d_out);
    ce = cudaGetLastError();
  }
  if (ce == cudaSuccess) ce = cudaMemcpy(out_index, d_out, sizeof(int32_t) * n_cells, cudaMemcpyDeviceToHost);
  if (ce != cudaSuccess) return done(fail(HELP_B200_ECUDA, "nearest_nodes", ce));
  return done(0);
}

// pack_weather for both entry points: the packed table is copied to `out_host` and/or handed over in `*keep_dev`
static int pack_weather_impl(const double* table, int64_t n_days, int32_t n_nodes, const int32_t* year_day0,
                             const int32_t* year_ndays, int32_t n_years, int32_t decimals, float* out_host, float** keep_dev) {
  if (!table || !year_day0 || !year_ndays || (!out_host && !keep_dev) || n_days < 1 || n_nodes < 1 || n_years < 1 || decimals < 0 || decimals > 15)
    return fail(HELP_B200_EINVAL, "pack_weather: bad argument");
  for (int y = 0; y < n_years; ++y)
    if (year_day0[y] < 0 || year_ndays[y] < 0 || year_ndays[y] > kDays || (int64_t)year_day0[y] + year_ndays[y] > n_days)
      return fail(HELP_B200_EINVAL, "pack_weather: a year's day range is outside the table");
  if (help_b200_device_count() < 1) return fail(HELP_B200_ENODEV, "no CUDA device (this engine has no CPU path)");
  double* d_tab = nullptr; float* d_out = nullptr; int32_t *d_d0 = nullptr, *d_nd = nullptr;
  int rc = 0;
  auto done = [&](int code) {
    cudaDeviceSynchronize();
    dev_free(d_tab); dev_free(d_d0); dev_free(d_nd);
    if (code == 0 && keep_dev) *keep_dev = d_out; else dev_free(d_out);
    return code;
  };
  const size_t n_in = (size_t)n_days * n_nodes, n_out = (size_t)n_nodes * n_years * kDays;
  if ((rc = dev_alloc(&d_tab, n_in)) || (rc = dev_alloc(&d_out, n_out)) || (rc = dev_alloc(&d_d0, (size_t)n_years)) ||
      (rc = dev_alloc(&d_nd, (size_t)n_years)))
    return done(rc);
  cudaError_t ce = cudaMemcpy(d_tab, table, sizeof(double) * n_in, cudaMemcpyHostToDevice);
  if (ce == cudaSuccess) ce = cudaMemcpy(d_d0, year_day0, sizeof(int32_t) * n_years, cudaMemcpyHostToDevice);
  if (ce == cudaSuccess) ce = cudaMemcpy(d_nd, year_ndays, sizeof(int32_t) * n_years, cudaMemcpyHostToDevice);
  if (ce == cudaSuccess) {
    double S = 1.0;
    for (int k = 0; k < decimals; ++k) S *= 10.0;
    const dim3 grid((kDays + 31) / 32, (unsigned)n_years, (unsigned)((n_nodes + 31) / 32));
    pack_weather_kernel<<<grid, dim3(32, 8)>>>(d_tab, (int64_t)n_nodes, d_d0, d_nd, n_years, S, d_out);
    ce = cudaGetLastError();
  }
  if (ce == cudaSuccess && out_host) ce = cudaMemcpy(out_host, d_out, sizeof(float) * n_out, cudaMemcpyDeviceToHost);
  if (ce != cudaSuccess) return done(fail(HELP_B200_ECUDA, "pack_weather", ce));
  return done(0);
}

extern "C" int help_b200_pack_weather(const double* table, int64_t n_days, int32_t n_nodes, const int32_t* year_day0,
                                      const int32_t* year_ndays, int32_t n_years, int32_t decimals, float* out) {
  return pack_weather_impl(table, n_days, n_nodes, year_day0, year_ndays, n_years, decimals, out, nullptr);
}

extern "C" int help_b200_pack_weather_resident(const double* table, int64_t n_days, int32_t n_nodes, const int32_t* year_day0,
                                               const int32_t* year_ndays, int32_t n_years, int32_t decimals, float** out_dev) {
  if (!out_dev) return fail(HELP_B200_EINVAL, "pack_weather_resident: out_dev is NULL");
  *out_dev = nullptr;
  return pack_weather_impl(table, n_days, n_nodes, year_day0, year_ndays, n_years, decimals, nullptr, out_dev);
}

extern "C" int help_b200_idfs_prescan_resident(const float* rad_dev, int32_t n_nodes, int32_t n_years, int32_t iu13, int32_t* out) {
  if (!rad_dev || !out || n_nodes < 1 || n_years < 1) return fail(HELP_B200_EINVAL, "idfs_prescan_resident: bad argument");
  if (help_b200_device_count() < 1) return fail(HELP_B200_ENODEV, "no CUDA device (this engine has no CPU path)");
  int32_t* d_out = nullptr;
  int rc = dev_alloc(&d_out, (size_t)2 * n_nodes);
  if (rc) return rc;
  idfs_prescan_kernel<<<(unsigned)((2 * n_nodes + 127) / 128), 128>>>(rad_dev, n_nodes, n_years, iu13, d_out);
  cudaError_t ce = cudaGetLastError();
  if (ce == cudaSuccess) ce = cudaMemcpy(out, d_out, sizeof(int32_t) * 2 * n_nodes, cudaMemcpyDeviceToHost);
  cudaDeviceSynchronize();
  dev_free(d_out);
  if (ce != cudaSuccess) return fail(HELP_B200_ECUDA, "idfs_prescan_resident", ce);
  return 0;
}

extern "C" void help_b200_device_free(void* p) {
  if (!p) return;
  cudaDeviceSynchronize();
  dev_free(p);
}

extern "C" int help_b200_parse_fields(const uint8_t* text, int64_t n_records, int32_t reclen, const int32_t* col,
                                      const int32_t* width, const int32_t* decimals, int32_t n_fields, double* out, uint8_t* flags) {
  if (!text || !col || !width || !decimals || !out || !flags || n_records < 0 || reclen < 1 || n_fields < 1 || n_fields > 64)
    return fail(HELP_B200_EINVAL, "parse_fields: bad argument");
  for (int f = 0; f < n_fields; ++f)
    if (col[f] < 0 || width[f] < 1 || col[f] >= reclen || decimals[f] < 0 || decimals[f] > 22)
      return fail(HELP_B200_EINVAL, "parse_fields: a field lies outside the record");
  if (help_b200_device_count() < 1) return fail(HELP_B200_ENODEV, "no CUDA device (this engine has no CPU path)");
  if (n_records == 0) return 0;
  uint8_t *d_text = nullptr, *d_flag = nullptr; int32_t* d_spec = nullptr; double* d_out = nullptr;
  int rc = 0;
  auto done = [&](int code) { cudaDeviceSynchronize(); dev_free(d_text); dev_free(d_flag); dev_free(d_spec); dev_free(d_out); return code; };
  const size_t nt = (size_t)n_records * reclen, nv = (size_t)n_records * n_fields;
  if ((rc = dev_alloc(&d_text, nt)) || (rc = dev_alloc(&d_flag, nv)) || (rc = dev_alloc(&d_spec, (size_t)3 * n_fields)) || (rc = dev_alloc(&d_out, nv)))
    return done(rc);
  cudaError_t ce = cudaMemcpy(d_text, text, nt, cudaMemcpyHostToDevice);
  if (ce == cudaSuccess) ce = cudaMemcpy(d_spec, col, sizeof(int32_t) * n_fields, cudaMemcpyHostToDevice);
  if (ce == cudaSuccess) ce = cudaMemcpy(d_spec + n_fields, width, sizeof(int32_t) * n_fields, cudaMemcpyHostToDevice);
  if (ce == cudaSuccess) ce = cudaMemcpy(d_spec + 2 * n_fields, decimals, sizeof(int32_t) * n_fields, cudaMemcpyHostToDevice);
  if (ce == cudaSuccess) {
    parse_fields_kernel<<<(unsigned)((nv + 255) / 256), 256>>>(d_text, n_records, reclen, d_spec, d_spec + n_fields, d_spec + 2 * n_fields,
                                                              n_fields, d_out, d_flag);
    ce = cudaGetLastError();
  }
  if (ce == cudaSuccess) ce = cudaMemcpy(out, d_out, sizeof(double) * nv, cudaMemcpyDeviceToHost);
  if (ce == cudaSuccess) ce = cudaMemcpy(flags, d_flag, nv, cudaMemcpyDeviceToHost);
  if (ce != cudaSuccess) return done(fail(HELP_B200_ECUDA, "parse_fields", ce));
  return done(0);
}

extern "C" int help_b200_run(const help_b200_batch* b, help_b200_result* r) {
  if (!r) return fail(HELP_B200_EINVAL, "result is NULL");
  help_b200_engine* e = nullptr;
  int rc = help_b200_engine_create(b, r->monthly != nullptr, r->raw_monthly != nullptr, &e);
  if (rc) return rc;
  rc = help_b200_engine_simulate(e, nullptr);
  if (!rc) rc = help_b200_engine_fetch(e, r);
  help_b200_engine_destroy(e);
  return rc;
}
