/* help_b200.h -- C ABI of the B200-native HELP water-balance engine.
 *
 * This is the drop-in boundary for the one hot path this repository
 * accelerates: PyHELP's per-cell daily HELP simulation,
 *
 *   run_help_allcells()                 reference pyhelp/processing.py:40-59
 *     -> HELP3O.run_simulation(...)     reference pyhelp/HELP3O.FOR:17-39 (f2py; meson.build:25-41)
 *     -> read_monthly_help_output()     reference pyhelp/processing.py:64-147
 *
 * The reference crosses that boundary once per cell (12-tuple in, one text
 * file out).  This ABI crosses it once per BATCH of cells: the caller hands
 * over the numeric content of the D4/D7/D13 weather files and of the D10/D11
 * records as structure-of-arrays buffers (already parsed with Fortran
 * formatted-read semantics, nothing else pre-computed) and receives the
 * arrays read_monthly_help_output() would have produced for every cell.
 *
 * Plain C: pointers and sizes only, no CUDA or torch types.  All functions
 * return 0 on success or a negative HELP_B200_E* code; none of them aborts
 * the process (the reference STOPs on a missing file, HELP3O.FOR:1202-1210).
 * The library never falls back to a CPU path: without a CUDA device every
 * compute entry returns HELP_B200_ENODEV.
 */
#ifndef HELP_B200_H
#define HELP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HELP_B200_ABI_VERSION 1
#define HELP_B200_MAX_LAYERS 20   /* HELP3O.FOR:62-66 (arrays of 20 layers)   */
#define HELP_B200_MAX_SEGS 67     /* HELP3O.FOR:70 (67 modelling segments)    */
#define HELP_B200_MAX_YEARS 100   /* HELP3O.FOR:22 (MXYR)                     */
#define HELP_B200_DAYS 368        /* row stride of one weather year (floats)  */

enum {
  HELP_B200_OK = 0,
  HELP_B200_EINVAL = -1,   /* bad argument / inconsistent sizes              */
  HELP_B200_ENODEV = -2,   /* no usable CUDA device                          */
  HELP_B200_ECUDA = -3,    /* CUDA runtime error (see help_b200_last_error)  */
  HELP_B200_ENOMEM = -4
};

/* ---- per-cell scalar fields: cell_f32[field][n_cells] ---------------------
 * D11 record 3 (HELP3O.FOR:1222-1225) and D10 records 2-3 (HELP3O.FOR:1264-1281),
 * values exactly as a Fortran formatted READ yields them (no unit conversion). */
enum {
  HB_ULAT = 0,   /* latitude, degrees                              (D11 F10.2) */
  HB_ULAI,       /* maximum leaf area index                        (D11 F7.0)  */
  HB_EDEPTH,     /* evaporative zone depth, cm if IU11=2 else in   (D11 F8.0)  */
  HB_WIND,       /* wind speed, km/h if IU11=2 else mph            (D11 F5.0)  */
  HB_HUM1, HB_HUM2, HB_HUM3, HB_HUM4, /* quarterly relative humidity, %       */
  HB_OSNO,       /* initial snow water                             (D10 F10.0) */
  HB_FRUNOF,     /* percent of area allowing runoff                (D10 F6.0)  */
  HB_CN2,        /* SCS curve number                               (D10 F7.0)  */
  HB_TFSOIL,     /* frozen-soil threshold, deg F (run_simulation arg TFSOIL)   */
  HB_NCELL_F32
};
enum {
  HB_IU11 = 0,   /* D11 units: 1 = IP, 2 = SI                                  */
  HB_IPL,        /* start of growing season, julian day                        */
  HB_IHV,        /* end of growing season                                      */
  HB_IU10,       /* D10 units                                                  */
  HB_IPRE,       /* 0: engine initialises moisture with a spin-up year         */
  HB_NLAY,       /* number of layers (1..20)                                   */
  HB_NODE_P,     /* row of this cell's D4 series in pre[]                      */
  HB_NODE_T,     /* row of this cell's D7 series in tmp[]                      */
  HB_NODE_R,     /* row of this cell's D13 series in rad[]                     */
  HB_NCELL_I32
};
/* ---- per-layer fields: layer_f32[field][layer][n_cells] --------------------
 * D10 layer records (HELP3O.FOR:1286-1300).                                   */
enum {
  HL_THICK = 0, HL_PORO, HL_FC, HL_WP, HL_SW,
  HL_XLENG, HL_SLOPE, HL_RECIR, HL_SUBIN, HL_PHOLE, HL_DEFEC, HL_TRANS,
  HL_NLAYER_F32
};
enum { HL_TYPE = 0, HL_ISOIL, HL_LAYR, HL_IPQ, HL_NLAYER_I32 };

typedef struct help_b200_batch {
  int32_t n_cells;
  int32_t n_years;        /* LMYR: years of output (a spin-up year is added when IPRE=0) */
  int32_t max_layers;     /* layer dimension of the layer_* arrays (1..20)    */
  int32_t n_nodes_p, n_nodes_t, n_nodes_r;
  /* weather, one row per distinct input file ("node"); value [node][year][day],
   * day 0-based, row stride HELP_B200_DAYS floats, as READCD reads them
   * (HELP3O.FOR:1100-1129), i.e. BEFORE the unit conversion of F:1138-1152.
   * pre / tmp / rad may be host pointers or device pointers on the engine's device
   * (help_b200_pack_weather_resident); every other member is a host pointer.   */
  const int32_t* years;   /* [n_years] calendar year of each block (leap = year%4==0, F:1048) */
  const float* pre;       /* [n_nodes_p][n_years][HELP_B200_DAYS]             */
  const float* tmp;       /* [n_nodes_t][n_years][HELP_B200_DAYS]             */
  const float* rad;       /* [n_nodes_r][n_years][HELP_B200_DAYS]             */
  const int32_t* iu4;     /* [n_nodes_p] units flag of each D4 file (1 IP, 2 SI) */
  const int32_t* iu7;     /* [n_nodes_t]                                      */
  const int32_t* iu13;    /* [n_nodes_r]                                      */
  const float* tm_normals;/* [n_nodes_t][12] monthly normal temperatures from the D7 header (F:1216) */
  const int32_t* idfs;    /* [n_nodes_r][2] days-to-thaw pre-scan of the whole D13 file
                             (F:1348-1375) for ULAT>=0 ([0]) and ULAT<0 ([1]) */
  /* cells */
  const float* cell_f32;    /* [HB_NCELL_F32][n_cells]                        */
  const int32_t* cell_i32;  /* [HB_NCELL_I32][n_cells]                        */
  const float* layer_f32;   /* [HL_NLAYER_F32][max_layers][n_cells]           */
  const int32_t* layer_i32; /* [HL_NLAYER_I32][max_layers][n_cells]           */
  const double* layer_rc;   /* [max_layers][n_cells] saturated K, cm/s (REAL*8, F:61) */
} help_b200_batch;

/* Result rows of `monthly` -- the keys of read_monthly_help_output()
 * (pyhelp/processing.py:139-146), mm, quantised to 0.001 mm like the
 * F7.3 text the reference prints and re-parses (HELP3O.FOR:6672-6682).        */
enum { HR_PRECIP = 0, HR_RUNOFF, HR_EVAPO, HR_SUBRUN1, HR_SUBRUN2, HR_PERCO, HR_RECHG, HR_NROWS };
/* Rows of `yearly`: sums over the 12 months, mm. */
enum { HY_PRECIP = 0, HY_RUNOFF, HY_EVAPO, HY_LATDRAIN, HY_RECHG, HY_NROWS };
/* Rows of `raw_monthly` (inches, the engine's REAL*4 accumulators before printing):
 * 0 PREM 1 RUNM 2 ETM 3..7 DRN1M..DRN5M 8..13 PRC1M..PRC6M                    */
#define HELP_B200_RAW_ROWS 14

/* per-cell status word */
#define HELP_B200_FLAG_NONCONV   0x1u  /* Newton solve hit 100 iterations (F:4481)     */
#define HELP_B200_FLAG_NAN       0x2u  /* a NaN reached the monthly output             */
#define HELP_B200_FLAG_OVERFLOW  0x4u  /* a value does not fit F7.3 (reference prints *******) */
#define HELP_B200_FLAG_RECIRC    0x8u  /* informational: the design recirculates drainage into a layer (RECIRC, HELP3O.FOR:5311-5440) */
#define HELP_B200_FLAG_BADCELL   0x10u /* malformed design (no layers, NSEG > capacity ...) */

typedef struct help_b200_result {
  float* monthly;      /* [HR_NROWS][n_years][12][n_cells]  or NULL            */
  float* yearly;       /* [HY_NROWS][n_years][n_cells]      or NULL            */
  float* raw_monthly;  /* [14][n_years][12][n_cells]        or NULL            */
  uint32_t* flags;     /* [n_cells]                         or NULL            */
  int32_t* topo;       /* [16][n_cells] NSEG, NSEG1..6, NSTEP1..6, IDFS, LP1, LD1 -- or NULL */
  /* timings filled by the call (milliseconds, CUDA events on the launch stream) */
  float ms_h2d, ms_setup, ms_simulate, ms_d2h;
  int32_t n_launches;  /* kernels launched by this call                        */
} help_b200_result;

/* ---- one-shot, host buffers in and out (what a binding calls) -------------
 * Replaces: the per-cell loop of run_help_allcells (processing.py:48-49).
 * Copies the batch to the current CUDA device, runs the set-up and simulation
 * kernels, copies the requested result arrays back.  Thread-safe per device.  */
int help_b200_run(const help_b200_batch* batch, help_b200_result* result);

/* ---- resident handle: upload once, simulate many times ---------------------
 * (bench.py's device-resident `value`; also lets a caller keep 5M-cell monthly
 * results on the device and fetch only the yearly summaries).                 */
typedef struct help_b200_engine help_b200_engine;
int help_b200_engine_create(const help_b200_batch* batch, int want_monthly, int want_raw,
                            help_b200_engine** out);
/* stream: a cudaStream_t cast to void* (NULL = the engine's own stream)        */
int help_b200_engine_simulate(help_b200_engine* e, void* stream);
/* device pointers of the result arrays (layouts as in help_b200_result), for
 * zero-copy hand-off to torch / NCCL; NULL when that array was not requested  */
int help_b200_engine_device_ptrs(help_b200_engine* e, void** monthly, void** yearly,
                                 void** raw_monthly, void** flags);
int help_b200_engine_fetch(help_b200_engine* e, help_b200_result* result);

/* The monthly result in the order of the reference's return value -- one block per cell, out[n_cells][HR_NROWS][n_years][12]
 * (pyhelp/processing.py:139-146 returns one dict of [n_years][12] arrays per cell) -- transposed on the device.  Host pointer. */
int help_b200_engine_fetch_monthly_by_cell(help_b200_engine* engine, float* out);
int help_b200_engine_timings(help_b200_engine* e, float* ms_setup, float* ms_simulate,
                             int32_t* n_launches);
void help_b200_engine_destroy(help_b200_engine* e);

/* ---- misc ------------------------------------------------------------------ */
int help_b200_abi_version(void);
int help_b200_device_count(void);
/* select the CUDA device later help_b200_run / engine_create calls of this thread use
 * (one process per GPU: pass LOCAL_RANK).  The reference picks its workers with
 * multiprocessing.Pool(ncore) (pyhelp/processing.py:43-47).                    */
int help_b200_set_device(int device);
const char* help_b200_last_error(void);
/* The library keeps freed device memory in the device's stream-ordered pool so that the next
 * batch re-uses it (the reference's workers likewise stay alive between cells,
 * pyhelp/processing.py:43-47); this hands it back to the driver.                            */
int help_b200_release_memory(void);
/* algorithmic bytes one simulate() moves:
 * 12*W*D + (56+36*L)*N + 336*N*Y, W = distinct nodes referenced, D = days     */
double help_b200_algorithmic_bytes(const help_b200_batch* batch);

/* diagnostic: evaluate on the device the transcendental function `fn` the engine uses
 * (pyhelp_b200/csrc/help_math.h; 0 pow(x,y) 1 exp 2 log 3 sin 4 cos 5 atan, 6..17 the REAL*4
 * wrappers acos tan pow exp log sin cos log10 sqrt x**8 x**7 x**4, 18 the dual-issue pow,
 * 19 the shared-logarithm pow pair, 20 x/y through the rounded reciprocal of y)
 * for n host values.  The
 * reference's `**`, EXP, ALOG ... resolve to libm (HELP3O.FOR:1307-1340, 4329-4480); this lets
 * a test prove the device functions return the same bits as the CPU oracle's.               */
int help_b200_math_eval(int fn, const double* x, const double* y, double* out, int64_t n);

/* ---- adjacent to the path -----------------------------------------
 * Nearest weather-grid node of every cell: replaces the per-cell loop of
 * HelpManager._generate_d4d7d13_input_files (reference pyhelp/managers.py:307-313), i.e.
 * argmin over nodes of calc_dist_from_coord (haversine, r = 6373 km, pyhelp/utils.py:81-96),
 * first minimum wins.  Degrees in, host pointers; out_index[n_cells] receives node indices.  */
int help_b200_nearest_nodes(int64_t n_cells, const double* cell_lat, const double* cell_lon,
                            int32_t n_nodes, const double* node_lat, const double* node_lon,
                            int32_t* out_index);

/* The reference's weather-file round trip for a whole table (8f row 1):
 * save_precip/airtemp/solrad_to_HELP print every daily value with '{:>5.1f}' / '{:>6.1f}' / '{:>6.2f}'
 * (pyhelp/weather_reader.py:31,44,57) and READCD reads the text back as REAL*4
 * (pyhelp/HELP3O.FOR:1100-1129).  table[n_days][n_nodes] float64 (mm, deg C or MJ/m2, host) ->
 * out[n_nodes][n_years][HELP_B200_DAYS] float32 = float32(float(format(x, '.{decimals}f'))), zero
 * beyond a year's days: the layout of help_b200_batch.pre/tmp/rad.  year_day0[y] / year_ndays[y] =
 * first row and number of rows of year y in the table.  Host pointers.                          */
int help_b200_pack_weather(const double* table, int64_t n_days, int32_t n_nodes, const int32_t* year_day0,
                           const int32_t* year_ndays, int32_t n_years, int32_t decimals, float* out);

/* The same, but the packed table stays in device memory: *out_dev receives a device pointer that may be given as
 * help_b200_batch.pre / .tmp / .rad of any number of engines on the same device (help_b200_engine_create copies it device to
 * device and converts its own copy) and is released with help_b200_device_free.  Saves the device -> host -> device round trip
 * of n_nodes x n_years x 368 x 4 bytes per table between packing and simulating.                                              */
int help_b200_pack_weather_resident(const double* table, int64_t n_days, int32_t n_nodes, const int32_t* year_day0,
                                    const int32_t* year_ndays, int32_t n_years, int32_t decimals, float** out_dev);

/* READIN's pre-scan of the radiation file for IDFS (pyhelp/HELP3O.FOR:1348-1375) on a device-resident radiation table
 * rad_dev[n_nodes][n_years][HELP_B200_DAYS] (raw file values: MJ/m2 with iu13 = 2, langleys with 1): out[n_nodes][2] (host) =
 * IDFS for a northern / southern site -- help_b200_batch.idfs.                                                                 */
int help_b200_idfs_prescan_resident(const float* rad_dev, int32_t n_nodes, int32_t n_years, int32_t iu13, int32_t* out);

/* Release a device buffer handed out by this library (help_b200_pack_weather_resident).  NULL is ignored.                      */
void help_b200_device_free(void* p);

/* Fortran formatted READ of fixed-width numeric fields for a whole matrix of records (8f row 1):
 * the reads READCD (pyhelp/HELP3O.FOR:1100-1129, FORMAT(I10,10F5.2) / (I5,10F6.1)) and READIN (HELP3O.FOR:1221-1225 D11,
 * :1260-1300 D10) do per cell, done once for all records of a drop-in run_help_allcells(cellparams) call
 * (pyhelp/processing.py:40-59 receives text).  text[n_records][reclen] bytes (host); field f = columns [col[f], col[f] +
 * width[f]) read as F<width>.<decimals[f]> (I<width> with decimals 0): blanks ignored, all-blank = 0, an explicit decimal
 * point overrides the implied decimals, optional E/D exponent.  out[n_fields][n_records] float64 = the correctly rounded
 * value; flags[n_fields][n_records] = 0, 1 (a valid number that needs more than one exact operation -- more than 15 digits
 * or a large exponent: convert it on the host) or 2 (not a number).                                                      */
int help_b200_parse_fields(const uint8_t* text, int64_t n_records, int32_t reclen, const int32_t* col, const int32_t* width,
                           const int32_t* decimals, int32_t n_fields, double* out, uint8_t* flags);

/* Daily channel: the reference's daily output (IOD = 1: OUTDAY,
 * HELP3O.FOR:5984-6029, parsed by read_daily_help_output, pyhelp/processing.py:150-220) for a
 * few selected cells -- a verification channel, not a bulk output.  select_daily() before
 * engine_simulate(); fetch_daily() afterwards fills out[n_selected][n_days][16] (double, inches,
 * unrounded; n_days = days of the output years, 365/366 by the MOD-4 rule) with the columns
 * precipitation, runoff, evapotranspiration, evaporative-zone water content, then head on the
 * liner / lateral drainage / percolation for subprofiles 1..3, air-temperature freeze flag,
 * soil-freeze flag, snow water.                                                            */
#define HELP_B200_DAILY_COLS 16
int help_b200_engine_select_daily(help_b200_engine* e, const int32_t* cells, int32_t n_selected);
int help_b200_engine_fetch_daily(help_b200_engine* e, double* out, int64_t* n_days);

/* The reductions PyHELP applies to the result, on the engine's device-
 * resident monthly buffer (engine created with want_monthly, after engine_simulate):
 *   HelpManager._post_process_output  (reference pyhelp/managers.py:448-506): a cell whose rechg
 *     holds a NaN is a NaN cell; the recharge of a cell with context == 2 (next to a stream)
 *     becomes subsurface runoff (subrun1 if its subrun2 is all zero, else subrun2);
 *   HelpOutput.calc_area_monthly_avg  (pyhelp/output.py:127-152): nansum over cells / Np;
 *   HelpOutput.calc_cells_yearly_avg  (pyhelp/output.py:171-203): per cell, mean over the years
 *     with year_from <= year <= year_to of the yearly sums.
 * context: int32[n_cells] in the caller's cell order, or NULL (no cell next to a stream).
 * Outputs (host, each may be NULL): area_monthly double[7][n_years][12] (rows HR_*),
 * cells_yearly double[7][n_cells], nan_mask uint8[n_cells] (1 = NaN cell, the reference's
 * idx_nan).                                                                               */
int help_b200_engine_post_process(help_b200_engine* e, const int32_t* context, int32_t year_from,
                                  int32_t year_to, double* area_monthly, double* cells_yearly,
                                  uint8_t* nan_mask);

#ifdef __cplusplus
}
#endif
#endif /* HELP_B200_H */
