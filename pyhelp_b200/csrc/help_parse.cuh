// help_parse.cuh -- Fortran formatted READ of fixed-width numeric fields, on the device.
//
// Everything the reference hands HELP3O is text: 37 records x 10 values per year and weather file (READCD,
// pyhelp/HELP3O.FOR:1100-1129: FORMAT(I10,10F5.2) / (I5,10F6.1)), the D11 record (F:1221-1225) and two records
// per layer of D10 (F:1260-1300).  A drop-in `run_help_allcells(cellparams)` has to read them back: 3.3 x 10^7
// fields for 1000 nodes x 30 yr, which numpy's string->float conversion does at ~7 x 10^6 fields/s on one host
// core -- longer than the simulation of 100 000 cells takes on the GPU.  parse_fields_kernel does the same
// reads for a whole byte matrix of records at once: one thread = one field of one record.
//
// Semantics of an Fw.d / Iw input field (what libgfortran does, and what pyhelp_b200/inputs.py::_to_real did):
// blanks are ignored (BN) and an all-blank field is 0; optional sign; digits with an optional decimal point; an
// optional exponent (E, D, e, d, or a bare sign); without a decimal point the last d digits of the significand are
// decimals.  The value is the correctly rounded double of the decimal: with a significand below 2^53 and a power
// of ten up to 10^22 both operands are exact and one IEEE multiplication or division rounds once.  Anything else
// (more digits, larger exponents, characters that are not part of a number) is flagged and converted on the host.
#pragma once
#include <stdint.h>

namespace helpb200 {

constexpr uint8_t kParseSlow = 1;     // valid number, but not one exact operation: host conversion
constexpr uint8_t kParseBad = 2;      // not a number

__device__ __forceinline__ double pow10_exact(int k) {   // 10^k, 0 <= k <= 22
  const double T[23] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15, 1e16,
                        1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
  return T[k];
}

// text[n_records][reclen]; field f occupies columns [col[f], col[f] + width[f]) and has dec[f] implied decimals;
// out[f][record] receives the value, flag[f][record] 0 / kParseSlow / kParseBad.
__global__ void parse_fields_kernel(const uint8_t* __restrict__ text, int64_t n_records, int reclen, const int32_t* __restrict__ col,
                                    const int32_t* __restrict__ width, const int32_t* __restrict__ dec, int n_fields,
                                    double* __restrict__ out, uint8_t* __restrict__ flag) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_records * n_fields) return;
  const int64_t rec = i / n_fields;
  const int f = (int)(i - rec * n_fields);
  const uint8_t* p = text + rec * reclen + col[f];
  int w = width[f];
  if (col[f] + w > reclen) w = reclen - col[f];
  uint64_t m = 0;
  // state 0: before the significand (a sign is allowed), 1: in the significand, 2: after the exponent letter or its
  // sign (a digit must follow), 3: in the exponent digits
  int ndig = 0, nfrac = 0, e10 = 0, esign = 1, state = 0;
  bool neg = false, point = false, any = false, bad = false, toolong = false, signseen = false, expsign = false;
  for (int k = 0; k < w; ++k) {
    const uint8_t c = p[k];
    if (c == ' ' || c == 0 || c == '\t' || c == '\r') continue;
    any = true;
    if (c >= '0' && c <= '9') {
      if (state <= 1) {
        state = 1;
        if (m > 1844674407370955160ull) toolong = true; else m = m * 10 + (uint64_t)(c - '0');
        ++ndig;
        if (point) ++nfrac;
      } else { state = 3; if (e10 < 100000) e10 = e10 * 10 + (c - '0'); }
    } else if (c == '.') {
      if (state > 1 || point) bad = true;
      point = true; state = 1;
    } else if (c == '+' || c == '-') {
      if (state == 0 && !signseen) { neg = (c == '-'); signseen = true; }
      else if (state == 1 && ndig > 0) { state = 2; expsign = true; esign = (c == '-') ? -1 : 1; }      // "1.5+3": exponent without a letter
      else if (state == 2 && !expsign) { expsign = true; esign = (c == '-') ? -1 : 1; }
      else bad = true;
    } else if (c == 'E' || c == 'e' || c == 'D' || c == 'd') {
      if (state == 1 && ndig > 0) state = 2; else bad = true;
    } else bad = true;
  }
  double v = 0.0;
  uint8_t fl = 0;
  if (any) {
    if (bad || ndig == 0 || state == 2) fl = kParseBad;
    else {
      int e = esign * e10 - nfrac - (point ? 0 : dec[f]);
      if (toolong || m >= (1ull << 53) || e > 22 || e < -22) fl = kParseSlow;
      else {
        const double md = (double)(int64_t)m;
        v = (e >= 0) ? __dmul_rn(md, pow10_exact(e)) : __ddiv_rn(md, pow10_exact(-e));
        if (neg) v = -v;
      }
    }
  }
  out[(int64_t)f * n_records + rec] = v;
  flag[(int64_t)f * n_records + rec] = fl;
}

}  // namespace helpb200
