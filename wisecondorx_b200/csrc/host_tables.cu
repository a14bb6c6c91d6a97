// Host-side text of the per-bin output table of `predict` (no device work).
//
// Reference code replaced: _generate_bins_bed (predict_output.py:59-84), one Python-formatted line per bin -- 2.06e5 lines
// and 4.1e5 float conversions per sample at 15 kb, the largest item of the `predict` command line after the reference
// file.  The line is  chr \t start \t end \t chr:start-end \t ratio \t zscore  with start = i * binsize + 1,
// end = (i + 1) * binsize, and a value of 0 printed as "nan" (predict_output.py:74-77).  The reference prints the values
// with str() of a NumPy float64, which is Python's repr(float): the SHORTEST decimal string that reads back as the same
// double, in fixed notation for 1e-4 <= |x| < 1e16 (with ".0" appended to integers) and as d.ddde[+-]XX (at least two
// exponent digits) otherwise.  std::to_chars yields the same shortest digits; format_repr lays them out by Python's
// rules.  tests/test_host_pin.py compares millions of values with repr() and the whole table with the Python writer.
#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>

#include "wcx_common.cuh"

namespace {

// writes Python's repr(x) at p, returns the end
char* format_repr(char* p, double x) {
  if (std::isnan(x)) { std::memcpy(p, "nan", 3); return p + 3; }
  if (std::isinf(x)) {
    if (x < 0) *p++ = '-';
    std::memcpy(p, "inf", 3);
    return p + 3;
  }
  if (std::signbit(x)) { *p++ = '-'; x = -x; }
  if (x == 0.0) { std::memcpy(p, "0.0", 3); return p + 3; }
  char buf[48];
  auto res = std::to_chars(buf, buf + sizeof(buf), x, std::chars_format::scientific);  // shortest round-trip digits
  char digits[24];
  int nd = 0, e10 = 0;
  const char* q = buf;
  for (; q < res.ptr && *q != 'e'; ++q)
    if (*q != '.') digits[nd++] = *q;
  if (q < res.ptr) {  // exponent
    ++q;
    bool neg = false;
    if (*q == '-') { neg = true; ++q; } else if (*q == '+') { ++q; }
    for (; q < res.ptr; ++q) e10 = e10 * 10 + (*q - '0');
    if (neg) e10 = -e10;
  }
  if (e10 < -4 || e10 >= 16) {  // float_repr_style "short": exponent form outside [1e-4, 1e16)
    *p++ = digits[0];
    if (nd > 1) {
      *p++ = '.';
      std::memcpy(p, digits + 1, (size_t)nd - 1);
      p += nd - 1;
    }
    *p++ = 'e';
    *p++ = e10 < 0 ? '-' : '+';
    int a = e10 < 0 ? -e10 : e10;
    if (a >= 100) { *p++ = (char)('0' + a / 100); a %= 100; *p++ = (char)('0' + a / 10); *p++ = (char)('0' + a % 10); }
    else { *p++ = (char)('0' + a / 10); *p++ = (char)('0' + a % 10); }
    return p;
  }
  if (e10 >= 0) {
    const int int_digits = e10 + 1;
    for (int i = 0; i < int_digits; ++i) *p++ = i < nd ? digits[i] : '0';
    *p++ = '.';
    if (nd > int_digits) {
      std::memcpy(p, digits + int_digits, (size_t)(nd - int_digits));
      p += nd - int_digits;
    } else {
      *p++ = '0';
    }
    return p;
  }
  *p++ = '0';
  *p++ = '.';
  for (int i = 0; i < -e10 - 1; ++i) *p++ = '0';
  std::memcpy(p, digits, (size_t)nd);
  return p + nd;
}

char* format_int(char* p, int64_t v) {
  auto res = std::to_chars(p, p + 24, v);
  return res.ptr;
}

}  // namespace

extern "C" int wcx_host_format_repr(const double* x, int64_t n, char* out, int64_t cap, int64_t* len) {
  if (n < 0 || !x || !out || !len || cap < n * 26) { wcx::set_error("wcx_host_format_repr: bad argument (26 bytes per value needed)"); return 1; }
  char* p = out;
  for (int64_t i = 0; i < n; ++i) {
    p = format_repr(p, x[i]);
    *p++ = '\n';
  }
  *len = p - out;
  return 0;
}

extern "C" int wcx_host_format_bins(const char* chr_name, int64_t binsize, const double* r, const double* z, int64_t n,
                                    char* out, int64_t cap, int64_t* len) {
  if (n < 0 || binsize < 0 || !chr_name || !out || !len || (n > 0 && (!r || !z))) {
    wcx::set_error("wcx_host_format_bins: bad argument");
    return 1;
  }
  const size_t ln = std::strlen(chr_name);
  if (ln > 16 || cap < n * (int64_t)(2 * ln + 4 * 20 + 2 * 25 + 8)) {
    wcx::set_error("wcx_host_format_bins: output buffer too small");
    return 1;
  }
  char* p = out;
  for (int64_t i = 0; i < n; ++i) {
    const int64_t s = i * binsize + 1, e = (i + 1) * binsize;
    std::memcpy(p, chr_name, ln); p += ln;
    *p++ = '\t'; p = format_int(p, s);
    *p++ = '\t'; p = format_int(p, e);
    *p++ = '\t';
    std::memcpy(p, chr_name, ln); p += ln;
    *p++ = ':'; p = format_int(p, s);
    *p++ = '-'; p = format_int(p, e);
    *p++ = '\t';
    if (r[i] == 0) { std::memcpy(p, "nan", 3); p += 3; } else { p = format_repr(p, r[i]); }  // predict_output.py:74-77
    *p++ = '\t';
    if (z[i] == 0) { std::memcpy(p, "nan", 3); p += 3; } else { p = format_repr(p, z[i]); }
    *p++ = '\n';
  }
  *len = p - out;
  return 0;
}
