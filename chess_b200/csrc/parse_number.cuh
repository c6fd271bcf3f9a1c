// Decimal text -> int / double for the sparse-matrix parser (ingest.cu), usable from host and
// device code.  The target is bit-identity with Python's int(str) / float(str), which the reference
// applies to every field (chess/helpers.py:316-326).  The conversion is correctly rounded or it
// gives up: anything it cannot decide with certainty is reported as PARSE_FALLBACK and the line is
// re-parsed on the host by Python itself.
//
//   1. w * 10^q with w < 2^53 and |q| <= 22: one IEEE multiplication or division of two exactly
//      representable numbers (Clinger's fast path) -- correctly rounded by construction.
//   2. otherwise w (<= 19 significant digits, normalised to 64 bits) times a 128-bit truncated
//      power of five (pow5_table.inc, generated with exact integer arithmetic).  The true product
//      lies in [Z, Z + 2) units of the 128-bit result; the rounding is taken only if no value in
//      that interval could change it (for 0 <= q <= 55 the table entry is exact, so is the product,
//      and ties are resolved to even).  Subnormal results are left to the host.
#pragma once
#include <stdint.h>

#include "pow5_table.inc"

#ifdef __CUDACC__
#define CHESS_HD __host__ __device__ __forceinline__
#else
#define CHESS_HD inline
#endif

enum { PARSE_OK = 0, PARSE_NONFINITE = 2, PARSE_FALLBACK = 3 };

CHESS_HD bool chess_is_space(char c) {  // str.rstrip() / float() whitespace (ASCII subset)
  return c == ' ' || c == '\t' || c == '\r' || c == '\n' || c == '\v' || c == '\f';
}

// [+]digits, at most 9 digits (regions are int32 bins); anything else -> fallback
CHESS_HD int chess_parse_index(const char *s, int len, int32_t *out) {
  int i = 0;
  if (len > 0 && s[0] == '+') i = 1;
  if (len - i < 1 || len - i > 9) return PARSE_FALLBACK;
  int32_t v = 0;
  for (; i < len; ++i) {
    const unsigned d = (unsigned)(s[i] - '0');
    if (d > 9u) return PARSE_FALLBACK;
    v = v * 10 + (int32_t)d;
  }
  *out = v;
  return PARSE_OK;
}

CHESS_HD void chess_mul64(uint64_t a, uint64_t b, uint64_t *hi, uint64_t *lo) {
#ifdef __CUDA_ARCH__
  *lo = a * b;
  *hi = __umul64hi(a, b);
#else
  const unsigned __int128 p = (unsigned __int128)a * b;
  *lo = (uint64_t)p;
  *hi = (uint64_t)(p >> 64);
#endif
}

CHESS_HD int chess_clz64(uint64_t v) {
#ifdef __CUDA_ARCH__
  return __clzll((long long)v);
#else
  return __builtin_clzll(v);
#endif
}

CHESS_HD double chess_bits_to_double(uint64_t b) {
#ifdef __CUDA_ARCH__
  return __longlong_as_double((long long)b);
#else
  double d;
  __builtin_memcpy(&d, &b, 8);
  return d;
#endif
}

// w * 10^q (w != 0) -> nearest double, or PARSE_FALLBACK / PARSE_NONFINITE (overflow to inf)
CHESS_HD int chess_decimal_to_double(uint64_t w, int q, const uint64_t *pow5, const short *exp5,
                                     const double *pow10, double *out) {
  if (w <= (1ull << 53) && q >= -22 && q <= 22) {
    const double m = (double)(long long)w;  // exact
    *out = q < 0 ? m / pow10[-q] : m * pow10[q];
    return PARSE_OK;
  }
  if (q < CHESS_POW5_QMIN) return PARSE_FALLBACK;      // underflow region: host decides
  if (q > CHESS_POW5_QMAX) return PARSE_NONFINITE;     // >= 1e309 -> inf
  const int lz = chess_clz64(w);
  const uint64_t wn = w << lz;
  const uint64_t th = pow5[2 * (q - CHESS_POW5_QMIN)], tl = pow5[2 * (q - CHESS_POW5_QMIN) + 1];
  // P = wn * T (192 bits) = upper:lower:lowest
  uint64_t h1, l1, h2, l2;
  chess_mul64(wn, th, &h1, &l1);
  chess_mul64(wn, tl, &h2, &l2);
  uint64_t upper = h1, lower = l1 + h2;
  if (lower < l1) upper += 1;
  const uint64_t lowest = l2;
  const int upperbit = (int)(upper >> 63);
  const int shift = 10 + upperbit;                      // keep 53 bits
  uint64_t m53 = upper >> shift;
  const uint64_t round_bit = (upper >> (shift - 1)) & 1ull;
  const uint64_t rest_mask = (1ull << (shift - 1)) - 1ull;
  const uint64_t rest_hi = upper & rest_mask;
  const bool exact = q >= 0 && q <= 55;                 // 5^q < 2^128: the table entry is exact
  bool sticky;
  if (exact) {
    sticky = (rest_hi | lower | lowest) != 0ull;
  } else {
    // the true product lies in [Z, Z + 2) units of `lower`
    if (rest_hi == rest_mask && lower >= 0xfffffffffffffffeull) return PARSE_FALLBACK;  // may carry
    if (round_bit && rest_hi == 0ull && lower == 0ull) return PARSE_FALLBACK;           // tie or just above
    sticky = true;  // an inexact entry means the product is not an exact multiple: something lies below
    if (rest_hi == 0ull && lower == 0ull) sticky = false;  // round bit is 0 here: rounds down either way
  }
  if (round_bit && (sticky || (m53 & 1ull))) m53 += 1;
  int e = 63 + upperbit - lz + (int)exp5[q - CHESS_POW5_QMIN] + q;  // value = 1.xxx * 2^e
  if (m53 == (1ull << 53)) { m53 >>= 1; e += 1; }
  const int biased = e + 1023;
  if (biased >= 2047) return PARSE_NONFINITE;
  if (biased <= 0) return PARSE_FALLBACK;               // subnormal: left to the host
  *out = chess_bits_to_double(((uint64_t)biased << 52) | (m53 & ((1ull << 52) - 1ull)));
  return PARSE_OK;
}

// float(str) for the plain grammar  [+-] (digits [. digits] | . digits) [(e|E) [+-] digits]  and
// the spellings of inf / nan; surrounding whitespace must already be stripped.
CHESS_HD int chess_parse_double(const char *s, int len, const uint64_t *pow5, const short *exp5,
                                const double *pow10, double *out) {
  int i = 0;
  bool neg = false;
  if (len > 0 && (s[0] == '+' || s[0] == '-')) { neg = s[0] == '-'; i = 1; }
  if (i >= len) return PARSE_FALLBACK;
  {  // inf / infinity / nan, any case
    const int r = len - i;
    auto low = [&](int k) { const char c = s[i + k]; return (char)((c >= 'A' && c <= 'Z') ? c + 32 : c); };
    if (r == 3 && low(0) == 'n' && low(1) == 'a' && low(2) == 'n') return PARSE_NONFINITE;
    if (r == 3 && low(0) == 'i' && low(1) == 'n' && low(2) == 'f') return PARSE_NONFINITE;
    if (r == 8 && low(0) == 'i' && low(1) == 'n' && low(2) == 'f' && low(3) == 'i' && low(4) == 'n' &&
        low(5) == 'i' && low(6) == 't' && low(7) == 'y')
      return PARSE_NONFINITE;
  }
  uint64_t w = 0;
  int ndig = 0;        // significant digits taken into w (after leading zeros)
  int q = 0;           // decimal exponent adjustment
  bool any_digit = false, seen_dot = false, dropped_nonzero = false;
  for (; i < len; ++i) {
    const char c = s[i];
    if (c == '.') {
      if (seen_dot) return PARSE_FALLBACK;
      seen_dot = true;
      continue;
    }
    const unsigned d = (unsigned)(c - '0');
    if (d > 9u) break;
    any_digit = true;
    if (ndig == 0 && d == 0u) {        // leading zero
      if (seen_dot) q -= 1;
      continue;
    }
    if (ndig < 19) {
      w = w * 10ull + d;
      ndig += 1;
      if (seen_dot) q -= 1;
    } else {                           // beyond 19 significant digits
      if (d != 0u) dropped_nonzero = true;
      if (!seen_dot) q += 1;
    }
  }
  if (!any_digit) return PARSE_FALLBACK;
  if (i < len) {
    if (s[i] != 'e' && s[i] != 'E') return PARSE_FALLBACK;
    ++i;
    bool eneg = false;
    if (i < len && (s[i] == '+' || s[i] == '-')) { eneg = s[i] == '-'; ++i; }
    if (i >= len) return PARSE_FALLBACK;
    int ev = 0;
    for (; i < len; ++i) {
      const unsigned d = (unsigned)(s[i] - '0');
      if (d > 9u) return PARSE_FALLBACK;
      if (ev < 100000) ev = ev * 10 + (int)d;
    }
    q += eneg ? -ev : ev;
  }
  if (dropped_nonzero) return PARSE_FALLBACK;  // more than 19 significant digits: host
  if (w == 0ull) {
    *out = neg ? -0.0 : 0.0;
    return PARSE_OK;
  }
  double v;
  const int rc = chess_decimal_to_double(w, q, pow5, exp5, pow10, &v);
  if (rc != PARSE_OK) return rc;
  *out = neg ? -v : v;
  return PARSE_OK;
}
