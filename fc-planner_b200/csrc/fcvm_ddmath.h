// fcvm_ddmath.h — asin / acos in double-double arithmetic, for the device-side pitch / yaw sums of getPose
// (viewpoint_manager.cpp:398-447: one asin and one acos per visible ray, accumulated in point order).
//
// Why this exists.  The reference calls glibc's asin / acos, which are NOT correctly rounded: on this image
// (glibc 2.39) 0.16 % of random arguments differ from the correctly rounded value, but only where the exact
// result lies within 0.021 ulp of a rounding midpoint (measured against libquadmath, tests/test_ddmath.py; the
// IBM Accurate Mathematical Library code behind them evaluates to ~70 bits and rounds once).  So a device
// implementation cannot reproduce glibc bit for bit by itself, but it can compute the CORRECTLY ROUNDED value
// together with a certificate: the double-double result (hi, lo) is accurate to < 2^-12 ulp, and when
// hi + lo is farther than kFlagMargin = 1/32 ulp from every rounding midpoint, glibc's value must equal hi.
// Rays inside the margin (6.25 % of them) are flagged and their arguments go to the host, which calls the
// real glibc function for exactly those (fcvm_host.cu, pose_sums_*).  Nothing here is approximate in effect:
// the summed pitch / yaw are bit-identical to the reference's, the host's libm work drops by ~16x (32x per function).
//
// Algorithm (x in [0, 1]; odd / reflection symmetry outside):  theta ~ asin(x) in plain double picks
// k = rint(64 theta), theta_k = k / 64 (exact), and with s = sqrt(1 - x^2) in double-double
//     d = sin(theta - theta_k) = x cos(theta_k) - s sin(theta_k)           (double-double, |d| <= 2^-6.9)
//     asin(x) = theta_k + asin(d),   asin(d) = d + d^3 (1/6 + 3/40 d^2 + 15/336 d^4 + ...)   (tail in plain double)
//     acos(x) = asin(s) with the roles of x and s swapped (full relative accuracy as x -> 1),  acos(-x) = pi - acos(x)
// sin / cos(theta_k) come from a 102-row double-double table generated with libquadmath
// (tools/gen_dd_tables.c).  Products use an explicit fused multiply-add (exact error term); -fmad=false /
// -ffp-contract=off only stop the COMPILER from fusing, explicit fma calls are unaffected.
#pragma once

#include <cmath>
#include <cstdint>
#include <cstring>

#include "fcvm_ddmath_tables.h"

#if defined(__CUDACC__)
#define FCVM_HD __host__ __device__ __forceinline__
#else
#define FCVM_HD inline
#endif

namespace fcvm {
namespace ddm {

static const double kTableHost[FCVM_DD_TABLE_ROWS * 4] = {FCVM_DD_TABLE_INIT};
#if defined(__CUDACC__)
static __device__ const double kTable[FCVM_DD_TABLE_ROWS * 4] = {FCVM_DD_TABLE_INIT};
#endif
#if defined(__CUDA_ARCH__)
#define FCVM_DD_TBL(i) __ldg(&kTable[i])
#define FCVM_DD_FMA(a, b, c) __fma_rn((a), (b), (c))
#else
#define FCVM_DD_TBL(i) kTableHost[i]
#define FCVM_DD_FMA(a, b, c) __builtin_fma((a), (b), (c))
#endif

struct DD { double hi, lo; };

FCVM_HD DD two_sum(double a, double b) {
  const double s = a + b, bb = s - a;
  return DD{s, (a - (s - bb)) + (b - bb)};
}
FCVM_HD DD fast_two_sum(double a, double b) {   // |a| >= |b|
  const double s = a + b;
  return DD{s, b - (s - a)};
}
FCVM_HD DD two_prod(double a, double b) {
  const double p = a * b;
  return DD{p, FCVM_DD_FMA(a, b, -p)};
}
FCVM_HD DD dd_add(DD a, DD b) {
  DD s = two_sum(a.hi, b.hi);
  const DD t = two_sum(a.lo, b.lo);
  s.lo += t.hi;
  s = fast_two_sum(s.hi, s.lo);
  s.lo += t.lo;
  return fast_two_sum(s.hi, s.lo);
}
FCVM_HD DD dd_neg(DD a) { return DD{-a.hi, -a.lo}; }
FCVM_HD DD dd_mul_d(DD a, double b) {
  DD p = two_prod(a.hi, b);
  p.lo += a.lo * b;
  return fast_two_sum(p.hi, p.lo);
}
FCVM_HD DD dd_mul(DD a, DD b) {
  DD p = two_prod(a.hi, b.hi);
  p.lo += a.hi * b.lo + a.lo * b.hi;
  return fast_two_sum(p.hi, p.lo);
}
// sqrt(1 - x^2) for |x| <= 1, relative error ~2^-100
FCVM_HD DD dd_sqrt_1mx2(double x) {
  const DD p = two_prod(x, x);
  DD t = two_sum(1.0, -p.hi);
  t.lo -= p.lo;
  t = fast_two_sum(t.hi, t.lo);
  if (!(t.hi > 0.0)) return DD{0.0, 0.0};
  const double s0 = sqrt(t.hi);
  const DD q = two_prod(s0, s0);
  const double r = ((t.hi - q.hi) - q.lo) + t.lo;
  return fast_two_sum(s0, r / (2.0 * s0));
}

// in ulps.  glibc 2.39's asin / acos (IBM Accurate Mathematical Library without its former slow paths) evaluate to well below
// 0.03 ulp before their single rounding: the largest distance from a midpoint at which their value was NOT the correctly rounded
// one is 0.0153 ulp (asin) / 0.0217 ulp (acos), and the maximum is reached within the first few million samples and does not move
// over billions (profiles/r2_ddmath_check.json) - a bound of the approximation, not a tail.  1/32 leaves 1.4x on top of it.
constexpr double kFlagMargin = 0.03125;

// true when hi + lo is within kFlagMargin ulp of a rounding midpoint next to hi (conservative at powers of two)
FCVM_HD bool near_midpoint(DD r) {
  const double a = fabs(r.hi);
  if (!(a > 0.0) || !(a < 1.0e300)) return !(r.hi == 0.0 && r.lo == 0.0);
  long long bits;
  memcpy(&bits, &a, sizeof(bits));
  bits -= 1;
  double below;
  memcpy(&below, &bits, sizeof(below));
  const double ulp_small = a - below;        // spacing on the side toward zero (half the spacing above at a power of two)
  return fabs(r.lo) >= (0.5 - kFlagMargin) * ulp_small;
}

// asin(d) - d for the small residual: d^3 q(d^2), plain double (|d| < 2^-6.9: relative 2^-52 of a term below 2^-16 |d|)
FCVM_HD double asin_tail(double d) {
  const double z = d * d;
  double q = 0.01735276442307692;                    // 10395/599040   (d^13)
  q = q * z + 0.022372159090909092;                  // 945/42240      (d^11)
  q = q * z + 0.030381944444444444;                  // 105/3456       (d^9)
  q = q * z + 0.044642857142857144;                  // 15/336         (d^7)
  q = q * z + 0.075;                                 // 3/40           (d^5)
  q = q * z + 0.16666666666666666;                   // 1/6            (d^3)
  return (d * z) * q;
}

// Core: y in [0, 1] and its complement c = sqrt(1 - y^2), both double-double; `theta` = plain-double estimate of asin(y).
// Returns k and d = sin(asin(y) - k/64) = y cos(k/64) - c sin(k/64).
FCVM_HD int reduce(DD y, DD c, double theta, DD& d) {
  int k = (int)rint(theta * 64.0);
  k = k < 0 ? 0 : (k > FCVM_DD_TABLE_ROWS - 1 ? FCVM_DD_TABLE_ROWS - 1 : k);
  const DD sk{FCVM_DD_TBL(4 * k + 0), FCVM_DD_TBL(4 * k + 1)};
  const DD ck{FCVM_DD_TBL(4 * k + 2), FCVM_DD_TBL(4 * k + 3)};
  d = dd_add(dd_mul(y, ck), dd_neg(dd_mul(c, sk)));
  return k;
}
// base + sign * (k/64 + d + asin_tail(d)): every term's absolute error is far below an ulp of the result, because the
// result is >= ~2^-7 whenever k > 0 and equals d (1 + O(d^2)) when k = 0
FCVM_HD DD finish(DD base, int k, DD d, double sign) {
  const double tail = asin_tail(d.hi) * sign;
  const DD t = two_sum(sign * ((double)k * 0.015625), sign * d.hi);
  const DD a = two_sum(base.hi, t.hi);
  const DD b = two_sum(a.hi, tail);
  const double lo = ((a.lo + b.lo) + t.lo) + (base.lo + sign * d.lo);
  return two_sum(b.hi, lo);
}

// Correctly rounded asin(x) with certificate.  *flag = true: the caller must obtain the value from the reference's libm.
// `theta0` = any plain-double approximation of asin(|x|) (a few ulps are plenty; it only selects the table row).
FCVM_HD double asin_cr(double x, double theta0, bool* flag) {
  const double ax = fabs(x);
  if (!(ax < 1.0)) {                      // |x| >= 1 or NaN: leave the exact semantics to libm
    *flag = true;
    return 0.0;
  }
  DD d;
  const int k = reduce(DD{ax, 0.0}, dd_sqrt_1mx2(ax), theta0, d);
  const DD r = finish(DD{0.0, 0.0}, k, d, 1.0);
  *flag = near_midpoint(r);
  return x < 0.0 ? -r.hi : r.hi;
}

// Correctly rounded acos(x) with certificate: acos(|x|) = asin(sqrt(1 - x^2)), whose complement is |x| itself, so the
// result keeps full RELATIVE accuracy as x -> 1; acos(-|x|) = pi - acos(|x|).  `theta0` ~ acos(|x|) in plain double.
FCVM_HD double acos_cr(double x, double theta0, bool* flag) {
  const double ax = fabs(x);
  if (!(ax < 1.0)) {
    *flag = true;
    return 0.0;
  }
  DD d;
  const int k = reduce(dd_sqrt_1mx2(ax), DD{ax, 0.0}, theta0, d);
  const DD r = x >= 0.0 ? finish(DD{0.0, 0.0}, k, d, 1.0) : finish(DD{FCVM_DD_PI_HI, FCVM_DD_PI_LO}, k, d, -1.0);
  *flag = near_midpoint(r);
  return r.hi;
}

}  // namespace ddm
}  // namespace fcvm
