// ddmath_check.cpp — CPU check of fc-planner_b200/csrc/fcvm_ddmath.h against libquadmath (accuracy of the double-double
// result) and against this box's glibc asin / acos (the certificate: outside the flag margin the correctly rounded value
// must equal glibc's).  Prints one JSON line.   usage: ddmath_check <n> <seed>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <quadmath.h>

#include "../fc-planner_b200/csrc/fcvm_ddmath.h"

using namespace fcvm::ddm;

static uint64_t s = 88172645463325252ull;
static inline uint64_t rnd() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }
static inline double u01() { return (double)(rnd() >> 11) / 9007199254740992.0; }

int main(int argc, char** argv) {
  const long n = argc > 1 ? atol(argv[1]) : 1000000;
  if (argc > 2) s ^= (uint64_t)atoll(argv[2]) * 0x9E3779B97F4A7C15ull;
  long flagged[2] = {0, 0}, bad_unflagged[2] = {0, 0}, glibc_not_cr[2] = {0, 0}, flagged_mismatch[2] = {0, 0};
  double max_err_ulp[2] = {0, 0}, worst_glibc_margin[2] = {0, 0};
  for (long i = 0; i < n; ++i) {
    double x;
    switch (i & 7) {
      case 0: x = 1.0 - ldexp(u01(), -(int)(rnd() % 50)); break;              // towards 1
      case 1: x = ldexp(u01(), -(int)(rnd() % 60)); break;                    // towards 0
      case 2: x = (double)(rnd() % 103) / 64.0 * 0.6 + ldexp(u01() - 0.5, -20); x = sin(x); break;   // near table nodes
      default: x = u01(); break;
    }
    if (x > 1.0) x = 1.0;
    if (rnd() & 1) x = -x;
    for (int w = 0; w < 2; ++w) {
      // the plain-double estimate the table index comes from: perturbed, it must not matter
      double th = w ? acos(fabs(x)) : asin(fabs(x));
      if (i & 8) th = w ? (double)acosf((float)fabs(x)) : (double)asinf((float)fabs(x));
      bool flag = false;
      const double v = w ? acos_cr(x, th, &flag) : asin_cr(x, th, &flag);
      const double g = w ? acos(x) : asin(x);
      if (fabs(x) >= 1.0) {            // always flagged: libm's own value is used
        if (!flag) bad_unflagged[w]++;
        continue;
      }
      const __float128 q = w ? acosq((__float128)x) : asinq((__float128)x);
      const double cr = (double)q;
      if (g != cr) {
        glibc_not_cr[w]++;
        const __float128 mid = ((__float128)g + (__float128)cr) / 2;
        const double m = (double)(fabsq(q - mid) / fabsq((__float128)g - (__float128)cr));
        if (m > worst_glibc_margin[w]) worst_glibc_margin[w] = m;
      }
      // accuracy of the double-double value behind v: recompute it here
      {
        DD d, r;
        const double ax = fabs(x);
        if (!w) { const int k = reduce(DD{ax, 0.0}, dd_sqrt_1mx2(ax), th, d); r = finish(DD{0.0, 0.0}, k, d, 1.0); }
        else { const int k = reduce(dd_sqrt_1mx2(ax), DD{ax, 0.0}, th, d); r = x >= 0.0 ? finish(DD{0.0, 0.0}, k, d, 1.0) : finish(DD{FCVM_DD_PI_HI, FCVM_DD_PI_LO}, k, d, -1.0); }
        const __float128 mine = (__float128)r.hi + (__float128)r.lo;
        const __float128 ref = (!w && x < 0) ? -q : q;
        double ulp = fabs(cr) > 0 ? fabs(nextafter(fabs(cr), INFINITY) - fabs(cr)) : 4.9e-324;
        const double e = (double)(fabsq(mine - ref) / ulp);
        if (e > max_err_ulp[w]) max_err_ulp[w] = e;
      }
      if (flag) {
        flagged[w]++;
        if (g != v) flagged_mismatch[w]++;
      } else if (g != v) {
        bad_unflagged[w]++;
        if (bad_unflagged[w] < 5) fprintf(stderr, "UNFLAGGED MISMATCH %s(%a): mine %a glibc %a cr %a\n", w ? "acos" : "asin", x, v, g, cr);
      }
    }
  }
  printf("{\"n\": %ld, \"flagged_asin\": %ld, \"flagged_acos\": %ld, \"unflagged_mismatch_asin\": %ld, \"unflagged_mismatch_acos\": %ld, "
         "\"glibc_not_cr_asin\": %ld, \"glibc_not_cr_acos\": %ld, \"flagged_mismatch_asin\": %ld, \"flagged_mismatch_acos\": %ld, "
         "\"max_err_ulp_asin\": %.3g, \"max_err_ulp_acos\": %.3g, \"worst_glibc_margin_asin\": %.4g, \"worst_glibc_margin_acos\": %.4g}\n",
         n, flagged[0], flagged[1], bad_unflagged[0], bad_unflagged[1], glibc_not_cr[0], glibc_not_cr[1], flagged_mismatch[0], flagged_mismatch[1],
         max_err_ulp[0], max_err_ulp[1], worst_glibc_margin[0], worst_glibc_margin[1]);
  return (bad_unflagged[0] || bad_unflagged[1]) ? 1 : 0;
}
