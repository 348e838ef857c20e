// ORACLE — TEST INFRASTRUCTURE ONLY.
// Deterministic, correctly rounded COS (Update_SUN) and Err**(1/ELO) (step controller) for the oracle.  The Fortran
// standard does not fix the bits of COS or `**`; the reference binary inherits them from whatever libm it links, and because
// the CBM-IV error estimate turns one ulp into a different accept/reject sequence the step counts depend on the libm
// (DESIGN.md section 5).  The oracle pins the two functions to the correctly rounded values (double-double evaluation, one
// rounding) -- the only implementation-independent definition, and what the IBM Accurate Mathematical Library of the golden
// binaries' glibc returned -- so its ISTATUS is the same on every host.  ORACLE_LIBM=1 in the environment selects the host
// libm's cos/pow instead (tests check both modes against the reference's golden files).
// Same algorithm and constants as the product's fatode_b200/csrc/portable_math.h; kept as a separate copy so nothing under
// oracle/ is used by the product (tests/test_portable_math.py compares the two bitwise on the GPU box).
#pragma once
#include <stdint.h>
#include <string.h>
#include <math.h>

#if defined(__CUDACC__)
#define FPM_HD __host__ __device__ __forceinline__
#else
#define FPM_HD static inline
#endif

namespace oracle_pm {

FPM_HD uint64_t bits_of(double x) { uint64_t u; memcpy(&u, &x, 8); return u; }
FPM_HD double from_bits(uint64_t u) { double x; memcpy(&x, &u, 8); return x; }

// ---- double-double helpers (error-free transformations; every operation is an IEEE operation, fma included)
struct dd { double hi, lo; };
FPM_HD dd two_sum(double a, double b) { double s = a + b; double bb = s - a; double e = (a - (s - bb)) + (b - bb); dd r = {s, e}; return r; }
FPM_HD dd fast_two_sum(double a, double b) { double s = a + b; double e = b - (s - a); dd r = {s, e}; return r; }   // |a| >= |b|
FPM_HD dd two_prod(double a, double b) { double p = a * b; double e = fma(a, b, -p); dd r = {p, e}; return r; }
FPM_HD dd dd_add(dd a, dd b) {
  dd s = two_sum(a.hi, b.hi), t = two_sum(a.lo, b.lo);
  s.lo = s.lo + t.hi;
  s = fast_two_sum(s.hi, s.lo);
  s.lo = s.lo + t.lo;
  return fast_two_sum(s.hi, s.lo);
}
FPM_HD dd dd_add_d(dd a, double b) {
  dd s = two_sum(a.hi, b);
  s.lo = s.lo + a.lo;
  return fast_two_sum(s.hi, s.lo);
}
FPM_HD dd dd_mul(dd a, dd b) {
  dd p = two_prod(a.hi, b.hi);
  p.lo = p.lo + (a.hi * b.lo + a.lo * b.hi);
  return fast_two_sum(p.hi, p.lo);
}

// cos(r) for the double-double r, |r| <= pi/4, z = r*r: Taylor series, the terms below 1e-9 in double, the rest in double-double
FPM_HD dd k_cos_dd(dd z) {
  const double zh = z.hi;
  const double tail = zh * (2.08767569878681e-09 + zh * (-1.1470745597729725e-11 + zh * (4.779477332387385e-14 + zh * (-1.5619206968586225e-16 +
                      zh * (4.110317623312165e-19 + zh * (-8.896791392450574e-22 + zh * 1.6117375710961184e-24))))));
  const dd D1 = {-0.5, 0.0}, D2 = {0.041666666666666664, 2.3129646346357427e-18}, D3 = {-0.001388888888888889, 5.300543954373577e-20},
           D4 = {2.48015873015873e-05, 2.1511947866775882e-23}, D5 = {-2.755731922398589e-07, -2.3767714622250297e-23};
  dd a = dd_add_d(D5, tail);
  a = dd_add(D4, dd_mul(z, a));
  a = dd_add(D3, dd_mul(z, a));
  a = dd_add(D2, dd_mul(z, a));
  a = dd_add(D1, dd_mul(z, a));
  a = dd_mul(z, a);
  return dd_add_d(a, 1.0);
}
// sin(r) for the double-double r, |r| <= pi/4, z = r*r
FPM_HD dd k_sin_dd(dd r, dd z) {
  const double zh = z.hi;
  const double tail = zh * (1.6059043836821613e-10 + zh * (-7.647163731819816e-13 + zh * (2.8114572543455206e-15 + zh * (-8.22063524662433e-18 +
                      zh * (1.9572941063391263e-20 + zh * -3.868170170630684e-23)))));
  const dd E1 = {-0.16666666666666666, -9.25185853854297e-18}, E2 = {0.008333333333333333, 1.1564823173178714e-19},
           E3 = {-0.0001984126984126984, -1.7209558293420705e-22}, E4 = {2.7557319223985893e-06, -1.858393274046472e-22},
           E5 = {-2.505210838544172e-08, 1.448814070935912e-24};
  dd a = dd_add_d(E5, tail);
  a = dd_add(E4, dd_mul(z, a));
  a = dd_add(E3, dd_mul(z, a));
  a = dd_add(E2, dd_mul(z, a));
  a = dd_add(E1, dd_mul(z, a));
  a = dd_mul(dd_mul(z, a), r);      // r * z * (...)
  return dd_add(r, a);
}
// cos(x), correctly rounded (see the header comment), for |x| <= 5*pi/4 (Update_SUN needs |x| <= pi): quadrant reduction with a
// three-part pi/2 (33 + 33 + 53 bits, the first two products with n exact), the kernels in double-double, one final rounding
FPM_HD double cos_cr(double x) {
  const double pio4 = 7.85398163397448278999e-01;
  const double P1 = 1.5707963267341256, P2 = 6.077100506303966e-11, P3 = 2.0222662487959506e-21;
  const double ax = fabs(x);
  const int n = (ax <= pio4) ? 0 : ((ax <= 3.0 * pio4) ? 1 : 2);
  dd r = {ax, 0.0};
  if (n) {
    const double z = ax - (double)n * P1;              // exact (33-bit constant, Sterbenz)
    r = two_sum(z, -((double)n * P2));                 // exact
    r.lo = r.lo - (double)n * P3;
    r = fast_two_sum(r.hi, r.lo);
  }
  const dd z = dd_mul(r, r);
  if (n == 1) return -k_sin_dd(r, z).hi;
  const double c = k_cos_dd(z).hi;
  return n ? -c : c;
}

// cube root, x > 0 normal (fdlibm s_cbrt construction: rational seed, 20-bit chop, one Newton step)
FPM_HD double cbrt_pos(double x) {
  const double C = 5.42857142857142815906e-01, D = -7.05306122448979611050e-01, E = 1.41428571428571436819e+00,
               F = 1.60714285714285720630e+00, G = 3.57142857142857150787e-01;
  uint32_t hx = (uint32_t)(bits_of(x) >> 32);
  double t = from_bits((uint64_t)(hx / 3u + 715094163u) << 32);
  double r = t * t / x;
  double s = C + r * t;
  t = t * (G + F / (s + E + D / s));
  t = from_bits(((bits_of(t) >> 32) + 1ull) << 32);    // chop to 20 bits, round up
  s = t * t;                                           // exact
  r = x / s;
  double w = t + t;
  r = (r - t) / (w + r);
  return t + t * r;
}

// x**fl(1/elo), correctly rounded, for the estimator orders of the Rosenbrock tableaux (ros_ELO = 2, 3 or 4; the reference
// evaluates Err**(ONE/ros_ELO)): a root to < 1 ulp, then one Newton correction from the exact residual (fma), one rounding
FPM_HD double root_elo(double x, double elo) {
  if (elo == 2.0) return sqrt(x);
  if (!(x > 1.0e-290 && x < 1.0e290)) return (elo == 4.0) ? sqrt(sqrt(x)) : pow(x, 1.0 / elo);   // never met: Err >= 1e-10
  if (elo == 4.0) {
    const double r0 = sqrt(sqrt(x));
    const dd p = two_prod(r0, r0);                      // r0^2 = p.hi + p.lo
    const dd q = two_prod(p.hi, p.hi);                  // p.hi^2 = q.hi + q.lo
    const double e = (x - q.hi) - (q.lo + 2.0 * p.hi * p.lo);   // x - r0^4 (x - q.hi is exact)
    return r0 + e * r0 / (4.0 * x);                     // Newton: delta = e / (4 r0^3)
  }
  if (elo == 3.0) {
    // x**fl(1/3) = cbrt(x) * exp(d ln x), d = fl(1/3) - 1/3 = -1.85e-17
    const double c0 = cbrt_pos(x);
    const dd p = two_prod(c0, c0);
    const dd q = two_prod(p.hi, c0);
    const double e = (x - q.hi) - (q.lo + p.lo * c0);   // x - c0^3
    const double d1 = e / (3.0 * p.hi);
    const uint64_t b = bits_of(x);
    int ex = (int)((b >> 52) & 0x7ff) - 1023;
    double m = from_bits((b & 0x000fffffffffffffull) | 0x3ff0000000000000ull);
    if (m > 1.4142135623730951) { m = 0.5 * m; ex += 1; }
    const double s = (m - 1.0) / (m + 1.0), s2 = s * s;   // ln m = 2 atanh(s), |s| <= 0.172: relative error < 1e-12
    const double lnm = 2.0 * s * (1.0 + s2 * (1.0 / 3.0 + s2 * (0.2 + s2 * (1.0 / 7.0 + s2 * (1.0 / 9.0 + s2 * (1.0 / 11.0 + s2 * (1.0 / 13.0
                       + s2 * (1.0 / 15.0))))))));
    const double lnx = (double)ex * 0.6931471805599453 + lnm;
    return c0 + (d1 + c0 * (-1.850371707708594e-17 * lnx));
  }
  return pow(x, 1.0 / elo);   // not reached by the shipped tableaux
}


// x**(-1/k) for a small positive integer-valued k: the SDIRK step controller's Err**(-1/rkELO) (rkELO = 2, 4) and the
// Newton-failure factor Qnewton**(-1/k), k = 1 + NewtonMaxit - NewtonIter, Qnewton in [1,10]
// (FWD/SDIRK/SDIRK_f90_Integrator.F90:548, 587).  Fixed-order IEEE operations only (host == device bits): roots by
// sqrt / the cbrt kernel where k allows, otherwise Newton's iteration on y**(-k) = x from the Bernoulli seed
// 1/(1+(x-1)/k) (monotone from below until the correction reaches round-off level; a few ulp).
FPM_HD double pow_neg_inv(double x, double k) {
  if (k == 1.0) return 1.0 / x;
  if (k == 2.0) return 1.0 / sqrt(x);
  if (k == 4.0) return 1.0 / sqrt(sqrt(x));
  if (k == 8.0) return 1.0 / sqrt(sqrt(sqrt(x)));
  if (k == 3.0) return 1.0 / root_elo(x, 3.0);
  if (k == 6.0) return 1.0 / sqrt(root_elo(x, 3.0));
  const int n = (int)k;
  double y = 1.0 / (1.0 + (x - 1.0) / k);
  for (int it = 0; it < 400; ++it) {
    double p = y, q = 1.0;          // y**n by square-and-multiply
    for (unsigned m = (unsigned)n; m; m >>= 1) { if (m & 1u) q = q * p; p = p * p; }
    const double d = (1.0 - x * q) / k;   // relative Newton correction: positive and decreasing until round-off level
    if (!(d > 4.0e-16)) break;
    y = y * (1.0 + d);
  }
  return y;
}

}  // namespace oracle_pm
