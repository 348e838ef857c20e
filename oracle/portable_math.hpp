// ORACLE — TEST INFRASTRUCTURE ONLY.
// Deterministic COS (Update_SUN) and Err**(1/ELO) (step controller) for the oracle.  The Fortran
// standard does not fix the bits of COS or `**`; the reference binary inherits them from whatever
// libm it links, and glibc itself returns different last bits on different CPUs (FMA vs non-FMA
// variants).  Because the CBM-IV error estimate turns one ulp into a different accept/reject
// sequence, the oracle pins these two functions to fixed-order IEEE kernels (< 1 ulp, fdlibm-style),
// so its ISTATUS is the same on every host.  ORACLE_LIBM=1 in the environment selects libm's
// cos/pow instead (tests check both modes against the reference's golden files).
// Same algorithm and constants as the product's fatode_b200/csrc/portable_math.h; kept as a
// separate copy so nothing under oracle/ is used by the product (tests compare the two bitwise).
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

// sin on [-pi/4, pi/4], argument x + y (y = tail)
FPM_HD double k_sin(double x, double y, int iy) {
  const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03, S3 = -1.98412698298579493134e-04,
               S4 = 2.75573137070700676789e-06, S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
  double z = x * x;
  double v = z * x;
  double r = S2 + z * (S3 + z * (S4 + z * (S5 + z * S6)));
  if (iy == 0) return x + v * (S1 + z * r);
  return x - ((z * (0.5 * y - v * r) - y) - v * S1);
}
// cos on [-pi/4, pi/4], argument x + y
FPM_HD double k_cos(double x, double y) {
  const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03, C3 = 2.48015872894767294178e-05,
               C4 = -2.75573143513906633035e-07, C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
  double ax = fabs(x);
  double z = x * x;
  double r = z * (C1 + z * (C2 + z * (C3 + z * (C4 + z * (C5 + z * C6)))));
  if (ax < 0.3) return 1.0 - (0.5 * z - (z * r - x * y));
  double qx;
  if (ax > 0.78125) qx = 0.28125;
  else qx = from_bits((bits_of(ax) - 0x0020000000000000ull) & 0xffffffff00000000ull);   // ~|x|/4, 32-bit mantissa
  double hz = 0.5 * z - qx;
  double a = 1.0 - qx;
  return a - (hz - (z * r - x * y));
}
// cos(x) for |x| <= ~3.93 (5*pi/4): quadrant reduction with a two-part pi/2
FPM_HD double cos_small(double x) {
  const double pio4 = 7.85398163397448278999e-01;
  const double pio2_1 = 1.57079632673412561417e+00;    // first 33 bits of pi/2
  const double pio2_1t = 6.07710050650619224932e-11;   // pi/2 - pio2_1
  double ax = fabs(x);
  if (ax <= pio4) return k_cos(ax, 0.0);
  int n = (ax <= 3.0 * pio4) ? 1 : 2;
  double z = ax - (double)n * pio2_1;                  // exact (33-bit constant, small n)
  double t = (double)n * pio2_1t;
  double r0 = z - t;
  double r1 = (z - r0) - t;                            // tail
  if (n == 1) return -k_sin(r0, r1, 1);
  return -k_cos(r0, r1);
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

// x**(1/elo) for the estimator orders used by the Rosenbrock tableaux (ros_ELO = 2, 3 or 4)
FPM_HD double root_elo(double x, double elo) {
  if (elo == 4.0) return sqrt(sqrt(x));
  if (elo == 3.0) {
    // the reference raises to fl(1/3) = 1/3 + d, d = -1.85e-17: x**fl(1/3) = cbrt(x)*(1 + d*ln x);
    // ln x is only needed to two digits here (exponent + linear mantissa term)
    const double c = cbrt_pos(x);
    const uint64_t b = bits_of(x);
    const double lg2 = (double)((int)((b >> 52) & 0x7ff) - 1023) + (from_bits((b & 0x000fffffffffffffull) | 0x3ff0000000000000ull) - 1.0);
    return c + c * (-1.850371707708594e-17 * (0.6931471805599453 * lg2));
  }
  if (elo == 2.0) return sqrt(x);
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
