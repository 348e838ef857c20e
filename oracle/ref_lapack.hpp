// ORACLE — TEST INFRASTRUCTURE ONLY (never linked into the product path).
//
// Restatement of the third-party dense kernels the reference's FULL_ALGEBRA
// path links (un-vendored netlib reference BLAS/LAPACK 3.x, see SURVEY.md §8c;
// call sites ADJ/ROS_ADJ/LS_Solver.F90:41 DGETRF, :50,:52 DGETRS;
// ADJ/RK_ADJ/LS_Solver.F90:55 ZGETRF, :104,:106 ZGETRS).  For N < 64 netlib
// DGETRF/ZGETRF run the unblocked DGETF2/ZGETF2 (NB=64 from ILAENV), which is
// what is restated here, loop order and all:
//   DGETF2: IDAMAX first-maximal pivot, DSWAP of full rows, reciprocal scaling
//           when |pivot| >= sfmin, DGER rank-1 update a(i,j) += x(i)*(-y(j)).
//   DGETRS: DLASWP + DTRSM(L,L,N,U) + DTRSM(L,U,N,N)   for 'N'
//           DTRSM(L,U,T,N) + DTRSM(L,L,T,U) + DLASWP^-1 for 'T'
// Matrices are column-major with leading dimension n; pivots are 1-based like
// LAPACK's IPIV.  Compile with -ffp-contract=off (no FMA) to match the
// reference's x86-64 SSE2 build.
#pragma once
#include <cmath>
#include <cfloat>

namespace oracle {

// DLAMCH('E') as used by the integrators (ADJ/ROS_ADJ/..:455): relative machine eps = 2^-53
static inline double dlamch_eps() { return DBL_EPSILON * 0.5; }
// DLAMCH('S') safe minimum (netlib dlamch: sfmin = tiny, bumped if 1/huge >= tiny)
static inline double dlamch_sfmin() {
  double sfmin = DBL_MIN, small = 1.0 / DBL_MAX;
  if (small >= sfmin) sfmin = small * (1.0 + dlamch_eps());
  return sfmin;
}

// Census of row interchanges (column j, pivot row jp != j; n <= 32) over everything factorised so far; read by
// tools/pivot_census.py to choose the candidate interchanges of the generated static-pattern factorisations.
struct PivotCensus { long d[32 * 32], z[32 * 32]; };
inline PivotCensus& pivot_census() { static PivotCensus c = {}; return c; }

// returns info (0 ok, j>0: U(j,j) exactly zero, first such j, 1-based)
static inline int dgetf2(int n, double* a, int* ipiv) {
  const double sfmin = dlamch_sfmin();
  int info = 0;
  for (int j = 0; j < n; ++j) {
    // IDAMAX over a(j:n-1, j): first index of max |.|
    int jp = j;
    double dmax = std::fabs(a[j + n * j]);
    for (int i = j + 1; i < n; ++i) {
      double v = std::fabs(a[i + n * j]);
      if (v > dmax) { dmax = v; jp = i; }
    }
    ipiv[j] = jp + 1;
    if (jp != j && n <= 32) {
#pragma omp atomic
      pivot_census().d[j * 32 + jp]++;
    }
    if (a[jp + n * j] != 0.0) {
      if (jp != j)
        for (int c = 0; c < n; ++c) { double t = a[j + n * c]; a[j + n * c] = a[jp + n * c]; a[jp + n * c] = t; }
      if (j < n - 1) {
        if (std::fabs(a[j + n * j]) >= sfmin) {
          double r = 1.0 / a[j + n * j];
          for (int i = j + 1; i < n; ++i) a[i + n * j] = r * a[i + n * j];   // DSCAL: da*dx(i)
        } else {
          for (int i = j + 1; i < n; ++i) a[i + n * j] = a[i + n * j] / a[j + n * j];
        }
      }
    } else if (info == 0) {
      info = j + 1;
    }
    if (j < n - 1) {
      // DGER(m-j, n-j, -1, x=a(j+1:,j), y=a(j,j+1:), A=a(j+1:,j+1:))
      for (int c = j + 1; c < n; ++c) {
        double yc = a[j + n * c];
        if (yc != 0.0) {
          double temp = -1.0 * yc;
          for (int i = j + 1; i < n; ++i) a[i + n * c] = a[i + n * c] + a[i + n * j] * temp;
        }
      }
    }
  }
  return info;
}

// DGETRS with one right-hand side
static inline void dgetrs(bool trans, int n, const double* a, const int* ipiv, double* b) {
  if (!trans) {
    for (int i = 0; i < n; ++i) {            // DLASWP forward
      int ip = ipiv[i] - 1;
      if (ip != i) { double t = b[i]; b[i] = b[ip]; b[ip] = t; }
    }
    for (int k = 0; k < n; ++k)              // DTRSM Left Lower NoTrans Unit
      if (b[k] != 0.0)
        for (int i = k + 1; i < n; ++i) b[i] = b[i] - b[k] * a[i + n * k];
    for (int k = n - 1; k >= 0; --k)         // DTRSM Left Upper NoTrans NonUnit
      if (b[k] != 0.0) {
        b[k] = b[k] / a[k + n * k];
        for (int i = 0; i < k; ++i) b[i] = b[i] - b[k] * a[i + n * k];
      }
  } else {
    for (int i = 0; i < n; ++i) {            // DTRSM Left Upper Trans NonUnit
      double temp = b[i];
      for (int k = 0; k < i; ++k) temp = temp - a[k + n * i] * b[k];
      temp = temp / a[i + n * i];
      b[i] = temp;
    }
    for (int i = n - 1; i >= 0; --i) {       // DTRSM Left Lower Trans Unit
      double temp = b[i];
      for (int k = i + 1; k < n; ++k) temp = temp - a[k + n * i] * b[k];
      b[i] = temp;
    }
    for (int i = n - 1; i >= 0; --i) {       // DLASWP backward
      int ip = ipiv[i] - 1;
      if (ip != i) { double t = b[i]; b[i] = b[ip]; b[ip] = t; }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// COMPLEX*16 kernels of the implicit Runge-Kutta families (FWD/RK/LS_Solver.F90:21-32 ZGETRF, :40-52 ZGETRS 'N').
// netlib ZGETF2: IZAMAX pivot on DCABS1 = |re|+|im| (first maximum), ZSWAP of full rows, ZSCAL by the reciprocal
// ONE/A(j,j) when ABS(A(j,j)) >= sfmin, ZGERU rank-1 update a(i,c) += x(i)*(-ONE*y(c)).
// Complex products are the plain four-multiplication form, quotients use the range-reduced (Smith) form gfortran
// emits inline under its default -fcx-fortran-rules; no FMA.
struct cplx { double re, im; };
static inline cplx cmul(cplx a, cplx b) { return cplx{a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
static inline cplx cadd(cplx a, cplx b) { return cplx{a.re + b.re, a.im + b.im}; }
static inline cplx csub(cplx a, cplx b) { return cplx{a.re - b.re, a.im - b.im}; }
static inline cplx cdiv(cplx n, cplx d) {   // (a + ib)/(c + id), GCC expand_complex_div_wide
  const double a = n.re, b = n.im, c = d.re, e = d.im;
  if (std::fabs(c) < std::fabs(e)) {
    double ratio = c / e, div = c * ratio + e;
    return cplx{(a * ratio + b) / div, (b * ratio - a) / div};
  } else {
    double ratio = e / c, div = e * ratio + c;
    return cplx{(b * ratio + a) / div, (b - a * ratio) / div};
  }
}
static inline bool cnonzero(cplx a) { return a.re != 0.0 || a.im != 0.0; }

static inline int zgetf2(int n, cplx* a, int* ipiv) {
  const double sfmin = dlamch_sfmin();
  int info = 0;
  for (int j = 0; j < n; ++j) {
    int jp = j;
    double dmax = std::fabs(a[j + n * j].re) + std::fabs(a[j + n * j].im);
    for (int i = j + 1; i < n; ++i) {
      double v = std::fabs(a[i + n * j].re) + std::fabs(a[i + n * j].im);
      if (v > dmax) { dmax = v; jp = i; }
    }
    ipiv[j] = jp + 1;
    if (jp != j && n <= 32) {
#pragma omp atomic
      pivot_census().z[j * 32 + jp]++;
    }
    if (cnonzero(a[jp + n * j])) {
      if (jp != j)
        for (int c = 0; c < n; ++c) { cplx t = a[j + n * c]; a[j + n * c] = a[jp + n * c]; a[jp + n * c] = t; }
      if (j < n - 1) {
        if (std::hypot(a[j + n * j].re, a[j + n * j].im) >= sfmin) {
          cplx r = cdiv(cplx{1.0, 0.0}, a[j + n * j]);
          for (int i = j + 1; i < n; ++i) a[i + n * j] = cmul(r, a[i + n * j]);   // ZSCAL: za*zx(i)
        } else {
          for (int i = j + 1; i < n; ++i) a[i + n * j] = cdiv(a[i + n * j], a[j + n * j]);
        }
      }
    } else if (info == 0) {
      info = j + 1;
    }
    if (j < n - 1) {
      for (int c = j + 1; c < n; ++c) {
        cplx yc = a[j + n * c];
        if (cnonzero(yc)) {
          cplx temp = cmul(cplx{-1.0, 0.0}, yc);
          for (int i = j + 1; i < n; ++i) a[i + n * c] = cadd(a[i + n * c], cmul(a[i + n * j], temp));
        }
      }
    }
  }
  return info;
}

// ZGETRS('N') with one right-hand side: ZLASWP + ZTRSM(L,L,N,U) + ZTRSM(L,U,N,N)
static inline void zgetrs_n(int n, const cplx* a, const int* ipiv, cplx* b) {
  for (int i = 0; i < n; ++i) {
    int ip = ipiv[i] - 1;
    if (ip != i) { cplx t = b[i]; b[i] = b[ip]; b[ip] = t; }
  }
  for (int k = 0; k < n; ++k)
    if (cnonzero(b[k]))
      for (int i = k + 1; i < n; ++i) b[i] = csub(b[i], cmul(b[k], a[i + n * k]));
  for (int k = n - 1; k >= 0; --k)
    if (cnonzero(b[k])) {
      b[k] = cdiv(b[k], a[k + n * k]);
      for (int i = 0; i < k; ++i) b[i] = csub(b[i], cmul(b[k], a[i + n * k]));
    }
}

// ZGETRS('T') with one right-hand side: ZTRSM(L,U,T,N) + ZTRSM(L,L,T,U) + ZLASWP(-1); plain transpose, no conjugation
static inline void zgetrs_t(int n, const cplx* a, const int* ipiv, cplx* b) {
  const cplx one = {1.0, 0.0};
  for (int i = 0; i < n; ++i) {
    cplx temp = cmul(one, b[i]);
    for (int k = 0; k < i; ++k) temp = csub(temp, cmul(a[k + n * i], b[k]));
    temp = cdiv(temp, a[i + n * i]);
    b[i] = temp;
  }
  for (int i = n - 1; i >= 0; --i) {
    cplx temp = cmul(one, b[i]);
    for (int k = i + 1; k < n; ++k) temp = csub(temp, cmul(a[k + n * i], b[k]));
    b[i] = temp;
  }
  for (int i = n - 1; i >= 0; --i) {
    int ip = ipiv[i] - 1;
    if (ip != i) { cplx t = b[i]; b[i] = b[ip]; b[ip] = t; }
  }
}

}  // namespace oracle
