// Small example models as per-thread device code for k_ros_fwd_cell (one system per thread), with netlib
// DGETF2 / DGETRS('N') on a dense N x N matrix held in the thread's shared-memory slots.
//
//   small_strato  EXAMPLES/small_strato/small_ros/User_Parameters.F90: FUN_U :10-37, JAC_U :39-85,
//                 Update_SUN :87-112, Update_RCONST :114-127; driver small_strato_dr.F90:50-58 (N = 5, NFIX = 2)
//   roberts       EXAMPLES/roberts/ROBERTS_ROS_ADJ/roberts_ros_adj_dr.F90: fun :97-110, jac :75-94 (N = 3)
// REAL(4) literals of the Fortran sources keep their single-precision rounding ((double)x.f);
// operation order is the source's, no FMA contraction in this translation unit.
//
// For N <= 5 full partial pivoting costs nothing: the whole DGETF2 (IDAMAX first-maximal pivot, DSWAP of full rows,
// reciprocal scaling when |pivot| >= DLAMCH('S'), DGER rank-1 update skipping zero y(j)) is unrolled per thread, the
// pivot vector is packed 4 bits per column in `swaps`.  A zero pivot returns LU_SINGULAR and the integrator's
// ros_PrepareMatrix retry (halve H, at most five times) runs inside the same kernel — these models have no other path.
#pragma once
#include <cfloat>
#include "fatode_common.cuh"
#include "model_cbm4.cuh"   // Update_SUN (same routine in all KPP examples)

namespace fatode {

// DGETF2 on a column-major N x N matrix, element (i,j) at a[(i + N*j) * SV]
template <int N, int SV>
__device__ __forceinline__ int dense_lu_factor(double* a, double* dinv, unsigned& swaps) {
  const double sfmin = DBL_MIN;
  unsigned piv = 0u;
  int info = 0;
#pragma unroll
  for (int j = 0; j < N; ++j) {
    int jp = j;
    double dmax = fabs(a[(j + N * j) * SV]);
#pragma unroll
    for (int i = j + 1; i < N; ++i) {
      const double v = fabs(a[(i + N * j) * SV]);
      if (v > dmax) { dmax = v; jp = i; }
    }
    piv |= (unsigned)jp << (4 * j);
    if (a[(jp + N * j) * SV] != 0.0) {
      if (jp != j) {
#pragma unroll
        for (int c = 0; c < N; ++c) {
          const double t = a[(j + N * c) * SV];
          a[(j + N * c) * SV] = a[(jp + N * c) * SV];
          a[(jp + N * c) * SV] = t;
        }
      }
      const double d = a[(j + N * j) * SV];
      dinv[j * SV] = __ddiv_rn(1.0, d);
      if (j < N - 1) {
        if (fabs(d) >= sfmin) {
          const double r = __ddiv_rn(1.0, d);
#pragma unroll
          for (int i = j + 1; i < N; ++i) a[(i + N * j) * SV] = __dmul_rn(r, a[(i + N * j) * SV]);
        } else {
#pragma unroll
          for (int i = j + 1; i < N; ++i) a[(i + N * j) * SV] = __ddiv_rn(a[(i + N * j) * SV], d);
        }
      }
    } else if (info == 0) {
      info = j + 1;
    }
    if (j < N - 1) {
#pragma unroll
      for (int c = j + 1; c < N; ++c) {
        const double yc = a[(j + N * c) * SV];
        if (yc != 0.0) {
          const double temp = __dmul_rn(-1.0, yc);
#pragma unroll
          for (int i = j + 1; i < N; ++i)
            a[(i + N * c) * SV] = __dadd_rn(a[(i + N * c) * SV], __dmul_rn(a[(i + N * j) * SV], temp));
        }
      }
    }
  }
  swaps = piv;
  return info == 0 ? LU_OK : LU_SINGULAR;
}

// DGETRS('N'): DLASWP, DTRSM(L,L,N,U), DTRSM(L,U,N,N)
template <int N, int SV>
__device__ __forceinline__ void dense_lu_solve(const double* a, unsigned swaps, double* b) {
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const int ip = (swaps >> (4 * i)) & 15;
    if (ip != i) { const double t = b[i * SV]; b[i * SV] = b[ip * SV]; b[ip * SV] = t; }
  }
#pragma unroll
  for (int k = 0; k < N; ++k) {
    const double bk = b[k * SV];
    if (bk != 0.0) {
#pragma unroll
      for (int i = k + 1; i < N; ++i) b[i * SV] = __dsub_rn(b[i * SV], __dmul_rn(bk, a[(i + N * k) * SV]));
    }
  }
#pragma unroll
  for (int k = N - 1; k >= 0; --k) {
    if (b[k * SV] != 0.0) {
      const double bk = __ddiv_rn(b[k * SV], a[(k + N * k) * SV]);
      b[k * SV] = bk;
#pragma unroll
      for (int i = 0; i < k; ++i) b[i * SV] = __dsub_rn(b[i * SV], __dmul_rn(bk, a[(i + N * k) * SV]));
    }
  }
}

// ---- COMPLEX*16 counterparts (FWD/RK/LS_Solver.F90:22-52): ZGETF2 / ZGETRS('N') on two planes ar/ai, netlib loop order,
// four-multiplication products and range-reduced (Smith) quotients as gfortran emits them (oracle/ref_lapack.hpp cdiv)
__device__ __forceinline__ void zdiv_dev(double a, double b, double c, double e, double& qr, double& qi) {
  if (fabs(c) < fabs(e)) {
    const double ratio = __ddiv_rn(c, e), div = __dadd_rn(__dmul_rn(c, ratio), e);
    qr = __ddiv_rn(__dadd_rn(__dmul_rn(a, ratio), b), div);
    qi = __ddiv_rn(__dsub_rn(__dmul_rn(b, ratio), a), div);
  } else {
    const double ratio = __ddiv_rn(e, c), div = __dadd_rn(__dmul_rn(e, ratio), c);
    qr = __ddiv_rn(__dadd_rn(__dmul_rn(b, ratio), a), div);
    qi = __ddiv_rn(__dsub_rn(b, __dmul_rn(a, ratio)), div);
  }
}
template <int N, int SV>
__device__ __forceinline__ int dense_zlu_factor(double* ar, double* ai, unsigned& swaps) {
  unsigned piv = 0u;
  int info = 0;
#pragma unroll
  for (int j = 0; j < N; ++j) {
    int jp = j;
    double dmax = __dadd_rn(fabs(ar[(j + N * j) * SV]), fabs(ai[(j + N * j) * SV]));
#pragma unroll
    for (int i = j + 1; i < N; ++i) {
      const double v = __dadd_rn(fabs(ar[(i + N * j) * SV]), fabs(ai[(i + N * j) * SV]));
      if (v > dmax) { dmax = v; jp = i; }
    }
    piv |= (unsigned)jp << (4 * j);
    if (ar[(jp + N * j) * SV] != 0.0 || ai[(jp + N * j) * SV] != 0.0) {
      if (jp != j) {
#pragma unroll
        for (int c = 0; c < N; ++c) {
          const double tr = ar[(j + N * c) * SV], ti = ai[(j + N * c) * SV];
          ar[(j + N * c) * SV] = ar[(jp + N * c) * SV]; ai[(j + N * c) * SV] = ai[(jp + N * c) * SV];
          ar[(jp + N * c) * SV] = tr; ai[(jp + N * c) * SV] = ti;
        }
      }
      if (j < N - 1) {
        const double pr = ar[(j + N * j) * SV], pi = ai[(j + N * j) * SV];
        if (hypot(pr, pi) >= DBL_MIN) {
          double rr, ri;
          zdiv_dev(1.0, 0.0, pr, pi, rr, ri);
#pragma unroll
          for (int i = j + 1; i < N; ++i) {
            const double xr = ar[(i + N * j) * SV], xi = ai[(i + N * j) * SV];
            ar[(i + N * j) * SV] = __dsub_rn(__dmul_rn(rr, xr), __dmul_rn(ri, xi));
            ai[(i + N * j) * SV] = __dadd_rn(__dmul_rn(rr, xi), __dmul_rn(ri, xr));
          }
        } else {
#pragma unroll
          for (int i = j + 1; i < N; ++i) {
            double qr, qi;
            zdiv_dev(ar[(i + N * j) * SV], ai[(i + N * j) * SV], pr, pi, qr, qi);
            ar[(i + N * j) * SV] = qr; ai[(i + N * j) * SV] = qi;
          }
        }
      }
    } else if (info == 0) {
      info = j + 1;
    }
    if (j < N - 1) {
#pragma unroll
      for (int c = j + 1; c < N; ++c) {
        const double yr = ar[(j + N * c) * SV], yi = ai[(j + N * c) * SV];
        if (yr != 0.0 || yi != 0.0) {
          const double tr = __dsub_rn(__dmul_rn(-1.0, yr), __dmul_rn(0.0, yi)), ti = __dadd_rn(__dmul_rn(-1.0, yi), __dmul_rn(0.0, yr));
#pragma unroll
          for (int i = j + 1; i < N; ++i) {
            const double xr = ar[(i + N * j) * SV], xi = ai[(i + N * j) * SV];
            ar[(i + N * c) * SV] = __dadd_rn(ar[(i + N * c) * SV], __dsub_rn(__dmul_rn(xr, tr), __dmul_rn(xi, ti)));
            ai[(i + N * c) * SV] = __dadd_rn(ai[(i + N * c) * SV], __dadd_rn(__dmul_rn(xr, ti), __dmul_rn(xi, tr)));
          }
        }
      }
    }
  }
  swaps = piv;
  return info == 0 ? LU_OK : LU_SINGULAR;
}
template <int N, int SV>
__device__ __forceinline__ void dense_zlu_solve(const double* ar, const double* ai, unsigned swaps, double* br, double* bi) {
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const int ip = (swaps >> (4 * i)) & 15;
    if (ip != i) {
      const double tr = br[i * SV], ti = bi[i * SV];
      br[i * SV] = br[ip * SV]; bi[i * SV] = bi[ip * SV];
      br[ip * SV] = tr; bi[ip * SV] = ti;
    }
  }
#pragma unroll
  for (int k = 0; k < N; ++k) {
    const double kr = br[k * SV], ki = bi[k * SV];
    if (kr != 0.0 || ki != 0.0) {
#pragma unroll
      for (int i = k + 1; i < N; ++i) {
        const double xr = ar[(i + N * k) * SV], xi = ai[(i + N * k) * SV];
        br[i * SV] = __dsub_rn(br[i * SV], __dsub_rn(__dmul_rn(kr, xr), __dmul_rn(ki, xi)));
        bi[i * SV] = __dsub_rn(bi[i * SV], __dadd_rn(__dmul_rn(kr, xi), __dmul_rn(ki, xr)));
      }
    }
  }
#pragma unroll
  for (int k = N - 1; k >= 0; --k) {
    if (br[k * SV] != 0.0 || bi[k * SV] != 0.0) {
      double kr, ki;
      zdiv_dev(br[k * SV], bi[k * SV], ar[(k + N * k) * SV], ai[(k + N * k) * SV], kr, ki);
      br[k * SV] = kr; bi[k * SV] = ki;
#pragma unroll
      for (int i = 0; i < k; ++i) {
        const double xr = ar[(i + N * k) * SV], xi = ai[(i + N * k) * SV];
        br[i * SV] = __dsub_rn(br[i * SV], __dsub_rn(__dmul_rn(kr, xr), __dmul_rn(ki, xi)));
        bi[i * SV] = __dsub_rn(bi[i * SV], __dadd_rn(__dmul_rn(kr, xi), __dmul_rn(ki, xr)));
      }
    }
  }
}
// dense policies: complex matrix from build_E (real plane, shift a) + imaginary plane (b on the diagonal)
#define FATODE_DENSE_COMPLEX_POLICY()                                                                                             \
  template <int SV>                                                                                                               \
  __device__ static void build_EZ(const double* y, const double* PH, const Const& mc, double a, double b, double* zr, double* zi) { \
    build_E<SV>(y, PH, mc, a, zr);                                                                                                \
    _Pragma("unroll") for (int j = 0; j < N; ++j)                                                                                 \
      _Pragma("unroll") for (int i = 0; i < N; ++i) zi[(i + N * j) * SV] = (i == j) ? b : 0.0;                                    \
  }                                                                                                                               \
  template <int SV>                                                                                                               \
  __device__ static int zlu_factor(double* zr, double* zi, unsigned& swaps) { return dense_zlu_factor<N, SV>(zr, zi, swaps); }    \
  template <int SV>                                                                                                               \
  __device__ static void zlu_solve(const double* zr, const double* zi, unsigned swaps, double* br, double* bi) {                  \
    dense_zlu_solve<N, SV>(zr, zi, swaps, br, bi);                                                                                \
  }

#define FATODE_F4(x) ((double)(x##f))

struct SmallConst {
  double fix[2];   // small_strato: FIX(1), FIX(2); roberts: unused
};

// ---------------------------------------------------------------------------------------------------------------
struct SmallStratoCell {
  static constexpr int N = 5, NP = 0, NLU = 25, NPH = 4;
  static constexpr bool HAS_GENERAL_PATH = false;
  typedef SmallConst Const;
  // Update_RCONST (:114-127): PH = RCONST(1), (3), (5), (10), the SUN-dependent ones
  __device__ static void photo(double t, const Const&, double* PH) {
    const double SUN = Cbm4Warp::sun(t);
    PH[0] = __dmul_rn(__dmul_rn(__dmul_rn(FATODE_F4(2.643E-10), SUN), SUN), SUN);
    PH[1] = __dmul_rn(FATODE_F4(6.120E-04), SUN);
    PH[2] = __dmul_rn(__dmul_rn(FATODE_F4(1.070E-03), SUN), SUN);
    PH[3] = __dmul_rn(FATODE_F4(1.289E-02), SUN);
  }
  template <int SV>
  __device__ static void fun(const double* V, const double* PH, const Const& mc, double* F) {   // FUN_U :10-37
    const double R2 = FATODE_F4(8.018E-17), R4 = FATODE_F4(1.576E-15), R6 = FATODE_F4(7.110E-11), R7 = FATODE_F4(1.200E-10),
                 R8 = FATODE_F4(6.062E-15), R9 = FATODE_F4(1.069E-11);
    const double v1 = V[0], v2 = V[1 * SV], v3 = V[2 * SV], v4 = V[3 * SV], v5 = V[4 * SV];
    const double A1 = __dmul_rn(PH[0], mc.fix[1]);
    const double A2 = __dmul_rn(__dmul_rn(R2, v2), mc.fix[1]);
    const double A3 = __dmul_rn(PH[1], v3);
    const double A4 = __dmul_rn(__dmul_rn(R4, v2), v3);
    const double A5 = __dmul_rn(PH[2], v3);
    const double A6 = __dmul_rn(__dmul_rn(R6, v1), mc.fix[0]);
    const double A7 = __dmul_rn(__dmul_rn(R7, v1), v3);
    const double A8 = __dmul_rn(__dmul_rn(R8, v3), v4);
    const double A9 = __dmul_rn(__dmul_rn(R9, v2), v5);
    const double A10 = __dmul_rn(PH[3], v5);
    F[0] = __dsub_rn(__dsub_rn(A5, A6), A7);
    F[1 * SV] = __dadd_rn(__dsub_rn(__dadd_rn(__dsub_rn(__dadd_rn(__dsub_rn(__dmul_rn(2.0, A1), A2), A3), A4), A6), A9), A10);
    F[2 * SV] = __dsub_rn(__dsub_rn(__dsub_rn(__dsub_rn(__dsub_rn(A2, A3), A4), A5), A7), A8);
    F[3 * SV] = __dadd_rn(__dadd_rn(-A8, A9), A10);
    F[4 * SV] = __dsub_rn(__dsub_rn(A8, A9), A10);
  }
  // JAC_U (:39-85, zeroes JF first) fused with lss_decomp's  e = -fjac; e(j,j) += 1/(h*gamma)  (FWD/ROS/LS_Solver.F90)
  template <int SV>
  __device__ static void build_E(const double* V, const double* PH, const Const& mc, double ghinv, double* E) {
    const double R2 = FATODE_F4(8.018E-17), R4 = FATODE_F4(1.576E-15), R6 = FATODE_F4(7.110E-11), R7 = FATODE_F4(1.200E-10),
                 R8 = FATODE_F4(6.062E-15), R9 = FATODE_F4(1.069E-11);
    const double v1 = V[0], v2 = V[1 * SV], v3 = V[2 * SV], v4 = V[3 * SV], v5 = V[4 * SV];
    const double B3 = __dmul_rn(R2, mc.fix[1]);
    const double B5 = PH[1];
    const double B6 = __dmul_rn(R4, v3);
    const double B7 = __dmul_rn(R4, v2);
    const double B8 = PH[2];
    const double B9 = __dmul_rn(R6, mc.fix[0]);
    const double B11 = __dmul_rn(R7, v3);
    const double B12 = __dmul_rn(R7, v1);
    const double B13 = __dmul_rn(R8, v4);
    const double B14 = __dmul_rn(R8, v3);
    const double B15 = __dmul_rn(R9, v5);
    const double B16 = __dmul_rn(R9, v2);
    const double B17 = PH[3];
    double J[25];
#pragma unroll
    for (int i = 0; i < 25; ++i) J[i] = 0.0;
#define JFX(i, j) J[(i - 1) + 5 * (j - 1)]
    JFX(1, 1) = __dsub_rn(-B9, B11);
    JFX(1, 3) = __dsub_rn(B8, B12);
    JFX(2, 1) = B9;
    JFX(2, 2) = __dsub_rn(__dsub_rn(-B3, B6), B15);
    JFX(2, 3) = __dsub_rn(B5, B7);
    JFX(2, 5) = __dadd_rn(-B16, B17);
    JFX(3, 1) = -B11;
    JFX(3, 2) = __dsub_rn(B3, B6);
    JFX(3, 3) = __dsub_rn(__dsub_rn(__dsub_rn(__dsub_rn(-B5, B7), B8), B12), B13);
    JFX(3, 4) = -B14;
    JFX(3, 5) = 0.0;
    JFX(4, 2) = B15;
    JFX(4, 3) = -B13;
    JFX(4, 4) = -B14;
    JFX(4, 5) = __dadd_rn(B16, B17);
    JFX(5, 2) = -B15;
    JFX(5, 3) = B13;
    JFX(5, 4) = B14;
    JFX(5, 5) = __dsub_rn(-B16, B17);
#undef JFX
#pragma unroll
    for (int j = 0; j < 5; ++j)
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        const double e = -J[i + 5 * j];
        E[(i + 5 * j) * SV] = (i == j) ? __dadd_rn(e, ghinv) : e;
      }
  }
  template <int SV>
  __device__ static int lu_factor(double* lu, double* dinv, unsigned& swaps) { return dense_lu_factor<N, SV>(lu, dinv, swaps); }
  template <int SV>
  __device__ static void lu_solve(const double* lu, unsigned swaps, double* b) { dense_lu_solve<N, SV>(lu, swaps, b); }
  FATODE_DENSE_COMPLEX_POLICY()
};

// ---------------------------------------------------------------------------------------------------------------
struct RobertsCell {
  static constexpr int N = 3, NP = 3, NLU = 9, NPH = 1;
  static constexpr bool HAS_GENERAL_PATH = false;
  typedef SmallConst Const;
  __device__ static void photo(double, const Const&, double* PH) { PH[0] = 0.0; }   // autonomous right-hand side
  template <int SV>
  __device__ static void fun(const double* y, const double*, const Const&, double* p) {   // :97-110
    const double c1 = FATODE_F4(0.04e0), c2 = FATODE_F4(1.0e4), c3 = FATODE_F4(3.0e7);
    const double y1 = y[0], y2 = y[1 * SV], y3 = y[2 * SV];
    const double p1 = __dadd_rn(__dmul_rn(-c1, y1), __dmul_rn(__dmul_rn(c2, y2), y3));
    const double p3 = __dmul_rn(__dmul_rn(c3, y2), y2);
    p[0] = p1;
    p[2 * SV] = p3;
    p[1 * SV] = __dsub_rn(-p1, p3);
  }
  // jac :75-94 does not write fjac(3,1), fjac(3,3): they keep the zero of the freshly allocated matrix
  template <int SV>
  __device__ static void build_E(const double* y, const double*, const Const&, double ghinv, double* E) {
    const double c1 = FATODE_F4(0.04e0), c2 = FATODE_F4(1.0e4), c3 = FATODE_F4(3.0e7);
    const double y2 = y[1 * SV], y3 = y[2 * SV];
    double J[9];
#define FJ(i, j) J[(i - 1) + 3 * (j - 1)]
    FJ(3, 1) = 0.0; FJ(3, 3) = 0.0;
    FJ(1, 1) = -c1;
    FJ(1, 2) = __dmul_rn(c2, y3);
    FJ(1, 3) = __dmul_rn(c2, y2);
    FJ(2, 1) = c1;
    FJ(2, 2) = __dsub_rn(__dmul_rn(-c2, y3), __dmul_rn(__dmul_rn(2.0, c3), y2));
    FJ(2, 3) = __dmul_rn(-c2, y2);
    FJ(3, 2) = __dmul_rn(__dmul_rn(2.0, c3), y2);
#undef FJ
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const double e = -J[i + 3 * j];
        E[(i + 3 * j) * SV] = (i == j) ? __dadd_rn(e, ghinv) : e;
      }
  }
  template <int SV>
  __device__ static int lu_factor(double* lu, double* dinv, unsigned& swaps) { return dense_lu_factor<N, SV>(lu, dinv, swaps); }
  template <int SV>
  __device__ static void lu_solve(const double* lu, unsigned swaps, double* b) { dense_lu_solve<N, SV>(lu, swaps, b); }
  FATODE_DENSE_COMPLEX_POLICY()
};

}  // namespace fatode
