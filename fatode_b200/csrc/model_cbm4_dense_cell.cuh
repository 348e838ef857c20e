// CBM-IV with GENERAL partial pivoting, one system per thread: the fallback policy of the implicit families
// (k_sdirk_fwd_cell / k_rk_fwd_cell).  Systems whose factorisations leave the static pivot pattern of Cbm4Cell
// (IERR_IRREGULAR from the first pass) are integrated again from their initial state with this policy: netlib DGETF2 /
// DGETRS('N') and ZGETF2 / ZGETRS('N') on dense column-major 32 x 32 matrices in the thread's shared-memory slots
// (element (i,j) at a[(i + 32 j) * SV]; the pivot row of column j is kept as a double at a[(1024 + j) * SV]), loop order,
// exact-zero skips and the reciprocal/sfmin branch as in oracle/ref_lapack.hpp.  Singular matrices return LU_SINGULAR and
// take the integrator's own retry path (halve H), exactly like the reference.  Slow (≈ 30 KB of shared memory per
// system) but only a few systems per thousand need it, and only at loose tolerances.
#pragma once
#include <cfloat>
#include "ros_cell_fwd.cuh"
#include "model_small_cell.cuh"

namespace fatode {

template <int SV>
__device__ __noinline__ int dense32_lu_factor(double* a) {
  constexpr int N = 32;
  int info = 0;
#pragma unroll 1
  for (int j = 0; j < N; ++j) {
    int jp = j;
    double dmax = fabs(a[(j + N * j) * SV]);
#pragma unroll 1
    for (int i = j + 1; i < N; ++i) {
      const double v = fabs(a[(i + N * j) * SV]);
      if (v > dmax) { dmax = v; jp = i; }
    }
    a[(N * N + j) * SV] = (double)jp;
    if (a[(jp + N * j) * SV] != 0.0) {
      if (jp != j) {
#pragma unroll 4
        for (int c = 0; c < N; ++c) {
          const double t = a[(j + N * c) * SV];
          a[(j + N * c) * SV] = a[(jp + N * c) * SV];
          a[(jp + N * c) * SV] = t;
        }
      }
      if (j < N - 1) {
        const double d = a[(j + N * j) * SV];
        if (fabs(d) >= DBL_MIN) {
          const double r = __ddiv_rn(1.0, d);
#pragma unroll 1
          for (int i = j + 1; i < N; ++i) a[(i + N * j) * SV] = __dmul_rn(r, a[(i + N * j) * SV]);
        } else {
#pragma unroll 1
          for (int i = j + 1; i < N; ++i) a[(i + N * j) * SV] = __ddiv_rn(a[(i + N * j) * SV], d);
        }
      }
    } else if (info == 0) {
      info = j + 1;
    }
    if (j < N - 1) {
#pragma unroll 1
      for (int c = j + 1; c < N; ++c) {
        const double yc = a[(j + N * c) * SV];
        if (yc != 0.0) {
          const double temp = __dmul_rn(-1.0, yc);
#pragma unroll 1
          for (int i = j + 1; i < N; ++i)
            a[(i + N * c) * SV] = __dadd_rn(a[(i + N * c) * SV], __dmul_rn(a[(i + N * j) * SV], temp));
        }
      }
    }
  }
  return info == 0 ? LU_OK : LU_SINGULAR;
}

template <int SV>
__device__ __noinline__ void dense32_lu_solve(const double* a, double* b) {
  constexpr int N = 32;
#pragma unroll 1
  for (int i = 0; i < N; ++i) {
    const int ip = (int)a[(N * N + i) * SV];
    if (ip != i) { const double t = b[i * SV]; b[i * SV] = b[ip * SV]; b[ip * SV] = t; }
  }
#pragma unroll 1
  for (int k = 0; k < N; ++k) {
    const double bk = b[k * SV];
    if (bk != 0.0) {
#pragma unroll 1
      for (int i = k + 1; i < N; ++i) b[i * SV] = __dsub_rn(b[i * SV], __dmul_rn(bk, a[(i + N * k) * SV]));
    }
  }
#pragma unroll 1
  for (int k = N - 1; k >= 0; --k) {
    if (b[k * SV] != 0.0) {
      const double bk = __ddiv_rn(b[k * SV], a[(k + N * k) * SV]);
      b[k * SV] = bk;
#pragma unroll 1
      for (int i = 0; i < k; ++i) b[i * SV] = __dsub_rn(b[i * SV], __dmul_rn(bk, a[(i + N * k) * SV]));
    }
  }
}

template <int SV>
__device__ __noinline__ int dense32_zlu_factor(double* ar, double* ai) {
  constexpr int N = 32;
  int info = 0;
#pragma unroll 1
  for (int j = 0; j < N; ++j) {
    int jp = j;
    double dmax = __dadd_rn(fabs(ar[(j + N * j) * SV]), fabs(ai[(j + N * j) * SV]));
#pragma unroll 1
    for (int i = j + 1; i < N; ++i) {
      const double v = __dadd_rn(fabs(ar[(i + N * j) * SV]), fabs(ai[(i + N * j) * SV]));
      if (v > dmax) { dmax = v; jp = i; }
    }
    ar[(N * N + j) * SV] = (double)jp;
    if (ar[(jp + N * j) * SV] != 0.0 || ai[(jp + N * j) * SV] != 0.0) {
      if (jp != j) {
#pragma unroll 2
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
#pragma unroll 1
          for (int i = j + 1; i < N; ++i) {
            const double xr = ar[(i + N * j) * SV], xi = ai[(i + N * j) * SV];
            ar[(i + N * j) * SV] = __dsub_rn(__dmul_rn(rr, xr), __dmul_rn(ri, xi));
            ai[(i + N * j) * SV] = __dadd_rn(__dmul_rn(rr, xi), __dmul_rn(ri, xr));
          }
        } else {
#pragma unroll 1
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
#pragma unroll 1
      for (int c = j + 1; c < N; ++c) {
        const double yr = ar[(j + N * c) * SV], yi = ai[(j + N * c) * SV];
        if (yr != 0.0 || yi != 0.0) {
          const double tr = __dsub_rn(__dmul_rn(-1.0, yr), __dmul_rn(0.0, yi)), ti = __dadd_rn(__dmul_rn(-1.0, yi), __dmul_rn(0.0, yr));
#pragma unroll 1
          for (int i = j + 1; i < N; ++i) {
            const double xr = ar[(i + N * j) * SV], xi = ai[(i + N * j) * SV];
            ar[(i + N * c) * SV] = __dadd_rn(ar[(i + N * c) * SV], __dsub_rn(__dmul_rn(xr, tr), __dmul_rn(xi, ti)));
            ai[(i + N * c) * SV] = __dadd_rn(ai[(i + N * c) * SV], __dadd_rn(__dmul_rn(xr, ti), __dmul_rn(xi, tr)));
          }
        }
      }
    }
  }
  return info == 0 ? LU_OK : LU_SINGULAR;
}

template <int SV>
__device__ __noinline__ void dense32_zlu_solve(const double* ar, const double* ai, double* br, double* bi) {
  constexpr int N = 32;
#pragma unroll 1
  for (int i = 0; i < N; ++i) {
    const int ip = (int)ar[(N * N + i) * SV];
    if (ip != i) {
      const double tr = br[i * SV], ti = bi[i * SV];
      br[i * SV] = br[ip * SV]; bi[i * SV] = bi[ip * SV];
      br[ip * SV] = tr; bi[ip * SV] = ti;
    }
  }
#pragma unroll 1
  for (int k = 0; k < N; ++k) {
    const double kr = br[k * SV], ki = bi[k * SV];
    if (kr != 0.0 || ki != 0.0) {
#pragma unroll 1
      for (int i = k + 1; i < N; ++i) {
        const double xr = ar[(i + N * k) * SV], xi = ai[(i + N * k) * SV];
        br[i * SV] = __dsub_rn(br[i * SV], __dsub_rn(__dmul_rn(kr, xr), __dmul_rn(ki, xi)));
        bi[i * SV] = __dsub_rn(bi[i * SV], __dadd_rn(__dmul_rn(kr, xi), __dmul_rn(ki, xr)));
      }
    }
  }
#pragma unroll 1
  for (int k = N - 1; k >= 0; --k) {
    if (br[k * SV] != 0.0 || bi[k * SV] != 0.0) {
      double kr, ki;
      zdiv_dev(br[k * SV], bi[k * SV], ar[(k + N * k) * SV], ai[(k + N * k) * SV], kr, ki);
      br[k * SV] = kr; bi[k * SV] = ki;
#pragma unroll 1
      for (int i = 0; i < k; ++i) {
        const double xr = ar[(i + N * k) * SV], xi = ai[(i + N * k) * SV];
        br[i * SV] = __dsub_rn(br[i * SV], __dsub_rn(__dmul_rn(kr, xr), __dmul_rn(ki, xi)));
        bi[i * SV] = __dsub_rn(bi[i * SV], __dadd_rn(__dmul_rn(kr, xi), __dmul_rn(ki, xr)));
      }
    }
  }
}

// DGETRS('N') / DGETRS('T') for the lane-per-column sweeps of the general-pivoting pass: ONE dense factorisation per warp in
// shared memory (a: column-major 32 x 32, pivot rows as doubles at a[1024 + j], warp-uniform reads), the lane's right-hand side
// in its shared-memory slot (b[i * SB]).  Loop order and exact-zero skips of netlib (oracle/ref_lapack.hpp::dgetrs).
template <int SB>
__device__ __noinline__ void dense32_solve_n_shared(const double* __restrict__ a, double* b) {
  constexpr int N = 32;
#pragma unroll 1
  for (int i = 0; i < N; ++i) {
    const int ip = (int)a[N * N + i];
    if (ip != i) { const double t = b[i * SB]; b[i * SB] = b[ip * SB]; b[ip * SB] = t; }
  }
#pragma unroll 1
  for (int k = 0; k < N; ++k) {
    const double bk = b[k * SB];
    if (bk != 0.0) {
#pragma unroll 4
      for (int i = k + 1; i < N; ++i) b[i * SB] = __dsub_rn(b[i * SB], __dmul_rn(bk, a[i + N * k]));
    }
  }
#pragma unroll 1
  for (int k = N - 1; k >= 0; --k) {
    if (b[k * SB] != 0.0) {
      const double bk = __ddiv_rn(b[k * SB], a[k + N * k]);
      b[k * SB] = bk;
#pragma unroll 4
      for (int i = 0; i < k; ++i) b[i * SB] = __dsub_rn(b[i * SB], __dmul_rn(bk, a[i + N * k]));
    }
  }
}
template <int SB>
__device__ __noinline__ void dense32_solve_t_shared(const double* __restrict__ a, double* b) {
  constexpr int N = 32;
#pragma unroll 1
  for (int i = 0; i < N; ++i) {            // U^T x = b, dot-product form
    double temp = b[i * SB];
#pragma unroll 4
    for (int k = 0; k < i; ++k) temp = __dsub_rn(temp, __dmul_rn(a[k + N * i], b[k * SB]));
    b[i * SB] = __ddiv_rn(temp, a[i + N * i]);
  }
#pragma unroll 1
  for (int i = N - 1; i >= 0; --i) {       // L^T x = b (unit diagonal)
    double temp = b[i * SB];
#pragma unroll 4
    for (int k = i + 1; k < N; ++k) temp = __dsub_rn(temp, __dmul_rn(a[k + N * i], b[k * SB]));
    b[i * SB] = temp;
  }
#pragma unroll 1
  for (int i = N - 1; i >= 0; --i) {       // DLASWP backward
    const int ip = (int)a[N * N + i];
    if (ip != i) { const double t = b[i * SB]; b[i * SB] = b[ip * SB]; b[ip * SB] = t; }
  }
}

// ZGETRS('T') (plain transpose) for the lane-per-column sweep of the general-pivoting pass: the dense complex factors of the warp
// in shared memory (zr / zi column-major 32 x 32, pivot rows as doubles at zr[1024 + j]), the lane's right-hand side (br, bi) in
// its shared-memory slots.  Loop order, four-multiplication products and Smith quotients of oracle/ref_lapack.hpp::zgetrs_t.
template <int SB>
__device__ __noinline__ void dense32_zsolve_t_shared(const double* __restrict__ zr, const double* __restrict__ zi, double* br, double* bi) {
  constexpr int N = 32;
#pragma unroll 1
  for (int i = 0; i < N; ++i) {
    const double b0r = br[i * SB], b0i = bi[i * SB];
    double tr = __dsub_rn(__dmul_rn(1.0, b0r), __dmul_rn(0.0, b0i)), ti = __dadd_rn(__dmul_rn(1.0, b0i), __dmul_rn(0.0, b0r));
#pragma unroll 2
    for (int k = 0; k < i; ++k) {
      const double ar = zr[k + N * i], ai = zi[k + N * i], xr = br[k * SB], xi = bi[k * SB];
      tr = __dsub_rn(tr, __dsub_rn(__dmul_rn(ar, xr), __dmul_rn(ai, xi)));
      ti = __dsub_rn(ti, __dadd_rn(__dmul_rn(ar, xi), __dmul_rn(ai, xr)));
    }
    double qr, qi;
    zdiv_dev(tr, ti, zr[i + N * i], zi[i + N * i], qr, qi);
    br[i * SB] = qr; bi[i * SB] = qi;
  }
#pragma unroll 1
  for (int i = N - 1; i >= 0; --i) {
    const double b0r = br[i * SB], b0i = bi[i * SB];
    double tr = __dsub_rn(__dmul_rn(1.0, b0r), __dmul_rn(0.0, b0i)), ti = __dadd_rn(__dmul_rn(1.0, b0i), __dmul_rn(0.0, b0r));
#pragma unroll 2
    for (int k = i + 1; k < N; ++k) {
      const double ar = zr[k + N * i], ai = zi[k + N * i], xr = br[k * SB], xi = bi[k * SB];
      tr = __dsub_rn(tr, __dsub_rn(__dmul_rn(ar, xr), __dmul_rn(ai, xi)));
      ti = __dsub_rn(ti, __dadd_rn(__dmul_rn(ar, xi), __dmul_rn(ai, xr)));
    }
    br[i * SB] = tr; bi[i * SB] = ti;
  }
#pragma unroll 1
  for (int i = N - 1; i >= 0; --i) {
    const int ip = (int)zr[N * N + i];
    if (ip != i) {
      double t = br[i * SB]; br[i * SB] = br[ip * SB]; br[ip * SB] = t;
      t = bi[i * SB]; bi[i * SB] = bi[ip * SB]; bi[ip * SB] = t;
    }
  }
}

struct Cbm4DenseCell : Cbm4Cell {
  static constexpr int NLU = 32 * 32 + 32;   // dense matrix + pivot rows
  static constexpr bool HAS_GENERAL_PATH = false;
  template <int SV>
  __device__ static void build_E(const double* y, const double* PH, const Const& mc, double ghinv, double* lu) {
    cbm4_cell::build_E_dense<SV>(y, mc.fix, PH, ghinv, lu);
  }
  template <int SV>
  __device__ static int lu_factor(double* lu, double*, unsigned& swaps) { swaps = 0u; return dense32_lu_factor<SV>(lu); }
  template <int SV>
  __device__ static void lu_solve(const double* lu, unsigned, double* b) { dense32_lu_solve<SV>(lu, b); }
  template <int SV>
  __device__ static void build_EZ(const double* y, const double* PH, const Const& mc, double a, double b, double* zr, double* zi) {
    cbm4_cell::build_E_dense<SV>(y, mc.fix, PH, a, zr);
#pragma unroll 1
    for (int q = 0; q < 1024; ++q) zi[q * SV] = ((q & 31) == (q >> 5)) ? b : 0.0;
  }
  template <int SV>
  __device__ static int zlu_factor(double* zr, double* zi, unsigned& swaps) { swaps = 0u; return dense32_zlu_factor<SV>(zr, zi); }
  template <int SV>
  __device__ static void zlu_solve(const double* zr, const double* zi, unsigned, double* br, double* bi) {
    dense32_zlu_solve<SV>(zr, zi, br, bi);
  }
};

}  // namespace fatode
