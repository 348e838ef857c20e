// The JAC callback of the shallow-water example on the device, matrix-free: the 27 structural entries of every row of
// JAC = -JF - JG (compute_JacF, EXAMPLES/swe/SWE_ERK_ADJ/swe2D_upwind.F90:483-2023, identical in SWE_ERK_TLM; generated formulas in
// gen/swe_jac_gen.h, tools/gen_swe_jac.py) and the two products the integrators take with the dense matrix upstream:
// DGEMV('T') in ERK_DadjInt (ADJ/ERK_ADJ/ERK_ADJ_f90_Integrator.F90:668) and MATMUL(FJAC, S_TLM) in the dense-Jacobian tangent
// linear mode (TLM/ERK_TLM/ERK_TLM_f90_Integrator.F90:473-474).  Used by the 320-thread one-member-per-CTA kernels (erk_swe.cuh,
// erk_adj_swe.cu): rows / columns tid + 320 q.
#pragma once
#include <cmath>

#define SWE_JAC_FN __device__ __forceinline__
#define SWE_JAC_POW15(x) pow((x), 1.5)
#include "gen/swe_jac_gen.h"

namespace fatode {

constexpr int SWE_NNZR = 27;                 // structural entries of a Jacobian row: 3 fields x (5 along i + 4 more along j)

__device__ __forceinline__ int swe_wrap(int a) { return a < 0 ? a + SWE_MX : (a >= SWE_MX ? a - SWE_MX : a); }

// phase 1: JV[e][r] for the thread's rows r = tid + 320 q.  Z (the stage state) in shared memory.
__device__ __forceinline__ void swe_jac_rows(const double* __restrict__ Z, double* __restrict__ JV, const double dx) {
  const double g = 9.80616;
#pragma unroll 1
  for (int q = 0; q < ERK_EPT; ++q) {
    const int r = threadIdx.x + ERK_THREADS * q;
    const int kf = r / SWE_CELLS;                       // uniform over the CTA for a given q (1600 = 5 x 320)
    const int c = r - kf * SWE_CELLS;
    const int a = c / SWE_MX, b = c - a * SWE_MX;
    double V[3][9], JF[3][5], JG[3][5];
    bool fe[3], ge[3];
#pragma unroll
    for (int f = 0; f < 3; ++f) {
      const double* Zf = Z + f * SWE_CELLS;
#pragma unroll
      for (int d = -2; d <= 2; ++d) V[f][d + 2] = Zf[swe_wrap(a + d) * SWE_MX + b];
      V[f][5] = Zf[a * SWE_MX + swe_wrap(b - 2)];
      V[f][6] = Zf[a * SWE_MX + swe_wrap(b - 1)];
      V[f][7] = Zf[a * SWE_MX + swe_wrap(b + 1)];
      V[f][8] = Zf[a * SWE_MX + swe_wrap(b + 2)];
    }
    {   // feigvalues / geigvalues of compute_F (:356-358, :441-443)
      const double hc = V[0][2];
      const double cs = sqrt(g * hc);
      const double uc = V[1][2] / hc, vc = V[2][2] / hc;
      fe[0] = uc >= 0.0; fe[1] = (uc - cs) >= 0.0; fe[2] = (uc + cs) >= 0.0;
      ge[0] = vc >= 0.0; ge[1] = (vc - cs) >= 0.0; ge[2] = (vc + cs) >= 0.0;
    }
    if (kf == 0) swe_jac::row0(V, fe, ge, dx, dx, g, JF, JG);
    else if (kf == 1) swe_jac::row1(V, fe, ge, dx, dx, g, JF, JG);
    else swe_jac::row2(V, fe, ge, dx, dx, g, JF, JG);
#pragma unroll
    for (int cf = 0; cf < 3; ++cf) {                    // JF = -JF - JG (:2019)
#pragma unroll
      for (int d = 0; d < 5; ++d) JV[(size_t)(cf * 9 + d) * SWE_N + r] = -JF[cf][d] - (d == 2 ? JG[cf][2] : 0.0);
      JV[(size_t)(cf * 9 + 5) * SWE_N + r] = -0.0 - JG[cf][0];
      JV[(size_t)(cf * 9 + 6) * SWE_N + r] = -0.0 - JG[cf][1];
      JV[(size_t)(cf * 9 + 7) * SWE_N + r] = -0.0 - JG[cf][3];
      JV[(size_t)(cf * 9 + 8) * SWE_N + r] = -0.0 - JG[cf][4];
    }
  }
}

// phase 2: out[c] = sum_r JAC(r, c) x[r] for the thread's columns; x in shared memory.  Row (rf; a - d, b) reaches column
// (cf; a, b) through its first-index entry d, row (rf; a, b - d) through its second-index entry d.
__device__ __forceinline__ void swe_jac_t_gather(const double* __restrict__ JV, const double* __restrict__ x, double* __restrict__ out) {
#pragma unroll 1
  for (int q = 0; q < ERK_EPT; ++q) {
    const int c = threadIdx.x + ERK_THREADS * q;
    const int cf = c / SWE_CELLS;
    const int cc = c - cf * SWE_CELLS;
    const int a = cc / SWE_MX, b = cc - a * SWE_MX;
    const double* Jc = JV + (size_t)(cf * 9) * SWE_N;
    double t = 0.0;
#pragma unroll
    for (int rf = 0; rf < 3; ++rf) {
      const int base = rf * SWE_CELLS;
      int r;
      r = base + swe_wrap(a - 2) * SWE_MX + b; t = t + Jc[(size_t)4 * SWE_N + r] * x[r];     // row two above: its entry d = +2
      r = base + swe_wrap(a - 1) * SWE_MX + b; t = t + Jc[(size_t)3 * SWE_N + r] * x[r];
      r = base + a * SWE_MX + swe_wrap(b - 2); t = t + Jc[(size_t)8 * SWE_N + r] * x[r];     // second index: d = +2
      r = base + a * SWE_MX + swe_wrap(b - 1); t = t + Jc[(size_t)7 * SWE_N + r] * x[r];
      r = base + a * SWE_MX + b;               t = t + Jc[(size_t)2 * SWE_N + r] * x[r];     // centre
      r = base + a * SWE_MX + swe_wrap(b + 1); t = t + Jc[(size_t)6 * SWE_N + r] * x[r];     // d = -1
      r = base + a * SWE_MX + swe_wrap(b + 2); t = t + Jc[(size_t)5 * SWE_N + r] * x[r];
      r = base + swe_wrap(a + 1) * SWE_MX + b; t = t + Jc[(size_t)1 * SWE_N + r] * x[r];
      r = base + swe_wrap(a + 2) * SWE_MX + b; t = t + Jc[(size_t)0 * SWE_N + r] * x[r];
    }
    out[c] = t;
  }
}

// y = JAC x for the thread's rows (MATMUL's column sweep: per row the columns ascending, exact in the interior); x in shared memory
__device__ __forceinline__ void swe_jac_n_gather(const double* __restrict__ JV, const double* __restrict__ x, double* __restrict__ out) {
#pragma unroll 1
  for (int q = 0; q < ERK_EPT; ++q) {
    const int r = threadIdx.x + ERK_THREADS * q;
    const int rf = r / SWE_CELLS;
    const int rc = r - rf * SWE_CELLS;
    const int a = rc / SWE_MX, b = rc - a * SWE_MX;
    double t = 0.0;
#pragma unroll
    for (int cf = 0; cf < 3; ++cf) {
      const double* Jc = JV + (size_t)(cf * 9) * SWE_N + r;
      const double* xc = x + cf * SWE_CELLS;
      t = t + Jc[(size_t)0 * SWE_N] * xc[swe_wrap(a - 2) * SWE_MX + b];
      t = t + Jc[(size_t)1 * SWE_N] * xc[swe_wrap(a - 1) * SWE_MX + b];
      t = t + Jc[(size_t)5 * SWE_N] * xc[a * SWE_MX + swe_wrap(b - 2)];
      t = t + Jc[(size_t)6 * SWE_N] * xc[a * SWE_MX + swe_wrap(b - 1)];
      t = t + Jc[(size_t)2 * SWE_N] * xc[a * SWE_MX + b];
      t = t + Jc[(size_t)7 * SWE_N] * xc[a * SWE_MX + swe_wrap(b + 1)];
      t = t + Jc[(size_t)8 * SWE_N] * xc[a * SWE_MX + swe_wrap(b + 2)];
      t = t + Jc[(size_t)3 * SWE_N] * xc[swe_wrap(a + 1) * SWE_MX + b];
      t = t + Jc[(size_t)4 * SWE_N] * xc[swe_wrap(a + 2) * SWE_MX + b];
    }
    out[r] = t;
  }
}

}  // namespace fatode
