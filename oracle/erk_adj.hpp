// ORACLE — TEST INFRASTRUCTURE ONLY (never linked into the product path).
//
// CPU restatement of the reference's explicit Runge-Kutta discrete adjoint and of the JAC callback of its shallow-water example:
//   ADJ/ERK_ADJ/ERK_ADJ_f90_Integrator.F90      INTEGRATE_ADJ :33-140 (the `> 0` override rule; without NP / Mu / JACP it calls
//                                               ERKADJ2), ERKADJ2 option parsing :1276-1560 (identical to FWD/ERK's), ERK_FwdInt
//                                               :1600-1740 (= ERK_Integrator plus ERK_Push(T, H, Y, Z) of every accepted step,
//                                               Z = the stage states), ERK_DadjInt (Lambda only) — per popped step, stages S..1:
//                                               JAC at the stage state, U_i = JAC^T (h sum_{j>i} a_ji U_j + h b_i Lambda) by
//                                               DGEMV('T'), then Lambda += sum_i U_i (:608-700 show the twin with Mu)
//   EXAMPLES/swe/SWE_ERK_ADJ/swe2D_erk_adj_dr.F90   jac :2-13 (vec2grid, compute_F for the eigenvalue signs, compute_JacF, fjac = JF),
//                                               adjinit :31-43 (Lambda = 0, Lambda(k,k) = 1)
//   EXAMPLES/swe/SWE_ERK_ADJ/swe2D_upwind.F90   compute_JacF :483-2023 through gen/swe_jac_gen.h (tools/gen_swe_jac.py), mapgr /
//                                               mapgrinv :128-176 (periodic neighbours)
// The reference builds the dense 4800 x 4800 matrix and multiplies with DGEMV('T'): y(j) = sum_i A(i,j) x(i), i ascending, exact
// zeros contribute nothing.  Here the 27 structural entries of each row are kept and scattered row by row, which performs the same
// additions in the same order for every column.
// PARITY UNPINNED by reference artefacts (the driver prints only): cross-checked against finite differences of the golden-pinned
// state sweep (the Jacobian, to the 1e-6 its six-digit literals allow) and against the tangent linear model by duality
// (tests/test_oracle_erk_adj.py).
#pragma once
#include <algorithm>
#include <cmath>
#include <utility>
#include <vector>
#include "erk.hpp"

namespace oracle {

struct ErkAdjSwe : Erk<Swe> {
  typedef Erk<Swe> Base;
  static constexpr int N = Swe::N;
  struct Entry { double T, H; std::vector<double> Y, Z; };
  std::vector<Entry> stack;
  explicit ErkAdjSwe(Swe& m) : Base(m) {}

  // ERK_FwdInt :1600-1740 — ERK_Integrator with the checkpoint of every accepted step
  int fwd(double* Y, double Tinitial, double Tfinal, const double* AbsTol, const double* RelTol) {
    const ErkTableauData& tb = cfg.tab;
    const int S = tb.S;
    std::vector<double> K((size_t)N * S), Z((size_t)N * S), TMP(N), SCAL(N);
    const double Roundoff = cfg.Roundoff;
    double T = Tinitial;
    const double Tdirection = (Tfinal - Tinitial) >= 0.0 ? 1.0 : -1.0;
    double H = std::fmax(std::fabs(cfg.Hmin), std::fabs(cfg.Hstart));
    if (std::fabs(H) <= 10.0 * Roundoff) H = 1.0e-6;
    H = std::fmin(std::fabs(H), cfg.Hmax);
    H = std::copysign(H, Tdirection);
    bool Reject = false, FirstStep = true;
    error_scale(AbsTol, RelTol, Y, SCAL.data());
    while ((Tfinal - T) * Tdirection - Roundoff > 0.0) {
      if (ISTATUS[Nstp] > cfg.Max_no_steps) return -6;
      if ((T + 0.1 * H == T) || (std::fabs(H) <= Roundoff)) return -7;
      for (int istage = 1; istage <= S; ++istage) {
        double* Ki = K.data() + (size_t)(istage - 1) * N;
        double* Zi = Z.data() + (size_t)(istage - 1) * N;
        for (int i = 0; i < N; ++i) Ki[i] = 0.0;
        for (int i = 0; i < N; ++i) Zi[i] = 0.0;
        for (int j = 1; j <= istage - 1; ++j) daxpy(N, H * tb.A[istage - 1][j - 1], K.data() + (size_t)(j - 1) * N, Zi);
        daxpy(N, 1.0, Y, Zi);
        mdl.fun(T + tb.C[istage - 1] * H, Zi, Ki);
        ISTATUS[Nfun]++;
      }
      ISTATUS[Nstp]++;
      for (int i = 0; i < N; ++i) TMP[i] = 0.0;
      for (int i = 1; i <= S; ++i)
        if (tb.E[i - 1] != 0.0) daxpy(N, H * tb.E[i - 1], K.data() + (size_t)(i - 1) * N, TMP.data());
      const double Err = error_norm(TMP.data(), SCAL.data());
      double Fac = cfg.FacSafe * o_pow_neg_inv(Err, tb.ELO);
      Fac = std::fmax(cfg.FacMin, std::fmin(cfg.FacMax, Fac));
      double Hnew = H * Fac;
      if (Err < 1.0) {
        FirstStep = false;
        ISTATUS[Nacc]++;
        if ((int)stack.size() + 1 > cfg.Max_no_steps) return -106;   // ERK_Push: "Push failed: buffer overflow", STOP (:766-770)
        stack.push_back(Entry{T, H, std::vector<double>(Y, Y + N), Z});
        T = T + H;
        for (int i = 1; i <= S; ++i)
          if (tb.B[i - 1] != 0.0) daxpy(N, H * tb.B[i - 1], K.data() + (size_t)(i - 1) * N, Y);
        error_scale(AbsTol, RelTol, Y, SCAL.data());
        Hnew = Tdirection * std::fmin(std::fabs(Hnew), cfg.Hmax);
        RSTATUS[Ntexit] = T;
        RSTATUS[Nhexit] = H;
        RSTATUS[Nhnew] = Hnew;
        if (Reject) Hnew = Tdirection * std::fmin(std::fabs(Hnew), std::fabs(H));
        Reject = false;
        if ((T + Hnew / cfg.Qmin - Tfinal) * Tdirection > 0.0) H = Tfinal - T;
        else H = Hnew;
      } else {
        if (FirstStep || Reject) H = cfg.FacRej * H; else H = Hnew;
        Reject = true;
        if (ISTATUS[Nacc] >= 1) ISTATUS[Nrej]++;
      }
    }
    return 1;
  }

  // ERK_DadjInt (Lambda only).  Lambda (N, NADJ) column-major.
  int dadj(int NADJ, double* Lambda) {
    const ErkTableauData& tb = cfg.tab;
    const int S = tb.S;
    std::vector<double> U((size_t)N * NADJ * S, 0.0), TMP(N);
    auto Uc = [&](int iadj, int istage) { return U.data() + ((size_t)istage * NADJ + iadj) * N; };
    SweJac J;
    while (!stack.empty()) {
      const Entry e = stack.back();
      stack.pop_back();
      const double H = e.H;
      ISTATUS[Nstp]++;
      for (int istage = S; istage >= 1; --istage) {
        J.build(mdl, e.Z.data() + (size_t)(istage - 1) * N);
        ISTATUS[Njac]++;
        for (int iadj = 0; iadj < NADJ; ++iadj) {
          for (int i = 0; i < N; ++i) TMP[i] = 0.0;
          for (int j = istage + 1; j <= S; ++j) daxpy(N, H * tb.A[j - 1][istage - 1], Uc(iadj, j - 1), TMP.data());
          daxpy(N, H * tb.B[istage - 1], Lambda + (size_t)iadj * N, TMP.data());
          J.mul_t(TMP.data(), Uc(iadj, istage - 1));
        }
      }
      for (int istage = 1; istage <= S; ++istage)
        for (int iadj = 0; iadj < NADJ; ++iadj) daxpy(N, 1.0, Uc(iadj, istage - 1), Lambda + (size_t)iadj * N);
    }
    return 1;
  }
};

// INTEGRATE_ADJ of ADJ/ERK_ADJ without parameters (ERKADJ2) for the shallow-water example
static inline int erk_adj_integrate_swe(Swe& mdl, int NADJ, double* Y, double* Lambda, double TIN, double TOUT, const double* ATOL,
                                        const double* RTOL, const int* ICNTRL_U, const double* RCNTRL_U, int* ISTATUS_U, double* RSTATUS_U) {
  int ICNTRL[20] = {0};
  double RCNTRL[20] = {0};
  if (ICNTRL_U) for (int i = 0; i < 20; ++i) if (ICNTRL_U[i] > 0) ICNTRL[i] = ICNTRL_U[i];
  if (RCNTRL_U) for (int i = 0; i < 20; ++i) if (RCNTRL_U[i] > 0) RCNTRL[i] = RCNTRL_U[i];
  ErkAdjSwe s(mdl);
  int Ierr = erk_parse(Swe::N, ICNTRL, RCNTRL, TIN, TOUT, ATOL, RTOL, false, s.cfg);
  if (Ierr >= 0) {
    Ierr = s.fwd(Y, TIN, TOUT, ATOL, RTOL);
    if (Ierr == -106) {   // the reference program STOPs in ERK_Push; restated as IERR -6 with Lambda untouched
      Ierr = -6;
    } else {              // a failed forward sweep still runs ADJINIT and the backward sweep over its accepted steps (:1563-1576)
      for (size_t q = 0; q < (size_t)Swe::N * NADJ; ++q) Lambda[q] = 0.0;
      for (int k = 0; k < NADJ; ++k) Lambda[(size_t)k * Swe::N + k] = 1.0;
      Ierr = s.dadj(NADJ, Lambda);
    }
  }
  if (ISTATUS_U) for (int i = 0; i < 20; ++i) ISTATUS_U[i] = s.ISTATUS[i];
  if (RSTATUS_U) for (int i = 0; i < 20; ++i) RSTATUS_U[i] = s.RSTATUS[i];
  return Ierr;
}

}  // namespace oracle
