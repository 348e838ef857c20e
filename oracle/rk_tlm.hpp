// ORACLE — TEST INFRASTRUCTURE ONLY (never linked into the product path).
//
// CPU restatement of the reference's tangent linear model of the 3-stage implicit Runge-Kutta family, FULL_ALGEBRA path:
//   TLM/RK_TLM/RK_TLM_f90_Integrator.F90   INTEGRATE_TLM :33-99 (presets ICNTRL(5) = 8, (7) = 1 direct TLM solve, (10) = 1 SDIRK
//                                          error estimator, (11) = 1 i.e. the CLASSIC controller (:385-389 tests == 0), the `> 0`
//                                          override rule: the modified-Newton TLM, ICNTRL(7) = 0, cannot be selected),
//                                          RungeKuttaTLM option handling :300-497 (Max_no_steps 200000, no Lobatto-3A),
//                                          RK_IntegratorTLM :505-965, RK_PrepareRHS_TLMdirect :1185-1227
//   TLM/RK_TLM/LS_Solver.F90               lapack_decomp_big / lapack_solve_big (the same routines as ADJ/RK_ADJ/LS_Solver.F90:59-80,
//                                          115-124), lss_mul_jac1..3 = matmul(fjac_i, g)
// Restated: the preset mode — direct TLM solve, truncation error of the state only (ICNTRL(12) = 0).  RK_IntegratorTLM is the
// forward sweep of ADJ/RK_ADJ (rk_adj.hpp::fwd_integrate: no FUN(T,Y) at the top of the loop, the SDIRK stage counts its own)
// with three differences: RK_PrepareRHS counts its FUN calls, the factorisation is never carried to the next step, and on every
// ACCEPTED step — before Y is advanced — the three stage Jacobians are evaluated (Njac += 3), e_big = I - H (A x J_b) is
// factorised (Ndec += 1) and every column solves Zbig = e_big^-1 (H A x J_b) Y_tlm (Nsol += 1), then Y_tlm += sum d_i Z_i.  The
// columns never feed back into the step sequence in this mode, so the restatement runs the state sweep with the tape of
// accepted steps (T, H, Y, Z_1..3: exactly what the reference has in hand at that point) and then the column pass over the tape.
// ICNTRL(12) /= 0 (RK_ErrorEstimate_TLM) returns -100.
// PARITY UNPINNED by reference artefacts: the reference ships no CBM-IV driver and no golden file for this family (its driver is
// the shallow-water example with a sparse solver).  Cross-checks (tests/test_rk_tlm.py): the state sweep against FWD/RK's
// golden-pinned restatement at identical tolerances, Y_tlm against finite differences of the state sweep, against the
// SDIRK_TLM / ROS_TLM restatements, and duality with the RK_ADJ restatement <Lambda(t0), dy0> = <Lambda(T), Y_tlm dy0>.
#pragma once
#include "rk_adj.hpp"

namespace oracle {

template <class Model>
struct RungeKuttaTlm : RungeKuttaAdj<Model> {
  typedef RungeKuttaAdj<Model> Base;
  static constexpr int N = Model::N;
  using Base::cfg; using Base::ISTATUS; using Base::mdl; using Base::tape; using Base::fjac1; using Base::fjac2; using Base::fjac3;
  using Base::e_big; using Base::ip_big;
  explicit RungeKuttaTlm(Model& m) : Base(m) { this->tlm_flavour = true; }

  static void mul_jac(const double* fj, double* z, const double* g) {   // matmul(fjac_i, g): libgfortran's column sweep
    for (int i = 0; i < N; ++i) z[i] = 0.0;
    for (int k = 0; k < N; ++k) for (int i = 0; i < N; ++i) z[i] += fj[i + N * k] * g[k];
  }

  // the TLM block of RK_IntegratorTLM :761-793 and the update :909-914, for every accepted step in forward order
  int tlm_direct(int NTLM, double* Y_TLM) {
    const RkTableauData& tb = cfg.tab;
    constexpr int NB = 3 * N;
    double TMP[N], F[N], Zbig[NB];
    e_big.resize((size_t)NB * NB);
    ip_big.resize(NB);
    for (size_t q = 0; q < tape.size(); ++q) {
      const typename Base::TapeEntry& e = tape[q];
      const double T = e.T, H = e.H;
      for (int i = 0; i < N; ++i) TMP[i] = e.Y[i] + e.Z[i];
      mdl.jac(T + tb.C[1] * H, TMP, fjac1.data());
      for (int i = 0; i < N; ++i) TMP[i] = e.Y[i] + e.Z[N + i];
      mdl.jac(T + tb.C[2] * H, TMP, fjac2.data());
      for (int i = 0; i < N; ++i) TMP[i] = e.Y[i] + e.Z[2 * N + i];
      mdl.jac(T + tb.C[3] * H, TMP, fjac3.data());
      ISTATUS[Njac] += 3;
      const double* FJ[3] = {fjac1.data(), fjac2.data(), fjac3.data()};
      for (int j = 0; j < N; ++j)
        for (int i = 0; i < N; ++i)
          for (int b = 0; b < 3; ++b)
            for (int a = 0; a < 3; ++a)
              e_big[(size_t)(a * N + i) + (size_t)NB * (b * N + j)] = -H * tb.A[a + 1][b + 1] * FJ[b][i + N * j];
      for (int i = 0; i < NB; ++i) e_big[(size_t)i + (size_t)NB * i] = e_big[(size_t)i + (size_t)NB * i] + 1;
      const int inf = dgetf2(NB, e_big.data(), ip_big.data());
      ISTATUS[Ndec]++;
      if (inf != 0) return -12;   // the reference STOPs here ("Big big guy is singular")
      for (int itlm = 0; itlm < NTLM; ++itlm) {
        double* Yt = Y_TLM + (size_t)itlm * N;
        for (int b = 0; b < 3; ++b) {   // RK_PrepareRHS_TLMdirect
          mul_jac(FJ[b], F, Yt);
          for (int a = 0; a < 3; ++a)
            for (int i = 0; i < N; ++i) {
              const double t = H * tb.A[a + 1][b + 1] * F[i];
              Zbig[a * N + i] = (b == 0) ? t : Zbig[a * N + i] + t;
            }
        }
        dgetrs(false, NB, e_big.data(), ip_big.data(), Zbig);
        ISTATUS[Nsol]++;
        for (int s = 1; s <= 3; ++s)
          if (tb.D[s] != 0.0) for (int i = 0; i < N; ++i) Yt[i] = Yt[i] + tb.D[s] * Zbig[(s - 1) * N + i];
      }
    }
    return 1;
  }
};

// INTEGRATE_TLM of TLM/RK_TLM (:33-99).  Y_TLM is (N,NTLM) column-major.
template <class Model>
int rk_tlm_integrate(Model& mdl, int NTLM, double* Y, double* Y_TLM, double TIN, double TOUT, const double* ATOL, const double* RTOL,
                     const int* ICNTRL_U, const double* RCNTRL_U, int* ISTATUS_U, double* RSTATUS_U) {
  int ICNTRL[20] = {0};
  double RCNTRL[20] = {0};
  ICNTRL[4] = 8; ICNTRL[6] = 1; ICNTRL[9] = 1; ICNTRL[10] = 1;
  if (ICNTRL_U) for (int i = 0; i < 20; ++i) if (ICNTRL_U[i] > 0) ICNTRL[i] = ICNTRL_U[i];
  if (RCNTRL_U) for (int i = 0; i < 20; ++i) if (RCNTRL_U[i] > 0) RCNTRL[i] = RCNTRL_U[i];
  RungeKuttaTlm<Model> s(mdl);
  int IERR = 0;
  if (ICNTRL[2] == 5) {                                 // no Lobatto-3A in this family (:321-334)
    IERR = -13;
    int ic2[20];
    for (int i = 0; i < 20; ++i) ic2[i] = ICNTRL[i];
    ic2[2] = 1;
    const int other = rk_parse(Model::N, ic2, RCNTRL, TIN, TOUT, ATOL, RTOL, s.cfg);
    if (other < 0) IERR = other;
  } else {
    IERR = rk_parse(Model::N, ICNTRL, RCNTRL, TIN, TOUT, ATOL, RTOL, s.cfg);
  }
  if (IERR >= 0 && ICNTRL[11] != 0) IERR = -100;       // TLM truncation-error control: not restated
  if (IERR >= 0) {
    IERR = s.fwd_integrate(Y, TIN, TOUT, ATOL, RTOL, /*discrete=*/true);
    // the reference advances the columns inside the accepted steps, so whatever was accepted before a failure is applied
    const int e2 = s.tlm_direct(NTLM, Y_TLM);
    if (IERR >= 0 && e2 < 0) IERR = e2;
  }
  if (ISTATUS_U) for (int i = 0; i < 20; ++i) ISTATUS_U[i] = s.ISTATUS[i];
  if (RSTATUS_U) for (int i = 0; i < 20; ++i) RSTATUS_U[i] = s.RSTATUS[i];
  return IERR;
}

}  // namespace oracle
