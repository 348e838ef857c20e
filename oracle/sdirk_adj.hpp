// ORACLE — TEST INFRASTRUCTURE ONLY (never linked into the product path).
//
// CPU restatement of the reference's discrete adjoint of the SDIRK family, FULL_ALGEBRA path:
//   ADJ/SDIRK_ADJ/SDIRK_ADJ_f90_Integrator.F90   INTEGRATE_ADJ :29-149 (presets ICNTRL(5) = 8, ICNTRL(7) = 1 direct adjoint solve,
//                                                ICNTRL(8) = 0; the `> 0` override rule), SDIRKADJ1 / SDIRKADJ2 option parsing
//                                                :153-540 / :1673-2055 (as FWD/SDIRK except Max_no_steps = 10000, starting values
//                                                always interpolated, DirectADJ = (ICNTRL(7) == 1), SaveLU always off),
//                                                SDIRK_FwdInt :545-781 (= SDIRK_Integrator + SDIRK_Push), SDIRK_DadjInt :786-995
//                                                (Lambda-only twin :2300-2489), SDIRK_PrepareMatrix :1199-1246,
//                                                SDIRK_Solve :1250-1268, coefficient sets :1272-1613 (gen/sdirk_adj_tableau_gen.h)
//   ADJ/SDIRK_ADJ/LS_Solver.F90                  lss_jac :308-312, lss_decomp, lss_solve (DGETRS 'N' / 'T'), lss_mul_jactr :345-348
// Restated: both adjoint solution modes (direct: one factorisation of 1/(h gamma) I - J(T + c_i h, Y + Z_i) per stage; Newton:
// simplified Newton iterations with the step's factorisation and the direct solve as the fall-back), without quadrature terms.
// Reference behaviour kept: the return code of a failed forward sweep is overwritten — ADJINIT and the backward sweep run over
// whatever was taped and INTEGRATE_ADJ returns 1 (:519-531).
#pragma once
#include <vector>
#include "sdirk.hpp"
#include "gen/sdirk_adj_tableau_gen.h"

namespace oracle {

template <class Model>
struct SdirkAdj : Sdirk<Model> {
  typedef Sdirk<Model> Base;
  static constexpr int N = Model::N;
  using Base::cfg; using Base::ISTATUS; using Base::RSTATUS; using Base::lss; using Base::mdl;
  std::vector<typename Base::TapeEntry> stack;
  SdirkAdjAB ab;
  bool DirectADJ = true;
  explicit SdirkAdj(Model& m) : Base(m) { this->tape = &stack; this->sng_keeps_jac = false; }

  void solve_t(bool trans, double H, double* RHS) {   // SDIRK_Solve :1250-1268
    const double HGammaInv = 1.0 / (H * cfg.tab.Gamma);
    for (int i = 0; i < N; ++i) RHS[i] = HGammaInv * RHS[i];
    lss.solve(trans, RHS);
    ISTATUS[Nsol]++;
  }
  static void daxpy(int n, double a, const double* x, double* y) {
    if (a == 0.0) return;
    for (int i = 0; i < n; ++i) y[i] = y[i] + a * x[i];
  }

  // SDIRK_DadjInt :786-995.  Lambda (N,NADJ), Mu (NP,NADJ) column-major; AbsTol_adj/RelTol_adj (N,NADJ) are read by the
  // Newton mode only.
  int dadj(int NADJ, int NP, double* Lambda, double* Mu, const double* AbsTol_adj, const double* RelTol_adj) {
    const SdirkTableauData& tb = cfg.tab;
    const int S = tb.S;
    std::vector<double> FPJAC((size_t)N * (NP > 0 ? NP : 1), 0.0), U((size_t)N * NADJ * 5, 0.0), V((size_t)(NP > 0 ? NP : 1) * NADJ * 5, 0.0);
    auto Uc = [&](int iadj, int istage) { return U.data() + ((size_t)istage * NADJ + iadj) * N; };          // U(1:N, iadj, istage)
    auto Vc = [&](int iadj, int istage) { return V.data() + ((size_t)istage * NADJ + iadj) * (NP > 0 ? NP : 1); };
    double TMP[N], G[N], SCAL[N], DU[N];
    double NewtonRate = 0, Theta = 0, NewtonPredictedErr, NewtonIncrement = 0, NewtonIncrementOld = 0, HGamma = 0;
    while (!stack.empty()) {
      const typename Base::TapeEntry e = stack.back();
      stack.pop_back();
      const double T = e.T;
      double H = e.H;
      const double* Y = e.Y;
      {   // IF (.NOT.SaveLU): SaveLU is always off
        bool SkipJac = false, SkipLU = false, Reject = false;
        if (this->prepare_matrix(H, T, Y, SkipJac, SkipLU, Reject) != 0) return -8;
      }
      ISTATUS[Nstp]++;
      for (int istage = S; istage >= 1; --istage) {
        for (int i = 0; i < N; ++i) TMP[i] = Y[i] + e.Z[istage - 1][i];
        mdl.jac(T + tb.C[istage - 1] * H, TMP, lss.fjac);
        ISTATUS[Njac]++;
        if (NP > 0) mdl.jacp(T + tb.C[istage - 1] * H, TMP, FPJAC.data());
        if (DirectADJ) {
          HGamma = 1.0 / (H * tb.Gamma);
          const int IER = lss.decomp(HGamma);
          ISTATUS[Ndec]++;
          if (IER != 0) return -8;
        }
        for (int iadj = 0; iadj < NADJ; ++iadj) {
          double* lam = Lambda + (size_t)iadj * N;
          if (!DirectADJ)
            for (int i = 0; i < N; ++i)
              SCAL[i] = (cfg.ITOL == 0) ? 1.0 / (AbsTol_adj[(size_t)iadj * N] + RelTol_adj[(size_t)iadj * N] * std::fabs(lam[i]))
                                        : 1.0 / (AbsTol_adj[(size_t)iadj * N + i] + RelTol_adj[(size_t)iadj * N + i] * std::fabs(lam[i]));
          for (int i = 0; i < N; ++i) G[i] = ab.B[istage - 1] * lam[i];
          for (int j = istage + 1; j <= S; ++j) daxpy(N, ab.A[j - 1][istage - 1], Uc(iadj, j - 1), G);
          lss.mul_jactr(TMP, G);
          for (int i = 0; i < N; ++i) G[i] = H * TMP[i];
          double* Ui = Uc(iadj, istage - 1);
          if (DirectADJ) {
            solve_t(true, H, G);
            for (int i = 0; i < N; ++i) Ui[i] = G[i];
          } else {
            for (int i = 0; i < N; ++i) Ui[i] = 0.0;
            bool NewtonDone = false;
            for (int NewtonIter = 1; NewtonIter <= cfg.NewtonMaxit; ++NewtonIter) {
              lss.mul_jactr(TMP, Ui);
              { const double hg = H * tb.Gamma; for (int i = 0; i < N; ++i) DU[i] = Ui[i] - hg * TMP[i] - G[i]; }
              solve_t(true, H, DU);
              NewtonIncrement = Base::error_norm(DU, SCAL);
              if (NewtonIter == 1) {
                Theta = std::fabs(cfg.ThetaMin);
                NewtonRate = 2.0;
              } else {
                Theta = NewtonIncrement / NewtonIncrementOld;
                if (Theta < 0.99) {
                  NewtonRate = Theta / (1.0 - Theta);
                  NewtonPredictedErr = NewtonIncrement * powi(Theta, cfg.NewtonMaxit - NewtonIter) / (1.0 - Theta);
                  if (NewtonPredictedErr >= cfg.NewtonTol) break;
                } else {
                  break;
                }
              }
              NewtonIncrementOld = NewtonIncrement;
              for (int i = 0; i < N; ++i) Ui[i] = Ui[i] - DU[i];
              NewtonDone = (NewtonRate * NewtonIncrement <= cfg.NewtonTol);
              if (NewtonIter >= 4 && NewtonDone) break;
            }
            if (!NewtonDone) {   // direct solution; HGamma keeps whatever an earlier direct solve left in it (:949)
              const int IER = lss.decomp(HGamma);
              ISTATUS[Ndec]++;
              if (IER != 0) return -8;
              solve_t(true, H, G);
              for (int i = 0; i < N; ++i) Ui[i] = G[i];
            }
          }
        }
        if (NP > 0 && Mu) {
          for (int iadj = 0; iadj < NADJ; ++iadj) {
            for (int i = 0; i < N; ++i) TMP[i] = 0.0;
            for (int j = istage; j <= S; ++j) daxpy(N, H * ab.A[j - 1][istage - 1], Uc(iadj, j - 1), TMP);
            daxpy(N, H * ab.B[istage - 1], Lambda + (size_t)iadj * N, TMP);
            double* Vi = Vc(iadj, istage - 1);
            for (int j = 0; j < NP; ++j) {   // DGEMV('T', alpha = 1, beta = 0)
              double t = 0.0;
              for (int i = 0; i < N; ++i) t = t + FPJAC[i + (size_t)N * j] * TMP[i];
              Vi[j] = 1.0 * t;
            }
          }
        }
      }
      for (int istage = 1; istage <= S; ++istage)
        for (int iadj = 0; iadj < NADJ; ++iadj) {
          double* lam = Lambda + (size_t)iadj * N;
          const double* Ui = Uc(iadj, istage - 1);
          for (int i = 0; i < N; ++i) lam[i] = lam[i] + Ui[i];
          if (NP > 0 && Mu) {
            double* mu = Mu + (size_t)iadj * NP;
            const double* Vi = Vc(iadj, istage - 1);
            for (int j = 0; j < NP; ++j) mu[j] = mu[j] + Vi[j];
          }
        }
    }
    return 1;
  }
};

// INTEGRATE_ADJ of ADJ/SDIRK_ADJ (:29-149).  Lambda (N,NADJ) column-major, Mu (NP,NADJ) or NULL; ADJINIT of the example drivers
// (Lambda = I, Mu = 0) runs after the forward sweep.  ATOL_adj/RTOL_adj may be NULL in the direct mode.
template <class Model>
int sdirk_adj_integrate(Model& mdl, int NP, int NADJ, double* Y, double* Lambda, double* Mu, double TIN, double TOUT,
                        const double* ATOL_adj, const double* RTOL_adj, const double* ATOL, const double* RTOL, const int* ICNTRL_U,
                        const double* RCNTRL_U, int* ISTATUS_U, double* RSTATUS_U) {
  int ICNTRL[20] = {0};
  double RCNTRL[20] = {0};
  ICNTRL[4] = 8; ICNTRL[6] = 1;
  if (ICNTRL_U) for (int i = 0; i < 20; ++i) if (ICNTRL_U[i] > 0) ICNTRL[i] = ICNTRL_U[i];
  if (RCNTRL_U) for (int i = 0; i < 20; ++i) if (RCNTRL_U[i] > 0) RCNTRL[i] = RCNTRL_U[i];
  SdirkAdj<Model> s(mdl);
  int ic2[20];
  for (int i = 0; i < 20; ++i) ic2[i] = ICNTRL[i];
  ic2[5] = 0;                                            // starting values are always interpolated in this family (:623)
  int Ierr = sdirk_parse(Model::N, ic2, RCNTRL, TIN, TOUT, ATOL, RTOL, s.cfg);
  if (ICNTRL[3] == 0) s.cfg.Max_no_steps = 10000;       // :375
  int method = (ICNTRL[2] == 1 || ICNTRL[2] == 2 || ICNTRL[2] == 3 || ICNTRL[2] == 5) ? ICNTRL[2] : 4;
  s.ab = SDIRK_ADJ_AB[method - 1];
  s.DirectADJ = (ICNTRL[6] == 1);
  const bool with_mu = (NP > 0 && Mu != nullptr);
  if (Ierr >= 0 && !s.DirectADJ && (!ATOL_adj || !RTOL_adj)) Ierr = -100;
  if (Ierr >= 0) {
    s.integrate(Y, TIN, TOUT, ATOL, RTOL);              // its return code is overwritten by the backward sweep's (:519-531)
    mdl.adjinit(NADJ, TOUT, Y, Lambda, with_mu ? Mu : nullptr);
    Ierr = s.dadj(NADJ, with_mu ? NP : 0, Lambda, with_mu ? Mu : nullptr, ATOL_adj, RTOL_adj);
  }
  if (ISTATUS_U) for (int i = 0; i < 20; ++i) ISTATUS_U[i] = s.ISTATUS[i];
  if (RSTATUS_U) for (int i = 0; i < 20; ++i) RSTATUS_U[i] = s.RSTATUS[i];
  return Ierr;
}

}  // namespace oracle
