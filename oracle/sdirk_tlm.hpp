// ORACLE — TEST INFRASTRUCTURE ONLY (never linked into the product path).
//
// CPU restatement of the reference's tangent linear model of the SDIRK family, FULL_ALGEBRA path:
//   TLM/SDIRK_TLM/SDIRK_TLM_f90_Integrator.F90   INTEGRATE_TLM :31-98 (presets ICNTRL(5) = 8, the `> 0` override rule),
//                                                SdirkTLM option parsing :236-455 (as FWD/SDIRK except: unknown
//                                                ICNTRL(3) selects Sdirk2a; ICNTRL(7) DirectTLM, ICNTRL(9) TLMNewtonEst,
//                                                ICNTRL(12) TLMtruncErr), SDIRK_IntegratorTLM :474-911,
//                                                SDIRK_ErrorNorm_TLM :947-960; coefficient sets identical to FWD/SDIRK's
//   TLM/SDIRK_TLM/LS_Solver.F90                  lss_jac1 :566-570, lss_mul_jac1 :555-558 (matmul(fjac1, g))
// Restated: the modified-Newton TLM (ICNTRL(7) = 0) with and without the TLM Newton estimate (ICNTRL(9)) and the TLM
// truncation error (ICNTRL(12)), and the direct TLM solve (ICNTRL(7) = 1, :332, :659-705: a second factorisation per stage at the
// stage point, LSS_Decomp_TLM / LSS_Solve_TLM).
// Differences from FWD/SDIRK that matter for ISTATUS: the LU is never reused across steps (:879), every stage of every
// attempt evaluates one more Jacobian (:701-702) and runs the TLM iterations of all NTLM columns — rejected attempts
// included — each iteration counting one Nsol (:727).
// PARITY UNPINNED by reference artefacts: the reference ships no CBM-IV driver and no golden file for this family (its TLM
// drivers exist for the shallow-water example only and write nothing).  The restatement is cross-validated against the ROS_TLM
// restatement, finite differences of its own state sweep and exact linearity in Y_tlm(t0) (tests/test_oracle_sdirk_tlm.py); its
// state sweep shares every routine with the golden-pinned FWD/SDIRK restatement (sdirk.hpp).
#pragma once
#include <vector>
#include "sdirk.hpp"

namespace oracle {

template <class Model>
struct SdirkTlm : Sdirk<Model> {
  typedef Sdirk<Model> Base;
  static constexpr int N = Model::N;
  using Base::cfg; using Base::ISTATUS; using Base::RSTATUS; using Base::lss; using Base::mdl;
  bool TLMNewtonEst = false, TLMtruncErr = false, DirectTLM = false;
  double e_tlm[Model::N * Model::N];
  int ip_tlm[Model::N];
  explicit SdirkTlm(Model& m) : Base(m) {}

  void scale_tlm(const double* AbsTol, const double* RelTol, const double* Y, double* SCAL) const {   // SDIRK_ErrorScale
    for (int i = 0; i < N; ++i)
      SCAL[i] = (cfg.ITOL == 0) ? 1.0 / (AbsTol[0] + RelTol[0] * std::fabs(Y[i])) : 1.0 / (AbsTol[i] + RelTol[i] * std::fabs(Y[i]));
  }

  // SDIRK_IntegratorTLM :474-911
  int integrate_tlm(int NTLM, double* Y, double* Y_TLM, double Tinitial, double Tfinal, const double* AbsTol, const double* RelTol,
                    const double* AbsTol_TLM, const double* RelTol_TLM) {
    const SdirkTableauData& tb = cfg.tab;
    const int S = tb.S;
    double Z[5][N], G[N], TMP[N], SCAL[N], DZ[N], SCAL_TLM[N], Yerr[N];
    std::vector<double> Z_TLM((size_t)N * 5 * NTLM, 0.0), Yerr_TLM((size_t)N * NTLM, 0.0);
    auto ztlm = [&](int istage, int itlm) { return Z_TLM.data() + ((size_t)itlm * 5 + istage) * N; };   // Z_TLM(1:N, istage, itlm)
    double NewtonRate = 0, Theta = 0, Hratio, NewtonPredictedErr, Qnewton, Err, Fac, Hnew, NewtonIncrement = 0, NewtonIncrementOld = 0,
           ThetaTLM;
    int saveNiter = 0;
    const double Roundoff = cfg.Roundoff;
    double T = Tinitial;
    const double Tdirection = (Tfinal - Tinitial) >= 0.0 ? 1.0 : -1.0;
    double H = std::fmax(std::fabs(cfg.Hmin), std::fabs(cfg.Hstart));
    if (std::fabs(H) <= 10.0 * Roundoff) H = 1.0e-6;
    H = std::fmin(std::fabs(H), cfg.Hmax);
    H = std::copysign(H, Tdirection);
    bool SkipLU = false, SkipJac = false, Reject = false, FirstStep = true;
    this->error_scale(AbsTol, RelTol, Y, SCAL);
    while ((Tfinal - T) * Tdirection - Roundoff > 0.0) {   // Tloop
      if (!SkipLU) {
        if (this->prepare_matrix(H, T, Y, SkipJac, SkipLU, Reject) != 0) return -8;
      }
      if (ISTATUS[Nstp] > cfg.Max_no_steps) return -6;
      if ((T + 0.1 * H == T) || (std::fabs(H) <= Roundoff)) return -7;
      bool cycle = false;
      for (int istage = 1; istage <= S && !cycle; ++istage) {
        double* Zi = Z[istage - 1];
        for (int i = 0; i < N; ++i) Zi[i] = 0.0;
        for (int i = 0; i < N; ++i) G[i] = 0.0;
        for (int j = 1; j <= istage - 1; ++j) {
          const double th = tb.Theta[istage - 1][j - 1];
          for (int i = 0; i < N; ++i) G[i] = G[i] + th * Z[j - 1][i];
          if (cfg.StartNewton) {
            const double al = tb.Alpha[istage - 1][j - 1];
            for (int i = 0; i < N; ++i) Zi[i] = Zi[i] + al * Z[j - 1][i];
          }
        }
        bool NewtonDone = false;
        Fac = 0.5;
        for (int NewtonIter = 1; NewtonIter <= cfg.NewtonMaxit; ++NewtonIter) {
          for (int i = 0; i < N; ++i) TMP[i] = Y[i] + Zi[i];
          mdl.fun(T + tb.C[istage - 1] * H, TMP, DZ);
          ISTATUS[Nfun]++;
          { const double a = H * tb.Gamma; for (int i = 0; i < N; ++i) DZ[i] = a * DZ[i]; }
          for (int i = 0; i < N; ++i) DZ[i] = DZ[i] + (-1.0) * Zi[i];
          for (int i = 0; i < N; ++i) DZ[i] = DZ[i] + 1.0 * G[i];
          this->solve(H, DZ);
          NewtonIncrement = Base::error_norm(DZ, SCAL);
          if (NewtonIter == 1) {
            Theta = std::fabs(cfg.ThetaMin);
            NewtonRate = 2.0;
          } else {
            Theta = NewtonIncrement / NewtonIncrementOld;
            if (Theta < 0.99) {
              NewtonRate = Theta / (1.0 - Theta);
              NewtonPredictedErr = NewtonIncrement * powi(Theta, cfg.NewtonMaxit - NewtonIter) / (1.0 - Theta);
              if (NewtonPredictedErr >= cfg.NewtonTol) {
                Qnewton = std::fmin(10.0, NewtonPredictedErr / cfg.NewtonTol);
                Fac = 0.8 * o_pow_neg_inv(Qnewton, (double)(1 + cfg.NewtonMaxit - NewtonIter));
                break;
              }
            } else {
              break;
            }
          }
          NewtonIncrementOld = NewtonIncrement;
          for (int i = 0; i < N; ++i) Zi[i] = Zi[i] + 1.0 * DZ[i];
          NewtonDone = (NewtonRate * NewtonIncrement <= cfg.NewtonTol);
          if (NewtonDone) { saveNiter = NewtonIter + 1; break; }
        }
        if (!NewtonDone) {
          H = Fac * H; Reject = true; SkipJac = true; SkipLU = false;
          cycle = true;
          break;
        }
        if (DirectTLM) {   // direct solution for the TLM variables (:659-697): a second factorisation, at the stage point
          for (int i = 0; i < N; ++i) TMP[i] = Y[i] + Zi[i];
          bool SkipJacT = false;
          int ConsecutiveSng = 0, IER = 1;
          double HGammaInv = 0.0;
          while (IER != 0) {
            HGammaInv = 1.0 / (H * tb.Gamma);
            if (!SkipJacT) { mdl.jac(T + tb.C[istage - 1] * H, TMP, lss.fjac1); ISTATUS[Njac]++; }
            for (int j = 0; j < N; ++j) {          // LSS_Decomp_TLM: e_tlm = HGammaInv I - fjac1, DGETRF
              for (int i = 0; i < N; ++i) e_tlm[i + N * j] = -lss.fjac1[i + N * j];
              e_tlm[j + N * j] = e_tlm[j + N * j] + HGammaInv;
            }
            IER = dgetf2(N, e_tlm, ip_tlm);
            ISTATUS[Ndec]++;
            if (IER != 0) {
              ISTATUS[Nsng]++;
              ConsecutiveSng++;
              if (ConsecutiveSng >= 6) return -8;   // "RETURN ! Failure" (the reference leaves IERR as it is)
              H = 0.5 * H; SkipJacT = true; SkipLU = false; Reject = true;
            }
          }
          for (int itlm = 0; itlm < NTLM; ++itlm) {
            double* Zt = ztlm(istage - 1, itlm);
            const double* Yt = Y_TLM + (size_t)itlm * N;
            for (int i = 0; i < N; ++i) G[i] = Yt[i];
            for (int j = 1; j <= istage - 1; ++j) {
              const double th = tb.Theta[istage - 1][j - 1];
              const double* Zj = ztlm(j - 1, itlm);
              for (int i = 0; i < N; ++i) G[i] = G[i] + th * Zj[i];
            }
            HGammaInv = 1.0 / (H * tb.Gamma);
            for (int i = 0; i < N; ++i) G[i] = HGammaInv * G[i];
            dgetrs(false, N, e_tlm, ip_tlm, G);
            ISTATUS[Nsol]++;
            for (int i = 0; i < N; ++i) Zt[i] = G[i] - Yt[i];
          }
          continue;
        }
        // modified-Newton TLM of this stage (:699-783)
        for (int i = 0; i < N; ++i) TMP[i] = Y[i] + Zi[i];
        mdl.jac(T + tb.C[istage - 1] * H, TMP, lss.fjac1);
        ISTATUS[Njac]++;
        for (int itlm = 0; itlm < NTLM && !cycle; ++itlm) {
          double* Zt = ztlm(istage - 1, itlm);
          const double* Yt = Y_TLM + (size_t)itlm * N;
          NewtonRate = std::pow(std::fmax(NewtonRate, Roundoff), 0.8);   // only read when TLMNewtonEst (then reset below)
          for (int i = 0; i < N; ++i) Zt[i] = 0.0;
          lss.mul_jac1(DZ, Yt);
          { const double a = H * tb.Gamma; for (int i = 0; i < N; ++i) G[i] = a * DZ[i]; }
          for (int j = 1; j <= istage - 1; ++j) {
            const double th = tb.Theta[istage - 1][j - 1];
            const double* Zj = ztlm(j - 1, itlm);
            for (int i = 0; i < N; ++i) G[i] = G[i] + th * Zj[i];
          }
          if (TLMNewtonEst) {
            NewtonDone = false;
            Fac = 0.5;
            scale_tlm(AbsTol_TLM + (size_t)itlm * N, RelTol_TLM + (size_t)itlm * N, Yt, SCAL_TLM);
          }
          for (int NewtonIterTLM = 1; NewtonIterTLM <= cfg.NewtonMaxit; ++NewtonIterTLM) {
            lss.mul_jac1(DZ, Zt);
            { const double a = H * tb.Gamma; for (int i = 0; i < N; ++i) DZ[i] = a * DZ[i] + G[i] - Zt[i]; }
            this->solve(H, DZ);
            if (TLMNewtonEst) {
              NewtonIncrement = Base::error_norm(DZ, SCAL_TLM);
              if (NewtonIterTLM <= 1) {
                ThetaTLM = std::fabs(cfg.ThetaMin);
                NewtonRate = 2.0;
              } else {
                ThetaTLM = NewtonIncrement / NewtonIncrementOld;
                if (ThetaTLM < 0.99) {
                  NewtonRate = ThetaTLM / (1.0 - ThetaTLM);
                  NewtonPredictedErr = NewtonIncrement * powi(ThetaTLM, cfg.NewtonMaxit - NewtonIterTLM) / (1.0 - ThetaTLM);
                  if (NewtonPredictedErr >= cfg.NewtonTol) {
                    Qnewton = std::fmin(10.0, NewtonPredictedErr / cfg.NewtonTol);
                    Fac = 0.8 * o_pow_neg_inv(Qnewton, (double)(1 + cfg.NewtonMaxit - NewtonIterTLM));
                    break;
                  }
                } else {
                  break;
                }
              }
              NewtonIncrementOld = NewtonIncrement;
            }
            for (int i = 0; i < N; ++i) Zt[i] = Zt[i] + 1.0 * DZ[i];
            if (TLMNewtonEst) {
              NewtonDone = (NewtonRate * NewtonIncrement <= cfg.NewtonTol);
              if (NewtonDone) break;
            } else {
              if (NewtonIterTLM >= saveNiter) break;
            }
          }
          if (TLMNewtonEst && !NewtonDone) {
            H = Fac * H; Reject = true; SkipJac = true; SkipLU = false;
            cycle = true;
          }
        }
      }
      if (cycle) continue;
      ISTATUS[Nstp]++;
      for (int i = 0; i < N; ++i) Yerr[i] = 0.0;
      for (int s = 1; s <= S; ++s)
        if (tb.E[s - 1] != 0.0) for (int i = 0; i < N; ++i) Yerr[i] = Yerr[i] + tb.E[s - 1] * Z[s - 1][i];
      this->solve(H, Yerr);
      Err = Base::error_norm(Yerr, SCAL);
      if (TLMtruncErr) {
        for (size_t q = 0; q < Yerr_TLM.size(); ++q) Yerr_TLM[q] = 0.0;
        for (int itlm = 0; itlm < NTLM; ++itlm) {
          double* Ye = Yerr_TLM.data() + (size_t)itlm * N;
          for (int j = 1; j <= S; ++j)
            if (tb.E[j - 1] != 0.0) { const double* Zj = ztlm(j - 1, itlm); for (int i = 0; i < N; ++i) Ye[i] = Ye[i] + tb.E[j - 1] * Zj[i]; }
          this->solve(H, Ye);
        }
        for (int itlm = 0; itlm < NTLM; ++itlm) {   // SDIRK_ErrorNorm_TLM: scaled by the error vector itself (:953-955)
          const double* Ye = Yerr_TLM.data() + (size_t)itlm * N;
          scale_tlm(AbsTol_TLM + (size_t)itlm * N, RelTol_TLM + (size_t)itlm * N, Ye, SCAL_TLM);
          Err = std::fmax(Err, Base::error_norm(Ye, SCAL_TLM));
        }
      }
      Fac = cfg.FacSafe * o_pow_neg_inv(Err, tb.ELO);
      Fac = std::fmax(cfg.FacMin, std::fmin(cfg.FacMax, Fac));
      Hnew = H * Fac;
      if (Err < 1.0) {
        FirstStep = false;
        ISTATUS[Nacc]++;
        T = T + H;
        for (int s = 1; s <= S; ++s)
          if (tb.D[s - 1] != 0.0) {
            for (int i = 0; i < N; ++i) Y[i] = Y[i] + tb.D[s - 1] * Z[s - 1][i];
            for (int itlm = 0; itlm < NTLM; ++itlm) {
              double* Yt = Y_TLM + (size_t)itlm * N;
              const double* Zs = ztlm(s - 1, itlm);
              for (int i = 0; i < N; ++i) Yt[i] = Yt[i] + tb.D[s - 1] * Zs[i];
            }
          }
        this->error_scale(AbsTol, RelTol, Y, SCAL);
        Hnew = Tdirection * std::fmin(std::fabs(Hnew), cfg.Hmax);
        RSTATUS[Ntexit] = T;
        RSTATUS[Nhexit] = H;
        RSTATUS[Nhnew] = Hnew;
        if (Reject) Hnew = Tdirection * std::fmin(std::fabs(Hnew), std::fabs(H));
        Reject = false;
        if ((T + Hnew / cfg.Qmin - Tfinal) * Tdirection > 0.0) {
          H = Tfinal - T;
        } else {
          Hratio = Hnew / H;
          SkipLU = false;   // "For TLM: do not skip LU" (:879)
          H = Hnew;
        }
        SkipJac = false;
      } else {
        if (FirstStep || Reject) H = cfg.FacRej * H; else H = Hnew;
        Reject = true; SkipJac = true; SkipLU = false;
        if (ISTATUS[Nacc] >= 1) ISTATUS[Nrej]++;
      }
    }
    (void)Hratio;
    return 1;
  }
};

// INTEGRATE_TLM of TLM/SDIRK_TLM (:31-98).  Y_TLM is (N,NTLM) column-major; AbsTol_TLM/RelTol_TLM (N,NTLM) may be NULL when
// neither ICNTRL(9) nor ICNTRL(12) is set.
template <class Model>
int sdirk_tlm_integrate(Model& mdl, int NTLM, double* Y, double* Y_TLM, double TIN, double TOUT, const double* ATOL_TLM,
                        const double* RTOL_TLM, const double* ATOL, const double* RTOL, const int* ICNTRL_U, const double* RCNTRL_U,
                        int* ISTATUS_U, double* RSTATUS_U) {
  int ICNTRL[20] = {0};
  double RCNTRL[20] = {0};
  ICNTRL[4] = 8;
  if (ICNTRL_U) for (int i = 0; i < 20; ++i) if (ICNTRL_U[i] > 0) ICNTRL[i] = ICNTRL_U[i];
  if (RCNTRL_U) for (int i = 0; i < 20; ++i) if (RCNTRL_U[i] > 0) RCNTRL[i] = RCNTRL_U[i];
  SdirkTlm<Model> s(mdl);
  int Ierr = sdirk_parse(Model::N, ICNTRL, RCNTRL, TIN, TOUT, ATOL, RTOL, s.cfg);
  if (ICNTRL[2] < 0 || ICNTRL[2] > 5) s.cfg.tab = SDIRK_TABLEAU[0];   // CASE DEFAULT: Sdirk2a (:262-275 of the TLM file)
  s.TLMNewtonEst = (ICNTRL[8] != 0);
  s.TLMtruncErr = (ICNTRL[11] != 0);
  s.DirectTLM = (ICNTRL[6] == 1);                                      // :332
  if (Ierr >= 0 && (s.TLMNewtonEst || s.TLMtruncErr) && (!ATOL_TLM || !RTOL_TLM)) Ierr = -100;
  if (Ierr >= 0) Ierr = s.integrate_tlm(NTLM, Y, Y_TLM, TIN, TOUT, ATOL, RTOL, ATOL_TLM, RTOL_TLM);
  if (ISTATUS_U) for (int i = 0; i < 20; ++i) ISTATUS_U[i] = s.ISTATUS[i];
  if (RSTATUS_U) for (int i = 0; i < 20; ++i) RSTATUS_U[i] = s.RSTATUS[i];
  return Ierr;
}

}  // namespace oracle
