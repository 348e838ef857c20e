// ORACLE — TEST INFRASTRUCTURE ONLY (never linked into the product path).
//
// CPU restatement of the reference's discrete adjoint of the fully implicit 3-stage Runge-Kutta methods
// (Radau-2A, Radau-1A, Lobatto-3C, Gauss), FULL_ALGEBRA path, with parameter sensitivities (RungeKuttaADJ1) or
// without (RungeKuttaADJ2: same code minus Mu):
//   ADJ/RK_ADJ/RK_ADJ_f90_Integrator.F90   INTEGRATE_ADJ :20-140 (presets ICNTRL(5)=8, (7)=1, (9)=2, (10)=1, (11)=1),
//                                          RungeKuttaADJ1 option parsing :389-610 and sequencing :611-655,
//                                          rk_DPush/rk_DPop :775-818, RK_FwdIntegrator :866-1226, rk_DadjInt :1229-1524,
//                                          RK_PrepareRHS :1707-1739, RK_PrepareRHS_Adj :1742-1784, RK_Decomp :1788-1813,
//                                          RK_Solve :1817-1848, RK_SolveTR :1852-1888; coefficient sets :1960-2459
//                                          (identical to FWD/RK's: checked entry by entry by tools/gen_rk_tableau.py)
//   ADJ/RK_ADJ/LS_Solver.F90               lapack_decomp(_cmp) :32-57, lapack_solve :83-92 (DGETRS 'n'/'t'),
//                                          lapack_solve_cmp :95-113 (ZGETRS 'n'/'t' — plain transpose),
//                                          lss_jac1..3 :704-720, lss_mul_jactr1..3 :835-848
// Built: the discrete adjoint with the fixed number of transposed simplified-Newton sweeps (ICNTRL(7) = 0/1, the preset
// and what cbm4_rk_adj_dr.F90 runs), ICNTRL(9) = 1 (forward only), and the direct solve (ICNTRL(7) = 2: lapack_decomp_big /
// lapack_solve_big :59-80,115-124, one DGETRF of the 3N x 3N system per step and one DGETRS('T') per column).  The adaptive
// solve (ICNTRL(7) = 3) and the quadrature terms are not restated: rk_adj_integrate returns -100.
// Differences from FWD/RK's forward sweep that matter for ISTATUS: no FUN(T,Y) at the top of the time loop; the SDIRK
// error stage evaluates and counts its own FUN(T,Y); ICNTRL(11) is preset to 1, i.e. the classic controller; default
// Max_no_steps = 10000 (= tape capacity); no Lobatto-3A.
#pragma once
#include <algorithm>
#include <vector>
#include "rk.hpp"

namespace oracle {

template <class Model>
struct RungeKuttaAdj : RungeKutta<Model> {
  typedef RungeKutta<Model> Base;
  static constexpr int N = Model::N;
  using Base::cfg; using Base::ISTATUS; using Base::RSTATUS; using Base::lss; using Base::mdl; using Base::e2; using Base::ip2;
  struct TapeEntry { double T, H, Y[N], Z[3 * N]; int NewIt; };
  std::vector<TapeEntry> tape;            // chk_T, chk_H, chk_Y, chk_Z, chk_NiT
  std::vector<double> fjac1, fjac2, fjac3, e_big;
  std::vector<int> ip_big;
  bool direct_solve = false;               // AdjointSolve == Solve_direct (ICNTRL(7) = 2): one 3N x 3N factorisation per step
  bool tlm_flavour = false;                // forward sweep as RK_IntegratorTLM runs it (rk_tlm.hpp): RK_PrepareRHS counts its three FUN
                                           // calls (TLM/RK_TLM :1120-1134) and the LU is never reused across steps (:934-935)
  explicit RungeKuttaAdj(Model& m) : Base(m), fjac1(N * N, 0.0), fjac2(N * N, 0.0), fjac3(N * N, 0.0) {}

  // RK_PrepareRHS :1707-1739 (no F0 argument, no Lobatto-3A)
  void prepare_rhs_adjfwd(double T, double H, const double* Y, const double* Z1, const double* Z2, const double* Z3, double* R1,
                          double* R2, double* R3) {
    const RkTableauData& tb = cfg.tab;
    double F[N], TMP[N];
    for (int i = 0; i < N; ++i) { R1[i] = Z1[i]; R2[i] = Z2[i]; R3[i] = Z3[i]; }
    auto axpy = [](double a, const double* x, double* y) { for (int i = 0; i < N; ++i) y[i] = y[i] + a * x[i]; };
    const double* Z[3] = {Z1, Z2, Z3};
    for (int s = 1; s <= 3; ++s) {
      for (int i = 0; i < N; ++i) TMP[i] = Y[i] + Z[s - 1][i];
      mdl.fun(T + tb.C[s] * H, TMP, F);
      if (tlm_flavour) ISTATUS[Nfun]++;
      axpy(-H * tb.A[1][s], F, R1); axpy(-H * tb.A[2][s], F, R2); axpy(-H * tb.A[3][s], F, R3);
    }
  }

  // RK_FwdIntegrator :866-1226 (SdirkError branch only: INTEGRATE_ADJ presets ICNTRL(10) = 1 and 0 cannot override it)
  int fwd_integrate(double* Y, double Tstart, double Tend, const double* AbsTol, const double* RelTol, bool discrete) {
    const RkTableauData& tb = cfg.tab;
    const double Roundoff = cfg.Roundoff;
    const int NewtonMaxit = cfg.NewtonMaxit;
    double Z1[N], Z2[N], Z3[N], Z4[N], SCAL[N], DZ1[N], DZ2[N], DZ3[N], DZ4[N], G[N], TMP[N], CONT[N][3];
    double Hacc = 0, Hnew = 0, Fac, FacGus, Theta = 0, Rerr, ErrOld = 0, NewtonRate = 0, NewtonIncrement = 0, Hratio, Qnewton,
           NewtonPredictedErr, NewtonIncrementOld = 0, ThetaSD;
    int NewtonIter = 0, Nconsecutive = 0, NewIt = 0;
    auto axpy = [](double a, const double* x, double* y) { for (int i = 0; i < N; ++i) y[i] = y[i] + a * x[i]; };
    double T = Tstart;
    const double Tdirection = (Tend - Tstart) >= 0.0 ? 1.0 : -1.0;
    double H = std::fmin(std::fmax(std::fabs(cfg.Hmin), std::fabs(cfg.Hstart)), cfg.Hmax);
    if (std::fabs(H) <= 10.0 * Roundoff) H = 1.0e-6;
    H = std::copysign(H, Tdirection);
    double Hold = H;
    bool Reject = false, FirstStep = true, SkipJac = false, SkipLU = false;
    if ((T + H * 1.0001 - Tend) * Tdirection >= 0.0) H = Tend - T;
    this->error_scale(AbsTol, RelTol, Y, SCAL);
    while ((Tend - T) * Tdirection - Roundoff > 0.0) {   // Tloop
      if (!SkipLU) {
        if (!SkipJac) { mdl.jac(T, Y, lss.fjac); ISTATUS[Njac]++; }
        int ISING = this->decomp(H);
        if (ISING != 0) {
          ISTATUS[Nsng]++; Nconsecutive++;
          if (Nconsecutive >= 5) return -12;
          H = H * 0.5; Reject = true; SkipJac = true; SkipLU = false;
          continue;
        } else {
          Nconsecutive = 0;
        }
      }
      ISTATUS[Nstp]++;
      if (ISTATUS[Nstp] > cfg.Max_no_steps) return -9;
      if (0.1 * std::fabs(H) <= std::fabs(T) * Roundoff) return -10;
      if (FirstStep || !cfg.StartNewton) {
        for (int i = 0; i < N; ++i) { Z1[i] = 0.0; Z2[i] = 0.0; Z3[i] = 0.0; }
      } else {
        this->interpolate_eval(H, Hold, Z1, Z2, Z3, CONT);
      }
      bool NewtonDone = false;
      Fac = 0.5;
      for (NewtonIter = 1; NewtonIter <= NewtonMaxit; ++NewtonIter) {
        prepare_rhs_adjfwd(T, H, Y, Z1, Z2, Z3, DZ1, DZ2, DZ3);
        this->solve(H, DZ1, DZ2, DZ3);
        {
          const double n1 = Base::error_norm(SCAL, DZ1), n2 = Base::error_norm(SCAL, DZ2), n3 = Base::error_norm(SCAL, DZ3);
          NewtonIncrement = std::sqrt((n1 * n1 + n2 * n2 + n3 * n3) / 3.0);
        }
        if (NewtonIter == 1) {
          Theta = std::fabs(cfg.ThetaMin);
          NewtonRate = 2.0;
        } else {
          Theta = NewtonIncrement / NewtonIncrementOld;
          if (Theta < 0.99) NewtonRate = Theta / (1.0 - Theta);
          else break;
          if (NewtonIter < NewtonMaxit) {
            NewtonPredictedErr = NewtonIncrement * powi(Theta, NewtonMaxit - NewtonIter) / (1.0 - Theta);
            if (NewtonPredictedErr >= cfg.NewtonTol) {
              Qnewton = std::fmin(10.0, NewtonPredictedErr / cfg.NewtonTol);
              Fac = 0.8 * o_pow_neg_inv(Qnewton, (double)(1 + NewtonMaxit - NewtonIter));
              break;
            }
          }
        }
        NewtonIncrementOld = std::fmax(NewtonIncrement, Roundoff);
        axpy(-1.0, DZ1, Z1); axpy(-1.0, DZ2, Z2); axpy(-1.0, DZ3, Z3);
        NewtonDone = (NewtonRate * NewtonIncrement <= cfg.NewtonTol);
        if (NewtonDone) { NewIt = NewtonIter; break; }
      }
      if (!NewtonDone) {
        H = Fac * H; Reject = true; SkipJac = true; SkipLU = false;
        continue;
      }
      // SDIRK error stage :1039-1094
      for (int i = 0; i < N; ++i) Z4[i] = Z3[i];
      mdl.fun(T, Y, DZ4);
      ISTATUS[Nfun]++;
      for (int i = 0; i < N; ++i) G[i] = 0.0;
      axpy(tb.Bgam[0] * H, DZ4, G);
      axpy(tb.Theta[1], Z1, G); axpy(tb.Theta[2], Z2, G); axpy(tb.Theta[3], Z3, G);
      NewtonDone = false;
      Fac = 0.5;
      for (NewtonIter = 1; NewtonIter <= NewtonMaxit; ++NewtonIter) {
        for (int i = 0; i < N; ++i) TMP[i] = Y[i] + Z4[i];
        mdl.fun(T + H, TMP, DZ4);
        ISTATUS[Nfun]++;
        axpy(-1.0 * tb.Gamma / H, Z4, DZ4);
        axpy(tb.Gamma / H, G, DZ4);
        lss.solve(false, DZ4);
        ISTATUS[Nsol]++;
        NewtonIncrement = Base::error_norm(SCAL, DZ4);
        if (NewtonIter == 1) {
          ThetaSD = std::fabs(cfg.ThetaMin);
          NewtonRate = 2.0;
        } else {
          ThetaSD = NewtonIncrement / NewtonIncrementOld;
          if (ThetaSD < 0.99) {
            NewtonRate = ThetaSD / (1.0 - ThetaSD);
            NewtonPredictedErr = NewtonIncrement * powi(ThetaSD, NewtonMaxit - NewtonIter) / (1.0 - ThetaSD);
            if (NewtonPredictedErr >= cfg.NewtonTol) {
              Qnewton = std::fmin(10.0, NewtonPredictedErr / cfg.NewtonTol);
              Fac = 0.8 * o_pow_neg_inv(Qnewton, (double)(1 + NewtonMaxit - NewtonIter));
              break;
            }
          } else {
            break;
          }
        }
        NewtonIncrementOld = NewtonIncrement;
        axpy(1.0, DZ4, Z4);
        NewtonDone = (NewtonRate * NewtonIncrement <= cfg.NewtonTol);
        if (NewtonDone) break;
      }
      if (!NewtonDone) {
        H = Fac * H; Reject = true; SkipJac = true; SkipLU = false;
        continue;
      }
      for (int i = 0; i < N; ++i) DZ4[i] = 0.0;
      if (tb.D[1] != 0.0) axpy(tb.D[1], Z1, DZ4);
      if (tb.D[2] != 0.0) axpy(tb.D[2], Z2, DZ4);
      if (tb.D[3] != 0.0) axpy(tb.D[3], Z3, DZ4);
      axpy(-1.0, Z4, DZ4);
      Rerr = Base::error_norm(SCAL, DZ4);
      Fac = o_pow_neg_inv(Rerr, tb.ELO) * std::fmin(cfg.FacSafe, (1.0 + 2 * NewtonMaxit) / (double)(NewtonIter + 2 * NewtonMaxit));
      Fac = std::fmin(cfg.FacMax, std::fmax(cfg.FacMin, Fac));
      Hnew = Fac * H;
      if (Rerr < 1.0) {   // accepted
        if (discrete) {
          TapeEntry e;
          e.T = T; e.H = H; e.NewIt = NewIt;
          for (int i = 0; i < N; ++i) { e.Y[i] = Y[i]; e.Z[i] = Z1[i]; e.Z[N + i] = Z2[i]; e.Z[2 * N + i] = Z3[i]; }
          tape.push_back(e);
        }
        FirstStep = false;
        ISTATUS[Nacc]++;
        if (cfg.Gustafsson) {
          if (ISTATUS[Nacc] > 1) {
            FacGus = cfg.FacSafe * (H / Hacc) * o_pow_neg_inv(Rerr * Rerr / ErrOld, 4.0);
            FacGus = std::fmin(cfg.FacMax, std::fmax(cfg.FacMin, FacGus));
            Fac = std::fmin(Fac, FacGus);
            Hnew = Fac * H;
          }
          Hacc = H;
          ErrOld = std::fmax(1.0e-2, Rerr);
        }
        Hold = H;
        T = T + H;
        if (tb.D[1] != 0.0) axpy(tb.D[1], Z1, Y);
        if (tb.D[2] != 0.0) axpy(tb.D[2], Z2, Y);
        if (tb.D[3] != 0.0) axpy(tb.D[3], Z3, Y);
        if (cfg.StartNewton) this->interpolate_make(Z1, Z2, Z3, CONT);
        this->error_scale(AbsTol, RelTol, Y, SCAL);
        RSTATUS[Ntexit] = T;
        RSTATUS[Nhnew] = Hnew;
        RSTATUS[Nhexit] = H;
        Hnew = Tdirection * std::fmin(std::fmax(std::fabs(Hnew), cfg.Hmin), cfg.Hmax);
        if (Reject) Hnew = Tdirection * std::fmin(std::fabs(Hnew), std::fabs(H));
        Reject = false;
        if ((T + Hnew / cfg.Qmin - Tend) * Tdirection >= 0.0) {
          H = Tend - T;
        } else {
          Hratio = Hnew / H;
          SkipLU = (Theta <= cfg.ThetaMin) && (Hratio >= cfg.Qmin) && (Hratio <= cfg.Qmax);
          if (tlm_flavour) SkipLU = false;   // "For TLM: do not skip LU"
          if (!SkipLU) H = Hnew;
        }
        SkipJac = false;
      } else {
        if (FirstStep || Reject) H = cfg.FacRej * H; else H = Hnew;
        Reject = true; SkipJac = true; SkipLU = false;
        if (ISTATUS[Nacc] >= 1) ISTATUS[Nrej]++;
      }
    }
    return 1;
  }

  static void mul_jactr(const double* fj, double* z, const double* g) {   // matmul(transpose(fjac_i), g), LS_Solver.F90:835-848
    for (int i = 0; i < N; ++i) { double s = 0.0; for (int k = 0; k < N; ++k) s += fj[k + N * i] * g[k]; z[i] = s; }
  }
  // RK_SolveTR :1852-1888
  void solve_tr(double H, double* R1, double* R2, double* R3) {
    const RkTableauData& tb = cfg.tab;
    for (int i = 0; i < N; ++i) {
      double x1 = R1[i] / H, x2 = R2[i] / H, x3 = R3[i] / H;
      R1[i] = tb.AinvT[0][0] * x1 + tb.AinvT[1][0] * x2 + tb.AinvT[2][0] * x3;
      R2[i] = tb.AinvT[0][1] * x1 + tb.AinvT[1][1] * x2 + tb.AinvT[2][1] * x3;
      R3[i] = tb.AinvT[0][2] * x1 + tb.AinvT[1][2] * x2 + tb.AinvT[2][2] * x3;
    }
    lss.solve(true, R1);
    cplx rhs[N];
    for (int i = 0; i < N; ++i) rhs[i] = cplx{R2[i], -R3[i]};
    zgetrs_t(N, e2, ip2, rhs);
    for (int i = 0; i < N; ++i) { R2[i] = rhs[i].re; R3[i] = -rhs[i].im; }
    for (int i = 0; i < N; ++i) {
      double x1 = R1[i], x2 = R2[i], x3 = R3[i];
      R1[i] = tb.Tinv[0][0] * x1 + tb.Tinv[1][0] * x2 + tb.Tinv[2][0] * x3;
      R2[i] = tb.Tinv[0][1] * x1 + tb.Tinv[1][1] * x2 + tb.Tinv[2][1] * x3;
      R3[i] = tb.Tinv[0][2] * x1 + tb.Tinv[1][2] * x2 + tb.Tinv[2][2] * x3;
    }
    ISTATUS[Nsol] += 2;
  }

  // rk_DadjInt :1229-1524, fixed number of sweeps (AdjointSolve == Solve_fixed), no quadrature
  int dadj(int NADJ, int NP, double* Lambda, double* Mu) {
    const RkTableauData& tb = cfg.tab;
    const int NewtonMaxit = cfg.NewtonMaxit;
    std::vector<double> FPJAC1((size_t)N * (NP > 0 ? NP : 1), 0.0), FPJAC2(FPJAC1), FPJAC3(FPJAC1);
    double U1[N], U2[N], U3[N], DU1[N], DU2[N], DU3[N], TMP[N], G1[N], G2[N], G3[N], F[N];
    std::vector<double> V1(NP > 0 ? NP : 1), V2(V1), V3(V1);
    auto axpy = [](int n, double a, const double* x, double* y) { for (int i = 0; i < n; ++i) y[i] = y[i] + a * x[i]; };
    while (!tape.empty()) {
      const TapeEntry e = tape.back();
      tape.pop_back();
      const double T = e.T, H = e.H;
      const double* Y = e.Y;
      const double* Zs = e.Z;
      mdl.jac(T, Y, lss.fjac);
      ISTATUS[Njac]++;
      this->decomp(H);
      ISTATUS[Njac] += 3;
      for (int i = 0; i < N; ++i) TMP[i] = Y[i] + Zs[i];
      mdl.jac(T + tb.C[1] * H, TMP, fjac1.data());
      if (NP > 0) mdl.jacp(T + tb.C[1] * H, TMP, FPJAC1.data());
      for (int i = 0; i < N; ++i) TMP[i] = Y[i] + Zs[N + i];
      mdl.jac(T + tb.C[2] * H, TMP, fjac2.data());
      if (NP > 0) mdl.jacp(T + tb.C[2] * H, TMP, FPJAC2.data());
      for (int i = 0; i < N; ++i) TMP[i] = Y[i] + Zs[2 * N + i];
      mdl.jac(T + tb.C[3] * H, TMP, fjac3.data());
      if (NP > 0) mdl.jacp(T + tb.C[3] * H, TMP, FPJAC3.data());
      const double* FJ[3] = {fjac1.data(), fjac2.data(), fjac3.data()};
      const double* FP[3] = {FPJAC1.data(), FPJAC2.data(), FPJAC3.data()};
      if (direct_solve) {   // LSS_Decomp_Big -> lapack_decomp_big (ADJ/RK_ADJ/LS_Solver.F90:59-80); its ISING is never looked at (:1330)
        constexpr int NB = 3 * N;
        e_big.resize((size_t)NB * NB);
        ip_big.resize(NB);
        for (int j = 0; j < N; ++j)
          for (int i = 0; i < N; ++i)
            for (int b = 0; b < 3; ++b)
              for (int a = 0; a < 3; ++a)
                e_big[(size_t)(a * N + i) + (size_t)NB * (b * N + j)] = -H * tb.A[a + 1][b + 1] * FJ[b][i + N * j];
        for (int i = 0; i < NB; ++i) e_big[(size_t)i + (size_t)NB * i] = e_big[(size_t)i + (size_t)NB * i] + 1;
        dgetf2(NB, e_big.data(), ip_big.data());   // DGETRF: blocked (NB = 64 < 96) in netlib, same operations per element in the same order
        ISTATUS[Ndec]++;
      }
      for (int iadj = 0; iadj < NADJ; ++iadj) {
        double* lam = Lambda + (size_t)iadj * N;
        for (int i = 0; i < N; ++i) { U1[i] = 0.0; U2[i] = 0.0; U3[i] = 0.0; G1[i] = 0.0; G2[i] = 0.0; G3[i] = 0.0; }
        mul_jactr(fjac1.data(), TMP, lam); axpy(N, -H * tb.B[1], TMP, G1);
        mul_jactr(fjac2.data(), TMP, lam); axpy(N, -H * tb.B[2], TMP, G2);
        mul_jactr(fjac3.data(), TMP, lam); axpy(N, -H * tb.B[3], TMP, G3);
        if (direct_solve) {   // :1466-1499: X = -G, one transposed solve with the big factors, U_s = X_s
          double X[3 * N];
          for (int i = 0; i < N; ++i) { X[i] = -G1[i]; X[N + i] = -G2[i]; X[2 * N + i] = -G3[i]; }
          dgetrs(true, 3 * N, e_big.data(), ip_big.data(), X);
          ISTATUS[Nsol]++;
          const double* Xs[3] = {X, X + N, X + 2 * N};
          if (NP > 0 && Mu) {
            double* mu = Mu + (size_t)iadj * NP;
            double* V[3] = {V1.data(), V2.data(), V3.data()};
            for (int s = 1; s <= 3; ++s) {
              for (int i = 0; i < N; ++i) TMP[i] = 0.0;
              axpy(N, H * tb.A[1][s], Xs[0], TMP); axpy(N, H * tb.A[2][s], Xs[1], TMP); axpy(N, H * tb.A[3][s], Xs[2], TMP);
              axpy(N, H * tb.B[s], lam, TMP);
              for (int j = 0; j < NP; ++j) {
                double t = 0.0;
                for (int i = 0; i < N; ++i) t = t + FP[s - 1][i + (size_t)N * j] * TMP[i];
                V[s - 1][j] = 1.0 * t;
              }
            }
            for (int i = 0; i < N; ++i) lam[i] = lam[i] + Xs[0][i] + Xs[1][i] + Xs[2][i];
            for (int j = 0; j < NP; ++j) mu[j] = mu[j] + V1[j] + V2[j] + V3[j];
          } else {
            for (int i = 0; i < N; ++i) lam[i] = lam[i] + Xs[0][i] + Xs[1][i] + Xs[2][i];
          }
          continue;
        }
        for (int NewtonIter = 1; NewtonIter <= NewtonMaxit; ++NewtonIter) {
          // RK_PrepareRHS_Adj :1742-1784
          double* U[3] = {U1, U2, U3};
          double* Gs[3] = {G1, G2, G3};
          double* R[3] = {DU1, DU2, DU3};
          for (int s = 1; s <= 3; ++s) {
            for (int i = 0; i < N; ++i) R[s - 1][i] = Gs[s - 1][i];
            for (int i = 0; i < N; ++i) F[i] = 0.0;
            axpy(N, -H * tb.A[1][s], U1, F); axpy(N, -H * tb.A[2][s], U2, F); axpy(N, -H * tb.A[3][s], U3, F);
            mul_jactr(FJ[s - 1], TMP, F);
            axpy(N, 1.0, U[s - 1], TMP);
            axpy(N, 1.0, TMP, R[s - 1]);
          }
          solve_tr(H, DU1, DU2, DU3);
          axpy(N, -1.0, DU1, U1); axpy(N, -1.0, DU2, U2); axpy(N, -1.0, DU3, U3);
          if (NewtonIter > std::max(e.NewIt + 1, 6)) break;
        }
        if (NP > 0 && Mu) {
          double* mu = Mu + (size_t)iadj * NP;
          double* V[3] = {V1.data(), V2.data(), V3.data()};
          for (int s = 1; s <= 3; ++s) {
            for (int i = 0; i < N; ++i) TMP[i] = 0.0;
            axpy(N, H * tb.A[1][s], U1, TMP); axpy(N, H * tb.A[2][s], U2, TMP); axpy(N, H * tb.A[3][s], U3, TMP);
            axpy(N, H * tb.B[s], lam, TMP);
            for (int j = 0; j < NP; ++j) {   // DGEMV('T', alpha = 1, beta = 0): y(j) = alpha * sum_i a(i,j) x(i)
              double t = 0.0;
              for (int i = 0; i < N; ++i) t = t + FP[s - 1][i + (size_t)N * j] * TMP[i];
              V[s - 1][j] = 1.0 * t;
            }
          }
          axpy(N, 1.0, U1, lam); axpy(N, 1.0, U2, lam); axpy(N, 1.0, U3, lam);
          axpy(NP, 1.0, V1.data(), mu); axpy(NP, 1.0, V2.data(), mu); axpy(NP, 1.0, V3.data(), mu);
        } else {
          axpy(N, 1.0, U1, lam); axpy(N, 1.0, U2, lam); axpy(N, 1.0, U3, lam);
        }
      }
      ISTATUS[Nstp]++;
    }
    return 1;
  }
};

// INTEGRATE_ADJ of ADJ/RK_ADJ (:20-140).  Lambda is (N,NADJ) column-major, Mu (NP,NADJ) or NULL; ADJINIT of the example
// drivers (Lambda = I, Mu = 0) runs after the forward sweep.
template <class Model>
int rk_adj_integrate(Model& mdl, int NP, int NADJ, double* Y, double* Lambda, double* Mu, double TIN, double TOUT,
                     const double* ATOL, const double* RTOL, const int* ICNTRL_U, const double* RCNTRL_U, int* ISTATUS_U,
                     double* RSTATUS_U) {
  int ICNTRL[20] = {0};
  double RCNTRL[20] = {0};
  ICNTRL[4] = 8; ICNTRL[6] = 1; ICNTRL[8] = 2; ICNTRL[9] = 1; ICNTRL[10] = 1;
  if (ICNTRL_U) for (int i = 0; i < 20; ++i) if (ICNTRL_U[i] > 0) ICNTRL[i] = ICNTRL_U[i];
  if (RCNTRL_U) for (int i = 0; i < 20; ++i) if (RCNTRL_U[i] > 0) RCNTRL[i] = RCNTRL_U[i];
  RungeKuttaAdj<Model> s(mdl);
  int IERR = 0;
  if (ICNTRL[2] == 5) {                                 // RK_ADJ has no Lobatto-3A (:411-423)
    IERR = -13;
    int ic2[20];
    for (int i = 0; i < 20; ++i) ic2[i] = ICNTRL[i];
    ic2[2] = 1;
    int other = rk_parse(Model::N, ic2, RCNTRL, TIN, TOUT, ATOL, RTOL, s.cfg);
    if (other < 0) IERR = other;                        // later checks overwrite IERR like the reference
  } else {
    IERR = rk_parse(Model::N, ICNTRL, RCNTRL, TIN, TOUT, ATOL, RTOL, s.cfg);
  }
  if (ICNTRL[3] == 0) s.cfg.Max_no_steps = 10000;       // :426-427
  bool discrete = true;
  if (ICNTRL[6] < 0 || ICNTRL[6] > 3) IERR = -9;        // ICNTRL(7) (:456-470): returns at once
  else if (ICNTRL[8] == 1) discrete = false;            // ICNTRL(9) = 1: no adjoint
  else if (ICNTRL[8] != 0 && ICNTRL[8] != 2) IERR = -9;
  const bool with_mu = (NP > 0 && Mu != nullptr);
  s.direct_solve = (ICNTRL[6] == 2);                       // Solve_direct (:458-459)
  if (IERR >= 0 && discrete && ICNTRL[6] == 3) IERR = -100;   // adaptive adjoint solve: not restated
  if (IERR >= 0) {
    IERR = s.fwd_integrate(Y, TIN, TOUT, ATOL, RTOL, discrete);
    if (IERR >= 0) {
      mdl.adjinit(NADJ, TOUT, Y, Lambda, with_mu ? Mu : nullptr);
      if (discrete) IERR = s.dadj(NADJ, with_mu ? NP : 0, Lambda, with_mu ? Mu : nullptr);
    }
  }
  if (ISTATUS_U) for (int i = 0; i < 20; ++i) ISTATUS_U[i] = s.ISTATUS[i];
  if (RSTATUS_U) for (int i = 0; i < 20; ++i) RSTATUS_U[i] = s.RSTATUS[i];
  return IERR;
}

}  // namespace oracle
