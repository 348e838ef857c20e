// ORACLE — TEST INFRASTRUCTURE ONLY (never linked into the product path).
//
// CPU restatement of the reference's forward fully implicit Runge-Kutta family (3-stage Radau-2A, Radau-1A, Lobatto-3C,
// Gauss, Lobatto-3A), FULL_ALGEBRA path:
//   FWD/RK/RK_f90_Integrator.F90   INTEGRATE :32-88, RungeKutta (option parsing) :92-459, RK_Integrator :461-780,
//                                  RK_ErrorScale :837-857, RK_Interpolate :877-913, RK_PrepareRHS :917-955,
//                                  RK_Decomp :959-988, RK_Solve :992-1023, RK_ErrorEstimate :1027-1083,
//                                  RK_ErrorNorm :1087-1101, coefficient sets :1104-1739 (gen/rk_tableau_gen.h)
//   FWD/RK/LS_Solver.F90           lapack_decomp :10-20 (e1 = -fjac, e1(j,j) += gamma/h; DGETRF),
//                                  lapack_decomp_cmp :22-32 (e2 = -fjac + CMPLX(alpha/h, beta/h); ZGETRF) — CMPLX without
//                                  a KIND is default (single precision) complex, so the shift is rounded to REAL(4)
//                                  (SURVEY fact 4), lapack_solve :34-38, lapack_solve_cmp :40-52 (ZGETRS 'N').
// Statement order, BLAS call order and counter updates follow the Fortran; compile with -ffp-contract=off.
// Elementary functions: x**(-1/k) and x**(-0.25d0) go through o_pow_neg_inv, Theta**n with an INTEGER exponent through
// powi (both in sdirk.hpp).
#pragma once
#include <cmath>
#include <cstring>
#include "ros.hpp"
#include "sdirk.hpp"
#include "gen/rk_tableau_gen.h"

namespace oracle {

struct RkConfig {
  RkTableauData tab;
  int ITOL, Max_no_steps, NewtonMaxit;
  bool StartNewton, Gustafsson, SdirkError;
  double Roundoff, Hmin, Hmax, Hstart, FacMin, FacMax, FacRej, FacSafe, ThetaMin, NewtonTol, Qmin, Qmax;
};

// RungeKutta :260-451.  Returns IERR (0 = fine so far); like the reference the checks keep going after an error and
// the last code wins.
static inline int rk_parse(int N, const int* ICNTRL, const double* RCNTRL, double T, double Tend, const double* AbsTol,
                           const double* RelTol, RkConfig& c) {
  int IERR = 0;
  c.ITOL = (ICNTRL[1] == 0) ? 1 : 0;
  c.SdirkError = (ICNTRL[9] != 0);
  int method = 0;
  switch (ICNTRL[2]) {
    case 1: method = 1; break;            // Radau2A
    case 0: case 2: method = 3; break;    // Lobatto3C (the actual default, fact 5)
    case 3: method = 4; break;            // Gauss
    case 4: method = 2; break;            // Radau1A
    case 5: method = 5; break;            // Lobatto3A
    default: IERR = -13;
  }
  if (method) c.tab = RK_TABLEAU[2 * (method - 1) + (c.SdirkError ? 1 : 0)];
  else std::memset(&c.tab, 0, sizeof c.tab);
  if (ICNTRL[3] == 0) c.Max_no_steps = 200000;
  else { c.Max_no_steps = ICNTRL[3]; if (c.Max_no_steps <= 0) IERR = -1; }
  if (ICNTRL[4] == 0) c.NewtonMaxit = 8;
  else { c.NewtonMaxit = ICNTRL[4]; if (c.NewtonMaxit <= 0) IERR = -2; }
  c.StartNewton = (ICNTRL[5] == 0);
  c.Gustafsson = (ICNTRL[10] == 0);
  c.Roundoff = dlamch_eps();
  c.Hmin = (RCNTRL[0] == 0.0) ? 0.0 : std::fmin(std::fabs(RCNTRL[0]), std::fabs(Tend - T));
  c.Hmax = (RCNTRL[1] == 0.0) ? std::fabs(Tend - T) : std::fmin(std::fabs(RCNTRL[1]), std::fabs(Tend - T));
  c.Hstart = (RCNTRL[2] == 0.0) ? 0.0 : std::fmin(std::fabs(RCNTRL[2]), std::fabs(Tend - T));
  c.FacMin = (RCNTRL[3] == 0.0) ? 0.2 : RCNTRL[3];
  c.FacMax = (RCNTRL[4] == 0.0) ? 8.0 : RCNTRL[4];
  c.FacRej = (RCNTRL[5] == 0.0) ? 0.1 : RCNTRL[5];
  c.FacSafe = (RCNTRL[6] == 0.0) ? 0.9 : RCNTRL[6];
  if ((c.FacMax < 1.0) || (c.FacMin > 1.0) || (c.FacSafe <= 1.0e-3) || (c.FacSafe >= 1.0)) IERR = -4;
  if (RCNTRL[7] == 0.0) c.ThetaMin = 1.0e-3;
  else { c.ThetaMin = RCNTRL[7]; if (c.ThetaMin <= 0.0 || c.ThetaMin >= 1.0) IERR = -5; }
  if (RCNTRL[8] == 0.0) c.NewtonTol = 3.0e-2;
  else { c.NewtonTol = RCNTRL[8]; if (c.NewtonTol <= c.Roundoff) IERR = -6; }
  c.Qmin = (RCNTRL[9] == 0.0) ? 1.0 : RCNTRL[9];
  c.Qmax = (RCNTRL[10] == 0.0) ? 1.2 : RCNTRL[10];
  if (c.Qmin > 1.0 || c.Qmax < 1.0) IERR = -7;
  if (c.ITOL == 0) {
    if (AbsTol[0] <= 0.0 || RelTol[0] <= 10.0 * c.Roundoff) IERR = -8;
  } else {
    for (int i = 0; i < N; ++i)
      if (AbsTol[i] <= 0.0 || RelTol[i] <= 10.0 * c.Roundoff) IERR = -8;
  }
  return IERR;
}

template <class Model>
struct RungeKutta {
  static constexpr int N = Model::N;
  enum { R2A = 1, R1A = 2, L3C = 3, GAU = 4, L3A = 5 };
  Model& mdl;
  RkConfig cfg;
  int ISTATUS[20];
  double RSTATUS[20];
  LssDense<N> lss;          // fjac, e1 (as e), ip1
  cplx e2[N * N];
  int ip2[N];
  explicit RungeKutta(Model& m) : mdl(m) {
    std::memset(ISTATUS, 0, sizeof ISTATUS); std::memset(RSTATUS, 0, sizeof RSTATUS);
    std::memset(e2, 0, sizeof e2); std::memset(ip2, 0, sizeof ip2);
  }

  void error_scale(const double* AbsTol, const double* RelTol, const double* Y, double* SCAL) const {   // :837-857
    for (int i = 0; i < N; ++i)
      SCAL[i] = (cfg.ITOL == 0) ? 1.0 / (AbsTol[0] + RelTol[0] * std::fabs(Y[i])) : 1.0 / (AbsTol[i] + RelTol[i] * std::fabs(Y[i]));
  }
  static double error_norm(const double* SCAL, const double* DY) {   // :1087-1101
    double s = 0.0;
    for (int i = 0; i < N; ++i) { double q = DY[i] * SCAL[i]; s = s + q * q; }
    return std::fmax(std::sqrt(s / (double)N), 1.0e-10);
  }
  // RK_Interpolate :877-913
  void interpolate_make(const double* Z1, const double* Z2, const double* Z3, double (*CONT)[3]) const {
    const double c1 = cfg.tab.C[1], c2 = cfg.tab.C[2], c3 = cfg.tab.C[3];
    const double den = (c3 - c2) * (c2 - c1) * (c1 - c3);
    for (int i = 0; i < N; ++i) {
      double s = -(c3 * c3) * c2 * Z1[i] + Z3[i] * c2 * (c1 * c1);
      s = s + (c2 * c2) * c3 * Z1[i];
      s = s - (c2 * c2) * c1 * Z3[i];
      s = s + (c3 * c3) * c1 * Z2[i];
      s = s - Z2[i] * c3 * (c1 * c1);
      CONT[i][0] = s / den - Z3[i];
      CONT[i][1] = -((c1 * c1) * (Z3[i] - Z2[i]) + (c2 * c2) * (Z1[i] - Z3[i]) + (c3 * c3) * (Z2[i] - Z1[i])) / den;
      CONT[i][2] = (c1 * (Z3[i] - Z2[i]) + c2 * (Z1[i] - Z3[i]) + c3 * (Z2[i] - Z1[i])) / den;
    }
  }
  void interpolate_eval(double H, double Hold, double* Z1, double* Z2, double* Z3, const double (*CONT)[3]) const {
    const double r = H / Hold;
    const double x1 = 1.0 + cfg.tab.C[1] * r, x2 = 1.0 + cfg.tab.C[2] * r, x3 = 1.0 + cfg.tab.C[3] * r;
    for (int i = 0; i < N; ++i) {
      Z1[i] = CONT[i][0] + x1 * (CONT[i][1] + x1 * CONT[i][2]);
      Z2[i] = CONT[i][0] + x2 * (CONT[i][1] + x2 * CONT[i][2]);
      Z3[i] = CONT[i][0] + x3 * (CONT[i][1] + x3 * CONT[i][2]);
    }
  }
  // RK_PrepareRHS :917-955 (its three FUN calls are not counted, fact 6)
  void prepare_rhs(double T, double H, const double* Y, const double* F0, const double* Z1, const double* Z2, const double* Z3,
                   double* R1, double* R2, double* R3) {
    const RkTableauData& tb = cfg.tab;
    double F[N], TMP[N];
    for (int i = 0; i < N; ++i) { R1[i] = Z1[i]; R2[i] = Z2[i]; R3[i] = Z3[i]; }
    auto axpy = [](double a, const double* x, double* y) { for (int i = 0; i < N; ++i) y[i] = y[i] + a * x[i]; };
    if (tb.method == L3A) {
      axpy(-H * tb.A[1][0], F0, R1); axpy(-H * tb.A[2][0], F0, R2); axpy(-H * tb.A[3][0], F0, R3);
    }
    const double* Z[3] = {Z1, Z2, Z3};
    for (int s = 1; s <= 3; ++s) {
      for (int i = 0; i < N; ++i) TMP[i] = Y[i] + Z[s - 1][i];
      mdl.fun(T + tb.C[s] * H, TMP, F);
      axpy(-H * tb.A[1][s], F, R1); axpy(-H * tb.A[2][s], F, R2); axpy(-H * tb.A[3][s], F, R3);
    }
  }
  // RK_Decomp :959-988 with lapack_decomp / lapack_decomp_cmp
  int decomp(double H) {
    const double Gamma = cfg.tab.Gamma / H, Alpha = cfg.tab.Alpha / H, Beta = cfg.tab.Beta / H;
    int ISING = lss.decomp(Gamma);
    if (ISING != 0) { ISTATUS[Ndec]++; return ISING; }
    const double a4 = (double)(float)Alpha, b4 = (double)(float)Beta;   // CMPLX(alpha,beta): default-kind complex
    for (int j = 0; j < N; ++j) {
      for (int i = 0; i < N; ++i) e2[i + N * j] = cplx{-lss.fjac[i + N * j], 0.0};
      e2[j + N * j] = cadd(e2[j + N * j], cplx{a4, b4});
    }
    ISING = zgetf2(N, e2, ip2);
    ISTATUS[Ndec]++;
    return ISING;
  }
  // RK_Solve :992-1023
  void solve(double H, double* R1, double* R2, double* R3) {
    const RkTableauData& tb = cfg.tab;
    for (int i = 0; i < N; ++i) {
      double x1 = R1[i] / H, x2 = R2[i] / H, x3 = R3[i] / H;
      R1[i] = tb.TinvAinv[0][0] * x1 + tb.TinvAinv[0][1] * x2 + tb.TinvAinv[0][2] * x3;
      R2[i] = tb.TinvAinv[1][0] * x1 + tb.TinvAinv[1][1] * x2 + tb.TinvAinv[1][2] * x3;
      R3[i] = tb.TinvAinv[2][0] * x1 + tb.TinvAinv[2][1] * x2 + tb.TinvAinv[2][2] * x3;
    }
    lss.solve(false, R1);
    cplx rhs[N];
    for (int i = 0; i < N; ++i) rhs[i] = cplx{R2[i], R3[i]};
    zgetrs_n(N, e2, ip2, rhs);
    for (int i = 0; i < N; ++i) { R2[i] = rhs[i].re; R3[i] = rhs[i].im; }
    for (int i = 0; i < N; ++i) {
      double x1 = R1[i], x2 = R2[i], x3 = R3[i];
      R1[i] = tb.T[0][0] * x1 + tb.T[0][1] * x2 + tb.T[0][2] * x3;
      R2[i] = tb.T[1][0] * x1 + tb.T[1][1] * x2 + tb.T[1][2] * x3;
      R3[i] = tb.T[2][0] * x1 + tb.T[2][1] * x2 + tb.T[2][2] * x3;
    }
    ISTATUS[Nsol] += 2;
  }
  // RK_ErrorEstimate :1027-1083 (classic estimator, ICNTRL(10) = 0 — not reachable through INTEGRATE, which presets 1)
  double error_estimate(double H, double T, const double* Y, const double* F0, const double* Z1, const double* Z2,
                        const double* Z3, const double* SCAL, bool FirstStep, bool Reject) {
    const RkTableauData& tb = cfg.tab;
    double F1[N], F2[N], TMP[N];
    const double HrkE1 = tb.E[1] / H, HrkE2 = tb.E[2] / H, HrkE3 = tb.E[3] / H;
    for (int i = 0; i < N; ++i) {
      F2[i] = HrkE1 * Z1[i] + HrkE2 * Z2[i] + HrkE3 * Z3[i];
      TMP[i] = tb.E[0] * F0[i] + F2[i];
    }
    lss.solve(false, TMP);
    ISTATUS[Nsol]++;
    if (tb.method == R1A || tb.method == GAU || tb.method == L3A) lss.solve(false, TMP);
    if (tb.method == GAU) lss.solve(false, TMP);
    double Err = error_norm(SCAL, TMP);
    if (Err < 1.0) return Err;
    if (FirstStep || Reject) {
      for (int i = 0; i < N; ++i) TMP[i] = Y[i] + TMP[i];
      mdl.fun(T, TMP, F1);
      ISTATUS[Nfun]++;
      for (int i = 0; i < N; ++i) TMP[i] = F1[i] + F2[i];
      lss.solve(false, TMP);
      ISTATUS[Nsol]++;
      Err = error_norm(SCAL, TMP);
    }
    return Err;
  }

  // RK_Integrator :461-780
  int integrate(double* Y, double T, double Tend, const double* AbsTol, const double* RelTol) {
    const RkTableauData& tb = cfg.tab;
    const double Roundoff = cfg.Roundoff;
    const int NewtonMaxit = cfg.NewtonMaxit;
    double Z1[N], Z2[N], Z3[N], Z4[N], SCAL[N], DZ1[N], DZ2[N], DZ3[N], DZ4[N], G[N], TMP[N], F0[N], CONT[N][3];
    double Hacc = 0, Hnew = 0, Fac, FacGus, Theta = 0, Err, ErrOld = 0, NewtonRate = 0, NewtonIncrement = 0, Hratio, Qnewton,
           NewtonPredictedErr, NewtonIncrementOld = 0, ThetaSD;
    int NewtonIter = 0, Nconsecutive = 0;
    auto axpy = [](double a, const double* x, double* y) { for (int i = 0; i < N; ++i) y[i] = y[i] + a * x[i]; };

    const double Tdirection = (Tend - T) >= 0.0 ? 1.0 : -1.0;
    double H = std::fmin(std::fmax(std::fabs(cfg.Hmin), std::fabs(cfg.Hstart)), cfg.Hmax);
    if (std::fabs(H) <= 10.0 * Roundoff) H = 1.0e-6;
    H = std::copysign(H, Tdirection);
    double Hold = H;
    bool Reject = false, FirstStep = true, SkipJac = false, SkipLU = false;
    if ((T + H * 1.0001 - Tend) * Tdirection >= 0.0) H = Tend - T;
    error_scale(AbsTol, RelTol, Y, SCAL);

    while ((Tend - T) * Tdirection - Roundoff > 0.0) {   // Tloop
      mdl.fun(T, Y, F0);
      ISTATUS[Nfun]++;
      if (!SkipLU) {
        if (!SkipJac) { mdl.jac(T, Y, lss.fjac); ISTATUS[Njac]++; }
        int ISING = decomp(H);
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
        interpolate_eval(H, Hold, Z1, Z2, Z3, CONT);
      }
      bool NewtonDone = false;
      Fac = 0.5;
      for (NewtonIter = 1; NewtonIter <= NewtonMaxit; ++NewtonIter) {   // NewtonLoop
        prepare_rhs(T, H, Y, F0, Z1, Z2, Z3, DZ1, DZ2, DZ3);
        solve(H, DZ1, DZ2, DZ3);
        {
          const double n1 = error_norm(SCAL, DZ1), n2 = error_norm(SCAL, DZ2), n3 = error_norm(SCAL, DZ3);
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
        if (NewtonDone) break;
      }
      if (!NewtonDone) {
        H = Fac * H; Reject = true; SkipJac = true; SkipLU = false;
        continue;
      }

      if (cfg.SdirkError) {   // SDIRK stage :604-666
        for (int i = 0; i < N; ++i) Z4[i] = Z3[i];
        for (int i = 0; i < N; ++i) G[i] = 0.0;
        if (tb.method != L3A) axpy(tb.Bgam[0] * H, F0, G);
        axpy(tb.Theta[1], Z1, G); axpy(tb.Theta[2], Z2, G); axpy(tb.Theta[3], Z3, G);
        NewtonDone = false;
        Fac = 0.5;
        for (NewtonIter = 1; NewtonIter <= NewtonMaxit; ++NewtonIter) {   // SDNewtonLoop
          for (int i = 0; i < N; ++i) TMP[i] = Y[i] + Z4[i];
          mdl.fun(T + H, TMP, DZ4);
          ISTATUS[Nfun]++;
          axpy(-1.0 * tb.Gamma / H, Z4, DZ4);
          axpy(tb.Gamma / H, G, DZ4);
          lss.solve(false, DZ4);
          ISTATUS[Nsol]++;
          NewtonIncrement = error_norm(SCAL, DZ4);
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
      }

      // error estimation :671-693
      if (cfg.SdirkError) {
        for (int i = 0; i < N; ++i) DZ4[i] = 0.0;
        if (tb.method == L3A) {
          for (int i = 0; i < N; ++i) DZ4[i] = H * tb.F[0] * F0[i];
          if (tb.F[1] != 0.0) axpy(tb.F[1], Z1, DZ4);
          if (tb.F[2] != 0.0) axpy(tb.F[2], Z2, DZ4);
          if (tb.F[3] != 0.0) axpy(tb.F[3], Z3, DZ4);
          for (int i = 0; i < N; ++i) TMP[i] = Y[i] + Z4[i];
          mdl.fun(T + H, TMP, DZ1);
          axpy(H * tb.Bgam[4], DZ1, DZ4);
        } else {
          if (tb.D[1] != 0.0) axpy(tb.D[1], Z1, DZ4);
          if (tb.D[2] != 0.0) axpy(tb.D[2], Z2, DZ4);
          if (tb.D[3] != 0.0) axpy(tb.D[3], Z3, DZ4);
          axpy(-1.0, Z4, DZ4);
        }
        Err = error_norm(SCAL, DZ4);
      } else {
        Err = error_estimate(H, T, Y, F0, Z1, Z2, Z3, SCAL, FirstStep, Reject);
      }

      // new step size :696-699 (NewtonIter keeps its value from the last Newton loop that ran)
      Fac = o_pow_neg_inv(Err, tb.ELO) * std::fmin(cfg.FacSafe, (1.0 + 2 * NewtonMaxit) / (double)(NewtonIter + 2 * NewtonMaxit));
      Fac = std::fmin(cfg.FacMax, std::fmax(cfg.FacMin, Fac));
      Hnew = Fac * H;

      if (Err < 1.0) {   // accepted
        FirstStep = false;
        ISTATUS[Nacc]++;
        if (cfg.Gustafsson) {
          if (ISTATUS[Nacc] > 1) {
            FacGus = cfg.FacSafe * (H / Hacc) * o_pow_neg_inv(Err * Err / ErrOld, 4.0);
            FacGus = std::fmin(cfg.FacMax, std::fmax(cfg.FacMin, FacGus));
            Fac = std::fmin(Fac, FacGus);
            Hnew = Fac * H;
          }
          Hacc = H;
          ErrOld = std::fmax(1.0e-2, Err);
        }
        Hold = H;
        T = T + H;
        if (tb.D[1] != 0.0) axpy(tb.D[1], Z1, Y);
        if (tb.D[2] != 0.0) axpy(tb.D[2], Z2, Y);
        if (tb.D[3] != 0.0) axpy(tb.D[3], Z3, Y);
        if (cfg.StartNewton) interpolate_make(Z1, Z2, Z3, CONT);
        error_scale(AbsTol, RelTol, Y, SCAL);
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
          if (!SkipLU) H = Hnew;
        }
        SkipJac = false;
      } else {   // rejected
        if (FirstStep || Reject) H = cfg.FacRej * H; else H = Hnew;
        Reject = true;
        SkipJac = true;
        SkipLU = false;
        if (ISTATUS[Nacc] >= 1) ISTATUS[Nrej]++;
      }
    }
    return 1;
  }
};

// INTEGRATE of FWD/RK (:32-88): presets ICNTRL(5) = 8, ICNTRL(10) = 1 (SDIRK error estimator); user values override the
// defaults only when > 0 — so the classic estimator and ICNTRL(6) = 0 cannot be selected through this interface.
// `raw` skips the presets/override rule and passes the arrays straight to RungeKutta (for the classic-estimator tests).
template <class Model>
int rk_integrate(Model& mdl, double TIN, double TOUT, double* VAR, const double* RTOL, const double* ATOL,
                 const int* ICNTRL_U, const double* RCNTRL_U, int* ISTATUS_U, double* RSTATUS_U, bool raw = false) {
  int ICNTRL[20] = {0};
  double RCNTRL[20] = {0};
  if (!raw) {
    ICNTRL[1] = 0; ICNTRL[4] = 8; ICNTRL[5] = 0; ICNTRL[9] = 1; ICNTRL[10] = 0;
    if (ICNTRL_U) for (int i = 0; i < 20; ++i) if (ICNTRL_U[i] > 0) ICNTRL[i] = ICNTRL_U[i];
    if (RCNTRL_U) for (int i = 0; i < 20; ++i) if (RCNTRL_U[i] > 0) RCNTRL[i] = RCNTRL_U[i];
  } else {
    if (ICNTRL_U) for (int i = 0; i < 20; ++i) ICNTRL[i] = ICNTRL_U[i];
    if (RCNTRL_U) for (int i = 0; i < 20; ++i) RCNTRL[i] = RCNTRL_U[i];
  }
  RungeKutta<Model> s(mdl);
  int IERR = rk_parse(Model::N, ICNTRL, RCNTRL, TIN, TOUT, ATOL, RTOL, s.cfg);
  if (IERR >= 0) IERR = s.integrate(VAR, TIN, TOUT, ATOL, RTOL);
  if (ISTATUS_U) for (int i = 0; i < 20; ++i) ISTATUS_U[i] = s.ISTATUS[i];
  if (RSTATUS_U) for (int i = 0; i < 20; ++i) RSTATUS_U[i] = s.RSTATUS[i];
  return IERR;
}

}  // namespace oracle
