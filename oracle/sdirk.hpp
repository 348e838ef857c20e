// ORACLE — TEST INFRASTRUCTURE ONLY (never linked into the product path).
//
// CPU restatement of the reference's forward SDIRK family, FULL_ALGEBRA path:
//   FWD/SDIRK/SDIRK_f90_Integrator.F90   INTEGRATE :25-90, SDIRK (option parsing) :92-425,
//                                        SDIRK_Integrator :430-657, SDIRK_ErrorScale :661-676,
//                                        SDIRK_ErrorNorm :680-691, SDIRK_PrepareMatrix :736-785, SDIRK_Solve :789-810,
//                                        coefficient sets :814-1159 (gen/sdirk_tableau_gen.h)
//   FWD/SDIRK/LS_Solver.F90              lapack_decomp / lapack_solve (same as FWD/ROS: e = -fjac, e(j,j) += 1/(h gamma))
// Statement order, BLAS call order and counter updates follow the Fortran; compile with -ffp-contract=off.
// Elementary functions: Err**(-1/rkELO), Qnewton**(-1/k) go through o_pow_neg_inv (portable kernels by default, libm pow
// in ORACLE_LIBM mode, see portable_math.hpp); Theta**n with an INTEGER exponent is gfortran's __builtin_powi
// (libgcc __powidf2: square-and-multiply), restated in powi().
#pragma once
#include <cmath>
#include <cstring>
#include <vector>
#include "ros.hpp"
#include "gen/sdirk_tableau_gen.h"

namespace oracle {

static inline double powi(double x, int m) {   // libgcc __powidf2
  unsigned n = m < 0 ? 0u - (unsigned)m : (unsigned)m;
  double y = (n % 2) ? x : 1.0;
  while (n >>= 1) {
    x = x * x;
    if (n % 2) y = y * x;
  }
  return m < 0 ? 1.0 / y : y;
}
// x**(-1/k) for a small positive integer-valued k (k = rkELO or 1+NewtonMaxit-NewtonIter)
static inline double o_pow_neg_inv(double x, double k) {
  return libm_mode() ? std::pow(x, -1.0 / k) : oracle_pm::pow_neg_inv(x, k);
}

struct SdirkConfig {
  SdirkTableauData tab;
  int ITOL, Max_no_steps, NewtonMaxit;
  bool StartNewton;
  double Roundoff, Hmin, Hmax, Hstart, FacMin, FacMax, FacRej, FacSafe, ThetaMin, NewtonTol, Qmin, Qmax;
};

// SDIRK :236-417.  Returns Ierr (0 = fine so far; the reference keeps going after an option error and returns the
// last code before integrating).
static inline int sdirk_parse(int N, const int* ICNTRL, const double* RCNTRL, double Tinitial, double Tfinal,
                              const double* AbsTol, const double* RelTol, SdirkConfig& c) {
  int Ierr = 0;
  c.ITOL = (ICNTRL[1] == 0) ? 1 : 0;
  int method;
  switch (ICNTRL[2]) {
    case 1: method = 1; break;
    case 2: method = 2; break;
    case 3: method = 3; break;
    case 0: case 4: method = 4; break;
    case 5: method = 5; break;
    default: method = 4;
  }
  c.tab = SDIRK_TABLEAU[method - 1];
  if (ICNTRL[3] == 0) c.Max_no_steps = 200000;
  else if (ICNTRL[3] > 0) c.Max_no_steps = ICNTRL[3];
  else { c.Max_no_steps = 0; Ierr = -1; }
  if (ICNTRL[4] == 0) c.NewtonMaxit = 8;
  else { c.NewtonMaxit = ICNTRL[4]; if (c.NewtonMaxit <= 0) Ierr = -2; }
  c.StartNewton = (ICNTRL[5] == 0);
  c.Roundoff = dlamch_eps();
  if (RCNTRL[0] == 0.0) c.Hmin = 0.0; else if (RCNTRL[0] > 0.0) c.Hmin = RCNTRL[0]; else { c.Hmin = 0.0; Ierr = -3; }
  if (RCNTRL[1] == 0.0) c.Hmax = std::fabs(Tfinal - Tinitial);
  else if (RCNTRL[1] > 0.0) c.Hmax = std::fmin(std::fabs(RCNTRL[1]), std::fabs(Tfinal - Tinitial));
  else { c.Hmax = std::fabs(Tfinal - Tinitial); Ierr = -3; }
  if (RCNTRL[2] == 0.0) c.Hstart = std::fmax(c.Hmin, c.Roundoff);
  else if (RCNTRL[2] > 0.0) c.Hstart = std::fmin(std::fabs(RCNTRL[2]), std::fabs(Tfinal - Tinitial));
  else { c.Hstart = 0.0; Ierr = -3; }
  // REAL(4) literals 0.2, 10.0, 0.1, 0.9 (:341-367)
  if (RCNTRL[3] == 0.0) c.FacMin = (double)0.2f; else if (RCNTRL[3] > 0.0) c.FacMin = RCNTRL[3]; else { c.FacMin = 0; Ierr = -4; }
  if (RCNTRL[4] == 0.0) c.FacMax = (double)10.0f; else if (RCNTRL[4] > 0.0) c.FacMax = RCNTRL[4]; else { c.FacMax = 0; Ierr = -4; }
  if (RCNTRL[5] == 0.0) c.FacRej = (double)0.1f; else if (RCNTRL[5] > 0.0) c.FacRej = RCNTRL[5]; else { c.FacRej = 0; Ierr = -4; }
  if (RCNTRL[6] == 0.0) c.FacSafe = (double)0.9f; else if (RCNTRL[6] > 0.0) c.FacSafe = RCNTRL[6]; else { c.FacSafe = 0; Ierr = -4; }
  c.ThetaMin = (RCNTRL[7] == 0.0) ? 1.0e-3 : RCNTRL[7];
  c.NewtonTol = (RCNTRL[8] == 0.0) ? 3.0e-2 : RCNTRL[8];
  c.Qmin = (RCNTRL[9] == 0.0) ? 1.0 : RCNTRL[9];
  c.Qmax = (RCNTRL[10] == 0.0) ? 1.2 : RCNTRL[10];
  if (c.ITOL == 0) {
    if (AbsTol[0] <= 0.0 || RelTol[0] <= 10.0 * c.Roundoff) Ierr = -5;
  } else {
    for (int i = 0; i < N; ++i)
      if (AbsTol[i] <= 0.0 || RelTol[i] <= 10.0 * c.Roundoff) Ierr = -5;
  }
  return Ierr;
}

template <class Model>
struct Sdirk {
  static constexpr int N = Model::N;
  Model& mdl;
  SdirkConfig cfg;
  int ISTATUS[20];
  double RSTATUS[20];
  LssDense<N> lss;
  // hooks of the adjoint family's forward sweep (SDIRK_FwdInt, ADJ/SDIRK_ADJ/SDIRK_ADJ_f90_Integrator.F90:545-781): every
  // accepted step is pushed as (T, H, Y, Z) before the update (SDIRK_Push :735), and its SDIRK_PrepareMatrix re-evaluates
  // the Jacobian after a singular factorisation (SkipJac = .FALSE., :1239; the forward family keeps it, :780)
  struct TapeEntry { double T, H, Y[N], Z[5][N]; };
  std::vector<TapeEntry>* tape = nullptr;
  bool sng_keeps_jac = true;
  explicit Sdirk(Model& m) : mdl(m) { std::memset(ISTATUS, 0, sizeof ISTATUS); std::memset(RSTATUS, 0, sizeof RSTATUS); }

  void error_scale(const double* AbsTol, const double* RelTol, const double* Y, double* SCAL) const {   // :661-676
    for (int i = 0; i < N; ++i)
      SCAL[i] = (cfg.ITOL == 0) ? 1.0 / (AbsTol[0] + RelTol[0] * std::fabs(Y[i])) : 1.0 / (AbsTol[i] + RelTol[i] * std::fabs(Y[i]));
  }
  static double error_norm(const double* Y, const double* SCAL) {   // :680-691
    double Err = 0.0;
    for (int i = 0; i < N; ++i) { double q = Y[i] * SCAL[i]; Err = Err + q * q; }
    return std::fmax(std::sqrt(Err / (double)N), 1.0e-10);
  }
  // SDIRK_PrepareMatrix :736-785
  int prepare_matrix(double& H, double T, const double* Y, bool& SkipJac, bool& SkipLU, bool& Reject) {
    int ConsecutiveSng = 0, ISING = 1;
    while (ISING != 0) {
      double HGammaInv = 1.0 / (H * cfg.tab.Gamma);
      if (!SkipJac) { mdl.jac(T, Y, lss.fjac); ISTATUS[Njac]++; }
      ISING = lss.decomp(HGammaInv);
      ISTATUS[Ndec]++;
      if (ISING != 0) {
        ISTATUS[Nsng]++;
        ConsecutiveSng++;
        if (ConsecutiveSng >= 6) return ISING;
        H = 0.5 * H;
        SkipJac = sng_keeps_jac; SkipLU = false; Reject = true;
      }
    }
    return 0;
  }
  void solve(double H, double* RHS) {   // SDIRK_Solve :789-810
    double HGammaInv = 1.0 / (H * cfg.tab.Gamma);
    for (int i = 0; i < N; ++i) RHS[i] = HGammaInv * RHS[i];
    lss.solve(false, RHS);
    ISTATUS[Nsol]++;
  }

  // SDIRK_Integrator :430-657
  int integrate(double* Y, double Tinitial, double Tfinal, const double* AbsTol, const double* RelTol) {
    const SdirkTableauData& tb = cfg.tab;
    const int S = tb.S;
    double Z[5][N], G[N], TMP[N], SCAL[N], RHS[N];
    double NewtonRate = 0, Theta = 0, Hratio, NewtonPredictedErr, Qnewton, Err, Fac, Hnew, NewtonIncrement = 0, NewtonIncrementOld = 0;
    const double Roundoff = cfg.Roundoff;
    double T = Tinitial;
    const double Tdirection = (Tfinal - Tinitial) >= 0.0 ? 1.0 : -1.0;   // SIGN(ONE, Tfinal-Tinitial)
    double H = std::fmax(std::fabs(cfg.Hmin), std::fabs(cfg.Hstart));
    if (std::fabs(H) <= 10.0 * Roundoff) H = 1.0e-6;
    H = std::fmin(std::fabs(H), cfg.Hmax);
    H = std::copysign(H, Tdirection);
    bool SkipLU = false, SkipJac = false, Reject = false, FirstStep = true;
    error_scale(AbsTol, RelTol, Y, SCAL);
    while ((Tfinal - T) * Tdirection - Roundoff > 0.0) {   // Tloop
      if (!SkipLU) {
        if (prepare_matrix(H, T, Y, SkipJac, SkipLU, Reject) != 0) return -8;
      }
      if (ISTATUS[Nstp] > cfg.Max_no_steps) return -6;
      if ((T + 0.1 * H == T) || (std::fabs(H) <= Roundoff)) return -7;
      bool cycle = false;
      for (int istage = 1; istage <= S; ++istage) {
        double* Zi = Z[istage - 1];
        for (int i = 0; i < N; ++i) Zi[i] = 0.0;
        for (int i = 0; i < N; ++i) G[i] = 0.0;
        if (istage > 1) {
          for (int j = 1; j <= istage - 1; ++j) {
            const double th = tb.Theta[istage - 1][j - 1];
            for (int i = 0; i < N; ++i) G[i] = G[i] + th * Z[j - 1][i];
            if (cfg.StartNewton) {
              const double al = tb.Alpha[istage - 1][j - 1];
              for (int i = 0; i < N; ++i) Zi[i] = Zi[i] + al * Z[j - 1][i];
            }
          }
        }
        bool NewtonDone = false;
        Fac = 0.5;
        for (int NewtonIter = 1; NewtonIter <= cfg.NewtonMaxit; ++NewtonIter) {
          for (int i = 0; i < N; ++i) TMP[i] = Y[i] + Zi[i];                    // WADD
          mdl.fun(T + tb.C[istage - 1] * H, TMP, RHS);
          ISTATUS[Nfun]++;
          { const double a = H * tb.Gamma; for (int i = 0; i < N; ++i) RHS[i] = a * RHS[i]; }   // DSCAL
          for (int i = 0; i < N; ++i) RHS[i] = RHS[i] + (-1.0) * Zi[i];          // DAXPY(-ONE, Z)
          for (int i = 0; i < N; ++i) RHS[i] = RHS[i] + 1.0 * G[i];              // DAXPY(ONE, G)
          solve(H, RHS);
          NewtonIncrement = error_norm(RHS, SCAL);
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
          for (int i = 0; i < N; ++i) Zi[i] = Zi[i] + 1.0 * RHS[i];
          NewtonDone = (NewtonRate * NewtonIncrement <= cfg.NewtonTol);
          if (NewtonDone) break;
        }
        if (!NewtonDone) {
          H = Fac * H; Reject = true;
          SkipJac = true; SkipLU = false;
          cycle = true;
          break;
        }
      }
      if (cycle) continue;
      // error estimation
      ISTATUS[Nstp]++;
      for (int i = 0; i < N; ++i) TMP[i] = 0.0;
      for (int s = 1; s <= S; ++s)
        if (tb.E[s - 1] != 0.0) for (int i = 0; i < N; ++i) TMP[i] = TMP[i] + tb.E[s - 1] * Z[s - 1][i];
      solve(H, TMP);
      Err = error_norm(TMP, SCAL);
      Fac = cfg.FacSafe * o_pow_neg_inv(Err, tb.ELO);
      Fac = std::fmax(cfg.FacMin, std::fmin(cfg.FacMax, Fac));
      Hnew = H * Fac;
      if (Err < 1.0) {
        FirstStep = false;
        ISTATUS[Nacc]++;
        if (tape) {
          tape->emplace_back();
          TapeEntry& e = tape->back();
          e.T = T; e.H = H;
          std::memcpy(e.Y, Y, sizeof e.Y);
          std::memcpy(e.Z, Z, sizeof e.Z);
        }
        T = T + H;
        for (int s = 1; s <= S; ++s)
          if (tb.D[s - 1] != 0.0) for (int i = 0; i < N; ++i) Y[i] = Y[i] + tb.D[s - 1] * Z[s - 1][i];
        error_scale(AbsTol, RelTol, Y, SCAL);
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
          SkipLU = ((Theta <= cfg.ThetaMin) && (Hratio >= cfg.Qmin) && (Hratio <= cfg.Qmax));
          if (!SkipLU) H = Hnew;
        }
        SkipJac = false;
      } else {
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

// INTEGRATE of FWD/SDIRK (:25-90): user values override the defaults only when > 0
template <class Model>
int sdirk_integrate(Model& mdl, double TIN, double TOUT, double* VAR, const double* RTOL, const double* ATOL,
                    const int* ICNTRL_U, const double* RCNTRL_U, int* ISTATUS_U, double* RSTATUS_U) {
  int ICNTRL[20] = {0};
  double RCNTRL[20] = {0};
  if (ICNTRL_U) for (int i = 0; i < 20; ++i) if (ICNTRL_U[i] > 0) ICNTRL[i] = ICNTRL_U[i];
  if (RCNTRL_U) for (int i = 0; i < 20; ++i) if (RCNTRL_U[i] > 0) RCNTRL[i] = RCNTRL_U[i];
  Sdirk<Model> s(mdl);
  int Ierr = sdirk_parse(Model::N, ICNTRL, RCNTRL, TIN, TOUT, ATOL, RTOL, s.cfg);
  if (Ierr >= 0) Ierr = s.integrate(VAR, TIN, TOUT, ATOL, RTOL);
  if (ISTATUS_U) for (int i = 0; i < 20; ++i) ISTATUS_U[i] = s.ISTATUS[i];
  if (RSTATUS_U) for (int i = 0; i < 20; ++i) RSTATUS_U[i] = s.RSTATUS[i];
  return Ierr;
}

}  // namespace oracle
