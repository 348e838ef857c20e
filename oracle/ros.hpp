// ORACLE — TEST INFRASTRUCTURE ONLY (never linked into the product path).
//
// CPU restatement of the reference's Rosenbrock family, FULL_ALGEBRA path:
//   FWD/ROS/ROS_f90_Integrator.F90            INTEGRATE      :32-90,  Rosenbrock :93-387,
//                                             ros_Integrator :433-629
//   ADJ/ROS_ADJ/ROS_ADJ_f90_Integrator.F90    INTEGRATE_ADJ  :34-156, RosenbrockADJ1 :159-613,
//                                             ros_FwdInt :877-1173, ros_DadjInt :1177-1441,
//                                             RosenbrockADJ2 :2486-.. (Lambda-only twin, :3508-3700)
//   ADJ/ROS_ADJ/LS_Solver.F90                 lapack_decomp :31-42, lapack_solve :45-54,
//                                             lss_jac_time_derivative :355-360, lss_mul_* :373-390
// Statement order, BLAS call order (DAXPY/DSCAL/DGEMV loops) and counter updates follow the
// Fortran line by line; compile with -ffp-contract=off.  Re-entrant: all state is local.
//
// Deliberate decisions on reference defects:
//  * `ELSEIF (Max_no_steps > 0)` (ADJ :417) reads an uninitialised variable; restated as the
//    evident intent `ICNTRL(4) > 0`.
//  * quadrature-cost path (Q/QFUN, ros_W) is out of scope (its ros_W weights come from a loop that never runs);
//    the continuous adjoints (ICNTRL(7) = 3, 4) are restated with their quirks (cadjoint, simple_cadjoint below).
//  * SaveLU (ICNTRL(8)) is honoured on read but never implemented upstream; only 0 is accepted.
#pragma once
#include <cmath>
#include <cstring>
#include <vector>
#include "ref_lapack.hpp"
#include "portable_math.hpp"
#include "gen/ros_tableau_gen.h"

namespace oracle {

enum { Nfun = 0, Njac = 1, Nstp = 2, Nacc = 3, Nrej = 4, Ndec = 5, Nsol = 6, Nsng = 7 };  // 0-based
enum { Ntexit = 0, Nhexit = 1, Nhnew = 2 };

enum RosFamily { ROS_FWD = 0, ROS_ADJ = 1, ROS_TLM = 2 };

struct RosConfig {
  RosTableauData tab;
  bool Autonomous, VectorTol;
  int Max_no_steps;
  double Roundoff, Hmin, Hmax, Hstart, FacMin, FacMax, FacRej, FacSafe;
  double DeltaMinStart;   // DeltaMin of the driver routine (start step floor)
  double DeltaMinFD;      // DeltaMin of ros_FunTimeDerivative
  double tenth;           // the 0.1 of the step-too-small test
};

// Option parsing shared by the three Rosenbrock drivers.  Returns IERR (1 = ok so far).
//  FWD: FWD/ROS/..:249-375   ADJ: ADJ/ROS_ADJ/..:383-535   TLM: TLM/ROS_TLM/..:270-400
static inline int ros_parse(RosFamily fam, int N, const int* ICNTRL, const double* RCNTRL, double Tstart,
                            double Tend, const double* AbsTol, const double* RelTol, RosConfig& c) {
  const bool sp = (fam == ROS_FWD);          // FWD/ROS spells its literals in REAL(4)
  c.Autonomous = !(ICNTRL[0] == 0);
  c.VectorTol = (ICNTRL[1] == 0);
  int UplimTol = c.VectorTol ? N : 1;
  int method;
  switch (ICNTRL[2]) {
    case 1: method = 1; break;
    case 2: method = 2; break;
    case 3: method = 3; break;
    case 4: method = 4; break;
    case 5: method = 5; break;
    case 0: method = (fam == ROS_FWD) ? 3 : (fam == ROS_ADJ ? 4 : 5); break;   // fact 5
    default: return -2;
  }
  c.tab = sp ? ROS_TABLEAU_E[method - 1] : ROS_TABLEAU_D[method - 1];
  if (ICNTRL[3] == 0) c.Max_no_steps = (fam == ROS_ADJ) ? 200000 - 1 : 200000;   // bufsize-1 (ADJ :416)
  else if (ICNTRL[3] > 0) c.Max_no_steps = ICNTRL[3];
  else return -1;
  if (fam == ROS_ADJ) {
    if (!(ICNTRL[5] == 0 || (ICNTRL[5] >= 1 && ICNTRL[5] <= 5))) return -2;      // :426-435
    if (!(ICNTRL[6] == 0 || (ICNTRL[6] >= 1 && ICNTRL[6] <= 4))) return -9;      // :438-446
  }
  c.Roundoff = dlamch_eps();
  if (RCNTRL[0] == 0.0) c.Hmin = 0.0; else if (RCNTRL[0] > 0.0) c.Hmin = RCNTRL[0]; else return -3;
  if (RCNTRL[1] == 0.0) c.Hmax = std::fabs(Tend - Tstart);
  else if (RCNTRL[1] > 0.0) c.Hmax = std::fmin(std::fabs(RCNTRL[1]), std::fabs(Tend - Tstart));
  else return -3;
  c.DeltaMinStart = sp ? (double)1.0E-5f : 1.0e-5;
  c.DeltaMinFD = sp ? (double)1.0E-6f : 1.0e-6;
  c.tenth = sp ? (double)0.1f : 0.1;
  if (RCNTRL[2] == 0.0) c.Hstart = std::fmax(c.Hmin, c.DeltaMinStart);
  else if (RCNTRL[2] > 0.0) c.Hstart = std::fmin(std::fabs(RCNTRL[2]), std::fabs(Tend - Tstart));
  else return -3;
  if (RCNTRL[3] == 0.0) c.FacMin = sp ? (double)0.2f : 0.2; else if (RCNTRL[3] > 0.0) c.FacMin = RCNTRL[3]; else return -4;
  if (RCNTRL[4] == 0.0) c.FacMax = 6.0; else if (RCNTRL[4] > 0.0) c.FacMax = RCNTRL[4]; else return -4;
  if (RCNTRL[5] == 0.0) c.FacRej = sp ? (double)0.1f : 0.1; else if (RCNTRL[5] > 0.0) c.FacRej = RCNTRL[5]; else return -4;
  if (RCNTRL[6] == 0.0) c.FacSafe = sp ? (double)0.9f : 0.9; else if (RCNTRL[6] > 0.0) c.FacSafe = RCNTRL[6]; else return -4;
  for (int i = 0; i < UplimTol; ++i)
    if (AbsTol[i] <= 0.0 || RelTol[i] <= 10.0 * c.Roundoff || RelTol[i] >= 1.0) return -5;
  return 1;
}

// ros_ErrorNorm: ADJ :1837-1866 / FWD :633-663
static inline double ros_error_norm(int N, const double* Y, const double* Ynew, const double* Yerr,
                                    const double* AbsTol, const double* RelTol, bool VectorTol) {
  double Err = 0.0;
  for (int i = 0; i < N; ++i) {
    double Ymax = std::fmax(std::fabs(Y[i]), std::fabs(Ynew[i]));
    double Scale = VectorTol ? AbsTol[i] + RelTol[i] * Ymax : AbsTol[0] + RelTol[0] * Ymax;
    double q = Yerr[i] / Scale;
    Err = Err + q * q;
  }
  Err = std::sqrt(Err / N);
  return std::fmax(Err, 1.0e-10);
}

struct RosTapeEntry { double T, H; std::vector<double> Ystage, K; };
struct RosCTapeEntry { double T, H; std::vector<double> Y, dY; };   // ros_CPush :785-805 (chk_d2Y is only read by ros_Hermite5, never reached)

// Per-call linear-solver state (module ls_solver/lapack in the reference)
template <int N>
struct LssDense {
  double fjac[N * N], fjac1[N * N], djdt[N * N], e[N * N];
  int ip[N];
  LssDense() { std::memset(this, 0, sizeof *this); }   // fresh ALLOCATE -> zero pages (fact 16)
  int decomp(double hgamma) {                          // lapack_decomp, LS_Solver.F90:31-42
    for (int j = 0; j < N; ++j) {
      for (int i = 0; i < N; ++i) e[i + N * j] = -fjac[i + N * j];
      e[j + N * j] = e[j + N * j] + hgamma;
    }
    return dgetf2(N, e, ip);
  }
  void solve(bool trans, double* rhs) { dgetrs(trans, N, e, ip, rhs); }    // :45-54
  void mul_jactr(double* z, const double* g) const {   // matmul(transpose(fjac),g), :378-381
    for (int i = 0; i < N; ++i) { double s = 0.0; for (int k = 0; k < N; ++k) s += fjac[k + N * i] * g[k]; z[i] = s; }
  }
  void mul_jac(double* z, const double* g) const {     // matmul(fjac,g): column-sweep as libgfortran
    for (int i = 0; i < N; ++i) z[i] = 0.0;
    for (int k = 0; k < N; ++k) for (int i = 0; i < N; ++i) z[i] += fjac[i + N * k] * g[k];
  }
  void mul_jac1(double* z, const double* g) const {
    for (int i = 0; i < N; ++i) z[i] = 0.0;
    for (int k = 0; k < N; ++k) for (int i = 0; i < N; ++i) z[i] += fjac1[i + N * k] * g[k];
  }
  void mul_djdt(double* z, const double* g) const {
    for (int i = 0; i < N; ++i) z[i] = 0.0;
    for (int k = 0; k < N; ++k) for (int i = 0; i < N; ++i) z[i] += djdt[i + N * k] * g[k];
  }
  void mul_djdttr(double* z, const double* g) const {  // :388-391
    for (int i = 0; i < N; ++i) { double s = 0.0; for (int k = 0; k < N; ++k) s += djdt[k + N * i] * g[k]; z[i] = s; }
  }
  void jac_time_derivative(double delta) {             // :355-360
    double a = -1.0;
    for (int i = 0; i < N * N; ++i) djdt[i] = djdt[i] + a * fjac[i];
    a = 1 / delta;
    for (int i = 0; i < N * N; ++i) djdt[i] = a * djdt[i];
  }
};

template <class Model>
struct Rosenbrock {
  static constexpr int N = Model::N;
  Model& mdl;
  RosConfig cfg;
  int ISTATUS[20];
  double RSTATUS[20];
  LssDense<N> lss;
  std::vector<RosTapeEntry> tape;   // chk_T, chk_H, chk_Y, chk_K
  std::vector<RosCTapeEntry> ctape;  // continuous adjoints: chk_T, chk_H, chk_Y, chk_dY
  explicit Rosenbrock(Model& m) : mdl(m) { std::memset(ISTATUS, 0, sizeof ISTATUS); std::memset(RSTATUS, 0, sizeof RSTATUS); }

  // ros_FunTimeDerivative: ADJ :1870-1890 / FWD :667-688
  void fun_time_derivative(double T, const double* Y, const double* Fcn0, double* dFdT) {
    double Delta = std::sqrt(cfg.Roundoff) * std::fmax(cfg.DeltaMinFD, std::fabs(T));
    mdl.fun(T + Delta, Y, dFdT);
    ISTATUS[Nfun]++;
    for (int i = 0; i < N; ++i) dFdT[i] = dFdT[i] + (-1.0) * Fcn0[i];
    double s = 1.0 / Delta;
    for (int i = 0; i < N; ++i) dFdT[i] = s * dFdT[i];
  }

  // ros_PrepareMatrix: ADJ :1894-1948 (counts Ndec) / FWD :692-751 (does not, fact 6)
  bool prepare_matrix(double& H, int Direction, double gam, bool count_dec) {
    int Nconsecutive = 0;
    bool Singular = true;
    while (Singular) {
      double ghinv = 1.0 / (Direction * H * gam);
      int ising = lss.decomp(ghinv);
      if (count_dec) ISTATUS[Ndec]++;
      if (ising == 0) {
        Singular = false;
      } else {
        ISTATUS[Nsng]++;
        Nconsecutive++;
        Singular = true;
        if (Nconsecutive <= 5) H = H * 0.5; else return true;
      }
    }
    return false;
  }

  // ros_Integrator (FWD :433-629) and ros_FwdInt (ADJ :877-1173); `adj` selects the family quirks:
  // tape push, Ndec/Nsol counting, RSTATUS(Nhexit)=H at loop top (ADJ :963).
  int forward(double* Y, double Tstart, double Tend, const double* AbsTol, const double* RelTol, bool adj,
              bool count_la, bool cadj = false) {   // cadj: AdjointType continuous / simple continuous (checkpoints :1058-1065, 1143-1157)
    const RosTableauData& tb = cfg.tab;
    const int S = tb.S;
    double Ynew[N], Fcn0[N], Fcn[N], dFdT[N], Yerr[N];
    std::vector<double> K(N * S), Ystage(N * S, 0.0);
    double H, Hnew, HC, HG, Fac, Tau, Err;
    int Direction;
    bool RejectLastH = false, RejectMoreH = false;
    const double Roundoff = cfg.Roundoff;

    double T = Tstart;
    RSTATUS[Nhexit] = 0.0;
    H = std::fmin(std::fmax(std::fabs(cfg.Hmin), std::fabs(cfg.Hstart)), std::fabs(cfg.Hmax));
    if (std::fabs(H) <= 10.0 * Roundoff) H = cfg.DeltaMinStart;
    Direction = (Tend >= Tstart) ? +1 : -1;
    H = Direction * H;

    while ((Direction > 0 && (T - Tend) + Roundoff <= 0.0) || (Direction < 0 && (Tend - T) + Roundoff <= 0.0)) {
      if (ISTATUS[Nstp] > cfg.Max_no_steps) return -6;
      if ((T + cfg.tenth * H) == T || H <= Roundoff) return -7;
      if (adj || cadj) RSTATUS[Nhexit] = H;   // ros_FwdInt :963
      H = std::fmin(H, std::fabs(Tend - T));
      mdl.fun(T, Y, Fcn0);
      ISTATUS[Nfun]++;
      if (!cfg.Autonomous) fun_time_derivative(T, Y, Fcn0, dFdT);
      mdl.jac(T, Y, lss.fjac);
      ISTATUS[Njac]++;
      for (;;) {   // UntilAccepted
        if (prepare_matrix(H, Direction, tb.Gamma[0], count_la)) return -8;
        for (int istage = 1; istage <= S; ++istage) {
          int ioffset = N * (istage - 1);
          if (istage == 1) {
            for (int i = 0; i < N; ++i) Fcn[i] = Fcn0[i];
            if (adj) { for (int i = 0; i < N; ++i) { Ystage[i] = Y[i]; Ynew[i] = Y[i]; } }
          } else if (tb.NewF[istage - 1]) {
            for (int i = 0; i < N; ++i) Ynew[i] = Y[i];
            for (int j = 1; j <= istage - 1; ++j) {
              double a = tb.A[(istage - 1) * (istage - 2) / 2 + j - 1];
              for (int i = 0; i < N; ++i) Ynew[i] = Ynew[i] + a * K[N * (j - 1) + i];   // DAXPY
            }
            Tau = T + tb.Alpha[istage - 1] * Direction * H;
            mdl.fun(Tau, Ynew, Fcn);
            ISTATUS[Nfun]++;
          }
          if (istage > 1 && adj) for (int i = 0; i < N; ++i) Ystage[ioffset + i] = Ynew[i];
          for (int i = 0; i < N; ++i) K[ioffset + i] = Fcn[i];
          for (int j = 1; j <= istage - 1; ++j) {
            HC = tb.C[(istage - 1) * (istage - 2) / 2 + j - 1] / (Direction * H);
            for (int i = 0; i < N; ++i) K[ioffset + i] = K[ioffset + i] + HC * K[N * (j - 1) + i];
          }
          if (!cfg.Autonomous && tb.Gamma[istage - 1] != 0.0) {
            HG = Direction * H * tb.Gamma[istage - 1];
            for (int i = 0; i < N; ++i) K[ioffset + i] = K[ioffset + i] + HG * dFdT[i];
          }
          lss.solve(false, &K[ioffset]);
          if (count_la) ISTATUS[Nsol]++;
        }
        for (int i = 0; i < N; ++i) Ynew[i] = Y[i];
        for (int j = 1; j <= S; ++j) {
          double m = tb.M[j - 1];
          for (int i = 0; i < N; ++i) Ynew[i] = Ynew[i] + m * K[N * (j - 1) + i];
        }
        for (int i = 0; i < N; ++i) Yerr[i] = 0.0;
        for (int j = 1; j <= S; ++j) {
          double e = tb.E[j - 1];
          for (int i = 0; i < N; ++i) Yerr[i] = Yerr[i] + e * K[N * (j - 1) + i];
        }
        Err = ros_error_norm(N, Y, Ynew, Yerr, AbsTol, RelTol, cfg.VectorTol);
        Fac = std::fmin(cfg.FacMax, std::fmax(cfg.FacMin, cfg.FacSafe / o_root(Err, tb.ELO)));
        Hnew = H * Fac;
        ISTATUS[Nstp]++;
        if (Err <= 1.0 || H <= cfg.Hmin) {
          ISTATUS[Nacc]++;
          if (adj) { tape.push_back(RosTapeEntry{T, H, Ystage, K}); }   // ros_DPush :735-755
          if (cadj) ctape.push_back(RosCTapeEntry{T, H, std::vector<double>(Y, Y + N), std::vector<double>(Fcn0, Fcn0 + N)});
          for (int i = 0; i < N; ++i) Y[i] = Ynew[i];
          T = T + Direction * H;
          Hnew = std::fmax(cfg.Hmin, std::fmin(Hnew, cfg.Hmax));
          if (RejectLastH) Hnew = std::fmin(Hnew, H);
          RSTATUS[Nhexit] = H;
          RSTATUS[Nhnew] = Hnew;
          RSTATUS[Ntexit] = T;
          RejectLastH = false;
          RejectMoreH = false;
          H = Hnew;
          break;
        } else {
          if (RejectMoreH) Hnew = H * cfg.FacRej;
          RejectMoreH = RejectLastH;
          RejectLastH = true;
          H = Hnew;
          if (ISTATUS[Nacc] >= 1) ISTATUS[Nrej]++;
        }
      }
    }
    if (cadj) {   // "Save last state: only needed for continuous adjoint" (:1143-1157)
      mdl.fun(T, Y, Fcn0);
      ISTATUS[Nfun]++;
      mdl.jac(T, Y, lss.fjac);
      ISTATUS[Njac]++;
      if (!cfg.Autonomous) fun_time_derivative(T, Y, Fcn0, dFdT);   // feeds chk_d2Y only; its FUN call is counted
      ctape.push_back(RosCTapeEntry{T, H, std::vector<double>(Y, Y + N), std::vector<double>(Fcn0, Fcn0 + N)});
    }
    return 1;
  }

  // ros_CadjInt :1447-1671, called as ros_CadjInt(NADJ, Lambda, Tend, Tstart, ...) (:588-592): the adjoint runs backwards in
  // time, so Direction = -1 and H = Direction*H is negative — and the first pass through the time loop stops at
  // `IF ( ((T+0.1d0*H) == T).OR.(H <= Roundoff) )` (:1517) with IERR = -7 before anything is computed.  The continuous adjoint of
  // the reference therefore returns ADJINIT's Lambda and "Step size too small" for every forward-in-time problem; restated as is.
  int cadjoint(double Tstart, double Tend) {
    RSTATUS[Nhexit] = 0.0;
    double H = std::fmin(std::fmax(std::fabs(cfg.Hmin), std::fabs(cfg.Hstart)), std::fabs(cfg.Hmax));
    if (std::fabs(H) <= 10.0 * cfg.Roundoff) H = 1.0e-5;
    const int Direction = (Tend >= Tstart) ? +1 : -1;
    H = Direction * H;
    const double T = Tstart;
    if (!((Direction > 0 && (T - Tend) + cfg.Roundoff <= 0.0) || (Direction < 0 && (Tend - T) + cfg.Roundoff <= 0.0))) return 1;
    if (ISTATUS[Nstp] > cfg.Max_no_steps) return -6;
    if ((T + 0.1 * H) == T || H <= cfg.Roundoff) return -7;
    return -100;   // a backward-in-time forward problem (Direction = +1 here): not restated
  }

  // ros_SimpleCadjInt :1675-1834: the adjoint ODE on the forward sweep's step sequence, no error control, with the method
  // CadjMethod selected by ICNTRL(6).  Restated with its quirks: the first stage multiplies by the (negated) Jacobian itself,
  // LSS_Mul_Jac (:1741), the later stages by its transpose (:1781); the stage Jacobians are taken at Hermite-interpolated states.
  int simple_cadjoint(int NADJ, double* Lambda, double Tstart, double Tend) {
    const RosTableauData& tb = cfg.tab;
    const int S = tb.S;
    const int Direction = (Tend >= Tstart) ? -1 : +1;
    std::vector<double> Ynew((size_t)N * NADJ), Fcn0((size_t)N * NADJ), Fcn((size_t)N * NADJ), K((size_t)N * S * NADJ), dFdT((size_t)N * NADJ);
    double Y0[N];
    const int LD = N * S;
    for (int istack = (int)ctape.size(); istack >= 2; --istack) {
      const RosCTapeEntry& hi = ctape[istack - 1];
      const RosCTapeEntry& lo = ctape[istack - 2];
      const double T = hi.T, H = lo.H;
      for (int i = 0; i < N; ++i) Y0[i] = hi.Y[i];
      mdl.jac(T, Y0, lss.fjac);
      ISTATUS[Njac]++;
      if (!cfg.Autonomous) {
        const double Delta = std::sqrt(cfg.Roundoff) * std::fmax(1.0e-5, std::fabs(T));
        mdl.jac(T + Delta, Y0, lss.djdt);
        ISTATUS[Njac]++;
        lss.jac_time_derivative(Delta);
        for (int m = 0; m < NADJ; ++m) {
          lss.mul_djdttr(&dFdT[(size_t)N * m], Lambda + (size_t)N * m);
          for (int i = 0; i < N; ++i) dFdT[(size_t)N * m + i] = (-1.0) * dFdT[(size_t)N * m + i];
        }
      }
      for (int i = 0; i < N * N; ++i) lss.fjac[i] = -lss.fjac[i];                    // LSS_Rev_Jac
      for (int m = 0; m < NADJ; ++m) lss.mul_jac(&Fcn0[(size_t)N * m], Lambda + (size_t)N * m);
      const double ghinv = 1.0 / (Direction * H * tb.Gamma[0]);
      const int j = lss.decomp(ghinv);
      ISTATUS[Ndec]++;
      if (j != 0) return -8;                                                            // the reference STOPs
      for (int istage = 1; istage <= S; ++istage) {
        const int ioffset = N * (istage - 1);
        if (istage == 1) {
          for (size_t q = 0; q < (size_t)N * NADJ; ++q) Fcn[q] = Fcn0[q];
        } else if (tb.NewF[istage - 1]) {
          for (size_t q = 0; q < (size_t)N * NADJ; ++q) Ynew[q] = Lambda[q];
          for (int jj = 1; jj <= istage - 1; ++jj) {
            const double a = tb.A[(istage - 1) * (istage - 2) / 2 + jj - 1];
            for (int m = 0; m < NADJ; ++m)
              for (int i = 0; i < N; ++i) Ynew[(size_t)N * m + i] = Ynew[(size_t)N * m + i] + a * K[(size_t)LD * m + N * (jj - 1) + i];
          }
          const double Tau = T + tb.Alpha[istage - 1] * Direction * H;
          hermite3(lo.T, hi.T, Tau, lo.Y.data(), hi.Y.data(), lo.dY.data(), hi.dY.data(), Y0);
          mdl.jac(Tau, Y0, lss.fjac1);
          ISTATUS[Njac]++;
          for (int i = 0; i < N * N; ++i) lss.fjac1[i] = -lss.fjac1[i];                // LSS_Rev_Jac1
          for (int m = 0; m < NADJ; ++m) {                                            // LSS_Mul_Jactr1
            const double* g = &Ynew[(size_t)N * m];
            double* z = &Fcn[(size_t)N * m];
            for (int i = 0; i < N; ++i) { double sm = 0.0; for (int k = 0; k < N; ++k) sm += lss.fjac1[k + N * i] * g[k]; z[i] = sm; }
          }
        }
        for (int m = 0; m < NADJ; ++m)
          for (int i = 0; i < N; ++i) K[(size_t)LD * m + ioffset + i] = Fcn[(size_t)N * m + i];
        for (int jj = 1; jj <= istage - 1; ++jj) {
          const double HC = tb.C[(istage - 1) * (istage - 2) / 2 + jj - 1] / (Direction * H);
          for (int m = 0; m < NADJ; ++m)
            for (int i = 0; i < N; ++i) K[(size_t)LD * m + ioffset + i] = K[(size_t)LD * m + ioffset + i] + HC * K[(size_t)LD * m + N * (jj - 1) + i];
        }
        if (!cfg.Autonomous && tb.Gamma[istage - 1] != 0.0) {
          const double HG = Direction * H * tb.Gamma[istage - 1];
          for (int m = 0; m < NADJ; ++m)
            for (int i = 0; i < N; ++i) K[(size_t)LD * m + ioffset + i] = K[(size_t)LD * m + ioffset + i] + HG * dFdT[(size_t)N * m + i];
        }
        for (int m = 0; m < NADJ; ++m) {
          lss.solve(true, &K[(size_t)LD * m + ioffset]);
          ISTATUS[Nsol]++;
        }
      }
      for (int m = 0; m < NADJ; ++m)
        for (int jj = 1; jj <= S; ++jj) {
          const double mj = tb.M[jj - 1];
          for (int i = 0; i < N; ++i) Lambda[(size_t)N * m + i] = Lambda[(size_t)N * m + i] + mj * K[(size_t)LD * m + N * (jj - 1) + i];
        }
    }
    return 1;
  }

  // ros_Hermite3 :1999-2042: cubic Hermite interpolant, coefficients assembled with DCOPY/DSCAL/DAXPY in the reference's order
  static void hermite3(double a, double b, double T, const double* Ya, const double* Yb, const double* Ja, const double* Jb, double* Y) {
    double amb[3];
    amb[0] = 1.0 / (a - b);
    for (int i = 1; i < 3; ++i) amb[i] = amb[i - 1] * amb[0];
    double C3[N], C4[N];
    // REAL(4) literals times DOUBLE PRECISION: the literal is converted exactly, the product is double
    const double m3 = -3.0 * amb[1], p3 = 3.0 * amb[1], p2a = 2.0 * amb[0];
    for (int i = 0; i < N; ++i) C3[i] = Ya[i];
    for (int i = 0; i < N; ++i) C3[i] = m3 * C3[i];
    for (int i = 0; i < N; ++i) C3[i] = C3[i] + p3 * Yb[i];
    for (int i = 0; i < N; ++i) C3[i] = C3[i] + p2a * Ja[i];
    for (int i = 0; i < N; ++i) C3[i] = C3[i] + amb[0] * Jb[i];
    const double m2 = -2.0 * amb[2], p2 = 2.0 * amb[2];
    for (int i = 0; i < N; ++i) C4[i] = Ya[i];
    for (int i = 0; i < N; ++i) C4[i] = m2 * C4[i];
    for (int i = 0; i < N; ++i) C4[i] = C4[i] + p2 * Yb[i];
    for (int i = 0; i < N; ++i) C4[i] = C4[i] + amb[1] * Ja[i];
    for (int i = 0; i < N; ++i) C4[i] = C4[i] + amb[1] * Jb[i];
    const double Tau = T - a;
    const double t3 = Tau * Tau * Tau, t2 = Tau * Tau;     // Tau**3, TAU**(j-1) with integer exponents: repeated products
    for (int i = 0; i < N; ++i) Y[i] = C4[i];
    for (int i = 0; i < N; ++i) Y[i] = t3 * Y[i];
    for (int i = 0; i < N; ++i) Y[i] = Y[i] + t2 * C3[i];
    for (int i = 0; i < N; ++i) Y[i] = Y[i] + Tau * Ja[i];
    for (int i = 0; i < N; ++i) Y[i] = Y[i] + 1.0 * Ya[i];
  }

  // ros_DadjInt: ADJ1 :1177-1441 (with Mu) / ADJ2 :3508-3700 (Lambda only), GetQuad = false
  int adjoint(int NADJ, int NP, double* Lambda, double* Mu, double Tstart, double Tend) {
    const RosTableauData& tb = cfg.tab;
    const int S = tb.S;
    const bool with_mu = (Mu != nullptr);
    std::vector<double> FPJAC0(with_mu ? N * NP : 0, 0.0), FPJAC1(with_mu ? N * NP : 0, 0.0);
    std::vector<double> U((size_t)N * S * NADJ), V((size_t)N * S * NADJ);
    std::vector<double> Tmp3(NP > 0 ? NP : 1);
    double Tmp[N], Tmp2[N];
    const int Direction = (Tend >= Tstart) ? +1 : -1;
    const int LD = N * S;
    const double DeltaMin = 1.0e-5;   // :1212
    while (!tape.empty()) {
      RosTapeEntry ent = std::move(tape.back());   // ros_DPop :759-782
      tape.pop_back();
      const double T = ent.T, H = ent.H;
      const double* Ystage = ent.Ystage.data();
      const double* K = ent.K.data();
      ISTATUS[Nstp]++;
      mdl.jac(T, Ystage, lss.fjac);
      ISTATUS[Njac]++;
      double Tau = 1.0 / (Direction * H * tb.Gamma[0]);
      ISTATUS[Njac]++;                              // sic (:1264)
      lss.decomp(Tau);
      ISTATUS[Ndec]++;
      for (int istage = S; istage >= 1; --istage) {
        int istart = N * (istage - 1);
        for (int m = 0; m < NADJ; ++m) {
          double* u = &U[istart + (size_t)LD * m];
          for (int i = 0; i < N; ++i) u[i] = Lambda[i + N * m];
          for (int i = 0; i < N; ++i) u[i] = tb.M[istage - 1] * u[i];
        }
        for (int j = istage + 1; j <= S; ++j) {
          int jstart = N * (j - 1);
          double HA = tb.A[(j - 1) * (j - 2) / 2 + istage - 1];
          double HC = tb.C[(j - 1) * (j - 2) / 2 + istage - 1] / (Direction * H);
          for (int m = 0; m < NADJ; ++m) {
            double* u = &U[istart + (size_t)LD * m];
            const double* vj = &V[jstart + (size_t)LD * m];
            const double* uj = &U[jstart + (size_t)LD * m];
            for (int i = 0; i < N; ++i) u[i] = u[i] + HA * vj[i];
            for (int i = 0; i < N; ++i) u[i] = u[i] + HC * uj[i];
          }
        }
        for (int m = 0; m < NADJ; ++m) {
          lss.solve(true, &U[istart + (size_t)LD * m]);
          ISTATUS[Nsol]++;
        }
        Tau = T + tb.Alpha[istage - 1] * Direction * H;
        mdl.jac(Tau, &Ystage[istart], lss.fjac);
        ISTATUS[Njac]++;
        for (int m = 0; m < NADJ; ++m) lss.mul_jactr(&V[istart + (size_t)LD * m], &U[istart + (size_t)LD * m]);
      }
      // Lambda (:1324-1339)
      for (int istage = 1; istage <= S; ++istage) {
        int istart = N * (istage - 1);
        for (int m = 0; m < NADJ; ++m) {
          double* lam = &Lambda[N * m];
          const double* v = &V[istart + (size_t)LD * m];
          for (int i = 0; i < N; ++i) lam[i] = lam[i] + 1.0 * v[i];
          mdl.hesstr_vec(T, Ystage, &U[istart + (size_t)LD * m], &K[istart], Tmp);
          for (int i = 0; i < N; ++i) lam[i] = lam[i] + 1.0 * Tmp[i];
        }
      }
      // Mu (:1342-1372)
      if (with_mu) {
        for (int istage = 1; istage <= S; ++istage) {
          int istart = N * (istage - 1);
          Tau = T + tb.Alpha[istage - 1] * Direction * H;
          mdl.jacp(Tau, &Ystage[istart], FPJAC1.data());
          for (int m = 0; m < NADJ; ++m) {
            const double* u = &U[istart + (size_t)LD * m];
            double* mu = &Mu[NP * m];
            for (int p = 0; p < NP; ++p) Tmp3[p] = 0.0;
            // DGEMV('T',N,NP,ONE,FPJAC1,N,u,1,Beta=1,Tmp3,1): y(j) += alpha*sum_i a(i,j)x(i)
            for (int p = 0; p < NP; ++p) {
              double temp = 0.0;
              for (int i = 0; i < N; ++i) temp = temp + FPJAC1[i + N * p] * u[i];
              Tmp3[p] = Tmp3[p] + 1.0 * temp;
            }
            for (int p = 0; p < NP; ++p) mu[p] = mu[p] + 1.0 * Tmp3[p];
            mdl.hesstr_vec_f_py(T, Ystage, u, &K[istart], Tmp3.data());
            for (int p = 0; p < NP; ++p) mu[p] = mu[p] + 1.0 * Tmp3[p];
          }
        }
      }
      // non-autonomous terms (:1377-1406)
      if (!cfg.Autonomous) {
        double Delta = std::sqrt(cfg.Roundoff) * std::fmax(DeltaMin, std::fabs(T));
        mdl.jac(T + Delta, Ystage, lss.djdt);
        ISTATUS[Njac]++;
        lss.jac_time_derivative(Delta);            // relies on fjac == J(T,Y0) (fact 12)
        if (with_mu) {
          mdl.jacp(T, Ystage, FPJAC0.data());
          mdl.jacp(T + Delta, Ystage, FPJAC1.data());
          double Alpha = -1.0;
          for (int i = 0; i < N * NP; ++i) FPJAC1[i] = FPJAC1[i] + Alpha * FPJAC0[i];
          Alpha = 1 / Delta;
          for (int i = 0; i < N * NP; ++i) FPJAC1[i] = Alpha * FPJAC1[i];
        }
        for (int m = 0; m < NADJ; ++m) {
          for (int i = 0; i < N; ++i) Tmp[i] = 0.0;
          for (int istage = 1; istage <= S; ++istage) {
            const double* u = &U[N * (istage - 1) + (size_t)LD * m];
            double g = tb.Gamma[istage - 1];
            if (g != 0.0)                           // reference DAXPY returns early when da == 0
              for (int i = 0; i < N; ++i) Tmp[i] = Tmp[i] + g * u[i];
          }
          if (with_mu) {
            double* mu = &Mu[NP * m];
            double Alpha = H;
            for (int p = 0; p < NP; ++p) {
              double temp = 0.0;
              for (int i = 0; i < N; ++i) temp = temp + FPJAC1[i + N * p] * Tmp[i];
              mu[p] = mu[p] + Alpha * temp;
            }
          }
          lss.mul_djdttr(Tmp2, Tmp);
          double* lam = &Lambda[N * m];
          for (int i = 0; i < N; ++i) lam[i] = lam[i] + H * Tmp2[i];
        }
      }
    }
    return 1;
  }

  // ros_TLM_Int: TLM/ROS_TLM/ROS_TLM_f90_Integrator.F90:462-707 (forward + tangent linear model in one sweep).
  // Y_tlm(N,NTLM) column-major.  TLMtruncErr: ICNTRL(12) != 0 (ros_ErrorNorm_tlm :745-769).
  int tlm(double* Y, int NTLM, double* Y_tlm, double Tstart, double Tend, const double* AbsTol, const double* RelTol,
          const double* AbsTol_tlm, const double* RelTol_tlm, bool TLMtruncErr) {
    const RosTableauData& tb = cfg.tab;
    const int S = tb.S;
    double Ynew[N], Fcn0[N], Fcn[N], dFdT[N], Yerr[N], Tmp[N];
    std::vector<double> K(N * S), Ynew_tlm((size_t)N * NTLM), Fcn0_tlm((size_t)N * NTLM), Fcn_tlm((size_t)N * NTLM),
        K_tlm((size_t)N * S * NTLM), Yerr_tlm((size_t)N * NTLM);
    double H, Hnew, HC, HG, Fac, Tau, Err;
    const double DeltaMin = 1.0e-5;   // :500
    bool RejectLastH = false, RejectMoreH = false;
    const double Roundoff = cfg.Roundoff;
    const int LD = N * S;
    double T = Tstart;
    RSTATUS[Nhexit] = 0.0;
    H = std::fmin(std::fmax(std::fabs(cfg.Hmin), std::fabs(cfg.Hstart)), std::fabs(cfg.Hmax));
    if (std::fabs(H) <= 10.0 * Roundoff) H = DeltaMin;
    const int Direction = (Tend >= Tstart) ? +1 : -1;
    H = Direction * H;
    while ((Direction > 0 && (T - Tend) + Roundoff <= 0.0) || (Direction < 0 && (Tend - T) + Roundoff <= 0.0)) {
      if (ISTATUS[Nstp] > cfg.Max_no_steps) return -6;
      if ((T + 0.1 * H) == T || H <= Roundoff) return -7;
      H = std::fmin(H, std::fabs(Tend - T));
      mdl.fun(T, Y, Fcn0);
      ISTATUS[Nfun]++;
      mdl.jac(T, Y, lss.fjac);                                      // LSS_Jac :543
      ISTATUS[Njac]++;
      for (int m = 0; m < NTLM; ++m) lss.mul_jac(&Fcn0_tlm[(size_t)N * m], &Y_tlm[(size_t)N * m]);   // :551-553
      if (!cfg.Autonomous) {
        fun_time_derivative(T, Y, Fcn0, dFdT);                       // :557
        double Delta = std::sqrt(Roundoff) * std::fmax(DeltaMin, std::fabs(T));
        mdl.jac(T + Delta, Y, lss.djdt);                             // LSS_Jac2 :559
        ISTATUS[Njac]++;
        lss.jac_time_derivative(Delta);                              // :561
      }
      for (;;) {   // UntilAccepted
        if (prepare_matrix(H, Direction, tb.Gamma[0], true)) return -8;
        for (int istage = 1; istage <= S; ++istage) {
          const int ioffset = N * (istage - 1);
          for (int i = 0; i < N; ++i) Ynew[i] = Y[i];
          for (size_t i = 0; i < (size_t)N * NTLM; ++i) Ynew_tlm[i] = Y_tlm[i];
          if (istage == 1) {
            for (int i = 0; i < N; ++i) Fcn[i] = Fcn0[i];
            for (size_t i = 0; i < (size_t)N * NTLM; ++i) Fcn_tlm[i] = Fcn0_tlm[i];
          } else if (tb.NewF[istage - 1]) {
            for (int j = 1; j <= istage - 1; ++j) {
              const double a = tb.A[(istage - 1) * (istage - 2) / 2 + j - 1];
              for (int i = 0; i < N; ++i) Ynew[i] = Ynew[i] + a * K[N * (j - 1) + i];
              for (int m = 0; m < NTLM; ++m)
                for (int i = 0; i < N; ++i)
                  Ynew_tlm[(size_t)N * m + i] = Ynew_tlm[(size_t)N * m + i] + a * K_tlm[(size_t)LD * m + N * (j - 1) + i];
            }
            Tau = T + tb.Alpha[istage - 1] * Direction * H;
            mdl.fun(Tau, Ynew, Fcn);
            ISTATUS[Nfun]++;
            mdl.jac(Tau, Ynew, lss.fjac1);                           // LSS_Jac1 :600
            ISTATUS[Njac]++;
            for (int m = 0; m < NTLM; ++m) lss.mul_jac1(&Fcn_tlm[(size_t)N * m], &Ynew_tlm[(size_t)N * m]);
          }
          for (int i = 0; i < N; ++i) K[ioffset + i] = Fcn[i];
          for (int m = 0; m < NTLM; ++m)
            for (int i = 0; i < N; ++i) K_tlm[(size_t)LD * m + ioffset + i] = Fcn_tlm[(size_t)N * m + i];
          for (int j = 1; j <= istage - 1; ++j) {
            HC = tb.C[(istage - 1) * (istage - 2) / 2 + j - 1] / (Direction * H);
            for (int i = 0; i < N; ++i) K[ioffset + i] = K[ioffset + i] + HC * K[N * (j - 1) + i];
            for (int m = 0; m < NTLM; ++m)
              for (int i = 0; i < N; ++i)
                K_tlm[(size_t)LD * m + ioffset + i] = K_tlm[(size_t)LD * m + ioffset + i] + HC * K_tlm[(size_t)LD * m + N * (j - 1) + i];
          }
          if (!cfg.Autonomous && tb.Gamma[istage - 1] != 0.0) {
            HG = Direction * H * tb.Gamma[istage - 1];
            for (int i = 0; i < N; ++i) K[ioffset + i] = K[ioffset + i] + HG * dFdT[i];
            for (int m = 0; m < NTLM; ++m) {
              lss.mul_djdt(Tmp, &Ynew_tlm[(size_t)N * m]);             // LSS_Mul_Jac2 :620
              for (int i = 0; i < N; ++i) K_tlm[(size_t)LD * m + ioffset + i] = K_tlm[(size_t)LD * m + ioffset + i] + HG * Tmp[i];
            }
          }
          lss.solve(false, &K[ioffset]);
          ISTATUS[Nsol]++;
          for (int m = 0; m < NTLM; ++m) {
            mdl.hess_vec(T, Y, &K[ioffset], &Y_tlm[(size_t)N * m], Tmp);   // :628
            for (int i = 0; i < N; ++i) K_tlm[(size_t)LD * m + ioffset + i] = K_tlm[(size_t)LD * m + ioffset + i] + 1.0 * Tmp[i];
            lss.solve(false, &K_tlm[(size_t)LD * m + ioffset]);
            ISTATUS[Nsol]++;
          }
        }
        for (int i = 0; i < N; ++i) Ynew[i] = Y[i];
        for (int j = 1; j <= S; ++j)
          for (int i = 0; i < N; ++i) Ynew[i] = Ynew[i] + tb.M[j - 1] * K[N * (j - 1) + i];
        for (int m = 0; m < NTLM; ++m) {
          for (int i = 0; i < N; ++i) Ynew_tlm[(size_t)N * m + i] = Y_tlm[(size_t)N * m + i];
          for (int j = 1; j <= S; ++j)
            for (int i = 0; i < N; ++i)
              Ynew_tlm[(size_t)N * m + i] = Ynew_tlm[(size_t)N * m + i] + tb.M[j - 1] * K_tlm[(size_t)LD * m + N * (j - 1) + i];
        }
        for (int i = 0; i < N; ++i) Yerr[i] = 0.0;
        for (int j = 1; j <= S; ++j)
          for (int i = 0; i < N; ++i) Yerr[i] = Yerr[i] + tb.E[j - 1] * K[N * (j - 1) + i];
        Err = ros_error_norm(N, Y, Ynew, Yerr, AbsTol, RelTol, cfg.VectorTol);
        if (TLMtruncErr) {
          for (size_t i = 0; i < (size_t)N * NTLM; ++i) Yerr_tlm[i] = 0.0;
          for (int m = 0; m < NTLM; ++m)
            for (int j = 1; j <= S; ++j)
              for (int i = 0; i < N; ++i)
                Yerr_tlm[(size_t)N * m + i] = Yerr_tlm[(size_t)N * m + i] + tb.E[j - 1] * K_tlm[(size_t)LD * m + N * (j - 1) + i];
          for (int m = 0; m < NTLM; ++m) {                            // ros_ErrorNorm_tlm :745-769
            const double t = ros_error_norm(N, &Y_tlm[(size_t)N * m], &Ynew_tlm[(size_t)N * m], &Yerr_tlm[(size_t)N * m],
                                            &AbsTol_tlm[(size_t)N * m], &RelTol_tlm[(size_t)N * m], cfg.VectorTol);
            Err = std::fmax(Err, t);
          }
          Err = std::fmax(Err, 1.0e-10);
        }
        Fac = std::fmin(cfg.FacMax, std::fmax(cfg.FacMin, cfg.FacSafe / o_root(Err, tb.ELO)));
        Hnew = H * Fac;
        ISTATUS[Nstp]++;
        if (Err <= 1.0 || H <= cfg.Hmin) {
          ISTATUS[Nacc]++;
          for (int i = 0; i < N; ++i) Y[i] = Ynew[i];
          for (size_t i = 0; i < (size_t)N * NTLM; ++i) Y_tlm[i] = Ynew_tlm[i];
          T = T + Direction * H;
          Hnew = std::fmax(cfg.Hmin, std::fmin(Hnew, cfg.Hmax));
          if (RejectLastH) Hnew = std::fmin(Hnew, H);
          RSTATUS[Nhexit] = H;
          RSTATUS[Nhnew] = Hnew;
          RSTATUS[Ntexit] = T;
          RejectLastH = false;
          RejectMoreH = false;
          H = Hnew;
          break;
        } else {
          if (RejectMoreH) Hnew = H * cfg.FacRej;
          RejectMoreH = RejectLastH;
          RejectLastH = true;
          H = Hnew;
          if (ISTATUS[Nacc] >= 1) ISTATUS[Nrej]++;
        }
      }
    }
    return 1;
  }
};

// ---------------------------------------------------------------------------
// Public entry points mirroring the reference API (arrays 0-based; ICNTRL_U/RCNTRL_U may be null)

// INTEGRATE of FWD/ROS (:32-90): user values override defaults only when > 0
template <class Model>
int ros_integrate(Model& mdl, double TIN, double TOUT, double* VAR, const double* RTOL, const double* ATOL,
                  const int* ICNTRL_U, const double* RCNTRL_U, int* ISTATUS_U, double* RSTATUS_U) {
  int ICNTRL[20] = {0};
  double RCNTRL[20] = {0};
  if (ICNTRL_U) for (int i = 0; i < 20; ++i) if (ICNTRL_U[i] > 0) ICNTRL[i] = ICNTRL_U[i];
  if (RCNTRL_U) for (int i = 0; i < 20; ++i) if (RCNTRL_U[i] > 0) RCNTRL[i] = RCNTRL_U[i];
  Rosenbrock<Model> r(mdl);
  int IERR = ros_parse(ROS_FWD, Model::N, ICNTRL, RCNTRL, TIN, TOUT, ATOL, RTOL, r.cfg);
  if (IERR == 1) IERR = r.forward(VAR, TIN, TOUT, ATOL, RTOL, /*adj=*/false, /*count_la=*/false);
  if (ISTATUS_U) for (int i = 0; i < 20; ++i) ISTATUS_U[i] = r.ISTATUS[i];
  if (RSTATUS_U) for (int i = 0; i < 20; ++i) RSTATUS_U[i] = r.RSTATUS[i];
  return IERR;
}

// INTEGRATE_ADJ of ADJ/ROS_ADJ (:34-156): user values override when >= 0; Mu == nullptr or NP == 0
// selects RosenbrockADJ2 (Lambda only).  Discrete adjoint (ICNTRL(7) in {0,2}) or none (1).
template <class Model>
int ros_integrate_adj(Model& mdl, int NP, int NADJ, double* Y, double* Lambda, double* Mu, double TIN, double TOUT,
                      const double* ATOL, const double* RTOL, const int* ICNTRL_U, const double* RCNTRL_U,
                      int* ISTATUS_U, double* RSTATUS_U) {
  int ICNTRL[20] = {0};
  double RCNTRL[20] = {0};
  if (ICNTRL_U) for (int i = 0; i < 20; ++i) if (ICNTRL_U[i] >= 0) ICNTRL[i] = ICNTRL_U[i];
  if (RCNTRL_U) for (int i = 0; i < 20; ++i) if (RCNTRL_U[i] >= 0) RCNTRL[i] = RCNTRL_U[i];
  Rosenbrock<Model> r(mdl);
  int IERR = ros_parse(ROS_ADJ, Model::N, ICNTRL, RCNTRL, TIN, TOUT, ATOL, RTOL, r.cfg);
  const int AdjointType = (ICNTRL[6] == 0) ? 2 : ICNTRL[6];
  const bool cadj = (AdjointType == 3 || AdjointType == 4);
  if (IERR == 1 && ICNTRL[7] != 0) IERR = -9;                            // SaveLU never implemented upstream
  const bool with_mu = (NP > 0 && Mu != nullptr);
  if (IERR == 1) IERR = r.forward(Y, TIN, TOUT, ATOL, RTOL, /*adj=*/AdjointType == 2, /*count_la=*/true, cadj);
  if (IERR == 1) {
    if (cadj) {   // the method that integrates the continuous adjoint: ICNTRL(6), default Ros4 (:426-435, 559-577)
      const int CadjMethod = (ICNTRL[5] == 0) ? 3 : ICNTRL[5];
      r.cfg.tab = ROS_TABLEAU_D[CadjMethod - 1];
    }
    mdl.adjinit(NADJ, TOUT, Y, Lambda, with_mu ? Mu : nullptr);          // :579
    if (AdjointType == 2) IERR = r.adjoint(NADJ, NP, Lambda, with_mu ? Mu : nullptr, TIN, TOUT);
    else if (AdjointType == 3) IERR = r.cadjoint(TOUT, TIN);             // called with (Tend, Tstart) (:588-592)
    else if (AdjointType == 4) IERR = r.simple_cadjoint(NADJ, Lambda, TIN, TOUT);
  }
  if (ISTATUS_U) for (int i = 0; i < 20; ++i) ISTATUS_U[i] = r.ISTATUS[i];
  if (RSTATUS_U) for (int i = 0; i < 20; ++i) RSTATUS_U[i] = r.RSTATUS[i];
  return IERR;
}

// INTEGRATE_TLM of TLM/ROS_TLM (:35-98): presets ICNTRL(1)=0, (2)=1, (3)=5, (12)=1, RCNTRL(3)=STEPMIN (a SAVEd
// variable that carries the last step size over to the NEXT call, :63,74,93; restated as its initial value 0: every
// system of a batch is a first call); user values override when >= 0.  ROS_TLM has no IERR_U; the code is returned.
template <class Model>
int ros_integrate_tlm(Model& mdl, int NTLM, double* Y, double* Y_tlm, double TIN, double TOUT, const double* ATOL_tlm,
                      const double* RTOL_tlm, const double* ATOL, const double* RTOL, const int* ICNTRL_U,
                      const double* RCNTRL_U, int* ISTATUS_U, double* RSTATUS_U) {
  int ICNTRL[20] = {0};
  double RCNTRL[20] = {0};
  ICNTRL[1] = 1; ICNTRL[2] = 5; ICNTRL[11] = 1;
  if (ICNTRL_U) for (int i = 0; i < 20; ++i) if (ICNTRL_U[i] >= 0) ICNTRL[i] = ICNTRL_U[i];
  if (RCNTRL_U) for (int i = 0; i < 20; ++i) if (RCNTRL_U[i] >= 0) RCNTRL[i] = RCNTRL_U[i];
  Rosenbrock<Model> r(mdl);
  int IERR = ros_parse(ROS_TLM, Model::N, ICNTRL, RCNTRL, TIN, TOUT, ATOL, RTOL, r.cfg);
  if (IERR == 1) IERR = r.tlm(Y, NTLM, Y_tlm, TIN, TOUT, ATOL, RTOL, ATOL_tlm, RTOL_tlm, ICNTRL[11] != 0);
  if (ISTATUS_U) for (int i = 0; i < 20; ++i) ISTATUS_U[i] = r.ISTATUS[i];
  if (RSTATUS_U) for (int i = 0; i < 20; ++i) RSTATUS_U[i] = r.RSTATUS[i];
  return IERR;
}

}  // namespace oracle
