// ORACLE — TEST INFRASTRUCTURE ONLY (never linked into the product path).
//
// CPU restatement of the model callbacks shipped with the reference examples.
// Every function cites the reference lines it follows.  REAL(4) literals of the
// Fortran sources are spelled F4(x) = (double)(float)x.
// All models are plain structs with per-instance state (re-entrant; the
// reference keeps this state in module globals).
#pragma once
#include <cmath>
#include <cstring>
#include <cstdlib>
#include "gen/cbm4_gen.h"
#include "portable_math.hpp"

namespace oracle {

// 0: deterministic kernels of portable_math.hpp (default), 1: libm cos/pow (see portable_math.hpp)
inline int& libm_mode() { static int m = (getenv("ORACLE_LIBM") && atoi(getenv("ORACLE_LIBM")) != 0) ? 1 : 0; return m; }
static inline double o_cos(double x) { return libm_mode() ? std::cos(x) : oracle_pm::cos_cr(x); }
static inline double o_root(double x, double elo) { return libm_mode() ? std::pow(x, 1.0 / elo) : oracle_pm::root_elo(x, elo); }

#define F4(x) ((double)(x##f))

// Update_SUN: EXAMPLES/cbm4/CBM4_ROS_ADJ/cbm4_Parameters.F90:196-220 (d0 literals there;
// small_strato's copy User_Parameters.F90:87-112 uses exactly representable REAL(4)
// literals 4.5, 19.5, 3600.0, 2.0, 1.0 => same values).
static inline double update_sun(double T) {
  const double PI = 3.14159265358979;
  const double SunRise = 4.5, SunSet = 19.5;
  double Thour = T / 3600.0;
  double Tlocal = Thour - (double)(((int)Thour / 24) * 24);   // INT() truncates; integer division
  if (Tlocal >= SunRise && Tlocal <= SunSet) {
    double Ttmp = (2.0 * Tlocal - SunRise - SunSet) / (SunSet - SunRise);
    if (Ttmp > 0) Ttmp = Ttmp * Ttmp; else Ttmp = -Ttmp * Ttmp;
    return (1.0 + o_cos(PI * Ttmp)) / 2.0;
  }
  return 0.0;
}

// ---------------------------------------------------------------------------
// CBM-IV (EXAMPLES/cbm4/CBM4_ROS_ADJ): N=32, NP=29 constant rate coefficients
struct Cbm4 {
  static constexpr int N = 32, NP = 29;
  double temp = 298.0;                         // cbm4_ros_adj_dr.F90:268
  double fix[1] = {(1.25e+8) * 2.55e10};       // :178,202
  double rconst[81];
  // jcoeff: cbm4_ros_adj_dr.F90:31-32 / :57-58 (1-based reaction numbers)
  static const int* jcoeff() {
    static const int jc[29] = {4, 11, 18, 21, 24, 25, 36, 37, 41, 44, 49, 50, 52, 54, 55,
                               59, 64, 65, 66, 67, 68, 70, 73, 75, 76, 77, 78, 79, 81};
    return jc;
  }
  Cbm4() {
    std::memset(rconst, 0, sizeof rconst);
    // constant rate coefficients as the driver sets them (dr.F90:204-232); never read by
    // FUN/JAC/Hessian (hard-coded there) but kept so RCONST is fully defined
    static const int idx[29] = {4, 11, 18, 21, 24, 25, 36, 37, 41, 44, 49, 50, 52, 54, 55,
                                59, 64, 65, 66, 67, 68, 70, 73, 75, 76, 77, 78, 79, 81};
    static const double val[29] = {9.3e-12, 2.2e-10, 1.3e-21, 4.39999e-40, 6.6e-12, 1e-20, 2.2e-13, 1e-11,
                                   6.3e-16, 2.5e-15, 2e-12, 6.5e-12, 8.1e-13, 1600.0, 1.5e-11, 7.7e-15,
                                   8.1e-12, 4.2, 4.1e-11, 2.2e-11, 1.4e-11, 3e-11, 1.7e-11, 1.8e-11,
                                   9.6e-11, 1.2e-17, 3.2e-13, 8.1e-12, 6.8e-13};
    for (int i = 0; i < 29; ++i) rconst[idx[i] - 1] = val[i];
  }
  void update(double t) {                      // update_sun(t); update_rconst()   (dr.F90:91-92 etc.)
    double sun = update_sun(t);
    cbm4_gen::update_rconst(sun, temp, rconst);
  }
  // fun: cbm4_ros_adj_dr.F90:83-98
  void fun(double t, const double* y, double* p) {
    update(t);
    cbm4_gen::fun(y, fix, rconst, p);
    p[30] = p[30] + 2.0e3;                     // p(31)
    p[25] = p[25] + 2.0e3;                     // p(26)
  }
  // jac: cbm4_ros_adj_dr.F90:69-80 (zeroes, then fills)
  void jac(double t, const double* y, double* fjac) {
    std::memset(fjac, 0, sizeof(double) * N * N);
    update(t);
    cbm4_gen::jac(y, fix, rconst, fjac);
  }
  // hesstr_vec: cbm4_ros_adj_dr.F90:2-17   htv = (f_yy x v)^T u
  void hesstr_vec(double t, const double* y, const double* u, const double* v, double* htv) {
    double hess[cbm4_gen::NHESS];
    update(t);
    cbm4_gen::hessian(y, fix, rconst, hess);
    cbm4_gen::hesstr_vec(hess, u, v, htv);
  }
  // HESS_VEC callback of INTEGRATE_TLM: hv = (f_yy x u) . v  (no cbm4 TLM driver upstream; see gen_cbm4.py)
  void hess_vec(double t, const double* y, const double* u, const double* v, double* hv) {
    double hess[cbm4_gen::NHESS];
    update(t);
    cbm4_gen::hessian(y, fix, rconst, hess);
    cbm4_gen::hess_vec(hess, u, v, hv);
  }
  // dFun_dRcoeff: cbm4_Parameters.F90:1135-1175
  void dfun_drcoeff(const double* v, double* dfdr) {
    double arp[81];
    cbm4_gen::reactant_prod(v, fix, arp);
    const int* jc = jcoeff();
    for (int j = 0; j < NP; ++j) {
      for (int i = 0; i < N; ++i) dfdr[i + N * j] = 0.0;
      double aj = arp[jc[j] - 1];
      for (int k = cbm4_gen::CCOL_STOICM[jc[j] - 1]; k <= cbm4_gen::CCOL_STOICM[jc[j]] - 1; ++k)
        dfdr[(cbm4_gen::IROW_STOICM[k - 1] - 1) + N * j] = cbm4_gen::STOICM[k - 1] * aj;
    }
  }
  // dJac_dRcoeff: cbm4_Parameters.F90:2330-2380
  void djac_drcoeff(const double* v, const double* u, double* djdr) {
    double jvrp[cbm4_gen::NJVRP];
    cbm4_gen::jac_reactant_prod(v, fix, jvrp);
    const int* jc = jcoeff();
    for (int j = 0; j < NP; ++j) {
      for (int i = 0; i < N; ++i) djdr[i + N * j] = 0.0;
      double aj = 0.0;
      for (int k = cbm4_gen::CROW_JVRP[jc[j] - 1]; k <= cbm4_gen::CROW_JVRP[jc[j]] - 1; ++k)
        aj = aj + jvrp[k - 1] * u[cbm4_gen::ICOL_JVRP[k - 1] - 1];
      for (int k = cbm4_gen::CCOL_STOICM[jc[j] - 1]; k <= cbm4_gen::CCOL_STOICM[jc[j]] - 1; ++k)
        djdr[(cbm4_gen::IROW_STOICM[k - 1] - 1) + N * j] = cbm4_gen::STOICM[k - 1] * aj;
    }
  }
  // jacp: cbm4_ros_adj_dr.F90:49-66
  void jacp(double t, const double* y, double* fpjac) {
    std::memset(fpjac, 0, sizeof(double) * N * NP);
    update(t);
    dfun_drcoeff(y, fpjac);
  }
  // hesstr_vec_f_py: cbm4_ros_adj_dr.F90:20-45   htvg = (f_py x k)^T u
  void hesstr_vec_f_py(double t, const double* y, const double* u, const double* k, double* htvg) {
    double djdr[N * NP];
    update(t);
    djac_drcoeff(y, k, djdr);
    for (int i = 0; i < NP; ++i) {
      htvg[i] = 0.0;
      for (int j = 0; j < N; ++j) htvg[i] = htvg[i] + djdr[j + N * i] * u[j];
    }
  }
  // adjinit: cbm4_ros_adj_dr.F90:101-117
  void adjinit(int nadj, double, const double*, double* lambda, double* mu) {
    for (int i = 0; i < N * nadj; ++i) lambda[i] = 0.0;
    for (int k = 0; k < nadj; ++k) lambda[k + N * k] = 1.0;
    if (mu) for (int i = 0; i < NP * nadj; ++i) mu[i] = 0.0;
  }
  // driver initial state: cbm4_ros_adj_dr.F90:233-264 (var(7) = 0.0 REAL(4) zero)
  static void initial_state(double* var) {
    static const double v[32] = {
        3.652686375061191e-02, 3.470772014720235e+11, 2.558736854382598e+03, 3.353179955426548e-23,
        5.287698343141944e-20, 1.307511063007438e+07, 0.0, 1.985180650951722e+01,
        3.839565210778183e+08, 2.675425503948970e+08, 1.593461800730453e-24, 1.191924698309334e+12,
        4.894101892069533e-05, 5.462490511678749e-21, 1.943997577140665e-23, 2.329863421033919e+12,
        2.106517721289787e-30, 5.518957651363696e+08, 2.633932009302457e-21, 8.118206770655152e+03,
        3.000632019118241e+10, 0.0, 0.0, 2.687919546821573e+02,
        2.061185883675469e+12, 1.561906017211229e+10, 3.671987743784878e+07, 9.888485531199334e+08,
        1.248793414327174e+04, 3.357869003391008e+07, 2.727638922137502e+09, 6.534644475169904e+00};
    std::memcpy(var, v, sizeof v);
  }
};

// ---------------------------------------------------------------------------
// small_strato (EXAMPLES/small_strato/small_ros{,_adj}/User_Parameters.F90): N=5, NFIX=2
struct SmallStrato {
  static constexpr int N = 5, NP = 0;
  double fix[2] = {F4(8.120E+16) * F4(1.000000e+00), F4(1.697E+16) * F4(1.000000e+00)};  // dr.F90:55-56
  double rconst[10];
  SmallStrato() { update(0.0); }
  // Update_RCONST: User_Parameters.F90:114-127 (all REAL(4) literals)
  void update(double t) {
    double SUN = update_sun(t);
    rconst[0] = ((F4(2.643E-10)) * SUN * SUN * SUN);
    rconst[1] = ((F4(8.018E-17)));
    rconst[2] = ((F4(6.120E-04)) * SUN);
    rconst[3] = ((F4(1.576E-15)));
    rconst[4] = ((F4(1.070E-03)) * SUN * SUN);
    rconst[5] = ((F4(7.110E-11)));
    rconst[6] = ((F4(1.200E-10)));
    rconst[7] = ((F4(6.062E-15)));
    rconst[8] = ((F4(1.069E-11)));
    rconst[9] = ((F4(1.289E-02)) * SUN);
  }
  // FUN_U: User_Parameters.F90:10-37
  void fun(double t, const double* V, double* FCT) {
    double A[10];
    const double* RCONST = rconst; const double* FIX = fix;
    update(t);
    A[0] = RCONST[0] * FIX[1];
    A[1] = RCONST[1] * V[1] * FIX[1];
    A[2] = RCONST[2] * V[2];
    A[3] = RCONST[3] * V[1] * V[2];
    A[4] = RCONST[4] * V[2];
    A[5] = RCONST[5] * V[0] * FIX[0];
    A[6] = RCONST[6] * V[0] * V[2];
    A[7] = RCONST[7] * V[2] * V[3];
    A[8] = RCONST[8] * V[1] * V[4];
    A[9] = RCONST[9] * V[4];
    FCT[0] = A[4] - A[5] - A[6];
    FCT[1] = 2 * A[0] - A[1] + A[2] - A[3] + A[5] - A[8] + A[9];
    FCT[2] = A[1] - A[2] - A[3] - A[4] - A[6] - A[7];
    FCT[3] = -A[7] + A[8] + A[9];
    FCT[4] = A[7] - A[8] - A[9];
  }
  // JAC_U: User_Parameters.F90:39-85 (zeroes JF itself, :50)
  void jac(double t, const double* V, double* JF) {
    double B[17];
    const double* RCONST = rconst; const double* FIX = fix;
    update(t);
    for (int i = 0; i < 25; ++i) JF[i] = 0.0;
#define JFX(i, j) JF[(i - 1) + 5 * (j - 1)]
    B[2] = RCONST[1] * FIX[1];
    B[4] = RCONST[2];
    B[5] = RCONST[3] * V[2];
    B[6] = RCONST[3] * V[1];
    B[7] = RCONST[4];
    B[8] = RCONST[5] * FIX[0];
    B[10] = RCONST[6] * V[2];
    B[11] = RCONST[6] * V[0];
    B[12] = RCONST[7] * V[3];
    B[13] = RCONST[7] * V[2];
    B[14] = RCONST[8] * V[4];
    B[15] = RCONST[8] * V[1];
    B[16] = RCONST[9];
    JFX(1, 1) = -B[8] - B[10];
    JFX(1, 3) = B[7] - B[11];
    JFX(2, 1) = B[8];
    JFX(2, 2) = -B[2] - B[5] - B[14];
    JFX(2, 3) = B[4] - B[6];
    JFX(2, 5) = -B[15] + B[16];
    JFX(3, 1) = -B[10];
    JFX(3, 2) = B[2] - B[5];
    JFX(3, 3) = -B[4] - B[6] - B[7] - B[11] - B[12];
    JFX(3, 4) = -B[13];
    JFX(3, 5) = 0;
    JFX(4, 2) = B[14];
    JFX(4, 3) = -B[12];
    JFX(4, 4) = -B[13];
    JFX(4, 5) = B[15] + B[16];
    JFX(5, 2) = -B[14];
    JFX(5, 3) = B[12];
    JFX(5, 4) = B[13];
    JFX(5, 5) = -B[15] - B[16];
#undef JFX
  }
  // HessTR_Vec_U: small_ros_adj/User_Parameters.F90:136-190 (uses RCONST as last updated;
  // entries 4,7,8,9 are time independent)
  void hesstr_vec(double, const double*, const double* U1, const double* U2, double* HTU) {
    double D2A[5], HESS[11];
    D2A[1] = rconst[3]; D2A[2] = rconst[6]; D2A[3] = rconst[7]; D2A[4] = rconst[8];
    HESS[1] = -D2A[2]; HESS[2] = -D2A[1]; HESS[3] = -D2A[4]; HESS[4] = -D2A[2]; HESS[5] = -D2A[1];
    HESS[6] = -D2A[3]; HESS[7] = D2A[4]; HESS[8] = -D2A[3]; HESS[9] = -D2A[4]; HESS[10] = D2A[3];
#define u1(i) U1[i - 1]
#define u2(i) U2[i - 1]
    HTU[0] = HESS[1] * (u1(1) * u2(3)) + HESS[4] * (u1(3) * u2(3));
    HTU[1] = HESS[2] * (u1(2) * u2(3)) + HESS[3] * (u1(2) * u2(5)) + HESS[5] * (u1(3) * u2(3))
           + HESS[7] * (u1(4) * u2(5)) + HESS[9] * (u1(5) * u2(5));
    HTU[2] = HESS[1] * (u1(1) * u2(1)) + HESS[2] * (u1(2) * u2(2)) + HESS[4] * (u1(3) * u2(1))
           + HESS[5] * (u1(3) * u2(2)) + HESS[6] * (u1(3) * u2(4))
           + HESS[8] * (u1(4) * u2(4)) + HESS[10] * (u1(5) * u2(4));
    HTU[3] = HESS[6] * (u1(3) * u2(3)) + HESS[8] * (u1(4) * u2(3)) + HESS[10] * (u1(5) * u2(3));
    HTU[4] = HESS[3] * (u1(2) * u2(2)) + HESS[7] * (u1(4) * u2(2)) + HESS[9] * (u1(5) * u2(2));
#undef u1
#undef u2
  }
  void jacp(double, const double*, double*) {}
  void hesstr_vec_f_py(double, const double*, const double*, const double*, double*) {}
  // adjinit: small_ros_adj/User_Parameters.F90:192-204
  void adjinit(int nadj, double, const double*, double* lambda, double*) {
    for (int i = 0; i < N * nadj; ++i) lambda[i] = 0.0;
    for (int k = 0; k < nadj; ++k) lambda[k + N * k] = 1.0;
  }
  // driver initial state: small_ros/small_strato_dr.F90:50-54 (REAL(4) literals * CFACTOR)
  static void initial_state(double* var) {
    const double cf = F4(1.000000e+00);
    var[0] = F4(9.906E+01) * cf; var[1] = F4(6.624E+08) * cf; var[2] = F4(5.326E+11) * cf;
    var[3] = F4(8.725E+08) * cf; var[4] = F4(2.240E+08) * cf;
  }
};

// ---------------------------------------------------------------------------
// Roberts (EXAMPLES/roberts/ROBERTS_ROS_ADJ/roberts_ros_adj_dr.F90): N=3, NP=3
struct Roberts {
  static constexpr int N = 3, NP = 3;
  const double c[3] = {F4(0.04e0), F4(1.0e4), F4(3.0e7)};        // :66-68, :82-84, :102
  // fun :97-110
  void fun(double, const double* y, double* p) {
    p[0] = -c[0] * y[0] + c[1] * y[1] * y[2];
    p[2] = c[2] * y[1] * y[1];
    p[1] = -p[0] - p[2];
  }
  // jac :75-94 — does NOT write fjac(3,1), fjac(3,3); caller's buffer is
  // zero-initialised once at LSS_Init time in this oracle
  void jac(double, const double* y, double* fjac) {
#define FJ(i, j) fjac[(i - 1) + 3 * (j - 1)]
    FJ(1, 1) = -c[0];
    FJ(1, 2) = c[1] * y[2];
    FJ(1, 3) = c[1] * y[1];
    FJ(2, 1) = c[0];
    FJ(2, 2) = -c[1] * y[2] - 2 * c[2] * y[1];
    FJ(2, 3) = -c[1] * y[1];
    FJ(3, 2) = 2 * c[2] * y[1];
#undef FJ
  }
  // hesstr_vec :59-72
  void hesstr_vec(double, const double*, const double* u, const double* k, double* tmp) {
    tmp[0] = 0;
    tmp[1] = 2 * c[2] * k[1] * (u[2] - u[1]) + c[1] * k[2] * (u[0] - u[1]);
    tmp[2] = c[1] * k[1] * (u[0] - u[1]);
  }
  // JACP :18-29 (writes 6 of 9 entries; the rest stay as initialised = 0)
  void jacp(double, const double* y, double* fpjac) {
#define FP(i, j) fpjac[(i - 1) + 3 * (j - 1)]
    FP(1, 1) = -y[0];
    FP(2, 1) = y[0];
    FP(1, 2) = y[1] * y[2];
    FP(2, 2) = -y[1] * y[2];
    FP(2, 3) = -y[1] * y[1];
    FP(3, 3) = y[1] * y[1];
#undef FP
  }
  // HESSTR_VEC_F_PY :31-40
  void hesstr_vec_f_py(double, const double* y, const double* u, const double* k, double* tmp) {
    tmp[0] = (u[1] - u[0]) * k[0];
    tmp[1] = (u[0] - u[1]) * (y[2] * k[1] + y[1] * k[2]);
    tmp[2] = 2 * y[1] * k[1] * (u[2] - u[1]);
  }
  // adjinit :124-146
  void adjinit(int nadj, double, const double* y, double* lambda, double* mu) {
    for (int i = 0; i < N * nadj; ++i) lambda[i] = 0.0;
    for (int k = 0; k < nadj; ++k) lambda[k + N * k] = y[k];
    if (mu) {
      for (int i = 0; i < NP * nadj; ++i) mu[i] = 0.0;
#define MU(i, j) mu[(i - 1) + 3 * (j - 1)]
      MU(1, 1) = -y[0] * y[0];
      MU(1, 2) = y[0] * y[1] * y[2];
      MU(2, 1) = y[0] * y[1];
      MU(2, 2) = -y[1] * y[1] * y[2];
      MU(2, 3) = -y[1] * y[1] * y[1];
      MU(3, 3) = y[1] * y[1] * y[2];
#undef MU
    }
  }
  static void initial_state(double* var) { var[0] = 1; var[1] = 0; var[2] = 0; }   // :194-196
};

}  // namespace oracle
