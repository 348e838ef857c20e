// ORACLE — TEST INFRASTRUCTURE ONLY (never linked into the product path).
//
// CPU restatement of the reference's explicit Runge-Kutta family and of the shallow-water example it is shipped with:
//   FWD/ERK/ERK_f90_Integrator.F90            INTEGRATE :28-90 (the `> 0` override rule), ERK (option parsing) :92-362,
//                                             ERK_Integrator :364-503, ERK_ErrorScale :507-522, ERK_ErrorNorm :526-537,
//                                             coefficient sets :580-923 (gen/erk_tableau_gen.h)
//   TLM/ERK_TLM/ERK_TLM_f90_Integrator.F90    INTEGRATE_TLM :28-93, ERKTLM options :95-389 (ICNTRL(5) = 0: finite-difference
//                                             Jacobian-vector products; ICNTRL(12): TLM truncation error; RCNTRL(12):
//                                             FDincrement), ERK_IntegratorTLM :391-576, ERK_ErrorNorm_TLM :613-626,
//                                             FDJACV :1071-1097 (forward difference; it perturbs the stage vector in place,
//                                             so with NTLM > 1 the perturbations of earlier columns stay in it)
//   EXAMPLES/swe/SWE_ERK/swe2D_upwind.F90     module constants :11-32, initializeGaussianGrid :86-126, grid2vec :267-289,
//                                             vec2grid :292-315, compute_F :318-567 (third-order upwind flux splitting,
//                                             periodic halo); FUN = vec2grid -> compute_F -> grid2vec
//                                             (EXAMPLES/swe/SWE_ERK/swe2D_erk_dr.F90:2-14)
// Pinned: FWD/ERK + the SWE model against the reference's golden file EXAMPLES/swe/SWE_ERK/swe_lsode_sol.txt (tests/golden/
// swe_lsode_sol.txt; relative L2 error 7e-3 / 2e-4 / 1e-6 at tolerances 1e-2 / 1e-4 / 1e-6, tests/test_oracle_erk.py).
// PARITY UNPINNED by reference artefacts for the ERK_TLM part (the TLM driver only prints): checked against finite
// differences of the pinned state sweep.
// The dense-Jacobian TLM mode (ICNTRL(5) /= 0: JAC + MATMUL(FJAC, S_TLM) per stage, :460-474) is restated for the shallow-water
// model with its JAC callback kept as sparse rows (SweJac below; MATMUL's column sweep = per row the columns ascending).
// Statement order and BLAS call order follow the Fortran (DAXPY with a zero factor returns at once, as netlib's does);
// compile with -ffp-contract=off.  Err**(-1/rkELO) goes through o_pow_neg_inv (portable kernel by default, libm pow in
// ORACLE_LIBM mode).
#pragma once
#include <cmath>
#include <cstring>
#include <memory>
#include <type_traits>
#include <vector>
#include <algorithm>
#include <utility>
#include "sdirk.hpp"
#include "gen/erk_tableau_gen.h"
#include "gen/swe_jac_gen.h"

namespace oracle {

// ---- shallow-water model (swe2Dxy_Parameters + compute_F)
struct Swe {
  static constexpr int Mx = 40, My = 40, CELLS = Mx * My, N = 3 * CELLS;
  static constexpr int GX = Mx + 4, GY = My + 4;
  static constexpr double gAcc = 9.80616;
  static double dx() { return (3.0 - (-3.0)) / (Mx - 1.0); }
  static double dy() { return (3.0 - (-3.0)) / (My - 1.0); }
  std::vector<double> U, F;   // (3, Mx+4, My+4) column-major: U(k,i,j) at k + 3*(i + GX*j), 0-based
  std::vector<double> feig, geig;   // feigvalues / geigvalues of compute_F, kept when the JAC callback asks (keep_eig)
  bool keep_eig = false;
  Swe() : U((size_t)3 * GX * GY, 0.0), F((size_t)3 * GX * GY, 0.0) {}
  static inline int at(int k, int i, int j) { return k + 3 * (i + GX * j); }   // 0-based k, i, j

  // initializeGaussianGrid :86-126 followed by the drivers' u = v = const (swe2D_erk_dr.F90:100-103) and grid2vec
  static void initial_state(double A, double u0, double v0, double* vec) {
    std::vector<double> g((size_t)3 * GX * GY, 0.0), x(GX, 0.0), y(GY, 0.0);
    for (int i = 3; i <= Mx + 2; ++i) x[i - 1] = -3.0 + (i - 2) * dx();
    for (int j = 3; j <= My + 2; ++j) y[j - 1] = -3.0 + (j - 2) * dy();
    for (int j = 3; j <= My + 2; ++j)
      for (int i = 3; i <= Mx + 2; ++i) g[at(0, i - 1, j - 1)] = A * std::exp(-x[j - 1] * x[j - 1] - y[i - 1] * y[i - 1]);
    for (int j = 0; j < GY; ++j)
      for (int i = 0; i < GX; ++i) {
        g[at(0, i, j)] = g[at(0, i, j)] + 100.0;
        g[at(1, i, j)] = u0;
        g[at(2, i, j)] = v0;
      }
    grid2vec(g.data(), vec);
  }
  static void grid2vec(const double* g, double* vec) {   // :267-289
    int indx = 0;
    for (int k = 0; k < 3; ++k)
      for (int i = 3; i <= Mx + 2; ++i)
        for (int j = 3; j <= My + 2; ++j) vec[indx++] = g[at(k, i - 1, j - 1)];
  }
  static void vec2grid(const double* vec, double* g) {   // :292-315
    for (size_t q = 0; q < (size_t)3 * GX * GY; ++q) g[q] = 0.0;
    int indx = 0;
    for (int k = 0; k < 3; ++k)
      for (int i = 3; i <= Mx + 2; ++i)
        for (int j = 3; j <= My + 2; ++j) g[at(k, i - 1, j - 1)] = vec[indx++];
  }
  static inline void mv3(const double (&M)[3][3], const double* d, double* out) {   // tmpMV loops
    for (int i1 = 0; i1 < 3; ++i1) {
      out[i1] = 0.0;
      for (int j1 = 0; j1 < 3; ++j1) out[i1] = out[i1] + M[i1][j1] * d[j1];
    }
  }
  void compute_F() {   // :318-567 (1-based i, j below; u(k,i,j) is 1-based in all three indices)
    double* Ug = U.data();
    auto u = [&](int k, int i, int j) -> double& { return Ug[at(k - 1, i - 1, j - 1)]; };
    for (int k = 1; k <= 3; ++k) {
      for (int j = 3; j <= My + 2; ++j) {
        u(k, 2, j) = u(k, Mx + 2, j);
        u(k, 1, j) = u(k, Mx + 1, j);
        u(k, Mx + 3, j) = u(k, 3, j);
        u(k, Mx + 4, j) = u(k, 4, j);
      }
      for (int i = 3; i <= Mx + 2; ++i) {
        u(k, i, 2) = u(k, i, My + 2);
        u(k, i, 1) = u(k, i, My + 1);
        u(k, i, My + 3) = u(k, i, 3);
        u(k, i, My + 4) = u(k, i, 4);
      }
    }
    const double g = gAcc, ddx = dx(), ddy = dy();
    if (keep_eig && feig.empty()) { feig.assign((size_t)3 * GX * GY, 0.0); geig.assign((size_t)3 * GX * GY, 0.0); }
    for (int j = 3; j <= My + 2; ++j)
      for (int i = 3; i <= Mx + 2; ++i) {
        double eigf[3], eigg[3], Fu1[3][3] = {{0}}, Fu2[3][3] = {{0}}, Fu3[3][3] = {{0}}, Gu1[3][3] = {{0}}, Gu2[3][3] = {{0}},
                                 Gu3[3][3] = {{0}};
        double dUdxMinus[3], dUdxPlus[3], dUdyMinus[3], dUdyPlus[3], tmpF[3], tmpG[3], tmpMV[3];
        eigf[0] = u(2, i, j) / u(1, i, j);
        eigf[1] = u(2, i, j) / u(1, i, j) - std::sqrt(gAcc * u(1, i, j));
        eigf[2] = u(2, i, j) / u(1, i, j) + std::sqrt(gAcc * u(1, i, j));
        const double hc = u(1, i, j), uc = u(2, i, j) / u(1, i, j), vc = u(3, i, j) / u(1, i, j);
        Fu1[2][0] = -vc;
        Fu1[2][2] = 1.0;
        double isqrt = 1.0 / std::sqrt(g * hc);
        Fu2[0][0] = 0.5 * (1.0 + uc * isqrt);
        Fu2[0][1] = -0.5 * isqrt;
        Fu2[1][0] = 0.5 * (uc * uc - g * hc) * isqrt;
        Fu2[1][1] = 0.5 - 0.5 * uc * isqrt;
        Fu2[2][0] = 0.5 * (1.0 + uc * isqrt) * vc;
        Fu2[2][1] = -0.5 * vc * isqrt;
        Fu3[0][0] = 0.5 - 0.5 * uc * isqrt;
        Fu3[0][1] = 0.5 * isqrt;
        Fu3[1][0] = 0.5 * (g * hc - uc * uc) * isqrt;
        Fu3[1][1] = 0.5 * (1.0 + uc * isqrt);
        Fu3[2][0] = 0.5 * (1.0 - uc * isqrt) * vc;
        Fu3[2][1] = 0.5 * vc * isqrt;
        for (int k = 1; k <= 3; ++k) {
          dUdxMinus[k - 1] = (-u(k, i + 2, j) + 6.0 * u(k, i + 1, j) - 3.0 * u(k, i, j) - 2.0 * u(k, i - 1, j)) / (6.0 * ddx);
          dUdxPlus[k - 1] = (2.0 * u(k, i + 1, j) + 3.0 * u(k, i, j) - 6.0 * u(k, i - 1, j) + u(k, i - 2, j)) / (6.0 * ddx);
        }
        mv3(Fu1, eigf[0] >= 0.0 ? dUdxPlus : dUdxMinus, tmpMV);
        for (int k = 0; k < 3; ++k) tmpF[k] = eigf[0] * tmpMV[k];
        mv3(Fu2, eigf[1] >= 0.0 ? dUdxPlus : dUdxMinus, tmpMV);
        for (int k = 0; k < 3; ++k) tmpF[k] = tmpF[k] + eigf[1] * tmpMV[k];
        mv3(Fu3, eigf[2] >= 0.0 ? dUdxPlus : dUdxMinus, tmpMV);
        for (int k = 0; k < 3; ++k) tmpF[k] = tmpF[k] + eigf[2] * tmpMV[k];
        eigg[0] = u(3, i, j) / u(1, i, j);
        eigg[1] = u(3, i, j) / u(1, i, j) - std::sqrt(gAcc * u(1, i, j));
        eigg[2] = u(3, i, j) / u(1, i, j) + std::sqrt(gAcc * u(1, i, j));
        Gu1[1][0] = -uc;
        Gu1[1][1] = 1.0;
        isqrt = 1.0 / std::sqrt(g * hc);
        Gu2[0][0] = 0.5 * (1.0 + vc * isqrt);
        Gu2[0][2] = -0.5 * isqrt;
        Gu2[1][0] = 0.5 * uc * (1.0 + vc * isqrt);
        Gu2[1][2] = -0.5 * uc * isqrt;
        Gu2[2][0] = 0.5 * (vc * vc - g * hc) * isqrt;
        Gu2[2][2] = 0.5 - 0.5 * vc * isqrt;
        Gu3[0][0] = 0.5 * (1.0 - vc * isqrt);
        Gu3[0][2] = 0.5 * isqrt;
        Gu3[1][0] = 0.5 * uc * (1.0 - vc * isqrt);
        Gu3[1][2] = 0.5 * uc * isqrt;
        Gu3[2][0] = 0.5 * (-vc * vc + g * hc) * isqrt;
        Gu3[2][2] = 0.5 + 0.5 * vc * isqrt;
        for (int k = 1; k <= 3; ++k) {
          dUdyMinus[k - 1] = (-u(k, i, j + 2) + 6.0 * u(k, i, j + 1) - 3.0 * u(k, i, j) - 2.0 * u(k, i, j - 1)) / (6.0 * ddy);
          dUdyPlus[k - 1] = (2.0 * u(k, i, j + 1) + 3.0 * u(k, i, j) - 6.0 * u(k, i, j - 1) + u(k, i, j - 2)) / (6.0 * ddy);
        }
        mv3(Gu1, eigg[0] >= 0.0 ? dUdyPlus : dUdyMinus, tmpMV);
        for (int k = 0; k < 3; ++k) tmpG[k] = eigg[0] * tmpMV[k];
        mv3(Gu2, eigg[1] >= 0.0 ? dUdyPlus : dUdyMinus, tmpMV);
        for (int k = 0; k < 3; ++k) tmpG[k] = tmpG[k] + eigg[1] * tmpMV[k];
        mv3(Gu3, eigg[2] >= 0.0 ? dUdyPlus : dUdyMinus, tmpMV);
        for (int k = 0; k < 3; ++k) tmpG[k] = tmpG[k] + eigg[2] * tmpMV[k];
        for (int k = 0; k < 3; ++k) F[at(k, i - 1, j - 1)] = -tmpF[k] - tmpG[k];
        if (keep_eig) for (int k = 0; k < 3; ++k) { feig[at(k, i - 1, j - 1)] = eigf[k]; geig[at(k, i - 1, j - 1)] = eigg[k]; }
      }
  }
  void fun(double /*t*/, const double* y, double* p) {   // swe2D_erk_dr.F90:2-14
    vec2grid(y, U.data());
    compute_F();
    grid2vec(F.data(), p);
  }
};

// ---- the JAC callback of the shallow-water drivers (EXAMPLES/swe/SWE_ERK_TLM/swe2D_erk_tlm_dr.F90:2-14, SWE_ERK_ADJ/
// swe2D_erk_adj_dr.F90:2-13: vec2grid, compute_F for the eigenvalue signs, compute_JacF, fjac = JF; the two examples carry the same
// compute_JacF) kept as the 27 structural entries of each row instead of the dense 4800 x 4800 matrix; gen/swe_jac_gen.h
// (tools/gen_swe_jac.py) holds the formula export statement by statement
struct SweJac {
  static constexpr int N = Swe::N, NNZR = 27, Mx = Swe::Mx, My = Swe::My;
  std::vector<int> col;       // [N][27]
  std::vector<double> val;    // [N][27]  JAC(p, col) = -JF - JG
  SweJac() : col((size_t)N * NNZR), val((size_t)N * NNZR) {}

  static void mapgrinv(int p, int& k, int& i, int& j) {   // 1-based p -> 1-based field, grid indices with the halo of two
    k = (p - 1) / (Mx * My) + 1;
    const int aux = (p - 1) % (Mx * My);
    i = aux / Mx + 1 + 2;
    j = aux % Mx + 1 + 2;
  }
  static int mapgr(int p, int dk, int di, int dj) {       // 1-based index of the periodic neighbour
    int z, x, y;
    mapgrinv(p, z, x, y);
    z += dk;
    x += di;
    if (x == 2) x = Mx + 2;
    if (x == 1) x = Mx + 1;
    if (x == Mx + 3) x = 3;
    if (x == Mx + 4) x = 4;
    y += dj;
    if (y == 2) y = My + 2;
    if (y == 1) y = My + 1;
    if (y == My + 3) y = 3;
    if (y == My + 4) y = 4;
    return (z - 1) * Mx * My + (x - 3) * Mx + y - 2;
  }
  // jac(n, t, y, fjac) of the driver
  void build(Swe& m, const double* y) {
    m.vec2grid(y, m.U.data());
    m.keep_eig = true;
    m.compute_F();
    const double* Ug = m.U.data();
    auto u = [&](int k, int i, int j) { return Ug[Swe::at(k - 1, i - 1, j - 1)]; };
    const double dx = Swe::dx(), dy = Swe::dy(), g = Swe::gAcc;
    for (int p = 1; p <= N; ++p) {
      int kf, i, j;
      mapgrinv(p, kf, i, j);
      double V[3][9], JF[3][5], JG[3][5];
      bool fe[3], ge[3];
      for (int f = 1; f <= 3; ++f) {
        for (int d = -2; d <= 2; ++d) V[f - 1][d + 2] = u(f, i + d, j);
        V[f - 1][5] = u(f, i, j - 2);
        V[f - 1][6] = u(f, i, j - 1);
        V[f - 1][7] = u(f, i, j + 1);
        V[f - 1][8] = u(f, i, j + 2);
      }
      for (int q = 0; q < 3; ++q) {
        fe[q] = m.feig[Swe::at(q, i - 1, j - 1)] >= 0.0;
        ge[q] = m.geig[Swe::at(q, i - 1, j - 1)] >= 0.0;
      }
      if (kf == 1) swe_jac::row0(V, fe, ge, dx, dy, g, JF, JG);
      else if (kf == 2) swe_jac::row1(V, fe, ge, dx, dy, g, JF, JG);
      else swe_jac::row2(V, fe, ge, dx, dy, g, JF, JG);
      int* c = col.data() + (size_t)(p - 1) * NNZR;
      double* v = val.data() + (size_t)(p - 1) * NNZR;
      int e = 0;
      for (int cf = 1; cf <= 3; ++cf) {
        for (int d = -2; d <= 2; ++d) {          // JF(p, ind) [+ JG at the centre]
          c[e] = mapgr(p, cf - kf, d, 0) - 1;
          v[e] = -JF[cf - 1][d + 2] - (d == 0 ? JG[cf - 1][2] : 0.0);
          ++e;
        }
        for (int d = -2; d <= 2; ++d) {
          if (d == 0) continue;
          c[e] = mapgr(p, cf - kf, 0, d) - 1;
          v[e] = -0.0 - JG[cf - 1][d + 2];
          ++e;
        }
      }
    }
  }
  // DGEMV('T', N, N, 1, FJAC, N, x, 1, 0, out, 1)
  void mul_t(const double* x, double* out) const {
    for (int i = 0; i < N; ++i) out[i] = 0.0;
    for (int r = 0; r < N; ++r) {
      const int* c = col.data() + (size_t)r * NNZR;
      const double* v = val.data() + (size_t)r * NNZR;
      for (int e = 0; e < NNZR; ++e) out[c[e]] = out[c[e]] + v[e] * x[r];
    }
  }
  // y = FJAC x (MATMUL / DGEMV('N') of the dense-Jacobian tangent linear mode): column sweep, i.e. per row ascending columns
  void mul_n(const double* x, double* out) const {
    std::vector<std::pair<int, double>> t(NNZR);
    for (int r = 0; r < N; ++r) {
      const int* c = col.data() + (size_t)r * NNZR;
      const double* v = val.data() + (size_t)r * NNZR;
      for (int e = 0; e < NNZR; ++e) t[e] = {c[e], v[e]};
      std::sort(t.begin(), t.end());
      double s = 0.0;
      for (int e = 0; e < NNZR; ++e) s = s + t[e].second * x[t[e].first];
      out[r] = s;
    }
  }
};

struct ErkConfig {
  ErkTableauData tab;
  int ITOL, Max_no_steps;
  bool FDAprox = true, TLMtruncErr = false;
  double Roundoff, Hmin, Hmax, Hstart, FacMin, FacMax, FacRej, FacSafe, Qmin, Qmax, FDincrement;
};

// ERK :229-358 / ERKTLM :238-389.  Returns Ierr (0 = fine; the reference keeps parsing after an option error).
static inline int erk_parse(int N, const int* ICNTRL, const double* RCNTRL, double Tinitial, double Tfinal, const double* AbsTol,
                            const double* RelTol, bool tlm, ErkConfig& c) {
  int Ierr = 0;
  c.ITOL = (ICNTRL[1] == 0) ? 1 : 0;
  int method = 4;
  switch (ICNTRL[2]) {
    case 1: method = 1; break;
    case 2: method = 2; break;
    case 3: method = 3; break;
    case 5: method = 5; break;
    case 6: method = 6; break;
    default: method = 4;   // 0, 4 and CASE DEFAULT: Dopri5
  }
  c.tab = ERK_TABLEAU[method - 1];
  if (ICNTRL[3] == 0) c.Max_no_steps = 10000;
  else if (ICNTRL[3] > 0) c.Max_no_steps = ICNTRL[3];
  else { c.Max_no_steps = 0; Ierr = -1; }
  if (tlm) {
    c.FDAprox = (ICNTRL[4] == 0);
    c.TLMtruncErr = (ICNTRL[11] != 0);
  }
  c.Roundoff = 1.1102230246251565e-16;   // DLAMCH('E')
  if (RCNTRL[0] == 0.0) c.Hmin = 0.0; else if (RCNTRL[0] > 0.0) c.Hmin = RCNTRL[0]; else { c.Hmin = 0.0; Ierr = -3; }
  const double span = std::fabs(Tfinal - Tinitial);
  if (RCNTRL[1] == 0.0) c.Hmax = span; else if (RCNTRL[1] > 0.0) c.Hmax = std::fmin(std::fabs(RCNTRL[1]), span); else { c.Hmax = span; Ierr = -3; }
  if (RCNTRL[2] == 0.0) c.Hstart = std::fmax(c.Hmin, c.Roundoff);
  else if (RCNTRL[2] > 0.0) c.Hstart = std::fmin(std::fabs(RCNTRL[2]), span);
  else { c.Hstart = 0.0; Ierr = -3; }
  if (RCNTRL[3] == 0.0) c.FacMin = (double)0.2f; else if (RCNTRL[3] > 0.0) c.FacMin = RCNTRL[3]; else { c.FacMin = 0.0; Ierr = -4; }
  if (RCNTRL[4] == 0.0) c.FacMax = (double)10.0f; else if (RCNTRL[4] > 0.0) c.FacMax = RCNTRL[4]; else { c.FacMax = 0.0; Ierr = -4; }
  if (RCNTRL[5] == 0.0) c.FacRej = (double)0.1f; else if (RCNTRL[5] > 0.0) c.FacRej = RCNTRL[5]; else { c.FacRej = 0.0; Ierr = -4; }
  if (RCNTRL[6] == 0.0) c.FacSafe = (double)0.9f; else if (RCNTRL[6] > 0.0) c.FacSafe = RCNTRL[6]; else { c.FacSafe = 0.0; Ierr = -4; }
  c.Qmin = (RCNTRL[9] == 0.0) ? 1.0 : RCNTRL[9];
  c.Qmax = (RCNTRL[10] == 0.0) ? 1.2 : RCNTRL[10];
  c.FDincrement = (RCNTRL[11] == 0.0) ? std::sqrt(std::fmax(RelTol[0], c.Roundoff)) : RCNTRL[11];
  if (c.ITOL == 0) {
    if (AbsTol[0] <= 0.0 || RelTol[0] <= 10.0 * c.Roundoff) Ierr = -5;
  } else {
    for (int i = 0; i < N; ++i)
      if (AbsTol[i] <= 0.0 || RelTol[i] <= 10.0 * c.Roundoff) Ierr = -5;
  }
  return Ierr;
}

template <class Model>
struct Erk {
  static constexpr int N = Model::N;
  Model& mdl;
  ErkConfig cfg;
  int ISTATUS[20];
  double RSTATUS[20];
  std::unique_ptr<SweJac> jac;   // dense-Jacobian tangent linear mode (shallow-water model only)
  explicit Erk(Model& m) : mdl(m) { std::memset(ISTATUS, 0, sizeof ISTATUS); std::memset(RSTATUS, 0, sizeof RSTATUS); }

  static void daxpy(int n, double a, const double* x, double* y) {
    if (a == 0.0) return;
    for (int i = 0; i < n; ++i) y[i] = y[i] + a * x[i];
  }
  void error_scale(const double* AbsTol, const double* RelTol, const double* Y, double* SCAL) const {
    for (int i = 0; i < N; ++i)
      SCAL[i] = (cfg.ITOL == 0) ? 1.0 / (AbsTol[0] + RelTol[0] * std::fabs(Y[i])) : 1.0 / (AbsTol[i] + RelTol[i] * std::fabs(Y[i]));
  }
  static double error_norm(const double* Y, const double* SCAL) {
    double Err = 0.0;
    for (int i = 0; i < N; ++i) { const double q = Y[i] * SCAL[i]; Err = Err + q * q; }
    return std::fmax(std::sqrt(Err / (double)N), 1.0e-10);
  }

  // ERK_Integrator (NTLM = 0) / ERK_IntegratorTLM (FD mode).  Y_TLM (N, NTLM) column-major.
  int integrate(int NTLM, double* Y, double* Y_TLM, double Tinitial, double Tfinal, const double* AbsTol, const double* RelTol,
                const double* AbsTol_TLM, const double* RelTol_TLM) {
    const ErkTableauData& tb = cfg.tab;
    const int S = tb.S;
    std::vector<double> K((size_t)N * S), TMP(N), SCAL(N), S_TLM(N), SCAL_TLM(N), K_TLM((size_t)N * S * std::max(NTLM, 1)),
        Yerr_TLM((size_t)N * std::max(NTLM, 1));
    auto Kt = [&](int istage, int itlm) { return K_TLM.data() + ((size_t)itlm * S + istage) * N; };
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
        for (int i = 0; i < N; ++i) Ki[i] = 0.0;
        for (int i = 0; i < N; ++i) TMP[i] = 0.0;
        for (int j = 1; j <= istage - 1; ++j) daxpy(N, H * tb.A[istage - 1][j - 1], K.data() + (size_t)(j - 1) * N, TMP.data());
        daxpy(N, 1.0, Y, TMP.data());
        mdl.fun(T + tb.C[istage - 1] * H, TMP.data(), Ki);
        ISTATUS[Nfun]++;
        if (NTLM > 0 && !cfg.FDAprox) {   // FJAC = 0; CALL JAC(NVAR, T + c H, TMP, FJAC)  (:460-464)
          if constexpr (std::is_same<Model, Swe>::value) { if (!jac) jac.reset(new SweJac); jac->build(mdl, TMP.data()); }
          ISTATUS[Njac]++;
        }
        for (int itlm = 0; itlm < NTLM; ++itlm) {
          for (int i = 0; i < N; ++i) S_TLM[i] = Y_TLM[(size_t)itlm * N + i];
          for (int j = 1; j <= istage - 1; ++j) daxpy(N, H * tb.A[istage - 1][j - 1], Kt(j - 1, itlm), S_TLM.data());
          if (!cfg.FDAprox) {             // K_TLM(:, istage, itlm) = MATMUL(FJAC, S_TLM)  (:473-474)
            if constexpr (std::is_same<Model, Swe>::value) jac->mul_n(S_TLM.data(), Kt(istage - 1, itlm));
            continue;
          }
          // FDJACV(N, T + c H, TMP, S_TLM, K(:,istage), K_TLM(:,istage,itlm)) :1071-1097
          double* Kc = Kt(istage - 1, itlm);
          daxpy(N, cfg.FDincrement, S_TLM.data(), TMP.data());
          mdl.fun(T + tb.C[istage - 1] * H, TMP.data(), Kc);
          ISTATUS[Nfun]++;
          daxpy(N, -1.0, Ki, Kc);
          { const double a = 1.0 / cfg.FDincrement; for (int i = 0; i < N; ++i) Kc[i] = a * Kc[i]; }
        }
      }
      ISTATUS[Nstp]++;
      for (int i = 0; i < N; ++i) TMP[i] = 0.0;
      for (int i = 1; i <= S; ++i)
        if (tb.E[i - 1] != 0.0) daxpy(N, H * tb.E[i - 1], K.data() + (size_t)(i - 1) * N, TMP.data());
      double Err = error_norm(TMP.data(), SCAL.data());
      if (NTLM > 0 && cfg.TLMtruncErr) {
        for (size_t q = 0; q < (size_t)N * NTLM; ++q) Yerr_TLM[q] = 0.0;
        for (int itlm = 0; itlm < NTLM; ++itlm)
          for (int j = 1; j <= S; ++j)
            if (tb.E[j - 1] != 0.0) daxpy(N, H * tb.E[j - 1], Kt(j - 1, itlm), Yerr_TLM.data() + (size_t)itlm * N);
        for (int itlm = 0; itlm < NTLM; ++itlm) {   // ERK_ErrorNorm_TLM: scaled by the error vector itself (:619-624)
          const double* Ye = Yerr_TLM.data() + (size_t)itlm * N;
          for (int i = 0; i < N; ++i)
            SCAL_TLM[i] = (cfg.ITOL == 0) ? 1.0 / (AbsTol_TLM[(size_t)itlm * N] + RelTol_TLM[(size_t)itlm * N] * std::fabs(Ye[i]))
                                          : 1.0 / (AbsTol_TLM[(size_t)itlm * N + i] + RelTol_TLM[(size_t)itlm * N + i] * std::fabs(Ye[i]));
          Err = std::fmax(Err, error_norm(Ye, SCAL_TLM.data()));
        }
      }
      double Fac = cfg.FacSafe * o_pow_neg_inv(Err, tb.ELO);
      Fac = std::fmax(cfg.FacMin, std::fmin(cfg.FacMax, Fac));
      double Hnew = H * Fac;
      if (Err < 1.0) {
        FirstStep = false;
        ISTATUS[Nacc]++;
        T = T + H;
        for (int i = 1; i <= S; ++i) {
          if (tb.B[i - 1] != 0.0) daxpy(N, H * tb.B[i - 1], K.data() + (size_t)(i - 1) * N, Y);
          for (int itlm = 0; itlm < NTLM; ++itlm) daxpy(N, H * tb.B[i - 1], Kt(i - 1, itlm), Y_TLM + (size_t)itlm * N);
        }
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
};

// INTEGRATE of FWD/ERK (NTLM = 0, Y_TLM = NULL) and INTEGRATE_TLM of TLM/ERK_TLM
template <class Model>
int erk_integrate(Model& mdl, int NTLM, double* Y, double* Y_TLM, double TIN, double TOUT, const double* ATOL_TLM, const double* RTOL_TLM,
                  const double* ATOL, const double* RTOL, const int* ICNTRL_U, const double* RCNTRL_U, int* ISTATUS_U, double* RSTATUS_U) {
  int ICNTRL[20] = {0};
  double RCNTRL[20] = {0};
  if (ICNTRL_U) for (int i = 0; i < 20; ++i) if (ICNTRL_U[i] > 0) ICNTRL[i] = ICNTRL_U[i];
  if (RCNTRL_U) for (int i = 0; i < 20; ++i) if (RCNTRL_U[i] > 0) RCNTRL[i] = RCNTRL_U[i];
  Erk<Model> s(mdl);
  const bool tlm = NTLM > 0;
  int Ierr = erk_parse(Model::N, ICNTRL, RCNTRL, TIN, TOUT, ATOL, RTOL, tlm, s.cfg);
  if (Ierr >= 0 && tlm && !s.cfg.FDAprox && !std::is_same<Model, Swe>::value) Ierr = -100;   // dense-Jacobian TLM mode: shallow-water model only
  if (Ierr >= 0 && tlm && s.cfg.TLMtruncErr && (!ATOL_TLM || !RTOL_TLM)) Ierr = -100;
  if (Ierr >= 0) Ierr = s.integrate(NTLM, Y, Y_TLM, TIN, TOUT, ATOL, RTOL, ATOL_TLM, RTOL_TLM);
  if (ISTATUS_U) for (int i = 0; i < 20; ++i) ISTATUS_U[i] = s.ISTATUS[i];
  if (RSTATUS_U) for (int i = 0; i < 20; ++i) RSTATUS_U[i] = s.RSTATUS[i];
  return Ierr;
}

}  // namespace oracle
