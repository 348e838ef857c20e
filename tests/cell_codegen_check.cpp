// TEST INFRASTRUCTURE.  Host-side bitwise check of the generated per-thread CBM-IV code
// (fatode_b200/csrc/gen/cbm4_cell_gen.h) against the oracle along real trajectories:
// every DGETF2 / DGETRS('N') call the oracle makes while integrating perturbed cells is
// replayed through cbm4_cell::lu_factor / lu_solve and compared bit for bit; FUN and the E
// matrix are compared at every step start.  Built and run by tests/test_cell_codegen.py.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <cmath>

static double g_rc[81], g_hess[258];
#define CBM4_CELL_FN static inline
#define CBM4_CELL_INL static inline
#define CBM4_CELL_RC(r) g_rc[r]
#define CBM4_CELL_HESS(h) g_hess[h]
#define CBM4_CELL_ST4(p, a, b, c, d) do { (p)[0] = (a); (p)[1] = (b); (p)[2] = (c); (p)[3] = (d); } while (0)
#include "../fatode_b200/csrc/gen/cbm4_cell_gen.h"

static long n_dom = 0, n_lu = 0, n_lu_bad = 0, n_irregular = 0, n_solve = 0, n_solve_bad = 0, n_swaps[4] = {0, 0, 0, 0};
static double cur_lu[cbm4_cell::NLU];
static double cur_dinv[32];
static unsigned cur_swaps = 0;
static bool cur_ok = false;
static double max_terr = 0.0;

static bool same(double a, double b) { return a == b || (std::isnan(a) && std::isnan(b)); }   // -0 == +0 accepted

// wrap the oracle's LAPACK restatement
#define dgetf2 dgetf2_oracle
#define dgetrs dgetrs_oracle
#define zgetf2 zgetf2_oracle
#define zgetrs_n zgetrs_n_oracle
#define zgetrs_t zgetrs_t_oracle
#include "../oracle/ref_lapack.hpp"
#undef dgetf2
#undef dgetrs
#undef zgetf2
#undef zgetrs_n
#undef zgetrs_t
static double max_zterr = 0.0;
static long n_zlu = 0, n_zlu_bad = 0, n_zirregular = 0, n_zsolve = 0, n_zsolve_bad = 0;
static double cur_zr[cbm4_cell::NLU], cur_zi[cbm4_cell::NLU];
static unsigned cur_zswaps = 0;
static bool cur_zok = false;
namespace oracle {
static inline int dgetf2(int n, double* a, int* ipiv) {
  using namespace cbm4_cell;
  // E must live on the union pattern
  bool inpat[32][32];
  std::memset(inpat, 0, sizeof inpat);
  for (int i = 0; i < 32; ++i)
    for (int e = LU_ROWPTR[i]; e < LU_ROWPTR[i + 1]; ++e) { inpat[i][LU_COL[e]] = true; cur_lu[e] = a[i + 32 * LU_COL[e]]; }
  for (int i = 0; i < 32; ++i)
    for (int c = 0; c < 32; ++c)
      if (!inpat[i][c] && a[i + 32 * c] != 0.0) { std::printf("E has an entry outside the pattern (%d,%d)\n", i, c); std::exit(2); }
  int info = dgetf2_oracle(n, a, ipiv);
  double* const dinv = cur_dinv;
  int bad = lu_factor<1>(cur_lu, dinv, cur_swaps);
  ++n_lu;
  // expected verdict: regular iff every interchange is one of the allowed ones
  bool regular = (info == 0);
  unsigned exp_sw = 0;
  for (int k = 0; k < 32; ++k)
    if (ipiv[k] != k + 1) {
      bool ok = false;
      for (int b = 0; b < NSWAP; ++b)
        if (SWAP_COL[b] == k && SWAP_ROW[b] == ipiv[k] - 1) { ok = true; exp_sw |= 1u << b; }
      if (!ok) regular = false;
    }
  cur_ok = (bad == 0);
  if (!regular) {
    ++n_irregular;
    if (bad == 0) { ++n_lu_bad; std::printf("lu_factor accepted an irregular pivot sequence\n"); }
    return info;
  }
  if (bad != 0) { ++n_lu_bad; std::printf("lu_factor rejected a regular factorisation (lu %ld)\n", n_lu); return info; }
  if (cur_swaps != exp_sw) { ++n_lu_bad; std::printf("swap flags %u != %u\n", cur_swaps, exp_sw); return info; }
  for (int b = 0; b < NSWAP; ++b) if (cur_swaps & (1u << b)) ++n_swaps[b];
  int nbad = 0;
  for (int i = 0; i < 32; ++i)
    for (int c = 0; c < 32; ++c) {
      const double ref = a[i + 32 * c];
      if (inpat[i][c]) {
        int e = -1;
        for (int q = LU_ROWPTR[i]; q < LU_ROWPTR[i + 1]; ++q) if (LU_COL[q] == c) e = q;
        if (!same(cur_lu[e], ref)) ++nbad;
      } else if (ref != 0.0) ++nbad;
    }
  for (int k = 0; k < 32; ++k) if (!same(dinv[k], 1.0 / a[k + 32 * k])) ++nbad;
  if (nbad) { ++n_lu_bad; std::printf("LU %ld: %d entries differ\n", n_lu, nbad); }
  if ((n_lu & 15) == 0) {   // transposed solve (backward sweep form) against DGETRS('T'), to rounding
    double xr[32], xn[32];
    for (int i = 0; i < 32; ++i) xr[i] = xn[i] = std::sin(1.0 + i * 0.37 + (double)(n_lu % 97));
    dgetrs_oracle(true, n, a, ipiv, xr);
    lut_solve(cur_lu, dinv, cur_swaps, xn);
    if (cur_swaps == LUT_DOM_SWAPS) {   // the dominant-outcome code must give the very same bits
      double xd[32];
      for (int i = 0; i < 32; ++i) xd[i] = std::sin(1.0 + i * 0.37 + (double)(n_lu % 97));
      lut_solve_dom(cur_lu, dinv, xd);
      ++n_dom;
      for (int i = 0; i < 32; ++i) if (!same(xd[i], xn[i])) { ++n_solve_bad; break; }
    }
    double sc = 0.0, er = 0.0;
    for (int i = 0; i < 32; ++i) { sc = std::fmax(sc, std::fabs(xr[i])); er = std::fmax(er, std::fabs(xr[i] - xn[i])); }
    max_terr = std::fmax(max_terr, er / sc);
    {   // reference-order transposed solve: bit for bit
      double ys[32], yref[32];
      for (int i = 0; i < 32; ++i) ys[i] = yref[i] = std::cos(1.0 + i * 0.77 + (double)(n_lu % 31));
      dgetrs_oracle(true, n, a, ipiv, yref);
      lut_solve_seq(cur_lu, dinv, cur_swaps, ys);
      for (int i = 0; i < 32; ++i) if (!same(ys[i], yref[i])) { ++n_solve_bad; break; }
    }
    {   // reference-order non-transposed register solve: bit for bit
      double ys[32], yref[32];
      for (int i = 0; i < 32; ++i) ys[i] = yref[i] = std::sin(2.0 + i * 0.41 + (double)(n_lu % 29));
      dgetrs_oracle(false, n, a, ipiv, yref);
      lun_solve_seq(cur_lu, dinv, cur_swaps, ys);
      for (int i = 0; i < 32; ++i) if (!same(ys[i], yref[i])) { ++n_solve_bad; break; }
    }
    // non-transposed register solve of the tangent-linear sweep against DGETRS('N')
    for (int i = 0; i < 32; ++i) xr[i] = xn[i] = std::cos(0.5 + i * 0.61 + (double)(n_lu % 89));
    dgetrs_oracle(false, n, a, ipiv, xr);
    lun_solve(cur_lu, dinv, cur_swaps, xn);
    sc = 0.0; er = 0.0;
    for (int i = 0; i < 32; ++i) { sc = std::fmax(sc, std::fabs(xr[i])); er = std::fmax(er, std::fabs(xr[i] - xn[i])); }
    max_terr = std::fmax(max_terr, er / sc);
  }
  return info;
}
static inline void dgetrs(bool trans, int n, const double* a, const int* ipiv, double* b) {
  if (trans || !cur_ok) { dgetrs_oracle(trans, n, a, ipiv, b); return; }
  double x[32];
  std::memcpy(x, b, sizeof x);
  dgetrs_oracle(trans, n, a, ipiv, b);
  double xr[32];
  std::memcpy(xr, x, sizeof xr);
  cbm4_cell::lu_solve<1>(cur_lu, cur_swaps, x);
  ++n_solve;
  for (int i = 0; i < 32; ++i) if (!same(x[i], b[i])) { ++n_solve_bad; break; }
  // the register form of the tensor-memory forward kernel (quotients through exact_div_safe): every solve of the trajectory
  cbm4_cell::lu_solve_reg<1>(cur_lu, cur_dinv, cur_swaps, xr);
  for (int i = 0; i < 32; ++i) if (!same(xr[i], b[i])) { ++n_solve_bad; break; }
}
// ZGETF2 / ZGETRS('N') of the implicit Runge-Kutta families through cbm4_cell::zlu_factor / zlu_solve
static inline int zgetf2(int n, cplx* a, int* ipiv) {
  using namespace cbm4_cell;
  bool inpat[32][32];
  std::memset(inpat, 0, sizeof inpat);
  for (int i = 0; i < 32; ++i)
    for (int e = LU_ROWPTR[i]; e < LU_ROWPTR[i + 1]; ++e) {
      inpat[i][LU_COL[e]] = true; cur_zr[e] = a[i + 32 * LU_COL[e]].re; cur_zi[e] = a[i + 32 * LU_COL[e]].im;
    }
  int info = zgetf2_oracle(n, a, ipiv);
  int bad = zlu_factor<1>(cur_zr, cur_zi, cur_zswaps);
  ++n_zlu;
  bool regular = (info == 0);
  unsigned exp_sw = 0;
  for (int k = 0; k < 32; ++k)
    if (ipiv[k] != k + 1) {
      bool ok = false;
      for (int b = 0; b < NSWAP; ++b)
        if (SWAP_COL[b] == k && SWAP_ROW[b] == ipiv[k] - 1) { ok = true; exp_sw |= 1u << b; }
      if (!ok) regular = false;
    }
  cur_zok = (bad == 0);
  if (!regular) {
    ++n_zirregular;
    if (bad == 0) { ++n_zlu_bad; std::printf("zlu_factor accepted an irregular pivot sequence\n"); }
    return info;
  }
  if (bad != 0) { ++n_zlu_bad; std::printf("zlu_factor rejected a regular factorisation (zlu %ld)\n", n_zlu); return info; }
  if (cur_zswaps != exp_sw) { ++n_zlu_bad; std::printf("complex swap flags %u != %u\n", cur_zswaps, exp_sw); return info; }
  int nbad = 0;
  for (int i = 0; i < 32; ++i)
    for (int c = 0; c < 32; ++c) {
      const cplx ref = a[i + 32 * c];
      if (inpat[i][c]) {
        int e = -1;
        for (int q = LU_ROWPTR[i]; q < LU_ROWPTR[i + 1]; ++q) if (LU_COL[q] == c) e = q;
        if (!same(cur_zr[e], ref.re) || !same(cur_zi[e], ref.im)) ++nbad;
      } else if (ref.re != 0.0 || ref.im != 0.0) ++nbad;
    }
  if (nbad) { ++n_zlu_bad; std::printf("ZLU %ld: %d entries differ\n", n_zlu, nbad); }
  if ((n_zlu & 15) == 0) {   // transposed complex solve of the RK adjoint sweep against ZGETRS('T'), to rounding
    cplx xr_[32];
    double xr[32], xi[32], zdr[32], zdi[32];
    for (int i = 0; i < 32; ++i) {
      xr[i] = std::sin(1.0 + i * 0.37 + (double)(n_zlu % 97)); xi[i] = std::cos(0.3 + i * 0.91 + (double)(n_zlu % 53));
      xr_[i] = cplx{xr[i], xi[i]};
      const cplx d = cdiv(cplx{1.0, 0.0}, a[i + 32 * i]);
      zdr[i] = d.re; zdi[i] = d.im;
    }
    zgetrs_t_oracle(n, a, ipiv, xr_);
    zlut_solve(cur_zr, cur_zi, zdr, zdi, cur_zswaps, xr, xi);
    double sc = 0.0, er = 0.0;
    for (int i = 0; i < 32; ++i) {
      sc = std::fmax(sc, std::fabs(xr_[i].re) + std::fabs(xr_[i].im));
      er = std::fmax(er, std::fabs(xr_[i].re - xr[i]) + std::fabs(xr_[i].im - xi[i]));
    }
    max_zterr = std::fmax(max_zterr, er / sc);
    // reference-order variant: bit for bit
    double yr[32], yi[32];
    cplx ref2[32];
    for (int i = 0; i < 32; ++i) { yr[i] = std::cos(2.0 + i * 0.53); yi[i] = std::sin(0.1 + i * 1.7); ref2[i] = cplx{yr[i], yi[i]}; }
    zgetrs_t_oracle(n, a, ipiv, ref2);
    double zrt[32], zdv[32], zrd[32];
    const unsigned zmask = zdiag_prep<1>(cur_zr, cur_zi, zrt, zdv, zrd);
    zlut_solve_seq(cur_zr, cur_zi, zrt, zdv, zrd, zmask, cur_zswaps, yr, yi);
    for (int i = 0; i < 32; ++i) if (!same(yr[i], ref2[i].re) || !same(yi[i], ref2[i].im)) { ++n_zsolve_bad; break; }
  }
  return info;
}
static inline void zgetrs_t(int n, const cplx* a, const int* ipiv, cplx* b) { zgetrs_t_oracle(n, a, ipiv, b); }
static inline void zgetrs_n(int n, const cplx* a, const int* ipiv, cplx* b) {
  if (!cur_zok) { zgetrs_n_oracle(n, a, ipiv, b); return; }
  double xr[32], xi[32];
  for (int i = 0; i < 32; ++i) { xr[i] = b[i].re; xi[i] = b[i].im; }
  zgetrs_n_oracle(n, a, ipiv, b);
  cbm4_cell::zlu_solve<1>(cur_zr, cur_zi, cur_zswaps, xr, xi);
  ++n_zsolve;
  for (int i = 0; i < 32; ++i) if (!same(xr[i], b[i].re) || !same(xi[i], b[i].im)) { ++n_zsolve_bad; break; }
}
}  // namespace oracle
#include "../oracle/models.hpp"
#include "../oracle/ros.hpp"
#include "../oracle/rk.hpp"
using namespace oracle;

static uint64_t splitmix(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull; x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

int main(int argc, char** argv) {
  const int ncells = argc > 1 ? std::atoi(argv[1]) : 6;
  const double hours = argc > 2 ? std::atof(argv[2]) : 72.0;
  const int family_adj = argc > 3 ? std::atoi(argv[3]) : 1;
  double y0[32];
  Cbm4::initial_state(y0);
  const double F01 = (double)0.1f;
  double rtol[32], atol[32];
  for (int i = 0; i < 32; ++i) { rtol[i] = F01 * 1e-4; atol[i] = F01 * 1e-6; }
  int ic[20] = {0};
  double rc[20] = {0};
  ic[2] = 5;
  long bad_fun = 0, n_fun = 0;
  for (int c = 0; c < ncells; ++c) {
    double y[32];
    for (int i = 0; i < 32; ++i) {
      double u = (double)(splitmix((uint64_t)c * 32 + i + 777) >> 11) / 9007199254740992.0 * 2 - 1;
      y[i] = y0[i] * (c ? 1 + 0.1 * u : 1.0);
    }
    // FUN / E / adjoint helper values at the start state and a few times
    Cbm4 m;
    for (int q = 0; q < 4; ++q) {
      const double t = 43200.0 + q * 17000.0;
      double f_ref[32], f_new[32], jac[1024], PH[cbm4_cell::NPHOTO];
      m.fun(t, y, f_ref);
      for (int r = 0; r < 81; ++r) g_rc[r] = m.rconst[r];
      for (int k = 0; k < cbm4_cell::NPHOTO; ++k) PH[k] = m.rconst[cbm4_cell::PHOTO_IDX[k]];
      cbm4_cell::fun<1>(y, m.fix[0], PH, f_new);
      f_new[30] = f_new[30] + 2.0e3; f_new[25] = f_new[25] + 2.0e3;
      ++n_fun;
      for (int i = 0; i < 32; ++i) if (!same(f_ref[i], f_new[i])) { ++bad_fun; break; }
      m.jac(t, y, jac);
      double lu[cbm4_cell::NLU], jv[cbm4_cell::NJ + 4];
      const double ghinv = 1.0 / (37.0 * 0.25);
      cbm4_cell::build_E<1>(y, m.fix[0], PH, ghinv, lu);
      for (int i = 0; i < 32; ++i)
        for (int e = cbm4_cell::LU_ROWPTR[i]; e < cbm4_cell::LU_ROWPTR[i + 1]; ++e) {
          const int cc = cbm4_cell::LU_COL[e];
          double ref = -jac[i + 32 * cc];
          if (i == cc) ref = ref + ghinv;
          if (!same(ref, lu[e])) { ++bad_fun; }
        }
      cbm4_cell::jac_values<1>(y, m.fix[0], PH, jv);
      {   // J^T x of the RK adjoint sweep against the dense product
        double x[32], out[32];
        for (int i = 0; i < 32; ++i) x[i] = std::sin(0.7 + 1.3 * i + q);
        cbm4_cell::jt_vec(jv, x, out);
        double sc = 0.0, er = 0.0;
        for (int cidx = 0; cidx < 32; ++cidx) {
          double sref = 0.0;
          for (int r = 0; r < 32; ++r) sref += jac[r + 32 * cidx] * x[r];
          sc = std::fmax(sc, std::fabs(sref)); er = std::fmax(er, std::fabs(sref - out[cidx]));
        }
        if (!(er <= 1e-12 * sc)) ++bad_fun;
        cbm4_cell::j_vec_seq(jv, x, out);         // matmul(fjac1, g) in column-sweep order = sequential row sums
        for (int r = 0; r < 32; ++r) {
          double sref = 0.0;
          for (int cidx = 0; cidx < 32; ++cidx) sref += jac[r + 32 * cidx] * x[cidx];
          if (!same(sref, out[r])) ++bad_fun;
        }
        cbm4_cell::jt_vec_seq(jv, x, out);        // reference order: bit for bit with the dense sequential dot products
        for (int cidx = 0; cidx < 32; ++cidx) {
          double sref = 0.0;
          for (int r = 0; r < 32; ++r) sref += jac[r + 32 * cidx] * x[r];
          if (!same(sref, out[cidx])) ++bad_fun;
        }
      }
      // entry order: by column then row over the structural non-zeros of CBM4_JAC
      double hs_ref[258], hs_new[258];
      cbm4_gen::hessian(y, m.fix, m.rconst, hs_ref);
      cbm4_cell::hess_values<1>(y, m.fix[0], PH, hs_new);
      for (int h = 0; h < 258; ++h) if (!same(hs_ref[h], hs_new[h])) ++bad_fun;
    }
    int is[20];
    double rs[20];
    int e;
    if (family_adj == 2) {
      // FWD/RK (real + complex factorisations): ICNTRL(3) cycles through the five methods
      int icr[20] = {0};
      const int meth[5] = {1, 2, 3, 4, 5};
      icr[2] = meth[c % 5];
      double rt[32], at[32];
      for (int i = 0; i < 32; ++i) { rt[i] = F01 * 1e-5; at[i] = F01 * 1e-7; }
      e = rk_integrate<Cbm4>(m, 43200.0, 43200.0 + hours * 3600.0, y, rt, at, icr, rc, is, rs);
    } else if (family_adj) {
      // ADJ-family forward sweep (DOUBLE PRECISION tableau), adjoint skipped: ICNTRL(7)=1 is "no adjoint"
      double lam[32 * 1];
      std::memset(lam, 0, sizeof lam);
      ic[6] = 1;
      e = ros_integrate_adj<Cbm4>(m, 0, 1, y, lam, nullptr, 43200.0, 43200.0 + hours * 3600.0, atol, rtol, ic, rc, is, rs);
    } else {
      e = ros_integrate<Cbm4>(m, 43200.0, 43200.0 + hours * 3600.0, y, rtol, atol, ic, rc, is, rs);
    }
    if (e != 1) { std::printf("oracle integration failed %d\n", e); return 3; }
  }
  std::printf("cells %d LUs %ld irregular %ld lu_mismatch %ld solves %ld solve_mismatch %ld fun_checks %ld fun_mismatch %ld swaps %ld %ld %ld %ld max_transposed_solve_err %.2e\n",
              ncells, n_lu, n_irregular, n_lu_bad, n_solve, n_solve_bad, n_fun, bad_fun, n_swaps[0], n_swaps[1], n_swaps[2],
              n_swaps[3], max_terr);
  if (family_adj == 2)
    std::printf("complex: LUs %ld irregular %ld lu_mismatch %ld solves %ld solve_mismatch %ld max_transposed_err %.2e\n", n_zlu,
                n_zirregular, n_zlu_bad, n_zsolve, n_zsolve_bad, max_zterr);
  if (family_adj == 2 && !(max_zterr < 1e-9)) return 5;
  if (family_adj == 2 && (n_zlu == 0 || n_zsolve == 0)) return 4;
  return (n_lu_bad || n_solve_bad || bad_fun || !(max_terr < 1e-9) || n_zlu_bad || n_zsolve_bad) ? 1 : 0;
}
