// ORACLE — TEST INFRASTRUCTURE ONLY.  C entry points (ctypes / bench cpu_baseline) over the
// C++ restatement in ros.hpp / models.hpp.  Cells are independent; OpenMP loops over them the way
// a user of the (serial, single-system) reference would loop over INTEGRATE_ADJ calls.
// Layout: cell-major ("array of systems"): y[cell][N], lambda[cell][NADJ][N] (= Fortran
// Lambda(N,NADJ) per cell), mu[cell][NADJ][NP], istatus[cell][20], rstatus[cell][20].
#include <omp.h>
#include <cstring>
#include <vector>
#include "models.hpp"
#include "ros.hpp"
#include "sdirk.hpp"
#include "rk.hpp"
#include "rk_adj.hpp"
#include "sdirk_tlm.hpp"
#include "rk_tlm.hpp"
#include "erk.hpp"
#include "erk_adj.hpp"
#include "sdirk_adj.hpp"

using namespace oracle;

enum { MODEL_ROBERTS = 1, MODEL_SMALL_STRATO = 2, MODEL_CBM4 = 3, MODEL_SWE = 4 };

template <class M>
static void adj_batch(long ncells, int np, int nadj, double* y, double* lambda, double* mu, double tin, double tout,
                      const double* atol, const double* rtol, const int* icntrl, const double* rcntrl, int* istatus,
                      double* rstatus, int* ierr, int nthreads) {
  const int N = M::N;
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
  for (long c = 0; c < ncells; ++c) {
    M mdl;
    ierr[c] = ros_integrate_adj<M>(mdl, np, nadj, y + c * N, lambda + c * (long)N * nadj,
                                   (np > 0 && mu) ? mu + c * (long)np * nadj : nullptr, tin, tout, atol, rtol, icntrl,
                                   rcntrl, istatus + c * 20, rstatus + c * 20);
  }
}

template <class M>
static void rk_adj_batch(long ncells, int np, int nadj, double* y, double* lambda, double* mu, double tin, double tout,
                         const double* atol, const double* rtol, const int* icntrl, const double* rcntrl, int* istatus,
                         double* rstatus, int* ierr, int nthreads) {
  const int N = M::N;
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
  for (long c = 0; c < ncells; ++c) {
    M mdl;
    ierr[c] = rk_adj_integrate<M>(mdl, np, nadj, y + c * N, lambda + c * (long)N * nadj,
                                  (np > 0 && mu) ? mu + c * (long)np * nadj : nullptr, tin, tout, atol, rtol, icntrl,
                                  rcntrl, istatus + c * 20, rstatus + c * 20);
  }
}

template <class M>
static void fwd_batch(long ncells, double* y, double tin, double tout, const double* atol, const double* rtol,
                      const int* icntrl, const double* rcntrl, int* istatus, double* rstatus, int* ierr, int nthreads) {
  const int N = M::N;
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
  for (long c = 0; c < ncells; ++c) {
    M mdl;
    ierr[c] = ros_integrate<M>(mdl, tin, tout, y + c * N, rtol, atol, icntrl, rcntrl, istatus + c * 20, rstatus + c * 20);
  }
}

extern "C" {

int oracle_max_threads() { return omp_get_max_threads(); }
// the correctly rounded kernels of portable_math.hpp themselves (tests: against a multi-precision reference, against the device)
void oracle_probe_math(int n, const double* x, double elo, double* out_cos, double* out_root) {
  for (int i = 0; i < n; ++i) { out_cos[i] = oracle_pm::cos_cr(x[i]); out_root[i] = oracle_pm::root_elo(std::fabs(x[i]), elo); }
}
int oracle_set_libm(int on) { int old = libm_mode(); libm_mode() = on ? 1 : 0; return old; }

int oracle_model_dims(int model, int* n, int* np) {
  switch (model) {
    case MODEL_ROBERTS: *n = Roberts::N; *np = Roberts::NP; return 0;
    case MODEL_SMALL_STRATO: *n = SmallStrato::N; *np = SmallStrato::NP; return 0;
    case MODEL_CBM4: *n = Cbm4::N; *np = Cbm4::NP; return 0;
  }
  return -1;
}

int oracle_initial_state(int model, double* y) {
  switch (model) {
    case MODEL_ROBERTS: Roberts::initial_state(y); return 0;
    case MODEL_SMALL_STRATO: SmallStrato::initial_state(y); return 0;
    case MODEL_CBM4: Cbm4::initial_state(y); return 0;
  }
  return -1;
}

// FWD/ROS INTEGRATE looped over cells
int oracle_ros_fwd_batch(int model, long ncells, double* y, double tin, double tout, const double* atol,
                         const double* rtol, const int* icntrl, const double* rcntrl, int* istatus, double* rstatus,
                         int* ierr, int nthreads) {
  if (nthreads <= 0) nthreads = omp_get_max_threads();
  switch (model) {
    case MODEL_ROBERTS: fwd_batch<Roberts>(ncells, y, tin, tout, atol, rtol, icntrl, rcntrl, istatus, rstatus, ierr, nthreads); return 0;
    case MODEL_SMALL_STRATO: fwd_batch<SmallStrato>(ncells, y, tin, tout, atol, rtol, icntrl, rcntrl, istatus, rstatus, ierr, nthreads); return 0;
    case MODEL_CBM4: fwd_batch<Cbm4>(ncells, y, tin, tout, atol, rtol, icntrl, rcntrl, istatus, rstatus, ierr, nthreads); return 0;
  }
  return -1;
}

// ADJ/ROS_ADJ INTEGRATE_ADJ looped over cells (mu == NULL or np == 0 -> RosenbrockADJ2)
int oracle_ros_adj_batch(int model, long ncells, int np, int nadj, double* y, double* lambda, double* mu, double tin,
                         double tout, const double* atol, const double* rtol, const int* icntrl, const double* rcntrl,
                         int* istatus, double* rstatus, int* ierr, int nthreads) {
  if (nthreads <= 0) nthreads = omp_get_max_threads();
  switch (model) {
    case MODEL_ROBERTS: adj_batch<Roberts>(ncells, np, nadj, y, lambda, mu, tin, tout, atol, rtol, icntrl, rcntrl, istatus, rstatus, ierr, nthreads); return 0;
    case MODEL_SMALL_STRATO: adj_batch<SmallStrato>(ncells, np, nadj, y, lambda, mu, tin, tout, atol, rtol, icntrl, rcntrl, istatus, rstatus, ierr, nthreads); return 0;
    case MODEL_CBM4: adj_batch<Cbm4>(ncells, np, nadj, y, lambda, mu, tin, tout, atol, rtol, icntrl, rcntrl, istatus, rstatus, ierr, nthreads); return 0;
  }
  return -1;
}

// INTEGRATE_ADJ of ADJ/RK_ADJ looped over cells (same layouts as oracle_ros_adj_batch)
int oracle_rk_adj_batch(int model, long ncells, int np, int nadj, double* y, double* lambda, double* mu, double tin,
                        double tout, const double* atol, const double* rtol, const int* icntrl, const double* rcntrl,
                        int* istatus, double* rstatus, int* ierr, int nthreads) {
  if (nthreads <= 0) nthreads = omp_get_max_threads();
  switch (model) {
    case MODEL_ROBERTS: rk_adj_batch<Roberts>(ncells, np, nadj, y, lambda, mu, tin, tout, atol, rtol, icntrl, rcntrl, istatus, rstatus, ierr, nthreads); return 0;
    case MODEL_SMALL_STRATO: rk_adj_batch<SmallStrato>(ncells, np, nadj, y, lambda, mu, tin, tout, atol, rtol, icntrl, rcntrl, istatus, rstatus, ierr, nthreads); return 0;
    case MODEL_CBM4: rk_adj_batch<Cbm4>(ncells, np, nadj, y, lambda, mu, tin, tout, atol, rtol, icntrl, rcntrl, istatus, rstatus, ierr, nthreads); return 0;
  }
  return -1;
}

// INTEGRATE_TLM of TLM/ROS_TLM looped over cells (cbm4 only: the model needs a HESS_VEC callback).
// y_tlm[cell][NTLM][N] = Y_tlm(N,NTLM) per cell; atol_tlm/rtol_tlm [NTLM][N] shared by the batch (may be NULL when
// ICNTRL(12) = 0).  icntrl == NULL means "ICNTRL_U not present" (the presets apply).
int oracle_ros_tlm_batch(int model, long ncells, int ntlm, double* y, double* y_tlm, double tin, double tout,
                         const double* atol_tlm, const double* rtol_tlm, const double* atol, const double* rtol,
                         const int* icntrl, const double* rcntrl, int* istatus, double* rstatus, int* ierr, int nthreads) {
  if (model != MODEL_CBM4) return -1;
  if (nthreads <= 0) nthreads = omp_get_max_threads();
  const int N = Cbm4::N;
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
  for (long c = 0; c < ncells; ++c) {
    Cbm4 mdl;
    ierr[c] = ros_integrate_tlm<Cbm4>(mdl, ntlm, y + c * N, y_tlm + c * (long)N * ntlm, tin, tout, atol_tlm, rtol_tlm, atol, rtol,
                                      icntrl, rcntrl, istatus + c * 20, rstatus + c * 20);
  }
  return 0;
}

// INTEGRATE_TLM of TLM/SDIRK_TLM looped over cells (layouts as oracle_ros_tlm_batch; any model: no Hessian needed)
int oracle_sdirk_tlm_batch(int model, long ncells, int ntlm, double* y, double* y_tlm, double tin, double tout,
                           const double* atol_tlm, const double* rtol_tlm, const double* atol, const double* rtol,
                           const int* icntrl, const double* rcntrl, int* istatus, double* rstatus, int* ierr, int nthreads) {
  if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
  for (long c = 0; c < ncells; ++c) {
    switch (model) {
      case MODEL_ROBERTS: { Roberts m; ierr[c] = sdirk_tlm_integrate<Roberts>(m, ntlm, y + c * Roberts::N, y_tlm + c * (long)Roberts::N * ntlm, tin, tout, atol_tlm, rtol_tlm, atol, rtol, icntrl, rcntrl, istatus + c * 20, rstatus + c * 20); break; }
      case MODEL_SMALL_STRATO: { SmallStrato m; ierr[c] = sdirk_tlm_integrate<SmallStrato>(m, ntlm, y + c * SmallStrato::N, y_tlm + c * (long)SmallStrato::N * ntlm, tin, tout, atol_tlm, rtol_tlm, atol, rtol, icntrl, rcntrl, istatus + c * 20, rstatus + c * 20); break; }
      case MODEL_CBM4: { Cbm4 m; ierr[c] = sdirk_tlm_integrate<Cbm4>(m, ntlm, y + c * Cbm4::N, y_tlm + c * (long)Cbm4::N * ntlm, tin, tout, atol_tlm, rtol_tlm, atol, rtol, icntrl, rcntrl, istatus + c * 20, rstatus + c * 20); break; }
      default: ierr[c] = -100;
    }
  }
  return 0;
}

// INTEGRATE_TLM of TLM/RK_TLM looped over cells (preset direct TLM solve); y_tlm[cell][NTLM][N]
int oracle_rk_tlm_batch(int model, long ncells, int ntlm, double* y, double* y_tlm, double tin, double tout, const double* atol_tlm,
                        const double* rtol_tlm, const double* atol, const double* rtol, const int* icntrl, const double* rcntrl,
                        int* istatus, double* rstatus, int* ierr, int nthreads) {
  (void)atol_tlm; (void)rtol_tlm;
  if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
  for (long c = 0; c < ncells; ++c) {
    switch (model) {
      case MODEL_ROBERTS: { Roberts m; ierr[c] = rk_tlm_integrate<Roberts>(m, ntlm, y + c * Roberts::N, y_tlm + c * (long)Roberts::N * ntlm, tin, tout, atol, rtol, icntrl, rcntrl, istatus + c * 20, rstatus + c * 20); break; }
      case MODEL_SMALL_STRATO: { SmallStrato m; ierr[c] = rk_tlm_integrate<SmallStrato>(m, ntlm, y + c * SmallStrato::N, y_tlm + c * (long)SmallStrato::N * ntlm, tin, tout, atol, rtol, icntrl, rcntrl, istatus + c * 20, rstatus + c * 20); break; }
      case MODEL_CBM4: { Cbm4 m; ierr[c] = rk_tlm_integrate<Cbm4>(m, ntlm, y + c * Cbm4::N, y_tlm + c * (long)Cbm4::N * ntlm, tin, tout, atol, rtol, icntrl, rcntrl, istatus + c * 20, rstatus + c * 20); break; }
      default: ierr[c] = -100;
    }
  }
  return 0;
}

// INTEGRATE of FWD/ERK (ntlm = 0, y_tlm = NULL) / INTEGRATE_TLM of TLM/ERK_TLM looped over systems; model 4 = the
// shallow-water example (N = 4800): y[member][N], y_tlm[member][NTLM][N]
int oracle_erk_batch(int model, long ncells, int ntlm, double* y, double* y_tlm, double tin, double tout, const double* atol_tlm,
                     const double* rtol_tlm, const double* atol, const double* rtol, const int* icntrl, const double* rcntrl,
                     int* istatus, double* rstatus, int* ierr, int nthreads) {
  if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
  for (long c = 0; c < ncells; ++c) {
    switch (model) {
      case MODEL_SWE: { Swe m; ierr[c] = erk_integrate<Swe>(m, ntlm, y + c * Swe::N, y_tlm ? y_tlm + c * (long)Swe::N * ntlm : nullptr, tin, tout, atol_tlm, rtol_tlm, atol, rtol, icntrl, rcntrl, istatus + c * 20, rstatus + c * 20); break; }
      case MODEL_SMALL_STRATO: { SmallStrato m; ierr[c] = erk_integrate<SmallStrato>(m, ntlm, y + c * SmallStrato::N, y_tlm ? y_tlm + c * (long)SmallStrato::N * ntlm : nullptr, tin, tout, atol_tlm, rtol_tlm, atol, rtol, icntrl, rcntrl, istatus + c * 20, rstatus + c * 20); break; }
      default: ierr[c] = -100;
    }
  }
  return 0;
}
// INTEGRATE_ADJ of ADJ/ERK_ADJ (no parameters: ERKADJ2) for the shallow-water example; lambda[member][NADJ][N]
int oracle_erk_adj_batch(int model, long ncells, int nadj, double* y, double* lambda, double tin, double tout, const double* atol,
                         const double* rtol, const int* icntrl, const double* rcntrl, int* istatus, double* rstatus, int* ierr,
                         int nthreads) {
  if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
  for (long c = 0; c < ncells; ++c) {
    if (model != MODEL_SWE) { ierr[c] = -100; continue; }
    Swe m;
    ierr[c] = erk_adj_integrate_swe(m, nadj, y + c * Swe::N, lambda + c * (long)Swe::N * nadj, tin, tout, atol, rtol, icntrl, rcntrl,
                                    istatus + c * 20, rstatus + c * 20);
  }
  return 0;
}
// the JAC callback of the shallow-water drivers as a matrix-free product: out = JAC(y)^T x (trans != 0) or JAC(y) x
int oracle_swe_jac_mul(const double* y, const double* x, int trans, double* out) {
  Swe m;
  SweJac J;
  J.build(m, y);
  if (trans) J.mul_t(x, out); else J.mul_n(x, out);
  return 0;
}
// initializeGaussianGrid(A) + u = u0, v = v0 + grid2vec (EXAMPLES/swe/SWE_ERK/swe2D_erk_dr.F90:100-103)
int oracle_swe_initial_state(double A, double u0, double v0, double* vec) { Swe::initial_state(A, u0, v0, vec); return 0; }
int oracle_swe_fun(const double* y, double* p) { Swe m; m.fun(0.0, y, p); return 0; }

// INTEGRATE_ADJ of ADJ/SDIRK_ADJ looped over cells (layouts as oracle_ros_adj_batch)
int oracle_sdirk_adj_batch(int model, long ncells, int np, int nadj, double* y, double* lambda, double* mu, double tin, double tout,
                           const double* atol, const double* rtol, const int* icntrl, const double* rcntrl, int* istatus,
                           double* rstatus, int* ierr, int nthreads) {
  if (nthreads <= 0) nthreads = omp_get_max_threads();
  if (model != MODEL_CBM4) { for (long c = 0; c < ncells; ++c) ierr[c] = -100; return 0; }
  const int N = Cbm4::N;
  std::vector<double> tol_adj((size_t)2 * N * nadj);   // the drivers pass rtol_adj = rtol, atol_adj = atol per column
  for (int m = 0; m < nadj; ++m)
    for (int i = 0; i < N; ++i) { tol_adj[(size_t)m * N + i] = atol[i]; tol_adj[(size_t)(nadj + m) * N + i] = rtol[i]; }
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
  for (long c = 0; c < ncells; ++c) {
    Cbm4 m;
    ierr[c] = sdirk_adj_integrate<Cbm4>(m, np, nadj, y + c * N, lambda + c * (long)N * nadj, (np > 0 && mu) ? mu + c * (long)np * nadj : nullptr,
                                        tin, tout, tol_adj.data(), tol_adj.data() + (size_t)N * nadj, atol, rtol, icntrl, rcntrl,
                                        istatus + c * 20, rstatus + c * 20);
  }
  return 0;
}

// the same with the caller's ATOL_adj / RTOL_adj (N, NADJ) — read by the simplified-Newton adjoint stages (ICNTRL(7) /= 1)
int oracle_sdirk_adj_batch_tol(int model, long ncells, int np, int nadj, double* y, double* lambda, double* mu, double tin, double tout,
                               const double* atol_adj, const double* rtol_adj, const double* atol, const double* rtol, const int* icntrl,
                               const double* rcntrl, int* istatus, double* rstatus, int* ierr, int nthreads) {
  if (nthreads <= 0) nthreads = omp_get_max_threads();
  if (model != MODEL_CBM4) { for (long c = 0; c < ncells; ++c) ierr[c] = -100; return 0; }
  const int N = Cbm4::N;
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
  for (long c = 0; c < ncells; ++c) {
    Cbm4 m;
    ierr[c] = sdirk_adj_integrate<Cbm4>(m, np, nadj, y + c * N, lambda + c * (long)N * nadj, (np > 0 && mu) ? mu + c * (long)np * nadj : nullptr,
                                        tin, tout, atol_adj, rtol_adj, atol, rtol, icntrl, rcntrl, istatus + c * 20, rstatus + c * 20);
  }
  return 0;
}

// FWD/SDIRK INTEGRATE looped over cells
int oracle_sdirk_fwd_batch(int model, long ncells, double* y, double tin, double tout, const double* atol, const double* rtol,
                           const int* icntrl, const double* rcntrl, int* istatus, double* rstatus, int* ierr, int nthreads) {
  if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
  for (long c = 0; c < ncells; ++c) {
    switch (model) {
      case MODEL_ROBERTS: { Roberts m; ierr[c] = sdirk_integrate<Roberts>(m, tin, tout, y + c * Roberts::N, rtol, atol, icntrl, rcntrl, istatus + c * 20, rstatus + c * 20); break; }
      case MODEL_SMALL_STRATO: { SmallStrato m; ierr[c] = sdirk_integrate<SmallStrato>(m, tin, tout, y + c * SmallStrato::N, rtol, atol, icntrl, rcntrl, istatus + c * 20, rstatus + c * 20); break; }
      case MODEL_CBM4: { Cbm4 m; ierr[c] = sdirk_integrate<Cbm4>(m, tin, tout, y + c * Cbm4::N, rtol, atol, icntrl, rcntrl, istatus + c * 20, rstatus + c * 20); break; }
      default: ierr[c] = -100;
    }
  }
  return 0;
}

// FWD/RK INTEGRATE looped over cells.  raw != 0 bypasses INTEGRATE's presets (calls RungeKutta with the arrays as given).
// stats (optional, 2 longs per cell): out[0] = bit mask of DGETF2 row interchanges jp != j that occurred anywhere is not
// tracked here; see oracle_rk_pivot_census for the pivot statistics used by the code generator.
int oracle_rk_fwd_batch(int model, long ncells, double* y, double tin, double tout, const double* atol, const double* rtol,
                        const int* icntrl, const double* rcntrl, int* istatus, double* rstatus, int* ierr, int raw, int nthreads) {
  if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
  for (long c = 0; c < ncells; ++c) {
    switch (model) {
      case MODEL_ROBERTS: { Roberts m; ierr[c] = rk_integrate<Roberts>(m, tin, tout, y + c * Roberts::N, rtol, atol, icntrl, rcntrl, istatus + c * 20, rstatus + c * 20, raw != 0); break; }
      case MODEL_SMALL_STRATO: { SmallStrato m; ierr[c] = rk_integrate<SmallStrato>(m, tin, tout, y + c * SmallStrato::N, rtol, atol, icntrl, rcntrl, istatus + c * 20, rstatus + c * 20, raw != 0); break; }
      case MODEL_CBM4: { Cbm4 m; ierr[c] = rk_integrate<Cbm4>(m, tin, tout, y + c * Cbm4::N, rtol, atol, icntrl, rcntrl, istatus + c * 20, rstatus + c * 20, raw != 0); break; }
      default: ierr[c] = -100;
    }
  }
  return 0;
}

}  // extern "C"

// unit-level probes for the parity tests
extern "C" {
// interchange census (see ref_lapack.hpp): copies the 32x32 count tables [column][pivot row], optionally clearing them
int oracle_pivot_census(long* d, long* z, int reset) {
  PivotCensus& c = pivot_census();
  if (d) std::memcpy(d, c.d, sizeof c.d);
  if (z) std::memcpy(z, c.z, sizeof c.z);
  if (reset) std::memset(&c, 0, sizeof c);
  return 0;
}
int oracle_cbm4_fun_jac(double t, const double* y, double* f, double* jac, double* hess) {
  Cbm4 m;
  m.fun(t, y, f);
  m.jac(t, y, jac);
  m.update(t);
  cbm4_gen::hessian(y, m.fix, m.rconst, hess);
  return 0;
}
// DGETF2 + DGETRS('N' and 'T'): a is column-major n x n (overwritten by L\U), ipiv 1-based
int oracle_dgetf2_solve(int n, double* a, int* ipiv, const double* b, double* x_n, double* x_t) {
  int info = dgetf2(n, a, ipiv);
  for (int i = 0; i < n; ++i) { x_n[i] = b[i]; x_t[i] = b[i]; }
  dgetrs(false, n, a, ipiv, x_n);
  dgetrs(true, n, a, ipiv, x_t);
  return info;
}
}
