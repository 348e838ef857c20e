// SDIRK forward sweep, ONE SYSTEM PER THREAD (FWD/SDIRK INTEGRATE).
//
// Replaces SDIRK_Integrator (FWD/SDIRK/SDIRK_f90_Integrator.F90:430-657) with SDIRK_PrepareMatrix :736-785, SDIRK_Solve
// :789-810, SDIRK_ErrorScale/ErrorNorm :661-691 and the LS_Solver DGETRF/DGETRS path, for models that provide the
// per-thread policy of ros_cell_fwd.cuh (fun, build_E, lu_factor, lu_solve).  Arithmetic and operation order are the
// reference's (no FMA contraction), so states and all ISTATUS counters are bit-identical to the oracle.
//
// Same mapping as k_ros_fwd_cell: the stage increments Z_1..Z_S, G, TMP, RHS, SCAL, Y and the L\U factors live in the
// thread's shared-memory slots (element stride = lanes per block), every lane pulls its next system from an atomic
// queue.  The time loop of the reference ("CYCLE Tloop" after a failed Newton iteration, LU reuse across accepted steps)
// is kept as is: one pass of the loop body = one attempt; the simplified Newton iterations of a stage are a
// data-dependent inner loop (lanes that have converged wait for the slowest of their warp).
// E = 1/(h*gamma) I - J is the same matrix family as in the Rosenbrock sweep, so the CBM-IV policy factorises it with
// the same static-pattern DGETF2; a system that needs any other pivot sequence ends with IERR = SDIRK_IERR_IRREGULAR, keeps
// its initial state and is integrated again by the host with the general-pivoting policy Cbm4DenseCell (cell_ids lists
// those systems).  The dense small-model policies handle every case themselves.
#pragma once
#include "fatode_common.cuh"
#include "ros_cell_fwd.cuh"
#include "portable_math.h"
#include "gen/sdirk_tableau_gen.h"
#include "sdirk_params.h"

namespace fatode {

constexpr int SDIRK_IERR_IRREGULAR = -20;

template <class Model, int LANES>
struct SdirkSmem {
  static constexpr int N_ = Model::N;
  static constexpr int OFF_Y = 0, OFF_G = N_, OFF_TMP = 2 * N_, OFF_RHS = 3 * N_, OFF_SCAL = 4 * N_, OFF_DINV = 5 * N_, OFF_Z = 6 * N_,
                       OFF_YJ = 11 * N_, OFF_LU = 12 * N_, DOUBLES = OFF_LU + Model::NLU;
  static constexpr size_t BYTES = sizeof(double) * DOUBLES * LANES;
};

// TLM = true: the state part of TLM/SDIRK_TLM (SDIRK_IntegratorTLM, TLM/SDIRK_TLM/SDIRK_TLM_f90_Integrator.F90:474-911) in
// its preset mode, where the tangent-linear columns do not influence the step control: the LU is never reused across
// steps (:879), every completed stage counts one more Jacobian (:701-702) and the NTLM * min(saveNiter, NewtonMaxit) solves
// of its fixed-count TLM iterations (:727, :769 — rejected attempts included), and every accepted step pushes
// (T, H, saveNiter of the stages, Y, Z_1..Z_S) to the system's tape slot: record r of queue slot s lives at
// tape[(s * tape_cap + r) * SD_REC].  The columns themselves are propagated afterwards (sdirk_cell_tlm.cuh).
// MODE 0: FWD/SDIRK.  MODE 1 (TLM): as described above.  MODE 2 (ADJ): SDIRK_FwdInt of ADJ/SDIRK_ADJ
// (ADJ/SDIRK_ADJ/SDIRK_ADJ_f90_Integrator.F90:545-781) = the forward family's loop with every accepted step pushed to the tape
// (:735) and its own SDIRK_PrepareMatrix, which re-evaluates the Jacobian after a singular factorisation (:1239).
constexpr int SD_MODE_FWD = 0, SD_MODE_TLM = 1, SD_MODE_ADJ = 2;

template <class Model, int LANES, int MODE = 0>
__global__ void __launch_bounds__(32, 1)
k_sdirk_fwd_cell(SdirkParams sp, typename Model::Const mc, long ncells, const double* __restrict__ y_in, double* __restrict__ y_out,
                 const double* __restrict__ atol_g, const double* __restrict__ rtol_g, int* __restrict__ istatus,
                 double* __restrict__ rstatus, int* __restrict__ ierr_out, unsigned long long* __restrict__ queue,
                 const int* __restrict__ cell_ids, double* __restrict__ tape, int tape_cap, int ntlm) {
  extern __shared__ __align__(16) double cell_smem[];
  constexpr int N = Model::N;
  constexpr int SV = LANES;
  typedef SdirkSmem<Model, LANES> L;
  if (threadIdx.x >= LANES) return;
  double* const base = cell_smem + threadIdx.x;
  double* const Y = base + L::OFF_Y * SV;
  double* const G = base + L::OFF_G * SV;
  double* const TMP = base + L::OFF_TMP * SV;
  double* const RHS = base + L::OFF_RHS * SV;
  double* const SCAL = base + L::OFF_SCAL * SV;
  double* const dinv = base + L::OFF_DINV * SV;
  double* const Z = base + L::OFF_Z * SV;       // Z[(s * N + i) * SV]
  double* const YJ = base + L::OFF_YJ * SV;     // state at which the stored Jacobian was evaluated (SkipJac keeps it)
  double* const lu = base + L::OFF_LU * SV;
  const int S = sp.S;
  const double Roundoff = sp.roundoff;
  const double Tfinal = sp.tend;
  const double Tdirection = (sp.tend - sp.tstart) >= 0.0 ? 1.0 : -1.0;
  double PH[Model::NPH];
  unsigned swaps = 0;

  auto error_scale = [&]() {   // SDIRK_ErrorScale :661-676
#pragma unroll 4
    for (int i = 0; i < N; ++i) {
      const int ti = sp.vector_tol ? i : 0;
      SCAL[i * SV] = __ddiv_rn(1.0, __dadd_rn(atol_g[ti], __dmul_rn(rtol_g[ti], fabs(Y[i * SV]))));
    }
  };
  auto error_norm = [&](const double* v) -> double {   // SDIRK_ErrorNorm :680-691
    double Err = 0.0;
#pragma unroll 4
    for (int i = 0; i < N; ++i) {
      const double q = __dmul_rn(v[i * SV], SCAL[i * SV]);
      Err = __dadd_rn(Err, __dmul_rn(q, q));
    }
    return fmax(sqrt(__ddiv_rn(Err, (double)N)), 1.0e-10);
  };
  auto solve = [&](double H, double* v) {   // SDIRK_Solve :789-810
    const double HGammaInv = __ddiv_rn(1.0, __dmul_rn(H, sp.Gamma));
#pragma unroll 4
    for (int i = 0; i < N; ++i) v[i * SV] = __dmul_rn(HGammaInv, v[i * SV]);
    Model::template lu_solve<SV>(lu, swaps, v);
  };

  for (;;) {
    const long slot = (long)atomicAdd(queue, 1ull);
    if (slot >= ncells) break;
    const long cell = cell_ids ? (long)cell_ids[slot] : slot;   // second pass: the systems the first pass handed back
#pragma unroll 4
    for (int i = 0; i < N; ++i) Y[i * SV] = y_in[cell * N + i];
    int nfun = 0, njac = 0, nstp = 0, nacc = 0, nrej = 0, ndec = 0, nsol = 0, nsng = 0;
    double texit = 0.0, hexit = 0.0, hnew_out = 0.0;
    double T = sp.tstart;
    double H = fmax(fabs(sp.hmin), fabs(sp.hstart));
    if (fabs(H) <= __dmul_rn(10.0, Roundoff)) H = 1.0e-6;
    H = fmin(fabs(H), sp.hmax);
    H = copysign(H, Tdirection);
    bool SkipLU = false, SkipJac = false, Reject = false, FirstStep = true;
    double Theta = 0.0, NewtonRate = 0.0, NewtonIncrement = 0.0, NewtonIncrementOld = 0.0;
    int ierr = 1;
    double TJ = T;
    error_scale();

    while (__dsub_rn(__dmul_rn(__dsub_rn(Tfinal, T), Tdirection), Roundoff) > 0.0) {   // Tloop
      if (!SkipLU) {   // SDIRK_PrepareMatrix :736-785
        int ConsecutiveSng = 0, code;
        for (;;) {
          const double HGammaInv = __ddiv_rn(1.0, __dmul_rn(H, sp.Gamma));
          if (!SkipJac) {                   // LSS_Jac(T,Y): remember where; the values are rebuilt inside build_E
            njac++;
            TJ = T;
#pragma unroll 4
            for (int i = 0; i < N; ++i) YJ[i * SV] = Y[i * SV];
          }
          // SkipJac keeps the STORED Jacobian, which after a reused LU may be older than the current (T, Y)
          Model::photo(TJ, mc, PH);
          Model::template build_E<SV>(YJ, PH, mc, HGammaInv, lu);
          code = Model::template lu_factor<SV>(lu, dinv, swaps);
          ndec++;
          if (code != LU_SINGULAR) break;
          nsng++;
          if (++ConsecutiveSng >= 6) break;
          H = __dmul_rn(0.5, H);
          SkipJac = (MODE != SD_MODE_ADJ); SkipLU = false; Reject = true;
        }
        if (code == LU_IRREGULAR) { ierr = SDIRK_IERR_IRREGULAR; break; }
        if (code == LU_SINGULAR) { ierr = -8; break; }
      }
      if (nstp > sp.max_no_steps) { ierr = -6; break; }
      if (__dadd_rn(T, __dmul_rn(0.1, H)) == T || fabs(H) <= Roundoff) { ierr = -7; break; }
      bool newton_failed = false;
      double Fac = 0.5;
      unsigned save_pack = 0u;
      for (int istage = 0; istage < S; ++istage) {
        double* Zi = Z + istage * N * SV;
#pragma unroll 4
        for (int i = 0; i < N; ++i) { Zi[i * SV] = 0.0; G[i * SV] = 0.0; }
        for (int j = 0; j < istage; ++j) {
          const double th = sp.Theta[istage][j], al = sp.Alpha[istage][j];
          const double* Zj = Z + j * N * SV;
#pragma unroll 4
          for (int i = 0; i < N; ++i) {
            const double zj = Zj[i * SV];
            G[i * SV] = __dadd_rn(G[i * SV], __dmul_rn(th, zj));
            if (sp.start_newton) Zi[i * SV] = __dadd_rn(Zi[i * SV], __dmul_rn(al, zj));
          }
        }
        bool NewtonDone = false;
        Fac = 0.5;
        Model::photo(__dadd_rn(T, __dmul_rn(sp.C[istage], H)), mc, PH);
        const double HG = __dmul_rn(H, sp.Gamma);
        for (int NewtonIter = 1; NewtonIter <= sp.newton_maxit; ++NewtonIter) {
#pragma unroll 4
          for (int i = 0; i < N; ++i) TMP[i * SV] = __dadd_rn(Y[i * SV], Zi[i * SV]);
          Model::template fun<SV>(TMP, PH, mc, RHS);
          nfun++;
#pragma unroll 4
          for (int i = 0; i < N; ++i)
            RHS[i * SV] = __dadd_rn(__dadd_rn(__dmul_rn(HG, RHS[i * SV]), -Zi[i * SV]), G[i * SV]);
          solve(H, RHS);
          nsol++;
          NewtonIncrement = error_norm(RHS);
          if (NewtonIter == 1) {
            Theta = fabs(sp.thetamin);
            NewtonRate = 2.0;
          } else {
            Theta = __ddiv_rn(NewtonIncrement, NewtonIncrementOld);
            if (Theta < 0.99) {
              NewtonRate = __ddiv_rn(Theta, __dsub_rn(1.0, Theta));
              const double pred = __ddiv_rn(__dmul_rn(NewtonIncrement, sdirk_powi(Theta, sp.newton_maxit - NewtonIter)), __dsub_rn(1.0, Theta));
              if (pred >= sp.newtontol) {   // predicted error too large
                const double Qnewton = fmin(10.0, __ddiv_rn(pred, sp.newtontol));
                Fac = __dmul_rn(0.8, fpm::pow_neg_inv(Qnewton, (double)(1 + sp.newton_maxit - NewtonIter)));
                break;
              }
            } else {
              break;                        // Theta too large
            }
          }
          NewtonIncrementOld = NewtonIncrement;
#pragma unroll 4
          for (int i = 0; i < N; ++i) Zi[i * SV] = __dadd_rn(Zi[i * SV], RHS[i * SV]);
          NewtonDone = (__dmul_rn(NewtonRate, NewtonIncrement) <= sp.newtontol);
          if (NewtonDone) {
            if constexpr (MODE == SD_MODE_TLM) {   // saveNiter = NewtonIter + 1 (:603); the TLM loop runs min(saveNiter, NewtonMaxit) iterations
              const int it = min(NewtonIter + 1, sp.newton_maxit);
              save_pack |= (unsigned)it << (4 * istage);
              njac++;
              nsol += ntlm * it;
            }
            break;
          }
        }
        if (!NewtonDone) { newton_failed = true; break; }
      }
      if (newton_failed) {   // CYCLE Tloop (:561-566)
        H = __dmul_rn(Fac, H);
        Reject = true; SkipJac = true; SkipLU = false;
        continue;
      }
      // ---- error estimation (:576-589)
      nstp++;
#pragma unroll 4
      for (int i = 0; i < N; ++i) TMP[i * SV] = 0.0;
      for (int s = 0; s < S; ++s) {
        const double e = sp.E[s];
        if (e != 0.0) {
          const double* Zs = Z + s * N * SV;
#pragma unroll 4
          for (int i = 0; i < N; ++i) TMP[i * SV] = __dadd_rn(TMP[i * SV], __dmul_rn(e, Zs[i * SV]));
        }
      }
      solve(H, TMP);
      nsol++;
      const double Err = error_norm(TMP);
      double F2 = __dmul_rn(sp.facsafe, fpm::pow_neg_inv(Err, sp.ELO));
      F2 = fmax(sp.facmin, fmin(sp.facmax, F2));
      double Hnew = __dmul_rn(H, F2);
      if (Err < 1.0) {   // accept (:594-631)
        if constexpr (MODE != SD_MODE_FWD) if (tape) {
          if (nacc >= tape_cap) { ierr = IERR_TAPE_OVERFLOW; break; }
          double* rec = tape + ((size_t)slot * tape_cap + nacc) * SD_REC;
          st_global_256(rec, T, H, (double)save_pack, (double)S);
#pragma unroll 2
          for (int i = 0; i < N; i += 4)
            st_global_256(rec + 4 + i, Y[i * SV], Y[(i + 1) * SV], Y[(i + 2) * SV], Y[(i + 3) * SV]);
          for (int s2 = 0; s2 < S; ++s2) {
            const double* Zs = Z + s2 * N * SV;
#pragma unroll 2
            for (int i = 0; i < N; i += 4)
              st_global_256(rec + 4 + N + s2 * N + i, Zs[i * SV], Zs[(i + 1) * SV], Zs[(i + 2) * SV], Zs[(i + 3) * SV]);
          }
        }
        FirstStep = false;
        nacc++;
        T = __dadd_rn(T, H);
        for (int s = 0; s < S; ++s) {
          const double d = sp.D[s];
          if (d != 0.0) {
            const double* Zs = Z + s * N * SV;
#pragma unroll 4
            for (int i = 0; i < N; ++i) Y[i * SV] = __dadd_rn(Y[i * SV], __dmul_rn(d, Zs[i * SV]));
          }
        }
        error_scale();
        Hnew = __dmul_rn(Tdirection, fmin(fabs(Hnew), sp.hmax));
        texit = T; hexit = H; hnew_out = Hnew;
        if (Reject) Hnew = __dmul_rn(Tdirection, fmin(fabs(Hnew), fabs(H)));
        Reject = false;
        if (__dmul_rn(__dsub_rn(__dadd_rn(T, __ddiv_rn(Hnew, sp.qmin)), Tfinal), Tdirection) > 0.0) {
          H = __dsub_rn(Tfinal, T);
        } else {
          const double Hratio = __ddiv_rn(Hnew, H);
          SkipLU = (MODE != SD_MODE_TLM) && (Theta <= sp.thetamin) && (Hratio >= sp.qmin) && (Hratio <= sp.qmax);   // TLM: never (:879)
          if (!SkipLU) H = Hnew;
        }
        SkipJac = false;
      } else {           // reject (:633-646)
        if (FirstStep || Reject) H = __dmul_rn(sp.facrej, H); else H = Hnew;
        Reject = true; SkipJac = true; SkipLU = false;
        if (nacc >= 1) nrej++;
      }
    }
    // ---- results
    ierr_out[cell] = ierr;
    if (ierr != SDIRK_IERR_IRREGULAR && ierr != IERR_TAPE_OVERFLOW) {   // a handed-back system keeps its initial state
#pragma unroll 4
      for (int i = 0; i < N; ++i) y_out[cell * N + i] = Y[i * SV];
    }
    int* is_ = istatus + cell * 20;
    for (int q = 0; q < 20; ++q) is_[q] = 0;
    is_[Nfun] = nfun; is_[Njac] = njac; is_[Nstp] = nstp; is_[Nacc] = nacc; is_[Nrej] = nrej;
    is_[Ndec] = ndec; is_[Nsol] = nsol; is_[Nsng] = nsng;
    double* rs = rstatus + cell * 20;
    for (int q = 0; q < 20; ++q) rs[q] = 0.0;
    rs[Ntexit] = texit; rs[Nhexit] = hexit; rs[Nhnew] = hnew_out;
  }
}

// host: option handling of SDIRK (:236-417).  Returns the reference's Ierr before integrating (0 = go on, < 0 = the
// last option error: the reference keeps parsing after an error and returns the last code)
// tlm != nullptr: INTEGRATE_TLM of TLM/SDIRK_TLM (same presets; an unknown ICNTRL(3) selects Sdirk2a there, :262-275);
// tlm[0..2] receive ICNTRL(7) (DirectTLM), ICNTRL(9) (TLMNewtonEst), ICNTRL(12) (TLMtruncErr).
inline int sdirk_parse_options(int N, const int* icntrl_u, const double* rcntrl_u, double tin, double tout, const double* atol,
                               const double* rtol, SdirkParams& sp, int* tlm = nullptr) {
  int ic[20];
  double rc[20];
  std::memset(ic, 0, sizeof ic);
  std::memset(rc, 0, sizeof rc);
  for (int i = 0; i < 20; ++i) {   // INTEGRATE :57-64: user values override only when > 0
    if (icntrl_u && icntrl_u[i] > 0) ic[i] = icntrl_u[i];
    if (rcntrl_u && rcntrl_u[i] > 0) rc[i] = rcntrl_u[i];
  }
  std::memset(&sp, 0, sizeof sp);
  int ierr = 0;
  sp.vector_tol = (ic[1] == 0);
  int method;
  switch (ic[2]) {
    case 1: method = 1; break;
    case 2: method = 2; break;
    case 3: method = 3; break;
    case 5: method = 5; break;
    default: method = 4;   // 0, 4 and anything else: Sdirk4a (:262-275)
  }
  if (tlm) {
    if (ic[2] < 0 || ic[2] > 5) method = 1;
    tlm[0] = ic[6]; tlm[1] = ic[8]; tlm[2] = ic[11];
  }
  const SdirkTableauData& t = SDIRK_TABLEAU[method - 1];
  sp.S = t.S; sp.Gamma = t.Gamma; sp.ELO = t.ELO;
  std::memcpy(sp.C, t.C, sizeof sp.C);
  std::memcpy(sp.D, t.D, sizeof sp.D);
  std::memcpy(sp.E, t.E, sizeof sp.E);
  std::memcpy(sp.Theta, t.Theta, sizeof sp.Theta);
  std::memcpy(sp.Alpha, t.Alpha, sizeof sp.Alpha);
  if (ic[3] == 0) sp.max_no_steps = 200000; else if (ic[3] > 0) sp.max_no_steps = ic[3]; else ierr = -1;
  if (ic[4] == 0) sp.newton_maxit = 8; else { sp.newton_maxit = ic[4]; if (sp.newton_maxit <= 0) ierr = -2; }
  sp.start_newton = (ic[5] == 0);
  sp.roundoff = DBL_EPSILON * 0.5;
  const double span = std::fabs(tout - tin);
  if (rc[0] == 0.0) sp.hmin = 0.0; else if (rc[0] > 0.0) sp.hmin = rc[0]; else ierr = -3;
  if (rc[1] == 0.0) sp.hmax = span; else if (rc[1] > 0.0) sp.hmax = std::fmin(std::fabs(rc[1]), span); else ierr = -3;
  if (rc[2] == 0.0) sp.hstart = std::fmax(sp.hmin, sp.roundoff); else if (rc[2] > 0.0) sp.hstart = std::fmin(std::fabs(rc[2]), span); else ierr = -3;
  // REAL(4) literals 0.2, 10.0, 0.1, 0.9 (:341-367)
  if (rc[3] == 0.0) sp.facmin = (double)0.2f; else if (rc[3] > 0.0) sp.facmin = rc[3]; else ierr = -4;
  if (rc[4] == 0.0) sp.facmax = (double)10.0f; else if (rc[4] > 0.0) sp.facmax = rc[4]; else ierr = -4;
  if (rc[5] == 0.0) sp.facrej = (double)0.1f; else if (rc[5] > 0.0) sp.facrej = rc[5]; else ierr = -4;
  if (rc[6] == 0.0) sp.facsafe = (double)0.9f; else if (rc[6] > 0.0) sp.facsafe = rc[6]; else ierr = -4;
  sp.thetamin = (rc[7] == 0.0) ? 1.0e-3 : rc[7];
  sp.newtontol = (rc[8] == 0.0) ? 3.0e-2 : rc[8];
  sp.qmin = (rc[9] == 0.0) ? 1.0 : rc[9];
  sp.qmax = (rc[10] == 0.0) ? 1.2 : rc[10];
  const int uplim = sp.vector_tol ? N : 1;
  for (int i = 0; i < uplim; ++i)
    if (atol[i] <= 0.0 || rtol[i] <= 10.0 * sp.roundoff) ierr = -5;
  sp.tstart = tin;
  sp.tend = tout;
  return ierr;
}

}  // namespace fatode
