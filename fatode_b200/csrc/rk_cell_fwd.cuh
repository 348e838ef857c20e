// Fully implicit 3-stage Runge-Kutta forward sweep (Radau-2A, Radau-1A, Lobatto-3C, Gauss, Lobatto-3A),
// ONE SYSTEM PER THREAD (FWD/RK INTEGRATE).
//
// Replaces RK_Integrator (FWD/RK/RK_f90_Integrator.F90:461-780) with RK_ErrorScale :837-857, RK_Interpolate :877-913,
// RK_PrepareRHS :917-955, RK_Decomp :959-988, RK_Solve :992-1023, RK_ErrorNorm :1087-1101 and the LS_Solver
// DGETRF/DGETRS + ZGETRF/ZGETRS path (FWD/RK/LS_Solver.F90:10-52).  Arithmetic and operation order are the reference's
// (no FMA contraction; the complex shift alpha/h + i beta/h is rounded to REAL(4) like the reference's CMPLX() without
// KIND), so states and all ISTATUS counters are bit-identical to the oracle.
//
// Same mapping as k_sdirk_fwd_cell: all vectors (Y, SCAL, F0, Z1..Z4, the Newton right-hand sides, G, the interpolant
// CONT) and the three factor planes (real L\U of gamma/h I - J; real and imaginary plane of the complex L\U of
// (alpha + i beta)/h I - J) live in the thread's shared-memory slots (element stride = lanes per block); every lane pulls
// its next system from an atomic queue.  The reference's time loop is kept as is, including "CYCLE Tloop" after a singular
// matrix or a failed Newton iteration, LU reuse across accepted steps (SkipLU) and Jacobian reuse after rejections
// (SkipJac: the stored Jacobian may be older than the current state, so the state it was evaluated at is kept).
// INTEGRATE presets ICNTRL(10) = 1, so the error estimate is always the embedded SDIRK stage (:604-683); the classic
// estimator (RK_ErrorEstimate) cannot be selected through this interface and is not built.
// The CBM-IV policy factorises both matrices with the generated static-pattern DGETF2 / ZGETF2; a system that needs any
// other pivot sequence ends with IERR = RK_IERR_IRREGULAR, keeps its initial state and is integrated again by the host
// with the general-pivoting policy Cbm4DenseCell (cell_ids lists those systems).  The dense small-model policies handle
// every case themselves.
#pragma once
#include "fatode_common.cuh"
#include "ros_cell_fwd.cuh"
#include "sdirk_cell_fwd.cuh"
#include "portable_math.h"
#include "gen/rk_tableau_gen.h"

namespace fatode {

constexpr int RK_IERR_IRREGULAR = -20;
enum { RK_R2A = 1, RK_R1A = 2, RK_L3C = 3, RK_GAU = 4, RK_L3A = 5 };

struct RkParams {
  int method;
  double T[3][3], TinvAinv[3][3], Tinv[3][3], AinvT[3][3], A[4][4], B[4], C[4], D[4], Bgam[5], Theta[4], F[5], Gamma, Alpha, Beta, ELO;
  int vector_tol, max_no_steps, newton_maxit, start_newton, gustafsson;
  int tlm;   // RK_IntegratorTLM flavour of the ADJ forward sweep: RK_PrepareRHS counts its FUN calls, the LU is never reused (TLM/RK_TLM :934-935, 1120-1134)
  double roundoff, hmin, hmax, hstart, facmin, facmax, facrej, facsafe, thetamin, newtontol, qmin, qmax;
  double tstart, tend;
};

template <class Model, int LANES>
struct RkSmem {
  static constexpr int N_ = Model::N;
  static constexpr int OFF_Y = 0, OFF_SCAL = N_, OFF_F0 = 2 * N_, OFF_Z1 = 3 * N_, OFF_Z2 = 4 * N_, OFF_Z3 = 5 * N_, OFF_R1 = 6 * N_,
                       OFF_R2 = 7 * N_, OFF_R3 = 8 * N_, OFF_Z4 = 9 * N_, OFF_G = 10 * N_, OFF_TMP = 11 * N_, OFF_F = 12 * N_,
                       OFF_CONT = 13 * N_, OFF_YJ = 16 * N_, OFF_DINV = 17 * N_, OFF_LU = 18 * N_, OFF_ZR = OFF_LU + Model::NLU,
                       OFF_ZI = OFF_ZR + Model::NLU, DOUBLES = OFF_ZI + Model::NLU;
  static constexpr size_t BYTES = sizeof(double) * DOUBLES * LANES;
};

// ADJ = true: the forward sweep of ADJ/RK_ADJ (RK_FwdIntegrator, ADJ/RK_ADJ/RK_ADJ_f90_Integrator.F90:866-1226): no FUN(T,Y)
// at the top of the time loop (the SDIRK error stage evaluates and counts its own), and every accepted step pushes
// (T, H, NewIt, Y, Z1, Z2, Z3) — rk_DPush :775-794 — to the system's tape slot: record r of queue slot s lives at
// tape[(s * tape_cap + r) * RK_REC].  A system with more than tape_cap accepted steps ends with IERR_TAPE_OVERFLOW and is
// run again by the host with a larger slot.
constexpr int RK_REC = 132;   // doubles per tape record: T, H, NewIt, pad, Y(32), Z1(32), Z2(32), Z3(32)

template <class Model, int LANES, bool ADJ = false>
__global__ void __launch_bounds__(32, 1)
k_rk_fwd_cell(RkParams rp, typename Model::Const mc, long ncells, const double* __restrict__ y_in, double* __restrict__ y_out,
              const double* __restrict__ atol_g, const double* __restrict__ rtol_g, int* __restrict__ istatus,
              double* __restrict__ rstatus, int* __restrict__ ierr_out, unsigned long long* __restrict__ queue,
              const int* __restrict__ cell_ids, double* __restrict__ tape, int tape_cap) {
  extern __shared__ __align__(16) double cell_smem[];
  constexpr int N = Model::N;
  constexpr int SV = LANES;
  typedef RkSmem<Model, LANES> L;
  if (threadIdx.x >= LANES) return;
  double* const base = cell_smem + threadIdx.x;
  double* const Y = base + L::OFF_Y * SV;
  double* const SCAL = base + L::OFF_SCAL * SV;
  double* const F0 = base + L::OFF_F0 * SV;
  double* const Z1 = base + L::OFF_Z1 * SV;
  double* const Z2 = base + L::OFF_Z2 * SV;
  double* const Z3 = base + L::OFF_Z3 * SV;
  double* const R1 = base + L::OFF_R1 * SV;     // DZ1; DZ4 in the SDIRK stage
  double* const R2 = base + L::OFF_R2 * SV;     // DZ2
  double* const R3 = base + L::OFF_R3 * SV;     // DZ3
  double* const Z4 = base + L::OFF_Z4 * SV;
  double* const G = base + L::OFF_G * SV;
  double* const TMP = base + L::OFF_TMP * SV;
  double* const F = base + L::OFF_F * SV;
  double* const CONT = base + L::OFF_CONT * SV;   // CONT[(k * N + i) * SV], k = 0..2
  double* const YJ = base + L::OFF_YJ * SV;       // state at which the stored Jacobian was evaluated
  double* const dinv = base + L::OFF_DINV * SV;
  double* const lu = base + L::OFF_LU * SV;
  double* const zr = base + L::OFF_ZR * SV;
  double* const zi = base + L::OFF_ZI * SV;
  const double Roundoff = rp.roundoff;
  const double Tend = rp.tend;
  const double Tdirection = (rp.tend - rp.tstart) >= 0.0 ? 1.0 : -1.0;
  const int NewtonMaxit = rp.newton_maxit;
  double PH[Model::NPH];
  unsigned swaps = 0, zswaps = 0;

  auto error_scale = [&]() {   // RK_ErrorScale :837-857
#pragma unroll 4
    for (int i = 0; i < N; ++i) {
      const int ti = rp.vector_tol ? i : 0;
      SCAL[i * SV] = __ddiv_rn(1.0, __dadd_rn(atol_g[ti], __dmul_rn(rtol_g[ti], fabs(Y[i * SV]))));
    }
  };
  auto error_norm = [&](const double* v) -> double {   // RK_ErrorNorm :1087-1101
    double s = 0.0;
#pragma unroll 4
    for (int i = 0; i < N; ++i) {
      const double q = __dmul_rn(v[i * SV], SCAL[i * SV]);
      s = __dadd_rn(s, __dmul_rn(q, q));
    }
    return fmax(sqrt(__ddiv_rn(s, (double)N)), 1.0e-10);
  };

  for (;;) {
    const long slot = (long)atomicAdd(queue, 1ull);
    if (slot >= ncells) break;
    const long cell = cell_ids ? (long)cell_ids[slot] : slot;   // second pass: the systems the first pass handed back
#pragma unroll 4
    for (int i = 0; i < N; ++i) Y[i * SV] = y_in[cell * N + i];
    int nfun = 0, njac = 0, nstp = 0, nacc = 0, nrej = 0, ndec = 0, nsol = 0, nsng = 0;
    double texit = 0.0, hexit = 0.0, hnew_out = 0.0;
    double T = rp.tstart;
    double H = fmin(fmax(fabs(rp.hmin), fabs(rp.hstart)), rp.hmax);
    if (fabs(H) <= __dmul_rn(10.0, Roundoff)) H = 1.0e-6;
    H = copysign(H, Tdirection);
    double Hold = H, Hacc = 0.0, ErrOld = 0.0;
    bool Reject = false, FirstStep = true, SkipJac = false, SkipLU = false;
    if (__dmul_rn(__dsub_rn(__dadd_rn(T, __dmul_rn(H, 1.0001)), Tend), Tdirection) >= 0.0) H = __dsub_rn(Tend, T);
    int Nconsecutive = 0;
    double Theta = 0.0, NewtonRate = 0.0, NewtonIncrement = 0.0, NewtonIncrementOld = 0.0;
    int ierr = 1;
    double TJ = T;
    error_scale();

    int NewIt = 0;
    while (__dsub_rn(__dmul_rn(__dsub_rn(Tend, T), Tdirection), Roundoff) > 0.0) {   // Tloop
      if constexpr (!ADJ) {
        Model::photo(T, mc, PH);
        Model::template fun<SV>(Y, PH, mc, F0);
        nfun++;
      }
      if (!SkipLU) {
        if (!SkipJac) {                      // LSS_Jac(T,Y): remember where; the values are rebuilt inside build_E
          njac++;
          TJ = T;
#pragma unroll 4
          for (int i = 0; i < N; ++i) YJ[i * SV] = Y[i * SV];
        }
        // RK_Decomp :959-988
        Model::photo(TJ, mc, PH);
        Model::template build_E<SV>(YJ, PH, mc, __ddiv_rn(rp.Gamma, H), lu);
        int code = Model::template lu_factor<SV>(lu, dinv, swaps);
        if (code == LU_OK) {
          const double a4 = (double)(float)__ddiv_rn(rp.Alpha, H), b4 = (double)(float)__ddiv_rn(rp.Beta, H);   // CMPLX(): REAL(4)
          Model::template build_EZ<SV>(YJ, PH, mc, a4, b4, zr, zi);
          code = Model::template zlu_factor<SV>(zr, zi, zswaps);
        }
        ndec++;
        if (code == LU_IRREGULAR) { ierr = RK_IERR_IRREGULAR; break; }
        if (code == LU_SINGULAR) {
          nsng++; Nconsecutive++;
          if (Nconsecutive >= 5) { ierr = -12; break; }
          H = __dmul_rn(H, 0.5); Reject = true; SkipJac = true; SkipLU = false;
          continue;
        }
        Nconsecutive = 0;
      }
      nstp++;
      if (nstp > rp.max_no_steps) { ierr = -9; break; }
      if (__dmul_rn(0.1, fabs(H)) <= __dmul_rn(fabs(T), Roundoff)) { ierr = -10; break; }

      // ---- starting values (:540-547)
      if (FirstStep || !rp.start_newton) {
#pragma unroll 4
        for (int i = 0; i < N; ++i) { Z1[i * SV] = 0.0; Z2[i * SV] = 0.0; Z3[i * SV] = 0.0; }
      } else {   // RK_Interpolate('eval')
        const double r = __ddiv_rn(H, Hold);
        const double x1 = __dadd_rn(1.0, __dmul_rn(rp.C[1], r)), x2 = __dadd_rn(1.0, __dmul_rn(rp.C[2], r)),
                     x3 = __dadd_rn(1.0, __dmul_rn(rp.C[3], r));
#pragma unroll 4
        for (int i = 0; i < N; ++i) {
          const double c1 = CONT[i * SV], c2 = CONT[(N + i) * SV], c3 = CONT[(2 * N + i) * SV];
          Z1[i * SV] = __dadd_rn(c1, __dmul_rn(x1, __dadd_rn(c2, __dmul_rn(x1, c3))));
          Z2[i * SV] = __dadd_rn(c1, __dmul_rn(x2, __dadd_rn(c2, __dmul_rn(x2, c3))));
          Z3[i * SV] = __dadd_rn(c1, __dmul_rn(x3, __dadd_rn(c2, __dmul_rn(x3, c3))));
        }
      }

      // ---- simplified Newton iterations on the three coupled stages (:553-596)
      bool NewtonDone = false;
      double Fac = 0.5;
      int NewtonIter;
      for (NewtonIter = 1; NewtonIter <= NewtonMaxit; ++NewtonIter) {
        // RK_PrepareRHS :917-955 (its FUN calls are not counted, except by the RK_TLM copy of the routine)
        if (rp.tlm) nfun += 3;
#pragma unroll 4
        for (int i = 0; i < N; ++i) { R1[i * SV] = Z1[i * SV]; R2[i * SV] = Z2[i * SV]; R3[i * SV] = Z3[i * SV]; }
        if (rp.method == RK_L3A) {
          const double a1 = __dmul_rn(-H, rp.A[1][0]), a2 = __dmul_rn(-H, rp.A[2][0]), a3 = __dmul_rn(-H, rp.A[3][0]);
#pragma unroll 4
          for (int i = 0; i < N; ++i) {
            const double f = F0[i * SV];
            R1[i * SV] = __dadd_rn(R1[i * SV], __dmul_rn(a1, f));
            R2[i * SV] = __dadd_rn(R2[i * SV], __dmul_rn(a2, f));
            R3[i * SV] = __dadd_rn(R3[i * SV], __dmul_rn(a3, f));
          }
        }
        for (int s = 1; s <= 3; ++s) {
          const double* Zs = (s == 1) ? Z1 : (s == 2) ? Z2 : Z3;
#pragma unroll 4
          for (int i = 0; i < N; ++i) TMP[i * SV] = __dadd_rn(Y[i * SV], Zs[i * SV]);
          Model::photo(__dadd_rn(T, __dmul_rn(rp.C[s], H)), mc, PH);
          Model::template fun<SV>(TMP, PH, mc, F);
          const double a1 = __dmul_rn(-H, rp.A[1][s]), a2 = __dmul_rn(-H, rp.A[2][s]), a3 = __dmul_rn(-H, rp.A[3][s]);
#pragma unroll 4
          for (int i = 0; i < N; ++i) {
            const double f = F[i * SV];
            R1[i * SV] = __dadd_rn(R1[i * SV], __dmul_rn(a1, f));
            R2[i * SV] = __dadd_rn(R2[i * SV], __dmul_rn(a2, f));
            R3[i * SV] = __dadd_rn(R3[i * SV], __dmul_rn(a3, f));
          }
        }
        // RK_Solve :992-1023
#pragma unroll 4
        for (int i = 0; i < N; ++i) {
          const double x1 = __ddiv_rn(R1[i * SV], H), x2 = __ddiv_rn(R2[i * SV], H), x3 = __ddiv_rn(R3[i * SV], H);
          R1[i * SV] = __dadd_rn(__dadd_rn(__dmul_rn(rp.TinvAinv[0][0], x1), __dmul_rn(rp.TinvAinv[0][1], x2)), __dmul_rn(rp.TinvAinv[0][2], x3));
          R2[i * SV] = __dadd_rn(__dadd_rn(__dmul_rn(rp.TinvAinv[1][0], x1), __dmul_rn(rp.TinvAinv[1][1], x2)), __dmul_rn(rp.TinvAinv[1][2], x3));
          R3[i * SV] = __dadd_rn(__dadd_rn(__dmul_rn(rp.TinvAinv[2][0], x1), __dmul_rn(rp.TinvAinv[2][1], x2)), __dmul_rn(rp.TinvAinv[2][2], x3));
        }
        Model::template lu_solve<SV>(lu, swaps, R1);
        Model::template zlu_solve<SV>(zr, zi, zswaps, R2, R3);
#pragma unroll 4
        for (int i = 0; i < N; ++i) {
          const double x1 = R1[i * SV], x2 = R2[i * SV], x3 = R3[i * SV];
          R1[i * SV] = __dadd_rn(__dadd_rn(__dmul_rn(rp.T[0][0], x1), __dmul_rn(rp.T[0][1], x2)), __dmul_rn(rp.T[0][2], x3));
          R2[i * SV] = __dadd_rn(__dadd_rn(__dmul_rn(rp.T[1][0], x1), __dmul_rn(rp.T[1][1], x2)), __dmul_rn(rp.T[1][2], x3));
          R3[i * SV] = __dadd_rn(__dadd_rn(__dmul_rn(rp.T[2][0], x1), __dmul_rn(rp.T[2][1], x2)), __dmul_rn(rp.T[2][2], x3));
        }
        nsol += 2;
        {
          const double n1 = error_norm(R1), n2 = error_norm(R2), n3 = error_norm(R3);
          NewtonIncrement = sqrt(__ddiv_rn(__dadd_rn(__dadd_rn(__dmul_rn(n1, n1), __dmul_rn(n2, n2)), __dmul_rn(n3, n3)), 3.0));
        }
        if (NewtonIter == 1) {
          Theta = fabs(rp.thetamin);
          NewtonRate = 2.0;
        } else {
          Theta = __ddiv_rn(NewtonIncrement, NewtonIncrementOld);
          if (Theta < 0.99) NewtonRate = __ddiv_rn(Theta, __dsub_rn(1.0, Theta));
          else break;                       // Theta too large
          if (NewtonIter < NewtonMaxit) {
            const double pred = __ddiv_rn(__dmul_rn(NewtonIncrement, sdirk_powi(Theta, NewtonMaxit - NewtonIter)), __dsub_rn(1.0, Theta));
            if (pred >= rp.newtontol) {     // predicted error too large
              const double Qnewton = fmin(10.0, __ddiv_rn(pred, rp.newtontol));
              Fac = __dmul_rn(0.8, fpm::pow_neg_inv(Qnewton, (double)(1 + NewtonMaxit - NewtonIter)));
              break;
            }
          }
        }
        NewtonIncrementOld = fmax(NewtonIncrement, Roundoff);
#pragma unroll 4
        for (int i = 0; i < N; ++i) {
          Z1[i * SV] = __dadd_rn(Z1[i * SV], __dmul_rn(-1.0, R1[i * SV]));
          Z2[i * SV] = __dadd_rn(Z2[i * SV], __dmul_rn(-1.0, R2[i * SV]));
          Z3[i * SV] = __dadd_rn(Z3[i * SV], __dmul_rn(-1.0, R3[i * SV]));
        }
        NewtonDone = (__dmul_rn(NewtonRate, NewtonIncrement) <= rp.newtontol);
        if (NewtonDone) { NewIt = NewtonIter; break; }
      }
      if (!NewtonDone) {   // CYCLE Tloop (:598-602)
        H = __dmul_rn(Fac, H); Reject = true; SkipJac = true; SkipLU = false;
        continue;
      }

      // ---- embedded SDIRK stage (:604-666)
      {
        if constexpr (ADJ) {   // ADJ/RK_ADJ :1041-1043
          Model::photo(T, mc, PH);
          Model::template fun<SV>(Y, PH, mc, F0);
          nfun++;
        }
        const double bg = __dmul_rn(rp.Bgam[0], H);
#pragma unroll 4
        for (int i = 0; i < N; ++i) {
          Z4[i * SV] = Z3[i * SV];
          double g = 0.0;
          if (rp.method != RK_L3A) g = __dadd_rn(g, __dmul_rn(bg, F0[i * SV]));
          g = __dadd_rn(g, __dmul_rn(rp.Theta[1], Z1[i * SV]));
          g = __dadd_rn(g, __dmul_rn(rp.Theta[2], Z2[i * SV]));
          g = __dadd_rn(g, __dmul_rn(rp.Theta[3], Z3[i * SV]));
          G[i * SV] = g;
        }
        NewtonDone = false;
        Fac = 0.5;
        Model::photo(__dadd_rn(T, H), mc, PH);
        const double s1 = __ddiv_rn(__dmul_rn(-1.0, rp.Gamma), H), s2 = __ddiv_rn(rp.Gamma, H);
        for (NewtonIter = 1; NewtonIter <= NewtonMaxit; ++NewtonIter) {
#pragma unroll 4
          for (int i = 0; i < N; ++i) TMP[i * SV] = __dadd_rn(Y[i * SV], Z4[i * SV]);
          Model::template fun<SV>(TMP, PH, mc, R1);
          nfun++;
#pragma unroll 4
          for (int i = 0; i < N; ++i)
            R1[i * SV] = __dadd_rn(__dadd_rn(R1[i * SV], __dmul_rn(s1, Z4[i * SV])), __dmul_rn(s2, G[i * SV]));
          Model::template lu_solve<SV>(lu, swaps, R1);
          nsol++;
          NewtonIncrement = error_norm(R1);
          if (NewtonIter == 1) {
            NewtonRate = 2.0;               // ThetaSD = ABS(ThetaMin) is never read
          } else {
            const double ThetaSD = __ddiv_rn(NewtonIncrement, NewtonIncrementOld);
            if (ThetaSD < 0.99) {
              NewtonRate = __ddiv_rn(ThetaSD, __dsub_rn(1.0, ThetaSD));
              const double pred = __ddiv_rn(__dmul_rn(NewtonIncrement, sdirk_powi(ThetaSD, NewtonMaxit - NewtonIter)), __dsub_rn(1.0, ThetaSD));
              if (pred >= rp.newtontol) {
                const double Qnewton = fmin(10.0, __ddiv_rn(pred, rp.newtontol));
                Fac = __dmul_rn(0.8, fpm::pow_neg_inv(Qnewton, (double)(1 + NewtonMaxit - NewtonIter)));
                break;
              }
            } else {
              break;
            }
          }
          NewtonIncrementOld = NewtonIncrement;
#pragma unroll 4
          for (int i = 0; i < N; ++i) Z4[i * SV] = __dadd_rn(Z4[i * SV], R1[i * SV]);
          NewtonDone = (__dmul_rn(NewtonRate, NewtonIncrement) <= rp.newtontol);
          if (NewtonDone) break;
        }
        if (!NewtonDone) {
          H = __dmul_rn(Fac, H); Reject = true; SkipJac = true; SkipLU = false;
          continue;
        }
      }

      // ---- error estimate (:671-690)
      if (rp.method == RK_L3A) {
        const double hf = __dmul_rn(H, rp.F[0]);
#pragma unroll 4
        for (int i = 0; i < N; ++i) {
          double d = __dmul_rn(hf, F0[i * SV]);
          if (rp.F[1] != 0.0) d = __dadd_rn(d, __dmul_rn(rp.F[1], Z1[i * SV]));
          if (rp.F[2] != 0.0) d = __dadd_rn(d, __dmul_rn(rp.F[2], Z2[i * SV]));
          if (rp.F[3] != 0.0) d = __dadd_rn(d, __dmul_rn(rp.F[3], Z3[i * SV]));
          R1[i * SV] = d;
          TMP[i * SV] = __dadd_rn(Y[i * SV], Z4[i * SV]);
        }
        Model::template fun<SV>(TMP, PH, mc, R2);   // PH still holds T+H; not counted (:681)
        const double hb = __dmul_rn(H, rp.Bgam[4]);
#pragma unroll 4
        for (int i = 0; i < N; ++i) R1[i * SV] = __dadd_rn(R1[i * SV], __dmul_rn(hb, R2[i * SV]));
      } else {
#pragma unroll 4
        for (int i = 0; i < N; ++i) {
          double d = 0.0;
          if (rp.D[1] != 0.0) d = __dadd_rn(d, __dmul_rn(rp.D[1], Z1[i * SV]));
          if (rp.D[2] != 0.0) d = __dadd_rn(d, __dmul_rn(rp.D[2], Z2[i * SV]));
          if (rp.D[3] != 0.0) d = __dadd_rn(d, __dmul_rn(rp.D[3], Z3[i * SV]));
          R1[i * SV] = __dadd_rn(d, __dmul_rn(-1.0, Z4[i * SV]));
        }
      }
      const double Err = error_norm(R1);

      // ---- new step size (:696-699)
      Fac = __dmul_rn(fpm::pow_neg_inv(Err, rp.ELO),
                      fmin(rp.facsafe, __ddiv_rn((double)(1 + 2 * NewtonMaxit), (double)(NewtonIter + 2 * NewtonMaxit))));
      Fac = fmin(rp.facmax, fmax(rp.facmin, Fac));
      double Hnew = __dmul_rn(Fac, H);

      if (Err < 1.0) {   // accepted (:704-746)
        if constexpr (ADJ) if (tape) {   // rk_DPush (tape == nullptr: ICNTRL(9) = 1, forward sweep only)
          if (nacc >= tape_cap) { ierr = IERR_TAPE_OVERFLOW; break; }
          double* rec = tape + ((size_t)slot * tape_cap + nacc) * RK_REC;
          st_global_256(rec, T, H, (double)NewIt, 0.0);
#pragma unroll 2
          for (int i = 0; i < N; i += 4) {
            st_global_256(rec + 4 + i, Y[i * SV], Y[(i + 1) * SV], Y[(i + 2) * SV], Y[(i + 3) * SV]);
            st_global_256(rec + 4 + N + i, Z1[i * SV], Z1[(i + 1) * SV], Z1[(i + 2) * SV], Z1[(i + 3) * SV]);
            st_global_256(rec + 4 + 2 * N + i, Z2[i * SV], Z2[(i + 1) * SV], Z2[(i + 2) * SV], Z2[(i + 3) * SV]);
            st_global_256(rec + 4 + 3 * N + i, Z3[i * SV], Z3[(i + 1) * SV], Z3[(i + 2) * SV], Z3[(i + 3) * SV]);
          }
        }
        FirstStep = false;
        nacc++;
        if (rp.gustafsson) {
          if (nacc > 1) {
            double FacGus = __dmul_rn(__dmul_rn(rp.facsafe, __ddiv_rn(H, Hacc)),
                                      fpm::pow_neg_inv(__ddiv_rn(__dmul_rn(Err, Err), ErrOld), 4.0));
            FacGus = fmin(rp.facmax, fmax(rp.facmin, FacGus));
            Fac = fmin(Fac, FacGus);
            Hnew = __dmul_rn(Fac, H);
          }
          Hacc = H;
          ErrOld = fmax(1.0e-2, Err);
        }
        Hold = H;
        T = __dadd_rn(T, H);
#pragma unroll 4
        for (int i = 0; i < N; ++i) {
          double y = Y[i * SV];
          if (rp.D[1] != 0.0) y = __dadd_rn(y, __dmul_rn(rp.D[1], Z1[i * SV]));
          if (rp.D[2] != 0.0) y = __dadd_rn(y, __dmul_rn(rp.D[2], Z2[i * SV]));
          if (rp.D[3] != 0.0) y = __dadd_rn(y, __dmul_rn(rp.D[3], Z3[i * SV]));
          Y[i * SV] = y;
        }
        if (rp.start_newton) {   // RK_Interpolate('make') :889-900
          const double c1 = rp.C[1], c2 = rp.C[2], c3 = rp.C[3];
          const double den = __dmul_rn(__dmul_rn(__dsub_rn(c3, c2), __dsub_rn(c2, c1)), __dsub_rn(c1, c3));
          const double c1s = __dmul_rn(c1, c1), c2s = __dmul_rn(c2, c2), c3s = __dmul_rn(c3, c3);
#pragma unroll 2
          for (int i = 0; i < N; ++i) {
            const double z1 = Z1[i * SV], z2 = Z2[i * SV], z3 = Z3[i * SV];
            double s = __dadd_rn(__dmul_rn(__dmul_rn(-c3s, c2), z1), __dmul_rn(__dmul_rn(z3, c2), c1s));
            s = __dadd_rn(s, __dmul_rn(__dmul_rn(c2s, c3), z1));
            s = __dsub_rn(s, __dmul_rn(__dmul_rn(c2s, c1), z3));
            s = __dadd_rn(s, __dmul_rn(__dmul_rn(c3s, c1), z2));
            s = __dsub_rn(s, __dmul_rn(__dmul_rn(z2, c3), c1s));
            CONT[i * SV] = __dsub_rn(__ddiv_rn(s, den), z3);
            const double d32 = __dsub_rn(z3, z2), d13 = __dsub_rn(z1, z3), d21 = __dsub_rn(z2, z1);
            CONT[(N + i) * SV] = __ddiv_rn(-__dadd_rn(__dadd_rn(__dmul_rn(c1s, d32), __dmul_rn(c2s, d13)), __dmul_rn(c3s, d21)), den);
            CONT[(2 * N + i) * SV] = __ddiv_rn(__dadd_rn(__dadd_rn(__dmul_rn(c1, d32), __dmul_rn(c2, d13)), __dmul_rn(c3, d21)), den);
          }
        }
        error_scale();
        texit = T; hnew_out = Hnew; hexit = H;
        Hnew = __dmul_rn(Tdirection, fmin(fmax(fabs(Hnew), rp.hmin), rp.hmax));
        if (Reject) Hnew = __dmul_rn(Tdirection, fmin(fabs(Hnew), fabs(H)));
        Reject = false;
        if (__dmul_rn(__dsub_rn(__dadd_rn(T, __ddiv_rn(Hnew, rp.qmin)), Tend), Tdirection) >= 0.0) {
          H = __dsub_rn(Tend, T);
        } else {
          const double Hratio = __ddiv_rn(Hnew, H);
          SkipLU = (Theta <= rp.thetamin) && (Hratio >= rp.qmin) && (Hratio <= rp.qmax);
          if (rp.tlm) SkipLU = false;   // "For TLM: do not skip LU"
          if (!SkipLU) H = Hnew;
        }
        SkipJac = false;
      } else {           // rejected (:748-757)
        if (FirstStep || Reject) H = __dmul_rn(rp.facrej, H); else H = Hnew;
        Reject = true; SkipJac = true; SkipLU = false;
        if (nacc >= 1) nrej++;
      }
    }
    // ---- results
    ierr_out[cell] = ierr;
    if (ierr != RK_IERR_IRREGULAR && ierr != IERR_TAPE_OVERFLOW) {   // a handed-back system keeps its initial state for its second pass
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

// host: INTEGRATE's presets and override rule (:58-74) followed by the option handling of RungeKutta (:260-451).
// Returns the reference's IERR before integrating (0 = go on, < 0 = the last option error: the reference keeps checking
// after an error and returns the last code).
// adj != nullptr: INTEGRATE_ADJ of ADJ/RK_ADJ instead (presets :63-70, no Lobatto-3A :411-423, Max_no_steps 10000 :426,
// ICNTRL(7) adjoint solve and ICNTRL(9) adjoint type :456-487); *adj receives ICNTRL(7) (0..3) and ICNTRL(9) (0..2).
struct RkAdjOptions { int solve, type; };
struct RkTlmOptions { int direct, newton_est, trunc_err; };   // ICNTRL(7), (9), (12) of TLM/RK_TLM
// tlm != nullptr: INTEGRATE_TLM of TLM/RK_TLM (presets :62-73: ICNTRL(5) = 8, (7) = 1 direct TLM solve, (10) = 1, (11) = 1 i.e. the
// classic controller; no Lobatto-3A :321-334; Max_no_steps 200000 :336).
inline int rk_parse_options(int N, const int* icntrl_u, const double* rcntrl_u, double tin, double tout, const double* atol,
                            const double* rtol, RkParams& rp, RkAdjOptions* adj = nullptr, RkTlmOptions* tlm = nullptr) {
  int ic[20];
  double rc[20];
  std::memset(ic, 0, sizeof ic);
  std::memset(rc, 0, sizeof rc);
  ic[4] = 8;    // ICNTRL(5): max no. of Newton iterations
  ic[9] = 1;    // ICNTRL(10): SDIRK error estimation
  if (adj) { ic[6] = 1; ic[8] = 2; ic[10] = 1; }   // fixed-sweep adjoint solve, discrete adjoint, classic controller
  if (tlm) { ic[6] = 1; ic[10] = 1; }              // direct TLM solve, classic controller
  for (int i = 0; i < 20; ++i) {
    if (icntrl_u && icntrl_u[i] > 0) ic[i] = icntrl_u[i];
    if (rcntrl_u && rcntrl_u[i] > 0) rc[i] = rcntrl_u[i];
  }
  std::memset(&rp, 0, sizeof rp);
  int ierr = 0;
  rp.vector_tol = (ic[1] == 0);
  int method = 0;
  switch (ic[2]) {
    case 1: method = RK_R2A; break;
    case 0: case 2: method = RK_L3C; break;   // the actual default
    case 3: method = RK_GAU; break;
    case 4: method = RK_R1A; break;
    case 5: if (adj || tlm) ierr = -13; else method = RK_L3A; break;
    default: ierr = -13;
  }
  if (method) {
    const RkTableauData& t = RK_TABLEAU[2 * (method - 1) + 1];   // SdirkError = .TRUE.
    rp.method = method;
    std::memcpy(rp.T, t.T, sizeof rp.T);
    std::memcpy(rp.TinvAinv, t.TinvAinv, sizeof rp.TinvAinv);
    std::memcpy(rp.Tinv, t.Tinv, sizeof rp.Tinv);
    std::memcpy(rp.AinvT, t.AinvT, sizeof rp.AinvT);
    std::memcpy(rp.B, t.B, sizeof rp.B);
    std::memcpy(rp.A, t.A, sizeof rp.A);
    std::memcpy(rp.C, t.C, sizeof rp.C);
    std::memcpy(rp.D, t.D, sizeof rp.D);
    std::memcpy(rp.Bgam, t.Bgam, sizeof rp.Bgam);
    std::memcpy(rp.Theta, t.Theta, sizeof rp.Theta);
    std::memcpy(rp.F, t.F, sizeof rp.F);
    rp.Gamma = t.Gamma; rp.Alpha = t.Alpha; rp.Beta = t.Beta; rp.ELO = t.ELO;
  }
  if (ic[3] == 0) rp.max_no_steps = adj ? 10000 : 200000; else { rp.max_no_steps = ic[3]; if (rp.max_no_steps <= 0) ierr = -1; }
  rp.newton_maxit = ic[4];
  if (rp.newton_maxit <= 0) ierr = -2;
  rp.start_newton = (ic[5] == 0);
  rp.gustafsson = (ic[10] == 0);
  rp.roundoff = DBL_EPSILON * 0.5;
  const double span = std::fabs(tout - tin);
  rp.hmin = (rc[0] == 0.0) ? 0.0 : std::fmin(std::fabs(rc[0]), span);
  rp.hmax = (rc[1] == 0.0) ? span : std::fmin(std::fabs(rc[1]), span);
  rp.hstart = (rc[2] == 0.0) ? 0.0 : std::fmin(std::fabs(rc[2]), span);
  rp.facmin = (rc[3] == 0.0) ? 0.2 : rc[3];
  rp.facmax = (rc[4] == 0.0) ? 8.0 : rc[4];
  rp.facrej = (rc[5] == 0.0) ? 0.1 : rc[5];
  rp.facsafe = (rc[6] == 0.0) ? 0.9 : rc[6];
  if (rp.facmax < 1.0 || rp.facmin > 1.0 || rp.facsafe <= 1.0e-3 || rp.facsafe >= 1.0) ierr = -4;
  if (rc[7] == 0.0) rp.thetamin = 1.0e-3; else { rp.thetamin = rc[7]; if (rp.thetamin <= 0.0 || rp.thetamin >= 1.0) ierr = -5; }
  if (rc[8] == 0.0) rp.newtontol = 3.0e-2; else { rp.newtontol = rc[8]; if (rp.newtontol <= rp.roundoff) ierr = -6; }
  rp.qmin = (rc[9] == 0.0) ? 1.0 : rc[9];
  rp.qmax = (rc[10] == 0.0) ? 1.2 : rc[10];
  if (rp.qmin > 1.0 || rp.qmax < 1.0) ierr = -7;
  const int uplim = rp.vector_tol ? N : 1;
  for (int i = 0; i < uplim; ++i)
    if (atol[i] <= 0.0 || rtol[i] <= 10.0 * rp.roundoff) ierr = -8;
  rp.tstart = tin;
  rp.tend = tout;
  if (tlm) { tlm->direct = (ic[6] != 0); tlm->newton_est = (ic[8] != 0); tlm->trunc_err = (ic[11] != 0); rp.tlm = 1; }
  if (adj) {
    adj->solve = ic[6];
    adj->type = ic[8];
    if (ic[6] < 0 || ic[6] > 3 || ic[8] < 0 || ic[8] > 2) ierr = -9;   // the reference returns at once with -9
  }
  return ierr;
}

}  // namespace fatode
