// Lock-step SDIRK tangent linear model: ONE kernel advances the state and all NTLM columns of a system together.
//
// Replaces SDIRK_IntegratorTLM (TLM/SDIRK_TLM/SDIRK_TLM_f90_Integrator.F90:474-911) in the modes the split state-sweep /
// preparation / column-sweep path of sdirk_cell_fwd.cuh + sdirk_cell_tlm.cuh cannot take, because there the columns (or a
// second factorisation they need) feed back into the step sequence:
//   * ICNTRL(7) /= 0  direct solution for the TLM variables (:659-705): a second LU, of 1/(h gamma) I - J(T + c_i h, Y + Z_i),
//                     per stage (LSS_Decomp_TLM / LSS_Solve_TLM, TLM/SDIRK_TLM/LS_Solver.F90:44-79); a singular one halves H
//                     in the middle of the step (:674-683);
//   * ICNTRL(9) /= 0  the columns' modified-Newton iterations are convergence-controlled (:735-741, :755-779, :785-800): a
//                     column that does not converge rejects the attempt;
//   * ICNTRL(12) /= 0 the columns' embedded error estimates enter the accept/reject decision (:825-837, SDIRK_ErrorNorm_TLM
//                     :947-960).
// Any combination; general partial pivoting throughout (no static pivot pattern, so no second pass).
//
// Mapping: one system per warp, as in ros_tlm_lock.cu.  State part: lane = component (mechanism evaluated once per warp from
// tables, E in a shared-memory tile, DGETF2 by warp arg-max, DGETRS('N') by shuffles).  Column part: lane m = column m of
// Y_tlm; its vectors Y_tlm, Z_tlm_1..S and G live in the lane's shared-memory slots, the working vectors in registers; the
// operands every column shares (the stage Jacobian, L\U of E and of the stage matrix re-ordered into pivoted row order) are
// read with broadcast loads.  The columns decide the step sequence, so their arithmetic is the reference's, operation by
// operation, no fma: matmul(fjac1, g) as sequential row sums over the 276 structural entries, DAXPYs in the Fortran's order,
// DGETRS('N') as DLASWP + the two column-sweep DTRSMs with true divisions.  State, ISTATUS, RSTATUS and Y_tlm are
// bit-identical to the CPU restatement (oracle/sdirk_tlm.hpp).
#include <cuda_runtime.h>
#include <algorithm>
#include <string>
#include "../../include/fatode_b200.h"
#include "units_internal.h"
#include "ros_warp_fwd.cuh"
#include "sdirk_params.h"

#define CBM4_CELL_FN __device__ __noinline__
#define CBM4_CELL_INL __device__ __forceinline__
#define CBM4_CELL_RC(r) 0.0
#define CBM4_CELL_HESS(h) 0.0
#define CBM4_CELL_ST4(p, a, b, c, d) ((void)0)
#include "gen/cbm4_cell_gen.h"

namespace fatode {

constexpr int SLK_WARPS = 2;                 // systems per CTA (one per warp)
constexpr int SLK_NV = 7;                    // per-lane vectors: Y_tlm, Z_tlm_1..5, G
constexpr int SLK_Y = 0, SLK_Z = 1, SLK_G = 6;

template <class Model>
struct SdLockSmem {
  static constexpr int OFF_MODEL = 0;
  static constexpr int OFF_JV = (Model::WS_DOUBLES + 1) / 2 * 2;        // J(T,Y)                          [276]
  static constexpr int OFF_J1 = OFF_JV + 276;                           // J(T + c_i H, Y + Z_i)           [276]
  static constexpr int OFF_A = OFF_J1 + 276;                            // E / L\U tile [32][LDA]
  static constexpr int OFF_LU = OFF_A + 32 * LDA;                       // L\U of E in pivoted row order, column-major [32][32]
  static constexpr int OFF_A2 = OFF_LU + 32 * 32;                       // stage matrix of the direct TLM solve, tile
  static constexpr int OFF_LU2 = OFF_A2 + 32 * LDA;                     // ... in pivoted row order
  static constexpr int OFF_WHO = OFF_LU2 + 32 * 32;                     // 32 ints
  static constexpr int OFF_WHO2 = OFF_WHO + 16;                         // 32 ints
  static constexpr int OFF_Z = OFF_WHO2 + 16;                           // Z_1..Z_5 of the state [5][32]
  static constexpr int OFF_PV = OFF_Z + 5 * 32;                         // per-lane vectors [SLK_NV][32][32 lanes]
  static constexpr int DOUBLES = OFF_PV + SLK_NV * 32 * 32;
};

struct SdLockArgs {
  long ncells; const int* ids; int ntlm; int direct, newton_est, trunc;
  double* y; double* y_tlm; const double* atol; const double* rtol; const double* atol_tlm; const double* rtol_tlm;
  int* istatus; double* rstatus; int* ierr; unsigned long long* queue;
};

// DGETRS('N') for one right-hand side held in registers (already in pivoted row order), L\U column-major in pivoted row order
__device__ __forceinline__ void sdlock_solve_n(const double* __restrict__ a, double (&x)[32]) {
#pragma unroll
  for (int k = 0; k < 31; ++k) {
    const double bk = x[k];
#pragma unroll
    for (int i = k + 1; i < 32; ++i) x[i] = x[i] - bk * a[i + 32 * k];
  }
#pragma unroll
  for (int k = 31; k >= 0; --k) {
    const double bk = (x[k] != 0.0) ? x[k] / a[k + 32 * k] : x[k];
    x[k] = bk;
#pragma unroll
    for (int i = 0; i < k; ++i) x[i] = x[i] - bk * a[i + 32 * k];
  }
}

template <class Model>
__global__ void __launch_bounds__(SLK_WARPS * 32, 1)
k_sdirk_tlm_lock(SdirkParams sp, typename Model::Const mc, SdLockArgs a) {
  extern __shared__ __align__(16) double smem[];
  typedef SdLockSmem<Model> L;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* const sm = smem + (size_t)warp * L::DOUBLES;
  double* const ws = sm + L::OFF_MODEL;
  double* const jv = sm + L::OFF_JV;
  double* const j1 = sm + L::OFF_J1;
  double* const A = sm + L::OFF_A;
  double* const LU = sm + L::OFF_LU;
  double* const A2 = sm + L::OFF_A2;
  double* const LU2 = sm + L::OFF_LU2;
  int* const who = reinterpret_cast<int*>(sm + L::OFF_WHO);
  int* const who2 = reinterpret_cast<int*>(sm + L::OFF_WHO2);
  double* const Zst = sm + L::OFF_Z + lane;           // Z_s of the lane's component: Zst[s * 32]
  double* const pv = sm + L::OFF_PV + lane;           // element i of the lane's vector v: pv[(v * 32 + i) * 32]
  const int S = sp.S;
  const int ntlm = a.ntlm;
  const bool active = lane < ntlm;
  const double Roundoff = sp.roundoff;
  const int ti = sp.vector_tol ? lane : 0;
  const double atol = a.atol[ti], rtol = a.rtol[ti];
  const double Tfinal = sp.tend;
  const double Tdirection = (sp.tend - sp.tstart) >= 0.0 ? 1.0 : -1.0;
  const double* const at = a.atol_tlm + (size_t)(active ? lane : 0) * 32;
  const double* const rt = a.rtol_tlm + (size_t)(active ? lane : 0) * 32;
  const int maxit = sp.newton_maxit;

  auto norm32 = [&](double v, double scal) -> double {   // SDIRK_ErrorNorm :933-945, components summed in order
    const double q = __dmul_rn(v, scal);
    const double q2 = __dmul_rn(q, q);
    double acc = 0.0;
#pragma unroll
    for (int i = 0; i < 32; ++i) acc = __dadd_rn(acc, shfl_d(q2, i));
    return fmax(sqrt(__ddiv_rn(acc, 32.0)), 1.0e-10);
  };

  for (;;) {
    unsigned long long slot = 0;
    if (lane == 0) slot = atomicAdd(a.queue, 1ull);
    slot = __shfl_sync(FATODE_FULL_MASK, slot, 0);
    if (slot >= (unsigned long long)a.ncells) break;
    const long cell = a.ids ? (long)a.ids[slot] : (long)slot;
    double y = a.y[cell * 32 + lane];
    double* const ycol = a.y_tlm + ((size_t)cell * ntlm + (active ? lane : 0)) * 32;
#pragma unroll 4
    for (int i = 0; i < 32; ++i) pv[(SLK_Y * 32 + i) * 32] = active ? ycol[i] : 0.0;
    int nfun = 0, njac = 0, nstp = 0, nacc = 0, nrej = 0, ndec = 0, nsol = 0, nsng = 0;
    double texit = 0.0, hexit = 0.0, hnew_out = 0.0;
    int pos = lane, pos2 = lane;
    LUMasks mk = {0u, 0u};
    Model::init(ws, mc, lane);
    double T = sp.tstart;
    double H = fmax(fabs(sp.hmin), fabs(sp.hstart));
    if (fabs(H) <= __dmul_rn(10.0, Roundoff)) H = 1.0e-6;
    H = fmin(fabs(H), sp.hmax);
    H = copysign(H, Tdirection);
    bool SkipLU = false, SkipJac = false, Reject = false, FirstStep = true;
    double NewtonRate = 0.0, NewtonIncrement = 0.0, NewtonIncrementOld = 0.0;
    int saveNiter = 0;
    int ierr = 1;
    double scal = __ddiv_rn(1.0, __dadd_rn(atol, __dmul_rn(rtol, fabs(y))));   // SDIRK_ErrorScale :914-929
    __syncwarp();

    while (__dsub_rn(__dmul_rn(__dsub_rn(Tfinal, T), Tdirection), Roundoff) > 0.0) {   // Tloop :520
      if (!SkipLU) {   // the inlined SDIRK_PrepareMatrix (:524-554)
        int ConsecutiveSng = 0, info = 1;
        while (info != 0) {
          const double HGammaInv = __ddiv_rn(1.0, __dmul_rn(H, sp.Gamma));
          if (!SkipJac) {
            Model::set_time(ws, T, lane);
            Model::set_y(ws, y, lane);
            Model::jac(ws, jv, lane);
            njac++;
            __syncwarp();
          }
          Model::template scatter_E<LDA>(jv, HGammaInv, A, lane);
          info = warp_lu_factor(A, who, pos, lane, &mk);
          ndec++;
          if (info != 0) {
            nsng++;
            if (++ConsecutiveSng >= 6) break;
            H = __dmul_rn(0.5, H);
            SkipJac = true; SkipLU = false; Reject = true;
          }
        }
        if (info != 0) { ierr = -8; break; }
        __syncwarp();
        {   // L\U in pivoted row order for the lanes' own triangular sweeps (row r of the copy = physical row who[r])
          const int src = who[lane];
#pragma unroll 8
          for (int k = 0; k < 32; ++k) LU[lane + 32 * k] = A[k * LDA + src];
        }
        __syncwarp();
      }
      if (nstp > sp.max_no_steps) { ierr = -6; break; }
      if (__dadd_rn(T, __dmul_rn(0.1, H)) == T || fabs(H) <= Roundoff) { ierr = -7; break; }

      bool cycle = false;
#pragma unroll 1
      for (int is = 0; is < S && !cycle; ++is) {   // istage = is + 1
        // ---- state: simplified Newton iterations of the stage (:563-650)
        double zi = 0.0, g = 0.0;
        for (int j = 0; j < is; ++j) {
          const double zj = Zst[j * 32];
          g = __dadd_rn(g, __dmul_rn(sp.Theta[is][j], zj));
          if (sp.start_newton) zi = __dadd_rn(zi, __dmul_rn(sp.Alpha[is][j], zj));
        }
        bool NewtonDone = false;
        double Fac = 0.5;
        Model::set_time(ws, __dadd_rn(T, __dmul_rn(sp.C[is], H)), lane);
        {
          const double HG = __dmul_rn(H, sp.Gamma);
          const double HGammaInv = __ddiv_rn(1.0, HG);
          for (int NewtonIter = 1; NewtonIter <= maxit; ++NewtonIter) {
            Model::set_y(ws, __dadd_rn(y, zi), lane);
            double dz = Model::fun(ws, mc, lane);
            nfun++;
            dz = __dadd_rn(__dadd_rn(__dmul_rn(HG, dz), -zi), g);
            dz = __dmul_rn(HGammaInv, dz);
            dz = warp_lu_solve(A, who, pos, mk, dz, lane);
            nsol++;
            NewtonIncrement = norm32(dz, scal);
            if (NewtonIter == 1) {
              NewtonRate = 2.0;
            } else {
              const double Theta = __ddiv_rn(NewtonIncrement, NewtonIncrementOld);
              if (Theta < 0.99) {
                NewtonRate = __ddiv_rn(Theta, __dsub_rn(1.0, Theta));
                const double pred = __ddiv_rn(__dmul_rn(NewtonIncrement, sdirk_powi(Theta, maxit - NewtonIter)), __dsub_rn(1.0, Theta));
                if (pred >= sp.newtontol) {
                  const double Qnewton = fmin(10.0, __ddiv_rn(pred, sp.newtontol));
                  Fac = __dmul_rn(0.8, fpm::pow_neg_inv(Qnewton, (double)(1 + maxit - NewtonIter)));
                  break;
                }
              } else {
                break;
              }
            }
            NewtonIncrementOld = NewtonIncrement;
            zi = __dadd_rn(zi, dz);
            NewtonDone = (__dmul_rn(NewtonRate, NewtonIncrement) <= sp.newtontol);
            if (NewtonDone) { saveNiter = NewtonIter + 1; break; }
          }
        }
        if (!NewtonDone) {
          H = __dmul_rn(Fac, H);
          Reject = true; SkipJac = true; SkipLU = false;
          cycle = true;
          break;
        }
        Zst[is * 32] = zi;
        Model::set_y(ws, __dadd_rn(y, zi), lane);     // TMP = Y + Z_i: the point of the stage Jacobian in both TLM flavours
        double* const zt = pv + (size_t)(SLK_Z + is) * 32 * 32;   // Z_tlm_i of the lane's column; scratch until it is written

        if (a.direct) {
          // ---- direct solution for the TLM variables (:659-705)
          bool SkipJacT = false;
          int ConsecutiveSng = 0, info = 1;
          while (info != 0) {
            const double HGammaInv = __ddiv_rn(1.0, __dmul_rn(H, sp.Gamma));
            if (!SkipJacT) {
              Model::jac(ws, j1, lane);
              njac++;
              __syncwarp();
            }
            Model::template scatter_E<LDA>(j1, HGammaInv, A2, lane);
            info = warp_lu_factor(A2, who2, pos2, lane, nullptr);
            ndec++;
            if (info != 0) {
              nsng++;
              if (++ConsecutiveSng >= 6) break;
              H = __dmul_rn(0.5, H);
              SkipJacT = true; SkipLU = false; Reject = true;
            }
          }
          if (info != 0) { ierr = -8; cycle = true; break; }
          __syncwarp();
          {
            const int src = who2[lane];
#pragma unroll 8
            for (int k = 0; k < 32; ++k) LU2[lane + 32 * k] = A2[k * LDA + src];
          }
          __syncwarp();
          if (active) {
            const double HGammaInv = 1.0 / (H * sp.Gamma);
            double x[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) x[i] = pv[(SLK_Y * 32 + i) * 32];
#pragma unroll 1
            for (int j = 0; j < is; ++j) {
              const double th = sp.Theta[is][j];
              const double* zj = pv + (size_t)(SLK_Z + j) * 32 * 32;
#pragma unroll
              for (int i = 0; i < 32; ++i) x[i] = x[i] + th * zj[i * 32];
            }
#pragma unroll
            for (int i = 0; i < 32; ++i) zt[i * 32] = HGammaInv * x[i];
#pragma unroll
            for (int r = 0; r < 32; ++r) x[r] = zt[who2[r] * 32];       // DLASWP
            sdlock_solve_n(LU2, x);
#pragma unroll
            for (int i = 0; i < 32; ++i) zt[i * 32] = x[i] - pv[(SLK_Y * 32 + i) * 32];
          }
          nsol += ntlm;
          __syncwarp();
          continue;
        }

        // ---- modified-Newton TLM of the stage (:707-803)
        Model::jac(ws, j1, lane);
        njac++;
        __syncwarp();
        int iters = 0;
        bool failed = false;
        double facl = 0.5;
        if (active) {
          const double HG = H * sp.Gamma;
          const double HGammaInv = 1.0 / (H * sp.Gamma);
          double x[32], o[32];
          double* const gv = pv + (size_t)SLK_G * 32 * 32;
#pragma unroll
          for (int i = 0; i < 32; ++i) x[i] = pv[(SLK_Y * 32 + i) * 32];
          cbm4_cell::j_vec_seq(j1, x, o);                                 // DZ = J1 Y_tlm
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = HG * o[i];
#pragma unroll 1
          for (int j = 0; j < is; ++j) {
            const double th = sp.Theta[is][j];
            const double* zj = pv + (size_t)(SLK_Z + j) * 32 * 32;
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = o[i] + th * zj[i * 32];
          }
#pragma unroll
          for (int i = 0; i < 32; ++i) { gv[i * 32] = o[i]; x[i] = 0.0; }
          double rate = 0.0, inc = 0.0, incOld = 0.0;
          bool done = false;
#pragma unroll 1
          for (int it = 1; it <= maxit; ++it) {
            asm volatile("" ::: "memory");   // keeps the loop-invariant operands (J1, L\U) in shared memory instead of spilled registers
            cbm4_cell::j_vec_seq(j1, x, o);                               // DZ = J1 Z_tlm
#pragma unroll
            for (int i = 0; i < 32; ++i) zt[i * 32] = HGammaInv * ((HG * o[i] + gv[i * 32]) - x[i]);
#pragma unroll
            for (int r = 0; r < 32; ++r) o[r] = zt[who[r] * 32];          // DLASWP
            asm volatile("" ::: "memory");
            sdlock_solve_n(LU, o);
            iters++;
            if (a.newton_est) {
              double e = 0.0;
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                const int t = sp.vector_tol ? i : 0;
                const double sc = 1.0 / (at[t] + rt[t] * fabs(pv[(SLK_Y * 32 + i) * 32]));
                const double q = o[i] * sc;
                e = e + q * q;
              }
              inc = fmax(sqrt(e / 32.0), 1.0e-10);
              if (it <= 1) {
                rate = 2.0;
              } else {
                const double th = inc / incOld;
                if (th < 0.99) {
                  rate = th / (1.0 - th);
                  const double pred = inc * sdirk_powi(th, maxit - it) / (1.0 - th);
                  if (pred >= sp.newtontol) {
                    const double Qn = fmin(10.0, pred / sp.newtontol);
                    facl = 0.8 * fpm::pow_neg_inv(Qn, (double)(1 + maxit - it));
                    break;
                  }
                } else {
                  break;
                }
              }
              incOld = inc;
            }
#pragma unroll
            for (int i = 0; i < 32; ++i) x[i] = x[i] + o[i];
            if (a.newton_est) {
              done = (rate * inc <= sp.newtontol);
              if (done) break;
            } else if (it >= saveNiter) {
              break;
            }
          }
#pragma unroll
          for (int i = 0; i < 32; ++i) zt[i * 32] = x[i];
          failed = a.newton_est && !done;
        }
        {   // the columns run one after the other in the reference: the first one that fails ends the attempt (:795-800)
          const unsigned fm = __ballot_sync(FATODE_FULL_MASK, failed);
          const int f = fm ? __ffs(fm) - 1 : 31;
          nsol += __reduce_add_sync(FATODE_FULL_MASK, (active && lane <= f) ? iters : 0);
          if (fm) {
            H = __dmul_rn(shfl_d(facl, f), H);
            Reject = true; SkipJac = true; SkipLU = false;
            cycle = true;
          }
        }
        __syncwarp();
      }
      if (ierr != 1) break;
      if (cycle) continue;

      // ---- error estimation (:812-837)
      nstp++;
      const double HGammaInv = __ddiv_rn(1.0, __dmul_rn(H, sp.Gamma));
      double Yerr = 0.0;
      for (int s = 0; s < S; ++s)
        if (sp.E[s] != 0.0) Yerr = __dadd_rn(Yerr, __dmul_rn(sp.E[s], Zst[s * 32]));
      Yerr = __dmul_rn(HGammaInv, Yerr);
      Yerr = warp_lu_solve(A, who, pos, mk, Yerr, lane);
      nsol++;
      double Err = norm32(Yerr, scal);
      if (a.trunc) {
        double e = 0.0;
        if (active) {
          double o[32];
          double* const gv = pv + (size_t)SLK_G * 32 * 32;
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = 0.0;
#pragma unroll 1
          for (int j = 0; j < S; ++j) {
            const double ej = sp.E[j];
            if (ej != 0.0) {
              const double* zj = pv + (size_t)(SLK_Z + j) * 32 * 32;
#pragma unroll
              for (int i = 0; i < 32; ++i) o[i] = o[i] + ej * zj[i * 32];
            }
          }
#pragma unroll
          for (int i = 0; i < 32; ++i) gv[i * 32] = HGammaInv * o[i];
#pragma unroll
          for (int r = 0; r < 32; ++r) o[r] = gv[who[r] * 32];            // DLASWP
          sdlock_solve_n(LU, o);
          double acc = 0.0;
#pragma unroll
          for (int i = 0; i < 32; ++i) {   // SDIRK_ErrorNorm_TLM: scaled with the error vector itself (:953-955)
            const int t = sp.vector_tol ? i : 0;
            const double sc = 1.0 / (at[t] + rt[t] * fabs(o[i]));
            const double q = o[i] * sc;
            acc = acc + q * q;
          }
          e = fmax(sqrt(acc / 32.0), 1.0e-10);
        }
        nsol += ntlm;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) e = fmax(e, shfl_d(e, lane ^ o));
        Err = fmax(Err, e);
      }
      double Fac = __dmul_rn(sp.facsafe, fpm::pow_neg_inv(Err, sp.ELO));
      Fac = fmax(sp.facmin, fmin(sp.facmax, Fac));
      double Hnew = __dmul_rn(H, Fac);
      if (Err < 1.0) {   // accept (:847-890)
        FirstStep = false;
        nacc++;
        T = __dadd_rn(T, H);
        for (int s = 0; s < S; ++s)
          if (sp.D[s] != 0.0) y = __dadd_rn(y, __dmul_rn(sp.D[s], Zst[s * 32]));
        if (active) {
#pragma unroll 2
          for (int i = 0; i < 32; ++i) {
            double v = pv[(SLK_Y * 32 + i) * 32];
            for (int s = 0; s < S; ++s)
              if (sp.D[s] != 0.0) v = v + sp.D[s] * pv[((SLK_Z + s) * 32 + i) * 32];
            pv[(SLK_Y * 32 + i) * 32] = v;
          }
        }
        scal = __ddiv_rn(1.0, __dadd_rn(atol, __dmul_rn(rtol, fabs(y))));
        Hnew = __dmul_rn(Tdirection, fmin(fabs(Hnew), sp.hmax));
        texit = T; hexit = H; hnew_out = Hnew;
        if (Reject) Hnew = __dmul_rn(Tdirection, fmin(fabs(Hnew), fabs(H)));
        Reject = false;
        if (__dmul_rn(__dsub_rn(__dadd_rn(T, __ddiv_rn(Hnew, sp.qmin)), Tfinal), Tdirection) > 0.0) {
          H = __dsub_rn(Tfinal, T);
        } else {
          SkipLU = false;   // "For TLM: do not skip LU" (:885)
          H = Hnew;
        }
        SkipJac = false;
      } else {           // reject (:892-903)
        if (FirstStep || Reject) H = __dmul_rn(sp.facrej, H); else H = Hnew;
        Reject = true; SkipJac = true; SkipLU = false;
        if (nacc >= 1) nrej++;
      }
      __syncwarp();
    }
    // ---- results (a failed system returns Y and Y_tlm as of its last accepted step, like the reference)
    a.y[cell * 32 + lane] = y;
    if (active) {
#pragma unroll 4
      for (int i = 0; i < 32; ++i) ycol[i] = pv[(SLK_Y * 32 + i) * 32];
    }
    if (lane < 20) {
      int v = 0;
      if (lane == Nfun) v = nfun; else if (lane == Njac) v = njac; else if (lane == Nstp) v = nstp;
      else if (lane == Nacc) v = nacc; else if (lane == Nrej) v = nrej; else if (lane == Ndec) v = ndec;
      else if (lane == Nsol) v = nsol; else if (lane == Nsng) v = nsng;
      a.istatus[cell * 20 + lane] = v;
      double r = 0.0;
      if (lane == Ntexit) r = texit; else if (lane == Nhexit) r = hexit; else if (lane == Nhnew) r = hnew_out;
      a.rstatus[cell * 20 + lane] = r;
    }
    if (lane == 0) a.ierr[cell] = ierr;
    __syncwarp();
  }
}

int sdirk_tlm_lock_device(const SdirkParams& sp, const Cbm4Const& mc, long ncells, const int* d_ids, int ntlm, bool direct,
                          bool newton_est, bool trunc, double* d_y, double* d_ytlm, const double* d_atol, const double* d_rtol,
                          const double* d_atol_tlm, const double* d_rtol_tlm, int* d_istatus, double* d_rstatus, int* d_ierr,
                          cudaStream_t st, UnitTiming& tm, std::string& err) {
  if (ncells <= 0) return 0;
  typedef Cbm4Warp Model;
  int dev = 0, sms = 0;
  UNIT_TRY(cudaGetDevice(&dev));
  UNIT_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const size_t smem = sizeof(double) * (size_t)SdLockSmem<Model>::DOUBLES * SLK_WARPS;
  UNIT_TRY(cudaFuncSetAttribute(k_sdirk_tlm_lock<Model>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  unsigned long long* d_queue = nullptr;
  UNIT_TRY(cudaMalloc(&d_queue, sizeof(unsigned long long)));
  struct Free { void* p; ~Free() { cudaFree(p); } } guard{d_queue};
  UNIT_TRY(cudaMemsetAsync(d_queue, 0, sizeof(unsigned long long), st));
  const bool tols = newton_est || trunc;
  SdLockArgs a{ncells, d_ids, ntlm, direct ? 1 : 0, newton_est ? 1 : 0, trunc ? 1 : 0, d_y, d_ytlm, d_atol, d_rtol,
               tols ? d_atol_tlm : d_atol, tols ? d_rtol_tlm : d_rtol, d_istatus, d_rstatus, d_ierr, d_queue};
  cudaEvent_t e0, e1;
  UNIT_TRY(cudaEventCreate(&e0));
  UNIT_TRY(cudaEventCreate(&e1));
  UNIT_TRY(cudaEventRecord(e0, st));
  const int blocks = (int)std::min<long>((long)sms, (ncells + SLK_WARPS - 1) / SLK_WARPS);
  k_sdirk_tlm_lock<Model><<<blocks, SLK_WARPS * 32, smem, st>>>(sp, mc, a);
  cudaError_t le = cudaGetLastError();
  cudaEventRecord(e1, st);
  cudaError_t se = cudaStreamSynchronize(st);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  UNIT_TRY(le);
  UNIT_TRY(se);
  tm.ms += ms;
  tm.launches += 1;
  return 0;
}

}  // namespace fatode
