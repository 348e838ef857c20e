// Backward sweep of the SDIRK discrete adjoint with SIMPLIFIED-NEWTON adjoint stages (ADJ/SDIRK_ADJ, ICNTRL(7) /= 1).
//
// Replaces SDIRK_DadjInt (ADJ/SDIRK_ADJ/SDIRK_ADJ_f90_Integrator.F90:786-995) in the mode the per-stage-factorisation kernels of
// sdirk_cell_adj.cuh do not cover: per popped step ONE factorisation, of 1/(h gamma) I - J(T, Y) (SDIRK_PrepareMatrix :838-845,
// singular retries included), and per stage i = S..1 and adjoint column
//   G = h J_i^T ( b_i Lambda + sum_{j>i} a_ji U_j ),   U_i from U <- U - E^-T [ (U - h gamma J_i^T U - G) / (h gamma) ]   (:880-935)
// iterated until NewtonRate * increment <= NewtonTol (at least four iterations), the increments measured in the column's own
// scale 1 / (ATOL_adj + RTOL_adj |Lambda|) (:869-871).  A column whose iteration does not converge is solved directly after
// LSS_Decomp(HGamma) (:938-951) — HGamma is a local the Newton mode never assigns (it is set at :855, inside the direct branch),
// so upstream this re-factorises garbage * I - J_i; restated with HGamma = 0 (a zero-initialised stack), as oracle/sdirk_adj.hpp
// does.  The new factors REPLACE the step's: the columns that follow, and the later stages of the step, iterate with them.
//
// Mapping: one system per warp, reading the forward sweep's tape records (T, H, Y, Z_1..Z_S) directly.  Per step / stage the warp
// evaluates the Jacobians from the mechanism tables and factorises with DGETF2's partial pivoting (general pivoting: one kernel for
// every system); then lane m owns adjoint column m: Lambda, Mu, U_1..U_S, G in the lane's shared-memory slots, the working
// vectors in registers, J_i and the factors (in pivoted row order) read with broadcast loads.  The columns run concurrently; the
// reference's sequential semantics are restored afterwards: if column f is the first to fail, columns < f stand, column f takes
// the direct solve with the new factors and the columns > f are iterated AGAIN with them.  Arithmetic in the reference's order
// (jt_vec_seq = DGEMV('T') row sums, DGETRS('T') in dot-product form with true divisions, no fma): Lambda and every counter are
// bit-identical to the CPU restatement, Mu to rounding (its rank-one accumulation is regrouped as in the other adjoint sweeps).
#include <cuda_runtime.h>
#include <algorithm>
#include <string>
#include "../../include/fatode_b200.h"
#include "units_internal.h"
#include "ros_warp_fwd.cuh"
#include "sdirk_params.h"

namespace fatode {
__device__ __forceinline__ void sdn_st4(double* p, double a, double b, double c, double d) { p[0] = a; p[1] = b; p[2] = c; p[3] = d; }
}  // namespace fatode
#define CBM4_CELL_FN __device__ __noinline__
#define CBM4_CELL_INL __device__ __forceinline__
#define CBM4_CELL_RC(r) 0.0
#define CBM4_CELL_HESS(h) 0.0
#define CBM4_CELL_ST4(p, a, b, c, d) fatode::sdn_st4(p, a, b, c, d)
#include "gen/cbm4_cell_gen.h"

namespace fatode {

constexpr int SDN_WARPS = 2;
constexpr int SDN_NV = 8;                    // per-lane vectors: Lambda, Mu, U_1..U_5, G
constexpr int SDN_LAM = 0, SDN_MU = 1, SDN_U = 2, SDN_G = 7;

template <class Model>
struct SdnSmem {
  static constexpr int OFF_MODEL = 0;
  static constexpr int OFF_JV = (Model::WS_DOUBLES + 1) / 2 * 2;        // J(T, Y)                [276]
  static constexpr int OFF_J1 = OFF_JV + 276;                           // J(T + c_i H, Y + Z_i)  [276]
  static constexpr int OFF_A = OFF_J1 + 276;                            // factorisation tile [32][LDA]
  static constexpr int OFF_LUE = OFF_A + 32 * LDA;                      // L\U of the step matrix in pivoted row order [32][32]
  static constexpr int OFF_LUF = OFF_LUE + 32 * 32;                     // L\U of the fall-back matrix
  static constexpr int OFF_WHO = OFF_LUF + 32 * 32;                     // scratch pivot rows of the factorisation (32 ints)
  static constexpr int OFF_WHOE = OFF_WHO + 16;
  static constexpr int OFF_WHOF = OFF_WHOE + 16;
  static constexpr int OFF_AP = OFF_WHOF + 16;                          // reactant products of JACP per stage [5][32]
  static constexpr int OFF_TMP = OFF_AP + 5 * 32;                       // Y + Z_i [32]
  static constexpr int OFF_PV = OFF_TMP + 32;                           // per-lane vectors [SDN_NV][32][32 lanes]
  static constexpr int DOUBLES = OFF_PV + SDN_NV * 32 * 32;
};

struct SdnArgs {
  int ncells;                 // systems of this batch
  const long long* rec_off;   // [ncells + 1]
  const int* slot;            // tape slot of each system
  const int* cell;            // its index in the caller's arrays
  const double* tape; int tape_cap;
  double* lambda; double* mu; const double* atol_adj; const double* rtol_adj;
  int* istatus; int* ierr; unsigned long long* queue;
  int nadj, with_mu, np;
};

// DGETRS('T') for one right-hand side in registers: U^T y = b and L^T z = y in netlib's dot-product form on the factors in
// pivoted row order; the caller scatters z back through the pivot rows (x[who[r]] = z[r], DLASWP backwards)
__device__ __forceinline__ void sdn_solve_t(const double* __restrict__ a, double (&x)[32]) {
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    double t = x[i];
#pragma unroll
    for (int k = 0; k < i; ++k) t = t - a[k + 32 * i] * x[k];
    x[i] = t / a[i + 32 * i];
  }
#pragma unroll
  for (int i = 31; i >= 0; --i) {
    double t = x[i];
#pragma unroll
    for (int k = i + 1; k < 32; ++k) t = t - a[k + 32 * i] * x[k];
    x[i] = t;
  }
}

template <class Model>
__global__ void __launch_bounds__(SDN_WARPS * 32, 1)
k_sdirk_adj_newton(SdirkParams sp, typename Model::Const mc, SdirkAdjCoef co, SdnArgs a) {
  extern __shared__ __align__(16) double smem[];
  typedef SdnSmem<Model> L;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* const sm = smem + (size_t)warp * L::DOUBLES;
  double* const ws = sm + L::OFF_MODEL;
  double* const jv = sm + L::OFF_JV;
  double* const j1 = sm + L::OFF_J1;
  double* const A = sm + L::OFF_A;
  double* const LUE = sm + L::OFF_LUE;
  double* const LUF = sm + L::OFF_LUF;
  int* const who = reinterpret_cast<int*>(sm + L::OFF_WHO);
  int* const whoE = reinterpret_cast<int*>(sm + L::OFF_WHOE);
  int* const whoF = reinterpret_cast<int*>(sm + L::OFF_WHOF);
  double* const AP = sm + L::OFF_AP;
  double* const tmpv = sm + L::OFF_TMP;
  double* const pv = sm + L::OFF_PV + lane;
  double* const LAM = pv + (size_t)SDN_LAM * 32 * 32;
  double* const MU = pv + (size_t)SDN_MU * 32 * 32;
  double* const U = pv + (size_t)SDN_U * 32 * 32;        // U[(s * 32 + i) * 32]
  double* const GV = pv + (size_t)SDN_G * 32 * 32;
  const int S = sp.S;
  const int nadj = a.nadj;
  const bool active = lane < nadj;
  const int maxit = sp.newton_maxit;
  const double* const at = a.atol_adj + (size_t)(active ? lane : 0) * 32;
  const double* const rt = a.rtol_adj + (size_t)(active ? lane : 0) * 32;

  for (;;) {
    unsigned long long q0 = 0;
    if (lane == 0) q0 = atomicAdd(a.queue, 1ull);
    q0 = __shfl_sync(FATODE_FULL_MASK, q0, 0);
    if (q0 >= (unsigned long long)a.ncells) break;
    const int sys = (int)q0;
    const long long nrec = a.rec_off[sys + 1] - a.rec_off[sys];
    const int cell = a.cell[sys];
    const double* const tape = a.tape + (size_t)a.slot[sys] * a.tape_cap * SD_REC;
    Model::init(ws, mc, lane);
    // ADJINIT of the example drivers: Lambda = I, Mu = 0
#pragma unroll 4
    for (int i = 0; i < 32; ++i) { LAM[i * 32] = (i == lane) ? 1.0 : 0.0; MU[i * 32] = 0.0; }
    int njac = 0, nstp = 0, ndec = 0, nsol = 0, nsng = 0;
    int ierr = 1;
    int pos = lane;
    __syncwarp();

    for (long long r = nrec - 1; r >= 0 && ierr == 1; --r) {
      const double* rec = tape + (size_t)r * SD_REC;
      const double T = rec[0];
      double H = rec[1];
      const double y = rec[4 + lane];
      {   // SDIRK_PrepareMatrix (:1209-1245): this family re-evaluates the Jacobian after a singular factorisation
        int ConsecutiveSng = 0, info = 1;
        while (info != 0) {
          const double HGammaInv = __ddiv_rn(1.0, __dmul_rn(H, sp.Gamma));
          Model::set_time(ws, T, lane);
          Model::set_y(ws, y, lane);
          Model::jac(ws, jv, lane);
          njac++;
          __syncwarp();
          Model::template scatter_E<LDA>(jv, HGammaInv, A, lane);
          info = warp_lu_factor(A, who, pos, lane, nullptr);
          ndec++;
          if (info != 0) {
            nsng++;
            if (++ConsecutiveSng >= 6) break;
            H = __dmul_rn(0.5, H);
          }
        }
        if (info != 0) { ierr = -8; break; }
        __syncwarp();
        const int src = who[lane];
        whoE[lane] = src;
#pragma unroll 8
        for (int k = 0; k < 32; ++k) LUE[lane + 32 * k] = A[k * LDA + src];
        __syncwarp();
      }
      nstp++;
      const double* curLU = LUE;
      const int* curWho = whoE;
      const double hg = H * sp.Gamma;
      const double hgi = 1.0 / (H * sp.Gamma);
#pragma unroll 1
      for (int s = S - 1; s >= 0 && ierr == 1; --s) {
        // ---- stage Jacobian and the reactant products of JACP at (T + c_i H, Y + Z_i)   (:849-857)
        const double ys = __dadd_rn(y, rec[4 + 32 + s * 32 + lane]);
        Model::set_time(ws, __dadd_rn(T, __dmul_rn(sp.C[s], H)), lane);
        Model::set_y(ws, ys, lane);
        tmpv[lane] = ys;
        Model::jac(ws, j1, lane);
        njac++;
        __syncwarp();
        if (lane == 0 && a.with_mu) cbm4_cell::arp_values<1>(tmpv, mc.fix, AP + s * 32);
        __syncwarp();
        double* const Us = U + (size_t)s * 32 * 32;
        // ---- G = H J_i^T ( b_i Lambda + sum_{j>i} a_ji U_j )   (:873-883)
        if (active) {
          double x[32], o[32];
          const double b = co.B[s];
#pragma unroll
          for (int i = 0; i < 32; ++i) x[i] = b * LAM[i * 32];
          for (int j = s + 1; j < S; ++j) {
            const double c = co.A[j][s];
            if (c != 0.0) {
              const double* Uj = U + (size_t)j * 32 * 32;
#pragma unroll
              for (int i = 0; i < 32; ++i) x[i] = x[i] + c * Uj[i * 32];
            }
          }
          asm volatile("" ::: "memory");
          cbm4_cell::jt_vec_seq(j1, x, o);
#pragma unroll
          for (int i = 0; i < 32; ++i) GV[i * 32] = H * o[i];
        }
        int start = 0;
        for (;;) {   // columns start..NADJ-1 with the factors in force; a failing column installs new ones for those after it
          int iters = 0;
          bool failed = false;
          if (active && lane >= start) {
            double x[32], o[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) x[i] = 0.0;
            double rate = 0.0, inc = 0.0, incOld = 0.0;
            bool done = false;
#pragma unroll 1
            for (int it = 1; it <= maxit; ++it) {
              asm volatile("" ::: "memory");   // keeps the loop-invariant operands (J_i, L\U) in shared memory instead of spilled registers
              cbm4_cell::jt_vec_seq(j1, x, o);
#pragma unroll
              for (int i = 0; i < 32; ++i) o[i] = hgi * ((x[i] - hg * o[i]) - GV[i * 32]);
              asm volatile("" ::: "memory");
              sdn_solve_t(curLU, o);
#pragma unroll
              for (int r2 = 0; r2 < 32; ++r2) Us[curWho[r2] * 32] = o[r2];      // DLASWP backwards
              iters++;
              double e = 0.0;
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                o[i] = Us[i * 32];
                const int t = sp.vector_tol ? i : 0;
                const double sc = 1.0 / (at[t] + rt[t] * fabs(LAM[i * 32]));
                const double qq = o[i] * sc;
                e = e + qq * qq;
              }
              inc = fmax(sqrt(e / 32.0), 1.0e-10);
              if (it == 1) {
                rate = 2.0;
              } else {
                const double th = inc / incOld;
                if (th < 0.99) {
                  rate = th / (1.0 - th);
                  const double pred = inc * sdirk_powi(th, maxit - it) / (1.0 - th);
                  if (pred >= sp.newtontol) break;
                } else {
                  break;
                }
              }
              incOld = inc;
#pragma unroll
              for (int i = 0; i < 32; ++i) x[i] = x[i] - o[i];
              done = (rate * inc <= sp.newtontol);
              if (it >= 4 && done) break;
            }
#pragma unroll
            for (int i = 0; i < 32; ++i) Us[i * 32] = x[i];
            failed = !done;
          }
          const unsigned fm = __ballot_sync(FATODE_FULL_MASK, failed);
          const int f = fm ? __ffs(fm) - 1 : 31;
          nsol += __reduce_add_sync(FATODE_FULL_MASK, (active && lane >= start && lane <= f) ? iters : 0);
          if (!fm) break;
          // ---- column f: direct solution after LSS_Decomp(HGamma), HGamma = 0 in this mode (:938-951)
          __syncwarp();
          Model::template scatter_E<LDA>(j1, 0.0, A, lane);
          const int info = warp_lu_factor(A, who, pos, lane, nullptr);
          ndec++;
          if (info != 0) { ierr = -8; break; }
          __syncwarp();
          {
            const int src = who[lane];
            whoF[lane] = src;
#pragma unroll 8
            for (int k = 0; k < 32; ++k) LUF[lane + 32 * k] = A[k * LDA + src];
          }
          __syncwarp();
          curLU = LUF;
          curWho = whoF;
          if (lane == f) {
            double o[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = hgi * GV[i * 32];
            asm volatile("" ::: "memory");
            sdn_solve_t(curLU, o);
#pragma unroll
            for (int r2 = 0; r2 < 32; ++r2) Us[curWho[r2] * 32] = o[r2];
          }
          nsol++;
          start = f + 1;
          __syncwarp();
        }
        __syncwarp();
      }
      if (ierr != 1) break;
      if (active) {
        // Mu += V_1 + .. + V_S with V_i = JACP_i^T ( sum_{j>=i} (H a_ji) U_j + (H b_i) Lambda )   (:961-983)
        if (a.with_mu) {
          double x[32];
#pragma unroll 1
          for (int s = 0; s < S; ++s) {
#pragma unroll
            for (int i = 0; i < 32; ++i) x[i] = 0.0;
            for (int j = s; j < S; ++j) {
              const double c = H * co.A[j][s];
              if (c != 0.0) {
                const double* Uj = U + (size_t)j * 32 * 32;
#pragma unroll
                for (int i = 0; i < 32; ++i) x[i] = x[i] + c * Uj[i * 32];
              }
            }
            {
              const double c = H * co.B[s];
              if (c != 0.0) {
#pragma unroll
                for (int i = 0; i < 32; ++i) x[i] = x[i] + c * LAM[i * 32];
              }
            }
            const double* APs = AP + s * 32;
            Cbm4Warp::stoich_cols(x, [&](auto pc, double spv) {
              constexpr int p = decltype(pc)::value;
              MU[p * 32] = fma(APs[p], spv, MU[p * 32]);
            });
          }
        }
        for (int s = 0; s < S; ++s) {
          const double* Us = U + (size_t)s * 32 * 32;
#pragma unroll 4
          for (int i = 0; i < 32; ++i) LAM[i * 32] = LAM[i * 32] + Us[i * 32];
        }
      }
      __syncwarp();
    }
    if (active) {
      double* lo = a.lambda + ((size_t)cell * nadj + lane) * 32;
#pragma unroll 4
      for (int i = 0; i < 32; ++i) lo[i] = LAM[i * 32];
      if (a.with_mu) {
        double* mo = a.mu + ((size_t)cell * nadj + lane) * a.np;
        for (int p = 0; p < a.np; ++p) mo[p] = MU[p * 32];
      }
    }
    if (lane == 0) {
      int* is_ = a.istatus + (size_t)cell * 20;
      is_[Nstp] += nstp;
      is_[Njac] += njac;
      is_[Ndec] += ndec;
      is_[Nsol] += nsol;
      is_[Nsng] += nsng;
      a.ierr[cell] = ierr;
    }
    __syncwarp();
  }
}

int sdirk_adj_newton_device(const SdirkParams& sp, const Cbm4Const& mc, const SdirkAdjCoef& co, int nsys, const long long* d_rec_off,
                            const int* d_slot, const int* d_cell, const double* d_tape, int tape_cap, int nadj, int with_mu, int np,
                            double* d_lambda, double* d_mu, const double* d_atol_adj, const double* d_rtol_adj, int* d_istatus,
                            int* d_ierr, unsigned long long* d_queue, cudaStream_t st, UnitTiming& tm, std::string& err) {
  if (nsys <= 0) return 0;
  typedef Cbm4Warp Model;
  int dev = 0, sms = 0;
  UNIT_TRY(cudaGetDevice(&dev));
  UNIT_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const size_t smem = sizeof(double) * (size_t)SdnSmem<Model>::DOUBLES * SDN_WARPS;
  UNIT_TRY(cudaFuncSetAttribute(k_sdirk_adj_newton<Model>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  UNIT_TRY(cudaMemsetAsync(d_queue, 0, sizeof(unsigned long long), st));
  SdnArgs a{nsys, d_rec_off, d_slot, d_cell, d_tape, tape_cap, d_lambda, d_mu, d_atol_adj, d_rtol_adj, d_istatus, d_ierr, d_queue,
            nadj, with_mu, np};
  cudaEvent_t e0, e1;
  UNIT_TRY(cudaEventCreate(&e0));
  UNIT_TRY(cudaEventCreate(&e1));
  UNIT_TRY(cudaEventRecord(e0, st));
  const int blocks = std::min(sms, (nsys + SDN_WARPS - 1) / SDN_WARPS);
  k_sdirk_adj_newton<Model><<<blocks, SDN_WARPS * 32, smem, st>>>(sp, mc, co, a);
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
