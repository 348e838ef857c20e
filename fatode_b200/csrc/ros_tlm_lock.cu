// Lock-step Rosenbrock tangent linear model: ONE kernel advances the state and all NTLM columns of a system together.
//
// Replaces ros_TLM_Int (TLM/ROS_TLM/ROS_TLM_f90_Integrator.F90:462-707) in the cases the split state-sweep / column-sweep path
// of ros_cell_fwd.cuh + ros_cell_adj.cuh (k_ros_tlm_cell) cannot take:
//   * TLM truncation-error control, ICNTRL(12) /= 0 — the PRESET of INTEGRATE_TLM (:70-72): the columns' error norms
//     (ros_ErrorNorm_tlm :745-769, used at :656-665) enter the accept/reject decision, so the state cannot run ahead;
//   * systems whose factorisations leave the static pivot pattern of the per-thread kernels (general partial pivoting here).
//
// Mapping: one system per warp.  The state part is the warp-cooperative general kernel of ros_warp_fwd.cuh (lane = component:
// mechanism evaluated once per warp from tables, E = 1/(h gamma) I - J in a shared-memory tile, DGETF2 with netlib's partial
// pivoting by warp arg-max, DGETRS('N') by shuffles).  The column part gives lane m column m of Y_tlm: its vectors Y_tlm, K_tlm_1..S,
// Fcn0_tlm, Fcn_tlm live in the lane's shared-memory slots, the working vector in registers; the operands every column shares —
// J(T,Y), dJ/dt, the stage Jacobian J(Tau, Ynew), the stage vector K_i of the state, L\U re-ordered into pivoted row order — are
// read from the warp's shared memory with broadcast loads.  Because the columns' error norms decide the step sequence their
// arithmetic is the reference's, operation by operation, no fma: matmul(fjac, g) as libgfortran's column sweep (= sequential row
// sums over the 276 structural entries; the others add exact zeros), DAXPYs in the Fortran's order, Hess_Vec as the restated
// callback, DGETRS('N') as DLASWP + the two column-sweep DTRSMs with true divisions.  State, ISTATUS, RSTATUS are bit-identical to
// the CPU restatement (oracle/ros.hpp::tlm), and so is Y_tlm.
#include <cuda_runtime.h>
#include <algorithm>
#include <string>
#include "../../include/fatode_b200.h"
#include "units_internal.h"
#include "ros_warp_fwd.cuh"

namespace fatode {
static __constant__ double c_lock_rc[81];     // time-independent rate coefficients (only hess_values reads them here)
static __constant__ double c_lock_hess[258];  // Hessian values: rate coefficients only for this mechanism (set per call)
__device__ __forceinline__ void lock_st4(double* p, double a, double b, double c, double d) { p[0] = a; p[1] = b; p[2] = c; p[3] = d; }
}  // namespace fatode
#define CBM4_CELL_FN __device__ __noinline__
#define CBM4_CELL_INL __device__ __forceinline__
#define CBM4_CELL_RC(r) fatode::c_lock_rc[r]
#define CBM4_CELL_HESS(h) fatode::c_lock_hess[h]
#define CBM4_CELL_ST4(p, a, b, c, d) fatode::lock_st4(p, a, b, c, d)
#include "gen/cbm4_cell_gen.h"

namespace fatode {

constexpr int LK_WARPS = 2;                 // systems per CTA (one per warp); 104 KB of shared memory each
constexpr int LK_NV = 9;                    // per-lane vectors: Y_tlm | Ynew_tlm (two slots, swapped on accept), K_tlm_1..6, Fcn0_tlm
constexpr int LK_YA = 0, LK_YB = 1, LK_K = 2, LK_F0 = 8;

template <class Model>
struct LockSmem {
  static constexpr int OFF_MODEL = 0;
  static constexpr int OFF_JV = (Model::WS_DOUBLES + 1) / 2 * 2;        // J(T,Y)            [276]
  static constexpr int OFF_DJ = OFF_JV + 276;                           // J(T+Delta,Y) -> dJ/dt  [276]
  static constexpr int OFF_J1 = OFF_DJ + 276;                           // J(Tau, Ynew)      [276]
  static constexpr int OFF_DJT = OFF_J1 + 276;                          // dJ/dt on the NJT time-dependent entries [32]
  static constexpr int OFF_A = OFF_DJT + 32;                            // E / L\U tile [32][LDA]
  static constexpr int OFF_LU = OFF_A + 32 * LDA;                       // L\U in pivoted row order, column-major [32][32]
  static constexpr int OFF_WHO = OFF_LU + 32 * 32;                      // 32 ints
  static constexpr int OFF_K = OFF_WHO + 16;                            // K_1..K_6 of the state [6][32]
  static constexpr int OFF_PV = OFF_K + 6 * 32;                         // per-lane vectors [LK_NV][32][32 lanes]
  static constexpr int DOUBLES = OFF_PV + LK_NV * 32 * 32;
};

struct LockArgs {
  long ncells; const int* ids; int ntlm; int trunc;
  double* y; double* y_tlm; const double* atol; const double* rtol; const double* atol_tlm; const double* rtol_tlm;
  int* istatus; double* rstatus; int* ierr; unsigned long long* queue;
};

// DGETRS('N') for one right-hand side held in registers, L\U in pivoted row order (column-major, warp-uniform reads): the two
// column-sweep DTRSMs of netlib.  A zero b_k is not skipped (netlib's skip only saves adding exact zeros).
__device__ __forceinline__ void lock_solve_n(const double* __restrict__ a, double (&x)[32]) {
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
__global__ void __launch_bounds__(LK_WARPS * 32, 1)
k_ros_tlm_lock(RosParams rp, typename Model::Const mc, LockArgs a) {
  extern __shared__ __align__(16) double smem[];
  typedef LockSmem<Model> L;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* const sm = smem + (size_t)warp * L::DOUBLES;
  double* const ws = sm + L::OFF_MODEL;
  double* const jv = sm + L::OFF_JV;
  double* const dj = sm + L::OFF_DJ;
  double* const j1 = sm + L::OFF_J1;
  double* const djt = sm + L::OFF_DJT;
  double* const A = sm + L::OFF_A;
  double* const LU = sm + L::OFF_LU;
  int* const who = reinterpret_cast<int*>(sm + L::OFF_WHO);
  double* const Ks = sm + L::OFF_K;
  double* const pv = sm + L::OFF_PV + lane;          // element i of the lane's vector v: pv[(v * 32 + i) * 32]
  const int S = rp.S;
  const int ntlm = a.ntlm;
  const bool active = lane < ntlm;
  const double Roundoff = rp.roundoff;
  const int ti = rp.vector_tol ? lane : 0;
  const double atol = a.atol[ti], rtol = a.rtol[ti];

  for (;;) {
    unsigned long long slot = 0;
    if (lane == 0) slot = atomicAdd(a.queue, 1ull);
    slot = __shfl_sync(FATODE_FULL_MASK, slot, 0);
    if (slot >= (unsigned long long)a.ncells) break;
    const long cell = a.ids ? (long)a.ids[slot] : (long)slot;
    double y = a.y[cell * 32 + lane];
    double* const ycol = a.y_tlm + ((size_t)cell * ntlm + (active ? lane : 0)) * 32;
    int ya = LK_YA, yb = LK_YB;                         // slot of Y_tlm / of the candidate Ynew_tlm
#pragma unroll 4
    for (int i = 0; i < 32; ++i) pv[(ya * 32 + i) * 32] = active ? ycol[i] : 0.0;
    int nfun = 0, njac = 0, nstp = 0, nacc = 0, nrej = 0, ndec = 0, nsol = 0, nsng = 0;
    double texit = 0.0, hexit = 0.0, hnew_out = 0.0;
    double Ynew = y, Fcn0, Fcn = 0.0, dFdT = 0.0;
    int pos = lane;
    LUMasks mk = {0u, 0u};
    Model::init(ws, mc, lane);
    double T = rp.tstart;
    const double Tend = rp.tend;
    double H = fmin(fmax(fabs(rp.hmin), fabs(rp.hstart)), fabs(rp.hmax));
    if (fabs(H) <= __dmul_rn(10.0, Roundoff)) H = rp.deltamin_start;
    const int Direction = (Tend >= rp.tstart) ? 1 : -1;
    const double dDir = (double)Direction;
    H = __dmul_rn(dDir, H);
    bool RejectLastH = false, RejectMoreH = false;
    int ierr = 1;
    for (int s = 0; s < 6; ++s) Ks[s * 32 + lane] = 0.0;
    __syncwarp();

    while ((Direction > 0 && __dadd_rn(__dsub_rn(T, Tend), Roundoff) <= 0.0) ||
           (Direction < 0 && __dadd_rn(__dsub_rn(Tend, T), Roundoff) <= 0.0)) {
      if (nstp > rp.max_no_steps) { ierr = -6; break; }
      if (__dadd_rn(T, __dmul_rn(rp.tenth, H)) == T || H <= Roundoff) { ierr = -7; break; }
      H = fmin(H, fabs(__dsub_rn(Tend, T)));
      // ---- step set-up (:539-562): Fcn0 = f(T,Y), J(T,Y), Fcn0_tlm = J Y_tlm, dF/dT and dJ/dt by forward differences
      Model::set_time(ws, T, lane);
      Model::set_y(ws, y, lane);
      Fcn0 = Model::fun(ws, mc, lane);
      nfun++;
      Model::jac(ws, jv, lane);
      njac++;
      __syncwarp();
      {
        double x[32], o[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) x[i] = pv[(ya * 32 + i) * 32];
        cbm4_cell::j_vec_seq(jv, x, o);
#pragma unroll
        for (int i = 0; i < 32; ++i) pv[(LK_F0 * 32 + i) * 32] = o[i];
      }
      if (!rp.autonomous) {
        const double Delta = __dmul_rn(sqrt(Roundoff), fmax(rp.deltamin_fd, fabs(T)));       // ros_FunTimeDerivative :557
        Model::set_time(ws, __dadd_rn(T, Delta), lane);
        dFdT = Model::fun(ws, mc, lane);
        nfun++;
        dFdT = __dadd_rn(dFdT, __dmul_rn(-1.0, Fcn0));
        dFdT = __dmul_rn(__ddiv_rn(1.0, Delta), dFdT);
        const double Delta2 = __dmul_rn(sqrt(Roundoff), fmax(rp.deltamin_start, fabs(T)));   // DeltaMin = 1e-5 (:500, :558)
        if (Delta2 != Delta) Model::set_time(ws, __dadd_rn(T, Delta2), lane);
        Model::jac(ws, dj, lane);                                                              // LSS_Jac2 :559
        njac++;
        __syncwarp();
        const double dinv = __ddiv_rn(1.0, Delta2);
        for (int e = lane; e < Model::NJ; e += 32)                                             // LSS_Jac_Time_Derivative :561
          dj[e] = __dmul_rn(dinv, __dadd_rn(dj[e], __dmul_rn(-1.0, jv[e])));
        __syncwarp();
        if (lane < Model::NJT) djt[lane] = dj[Model::jac_time_entry(lane)];
        __syncwarp();
      }
      bool accepted = false;
      while (!accepted) {   // UntilAccepted
        {   // ros_PrepareMatrix (:822-858)
          int Nconsecutive = 0;
          bool Singular = true;
          while (Singular) {
            const double ghinv = __ddiv_rn(1.0, __dmul_rn(__dmul_rn(dDir, H), rp.Gamma[0]));
            Model::template scatter_E<LDA>(jv, ghinv, A, lane);
            const int info = warp_lu_factor(A, who, pos, lane, &mk);
            ndec++;
            if (info == 0) {
              Singular = false;
            } else {
              nsng++;
              Nconsecutive++;
              if (Nconsecutive <= 5) H = __dmul_rn(H, 0.5); else break;
            }
          }
          if (Singular) { ierr = -8; break; }
        }
        __syncwarp();
        {   // L\U in pivoted row order for the lanes' own triangular sweeps (row r of the copy = physical row who[r])
          const int src = who[lane];
#pragma unroll 8
          for (int k = 0; k < 32; ++k) LU[lane + 32 * k] = A[k * LDA + src];
        }
        __syncwarp();
        const double dH = __dmul_rn(dDir, H);
#pragma unroll 1
        for (int is = 0; is < S; ++is) {          // istage = is + 1
          const bool newf = (is > 0) && rp.NewF[is];
          // ---- state: stage argument, function and (new for the TLM) the stage Jacobian
          if (is == 0) {
            Fcn = Fcn0;
          } else if (newf) {
            Ynew = y;
            for (int j = 0; j < is; ++j) {
              const double aij = rp.A[is * (is - 1) / 2 + j];
              if (aij != 0.0) Ynew = __dadd_rn(Ynew, __dmul_rn(aij, Ks[j * 32 + lane]));
            }
            const double Tau = __dadd_rn(T, __dmul_rn(__dmul_rn(rp.Alpha[is], dDir), H));
            Model::set_time(ws, Tau, lane);
            Model::set_y(ws, Ynew, lane);
            Fcn = Model::fun(ws, mc, lane);
            nfun++;
            Model::jac(ws, j1, lane);             // LSS_Jac1 :600
            njac++;
            __syncwarp();
          }
          double k = Fcn;
          for (int j = 0; j < is; ++j) {
            const double HC = __ddiv_rn(rp.C[is * (is - 1) / 2 + j], dH);
            if (HC != 0.0) k = __dadd_rn(k, __dmul_rn(HC, Ks[j * 32 + lane]));
          }
          const bool timeterm = !rp.autonomous && rp.Gamma[is] != 0.0;
          const double HG = __dmul_rn(dH, rp.Gamma[is]);
          if (timeterm) k = __dadd_rn(k, __dmul_rn(HG, dFdT));
          k = warp_lu_solve(A, who, pos, mk, k, lane);
          nsol++;
          Ks[is * 32 + lane] = k;
          __syncwarp();
          // ---- columns (lane = column): K_tlm_i = E^-1 [ Fcn_tlm + sum HC_j K_tlm_j + HG dJ/dt Ynew_tlm + Hess_Vec(K_i, Y_tlm) ]
          {
            double x[32], o[32];
            double* const kt = pv + (size_t)(LK_K + is) * 32 * 32;
#pragma unroll
            for (int i = 0; i < 32; ++i) x[i] = pv[(ya * 32 + i) * 32];          // Ynew_tlm = Y_tlm (:586)
            if (is == 0) {
#pragma unroll
              for (int i = 0; i < 32; ++i) o[i] = pv[(LK_F0 * 32 + i) * 32];     // Fcn_tlm = Fcn0_tlm
#pragma unroll
              for (int i = 0; i < 32; ++i) pv[(yb * 32 + i) * 32] = o[i];
            } else if (newf) {
#pragma unroll 1
              for (int j = 0; j < is; ++j) {
                const double aij = rp.A[is * (is - 1) / 2 + j];
                const double* kj = pv + (size_t)(LK_K + j) * 32 * 32;
#pragma unroll
                for (int i = 0; i < 32; ++i) x[i] = x[i] + aij * kj[i * 32];
              }
              cbm4_cell::j_vec_seq(j1, x, o);                                     // Fcn_tlm = J(Tau, Ynew) Ynew_tlm (:602-604)
#pragma unroll
              for (int i = 0; i < 32; ++i) pv[(yb * 32 + i) * 32] = o[i];        // kept for stages without a new function
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) o[i] = pv[(yb * 32 + i) * 32];        // Fcn_tlm of the last new-function stage
            }
#pragma unroll 1
            for (int j = 0; j < is; ++j) {
              const double HC = __ddiv_rn(rp.C[is * (is - 1) / 2 + j], dH);
              const double* kj = pv + (size_t)(LK_K + j) * 32 * 32;
#pragma unroll
              for (int i = 0; i < 32; ++i) o[i] = o[i] + HC * kj[i * 32];
            }
            if (timeterm) cbm4_cell::dj_vec_seq(djt, HG, x, o);                  // + HG * (dJ/dt Ynew_tlm)  (:617-624)
#pragma unroll
            for (int i = 0; i < 32; ++i) kt[i * 32] = o[i];
#pragma unroll
            for (int i = 0; i < 32; ++i) x[i] = pv[(ya * 32 + i) * 32];
            cbm4_cell::hess_vec_seq(Ks + is * 32, x, o);                          // Hess_Vec(T, Y, K_i, Y_tlm) (:628)
#pragma unroll
            for (int i = 0; i < 32; ++i) kt[i * 32] = kt[i * 32] + o[i];         // DAXPY with ONE
#pragma unroll
            for (int r = 0; r < 32; ++r) x[r] = kt[who[r] * 32];                  // DLASWP
            lock_solve_n(LU, x);
#pragma unroll
            for (int i = 0; i < 32; ++i) kt[i * 32] = x[i];
            nsol += ntlm;
          }
          __syncwarp();
        }
        // ---- new solution and error estimates
        Ynew = y;
        for (int j = 0; j < S; ++j)
          if (rp.M[j] != 0.0) Ynew = __dadd_rn(Ynew, __dmul_rn(rp.M[j], Ks[j * 32 + lane]));
        double Yerr = 0.0;
        for (int j = 0; j < S; ++j)
          if (rp.E[j] != 0.0) Yerr = __dadd_rn(Yerr, __dmul_rn(rp.E[j], Ks[j * 32 + lane]));
        double Err;
        {
          const double Ymax = fmax(fabs(y), fabs(Ynew));
          const double Scale = __dadd_rn(atol, __dmul_rn(rtol, Ymax));
          const double q = __ddiv_rn(Yerr, Scale);
          const double q2 = __dmul_rn(q, q);
          double acc = 0.0;
#pragma unroll
          for (int i = 0; i < 32; ++i) acc = __dadd_rn(acc, shfl_d(q2, i));
          Err = fmax(sqrt(__ddiv_rn(acc, 32.0)), 1.0e-10);
        }
        {   // the lane's column: Ynew_tlm into the candidate slot, its error norm when the columns are controlled
          double acc = 0.0;
          const double* at = a.atol_tlm + (size_t)(active ? lane : 0) * 32;
          const double* rt = a.rtol_tlm + (size_t)(active ? lane : 0) * 32;
#pragma unroll 2
          for (int i = 0; i < 32; ++i) {
            const double y0 = pv[(ya * 32 + i) * 32];
            double yn = y0, ye = 0.0;
            for (int j = 0; j < S; ++j) {
              const double kj = pv[((LK_K + j) * 32 + i) * 32];
              yn = yn + rp.M[j] * kj;
              ye = ye + rp.E[j] * kj;
            }
            pv[(yb * 32 + i) * 32] = yn;
            if (a.trunc) {
              const int t = rp.vector_tol ? i : 0;
              const double Ymax = fmax(fabs(y0), fabs(yn));
              const double Scale = at[t] + rt[t] * Ymax;
              const double q = ye / Scale;
              acc = acc + q * q;
            }
          }
          if (a.trunc) {
            double e = active ? fmax(sqrt(acc / 32.0), 1.0e-10) : 0.0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) e = fmax(e, shfl_d(e, lane ^ o));
            Err = fmax(fmax(Err, e), 1.0e-10);
          }
        }
        const double Fac = fmin(rp.facmax, fmax(rp.facmin, __ddiv_rn(rp.facsafe, fpm::root_elo(Err, rp.ELO))));
        double Hnew = __dmul_rn(H, Fac);
        nstp++;
        if (Err <= 1.0 || H <= rp.hmin) {
          nacc++;
          y = Ynew;
          { const int t = ya; ya = yb; yb = t; }     // Y_tlm = Ynew_tlm
          T = __dadd_rn(T, dH);
          Hnew = fmax(rp.hmin, fmin(Hnew, rp.hmax));
          if (RejectLastH) Hnew = fmin(Hnew, H);
          hexit = H; hnew_out = Hnew; texit = T;
          RejectLastH = false; RejectMoreH = false;
          H = Hnew;
          accepted = true;
        } else {
          if (RejectMoreH) Hnew = __dmul_rn(H, rp.facrej);
          RejectMoreH = RejectLastH;
          RejectLastH = true;
          H = Hnew;
          if (nacc >= 1) nrej++;
        }
        __syncwarp();
      }
      if (ierr != 1) break;
    }
    // ---- results (a failed system returns Y and Y_tlm as of its last accepted step, like the reference)
    a.y[cell * 32 + lane] = y;
    if (active) {
#pragma unroll 4
      for (int i = 0; i < 32; ++i) ycol[i] = pv[(ya * 32 + i) * 32];
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

__global__ void k_lock_hess(double fix, double* __restrict__ out) {
  double y[32], ph[cbm4_cell::NPHOTO];
  for (int i = 0; i < 32; ++i) y[i] = 0.0;
  for (int i = 0; i < cbm4_cell::NPHOTO; ++i) ph[i] = 0.0;
  cbm4_cell::hess_values<1>(y, fix, ph, out);
}

int ros_tlm_lock_device(const RosParams& rp, const Cbm4Const& mc, long ncells, const int* d_ids, int ntlm, bool trunc,
                        double* d_y, double* d_ytlm, const double* d_atol, const double* d_rtol, const double* d_atol_tlm,
                        const double* d_rtol_tlm, int* d_istatus, double* d_rstatus, int* d_ierr, cudaStream_t st,
                        UnitTiming& tm, std::string& err) {
  if (ncells <= 0) return 0;
  typedef Cbm4Warp Model;
  int dev = 0, sms = 0;
  UNIT_TRY(cudaGetDevice(&dev));
  UNIT_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const size_t smem = sizeof(double) * (size_t)LockSmem<Model>::DOUBLES * LK_WARPS;
  UNIT_TRY(cudaFuncSetAttribute(k_ros_tlm_lock<Model>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  double* d_scratch = nullptr;                  // [0..257] Hessian staging, then the work queue
  UNIT_TRY(cudaMalloc(&d_scratch, sizeof(double) * 264));
  struct Free { double* p; ~Free() { cudaFree(p); } } guard{d_scratch};
  UNIT_TRY(cudaMemcpyToSymbolAsync(c_lock_rc, mc.rc_base, sizeof(double) * 81, 0, cudaMemcpyHostToDevice, st));
  k_lock_hess<<<1, 1, 0, st>>>(mc.fix, d_scratch);
  UNIT_TRY(cudaGetLastError());
  UNIT_TRY(cudaMemcpyToSymbolAsync(c_lock_hess, d_scratch, sizeof(double) * 258, 0, cudaMemcpyDeviceToDevice, st));
  unsigned long long* d_queue = reinterpret_cast<unsigned long long*>(d_scratch + 260);
  UNIT_TRY(cudaMemsetAsync(d_queue, 0, sizeof(unsigned long long), st));
  LockArgs a{ncells, d_ids, ntlm, trunc ? 1 : 0, d_y, d_ytlm, d_atol, d_rtol, trunc ? d_atol_tlm : d_atol, trunc ? d_rtol_tlm : d_rtol,
             d_istatus, d_rstatus, d_ierr, d_queue};
  cudaEvent_t e0, e1;
  UNIT_TRY(cudaEventCreate(&e0));
  UNIT_TRY(cudaEventCreate(&e1));
  UNIT_TRY(cudaEventRecord(e0, st));
  const int blocks = (int)std::min<long>((long)sms, (ncells + LK_WARPS - 1) / LK_WARPS);
  k_ros_tlm_lock<Model><<<blocks, LK_WARPS * 32, smem, st>>>(rp, mc, a);
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
  tm.launches += 2;
  return 0;
}

}  // namespace fatode
