// Forward adaptive Rosenbrock sweep, one cell per warp, lane i = component i (N = 32 models).
//
// Replaces, for a batch of cells, the reference's per-system step loop
//   ros_Integrator  FWD/ROS/ROS_f90_Integrator.F90:433-629        (family FAM_FWD)
//   ros_FwdInt      ADJ/ROS_ADJ/ROS_ADJ_f90_Integrator.F90:877-1173 (family FAM_ADJ, tape push :735-755)
// including ros_FunTimeDerivative (:1870-1890), ros_PrepareMatrix (:1894-1948), the
// LSS_Decomp/LSS_Solve FULL_ALGEBRA path (LS_Solver.F90:31-54 -> DGETRF/DGETRS 'N') and
// ros_ErrorNorm (:1837-1866).
//
// Linear algebra: lane i keeps row i of E = 1/(h*gamma) I - J in 32 FP64 registers.  Partial
// pivoting is implicit (rows never move; each lane tracks the logical position `pos` of its
// row), the pivot search is a warp arg-max with LAPACK's IDAMAX tie rule (first maximal), the
// pivot row travels by shuffles.  Every element sees the same operations in the same order as
// netlib's unblocked DGETF2 (reciprocal scaling, a(i,j) += l_i * (-u_j)) and column-sweep DTRSM,
// written with unfused multiplies and adds, so the factors and solutions are bit-identical to
// the reference's FULL_ALGEBRA arithmetic.
#pragma once
#include <float.h>
#include "fatode_common.cuh"

namespace fatode {

struct WarpLU {
  double a[32];   // row `lane` of L\U (unit lower part = multipliers)
  int pos;        // logical row index of this lane's row after pivoting
  int who;        // lane k holds who[k] = physical lane that owns logical row k
  int info;       // 0, or first k+1 with an exactly zero pivot (DGETF2 INFO)
};

// warp arg-max of |v| over lanes with active==true; ties -> smallest key.  returns winning lane.
__device__ __forceinline__ int warp_argmax_abs(double v, int key, bool active, int lane) {
  // order by (|v| as ordered 64-bit pattern, then 31-key)
  unsigned long long bits = active ? ((unsigned long long)__double_as_longlong(fabs(v))) : 0ull;
  unsigned hi = (unsigned)(bits >> 32), lo = (unsigned)bits;
  unsigned mhi = __reduce_max_sync(FATODE_FULL_MASK, active ? hi : 0u);
  bool c1 = active && hi == mhi;
  unsigned mlo = __reduce_max_sync(FATODE_FULL_MASK, c1 ? lo : 0u);
  bool c2 = c1 && lo == mlo;
  unsigned mk = __reduce_min_sync(FATODE_FULL_MASK, c2 ? (unsigned)key : 0xffffffffu);
  bool win = c2 && (unsigned)key == mk;
  unsigned b = __ballot_sync(FATODE_FULL_MASK, win);
  return __ffs(b) - 1;
}

// DGETF2 on the register-resident matrix (netlib LAPACK 3.x unblocked LU, see file header)
__device__ __forceinline__ void warp_lu_factor(WarpLU& lu, int lane) {
  const double sfmin = DBL_MIN;   // DLAMCH('S') for IEEE double (1/huge < tiny)
  lu.pos = lane;
  lu.who = lane;
  lu.info = 0;
#pragma unroll
  for (int k = 0; k < 32; ++k) {
    bool cand = lu.pos >= k;
    int p = warp_argmax_abs(lu.a[k], lu.pos, cand, lane);     // physical lane of the pivot row
    int pos_p = __shfl_sync(FATODE_FULL_MASK, lu.pos, p);
    double piv = shfl_d(lu.a[k], p);
    // swap logical positions k <-> pos_p  (DSWAP of rows; data stays in place)
    if (lu.pos == k) lu.pos = pos_p;
    if (lane == p) lu.pos = k;
    if (piv != 0.0) {
      if (k < 31) {
        bool below = lu.pos > k;
        if (fabs(piv) >= sfmin) {
          double r = __ddiv_rn(1.0, piv);
          if (below) lu.a[k] = __dmul_rn(r, lu.a[k]);
        } else {
          if (below) lu.a[k] = __ddiv_rn(lu.a[k], piv);
        }
      }
    } else if (lu.info == 0) {
      lu.info = k + 1;
    }
    if (k < 31) {
      bool below = lu.pos > k;
#pragma unroll
      for (int j = k + 1; j < 32; ++j) {
        double u = shfl_d(lu.a[j], p);
        if (below && u != 0.0) lu.a[j] = __dadd_rn(lu.a[j], __dmul_rn(lu.a[k], -u));
      }
    }
  }
  // directory: who[k] = lane whose pos == k
  {
    int w = 0;
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      unsigned b = __ballot_sync(FATODE_FULL_MASK, lu.pos == k);
      if (lane == k) w = __ffs(b) - 1;
    }
    lu.who = w;
  }
}

// DGETRS 'N' with one right-hand side; lane i passes b_i and receives x_i
__device__ __forceinline__ double warp_lu_solve(const WarpLU& lu, double b, int lane) {
  // row interchanges are implicit: lane's b belongs to logical row lu.pos
#pragma unroll
  for (int k = 0; k < 31; ++k) {                 // DTRSM Left Lower NoTrans Unit
    int src = __shfl_sync(FATODE_FULL_MASK, lu.who, k);
    double bk = shfl_d(b, src);
    if (lu.pos > k && bk != 0.0) b = __dsub_rn(b, __dmul_rn(bk, lu.a[k]));
  }
#pragma unroll
  for (int k = 31; k >= 0; --k) {                // DTRSM Left Upper NoTrans NonUnit
    int src = __shfl_sync(FATODE_FULL_MASK, lu.who, k);
    double d = shfl_d(lu.a[k], src);
    double bk = shfl_d(b, src);
    if (bk != 0.0) {
      bk = __ddiv_rn(bk, d);
      if (lane == src) b = bk;
      if (lu.pos < k) b = __dsub_rn(b, __dmul_rn(bk, lu.a[k]));
    }
  }
  // unknown r sits in the lane that owns logical row r
  return shfl_d(b, lu.who);
}

struct FwdResult {
  int ierr;
  int istatus[8];
  double texit, hexit, hnew;
};

// One cell.  `y` is lane's component (in/out).  TAPE: entries are appended at tape_base (per cell),
// up to tape_cap entries; exceeding the capacity ends the cell with IERR = -6.
template <class Model, bool TAPE>
__device__ void ros_forward_warp(const RosParams& rp, const typename Model::Const& mc, double* ws, double* tile,
                                 double& y, double atol, double rtol, double* tape_base, long tape_cap,
                                 FwdResult& res, int lane) {
  const int S = rp.S;
  const bool adj = (rp.family == FAM_ADJ);
  const bool count_la = (rp.family != FAM_FWD);
  const double Roundoff = rp.roundoff;
  const int NE = tape_entry_doubles(S, 32);
  int nfun = 0, njac = 0, nstp = 0, nacc = 0, nrej = 0, ndec = 0, nsol = 0, nsng = 0;
  double K[6], Ystage[6];
  double Ynew = y, Fcn0, Fcn = 0.0, dFdT = 0.0;
  WarpLU lu;
  double* jv = ws + Model::WS_DOUBLES;   // [NJ] Jacobian values

  Model::init(ws, mc, lane);
  double T = rp.tstart;
  const double Tend = rp.tend;
  res.hexit = 0.0; res.texit = 0.0; res.hnew = 0.0;
  double H = fmin(fmax(fabs(rp.hmin), fabs(rp.hstart)), fabs(rp.hmax));
  if (fabs(H) <= __dmul_rn(10.0, Roundoff)) H = rp.deltamin_start;
  const int Direction = (Tend >= rp.tstart) ? 1 : -1;
  const double dDir = (double)Direction;
  H = __dmul_rn(dDir, H);
  bool RejectLastH = false, RejectMoreH = false;
  int ierr = 1;
#pragma unroll
  for (int s = 0; s < 6; ++s) { K[s] = 0.0; Ystage[s] = 0.0; }

  while ((Direction > 0 && __dadd_rn(__dsub_rn(T, Tend), Roundoff) <= 0.0) ||
         (Direction < 0 && __dadd_rn(__dsub_rn(Tend, T), Roundoff) <= 0.0)) {
    if (nstp > rp.max_no_steps) { ierr = -6; break; }
    if (__dadd_rn(T, __dmul_rn(rp.tenth, H)) == T || H <= Roundoff) { ierr = -7; break; }
    if (adj) res.hexit = H;
    H = fmin(H, fabs(__dsub_rn(Tend, T)));
    // Fcn0 = f(T, Y)
    Model::set_time(ws, T, lane);
    Model::set_y(ws, y, lane);
    Fcn0 = Model::fun(ws, mc, lane);
    nfun++;
    // J(T, Y) (same operands)
    Model::jac(ws, jv, lane);
    njac++;
    if (!rp.autonomous) {   // ros_FunTimeDerivative
      double Delta = __dmul_rn(sqrt(Roundoff), fmax(rp.deltamin_fd, fabs(T)));
      Model::set_time(ws, __dadd_rn(T, Delta), lane);
      dFdT = Model::fun(ws, mc, lane);
      nfun++;
      dFdT = __dadd_rn(dFdT, __dmul_rn(-1.0, Fcn0));
      dFdT = __dmul_rn(__ddiv_rn(1.0, Delta), dFdT);
    }
    bool accepted = false;
    while (!accepted) {   // UntilAccepted
      // ros_PrepareMatrix
      {
        int Nconsecutive = 0;
        bool Singular = true;
        while (Singular) {
          double ghinv = __ddiv_rn(1.0, __dmul_rn(__dmul_rn(dDir, H), rp.Gamma[0]));
          Model::scatter_E(jv, ghinv, tile, lane);
#pragma unroll
          for (int j = 0; j < 32; ++j) lu.a[j] = tile[j * 32 + lane];
          __syncwarp();
          warp_lu_factor(lu, lane);
          if (count_la) ndec++;
          if (lu.info == 0) {
            Singular = false;
          } else {
            nsng++;
            Nconsecutive++;
            if (Nconsecutive <= 5) H = __dmul_rn(H, 0.5); else break;
          }
        }
        if (Singular) { ierr = -8; break; }
      }
#pragma unroll
      for (int is = 0; is < 6; ++is) {          // istage = is + 1; fully unrolled so K/Ystage stay in registers
        if (is < S) {
          if (is == 0) {
            Fcn = Fcn0;
            if (adj) { Ystage[0] = y; Ynew = y; }
          } else if (rp.NewF[is]) {
            Ynew = y;
#pragma unroll
            for (int j = 0; j < is; ++j) {
              double a = rp.A[is * (is - 1) / 2 + j];
              if (a != 0.0) Ynew = __dadd_rn(Ynew, __dmul_rn(a, K[j]));     // DAXPY (returns when da == 0)
            }
            double Tau = __dadd_rn(T, __dmul_rn(__dmul_rn(rp.Alpha[is], dDir), H));
            Model::set_time(ws, Tau, lane);
            Model::set_y(ws, Ynew, lane);
            Fcn = Model::fun(ws, mc, lane);
            nfun++;
          }
          if (is > 0 && adj) Ystage[is] = Ynew;
          double k = Fcn;
#pragma unroll
          for (int j = 0; j < is; ++j) {
            double HC = __ddiv_rn(rp.C[is * (is - 1) / 2 + j], __dmul_rn(dDir, H));
            if (HC != 0.0) k = __dadd_rn(k, __dmul_rn(HC, K[j]));
          }
          if (!rp.autonomous && rp.Gamma[is] != 0.0) {
            double HG = __dmul_rn(__dmul_rn(dDir, H), rp.Gamma[is]);
            k = __dadd_rn(k, __dmul_rn(HG, dFdT));
          }
          k = warp_lu_solve(lu, k, lane);
          if (count_la) nsol++;
          K[is] = k;
        }
      }
      Ynew = y;
#pragma unroll
      for (int j = 0; j < 6; ++j)
        if (j < S && rp.M[j] != 0.0) Ynew = __dadd_rn(Ynew, __dmul_rn(rp.M[j], K[j]));
      double Yerr = 0.0;
#pragma unroll
      for (int j = 0; j < 6; ++j)
        if (j < S && rp.E[j] != 0.0) Yerr = __dadd_rn(Yerr, __dmul_rn(rp.E[j], K[j]));
      // ros_ErrorNorm: sequential sum over i = 1..N in the reference -> lane 0 accumulates in order
      double Err;
      {
        double Ymax = fmax(fabs(y), fabs(Ynew));
        double Scale = __dadd_rn(atol, __dmul_rn(rtol, Ymax));
        double q = __ddiv_rn(Yerr, Scale);
        double q2 = __dmul_rn(q, q);
        double acc = 0.0;
#pragma unroll
        for (int i = 0; i < 32; ++i) acc = __dadd_rn(acc, shfl_d(q2, i));
        Err = fmax(sqrt(__ddiv_rn(acc, 32.0)), 1.0e-10);
      }
      double Fac = fmin(rp.facmax, fmax(rp.facmin, __ddiv_rn(rp.facsafe, pow(Err, __ddiv_rn(1.0, rp.ELO)))));
      double Hnew = __dmul_rn(H, Fac);
      nstp++;
      if (Err <= 1.0 || H <= rp.hmin) {
        nacc++;
        if (TAPE) {
          if (nacc > tape_cap) { ierr = -6; break; }
          double* e = tape_base + (long)(nacc - 1) * NE;
          if (lane == 0) { e[0] = T; e[1] = H; }
#pragma unroll
          for (int s = 0; s < 6; ++s)
            if (s < S) {
              e[4 + s * 32 + lane] = Ystage[s];
              e[4 + (S + s) * 32 + lane] = K[s];
            }
        }
        y = Ynew;
        T = __dadd_rn(T, __dmul_rn(dDir, H));
        Hnew = fmax(rp.hmin, fmin(Hnew, rp.hmax));
        if (RejectLastH) Hnew = fmin(Hnew, H);
        res.hexit = H; res.hnew = Hnew; res.texit = T;
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
    }
    if (ierr != 1) break;
  }
  res.ierr = ierr;
  res.istatus[Nfun] = nfun; res.istatus[Njac] = njac; res.istatus[Nstp] = nstp; res.istatus[Nacc] = nacc;
  res.istatus[Nrej] = nrej; res.istatus[Ndec] = ndec; res.istatus[Nsol] = nsol; res.istatus[Nsng] = nsng;
}

}  // namespace fatode
