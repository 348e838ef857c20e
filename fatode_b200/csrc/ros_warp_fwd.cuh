// Forward adaptive Rosenbrock sweep, one cell per warp, lane i = component i (N = 32 models).
//
// Replaces, for a batch of cells, the reference's per-system step loop
//   ros_Integrator  FWD/ROS/ROS_f90_Integrator.F90:433-629        (family FAM_FWD)
//   ros_FwdInt      ADJ/ROS_ADJ/ROS_ADJ_f90_Integrator.F90:877-1173 (family FAM_ADJ, tape push :735-755)
// including ros_FunTimeDerivative (:1870-1890), ros_PrepareMatrix (:1894-1948), the
// LSS_Decomp/LSS_Solve FULL_ALGEBRA path (LS_Solver.F90:31-54 -> DGETRF/DGETRS 'N') and
// ros_ErrorNorm (:1837-1866).
//
// Linear algebra: E = 1/(h*gamma) I - J lives in the warp's shared-memory slice, column-major with a
// padded leading dimension (33) so that both "lane = row" and "lane = column" accesses are
// bank-conflict free.  Partial pivoting is implicit: rows never move, each lane tracks the logical
// position `pos` of its row and who[k] names the physical row sitting at position k.  The pivot
// search is a warp arg-max with LAPACK's IDAMAX tie rule (first maximal).  The rank-1 update visits
// only the non-zero entries of the pivot row (a warp ballot builds the mask) — netlib's DGER skips
// those columns too — so the sparsity of the chemical Jacobian is exploited without changing a bit:
// every element sees the same operations in the same order as the unblocked DGETF2 (reciprocal
// scaling, a(i,j) += l_i * (-u_j)) and the column-sweep DTRSMs of DGETRS, written with unfused
// multiplies and adds.  The factors and solutions are bit-identical to the FULL_ALGEBRA arithmetic
// (tests/test_units_gpu.py checks this against the oracle's netlib restatement).
#pragma once
#include <float.h>
#include "fatode_common.cuh"
#include "portable_math.h"

namespace fatode {

constexpr int LDA = 33;   // padded leading dimension of the 32x32 tile

// warp arg-max of |v| over lanes with active==true; ties -> smallest key.  returns winning lane.
__device__ __forceinline__ int warp_argmax_abs(double v, int key, bool active) {
  unsigned long long bits = active ? ((unsigned long long)__double_as_longlong(fabs(v))) : 0ull;
  unsigned hi = (unsigned)(bits >> 32), lo = (unsigned)bits;
  unsigned mhi = __reduce_max_sync(FATODE_FULL_MASK, active ? hi : 0u);
  bool c1 = active && hi == mhi;
  unsigned mlo = __reduce_max_sync(FATODE_FULL_MASK, c1 ? lo : 0u);
  bool c2 = c1 && lo == mlo;
  unsigned mk = __reduce_min_sync(FATODE_FULL_MASK, c2 ? (unsigned)key : 0xffffffffu);
  bool win = c2 && (unsigned)key == mk;
  return __ffs(__ballot_sync(FATODE_FULL_MASK, win)) - 1;
}

// Column-occupancy masks of the factors: bit k of lnz / unz is set when column k of L (below the diagonal) /
// of U (above the diagonal) holds a non-zero.  The triangular sweeps skip empty columns — netlib's DTRSM
// would add exact zeros there — and for CBM-IV about half of U's columns are empty.
struct LUMasks { unsigned lnz, unz; };

// DGETF2 on the shared-memory tile A (column-major, ld = LDA, lane = physical row).
// Outputs: pos (per lane), who[0..31] in shared memory; returns INFO (0 or first zero pivot, 1-based).
__device__ __forceinline__ int warp_lu_factor(double* __restrict__ A, int* __restrict__ who, int& pos, int lane,
                                              LUMasks* masks = nullptr) {
  const double sfmin = DBL_MIN;   // DLAMCH('S') for IEEE double (1/huge < tiny)
  pos = lane;
  int info = 0;
  unsigned lnz = 0, unz = 0;
  for (int k = 0; k < 32; ++k) {
    const double v = A[k * LDA + lane];
    if (__ballot_sync(FATODE_FULL_MASK, pos < k && v != 0.0)) unz |= 1u << k;   // column k of U is final here
    const int p = warp_argmax_abs(v, pos, pos >= k);             // physical row of the pivot
    const int pos_p = __shfl_sync(FATODE_FULL_MASK, pos, p);
    const double piv = shfl_d(v, p);
    if (pos == k) pos = pos_p;                                   // DSWAP of rows k <-> pos_p (logical only)
    if (lane == p) pos = k;
    if (lane == 0) who[k] = p;
    if (piv != 0.0) {
      if (k < 31) {
        const bool below = pos > k;
        double l;
        if (fabs(piv) >= sfmin) l = __dmul_rn(__ddiv_rn(1.0, piv), v);   // DSCAL by the reciprocal
        else l = __ddiv_rn(v, piv);
        if (below) A[k * LDA + lane] = l;
        if (__ballot_sync(FATODE_FULL_MASK, below && l != 0.0)) lnz |= 1u << k;
        const double urow = A[lane * LDA + p];                   // lane = column: pivot row entry
        unsigned mask = __ballot_sync(FATODE_FULL_MASK, urow != 0.0 && lane > k);
        while (mask) {                                           // DGER over the non-zero columns
          const int j = __ffs(mask) - 1;
          mask &= mask - 1;
          const double u = A[j * LDA + p];
          if (below) {
            double* a = A + j * LDA + lane;
            *a = __dadd_rn(*a, __dmul_rn(l, -u));
          }
        }
      }
    } else if (info == 0) {
      info = k + 1;
    }
    __syncwarp();
  }
  if (masks) { masks->lnz = lnz; masks->unz = unz; }
  return info;
}

// DGETRS 'N' with one right-hand side; lane i passes b_i and receives x_i
__device__ __forceinline__ double warp_lu_solve(const double* __restrict__ A, const int* __restrict__ who, int pos,
                                                LUMasks mk, double b, int lane) {
  // row interchanges are implicit: lane's b belongs to logical row pos
  for (unsigned m = mk.lnz; m; m &= m - 1) {     // DTRSM Left Lower NoTrans Unit, non-empty columns ascending
    const int k = __ffs(m) - 1;
    const double bk = shfl_d(b, who[k]);
    if (bk != 0.0) {
      const double l = A[k * LDA + lane];
      if (pos > k) b = __dsub_rn(b, __dmul_rn(bk, l));
    }
  }
  for (unsigned m = mk.unz; m;) {                // DTRSM Left Upper NoTrans NonUnit, non-empty columns descending
    const int k = 31 - __clz(m);
    m &= ~(1u << k);
    const int src = who[k];
    double bk = shfl_d(b, src);
    if (bk != 0.0) {
      bk = __ddiv_rn(bk, A[k * LDA + src]);
      const double u = A[k * LDA + lane];
      if (lane == src) b = bk;
      if (pos < k) b = __dsub_rn(b, __dmul_rn(bk, u));
    }
  }
  // rows whose column of U is empty above the diagonal: nobody consumes x_k, divide them all at once
  if (!((mk.unz >> pos) & 1u) && b != 0.0) b = __ddiv_rn(b, A[pos * LDA + lane]);
  // unknown r sits in the lane that owns logical row r
  return shfl_d(b, who[lane]);
}

struct FwdResult {
  int ierr;
  int nfun, njac, nstp, nacc, nrej, ndec, nsol, nsng;
  double texit, hexit, hnew;
};

// Per-warp shared-memory slice of the forward sweep (doubles)
template <class Model>
struct FwdSmem {
  static constexpr int OFF_MODEL = 0;
  static constexpr int OFF_JV = Model::WS_DOUBLES;                       // [NJ] J(T,Y) values
  static constexpr int OFF_A = OFF_JV + ((Model::NJ + 1) / 2) * 2;       // [32][LDA] E / L\U
  static constexpr int OFF_WHO = OFF_A + 32 * LDA;                       // 32 ints
  static constexpr int OFF_K = OFF_WHO + 16;                             // [6][32]
  static constexpr int OFF_YS = OFF_K + 6 * 32;                          // [6][32] stage states (tape)
  static constexpr int DOUBLES = OFF_YS + 6 * 32;
};

// One cell.  `y` is lane's component (in/out).  TAPE: entries are appended at tape_base (per cell),
// up to tape_cap entries; exceeding the capacity ends the cell with IERR = -6.
template <class Model, bool TAPE>
__device__ void ros_forward_warp(const RosParams& rp, const typename Model::Const& mc, double* sm, double& y,
                                 double atol, double rtol, double* tape_base, long tape_cap, FwdResult& res, int lane) {
  typedef FwdSmem<Model> L;
  const int S = rp.S;
  const bool adj = (rp.family == FAM_ADJ);
  const bool count_la = (rp.family != FAM_FWD);
  const double Roundoff = rp.roundoff;
  const int NE = tape_entry_doubles(S, 32);
  double* ws = sm + L::OFF_MODEL;
  double* jv = sm + L::OFF_JV;
  double* A = sm + L::OFF_A;
  int* who = reinterpret_cast<int*>(sm + L::OFF_WHO);
  double* Ks = sm + L::OFF_K;
  double* Ys = sm + L::OFF_YS;
  int nfun = 0, njac = 0, nstp = 0, nacc = 0, nrej = 0, ndec = 0, nsol = 0, nsng = 0;
  double Ynew = y, Fcn0, Fcn = 0.0, dFdT = 0.0;
  int pos = lane;
  LUMasks mk = {0u, 0u};

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
  for (int s = 0; s < 6; ++s) { Ks[s * 32 + lane] = 0.0; Ys[s * 32 + lane] = 0.0; }
  __syncwarp();

  while ((Direction > 0 && __dadd_rn(__dsub_rn(T, Tend), Roundoff) <= 0.0) ||
         (Direction < 0 && __dadd_rn(__dsub_rn(Tend, T), Roundoff) <= 0.0)) {
    if (nstp > rp.max_no_steps) { ierr = -6; break; }
    if (__dadd_rn(T, __dmul_rn(rp.tenth, H)) == T || H <= Roundoff) { ierr = -7; break; }
    if (adj) res.hexit = H;
    H = fmin(H, fabs(__dsub_rn(Tend, T)));
    // Fcn0 = f(T, Y);  J(T, Y) (same operands)
    Model::set_time(ws, T, lane);
    Model::set_y(ws, y, lane);
    Fcn0 = Model::fun(ws, mc, lane);
    nfun++;
    Model::jac(ws, jv, lane);
    njac++;
    if (!rp.autonomous) {   // ros_FunTimeDerivative
      const double Delta = __dmul_rn(sqrt(Roundoff), fmax(rp.deltamin_fd, fabs(T)));
      Model::set_time(ws, __dadd_rn(T, Delta), lane);
      dFdT = Model::fun(ws, mc, lane);
      nfun++;
      dFdT = __dadd_rn(dFdT, __dmul_rn(-1.0, Fcn0));
      dFdT = __dmul_rn(__ddiv_rn(1.0, Delta), dFdT);
    }
    bool accepted = false;
    while (!accepted) {   // UntilAccepted
      {   // ros_PrepareMatrix
        int Nconsecutive = 0;
        bool Singular = true;
        while (Singular) {
          const double ghinv = __ddiv_rn(1.0, __dmul_rn(__dmul_rn(dDir, H), rp.Gamma[0]));
          Model::template scatter_E<LDA>(jv, ghinv, A, lane);
          const int info = warp_lu_factor(A, who, pos, lane, &mk);
          if (count_la) ndec++;
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
      const double dH = __dmul_rn(dDir, H);
      for (int is = 0; is < S; ++is) {          // istage = is + 1
        if (is == 0) {
          Fcn = Fcn0;
          if (adj) Ynew = y;
        } else if (rp.NewF[is]) {
          Ynew = y;
          for (int j = 0; j < is; ++j) {
            const double a = rp.A[is * (is - 1) / 2 + j];
            if (a != 0.0) Ynew = __dadd_rn(Ynew, __dmul_rn(a, Ks[j * 32 + lane]));     // DAXPY (returns when da == 0)
          }
          const double Tau = __dadd_rn(T, __dmul_rn(__dmul_rn(rp.Alpha[is], dDir), H));
          Model::set_time(ws, Tau, lane);
          Model::set_y(ws, Ynew, lane);
          Fcn = Model::fun(ws, mc, lane);
          nfun++;
        }
        if (adj) Ys[is * 32 + lane] = Ynew;     // stage state, saved even if Ynew was not refreshed (:1015-1018)
        double k = Fcn;
        for (int j = 0; j < is; ++j) {
          const double HC = __ddiv_rn(rp.C[is * (is - 1) / 2 + j], dH);
          if (HC != 0.0) k = __dadd_rn(k, __dmul_rn(HC, Ks[j * 32 + lane]));
        }
        if (!rp.autonomous && rp.Gamma[is] != 0.0) {
          const double HG = __dmul_rn(dH, rp.Gamma[is]);
          k = __dadd_rn(k, __dmul_rn(HG, dFdT));
        }
        k = warp_lu_solve(A, who, pos, mk, k, lane);
        if (count_la) nsol++;
        Ks[is * 32 + lane] = k;
      }
      Ynew = y;
      for (int j = 0; j < S; ++j)
        if (rp.M[j] != 0.0) Ynew = __dadd_rn(Ynew, __dmul_rn(rp.M[j], Ks[j * 32 + lane]));
      double Yerr = 0.0;
      for (int j = 0; j < S; ++j)
        if (rp.E[j] != 0.0) Yerr = __dadd_rn(Yerr, __dmul_rn(rp.E[j], Ks[j * 32 + lane]));
      // ros_ErrorNorm: the reference sums i = 1..N sequentially -> same order here
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
      const double Fac = fmin(rp.facmax, fmax(rp.facmin, __ddiv_rn(rp.facsafe, fpm::root_elo(Err, rp.ELO))));
      double Hnew = __dmul_rn(H, Fac);
      nstp++;
      if (Err <= 1.0 || H <= rp.hmin) {
        nacc++;
        if (TAPE) {
          if (nacc > tape_cap) { ierr = -6; break; }
          double* e = tape_base + (long)(nacc - 1) * NE;
          if (lane == 0) { e[0] = T; e[1] = H; }
          for (int s = 0; s < S; ++s) {
            e[4 + s * 32 + lane] = Ys[s * 32 + lane];
            e[4 + (S + s) * 32 + lane] = Ks[s * 32 + lane];
          }
        }
        y = Ynew;
        T = __dadd_rn(T, dH);
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
  res.nfun = nfun; res.njac = njac; res.nstp = nstp; res.nacc = nacc;
  res.nrej = nrej; res.ndec = ndec; res.nsol = nsol; res.nsng = nsng;
}

}  // namespace fatode
