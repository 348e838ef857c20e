// Rosenbrock discrete adjoint for the SMALL example models (small_strato N = 5, roberts N = 3): INTEGRATE_ADJ of ADJ/ROS_ADJ
// with everything of one system in ONE THREAD.
//
// Replaces, for these models, RosenbrockADJ1 / RosenbrockADJ2 (ADJ/ROS_ADJ/ROS_ADJ_f90_Integrator.F90:159, 2485): ros_FwdInt
// (:877-1173) with the trajectory tape (ros_DPush :735-755: T, H, Ystage(N*S), K(N*S) per accepted step), ADJINIT, ros_DadjInt
// (:1177-1441 with Mu, Lambda-only twin :3508-3700).  The reference example: EXAMPLES/small_strato/small_ros_adj
// (small_strato_dr.F90:73-76: NADJ = 5, Lambda only, golden rkadj_ADJ_results.m) and EXAMPLES/roberts/ROBERTS_ROS_ADJ (NP = 3).
//
// With N <= 5 a system's whole working set (state, stage vectors, dense LU, U and V of all stages and columns) is a few hundred
// doubles: one thread integrates its system forward, pushes the tape to its slab in HBM, and walks it backwards itself.
// Linear algebra is netlib DGETF2 / DGETRS('N'|'T') unrolled per thread (full partial pivoting, the singular-matrix retry of
// ros_PrepareMatrix included); every operation is the reference's, in the reference's order, without FMA contraction, so the
// forward state and ISTATUS are bit-identical to the FWD kernels and Lambda / Mu are bit-identical to the CPU restatement.
// Two launches per wave of systems: a counting pass (forward sweep without tape -> accepted steps per system), an exclusive
// scan on the host, then the taped forward + backward pass with exact tape offsets.  No quadrature terms (GetQuad).
#include <cuda_runtime.h>
#include <algorithm>
#include <cfloat>
#include <string>
#include <vector>
#include "../../include/fatode_b200.h"
#include "units_internal.h"
#include "portable_math.h"

namespace fatode {
namespace {

#define F4(x) ((double)(x##f))

__device__ __forceinline__ double small_sun(double T) { return Cbm4Warp::sun(T); }   // Update_SUN: the same routine in all KPP examples

// ---- small_strato: EXAMPLES/small_strato/small_ros_adj/User_Parameters.F90 (FUN_U :10-37, JAC_U :39-85, Update_RCONST
// :114-127, HessTR_Vec_U :136-190, adjinit :192-204); REAL(4) literals keep their single-precision rounding
struct SmallStratoAdj {
  static constexpr int N = 5, NP = 0;
  double fix0, fix1;
  double rc[10];
  __device__ void init(const double* fix) { fix0 = fix[0]; fix1 = fix[1]; update(0.0); }
  __device__ void update(double t) {
    const double SUN = small_sun(t);
    rc[0] = ((F4(2.643E-10)) * SUN * SUN * SUN);
    rc[1] = F4(8.018E-17);
    rc[2] = ((F4(6.120E-04)) * SUN);
    rc[3] = F4(1.576E-15);
    rc[4] = ((F4(1.070E-03)) * SUN * SUN);
    rc[5] = F4(7.110E-11);
    rc[6] = F4(1.200E-10);
    rc[7] = F4(6.062E-15);
    rc[8] = F4(1.069E-11);
    rc[9] = ((F4(1.289E-02)) * SUN);
  }
  __device__ void fun(double t, const double* V, double* FCT) {
    double A[10];
    update(t);
    A[0] = rc[0] * fix1;
    A[1] = rc[1] * V[1] * fix1;
    A[2] = rc[2] * V[2];
    A[3] = rc[3] * V[1] * V[2];
    A[4] = rc[4] * V[2];
    A[5] = rc[5] * V[0] * fix0;
    A[6] = rc[6] * V[0] * V[2];
    A[7] = rc[7] * V[2] * V[3];
    A[8] = rc[8] * V[1] * V[4];
    A[9] = rc[9] * V[4];
    FCT[0] = A[4] - A[5] - A[6];
    FCT[1] = 2 * A[0] - A[1] + A[2] - A[3] + A[5] - A[8] + A[9];
    FCT[2] = A[1] - A[2] - A[3] - A[4] - A[6] - A[7];
    FCT[3] = -A[7] + A[8] + A[9];
    FCT[4] = A[7] - A[8] - A[9];
  }
  __device__ void jac(double t, const double* V, double* JF) {     // column-major 5 x 5, zeroed by the callback (:50)
    double B[17];
    update(t);
    for (int i = 0; i < 25; ++i) JF[i] = 0.0;
#define JFX(i, j) JF[(i - 1) + 5 * (j - 1)]
    B[2] = rc[1] * fix1;
    B[4] = rc[2];
    B[5] = rc[3] * V[2];
    B[6] = rc[3] * V[1];
    B[7] = rc[4];
    B[8] = rc[5] * fix0;
    B[10] = rc[6] * V[2];
    B[11] = rc[6] * V[0];
    B[12] = rc[7] * V[3];
    B[13] = rc[7] * V[2];
    B[14] = rc[8] * V[4];
    B[15] = rc[8] * V[1];
    B[16] = rc[9];
    JFX(1, 1) = -B[8] - B[10];
    JFX(1, 3) = B[7] - B[11];
    JFX(2, 1) = B[8];
    JFX(2, 2) = -B[2] - B[5] - B[14];
    JFX(2, 3) = B[4] - B[6];
    JFX(2, 5) = -B[15] + B[16];
    JFX(3, 1) = -B[10];
    JFX(3, 2) = B[2] - B[5];
    JFX(3, 3) = -B[4] - B[6] - B[7] - B[11] - B[12];
    JFX(3, 4) = -B[13];
    JFX(3, 5) = 0;
    JFX(4, 2) = B[14];
    JFX(4, 3) = -B[12];
    JFX(4, 4) = -B[13];
    JFX(4, 5) = B[15] + B[16];
    JFX(5, 2) = -B[14];
    JFX(5, 3) = B[12];
    JFX(5, 4) = B[13];
    JFX(5, 5) = -B[15] - B[16];
#undef JFX
  }
  // HTU = (f_yy x U2)^T U1; only time-independent rate constants enter
  __device__ void hesstr_vec(double, const double*, const double* U1, const double* U2, double* HTU) {
    double D2A[5], HESS[11];
    D2A[1] = rc[3]; D2A[2] = rc[6]; D2A[3] = rc[7]; D2A[4] = rc[8];
    HESS[1] = -D2A[2]; HESS[2] = -D2A[1]; HESS[3] = -D2A[4]; HESS[4] = -D2A[2]; HESS[5] = -D2A[1];
    HESS[6] = -D2A[3]; HESS[7] = D2A[4]; HESS[8] = -D2A[3]; HESS[9] = -D2A[4]; HESS[10] = D2A[3];
#define u1(i) U1[i - 1]
#define u2(i) U2[i - 1]
    HTU[0] = HESS[1] * (u1(1) * u2(3)) + HESS[4] * (u1(3) * u2(3));
    HTU[1] = HESS[2] * (u1(2) * u2(3)) + HESS[3] * (u1(2) * u2(5)) + HESS[5] * (u1(3) * u2(3))
           + HESS[7] * (u1(4) * u2(5)) + HESS[9] * (u1(5) * u2(5));
    HTU[2] = HESS[1] * (u1(1) * u2(1)) + HESS[2] * (u1(2) * u2(2)) + HESS[4] * (u1(3) * u2(1))
           + HESS[5] * (u1(3) * u2(2)) + HESS[6] * (u1(3) * u2(4))
           + HESS[8] * (u1(4) * u2(4)) + HESS[10] * (u1(5) * u2(4));
    HTU[3] = HESS[6] * (u1(3) * u2(3)) + HESS[8] * (u1(4) * u2(3)) + HESS[10] * (u1(5) * u2(3));
    HTU[4] = HESS[3] * (u1(2) * u2(2)) + HESS[7] * (u1(4) * u2(2)) + HESS[9] * (u1(5) * u2(2));
#undef u1
#undef u2
  }
  __device__ void jacp(double, const double*, double*) {}
  __device__ void hesstr_vec_f_py(double, const double*, const double*, const double*, double*) {}
  __device__ void adjinit(int nadj, const double*, double* lambda, double*) {
    for (int i = 0; i < N * nadj; ++i) lambda[i] = 0.0;
    for (int k = 0; k < nadj && k < N; ++k) lambda[k + N * k] = 1.0;
  }
};

// ---- roberts: EXAMPLES/roberts/ROBERTS_ROS_ADJ/roberts_ros_adj_dr.F90 (fun :97-110, jac :75-94, hesstr_vec :59-72, JACP :18-29,
// HESSTR_VEC_F_PY :31-40, adjinit :124-146)
struct RobertsAdj {
  static constexpr int N = 3, NP = 3;
  double c0, c1, c2;
  __device__ void init(const double*) { c0 = F4(0.04e0); c1 = F4(1.0e4); c2 = F4(3.0e7); }
  __device__ void fun(double, const double* y, double* p) {
    p[0] = -c0 * y[0] + c1 * y[1] * y[2];
    p[2] = c2 * y[1] * y[1];
    p[1] = -p[0] - p[2];
  }
  // the callback does NOT write fjac(3,1) and fjac(3,3): they keep the zeros of the fresh allocation
  __device__ void jac(double, const double* y, double* fjac) {
#define FJ(i, j) fjac[(i - 1) + 3 * (j - 1)]
    FJ(1, 1) = -c0;
    FJ(1, 2) = c1 * y[2];
    FJ(1, 3) = c1 * y[1];
    FJ(2, 1) = c0;
    FJ(2, 2) = -c1 * y[2] - 2 * c2 * y[1];
    FJ(2, 3) = -c1 * y[1];
    FJ(3, 2) = 2 * c2 * y[1];
#undef FJ
  }
  __device__ void hesstr_vec(double, const double*, const double* u, const double* k, double* tmp) {
    tmp[0] = 0;
    tmp[1] = 2 * c2 * k[1] * (u[2] - u[1]) + c1 * k[2] * (u[0] - u[1]);
    tmp[2] = c1 * k[1] * (u[0] - u[1]);
  }
  __device__ void jacp(double, const double* y, double* fpjac) {    // writes 6 of 9 entries; the rest stay 0
#define FP(i, j) fpjac[(i - 1) + 3 * (j - 1)]
    FP(1, 1) = -y[0];
    FP(2, 1) = y[0];
    FP(1, 2) = y[1] * y[2];
    FP(2, 2) = -y[1] * y[2];
    FP(2, 3) = -y[1] * y[1];
    FP(3, 3) = y[1] * y[1];
#undef FP
  }
  __device__ void hesstr_vec_f_py(double, const double* y, const double* u, const double* k, double* tmp) {
    tmp[0] = (u[1] - u[0]) * k[0];
    tmp[1] = (u[0] - u[1]) * (y[2] * k[1] + y[1] * k[2]);
    tmp[2] = 2 * y[1] * k[1] * (u[2] - u[1]);
  }
  __device__ void adjinit(int nadj, const double* y, double* lambda, double* mu) {
    for (int i = 0; i < N * nadj; ++i) lambda[i] = 0.0;
    for (int k = 0; k < nadj && k < N; ++k) lambda[k + N * k] = y[k];
    if (mu) {
      for (int i = 0; i < NP * nadj; ++i) mu[i] = 0.0;
#define MU(i, j) if (j <= nadj) mu[(i - 1) + 3 * (j - 1)]
      MU(1, 1) = -y[0] * y[0];
      MU(1, 2) = y[0] * y[1] * y[2];
      MU(2, 1) = y[0] * y[1];
      MU(2, 2) = -y[1] * y[1] * y[2];
      MU(2, 3) = -y[1] * y[1] * y[1];
      MU(3, 3) = y[1] * y[1] * y[2];
#undef MU
    }
  }
};

// ---- netlib DGETF2 / DGETRS on a thread-local column-major N x N matrix (ADJ/ROS_ADJ/LS_Solver.F90:31-54)
template <int N>
__device__ int dgetf2_t(double* a, int* ipiv) {
  int info = 0;
  for (int j = 0; j < N; ++j) {
    int jp = j;
    double dmax = fabs(a[j + N * j]);
    for (int i = j + 1; i < N; ++i) {
      const double v = fabs(a[i + N * j]);
      if (v > dmax) { dmax = v; jp = i; }
    }
    ipiv[j] = jp;
    if (a[jp + N * j] != 0.0) {
      if (jp != j)
        for (int c = 0; c < N; ++c) { const double t = a[j + N * c]; a[j + N * c] = a[jp + N * c]; a[jp + N * c] = t; }
      if (j < N - 1) {
        if (fabs(a[j + N * j]) >= DBL_MIN) {
          const double r = 1.0 / a[j + N * j];
          for (int i = j + 1; i < N; ++i) a[i + N * j] = r * a[i + N * j];
        } else {
          for (int i = j + 1; i < N; ++i) a[i + N * j] = a[i + N * j] / a[j + N * j];
        }
      }
    } else if (info == 0) {
      info = j + 1;
    }
    if (j < N - 1)
      for (int c = j + 1; c < N; ++c) {           // DGER: skips a zero y(c)
        const double yc = a[j + N * c];
        if (yc != 0.0) {
          const double t = -1.0 * yc;
          for (int i = j + 1; i < N; ++i) a[i + N * c] = a[i + N * c] + a[i + N * j] * t;
        }
      }
  }
  return info;
}
template <int N>
__device__ void dgetrs_t(bool trans, const double* a, const int* ipiv, double* b) {
  if (!trans) {
    for (int i = 0; i < N; ++i) { const int ip = ipiv[i]; if (ip != i) { const double t = b[i]; b[i] = b[ip]; b[ip] = t; } }
    for (int k = 0; k < N; ++k)
      if (b[k] != 0.0)
        for (int i = k + 1; i < N; ++i) b[i] = b[i] - b[k] * a[i + N * k];
    for (int k = N - 1; k >= 0; --k)
      if (b[k] != 0.0) {
        b[k] = b[k] / a[k + N * k];
        for (int i = 0; i < k; ++i) b[i] = b[i] - b[k] * a[i + N * k];
      }
  } else {
    for (int i = 0; i < N; ++i) {
      double temp = b[i];
      for (int k = 0; k < i; ++k) temp = temp - a[k + N * i] * b[k];
      temp = temp / a[i + N * i];
      b[i] = temp;
    }
    for (int i = N - 1; i >= 0; --i) {
      double temp = b[i];
      for (int k = i + 1; k < N; ++k) temp = temp - a[k + N * i] * b[k];
      b[i] = temp;
    }
    for (int i = N - 1; i >= 0; --i) { const int ip = ipiv[i]; if (ip != i) { const double t = b[i]; b[i] = b[ip]; b[ip] = t; } }
  }
}

constexpr int SA_MAXADJ = 8;    // adjoint columns per call for the small models (the examples use NADJ = N)

struct SmallAdjArgs {
  long ncells; int nadj; int with_mu; int tape_pass;
  double* y; double* lambda; double* mu; const double* atol; const double* rtol;
  int* istatus; double* rstatus; int* ierr; int* nacc; const long long* tape_off; double* tape;
  double fix[2];
};

// MODE: the counting pass integrates forward only (outputs: nacc, ierr); the tape pass repeats the forward sweep with the
// tape, then runs ADJINIT and the backward sweep.
template <class Model>
__global__ void __launch_bounds__(64)
k_ros_small_adj(RosParams rp, SmallAdjArgs a) {
  constexpr int N = Model::N, NP = Model::NP;
  const long cell = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= a.ncells) return;
  if (a.tape_pass && a.ierr[cell] != 1) return;      // a failed forward sweep returns before ADJINIT (:585-600): Y untouched below
  Model mdl;
  mdl.init(a.fix);
  const int S = rp.S;
  const int ES = 2 + 2 * N * S;                       // tape entry: T, H, Ystage(N*S), K(N*S)
  double* tape = a.tape_pass ? a.tape + (size_t)a.tape_off[cell] * ES : nullptr;
  double Y[N], Ynew[N], Fcn0[N], Fcn[N], dFdT[N], Yerr[N], K[N * 6], Ystage[N * 6];
  double fjac[N * N], E[N * N];
  int ip[N];
  for (int i = 0; i < N * N; ++i) fjac[i] = 0.0;      // fresh ALLOCATE -> zero pages (entries a callback never writes stay 0)
  for (int i = 0; i < N * 6; ++i) Ystage[i] = 0.0;
  for (int i = 0; i < N; ++i) Y[i] = a.y[cell * N + i];
  int nfun = 0, njac = 0, nstp = 0, nacc = 0, nrej = 0, ndec = 0, nsol = 0, nsng = 0;
  double texit = 0.0, hexit = 0.0, hnew_out = 0.0;
  const double Roundoff = rp.roundoff, Tstart = rp.tstart, Tend = rp.tend;
  const int Direction = (Tend >= Tstart) ? 1 : -1;
  double T = Tstart;
  double H = fmin(fmax(fabs(rp.hmin), fabs(rp.hstart)), fabs(rp.hmax));
  if (fabs(H) <= 10.0 * Roundoff) H = rp.deltamin_start;
  H = Direction * H;
  bool RejectLastH = false, RejectMoreH = false;
  int ierr = 1;
  // ---------------------------------------------------------------- ros_FwdInt (:877-1173)
  while ((Direction > 0 && (T - Tend) + Roundoff <= 0.0) || (Direction < 0 && (Tend - T) + Roundoff <= 0.0)) {
    if (nstp > rp.max_no_steps) { ierr = -6; break; }
    if ((T + rp.tenth * H) == T || H <= Roundoff) { ierr = -7; break; }
    hexit = H;
    H = fmin(H, fabs(Tend - T));
    mdl.fun(T, Y, Fcn0);
    nfun++;
    if (!rp.autonomous) {
      const double Delta = sqrt(Roundoff) * fmax(rp.deltamin_fd, fabs(T));
      mdl.fun(T + Delta, Y, dFdT);
      nfun++;
      for (int i = 0; i < N; ++i) dFdT[i] = dFdT[i] + (-1.0) * Fcn0[i];
      const double s = 1.0 / Delta;
      for (int i = 0; i < N; ++i) dFdT[i] = s * dFdT[i];
    }
    mdl.jac(T, Y, fjac);
    njac++;
    bool accepted = false;
    while (!accepted) {
      {   // ros_PrepareMatrix (:1894-1948)
        int Nconsecutive = 0;
        bool Singular = true;
        while (Singular) {
          const double ghinv = 1.0 / (Direction * H * rp.Gamma[0]);
          for (int j = 0; j < N; ++j) {
            for (int i = 0; i < N; ++i) E[i + N * j] = -fjac[i + N * j];
            E[j + N * j] = E[j + N * j] + ghinv;
          }
          const int ising = dgetf2_t<N>(E, ip);
          ndec++;
          if (ising == 0) Singular = false;
          else {
            nsng++;
            Nconsecutive++;
            if (Nconsecutive <= 5) H = H * 0.5; else break;
          }
        }
        if (Singular) { ierr = -8; break; }
      }
      for (int istage = 1; istage <= S; ++istage) {
        const int ioffset = N * (istage - 1);
        if (istage == 1) {
          for (int i = 0; i < N; ++i) { Fcn[i] = Fcn0[i]; Ystage[i] = Y[i]; Ynew[i] = Y[i]; }
        } else if (rp.NewF[istage - 1]) {
          for (int i = 0; i < N; ++i) Ynew[i] = Y[i];
          for (int j = 1; j <= istage - 1; ++j) {
            const double aa = rp.A[(istage - 1) * (istage - 2) / 2 + j - 1];
            for (int i = 0; i < N; ++i) Ynew[i] = Ynew[i] + aa * K[N * (j - 1) + i];
          }
          const double Tau = T + rp.Alpha[istage - 1] * Direction * H;
          mdl.fun(Tau, Ynew, Fcn);
          nfun++;
        }
        if (istage > 1) for (int i = 0; i < N; ++i) Ystage[ioffset + i] = Ynew[i];
        for (int i = 0; i < N; ++i) K[ioffset + i] = Fcn[i];
        for (int j = 1; j <= istage - 1; ++j) {
          const double HC = rp.C[(istage - 1) * (istage - 2) / 2 + j - 1] / (Direction * H);
          for (int i = 0; i < N; ++i) K[ioffset + i] = K[ioffset + i] + HC * K[N * (j - 1) + i];
        }
        if (!rp.autonomous && rp.Gamma[istage - 1] != 0.0) {
          const double HG = Direction * H * rp.Gamma[istage - 1];
          for (int i = 0; i < N; ++i) K[ioffset + i] = K[ioffset + i] + HG * dFdT[i];
        }
        dgetrs_t<N>(false, E, ip, &K[ioffset]);
        nsol++;
      }
      for (int i = 0; i < N; ++i) Ynew[i] = Y[i];
      for (int j = 1; j <= S; ++j) {
        const double m = rp.M[j - 1];
        for (int i = 0; i < N; ++i) Ynew[i] = Ynew[i] + m * K[N * (j - 1) + i];
      }
      for (int i = 0; i < N; ++i) Yerr[i] = 0.0;
      for (int j = 1; j <= S; ++j) {
        const double e = rp.E[j - 1];
        for (int i = 0; i < N; ++i) Yerr[i] = Yerr[i] + e * K[N * (j - 1) + i];
      }
      double Err = 0.0;
      for (int i = 0; i < N; ++i) {
        const double Ymax = fmax(fabs(Y[i]), fabs(Ynew[i]));
        const int ti = rp.vector_tol ? i : 0;
        const double Scale = a.atol[ti] + a.rtol[ti] * Ymax;
        const double q = Yerr[i] / Scale;
        Err = Err + q * q;
      }
      Err = fmax(sqrt(Err / N), 1.0e-10);
      const double Fac = fmin(rp.facmax, fmax(rp.facmin, rp.facsafe / fpm::root_elo(Err, rp.ELO)));
      double Hnew = H * Fac;
      nstp++;
      if (Err <= 1.0 || H <= rp.hmin) {
        nacc++;
        if (tape) {                                // ros_DPush
          double* e = tape + (size_t)(nacc - 1) * ES;
          e[0] = T; e[1] = H;
          for (int i = 0; i < N * S; ++i) { e[2 + i] = Ystage[i]; e[2 + N * S + i] = K[i]; }
        }
        for (int i = 0; i < N; ++i) Y[i] = Ynew[i];
        T = T + Direction * H;
        Hnew = fmax(rp.hmin, fmin(Hnew, rp.hmax));
        if (RejectLastH) Hnew = fmin(Hnew, H);
        hexit = H; hnew_out = Hnew; texit = T;
        RejectLastH = false; RejectMoreH = false;
        H = Hnew;
        accepted = true;
      } else {
        if (RejectMoreH) Hnew = H * rp.facrej;
        RejectMoreH = RejectLastH;
        RejectLastH = true;
        H = Hnew;
        if (nacc >= 1) nrej++;
      }
    }
    if (ierr != 1) break;
  }
  if (!a.tape_pass) {   // counting pass: how much tape the system needs, and whether its forward sweep succeeds
    a.nacc[cell] = nacc;
    a.ierr[cell] = ierr;
    if (ierr != 1) {    // the reference returns before ADJINIT: Y as of the last accepted step, Lambda / Mu untouched
      for (int i = 0; i < N; ++i) a.y[cell * N + i] = Y[i];
      int* is_ = a.istatus + cell * 20;
      for (int q = 0; q < 20; ++q) is_[q] = 0;
      is_[Nfun] = nfun; is_[Njac] = njac; is_[Nstp] = nstp; is_[Nacc] = nacc; is_[Nrej] = nrej; is_[Ndec] = ndec; is_[Nsol] = nsol; is_[Nsng] = nsng;
      double* rs = a.rstatus + cell * 20;
      for (int q = 0; q < 20; ++q) rs[q] = 0.0;
      rs[Ntexit] = texit; rs[Nhexit] = hexit; rs[Nhnew] = hnew_out;
    }
    return;
  }
  // ---------------------------------------------------------------- ADJINIT + ros_DadjInt (:1177-1441)
  const int NADJ = a.nadj;
  const bool with_mu = a.with_mu != 0 && NP > 0;
  double Lambda[N * SA_MAXADJ], Mu[(NP > 0 ? NP : 1) * SA_MAXADJ];
  double U[N * 6 * SA_MAXADJ], V[N * 6 * SA_MAXADJ];
  double djdt[N * N], FPJAC0[N * (NP > 0 ? NP : 1)], FPJAC1[N * (NP > 0 ? NP : 1)], Tmp[N], Tmp2[N], Tmp3[NP > 0 ? NP : 1];
  for (int i = 0; i < N * N; ++i) djdt[i] = 0.0;
  for (int i = 0; i < N * (NP > 0 ? NP : 1); ++i) { FPJAC0[i] = 0.0; FPJAC1[i] = 0.0; }
  mdl.adjinit(NADJ, Y, Lambda, with_mu ? Mu : nullptr);
  const int LD = N * S;
  const double DeltaMin = 1.0e-5;
  for (int st = nacc - 1; st >= 0; --st) {             // ros_DPop
    const double* e = tape + (size_t)st * ES;
    const double Tt = e[0], Ht = e[1];
    const double* Yst = e + 2;
    const double* Kt = e + 2 + N * S;
    nstp++;
    double Y0[N];
    for (int i = 0; i < N; ++i) Y0[i] = Yst[i];
    mdl.jac(Tt, Y0, fjac);
    njac++;
    double Tau = 1.0 / (Direction * Ht * rp.Gamma[0]);
    njac++;                                            // sic (:1264)
    for (int j = 0; j < N; ++j) {
      for (int i = 0; i < N; ++i) E[i + N * j] = -fjac[i + N * j];
      E[j + N * j] = E[j + N * j] + Tau;
    }
    dgetf2_t<N>(E, ip);
    ndec++;
    for (int istage = S; istage >= 1; --istage) {
      const int istart = N * (istage - 1);
      for (int m = 0; m < NADJ; ++m) {
        double* u = &U[istart + LD * m];
        for (int i = 0; i < N; ++i) u[i] = Lambda[i + N * m];
        for (int i = 0; i < N; ++i) u[i] = rp.M[istage - 1] * u[i];
      }
      for (int j = istage + 1; j <= S; ++j) {
        const int jstart = N * (j - 1);
        const double HA = rp.A[(j - 1) * (j - 2) / 2 + istage - 1];
        const double HC = rp.C[(j - 1) * (j - 2) / 2 + istage - 1] / (Direction * Ht);
        for (int m = 0; m < NADJ; ++m) {
          double* u = &U[istart + LD * m];
          const double* vj = &V[jstart + LD * m];
          const double* uj = &U[jstart + LD * m];
          for (int i = 0; i < N; ++i) u[i] = u[i] + HA * vj[i];
          for (int i = 0; i < N; ++i) u[i] = u[i] + HC * uj[i];
        }
      }
      for (int m = 0; m < NADJ; ++m) {
        dgetrs_t<N>(true, E, ip, &U[istart + LD * m]);
        nsol++;
      }
      Tau = Tt + rp.Alpha[istage - 1] * Direction * Ht;
      double Yi[N];
      for (int i = 0; i < N; ++i) Yi[i] = Yst[istart + i];
      mdl.jac(Tau, Yi, fjac);
      njac++;
      for (int m = 0; m < NADJ; ++m) {                  // V = matmul(transpose(fjac), U)
        const double* g = &U[istart + LD * m];
        double* z = &V[istart + LD * m];
        for (int i = 0; i < N; ++i) { double s = 0.0; for (int k = 0; k < N; ++k) s += fjac[k + N * i] * g[k]; z[i] = s; }
      }
    }
    for (int istage = 1; istage <= S; ++istage) {       // Lambda (:1324-1339)
      const int istart = N * (istage - 1);
      double Ki[N];
      for (int i = 0; i < N; ++i) Ki[i] = Kt[istart + i];
      for (int m = 0; m < NADJ; ++m) {
        double* lam = &Lambda[N * m];
        const double* v = &V[istart + LD * m];
        for (int i = 0; i < N; ++i) lam[i] = lam[i] + 1.0 * v[i];
        mdl.hesstr_vec(Tt, Y0, &U[istart + LD * m], Ki, Tmp);
        for (int i = 0; i < N; ++i) lam[i] = lam[i] + 1.0 * Tmp[i];
      }
    }
    if (with_mu) {                                      // Mu (:1342-1372)
      for (int istage = 1; istage <= S; ++istage) {
        const int istart = N * (istage - 1);
        Tau = Tt + rp.Alpha[istage - 1] * Direction * Ht;
        double Yi[N], Ki[N];
        for (int i = 0; i < N; ++i) { Yi[i] = Yst[istart + i]; Ki[i] = Kt[istart + i]; }
        mdl.jacp(Tau, Yi, FPJAC1);
        for (int m = 0; m < NADJ; ++m) {
          const double* u = &U[istart + LD * m];
          double* mu = &Mu[NP * m];
          for (int p = 0; p < NP; ++p) Tmp3[p] = 0.0;
          for (int p = 0; p < NP; ++p) {
            double temp = 0.0;
            for (int i = 0; i < N; ++i) temp = temp + FPJAC1[i + N * p] * u[i];
            Tmp3[p] = Tmp3[p] + 1.0 * temp;
          }
          for (int p = 0; p < NP; ++p) mu[p] = mu[p] + 1.0 * Tmp3[p];
          mdl.hesstr_vec_f_py(Tt, Y0, u, Ki, Tmp3);
          for (int p = 0; p < NP; ++p) mu[p] = mu[p] + 1.0 * Tmp3[p];
        }
      }
    }
    if (!rp.autonomous) {                               // (:1377-1406)
      const double Delta = sqrt(Roundoff) * fmax(DeltaMin, fabs(Tt));
      mdl.jac(Tt + Delta, Y0, djdt);
      njac++;
      {   // LSS_Jac_Time_Derivative: relies on fjac == J(T_1, Y_1), the last Jacobian evaluated above (istage = 1)
        const double al = -1.0;
        for (int i = 0; i < N * N; ++i) djdt[i] = djdt[i] + al * fjac[i];
        const double sc = 1 / Delta;
        for (int i = 0; i < N * N; ++i) djdt[i] = sc * djdt[i];
      }
      if (with_mu) {
        mdl.jacp(Tt, Y0, FPJAC0);
        mdl.jacp(Tt + Delta, Y0, FPJAC1);
        double Alpha = -1.0;
        for (int i = 0; i < N * NP; ++i) FPJAC1[i] = FPJAC1[i] + Alpha * FPJAC0[i];
        Alpha = 1 / Delta;
        for (int i = 0; i < N * NP; ++i) FPJAC1[i] = Alpha * FPJAC1[i];
      }
      for (int m = 0; m < NADJ; ++m) {
        for (int i = 0; i < N; ++i) Tmp[i] = 0.0;
        for (int istage = 1; istage <= S; ++istage) {
          const double* u = &U[N * (istage - 1) + LD * m];
          const double g = rp.Gamma[istage - 1];
          if (g != 0.0)
            for (int i = 0; i < N; ++i) Tmp[i] = Tmp[i] + g * u[i];
        }
        if (with_mu) {
          double* mu = &Mu[NP * m];
          for (int p = 0; p < NP; ++p) {
            double temp = 0.0;
            for (int i = 0; i < N; ++i) temp = temp + FPJAC1[i + N * p] * Tmp[i];
            mu[p] = mu[p] + Ht * temp;
          }
        }
        for (int i = 0; i < N; ++i) { double s = 0.0; for (int k = 0; k < N; ++k) s += djdt[k + N * i] * Tmp[k]; Tmp2[i] = s; }
        double* lam = &Lambda[N * m];
        for (int i = 0; i < N; ++i) lam[i] = lam[i] + Ht * Tmp2[i];
      }
    }
  }
  // ---------------------------------------------------------------- results
  for (int i = 0; i < N; ++i) a.y[cell * N + i] = Y[i];
  for (int m = 0; m < NADJ; ++m)
    for (int i = 0; i < N; ++i) a.lambda[((size_t)cell * NADJ + m) * N + i] = Lambda[i + N * m];
  if (with_mu && a.mu)
    for (int m = 0; m < NADJ; ++m)
      for (int p = 0; p < NP; ++p) a.mu[((size_t)cell * NADJ + m) * NP + p] = Mu[p + NP * m];
  int* is_ = a.istatus + cell * 20;
  for (int q = 0; q < 20; ++q) is_[q] = 0;
  is_[Nfun] = nfun; is_[Njac] = njac; is_[Nstp] = nstp; is_[Nacc] = nacc; is_[Nrej] = nrej; is_[Ndec] = ndec; is_[Nsol] = nsol; is_[Nsng] = nsng;
  double* rs = a.rstatus + cell * 20;
  for (int q = 0; q < 20; ++q) rs[q] = 0.0;
  rs[Ntexit] = texit; rs[Nhexit] = hexit; rs[Nhnew] = hnew_out;
  a.ierr[cell] = 1;
}

template <class Model>
int small_adj_run(const RosParams& rp, const double* fix, long ncells, int nadj, bool with_mu, double* d_y, double* d_lambda,
                  double* d_mu, const double* d_atol, const double* d_rtol, int* d_istatus, double* d_rstatus, int* d_ierr,
                  size_t tape_budget, cudaStream_t st, UnitTiming& tm, std::string& err) {
  constexpr int N = Model::N;
  const int ES = 2 + 2 * N * rp.S;
  int* d_nacc = nullptr;
  long long* d_off = nullptr;
  double* d_tape = nullptr;
  size_t tape_cap = 0;
  struct Guard { int** a; long long** b; double** c; ~Guard() { cudaFree(*a); cudaFree(*b); cudaFree(*c); } } guard{&d_nacc, &d_off, &d_tape};
  UNIT_TRY(cudaMalloc(&d_nacc, sizeof(int) * (size_t)ncells));
  UNIT_TRY(cudaMalloc(&d_off, sizeof(long long) * (size_t)ncells));
  cudaEvent_t e0, e1;
  UNIT_TRY(cudaEventCreate(&e0));
  UNIT_TRY(cudaEventCreate(&e1));
  struct EvGuard { cudaEvent_t a, b; ~EvGuard() { cudaEventDestroy(a); cudaEventDestroy(b); } } evg{e0, e1};
  UNIT_TRY(cudaEventRecord(e0, st));
  SmallAdjArgs a{ncells, nadj, with_mu ? 1 : 0, 0, d_y, d_lambda, d_mu, d_atol, d_rtol, d_istatus, d_rstatus, d_ierr, d_nacc, d_off,
                 nullptr, {fix[0], fix[1]}};
  const unsigned blocks = (unsigned)((ncells + 63) / 64);
  k_ros_small_adj<Model><<<blocks, 64, 0, st>>>(rp, a);      // counting pass over all systems
  UNIT_TRY(cudaGetLastError());
  std::vector<int> h_nacc(ncells), h_ierr(ncells);
  UNIT_TRY(cudaMemcpyAsync(h_nacc.data(), d_nacc, sizeof(int) * (size_t)ncells, cudaMemcpyDeviceToHost, st));
  UNIT_TRY(cudaMemcpyAsync(h_ierr.data(), d_ierr, sizeof(int) * (size_t)ncells, cudaMemcpyDeviceToHost, st));
  UNIT_TRY(cudaStreamSynchronize(st));
  tm.launches += 1;
  // waves of consecutive systems whose tapes fit the budget together; the tape pass runs on a window [c0, c1)
  std::vector<long long> h_off(ncells);
  long c0 = 0;
  while (c0 < ncells) {
    long c1 = c0;
    size_t steps = 0;
    while (c1 < ncells) {
      const size_t need = (h_ierr[c1] == 1) ? (size_t)h_nacc[c1] : 0;
      if (c1 > c0 && (steps + need) * ES * sizeof(double) > tape_budget) break;
      h_off[c1] = (long long)steps;
      steps += need;
      ++c1;
    }
    const size_t bytes = std::max<size_t>(steps, 1) * ES * sizeof(double);
    if (bytes > tape_cap) {
      cudaFree(d_tape);
      d_tape = nullptr;
      cudaError_t ce = cudaMalloc(&d_tape, bytes);
      if (ce != cudaSuccess) { err = "trajectory tape of the small-model adjoint does not fit in device memory"; return -104; }
      tape_cap = bytes;
    }
    UNIT_TRY(cudaMemcpyAsync(d_off + c0, h_off.data() + c0, sizeof(long long) * (size_t)(c1 - c0), cudaMemcpyHostToDevice, st));
    SmallAdjArgs b = a;
    b.tape_pass = 1;
    b.ncells = c1 - c0;
    b.y = d_y + (size_t)c0 * N;
    b.lambda = d_lambda + (size_t)c0 * nadj * N;
    b.mu = d_mu ? d_mu + (size_t)c0 * nadj * Model::NP : nullptr;
    b.istatus = d_istatus + (size_t)c0 * 20;
    b.rstatus = d_rstatus + (size_t)c0 * 20;
    b.ierr = d_ierr + c0;
    b.nacc = d_nacc + c0;
    b.tape_off = d_off + c0;
    b.tape = d_tape;
    k_ros_small_adj<Model><<<(unsigned)((c1 - c0 + 63) / 64), 64, 0, st>>>(rp, b);
    UNIT_TRY(cudaGetLastError());
    UNIT_TRY(cudaStreamSynchronize(st));
    tm.launches += 1;
    c0 = c1;
  }
  UNIT_TRY(cudaEventRecord(e1, st));
  UNIT_TRY(cudaStreamSynchronize(st));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  tm.ms += ms;
  return 0;
}

}  // namespace

// Continuous adjoint (ICNTRL(7) = 3) after a forward-only sweep: what RosenbrockADJ1/2 leave behind.  The forward sweep in this
// mode evaluates FUN, JAC (and the time derivative) once more at TOUT for its last checkpoint (:1143-1157); ADJINIT runs; then
// ros_CadjInt is entered as ros_CadjInt(NADJ, Lambda, Tend, Tstart, ...) (:588-592), i.e. backwards in time with a NEGATIVE H, and
// its first statement inside the time loop, IF (((T+0.1d0*H) == T).OR.(H <= Roundoff)) (:1517), returns IERR = -7 (-6 when the
// forward sweep used up Max_no_steps) before anything is integrated: Lambda = ADJINIT, RSTATUS(Nhexit) = 0.
template <class Model>
__global__ void k_ros_small_cadj_epilogue(long ncells, int nadj, int with_mu, int autonomous, int max_no_steps, const double* y,
                                          double* lambda, double* mu, int* istatus, double* rstatus, int* ierr, double fix0, double fix1) {
  constexpr int N = Model::N, NP = Model::NP;
  const long cell = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= ncells || ierr[cell] != 1) return;
  Model mdl;
  const double fix[2] = {fix0, fix1};
  mdl.init(fix);
  double Y[N], Lambda[N * SA_MAXADJ], Mu[(NP > 0 ? NP : 1) * SA_MAXADJ];
  for (int i = 0; i < N; ++i) Y[i] = y[cell * N + i];
  const bool wm = with_mu != 0 && NP > 0;
  mdl.adjinit(nadj, Y, Lambda, wm ? Mu : nullptr);
  for (int m = 0; m < nadj; ++m)
    for (int i = 0; i < N; ++i) lambda[((size_t)cell * nadj + m) * N + i] = Lambda[i + N * m];
  if (wm && mu)
    for (int m = 0; m < nadj; ++m)
      for (int p = 0; p < NP; ++p) mu[((size_t)cell * nadj + m) * NP + p] = Mu[p + NP * m];
  int* is_ = istatus + cell * 20;
  is_[Nfun] += autonomous ? 1 : 2;
  is_[Njac] += 1;
  rstatus[cell * 20 + Nhexit] = 0.0;
  ierr[cell] = (is_[Nstp] > max_no_steps) ? -6 : -7;
}

int ros_small_cadj_epilogue_device(int model, const RosParams& rp, const double* fix, long ncells, int nadj, bool with_mu, const double* d_y,
                                   double* d_lambda, double* d_mu, int* d_istatus, double* d_rstatus, int* d_ierr, cudaStream_t st) {
  if (nadj < 1 || nadj > SA_MAXADJ) return -105;
  const unsigned blocks = (unsigned)((ncells + 63) / 64);
  if (model == FATODE_MODEL_SMALL_STRATO)
    k_ros_small_cadj_epilogue<SmallStratoAdj><<<blocks, 64, 0, st>>>(ncells, nadj, 0, rp.autonomous, rp.max_no_steps, d_y, d_lambda, d_mu,
                                                                     d_istatus, d_rstatus, d_ierr, fix[0], fix[1]);
  else
    k_ros_small_cadj_epilogue<RobertsAdj><<<blocks, 64, 0, st>>>(ncells, nadj, with_mu ? 1 : 0, rp.autonomous, rp.max_no_steps, d_y, d_lambda,
                                                                 d_mu, d_istatus, d_rstatus, d_ierr, fix[0], fix[1]);
  return cudaGetLastError() == cudaSuccess ? 0 : -103;
}

// model: FATODE_MODEL_SMALL_STRATO or FATODE_MODEL_ROBERTS; device arrays; y is read as y0 and overwritten with y(TOUT)
int ros_small_adj_device(int model, const RosParams& rp, const double* fix, long ncells, int nadj, bool with_mu, double* d_y,
                         double* d_lambda, double* d_mu, const double* d_atol, const double* d_rtol, int* d_istatus,
                         double* d_rstatus, int* d_ierr, size_t tape_budget, cudaStream_t st, UnitTiming& tm, std::string& err) {
  if (nadj < 1 || nadj > SA_MAXADJ) { err = "the small models take 1..8 adjoint columns per call"; return -105; }
  if (tape_budget == 0) {
    size_t fr = 0, tot = 0;
    UNIT_TRY(cudaMemGetInfo(&fr, &tot));
    tape_budget = (size_t)(0.5 * (double)fr);
  }
  if (model == FATODE_MODEL_SMALL_STRATO)
    return small_adj_run<SmallStratoAdj>(rp, fix, ncells, nadj, false, d_y, d_lambda, d_mu, d_atol, d_rtol, d_istatus, d_rstatus,
                                         d_ierr, tape_budget, st, tm, err);
  return small_adj_run<RobertsAdj>(rp, fix, ncells, nadj, with_mu, d_y, d_lambda, d_mu, d_atol, d_rtol, d_istatus, d_rstatus, d_ierr,
                                   tape_budget, st, tm, err);
}

}  // namespace fatode
