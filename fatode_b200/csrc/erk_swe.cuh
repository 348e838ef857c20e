// Explicit Runge-Kutta family on the shallow-water ensemble: FWD/ERK INTEGRATE and TLM/ERK_TLM INTEGRATE_TLM.
//
// Replaces ERK_Integrator (FWD/ERK/ERK_f90_Integrator.F90:364-503), ERK_IntegratorTLM in its preset finite-difference mode
// (TLM/ERK_TLM/ERK_TLM_f90_Integrator.F90:391-576, FDJACV :1071-1097) and the model callback FUN = vec2grid -> compute_F ->
// grid2vec of the SWE example (EXAMPLES/swe/SWE_ERK/swe2D_upwind.F90:267-567, swe2D_erk_dr.F90:2-14).
//
// Mapping: one ensemble member per CTA (320 threads, 5 grid cells = 15 vector components per thread), members handed out by an
// atomic queue.  The member's state Y and the stage vector TMP live in shared memory (2 x 38.4 KB; two CTAs per SM); the
// stage derivatives K_1..K_S (and K_tlm) go to a per-CTA scratch in global memory that is reused by every member the CTA
// integrates, so it stays L2-resident: HBM sees each member's Y once in and once out.  The upwind stencil reads its 27
// neighbours from shared memory with periodic index wrap (the reference's halo copy, :338-350).
//
// Parity: every floating-point operation is the reference's, in the reference's order (nvcc -fmad=false): stage sums as the
// sequence of DAXPYs (a zero factor is skipped, as netlib's DAXPY returns at once), the flux splitting as written (terms with a
// structurally zero coefficient dropped: they only add a signed zero), division by the constants 6 dx, 6 dy through the
// correctly rounded reciprocal-and-correct form, and the error norm as ONE sequential sum over the 4800 components (thread 0;
// the other CTA of the SM computes meanwhile).  State, ISTATUS and RSTATUS are bit-identical to the oracle.
#pragma once
#include <cstring>
#include <cmath>
#include "portable_math.h"
#include "gen/erk_tableau_gen.h"

namespace fatode {

constexpr int SWE_MX = 40, SWE_CELLS = 1600, SWE_N = 4800;
constexpr int ERK_THREADS = 320, ERK_CPT = SWE_CELLS / ERK_THREADS;   // cells per thread
constexpr int ERK_EPT = SWE_N / ERK_THREADS;                            // vector components per thread
static_assert(ERK_CPT * ERK_THREADS == SWE_CELLS, "cells per thread");
constexpr size_t ERK_SMEM = sizeof(double) * (2 * SWE_N + 8);

}  // namespace fatode
#include "swe_jac_dev.cuh"   // the example's JAC callback, matrix-free (dense-Jacobian TLM mode here, ERK_ADJ in erk_adj_swe.cu)
namespace fatode {

struct ErkParams {
  int S;
  double ELO, A[12][12], B[12], C[12], E[12];
  int vector_tol, max_no_steps;
  double roundoff, hmin, hmax, hstart, facmin, facmax, facrej, facsafe, qmin, fdinc;
  double tin, tout;
  double c6x, c6y, r6x, r6y;   // 6 dx, 6 dy and their correctly rounded reciprocals
};

__device__ __forceinline__ double erk_cdiv(const double a, const double b, const double rb) {   // RN(a / b) from rb = RN(1 / b)
  const double q = a * rb;
  return fma(fma(-q, b, a), rb, q);
}

// compute_F for the grid cell (a, b) = (i - 3, j - 3) of the reference, U(k, i, j) at T[k * 1600 + a * 40 + b]
__device__ __forceinline__ void swe_rhs(const double* __restrict__ T, const int a, const int b, const ErkParams& p, double& f0,
                                        double& f1, double& f2) {
  const double g = 9.80616;
  const int ap1 = (a + 1 < SWE_MX) ? a + 1 : a + 1 - SWE_MX, ap2 = (a + 2 < SWE_MX) ? a + 2 : a + 2 - SWE_MX;
  const int am1 = (a >= 1) ? a - 1 : a - 1 + SWE_MX, am2 = (a >= 2) ? a - 2 : a - 2 + SWE_MX;
  const int bp1 = (b + 1 < SWE_MX) ? b + 1 : b + 1 - SWE_MX, bp2 = (b + 2 < SWE_MX) ? b + 2 : b + 2 - SWE_MX;
  const int bm1 = (b >= 1) ? b - 1 : b - 1 + SWE_MX, bm2 = (b >= 2) ? b - 2 : b - 2 + SWE_MX;
  double uc0[3], dxm[3], dxp[3], dym[3], dyp[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const double* Tk = T + k * SWE_CELLS;
    const double c = Tk[a * SWE_MX + b];
    const double xp2 = Tk[ap2 * SWE_MX + b], xp1 = Tk[ap1 * SWE_MX + b], xm1 = Tk[am1 * SWE_MX + b], xm2 = Tk[am2 * SWE_MX + b];
    const double yp2 = Tk[a * SWE_MX + bp2], yp1 = Tk[a * SWE_MX + bp1], ym1 = Tk[a * SWE_MX + bm1], ym2 = Tk[a * SWE_MX + bm2];
    uc0[k] = c;
    dxm[k] = erk_cdiv(-xp2 + 6.0 * xp1 - 3.0 * c - 2.0 * xm1, p.c6x, p.r6x);
    dxp[k] = erk_cdiv(2.0 * xp1 + 3.0 * c - 6.0 * xm1 + xm2, p.c6x, p.r6x);
    dym[k] = erk_cdiv(-yp2 + 6.0 * yp1 - 3.0 * c - 2.0 * ym1, p.c6y, p.r6y);
    dyp[k] = erk_cdiv(2.0 * yp1 + 3.0 * c - 6.0 * ym1 + ym2, p.c6y, p.r6y);
  }
  const double hc = uc0[0], uc = uc0[1] / hc, vc = uc0[2] / hc;
  const double cs = sqrt(g * hc), isqrt = 1.0 / cs;
  // x direction
  double tF0, tF1, tF2;
  {
    const double e0 = uc, e1 = uc - cs, e2 = uc + cs;
    const bool s0 = (e0 >= 0.0), s1 = (e1 >= 0.0), s2 = (e2 >= 0.0);
    const double d0[3] = {s0 ? dxp[0] : dxm[0], s0 ? dxp[1] : dxm[1], s0 ? dxp[2] : dxm[2]};
    const double d1[3] = {s1 ? dxp[0] : dxm[0], s1 ? dxp[1] : dxm[1], s1 ? dxp[2] : dxm[2]};
    const double d2[3] = {s2 ? dxp[0] : dxm[0], s2 ? dxp[1] : dxm[1], s2 ? dxp[2] : dxm[2]};
    const double m1_2 = (-vc) * d0[0] + d0[2];
    const double A00 = 0.5 * (1.0 + uc * isqrt), A01 = -0.5 * isqrt, A10 = 0.5 * (uc * uc - g * hc) * isqrt,
                 A11 = 0.5 - 0.5 * uc * isqrt, A20 = 0.5 * (1.0 + uc * isqrt) * vc, A21 = -0.5 * vc * isqrt;
    const double m2_0 = A00 * d1[0] + A01 * d1[1], m2_1 = A10 * d1[0] + A11 * d1[1], m2_2 = A20 * d1[0] + A21 * d1[1];
    const double B00 = 0.5 - 0.5 * uc * isqrt, B01 = 0.5 * isqrt, B10 = 0.5 * (g * hc - uc * uc) * isqrt,
                 B11 = 0.5 * (1.0 + uc * isqrt), B20 = 0.5 * (1.0 - uc * isqrt) * vc, B21 = 0.5 * vc * isqrt;
    const double m3_0 = B00 * d2[0] + B01 * d2[1], m3_1 = B10 * d2[0] + B11 * d2[1], m3_2 = B20 * d2[0] + B21 * d2[1];
    tF0 = e1 * m2_0 + e2 * m3_0;
    tF1 = e1 * m2_1 + e2 * m3_1;
    tF2 = (e0 * m1_2 + e1 * m2_2) + e2 * m3_2;
  }
  // y direction
  double tG0, tG1, tG2;
  {
    const double e0 = vc, e1 = vc - cs, e2 = vc + cs;
    const bool s0 = (e0 >= 0.0), s1 = (e1 >= 0.0), s2 = (e2 >= 0.0);
    const double d0[3] = {s0 ? dyp[0] : dym[0], s0 ? dyp[1] : dym[1], s0 ? dyp[2] : dym[2]};
    const double d1[3] = {s1 ? dyp[0] : dym[0], s1 ? dyp[1] : dym[1], s1 ? dyp[2] : dym[2]};
    const double d2[3] = {s2 ? dyp[0] : dym[0], s2 ? dyp[1] : dym[1], s2 ? dyp[2] : dym[2]};
    const double m1_1 = (-uc) * d0[0] + d0[1];
    const double A00 = 0.5 * (1.0 + vc * isqrt), A02 = -0.5 * isqrt, A10 = 0.5 * uc * (1.0 + vc * isqrt), A12 = -0.5 * uc * isqrt,
                 A20 = 0.5 * (vc * vc - g * hc) * isqrt, A22 = 0.5 - 0.5 * vc * isqrt;
    const double m2_0 = A00 * d1[0] + A02 * d1[2], m2_1 = A10 * d1[0] + A12 * d1[2], m2_2 = A20 * d1[0] + A22 * d1[2];
    const double B00 = 0.5 * (1.0 - vc * isqrt), B02 = 0.5 * isqrt, B10 = 0.5 * uc * (1.0 - vc * isqrt), B12 = 0.5 * uc * isqrt,
                 B20 = 0.5 * (-vc * vc + g * hc) * isqrt, B22 = 0.5 + 0.5 * vc * isqrt;
    const double m3_0 = B00 * d2[0] + B02 * d2[2], m3_1 = B10 * d2[0] + B12 * d2[2], m3_2 = B20 * d2[0] + B22 * d2[2];
    tG0 = e1 * m2_0 + e2 * m3_0;
    tG1 = (e0 * m1_1 + e1 * m2_1) + e2 * m3_1;
    tG2 = e1 * m2_2 + e2 * m3_2;
  }
  f0 = -tF0 - tG0;
  f1 = -tF1 - tG1;
  f2 = -tF2 - tG2;
}

// FUN over the member's grid: K[v] = f(TMP)[v] for the thread's cells.  SUB: K = (f - F0) * (1 / FDincrement) (FDJACV).
template <bool SUB>
__device__ __forceinline__ void swe_fun_cells(const double* __restrict__ TMP, const ErkParams& p, double* __restrict__ K,
                                              const double* __restrict__ F0, const double rinc) {
#pragma unroll 1
  for (int m = 0; m < ERK_CPT; ++m) {
    const int c = threadIdx.x + ERK_THREADS * m;
    const int a = c / SWE_MX, b = c - a * SWE_MX;
    double f0, f1, f2;
    swe_rhs(TMP, a, b, p, f0, f1, f2);
    if constexpr (SUB) {
      f0 = rinc * (f0 - F0[c]);
      f1 = rinc * (f1 - F0[SWE_CELLS + c]);
      f2 = rinc * (f2 - F0[2 * SWE_CELLS + c]);
    }
    K[c] = f0;
    K[SWE_CELLS + c] = f1;
    K[2 * SWE_CELLS + c] = f2;
  }
}

struct ErkArgs {
  long nmembers;
  const double* y_in;
  double* y_out;
  double* y_tlm;               // [member][ntlm][N] (TLM only)
  const double* atol;
  const double* rtol;
  int* istatus;
  double* rstatus;
  int* ierr;
  unsigned long long* queue;
  double* scratch;             // per CTA: S * N (+ ntlm * S * N) doubles
  int ntlm;
  int tlm_trunc;               // ICNTRL(12) /= 0: the columns' truncation errors enter the step control
  const double* atol_tlm;      // [ntlm][N] (read with tlm_trunc only)
  const double* rtol_tlm;
  int tlm_dense;               // ICNTRL(5) /= 0: K_tlm = MATMUL(JAC, S_tlm) with the example's JAC callback (scratch: + 27 * N)
};

template <bool TLM>
__global__ void __launch_bounds__(ERK_THREADS, 2)
k_erk_swe(ErkParams p, ErkArgs a) {
  extern __shared__ __align__(16) double erk_smem[];
  double* const Ys = erk_smem;
  double* const TMP = erk_smem + SWE_N;
  double* const red = erk_smem + 2 * SWE_N;       // [0] = Err, [1] = member slot (as double bits)
  const int tid = threadIdx.x;
  const int S = p.S;
  const int ntlm = TLM ? a.ntlm : 0;
  double* const K = a.scratch + (size_t)blockIdx.x * ((size_t)S * SWE_N * (1 + ntlm) + (size_t)((TLM && a.tlm_dense) ? SWE_NNZR * SWE_N : 0));
  double* const KT = K + (size_t)S * SWE_N;       // K_tlm[col][stage][N]
  double* const JV = KT + (size_t)ntlm * S * SWE_N;   // [27][N] rows of JAC at the stage state (dense-Jacobian TLM mode)
  const bool dense = TLM && a.tlm_dense;
  const double dxg = p.c6x / 6.0;
  int njac = 0;
  const double Roundoff = p.roundoff;
  const double Tdirection = (p.tout - p.tin) >= 0.0 ? 1.0 : -1.0;
  const double rinc = TLM ? 1.0 / p.fdinc : 0.0;
  for (;;) {
    __syncthreads();
    if (tid == 0) reinterpret_cast<unsigned long long*>(red)[1] = atomicAdd(a.queue, 1ull);
    __syncthreads();
    const unsigned long long member = reinterpret_cast<unsigned long long*>(red)[1];
    if (member >= (unsigned long long)a.nmembers) break;
    const double* yin = a.y_in + (size_t)member * SWE_N;
    double* const ytl = TLM ? a.y_tlm + (size_t)member * SWE_N * ntlm : nullptr;
    for (int v = tid; v < SWE_N; v += ERK_THREADS) Ys[v] = yin[v];
    int nfun = 0, nstp = 0, nacc = 0, nrej = 0, ierr = 1;
    njac = 0;
    double T = p.tin, texit = 0.0, hexit = 0.0, hnew_out = 0.0;
    double H = fmax(fabs(p.hmin), fabs(p.hstart));
    if (fabs(H) <= 10.0 * Roundoff) H = 1.0e-6;
    H = fmin(fabs(H), p.hmax);
    H = copysign(H, Tdirection);
    bool Reject = false, FirstStep = true;
    __syncthreads();
    while ((p.tout - T) * Tdirection - Roundoff > 0.0) {
      if (nstp > p.max_no_steps) { ierr = -6; break; }
      if ((T + 0.1 * H == T) || (fabs(H) <= Roundoff)) { ierr = -7; break; }
#pragma unroll 1
      for (int is = 0; is < S; ++is) {
        // TMP = sum_j (H a_ij) K_j + Y   (:411-421)
        {   // stage-major so that the thread's 15 loads of one K_j are in flight together; per element the order is the DAXPYs'
          double t[ERK_EPT];
#pragma unroll
          for (int q = 0; q < ERK_EPT; ++q) t[q] = 0.0;
          for (int j = 0; j < is; ++j) {
            const double c = H * p.A[is][j];
            if (c != 0.0) {
              const double* Kj = K + (size_t)j * SWE_N + tid;
#pragma unroll
              for (int q = 0; q < ERK_EPT; ++q) t[q] = t[q] + c * Kj[q * ERK_THREADS];
            }
          }
#pragma unroll
          for (int q = 0; q < ERK_EPT; ++q) TMP[tid + q * ERK_THREADS] = t[q] + Ys[tid + q * ERK_THREADS];
        }
        __syncthreads();
        double* Ki = K + (size_t)is * SWE_N;
        swe_fun_cells<false>(TMP, p, Ki, nullptr, 0.0);
        nfun++;
        if constexpr (TLM) {
          if (dense) {   // FJAC = JAC(T + c H, TMP); K_TLM(:, istage, itlm) = MATMUL(FJAC, S_TLM)  (:460-474)
            swe_jac_rows(TMP, JV, dxg);
            njac++;
            for (int col = 0; col < ntlm; ++col) {
              __syncthreads();   // every thread has finished reading TMP (the rows, the previous column's product)
              for (int v = tid; v < SWE_N; v += ERK_THREADS) {
                double s = ytl[(size_t)col * SWE_N + v];
                for (int j = 0; j < is; ++j) {
                  const double c = H * p.A[is][j];
                  if (c != 0.0) s = s + c * KT[((size_t)col * S + j) * SWE_N + v];
                }
                TMP[v] = s;
              }
              __syncthreads();
              swe_jac_n_gather(JV, TMP, KT + ((size_t)col * S + is) * SWE_N);
            }
          } else
          for (int col = 0; col < ntlm; ++col) {
            __syncthreads();   // every thread has finished reading TMP
            double* Kc = KT + ((size_t)col * S + is) * SWE_N;
            for (int v = tid; v < SWE_N; v += ERK_THREADS) {
              double s = ytl[(size_t)col * SWE_N + v];
              for (int j = 0; j < is; ++j) {
                const double c = H * p.A[is][j];
                if (c != 0.0) s = s + c * KT[((size_t)col * S + j) * SWE_N + v];
              }
              TMP[v] = TMP[v] + p.fdinc * s;       // FDJACV perturbs the stage vector in place (:1088)
            }
            __syncthreads();
            swe_fun_cells<true>(TMP, p, Kc, Ki, rinc);
            nfun++;
          }
        }
        __syncthreads();
      }
      nstp++;
      // error estimate (:430-438): TMP <- (scaled error component)^2, then one sequential sum
      {
        double t[ERK_EPT];
#pragma unroll
        for (int q = 0; q < ERK_EPT; ++q) t[q] = 0.0;
        for (int i = 0; i < S; ++i)
          if (p.E[i] != 0.0) {
            const double c = H * p.E[i];
            if (c != 0.0) {
              const double* Ki = K + (size_t)i * SWE_N + tid;
#pragma unroll
              for (int q = 0; q < ERK_EPT; ++q) t[q] = t[q] + c * Ki[q * ERK_THREADS];
            }
          }
#pragma unroll
        for (int q = 0; q < ERK_EPT; ++q) {
          const int v = tid + q * ERK_THREADS;
          const int vt = p.vector_tol ? v : 0;
          const double scal = 1.0 / (a.atol[vt] + a.rtol[vt] * fabs(Ys[v]));
          const double w = t[q] * scal;
          TMP[v] = w * w;
        }
      }
      __syncthreads();
      if (tid == 0) {
        double e = 0.0;
#pragma unroll 8
        for (int v = 0; v < SWE_N; ++v) e = e + TMP[v];
        red[0] = fmax(sqrt(e / (double)SWE_N), 1.0e-10);
      }
      __syncthreads();
      if constexpr (TLM) {
        if (a.tlm_trunc) {   // ERK_ErrorNorm_TLM (TLM/ERK_TLM/ERK_TLM_f90_Integrator.F90:493-501, 613-626): scaled by the error vector itself
          for (int col = 0; col < ntlm; ++col) {
            double t[ERK_EPT];
#pragma unroll
            for (int q = 0; q < ERK_EPT; ++q) t[q] = 0.0;
            for (int i = 0; i < S; ++i)
              if (p.E[i] != 0.0) {
                const double c = H * p.E[i];
                if (c != 0.0) {
                  const double* Ki = KT + ((size_t)col * S + i) * SWE_N + tid;
#pragma unroll
                  for (int q = 0; q < ERK_EPT; ++q) t[q] = t[q] + c * Ki[q * ERK_THREADS];
                }
              }
#pragma unroll
            for (int q = 0; q < ERK_EPT; ++q) {
              const int v = tid + q * ERK_THREADS;
              const size_t vt = (size_t)col * SWE_N + (p.vector_tol ? v : 0);
              const double scal = 1.0 / (a.atol_tlm[vt] + a.rtol_tlm[vt] * fabs(t[q]));
              const double w = t[q] * scal;
              TMP[v] = w * w;
            }
            __syncthreads();
            if (tid == 0) {
              double e = 0.0;
#pragma unroll 8
              for (int v = 0; v < SWE_N; ++v) e = e + TMP[v];
              red[0] = fmax(red[0], fmax(sqrt(e / (double)SWE_N), 1.0e-10));
            }
            __syncthreads();
          }
        }
      }
      const double Err = red[0];
      double Fac = p.facsafe * fpm::pow_neg_inv(Err, p.ELO);
      Fac = fmax(p.facmin, fmin(p.facmax, Fac));
      double Hnew = H * Fac;
      if (Err < 1.0) {
        FirstStep = false;
        nacc++;
        T = T + H;
        {
          double t[ERK_EPT];
#pragma unroll
          for (int q = 0; q < ERK_EPT; ++q) t[q] = Ys[tid + q * ERK_THREADS];
          for (int i = 0; i < S; ++i)
            if (p.B[i] != 0.0) {
              const double c = H * p.B[i];
              if (c != 0.0) {
                const double* Ki = K + (size_t)i * SWE_N + tid;
#pragma unroll
                for (int q = 0; q < ERK_EPT; ++q) t[q] = t[q] + c * Ki[q * ERK_THREADS];
              }
            }
#pragma unroll
          for (int q = 0; q < ERK_EPT; ++q) Ys[tid + q * ERK_THREADS] = t[q];
        }
        for (int v = tid; v < SWE_N; v += ERK_THREADS) {
          if constexpr (TLM) {
            for (int col = 0; col < ntlm; ++col) {
              double s = ytl[(size_t)col * SWE_N + v];
              for (int i = 0; i < S; ++i) {
                const double c = H * p.B[i];
                if (c != 0.0) s = s + c * KT[((size_t)col * S + i) * SWE_N + v];
              }
              ytl[(size_t)col * SWE_N + v] = s;
            }
          }
        }
        Hnew = Tdirection * fmin(fabs(Hnew), p.hmax);
        texit = T; hexit = H; hnew_out = Hnew;
        if (Reject) Hnew = Tdirection * fmin(fabs(Hnew), fabs(H));
        Reject = false;
        if ((T + Hnew / p.qmin - p.tout) * Tdirection > 0.0) H = p.tout - T;
        else H = Hnew;
      } else {
        if (FirstStep || Reject) H = p.facrej * H; else H = Hnew;
        Reject = true;
        if (nacc >= 1) nrej++;
      }
      __syncthreads();
    }
    __syncthreads();
    double* yout = a.y_out + (size_t)member * SWE_N;
    for (int v = tid; v < SWE_N; v += ERK_THREADS) yout[v] = Ys[v];
    if (tid == 0) {
      a.ierr[member] = ierr;
      int* ist = a.istatus + member * 20;
      for (int i = 0; i < 20; ++i) ist[i] = 0;
      ist[0] = nfun; ist[1] = njac; ist[2] = nstp; ist[3] = nacc; ist[4] = nrej;
      double* rst = a.rstatus + member * 20;
      for (int i = 0; i < 20; ++i) rst[i] = 0.0;
      rst[0] = texit; rst[1] = hexit; rst[2] = hnew_out;
    }
  }
}

// host: option handling of ERK (:229-358) / ERKTLM (:238-389).  Returns the reference's Ierr before integrating (0 = go on).
// tlm != nullptr: tlm[0] = ICNTRL(5) (0 = finite-difference Jacobian-vector products), tlm[1] = ICNTRL(12) (TLMtruncErr).
inline int erk_parse_options(int N, const int* icntrl_u, const double* rcntrl_u, double tin, double tout, const double* atol,
                             const double* rtol, ErkParams& p, int* tlm = nullptr) {
  int ic[20];
  double rc[20];
  std::memset(ic, 0, sizeof ic);
  std::memset(rc, 0, sizeof rc);
  for (int i = 0; i < 20; ++i) {   // INTEGRATE :56-61: user values override only when > 0
    if (icntrl_u && icntrl_u[i] > 0) ic[i] = icntrl_u[i];
    if (rcntrl_u && rcntrl_u[i] > 0) rc[i] = rcntrl_u[i];
  }
  std::memset(&p, 0, sizeof p);
  int ierr = 0;
  p.vector_tol = (ic[1] == 0);
  int method = 4;   // 0, 4 and CASE DEFAULT: Dopri5 (:229-244)
  if (ic[2] == 1 || ic[2] == 2 || ic[2] == 3 || ic[2] == 5 || ic[2] == 6) method = ic[2];
  const ErkTableauData& t = ERK_TABLEAU[method - 1];
  p.S = t.S; p.ELO = t.ELO;
  std::memcpy(p.A, t.A, sizeof p.A);
  std::memcpy(p.B, t.B, sizeof p.B);
  std::memcpy(p.C, t.C, sizeof p.C);
  std::memcpy(p.E, t.E, sizeof p.E);
  if (ic[3] == 0) p.max_no_steps = 10000; else if (ic[3] > 0) p.max_no_steps = ic[3]; else ierr = -1;
  if (tlm) { tlm[0] = ic[4]; tlm[1] = ic[11]; }
  p.roundoff = 1.1102230246251565e-16;   // DLAMCH('E')
  const double span = std::fabs(tout - tin);
  if (rc[0] == 0.0) p.hmin = 0.0; else if (rc[0] > 0.0) p.hmin = rc[0]; else ierr = -3;
  if (rc[1] == 0.0) p.hmax = span; else if (rc[1] > 0.0) p.hmax = std::fmin(std::fabs(rc[1]), span); else ierr = -3;
  if (rc[2] == 0.0) p.hstart = std::fmax(p.hmin, p.roundoff); else if (rc[2] > 0.0) p.hstart = std::fmin(std::fabs(rc[2]), span); else ierr = -3;
  // REAL(4) literals of the reference (:291-317): 0.2, 10.0, 0.1, 0.9
  if (rc[3] == 0.0) p.facmin = (double)0.2f; else if (rc[3] > 0.0) p.facmin = rc[3]; else ierr = -4;
  if (rc[4] == 0.0) p.facmax = (double)10.0f; else if (rc[4] > 0.0) p.facmax = rc[4]; else ierr = -4;
  if (rc[5] == 0.0) p.facrej = (double)0.1f; else if (rc[5] > 0.0) p.facrej = rc[5]; else ierr = -4;
  if (rc[6] == 0.0) p.facsafe = (double)0.9f; else if (rc[6] > 0.0) p.facsafe = rc[6]; else ierr = -4;
  p.qmin = (rc[9] == 0.0) ? 1.0 : rc[9];
  p.fdinc = (rc[11] == 0.0) ? std::sqrt(std::fmax(rtol[0], p.roundoff)) : rc[11];   // TLM/ERK_TLM :359-363
  if (!p.vector_tol) {
    if (atol[0] <= 0.0 || rtol[0] <= 10.0 * p.roundoff) ierr = -5;
  } else {
    for (int i = 0; i < N; ++i)
      if (atol[i] <= 0.0 || rtol[i] <= 10.0 * p.roundoff) ierr = -5;
  }
  p.tin = tin; p.tout = tout;
  const double dx = (3.0 - (-3.0)) / (SWE_MX - 1.0);   // swe2D_upwind.F90:31-32 (dx = dy)
  p.c6x = 6.0 * dx; p.c6y = 6.0 * dx;
  p.r6x = 1.0 / p.c6x; p.r6y = 1.0 / p.c6y;
  return ierr;
}

}  // namespace fatode
