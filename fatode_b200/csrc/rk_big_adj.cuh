// Discrete adjoint of the 3-stage implicit Runge-Kutta methods, DIRECT adjoint solve (ADJ/RK_ADJ, ICNTRL(7) = 2).
//
// Replaces the AdjointSolve == Solve_direct branch of rk_DadjInt (ADJ/RK_ADJ/RK_ADJ_f90_Integrator.F90:1328-1332, 1466-1499)
// with LSS_Decomp_Big / LSS_Solve_Big -> lapack_decomp_big / lapack_solve_big (ADJ/RK_ADJ/LS_Solver.F90:59-80, 115-124): per
// accepted step ONE factorisation of the 3N x 3N matrix
//     e_big(a N + i, b N + j) = -H rkA(a,b) J_b(i,j) + delta,          J_b = J(T + c_b H, Y + Z_b),
// by DGETRF (netlib's blocked path for 96 > NB = 64 performs, element by element, the operations of the unblocked DGETF2 in the
// same order: DTRSM/DGEMM of the reference BLAS subtract the products k = 1, 2, ... one after the other), and per adjoint
// column one DGETRS('T') of X = -(G_1, G_2, G_3), G_s = -H b_s J_s^T Lambda; then U_s = X_s.
//
// One system per CTA of three warps.  The factorisation is block-cooperative in shared memory (96 x 96 doubles, column
// major): IDAMAX by a block arg-max with netlib's first-maximum rule, DSWAP of the two full rows by 96 threads, DSCAL by the
// reciprocal (or the division below DLAMCH('S')), DGER with the rows of the trailing block on the lanes and its columns dealt
// round-robin to the three warps.  The solves give lane m of warp 0 column m: its 96-vector lives in the lane's shared-memory
// slots, Lambda and Mu in its registers, the factors are read with warp-uniform addresses; DTRSM('L','U','T','N'),
// DTRSM('L','L','T','U') and DLASWP(-1) in netlib's dot-product loops, no fma: Lambda is bit-identical to the CPU restatement
// (oracle/rk_adj.hpp), Mu to rounding (the stoichiometric regrouping of JACP^T shared with k_rk_adj_cell).
// The records come from the preparation kernels of the fixed-sweep mode (static-pattern or dense): only T, H, the three stage
// Jacobians and the reactant products are read.
#pragma once
#include "rk_cell_adj.cuh"

namespace fatode {

constexpr int RB_N = 96;                                  // 3 N
constexpr int RB_THREADS = 96;
constexpr int RB_TAIL = 3 * 276 + 3 * 32;                 // stage Jacobians + reactant products: contiguous tail of both record kinds
constexpr int RB_OFF_E = 0;
constexpr int RB_OFF_X = RB_OFF_E + RB_N * RB_N;          // [96 rows][32 lanes]
constexpr int RB_OFF_REC = RB_OFF_X + RB_N * 32;          // hdr (4) + tail
constexpr int RB_OFF_PIV = RB_OFF_REC + 4 + RB_TAIL;      // 96 ints
constexpr int RB_OFF_RED = RB_OFF_PIV + RB_N / 2;         // arg-max scratch of the three warps + queue slot
constexpr int RB_DOUBLES = RB_OFF_RED + 8;
constexpr size_t RK_ADJ_BIG_SMEM = sizeof(double) * (size_t)RB_DOUBLES;   // 106 KB: two systems per SM

struct RkBigRecord { int stride, off_tail; };             // doubles per preparation record, offset of its J | P tail

__global__ void __launch_bounds__(RB_THREADS, 2)
k_rk_adj_big(RkParams rp, RkAdjArgs a, RkBigRecord rl) {
  extern __shared__ __align__(16) double cell_smem[];
  constexpr int N = 32, NB = RB_N;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  double* const E = cell_smem + RB_OFF_E;
  double* const X = cell_smem + RB_OFF_X + lane;          // this lane's vector: row r at X[r * 32]
  double* const REC = cell_smem + RB_OFF_REC;             // [0..3] header, then J_1..3 (276 each), then the reactant products (32 each)
  int* const PIV = reinterpret_cast<int*>(cell_smem + RB_OFF_PIV);
  double* const RED = cell_smem + RB_OFF_RED;             // [0..2] warp maxima, [3..5] their rows (as doubles), [6] queue slot
  const double sfmin = 2.2250738585072014e-308;           // DLAMCH('S')
  const bool active = (warp == 0) && lane < a.nadj;
  for (;;) {
    __syncthreads();
    if (tid == 0) RED[6] = (double)atomicAdd(a.queue, 1ull);
    __syncthreads();
    const unsigned long long slot = (unsigned long long)RED[6];
    if (slot >= (unsigned long long)a.ncells) break;
    const long long r0 = a.rec_off[slot], r1 = a.rec_off[slot + 1];
    const int cell = a.cell[slot];
    double lam[N], mu[N];
#pragma unroll
    for (int i = 0; i < N; ++i) { lam[i] = (i == lane) ? 1.0 : 0.0; mu[i] = 0.0; }   // ADJINIT of the example drivers
    for (long long q = r1 - 1; q >= r0; --q) {
      __syncthreads();
      {   // header + the J | P tail of the record
        const double* src = a.fat + (size_t)q * rl.stride;
        if (tid < 4) REC[tid] = src[tid];
        for (int e = tid; e < RB_TAIL; e += RB_THREADS) REC[4 + e] = __ldg(src + rl.off_tail + e);
      }
      __syncthreads();
      const double H = REC[1];
      // lapack_decomp_big: e_big = -H rkA(a,b) fjac_b + I.  Entries outside the Jacobian pattern are the signed zeros of the product.
      for (int idx = tid; idx < NB * NB; idx += RB_THREADS) {
        const int i = idx % NB, j = idx / NB;
        E[idx] = (-H * rp.A[i / N + 1][j / N + 1]) * 0.0;
      }
      __syncthreads();
      for (int e = tid; e < 3 * 276; e += RB_THREADS) {
        const int b = e / 276, ee = e - b * 276;
        const int r = cbm4_dev::JAC_ROW[ee], c = cbm4_dev::JAC_COL[ee];
        const double v = REC[4 + e];
#pragma unroll
        for (int ab = 0; ab < 3; ++ab) E[(ab * N + r) + NB * (b * N + c)] = (-H * rp.A[ab + 1][b + 1]) * v;
      }
      __syncthreads();
      E[tid + NB * tid] = E[tid + NB * tid] + 1;
      __syncthreads();
      // DGETF2 (oracle/ref_lapack.hpp::dgetf2 operation by operation)
      for (int k = 0; k < NB; ++k) {
        {   // IDAMAX over rows k..95 of column k: first index of the maximum modulus
          double v = (tid >= k) ? fabs(E[tid + NB * k]) : -1.0;
          int row = tid;
#pragma unroll
          for (int off = 16; off > 0; off >>= 1) {
            const double ov = __shfl_down_sync(0xffffffffu, v, off);
            const int orow = __shfl_down_sync(0xffffffffu, row, off);
            if (ov > v || (ov == v && orow < row)) { v = ov; row = orow; }
          }
          if (lane == 0) { RED[warp] = v; RED[3 + warp] = (double)row; }
        }
        __syncthreads();
        int jp;
        {
          double v = RED[0]; jp = (int)RED[3];
          if (RED[1] > v) { v = RED[1]; jp = (int)RED[4]; }
          if (RED[2] > v) { v = RED[2]; jp = (int)RED[5]; }
        }
        const double pv = E[jp + NB * k];
        if (tid == 0) PIV[k] = jp;
        __syncthreads();                                   // every thread has read RED and the pivot before rows move
        if (pv != 0.0) {
          if (jp != k) {                                   // DSWAP of the two full rows: thread = column
            const double t = E[k + NB * tid];
            E[k + NB * tid] = E[jp + NB * tid];
            E[jp + NB * tid] = t;
          }
          __syncthreads();
          if (k < NB - 1 && tid > k) {                     // DSCAL by the reciprocal, or the division below sfmin
            const double d = E[k + NB * k];
            if (fabs(d) >= sfmin) { const double r = 1.0 / d; E[tid + NB * k] = r * E[tid + NB * k]; }
            else E[tid + NB * k] = E[tid + NB * k] / d;
          }
        }
        __syncthreads();
        if (k < NB - 1) {   // DGER: rows of the trailing block on the lanes (three per lane), its columns dealt to the warps
          const double l0 = E[lane + NB * k], l1 = E[lane + 32 + NB * k], l2 = E[lane + 64 + NB * k];   // x = column k of L
          const bool r0 = lane > k, r1 = lane + 32 > k, r2 = lane + 64 > k;
          for (int c = k + 1 + warp; c < NB; c += 3) {
            const double yc = E[k + NB * c];
            if (yc != 0.0) {
              const double temp = -1.0 * yc;
              double* col = E + NB * c;
              const double a0 = col[lane], a1 = col[lane + 32], a2 = col[lane + 64];
              if (r0) col[lane] = a0 + l0 * temp;
              if (r1) col[lane + 32] = a1 + l1 * temp;
              if (r2) col[lane + 64] = a2 + l2 * temp;
            }
          }
        }
        __syncthreads();
      }
      if (!active) continue;
      // G_s = -H b_s J_s^T Lambda (:1337-1347), X = -G (:1468-1470)
      {
        double o[N];
#pragma unroll 1
        for (int s = 0; s < 3; ++s) {
          cbm4_cell::jt_vec_seq(REC + 4 + s * 276, lam, o);
          const double c = -H * rp.B[s + 1];
#pragma unroll
          for (int i = 0; i < N; ++i) X[(s * N + i) * 32] = -(0.0 + c * o[i]);   // G = 0 + c J^T Lambda (SET2ZERO, DAXPY), X = -G
        }
      }
      // DGETRS('T'): DTRSM(L,U,T,N), DTRSM(L,L,T,U), DLASWP(-1) in netlib's dot-product loops (an update-as-you-go U^T solve has the
      // same operations per element in the same order and was measured 20 % slower: one more shared-memory store per term)
      for (int i = 0; i < NB; ++i) {
        double temp = X[i * 32];
        const double* col = E + NB * i;
#pragma unroll 8
        for (int k = 0; k < i; ++k) temp = temp - col[k] * X[k * 32];
        X[i * 32] = temp / col[i];
      }
      for (int i = NB - 1; i >= 0; --i) {            // L^T: element i takes its products in ascending k, only the dot form has that order
        double temp = X[i * 32];
        const double* col = E + NB * i;
#pragma unroll 8
        for (int k = i + 1; k < NB; ++k) temp = temp - col[k] * X[k * 32];
        X[i * 32] = temp;
      }
      for (int i = NB - 1; i >= 0; --i) {
        const int ip = PIV[i];
        if (ip != i) { const double t = X[i * 32]; X[i * 32] = X[ip * 32]; X[ip * 32] = t; }
      }
      // Mu += sum_s JACP_s^T (H sum_j a_js X_j + H b_s Lambda)   (:1472-1494)
      if (a.with_mu) {
        double x[N];
#pragma unroll 1
        for (int s = 0; s < 3; ++s) {
          const double c1 = H * rp.A[1][s + 1], c2 = H * rp.A[2][s + 1], c3 = H * rp.A[3][s + 1], cb = H * rp.B[s + 1];
#pragma unroll
          for (int i = 0; i < N; ++i) x[i] = ((c1 * X[i * 32] + c2 * X[(N + i) * 32]) + c3 * X[(2 * N + i) * 32]) + cb * lam[i];
          const double* AP = REC + 4 + 3 * 276 + s * 32;
          Cbm4Warp::stoich_cols(x, [&](auto pc, double sp) {
            constexpr int p = decltype(pc)::value;
            mu[p] = fma(AP[p], sp, mu[p]);
          });
        }
      }
#pragma unroll
      for (int i = 0; i < N; ++i) lam[i] = ((lam[i] + X[i * 32]) + X[(N + i) * 32]) + X[(2 * N + i) * 32];   // :1496
    }
    if (active) {   // Lambda(N,NADJ), Mu(NP,NADJ) column-major per system
      double* lo = a.lambda + ((size_t)cell * a.nadj + lane) * N;
#pragma unroll
      for (int i = 0; i < N; ++i) lo[i] = lam[i];
      if (a.with_mu) {
        double* mo = a.mu + ((size_t)cell * a.nadj + lane) * a.np;
#pragma unroll
        for (int p = 0; p < N; ++p) if (p < a.np) mo[p] = mu[p];
      }
    }
    if (tid == 0) {   // backward counters: Njac 1 + 3, RK_Decomp and LSS_Decomp_Big, one LSS_Solve_Big per column (:1290-1299, 1331, 1472)
      const int A = (int)(r1 - r0);
      int* is_ = a.istatus + (size_t)cell * 20;
      is_[Njac] += 4 * A;
      is_[Ndec] += 2 * A;
      is_[Nstp] += A;
      is_[Nsol] += A * a.nadj;
    }
  }
}

}  // namespace fatode
