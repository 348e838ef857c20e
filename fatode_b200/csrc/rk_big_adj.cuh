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

// lapack_decomp_big for the step whose header and J | P tail sit in REC: assembles e_big = I - H (A x J_b) in E (96 x 96, column
// major) and factorises it in place like DGETF2, block-cooperatively (all 96 threads).  Returns netlib's INFO (0, or k + 1 of the
// first exactly zero pivot); the interchanges go to PIV (0-based).
__device__ __forceinline__ int rb_decomp_big(const RkParams& rp, double* __restrict__ E, const double* __restrict__ REC,
                                             int* __restrict__ PIV, double* __restrict__ RED, double H, int tid, int warp, int lane) {
  constexpr int N = 32, NB = RB_N;
  const double sfmin = 2.2250738585072014e-308;           // DLAMCH('S')
  int info = 0;
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
    if (pv == 0.0 && info == 0) info = k + 1;
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
  return info;
}

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
      rb_decomp_big(rp, E, REC, PIV, RED, H, tid, warp, lane);   // ISING of LSS_Decomp_Big is never looked at (:1330)
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

// =====================================================================================================================
// Tangent linear model of the 3-stage implicit Runge-Kutta methods, the reference's preset DIRECT TLM solve (TLM/RK_TLM
// INTEGRATE_TLM, ICNTRL(7) = 1; RK_IntegratorTLM, TLM/RK_TLM/RK_TLM_f90_Integrator.F90:761-793, 909-914).
// On every accepted step the reference evaluates the three stage Jacobians, factorises e_big = I - H (A x J_b)
// (LSS_Decomp_Big) and solves, per column, e_big Zbig = H (A x J_b) Y_tlm (RK_PrepareRHS_TLMdirect :1185-1227, LSS_Solve_Big
// with Transp = .FALSE.); then Y_tlm += sum_i d_i Z_i.  With the state's error control only (ICNTRL(12) = 0, the preset) the
// columns never influence the step sequence, so the state sweep (k_rk_fwd_cell<.., ADJ = true> with RkParams::tlm set: the
// RK_IntegratorTLM flavour) tapes the accepted steps and this kernel walks the preparation records FORWARDS.
// Same mapping as k_rk_adj_big: one system per CTA of three warps, block-cooperative DGETF2, lane m of warp 0 = column m of
// Y_tlm (in registers), its 96-vector Zbig in shared-memory slots.  matmul(fjac_b, Y_tlm) as the sequential row sums of
// j_vec_seq, DGETRS('N') as DLASWP + netlib's two column-sweep DTRSMs (zero right-hand-side entries skipped like netlib; they only
// skip exact zeros), no fma: Y_tlm is bit-identical to the CPU restatement (oracle/rk_tlm.hpp).  A singular e_big (the
// reference STOPs: "Big big guy is singular") ends the system with IERR -12.
__global__ void __launch_bounds__(RB_THREADS, 2)
k_rk_tlm_big(RkParams rp, RkAdjArgs a, RkBigRecord rl) {
  extern __shared__ __align__(16) double cell_smem[];
  constexpr int N = 32, NB = RB_N;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  double* const E = cell_smem + RB_OFF_E;
  double* const REC = cell_smem + RB_OFF_REC;
  int* const PIV = reinterpret_cast<int*>(cell_smem + RB_OFF_PIV);
  double* const RED = cell_smem + RB_OFF_RED;
  const int ntlm = a.nadj;
  const bool active = (warp == 0) && lane < ntlm;
  for (;;) {
    __syncthreads();
    if (tid == 0) RED[6] = (double)atomicAdd(a.queue, 1ull);
    __syncthreads();
    const unsigned long long slot = (unsigned long long)RED[6];
    if (slot >= (unsigned long long)a.ncells) break;
    const long long r0 = a.rec_off[slot], r1 = a.rec_off[slot + 1];
    const int cell = a.cell[slot];
    double* const ycol = a.lambda + ((size_t)cell * ntlm + (lane < ntlm ? lane : 0)) * N;   // Y_tlm(:, m) of the caller
    double yt[N];
#pragma unroll
    for (int i = 0; i < N; ++i) yt[i] = active ? ycol[i] : 0.0;
    long long done = 0;
    int info = 0;
    for (long long q = r0; q < r1; ++q) {
      __syncthreads();
      {
        const double* src = a.fat + (size_t)q * rl.stride;
        if (tid < 4) REC[tid] = src[tid];
        for (int e = tid; e < RB_TAIL; e += RB_THREADS) REC[4 + e] = __ldg(src + rl.off_tail + e);
      }
      __syncthreads();
      const double H = REC[1];
      info = rb_decomp_big(rp, E, REC, PIV, RED, H, tid, warp, lane);
      ++done;
      if (info != 0) break;                               // uniform: every thread computed the same INFO
      if (!active) continue;
      double* __restrict__ const X = cell_smem + RB_OFF_X + lane;   // Zbig of this lane: row r at X[r * 32]
      const double* __restrict__ const Ef = E;
      {   // RK_PrepareRHS_TLMdirect: Zbig_a = sum_b (H rkA(a,b)) (J_b Y_tlm), b = 1, 2, 3 in this order
        double o[N];
#pragma unroll 1
        for (int b = 0; b < 3; ++b) {
          cbm4_cell::j_vec_seq(REC + 4 + b * 276, yt, o);
          const double c1 = H * rp.A[1][b + 1], c2 = H * rp.A[2][b + 1], c3 = H * rp.A[3][b + 1];
          if (b == 0) {
#pragma unroll
            for (int i = 0; i < N; ++i) { X[i * 32] = c1 * o[i]; X[(N + i) * 32] = c2 * o[i]; X[(2 * N + i) * 32] = c3 * o[i]; }
          } else {
#pragma unroll
            for (int i = 0; i < N; ++i) {
              X[i * 32] = X[i * 32] + c1 * o[i];
              X[(N + i) * 32] = X[(N + i) * 32] + c2 * o[i];
              X[(2 * N + i) * 32] = X[(2 * N + i) * 32] + c3 * o[i];
            }
          }
        }
      }
      // DGETRS('N'): DLASWP, DTRSM(L,L,N,U), DTRSM(L,U,N,N)
      for (int i = 0; i < NB; ++i) {
        const int ip = PIV[i];
        if (ip != i) { const double t = X[i * 32]; X[i * 32] = X[ip * 32]; X[ip * 32] = t; }
      }
      // netlib's column sweeps subtract b_k A(i,k) from element i for k = 1, 2, ... (L) resp. k = M, M-1, ... (U); the row-wise loops
      // below perform the same subtractions on every element in the same order with one running register per element instead
      // of a shared-memory read-modify-write per term (measured 5.7 -> 3.6 s per 9 472 systems).  netlib skips b_k = 0: that
      // only avoids subtracting an exact zero.
      for (int i = 1; i < NB; ++i) {
        double temp = X[i * 32];
        const double* __restrict__ row = Ef + i;
#pragma unroll 8
        for (int k = 0; k < i; ++k) temp = temp - X[k * 32] * row[NB * k];
        X[i * 32] = temp;
      }
      for (int i = NB - 1; i >= 0; --i) {
        double temp = X[i * 32];
        const double* __restrict__ row = Ef + i;
#pragma unroll 8
        for (int k = NB - 1; k > i; --k) temp = temp - X[k * 32] * row[NB * k];
        X[i * 32] = (temp != 0.0) ? temp / row[NB * i] : temp;
      }
      // Y_tlm <- Y_tlm + sum_i d_i Z_i   (:909-914)
#pragma unroll 1
      for (int s = 1; s <= 3; ++s) {
        const double d = rp.D[s];
        if (d != 0.0) {
#pragma unroll
          for (int i = 0; i < N; ++i) yt[i] = yt[i] + d * X[((s - 1) * N + i) * 32];
        }
      }
    }
    if (active) {
#pragma unroll
      for (int i = 0; i < N; ++i) ycol[i] = yt[i];
    }
    if (tid == 0) {   // counters of the TLM block (:768, 773, 782): per accepted step Njac 3, Ndec 1, Nsol NTLM
      int* is_ = a.istatus + (size_t)cell * 20;
      is_[Njac] += 3 * (int)done;
      is_[Ndec] += (int)done;
      is_[Nsol] += (int)(info != 0 ? done - 1 : done) * ntlm;
      if (info != 0 && a.ierr && a.ierr[cell] == 1) a.ierr[cell] = -12;
    }
  }
}

}  // namespace fatode
