// Discrete adjoint of the 3-stage implicit Runge-Kutta methods for CBM-IV (ADJ/RK_ADJ INTEGRATE_ADJ, backward sweep).
//
// Replaces rk_DadjInt (ADJ/RK_ADJ/RK_ADJ_f90_Integrator.F90:1229-1524) in its preset mode: a fixed number of transposed
// simplified-Newton sweeps per adjoint column (AdjointSolve = Solve_fixed, ICNTRL(7) = 0/1), with RK_PrepareRHS_Adj
// :1742-1784, RK_SolveTR :1852-1888, RK_Decomp :1788-1813 and the LS_Solver pieces lss_jac1..3, lss_mul_jactr1..3,
// DGETRS('T') / ZGETRS('T') (ADJ/RK_ADJ/LS_Solver.F90:83-113,704-720,835-848).
//
// Two kernels after the taping forward sweep (k_rk_fwd_cell<.., ADJ = true>):
//   k_rk_adj_prep   one THREAD per accepted step of any system: everything of the step that does not depend on the
//                   adjoint column — J(T,Y), the real and the complex factorisation for this step's H (generated
//                   static-pattern DGETF2/ZGETF2 in the thread's shared-memory slots), the reciprocals of the complex
//                   diagonal, the three stage Jacobians J(T + c_i H, Y + Z_i) and the reactant products of JACP at the
//                   three stages — written as one contiguous 17 KB record with full-sector stores.
//   k_rk_adj_cell   one WARP per system, lane = adjoint column (NADJ <= 32).  The warp walks the system's records
//                   backwards, stages each record in shared memory, and every lane runs its column's sweeps: J_i^T products
//                   and the transposed solves in registers (generated jt_vec / lut_solve / zlut_solve read the record with
//                   warp-uniform addresses), the column's vectors Lambda, Mu, U1..3, G1..3, R1..3 in its shared-memory
//                   slots.  The update Lambda += U1 + U2 + U3 cancels up to nine digits for stiff components, which
//                   amplifies any rounding difference of a reordered evaluation beyond the parity bound, so the sweep
//                   performs the reference's operations in the reference's order (jt_vec_seq, lut_solve_seq,
//                   zlut_solve_seq: no fma, true divisions): Lambda is reproduced to the last bit, Mu to rounding.
// The factorisations of the backward sweep are accuracy-, not parity-critical, so the restricted pivot search of the
// generated code is accepted as is (its pivots are the reference's whenever the step stayed on the static pattern).
#pragma once
#include "rk_cell_fwd.cuh"
#include "ros_cell_adj.cuh"
#include "tmem.cuh"

namespace fatode {

// fat record (doubles); every section starts 32-byte aligned
constexpr int RA_HDR = 0;                    // T, H, NewIt | swaps << 8 | zswaps << 16, branch mask of the complex pivots
constexpr int RA_LU = 4;                     // 367 (+1)
constexpr int RA_DINV = RA_LU + 368;         // 32
constexpr int RA_ZR = RA_DINV + 32;          // 367 (+1)
constexpr int RA_ZI = RA_ZR + 368;           // 367 (+1)
constexpr int RA_ZRATIO = RA_ZI + 368;       // 32: pivot-only parts of the Smith divisions (cbm4_cell::zdiag_prep)
constexpr int RA_ZDIV = RA_ZRATIO + 32;      // 32
constexpr int RA_ZRDIV = RA_ZDIV + 32;       // 32
constexpr int RA_J = RA_ZRDIV + 32;          // 3 x 276
constexpr int RA_P = RA_J + 3 * 276;         // 3 x 32 (29 used)
constexpr int RA_REC = RA_P + 3 * 32;        // 2160 doubles = 17 280 B
static_assert(RA_REC % 4 == 0 && cbm4_cell::NLU == 367 && cbm4_cell::NJ == 276, "record layout");

constexpr int RA_PREP_LANES = 22;

struct RkAdjPrepArgs {
  long nrec;                 // records in this batch
  int ncells;                // systems in this batch
  const long long* rec_off;  // [ncells + 1] exclusive prefix of accepted steps
  const int* slot;           // [ncells] tape slot of each system
  const double* tape; int tape_cap;
  double* fat;
};

__global__ void __launch_bounds__(32, 1)
k_rk_adj_prep(RkParams rp, Cbm4Const mc, RkAdjPrepArgs a) {
  extern __shared__ __align__(16) double cell_smem[];
  constexpr int SV = RA_PREP_LANES, N = 32, NLU = cbm4_cell::NLU;
  if (threadIdx.x >= SV) return;
  double* const base = cell_smem + threadIdx.x;
  double* const lu = base;
  double* const dinv = lu + NLU * SV;
  double* const zr = dinv + N * SV;
  double* const zi = zr + NLU * SV;
  double* const Yv = zi + NLU * SV;
  double* const Tv = Yv + N * SV;
  double PH[Cbm4Cell::NPH];
  for (long q = (long)blockIdx.x * SV + threadIdx.x; q < a.nrec; q += (long)gridDim.x * SV) {
    int lo = 0, hi = a.ncells;               // system with rec_off[c] <= q < rec_off[c + 1]
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (a.rec_off[mid] <= q) lo = mid; else hi = mid;
    }
    const long r = q - a.rec_off[lo];
    const double* rec = a.tape + ((size_t)a.slot[lo] * a.tape_cap + r) * RK_REC;
    double* out = a.fat + (size_t)q * RA_REC;
    const double T = rec[0], H = rec[1];
#pragma unroll 4
    for (int i = 0; i < N; ++i) Yv[i * SV] = rec[4 + i];
    unsigned swaps = 0, zswaps = 0;
    Cbm4Cell::photo(T, mc, PH);
    Cbm4Cell::build_E<SV>(Yv, PH, mc, rp.Gamma / H, lu);
    cbm4_cell::lu_factor<SV>(lu, dinv, swaps);
    const double a4 = (double)(float)(rp.Alpha / H), b4 = (double)(float)(rp.Beta / H);
    Cbm4Cell::build_EZ<SV>(Yv, PH, mc, a4, b4, zr, zi);
    cbm4_cell::zlu_factor<SV>(zr, zi, zswaps);
    const unsigned zmask = cbm4_cell::zdiag_prep<SV>(zr, zi, out + RA_ZRATIO, out + RA_ZDIV, out + RA_ZRDIV);
    st_global_256(out + RA_HDR, T, H, (double)((unsigned)rec[2] | (swaps << 8) | (zswaps << 16)), (double)zmask);
#pragma unroll 2
    for (int e = 0; e < 364; e += 4) {
      st_global_256(out + RA_LU + e, lu[e * SV], lu[(e + 1) * SV], lu[(e + 2) * SV], lu[(e + 3) * SV]);
      st_global_256(out + RA_ZR + e, zr[e * SV], zr[(e + 1) * SV], zr[(e + 2) * SV], zr[(e + 3) * SV]);
      st_global_256(out + RA_ZI + e, zi[e * SV], zi[(e + 1) * SV], zi[(e + 2) * SV], zi[(e + 3) * SV]);
    }
    st_global_256(out + RA_LU + 364, lu[364 * SV], lu[365 * SV], lu[366 * SV], 0.0);
    st_global_256(out + RA_ZR + 364, zr[364 * SV], zr[365 * SV], zr[366 * SV], 0.0);
    st_global_256(out + RA_ZI + 364, zi[364 * SV], zi[365 * SV], zi[366 * SV], 0.0);
#pragma unroll 1
    for (int k = 0; k < N; k += 4)
      st_global_256(out + RA_DINV + k, dinv[k * SV], dinv[(k + 1) * SV], dinv[(k + 2) * SV], dinv[(k + 3) * SV]);
#pragma unroll 1
    for (int s = 1; s <= 3; ++s) {
      const double* Z = rec + 4 + s * N;
#pragma unroll 4
      for (int i = 0; i < N; ++i) Tv[i * SV] = Yv[i * SV] + Z[i];          // WADD(Y, Z_s)
      Cbm4Cell::photo(T + rp.C[s] * H, mc, PH);
      cbm4_cell::jac_values<SV>(Tv, mc.fix, PH, out + RA_J + (s - 1) * 276);
      cbm4_cell::arp_values<SV>(Tv, mc.fix, out + RA_P + (s - 1) * 32);
    }
  }
}

struct RkAdjArgs {
  int ncells;                 // systems of this batch (all integrated successfully)
  const long long* rec_off;   // [ncells + 1]
  const int* cell;            // [ncells] index of the system in the caller's arrays
  const double* fat;
  double* lambda; double* mu; int* istatus;
  unsigned long long* queue;
  int nadj, with_mu, np;
  int* ierr = nullptr;        // per-system IERR (only the tangent-linear big-solve kernel writes it)
};

constexpr int RA_VEC = 11;   // per-lane vectors: LAM, MU, U1..3, G1..3, R1..3
constexpr size_t RK_ADJ_SMEM = sizeof(double) * ((size_t)RA_VEC * 32 * 32 + RA_REC);

__global__ void __launch_bounds__(32, 2)
k_rk_adj_cell(RkParams rp, RkAdjArgs a) {
  extern __shared__ __align__(16) double cell_smem[];
  constexpr int N = 32, SV = 32;
  const int lane = threadIdx.x;
  double* const rec = cell_smem;                         // staged record (warp-uniform reads)
  double* const pv = cell_smem + RA_REC + lane;          // this lane's vectors, element stride 32
  double* const LAM = pv;
  double* const MU = pv + 1 * N * SV;
  double* const U1 = pv + 2 * N * SV;
  double* const U2 = pv + 3 * N * SV;
  double* const U3 = pv + 4 * N * SV;
  double* const G1 = pv + 5 * N * SV;
  double* const G2 = pv + 6 * N * SV;
  double* const G3 = pv + 7 * N * SV;
  double* const R1 = pv + 8 * N * SV;
  double* const R2 = pv + 9 * N * SV;
  double* const R3 = pv + 10 * N * SV;
  double* const Us[3] = {U1, U2, U3};
  double* const Gs[3] = {G1, G2, G3};
  double* const Rs[3] = {R1, R2, R3};
  const bool active = lane < a.nadj;
  for (;;) {
    unsigned long long slot = 0;
    if (lane == 0) slot = atomicAdd(a.queue, 1ull);
    slot = __shfl_sync(0xffffffffu, slot, 0);
    if (slot >= (unsigned long long)a.ncells) break;
    const long long r0 = a.rec_off[slot], r1 = a.rec_off[slot + 1];
    const int cell = a.cell[slot];
    // ADJINIT of the example drivers: Lambda = I, Mu = 0
#pragma unroll 4
    for (int i = 0; i < N; ++i) { LAM[i * SV] = (i == lane) ? 1.0 : 0.0; MU[i * SV] = 0.0; }
    long nsol = 0;
    for (long long q = r1 - 1; q >= r0; --q) {
      __syncwarp();
      {   // stage the record: 2128 doubles, 16 B per lane per access
        const double2* src = reinterpret_cast<const double2*>(a.fat + (size_t)q * RA_REC);
        double2* dst = reinterpret_cast<double2*>(rec);
        for (int e = lane; e < RA_REC / 2; e += 32) dst[e] = __ldg(src + e);
      }
      __syncwarp();
      const double H = rec[RA_HDR + 1];
      const unsigned hw = (unsigned)rec[RA_HDR + 2];
      const int NewIt = (int)(hw & 0xffu);
      const unsigned swaps = (hw >> 8) & 0xffu, zswaps = hw >> 16;
      const unsigned zmask = (unsigned)rec[RA_HDR + 3];
      const double rH = 1.0 / H;
      const int sweeps = min(rp.newton_maxit, max(NewIt + 1, 6) + 1);
      nsol += 2L * sweeps * a.nadj;
      if (!active) continue;
      double x[32], o[32];
      // G_s = -H b_s J_s^T Lambda   (:1337-1347)
#pragma unroll
      for (int i = 0; i < N; ++i) x[i] = LAM[i * SV];
#pragma unroll 1
      for (int s = 0; s < 3; ++s) {
        cbm4_cell::jt_vec_seq(rec + RA_J + s * 276, x, o);
        const double c = -H * rp.B[s + 1];
        double* G = Gs[s];
        double* U = Us[s];
#pragma unroll
        for (int i = 0; i < N; ++i) { G[i * SV] = c * o[i]; U[i * SV] = 0.0; }
      }
#pragma unroll 1
      for (int it = 0; it < sweeps; ++it) {
        // RK_PrepareRHS_Adj: R_s = G_s + (U_s + J_s^T (-H sum_j a_js U_j))
#pragma unroll 1
        for (int s = 0; s < 3; ++s) {
          const double c1 = -H * rp.A[1][s + 1], c2 = -H * rp.A[2][s + 1], c3 = -H * rp.A[3][s + 1];
#pragma unroll
          for (int i = 0; i < N; ++i) x[i] = (c1 * U1[i * SV] + c2 * U2[i * SV]) + c3 * U3[i * SV];
          cbm4_cell::jt_vec_seq(rec + RA_J + s * 276, x, o);
          double* R = Rs[s];
          const double* G = Gs[s];
          const double* U = Us[s];
#pragma unroll
          for (int i = 0; i < N; ++i) R[i * SV] = G[i * SV] + (U[i * SV] + o[i]);
        }
        // RK_SolveTR
#pragma unroll 4
        for (int i = 0; i < N; ++i) {
          const double x1 = cbm4_cell::exact_div(R1[i * SV], H, rH), x2 = cbm4_cell::exact_div(R2[i * SV], H, rH),
                       x3 = cbm4_cell::exact_div(R3[i * SV], H, rH);
          R1[i * SV] = rp.AinvT[0][0] * x1 + rp.AinvT[1][0] * x2 + rp.AinvT[2][0] * x3;
          R2[i * SV] = rp.AinvT[0][1] * x1 + rp.AinvT[1][1] * x2 + rp.AinvT[2][1] * x3;
          R3[i * SV] = rp.AinvT[0][2] * x1 + rp.AinvT[1][2] * x2 + rp.AinvT[2][2] * x3;
        }
#pragma unroll
        for (int i = 0; i < N; ++i) x[i] = R1[i * SV];
        cbm4_cell::lut_solve_seq(rec + RA_LU, rec + RA_DINV, swaps, x);
#pragma unroll
        for (int i = 0; i < N; ++i) { R1[i * SV] = x[i]; x[i] = R2[i * SV]; o[i] = -R3[i * SV]; }
        cbm4_cell::zlut_solve_seq(rec + RA_ZR, rec + RA_ZI, rec + RA_ZRATIO, rec + RA_ZDIV, rec + RA_ZRDIV, zmask, zswaps, x, o);
#pragma unroll 4
        for (int i = 0; i < N; ++i) {
          const double x1 = R1[i * SV], x2 = x[i], x3 = -o[i];
          const double d1 = rp.Tinv[0][0] * x1 + rp.Tinv[1][0] * x2 + rp.Tinv[2][0] * x3;
          const double d2 = rp.Tinv[0][1] * x1 + rp.Tinv[1][1] * x2 + rp.Tinv[2][1] * x3;
          const double d3 = rp.Tinv[0][2] * x1 + rp.Tinv[1][2] * x2 + rp.Tinv[2][2] * x3;
          U1[i * SV] -= d1; U2[i * SV] -= d2; U3[i * SV] -= d3;
        }
      }
      // Mu += sum_s JACP_s^T (H sum_j a_js U_j + H b_s Lambda)   (:1424-1464)
      if (a.with_mu) {
#pragma unroll 1
        for (int s = 0; s < 3; ++s) {
          const double c1 = H * rp.A[1][s + 1], c2 = H * rp.A[2][s + 1], c3 = H * rp.A[3][s + 1], cb = H * rp.B[s + 1];
#pragma unroll
          for (int i = 0; i < N; ++i) x[i] = ((c1 * U1[i * SV] + c2 * U2[i * SV]) + c3 * U3[i * SV]) + cb * LAM[i * SV];
          const double* AP = rec + RA_P + s * 32;
          Cbm4Warp::stoich_cols(x, [&](auto pc, double sp) {
            constexpr int p = decltype(pc)::value;
            MU[p * SV] = fma(AP[p], sp, MU[p * SV]);
          });
        }
      }
#pragma unroll 4
      for (int i = 0; i < N; ++i) LAM[i * SV] = ((LAM[i * SV] + U1[i * SV]) + U2[i * SV]) + U3[i * SV];
    }
    // results: Lambda(N,NADJ), Mu(NP,NADJ) column-major per system
    if (active) {
      double* lo = a.lambda + ((size_t)cell * a.nadj + lane) * N;
#pragma unroll 4
      for (int i = 0; i < N; ++i) lo[i] = LAM[i * SV];
      if (a.with_mu) {
        double* mo = a.mu + ((size_t)cell * a.nadj + lane) * a.np;
        for (int p = 0; p < a.np; ++p) mo[p] = MU[p * SV];
      }
    }
    if (lane == 0) {   // backward counters (:1290-1299, 1511, RK_Decomp, RK_SolveTR)
      const int A = (int)(r1 - r0);
      int* is_ = a.istatus + (size_t)cell * 20;
      is_[Njac] += 4 * A;
      is_[Ndec] += A;
      is_[Nstp] += A;
      is_[Nsol] += (int)nsol;
    }
  }
}


// ---- tensor-memory variant of the backward sweep (the shipped one for the static-pattern path): Lambda, Mu, U1..3 and G1..3
// (8 x 32 = 256 doubles, exactly a lane's share of tensor memory) live in tensor memory, the right-hand sides R1..3 and the
// staged record in shared memory (41 KB per warp): one CTA of four sweep warps per SM instead of two one-warp CTAs.
// Same floating-point operations in the same order as k_rk_adj_cell (sums accumulated term by term): Lambda stays bit-identical.
constexpr uint32_t RA_TM_LAM = 0, RA_TM_MU = 64, RA_TM_U = 128, RA_TM_G = 320;   // 32-bit columns; U_s / G_s at +64 s
constexpr int RA_TM_WARPS = 4;
constexpr int RA_TM_WARP_DOUBLES = RA_REC + 3 * 32 * 32;
constexpr size_t RK_ADJ_TM_SMEM = sizeof(double) * (size_t)RA_TM_WARPS * RA_TM_WARP_DOUBLES;

__global__ void __launch_bounds__(RA_TM_WARPS * 32, 1)
k_rk_adj_cell_tm(RkParams rp, RkAdjArgs a) {
  extern __shared__ __align__(16) double cell_smem[];
  __shared__ uint32_t tmem_base_s;
  constexpr int N = 32, SV = 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* const rec = cell_smem + (size_t)warp * RA_TM_WARP_DOUBLES;
  double* const pv = rec + RA_REC + lane;
  double* const R1 = pv;
  double* const R2 = pv + 1 * N * SV;
  double* const R3 = pv + 2 * N * SV;
  double* const Rs[3] = {R1, R2, R3};
  if (warp == 0) tmem::alloc512(&tmem_base_s);
  tmem::fence_before_sync();
  __syncthreads();
  tmem::fence_after_sync();
  const uint32_t tm = tmem_base_s + ((uint32_t)(warp * 32) << 16);
  const bool active = lane < a.nadj;
  for (;;) {
    unsigned long long slot = 0;
    if (lane == 0) slot = atomicAdd(a.queue, 1ull);
    slot = __shfl_sync(0xffffffffu, slot, 0);
    if (slot >= (unsigned long long)a.ncells) break;
    const long long r0 = a.rec_off[slot], r1 = a.rec_off[slot + 1];
    const int cell = a.cell[slot];
    double x[32], o[32];
#pragma unroll
    for (int i = 0; i < N; ++i) { x[i] = (i == lane) ? 1.0 : 0.0; o[i] = 0.0; }
    tmem::st32(tm + RA_TM_LAM, x);
    tmem::st32(tm + RA_TM_MU, o);
    tmem::wait_st();
    long nsol = 0;
    for (long long q = r1 - 1; q >= r0; --q) {
      __syncwarp();
      {
        const double2* src = reinterpret_cast<const double2*>(a.fat + (size_t)q * RA_REC);
        double2* dst = reinterpret_cast<double2*>(rec);
        for (int e = lane; e < RA_REC / 2; e += 32) dst[e] = __ldg(src + e);
      }
      __syncwarp();
      const double H = rec[RA_HDR + 1];
      const unsigned hw = (unsigned)rec[RA_HDR + 2];
      const int NewIt = (int)(hw & 0xffu);
      const unsigned swaps = (hw >> 8) & 0xffu, zswaps = hw >> 16;
      const unsigned zmask = (unsigned)rec[RA_HDR + 3];
      const double rH = 1.0 / H;
      const int sweeps = min(rp.newton_maxit, max(NewIt + 1, 6) + 1);
      nsol += 2L * sweeps * a.nadj;
      // G_s = -H b_s J_s^T Lambda, U_s = 0   (:1337-1347)
      tmem::ld32(tm + RA_TM_LAM, x);
#pragma unroll 1
      for (int s = 0; s < 3; ++s) {
        cbm4_cell::jt_vec_seq(rec + RA_J + s * 276, x, o);
        const double c = -H * rp.B[s + 1];
#pragma unroll
        for (int i = 0; i < N; ++i) o[i] = c * o[i];
        tmem::st32(tm + RA_TM_G + 64 * s, o);
#pragma unroll
        for (int i = 0; i < N; ++i) o[i] = 0.0;
        tmem::st32(tm + RA_TM_U + 64 * s, o);
      }
      tmem::wait_st();
#pragma unroll 1
      for (int it = 0; it < sweeps; ++it) {
        // RK_PrepareRHS_Adj: R_s = G_s + (U_s + J_s^T (-H sum_j a_js U_j))
#pragma unroll 1
        for (int s = 0; s < 3; ++s) {
          const double c1 = -H * rp.A[1][s + 1], c2 = -H * rp.A[2][s + 1], c3 = -H * rp.A[3][s + 1];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            double t[16];
            tmem::ld16(tm + RA_TM_U + 32 * h, t);
#pragma unroll
            for (int i = 0; i < 16; ++i) x[16 * h + i] = c1 * t[i];
            tmem::ld16(tm + RA_TM_U + 64 + 32 * h, t);
#pragma unroll
            for (int i = 0; i < 16; ++i) x[16 * h + i] = x[16 * h + i] + c2 * t[i];
            tmem::ld16(tm + RA_TM_U + 128 + 32 * h, t);
#pragma unroll
            for (int i = 0; i < 16; ++i) x[16 * h + i] = x[16 * h + i] + c3 * t[i];
          }
          cbm4_cell::jt_vec_seq(rec + RA_J + s * 276, x, o);
          double* R = Rs[s];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            double g[16], u[16];
            tmem::ld16(tm + RA_TM_G + 64 * s + 32 * h, g);
            tmem::ld16(tm + RA_TM_U + 64 * s + 32 * h, u);
#pragma unroll
            for (int i = 0; i < 16; ++i) R[(16 * h + i) * SV] = g[i] + (u[i] + o[16 * h + i]);
          }
        }
        // RK_SolveTR
#pragma unroll 4
        for (int i = 0; i < N; ++i) {
          const double x1 = cbm4_cell::exact_div(R1[i * SV], H, rH), x2 = cbm4_cell::exact_div(R2[i * SV], H, rH),
                       x3 = cbm4_cell::exact_div(R3[i * SV], H, rH);
          R1[i * SV] = rp.AinvT[0][0] * x1 + rp.AinvT[1][0] * x2 + rp.AinvT[2][0] * x3;
          R2[i * SV] = rp.AinvT[0][1] * x1 + rp.AinvT[1][1] * x2 + rp.AinvT[2][1] * x3;
          R3[i * SV] = rp.AinvT[0][2] * x1 + rp.AinvT[1][2] * x2 + rp.AinvT[2][2] * x3;
        }
#pragma unroll
        for (int i = 0; i < N; ++i) x[i] = R1[i * SV];
        cbm4_cell::lut_solve_seq(rec + RA_LU, rec + RA_DINV, swaps, x);
#pragma unroll
        for (int i = 0; i < N; ++i) { R1[i * SV] = x[i]; x[i] = R2[i * SV]; o[i] = -R3[i * SV]; }
        cbm4_cell::zlut_solve_seq(rec + RA_ZR, rec + RA_ZI, rec + RA_ZRATIO, rec + RA_ZDIV, rec + RA_ZRDIV, zmask, zswaps, x, o);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          double u1[16], u2[16], u3[16];
          tmem::ld16(tm + RA_TM_U + 32 * h, u1);
          tmem::ld16(tm + RA_TM_U + 64 + 32 * h, u2);
          tmem::ld16(tm + RA_TM_U + 128 + 32 * h, u3);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int e = 16 * h + i;
            const double x1 = R1[e * SV], x2 = x[e], x3 = -o[e];
            const double d1 = rp.Tinv[0][0] * x1 + rp.Tinv[1][0] * x2 + rp.Tinv[2][0] * x3;
            const double d2 = rp.Tinv[0][1] * x1 + rp.Tinv[1][1] * x2 + rp.Tinv[2][1] * x3;
            const double d3 = rp.Tinv[0][2] * x1 + rp.Tinv[1][2] * x2 + rp.Tinv[2][2] * x3;
            u1[i] -= d1; u2[i] -= d2; u3[i] -= d3;
          }
          tmem::st16(tm + RA_TM_U + 32 * h, u1);
          tmem::st16(tm + RA_TM_U + 64 + 32 * h, u2);
          tmem::st16(tm + RA_TM_U + 128 + 32 * h, u3);
        }
        tmem::wait_st();
      }
      // Mu += sum_s JACP_s^T (H sum_j a_js U_j + H b_s Lambda)   (:1424-1464)
      if (a.with_mu) {
        tmem::ld32(tm + RA_TM_MU, o);
#pragma unroll 1
        for (int s = 0; s < 3; ++s) {
          const double c1 = H * rp.A[1][s + 1], c2 = H * rp.A[2][s + 1], c3 = H * rp.A[3][s + 1], cb = H * rp.B[s + 1];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            double t[16];
            tmem::ld16(tm + RA_TM_U + 32 * h, t);
#pragma unroll
            for (int i = 0; i < 16; ++i) x[16 * h + i] = c1 * t[i];
            tmem::ld16(tm + RA_TM_U + 64 + 32 * h, t);
#pragma unroll
            for (int i = 0; i < 16; ++i) x[16 * h + i] = x[16 * h + i] + c2 * t[i];
            tmem::ld16(tm + RA_TM_U + 128 + 32 * h, t);
#pragma unroll
            for (int i = 0; i < 16; ++i) x[16 * h + i] = x[16 * h + i] + c3 * t[i];
            tmem::ld16(tm + RA_TM_LAM + 32 * h, t);
#pragma unroll
            for (int i = 0; i < 16; ++i) x[16 * h + i] = x[16 * h + i] + cb * t[i];
          }
          const double* AP = rec + RA_P + s * 32;
          Cbm4Warp::stoich_cols(x, [&](auto pc, double sp) {
            constexpr int p = decltype(pc)::value;
            o[p] = fma(AP[p], sp, o[p]);
          });
        }
        tmem::st32(tm + RA_TM_MU, o);
      }
      tmem::ld32(tm + RA_TM_LAM, x);
#pragma unroll 1
      for (int s = 0; s < 3; ++s) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          double t[16];
          tmem::ld16(tm + RA_TM_U + 64 * s + 32 * h, t);
#pragma unroll
          for (int i = 0; i < 16; ++i) x[16 * h + i] = x[16 * h + i] + t[i];
        }
      }
      tmem::st32(tm + RA_TM_LAM, x);
      tmem::wait_st();
    }
    tmem::ld32(tm + RA_TM_LAM, x);
    tmem::ld32(tm + RA_TM_MU, o);
    if (active) {
      double* lo = a.lambda + ((size_t)cell * a.nadj + lane) * N;
#pragma unroll
      for (int i = 0; i < N; ++i) lo[i] = x[i];
      if (a.with_mu) {
        double* mo = a.mu + ((size_t)cell * a.nadj + lane) * a.np;
#pragma unroll
        for (int p = 0; p < 32; ++p) if (p < a.np) mo[p] = o[p];
      }
    }
    if (lane == 0) {   // backward counters (:1290-1299, 1511, RK_Decomp, RK_SolveTR)
      const int A = (int)(r1 - r0);
      int* is_ = a.istatus + (size_t)cell * 20;
      is_[Njac] += 4 * A;
      is_[Ndec] += A;
      is_[Nstp] += A;
      is_[Nsol] += (int)nsol;
    }
  }
  tmem::fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem::dealloc512(tmem_base_s);
}

}  // namespace fatode
