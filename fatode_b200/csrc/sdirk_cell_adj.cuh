// Discrete adjoint of the SDIRK family for CBM-IV (ADJ/SDIRK_ADJ INTEGRATE_ADJ, backward sweep).
//
// Replaces SDIRK_DadjInt (ADJ/SDIRK_ADJ/SDIRK_ADJ_f90_Integrator.F90:786-995; Lambda-only twin :2300-2489) in the reference's
// preset mode, the direct adjoint solve (ICNTRL(7) = 1): per popped step and per stage i = S..1 one factorisation of
// 1/(h gamma) I - J(T + c_i h, Y + Z_i) (LSS_Jac + LSS_Decomp :859-878), then per adjoint column
//   G = h J_i^T ( b_i Lambda + sum_{j>i} a_ji U_j ),  U_i = (E_i^T)^-1 G / (h gamma)            (:873-889)
//   V_i = JACP_i^T ( h sum_{j>=i} a_ji U_j + h b_i Lambda )                                       (:961-969)
// and after the stages Lambda += sum_i U_i, Mu += sum_i V_i (:975-983).
//
// Two kernels after the taping forward sweep (k_sdirk_fwd_cell<.., SD_MODE_ADJ>):
//   k_sdirk_adj_prep   one THREAD per (accepted step, stage): the stage Jacobian, its factorisation and the reactant products of
//                      JACP at the stage state — everything that does not depend on the adjoint column — written into the
//                      step's contiguous record with full-sector stores.
//   k_sdirk_adj_cell   one WARP per system, lane = adjoint column (NADJ <= 32), walking the system's records backwards; each
//                      stage's 5.4 KB chunk is staged in shared memory (warp-uniform reads), the lane's vectors Lambda, Mu,
//                      U_1..U_S in its shared-memory slots, the working vectors in registers.
// As in the Runge-Kutta adjoint (rk_cell_adj.cuh) Lambda += sum U_i cancels for stiff components, so the sweep performs the
// reference's operations in the reference's order (jt_vec_seq, lut_solve_seq, no fma): Lambda is reproduced to the last bit,
// Mu to rounding.  The step's own SDIRK_PrepareMatrix (:838-845) only contributes counters in this mode (its factors are
// overwritten by the first stage's) and is not recomputed.
#pragma once
#include "sdirk_cell_fwd.cuh"
#include "rk_cell_adj.cuh"
#include "tmem.cuh"
#include "gen/sdirk_adj_tableau_gen.h"

namespace fatode {

constexpr int SA_HDR = 0;                    // T, H, S, pad
constexpr int SA_P = 4;                      // 5 x 32 (29 used): reactant products of JACP per stage
constexpr int SA_STAGE = SA_P + 5 * 32;      // per stage: LU 367 (+1), DINV 32, J 276, swaps (+3 pad)
constexpr int SA_LU = 0, SA_DINV = 368, SA_J = 400, SA_SW = 676;
constexpr int SA_CHUNK = 680;
constexpr int SA_REC = SA_STAGE + 5 * SA_CHUNK;   // 3564 doubles = 28 512 B
static_assert(SA_REC % 4 == 0 && SA_CHUNK % 4 == 0 && cbm4_cell::NLU == 367 && cbm4_cell::NJ == 276, "record layout");
// 30, not 32: every preparation kernel keeps its own instantiation of the generated per-thread routines (ptxas 12.9 narrowed the
// 32-byte stores of jac_values<32> in k_sdirk_tlm_prep once a second kernel shared that instantiation; tests/test_sass_guards.py)
constexpr int SA_PREP_LANES = 30;


struct SdirkAdjPrepArgs {
  long nrec;                 // records (accepted steps) in this batch
  int ncells;
  const long long* rec_off;  // [ncells + 1]
  const int* slot;           // [ncells] tape slot of each system
  const double* tape; int tape_cap;
  double* fat;
};

__global__ void __launch_bounds__(32, 1)
k_sdirk_adj_prep(SdirkParams sp, Cbm4Const mc, SdirkAdjPrepArgs a) {
  extern __shared__ __align__(16) double cell_smem[];
  constexpr int SV = SA_PREP_LANES, N = 32, NLU = cbm4_cell::NLU;
  if (threadIdx.x >= SV) return;
  double* const base = cell_smem + threadIdx.x;
  double* const lu = base;
  double* const dinv = lu + NLU * SV;
  double* const Tv = dinv + N * SV;
  double PH[Cbm4Cell::NPH];
  const int S = sp.S;
  const long nwork = a.nrec * S;
  for (long w = (long)blockIdx.x * SV + threadIdx.x; w < nwork; w += (long)gridDim.x * SV) {
    const long q = w / S;
    const int s = (int)(w - q * S);
    int lo = 0, hi = a.ncells;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (a.rec_off[mid] <= q) lo = mid; else hi = mid;
    }
    const long r = q - a.rec_off[lo];
    const double* rec = a.tape + ((size_t)a.slot[lo] * a.tape_cap + r) * SD_REC;
    double* out = a.fat + (size_t)q * SA_REC;
    const double T = rec[0], H = rec[1];
    const double* Z = rec + 4 + N + s * N;
#pragma unroll 4
    for (int i = 0; i < N; ++i) Tv[i * SV] = rec[4 + i] + Z[i];            // TMP = Y + Z_i (:853)
    if (s == 0) st_global_256(out + SA_HDR, T, H, (double)S, 0.0);
    double* ch = out + SA_STAGE + s * SA_CHUNK;
    unsigned swaps = 0;
    Cbm4Cell::photo(T + sp.C[s] * H, mc, PH);
    cbm4_cell::jac_values<SV>(Tv, mc.fix, PH, ch + SA_J);
    cbm4_cell::arp_values<SV>(Tv, mc.fix, out + SA_P + s * 32);
    Cbm4Cell::build_E<SV>(Tv, PH, mc, 1.0 / (H * sp.Gamma), lu);
    cbm4_cell::lu_factor<SV>(lu, dinv, swaps);
#pragma unroll 2
    for (int e = 0; e < 364; e += 4)
      st_global_256(ch + SA_LU + e, lu[e * SV], lu[(e + 1) * SV], lu[(e + 2) * SV], lu[(e + 3) * SV]);
    st_global_256(ch + SA_LU + 364, lu[364 * SV], lu[365 * SV], lu[366 * SV], 0.0);
#pragma unroll 1
    for (int k = 0; k < N; k += 4)
      st_global_256(ch + SA_DINV + k, dinv[k * SV], dinv[(k + 1) * SV], dinv[(k + 2) * SV], dinv[(k + 3) * SV]);
    st_global_256(ch + SA_SW, (double)swaps, 0.0, 0.0, 0.0);
  }
}

struct SdirkAdjArgs {
  int ncells;
  const long long* rec_off;
  const int* cell;
  const double* fat;
  double* lambda; double* mu; int* istatus; int* ierr;
  unsigned long long* queue;
  int nadj, with_mu, np;
};

constexpr int SA_VEC = 7;   // per-lane vectors: LAM, MU, U_1..U_5
constexpr size_t SDIRK_ADJ_SMEM = sizeof(double) * ((size_t)SA_VEC * 32 * 32 + SA_STAGE + SA_CHUNK);

__global__ void __launch_bounds__(32, 3)
k_sdirk_adj_cell(SdirkParams sp, SdirkAdjCoef co, SdirkAdjArgs a) {
  extern __shared__ __align__(16) double cell_smem[];
  constexpr int N = 32, SV = 32;
  const int lane = threadIdx.x;
  double* const hdr = cell_smem;                         // header + reactant products of the step (warp-uniform reads)
  double* const ch = cell_smem + SA_STAGE;               // the current stage's chunk
  double* const pv = cell_smem + SA_STAGE + SA_CHUNK + lane;
  double* const LAM = pv;
  double* const MU = pv + 1 * N * SV;
  double* const U = pv + 2 * N * SV;                     // U[(s * N + i) * SV]
  const bool active = lane < a.nadj;
  const int S = sp.S;
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
    for (long long q = r1 - 1; q >= r0; --q) {
      const double* rec = a.fat + (size_t)q * SA_REC;
      __syncwarp();
      {
        const double2* src = reinterpret_cast<const double2*>(rec);
        double2* dst = reinterpret_cast<double2*>(hdr);
        for (int e = lane; e < SA_STAGE / 2; e += 32) dst[e] = __ldg(src + e);
      }
      double H = 0.0, hgi = 0.0;
      double x[32], o[32];
#pragma unroll 1
      for (int s = S - 1; s >= 0; --s) {
        __syncwarp();
        {
          const double2* src = reinterpret_cast<const double2*>(rec + SA_STAGE + s * SA_CHUNK);
          double2* dst = reinterpret_cast<double2*>(ch);
          for (int e = lane; e < SA_CHUNK / 2; e += 32) dst[e] = __ldg(src + e);
        }
        __syncwarp();
        if (!active) continue;
        H = hdr[SA_HDR + 1];
        hgi = 1.0 / (H * sp.Gamma);
        const unsigned swaps = (unsigned)ch[SA_SW];
        // G = b_i Lambda + sum_{j>i} a_ji U_j   (:873-878; DAXPY with a zero factor returns at once)
        {
          const double b = co.B[s];
#pragma unroll
          for (int i = 0; i < N; ++i) x[i] = b * LAM[i * SV];
        }
        for (int j = s + 1; j < S; ++j) {
          const double c = co.A[j][s];
          if (c != 0.0) {
            const double* Uj = U + j * N * SV;
#pragma unroll
            for (int i = 0; i < N; ++i) x[i] = x[i] + c * Uj[i * SV];
          }
        }
        asm volatile("" ::: "memory");
        cbm4_cell::jt_vec_seq(ch + SA_J, x, o);            // TMP = J_i^T G
#pragma unroll
        for (int i = 0; i < N; ++i) o[i] = hgi * (H * o[i]);   // G = H TMP, then SDIRK_Solve's scaling by 1/(H gamma)
        asm volatile("" ::: "memory");
        cbm4_cell::lut_solve_seq(ch + SA_LU, ch + SA_DINV, swaps, o);
        double* Us = U + s * N * SV;
#pragma unroll
        for (int i = 0; i < N; ++i) Us[i * SV] = o[i];
      }
      if (!active) continue;
      // Mu += V_1 + .. + V_S with V_i = JACP_i^T ( sum_{j>=i} (H a_ji) U_j + (H b_i) Lambda )   (:961-983)
      if (a.with_mu) {
#pragma unroll 1
        for (int s = 0; s < S; ++s) {
#pragma unroll
          for (int i = 0; i < N; ++i) x[i] = 0.0;
          for (int j = s; j < S; ++j) {
            const double c = H * co.A[j][s];
            if (c != 0.0) {
              const double* Uj = U + j * N * SV;
#pragma unroll
              for (int i = 0; i < N; ++i) x[i] = x[i] + c * Uj[i * SV];
            }
          }
          {
            const double c = H * co.B[s];
            if (c != 0.0) {
#pragma unroll
              for (int i = 0; i < N; ++i) x[i] = x[i] + c * LAM[i * SV];
            }
          }
          const double* AP = hdr + SA_P + s * 32;
          Cbm4Warp::stoich_cols(x, [&](auto pc, double spv) {
            constexpr int p = decltype(pc)::value;
            MU[p * SV] = fma(AP[p], spv, MU[p * SV]);
          });
        }
      }
      for (int s = 0; s < S; ++s) {
        const double* Us = U + s * N * SV;
#pragma unroll 4
        for (int i = 0; i < N; ++i) LAM[i * SV] = LAM[i * SV] + Us[i * SV];
      }
    }
    if (active) {
      double* lo = a.lambda + ((size_t)cell * a.nadj + lane) * N;
#pragma unroll 4
      for (int i = 0; i < N; ++i) lo[i] = LAM[i * SV];
      if (a.with_mu) {
        double* mo = a.mu + ((size_t)cell * a.nadj + lane) * a.np;
        for (int p = 0; p < a.np; ++p) mo[p] = MU[p * SV];
      }
    }
    if (lane == 0) {   // backward counters (:838-847, 855, 866, SDIRK_Solve); the return code is the backward sweep's (:519-531)
      const int A = (int)(r1 - r0);
      int* is_ = a.istatus + (size_t)cell * 20;
      is_[Nstp] += A;
      is_[Njac] += (1 + S) * A;
      is_[Ndec] += (1 + S) * A;
      is_[Nsol] += S * a.nadj * A;
      a.ierr[cell] = 1;
    }
  }
}


// ---- tensor-memory variant of the backward sweep (the shipped one for the static-pattern path).
// The lane's seven vectors (Lambda, Mu, U_1..U_5 = 224 doubles) fit the 256 doubles of tensor memory a lane owns, the way the
// Rosenbrock sweeps keep their running sums (tmem.cuh): shared memory then only stages the records (6.75 KB per warp), and an
// SM runs FOUR sweep warps — one CTA of four warps, each with its quarter of the tensor memory and its own system — instead
// of three.  tcgen05.ld/st are warp-wide, so idle lanes (column >= NADJ) run along and are dropped at the output.  Same
// floating-point operations in the same order as k_sdirk_adj_cell: Lambda stays bit-identical.
constexpr uint32_t SA_TM_LAM = 0, SA_TM_MU = 64, SA_TM_U = 128;       // 32-bit columns; U_s at SA_TM_U + 64 s
constexpr int SA_TM_WARPS = 4;
constexpr size_t SDIRK_ADJ_TM_SMEM = sizeof(double) * (size_t)SA_TM_WARPS * (SA_STAGE + SA_CHUNK);

__global__ void __launch_bounds__(SA_TM_WARPS * 32, 1)
k_sdirk_adj_cell_tm(SdirkParams sp, SdirkAdjCoef co, SdirkAdjArgs a) {
  extern __shared__ __align__(16) double cell_smem[];
  __shared__ uint32_t tmem_base_s;
  constexpr int N = 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* const hdr = cell_smem + (size_t)warp * (SA_STAGE + SA_CHUNK);
  double* const ch = hdr + SA_STAGE;
  if (warp == 0) tmem::alloc512(&tmem_base_s);
  tmem::fence_before_sync();
  __syncthreads();
  tmem::fence_after_sync();
  const uint32_t tm = tmem_base_s + ((uint32_t)(warp * 32) << 16);
  const bool active = lane < a.nadj;
  const int S = sp.S;
  for (;;) {
    unsigned long long slot = 0;
    if (lane == 0) slot = atomicAdd(a.queue, 1ull);
    slot = __shfl_sync(0xffffffffu, slot, 0);
    if (slot >= (unsigned long long)a.ncells) break;
    const long long r0 = a.rec_off[slot], r1 = a.rec_off[slot + 1];
    const int cell = a.cell[slot];
    double x[32], o[32];
    // ADJINIT of the example drivers: Lambda = I, Mu = 0
#pragma unroll
    for (int i = 0; i < N; ++i) { x[i] = (i == lane) ? 1.0 : 0.0; o[i] = 0.0; }
    tmem::st32(tm + SA_TM_LAM, x);
    tmem::st32(tm + SA_TM_MU, o);
    tmem::wait_st();
    for (long long q = r1 - 1; q >= r0; --q) {
      const double* rec = a.fat + (size_t)q * SA_REC;
      __syncwarp();
      {
        const double2* src = reinterpret_cast<const double2*>(rec);
        double2* dst = reinterpret_cast<double2*>(hdr);
        for (int e = lane; e < SA_STAGE / 2; e += 32) dst[e] = __ldg(src + e);
      }
      double H = 0.0;
#pragma unroll 1
      for (int s = S - 1; s >= 0; --s) {
        __syncwarp();
        {
          const double2* src = reinterpret_cast<const double2*>(rec + SA_STAGE + s * SA_CHUNK);
          double2* dst = reinterpret_cast<double2*>(ch);
          for (int e = lane; e < SA_CHUNK / 2; e += 32) dst[e] = __ldg(src + e);
        }
        __syncwarp();
        H = hdr[SA_HDR + 1];
        const double hgi = 1.0 / (H * sp.Gamma);
        const unsigned swaps = (unsigned)ch[SA_SW];
        tmem::ld32(tm + SA_TM_LAM, x);
        {
          const double b = co.B[s];
#pragma unroll
          for (int i = 0; i < N; ++i) x[i] = b * x[i];
        }
        for (int j = s + 1; j < S; ++j) {
          const double c = co.A[j][s];
          if (c != 0.0) {
            double t[16];
            tmem::ld16(tm + SA_TM_U + 64 * j, t);
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = x[i] + c * t[i];
            tmem::ld16(tm + SA_TM_U + 64 * j + 32, t);
#pragma unroll
            for (int i = 0; i < 16; ++i) x[16 + i] = x[16 + i] + c * t[i];
          }
        }
        asm volatile("" ::: "memory");
        cbm4_cell::jt_vec_seq(ch + SA_J, x, o);
#pragma unroll
        for (int i = 0; i < N; ++i) o[i] = hgi * (H * o[i]);
        asm volatile("" ::: "memory");
        cbm4_cell::lut_solve_seq(ch + SA_LU, ch + SA_DINV, swaps, o);
        tmem::st32(tm + SA_TM_U + 64 * s, o);
        tmem::wait_st();
      }
      // Mu += V_1 + .. + V_S   (:961-983)
      if (a.with_mu) {
        tmem::ld32(tm + SA_TM_MU, o);
#pragma unroll 1
        for (int s = 0; s < S; ++s) {
#pragma unroll
          for (int i = 0; i < N; ++i) x[i] = 0.0;
          for (int j = s; j < S; ++j) {
            const double c = H * co.A[j][s];
            if (c != 0.0) {
              double t[16];
              tmem::ld16(tm + SA_TM_U + 64 * j, t);
#pragma unroll
              for (int i = 0; i < 16; ++i) x[i] = x[i] + c * t[i];
              tmem::ld16(tm + SA_TM_U + 64 * j + 32, t);
#pragma unroll
              for (int i = 0; i < 16; ++i) x[16 + i] = x[16 + i] + c * t[i];
            }
          }
          {
            const double c = H * co.B[s];
            if (c != 0.0) {
              double t[16];
              tmem::ld16(tm + SA_TM_LAM, t);
#pragma unroll
              for (int i = 0; i < 16; ++i) x[i] = x[i] + c * t[i];
              tmem::ld16(tm + SA_TM_LAM + 32, t);
#pragma unroll
              for (int i = 0; i < 16; ++i) x[16 + i] = x[16 + i] + c * t[i];
            }
          }
          const double* AP = hdr + SA_P + s * 32;
          Cbm4Warp::stoich_cols(x, [&](auto pc, double spv) {
            constexpr int p = decltype(pc)::value;
            o[p] = fma(AP[p], spv, o[p]);
          });
        }
        tmem::st32(tm + SA_TM_MU, o);
      }
      tmem::ld32(tm + SA_TM_LAM, x);
      for (int s = 0; s < S; ++s) {
        double t[16];
        tmem::ld16(tm + SA_TM_U + 64 * s, t);
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = x[i] + t[i];
        tmem::ld16(tm + SA_TM_U + 64 * s + 32, t);
#pragma unroll
        for (int i = 0; i < 16; ++i) x[16 + i] = x[16 + i] + t[i];
      }
      tmem::st32(tm + SA_TM_LAM, x);
      tmem::wait_st();
    }
    tmem::ld32(tm + SA_TM_LAM, x);
    tmem::ld32(tm + SA_TM_MU, o);
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
    if (lane == 0) {
      const int A = (int)(r1 - r0);
      int* is_ = a.istatus + (size_t)cell * 20;
      is_[Nstp] += A;
      is_[Njac] += (1 + S) * A;
      is_[Ndec] += (1 + S) * A;
      is_[Nsol] += S * a.nadj * A;
      a.ierr[cell] = 1;
    }
  }
  tmem::fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem::dealloc512(tmem_base_s);
}

}  // namespace fatode
