// Tangent linear model of the SDIRK family for CBM-IV (TLM/SDIRK_TLM INTEGRATE_TLM), column propagation.
//
// Replaces the TLM part of SDIRK_IntegratorTLM (TLM/SDIRK_TLM/SDIRK_TLM_f90_Integrator.F90:699-783, 851-862) in the
// reference's preset mode: modified Newton (ICNTRL(7) = 0) with the iteration count of the state stage plus one
// (saveNiter, :603, :769), no TLM influence on the step control (ICNTRL(9) = ICNTRL(12) = 0).  In that mode the columns
// never feed back into the state, so the state sweep (k_sdirk_fwd_cell<.., TLM = true>) runs first and tapes every
// accepted step; two kernels then propagate the columns:
//   k_sdirk_tlm_prep   one THREAD per (accepted step, part): the factorisation of 1/(H gamma) I - J(T,Y) (what the accepted attempt
//                      used: the LU is never carried across steps in this family and a rejected attempt leaves (T,Y)
//                      unchanged) and the S stage Jacobians J(T + c_i H, Y + Z_i) (LSS_Jac1), one contiguous record.
//   k_sdirk_tlm_cell   one WARP per system, lane = TLM column, walking the records forwards: per stage
//                      G = H gamma J_i Y_tlm + sum_j theta_ij Z_tlm_j and the fixed number of iterations
//                      Z_tlm += (E)^-1 [ (H gamma J_i Z_tlm + G - Z_tlm) / (H gamma) ], then Y_tlm += sum_i d_i Z_tlm_i.
// Like the Runge-Kutta adjoint sweep (rk_cell_adj.cuh) the update cancels for stiff components, so the sweep performs the
// reference's operations in the reference's order (j_vec_seq, lun_solve_seq, no fma): Y_tlm is reproduced to the last bit.
#pragma once
#include "sdirk_cell_fwd.cuh"
#include "rk_cell_adj.cuh"
#include "tmem.cuh"

namespace fatode {

constexpr int ST_HDR = 0;                 // T, H, packed iteration counts, S
constexpr int ST_LU = 4;                  // 367 (+1)
constexpr int ST_DINV = ST_LU + 368;      // 32
constexpr int ST_J = ST_DINV + 32;        // 5 x 276
constexpr int ST_REC = ST_J + 5 * 276;    // 1784 doubles = 14 272 B
static_assert(ST_REC % 4 == 0, "record layout");
constexpr int ST_PREP_LANES = 32;

struct SdirkTlmPrepArgs {
  long nrec;
  int ncells;
  const long long* rec_off;
  const int* slot;
  const double* tape; int tape_cap;
  double* fat;
};

// Work item = (accepted step, part): parts 0..S-1 are the stage Jacobians, part S the factorisation.  (One item per step with
// the stages in a loop was miscompiled by ptxas 12.9 when a second kernel of the same shape joined the translation unit: the
// 32-byte stores inside the looped jac_values call were narrowed to their first element.  tests/test_sass_guards.py watches
// the store widths of the three preparation kernels.)
__global__ void __launch_bounds__(32, 1)
k_sdirk_tlm_prep(SdirkParams sp, Cbm4Const mc, SdirkTlmPrepArgs a) {
  extern __shared__ __align__(16) double cell_smem[];
  constexpr int SV = ST_PREP_LANES, N = 32, NLU = cbm4_cell::NLU;
  double* const base = cell_smem + threadIdx.x;
  double* const lu = base;
  double* const dinv = lu + NLU * SV;
  double* const Tv = dinv + N * SV;
  double PH[Cbm4Cell::NPH];
  const int S = sp.S, P = S + 1;
  const long nwork = a.nrec * P;
  // items 0 .. nrec-1 are the factorisations, the rest the stage Jacobians: the lanes of a warp do the same kind of work
  for (long w = (long)blockIdx.x * SV + threadIdx.x; w < nwork; w += (long)gridDim.x * SV) {
    const long q = (w < a.nrec) ? w : (w - a.nrec) / S;
    const int part = (w < a.nrec) ? S : (int)((w - a.nrec) - q * S);
    int lo = 0, hi = a.ncells;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (a.rec_off[mid] <= q) lo = mid; else hi = mid;
    }
    const long r = q - a.rec_off[lo];
    const double* rec = a.tape + ((size_t)a.slot[lo] * a.tape_cap + r) * SD_REC;
    double* out = a.fat + (size_t)q * ST_REC;
    const double T = rec[0], H = rec[1];
    if (part < S) {
      const double* Z = rec + 4 + N + part * N;
#pragma unroll 4
      for (int i = 0; i < N; ++i) Tv[i * SV] = rec[4 + i] + Z[i];          // TMP = Y + Z_i (:700)
      Cbm4Cell::photo(T + sp.C[part] * H, mc, PH);
      cbm4_cell::jac_values<SV>(Tv, mc.fix, PH, out + ST_J + part * 276);
    } else {
#pragma unroll 4
      for (int i = 0; i < N; ++i) Tv[i * SV] = rec[4 + i];
      unsigned swaps = 0;
      Cbm4Cell::photo(T, mc, PH);
      Cbm4Cell::build_E<SV>(Tv, PH, mc, 1.0 / (H * sp.Gamma), lu);
      cbm4_cell::lu_factor<SV>(lu, dinv, swaps);
      st_global_256(out + ST_HDR, T, H, (double)((unsigned)rec[2] | (swaps << 24)), rec[3]);
#pragma unroll 2
      for (int e = 0; e < 364; e += 4)
        st_global_256(out + ST_LU + e, lu[e * SV], lu[(e + 1) * SV], lu[(e + 2) * SV], lu[(e + 3) * SV]);
      st_global_256(out + ST_LU + 364, lu[364 * SV], lu[365 * SV], lu[366 * SV], 0.0);
#pragma unroll 1
      for (int k = 0; k < N; k += 4)
        st_global_256(out + ST_DINV + k, dinv[k * SV], dinv[(k + 1) * SV], dinv[(k + 2) * SV], dinv[(k + 3) * SV]);
    }
  }
}

struct SdirkTlmArgs {
  int ncells;
  const long long* rec_off;
  const int* cell;
  const double* fat;
  double* y_tlm;
  unsigned long long* queue;
  int ntlm;
};

constexpr int ST_VEC = 7;   // per-lane vectors: Y_tlm, Z_tlm(5), G
constexpr size_t SDIRK_TLM_SMEM = sizeof(double) * ((size_t)ST_VEC * 32 * 32 + ST_REC);

__global__ void __launch_bounds__(32, 3)
k_sdirk_tlm_cell(SdirkParams sp, SdirkTlmArgs a) {
  extern __shared__ __align__(16) double cell_smem[];
  constexpr int N = 32, SV = 32;
  const int lane = threadIdx.x;
  double* const rec = cell_smem;
  double* const pv = cell_smem + ST_REC + lane;
  double* const YT = pv;
  double* const ZT = pv + N * SV;          // ZT[(s * N + i) * SV]
  double* const G = pv + 6 * N * SV;
  const bool active = lane < a.ntlm;
  const int S = sp.S;
  for (;;) {
    unsigned long long slot = 0;
    if (lane == 0) slot = atomicAdd(a.queue, 1ull);
    slot = __shfl_sync(0xffffffffu, slot, 0);
    if (slot >= (unsigned long long)a.ncells) break;
    const long long r0 = a.rec_off[slot], r1 = a.rec_off[slot + 1];
    const int cell = a.cell[slot];
    double* const ycol = a.y_tlm + ((size_t)cell * a.ntlm + (active ? lane : 0)) * N;
    if (active) {
#pragma unroll 4
      for (int i = 0; i < N; ++i) YT[i * SV] = ycol[i];
    }
    for (long long q = r0; q < r1; ++q) {
      __syncwarp();
      {
        const double2* src = reinterpret_cast<const double2*>(a.fat + (size_t)q * ST_REC);
        double2* dst = reinterpret_cast<double2*>(rec);
        for (int e = lane; e < ST_REC / 2; e += 32) dst[e] = __ldg(src + e);
      }
      __syncwarp();
      if (!active) continue;
      const double H = rec[ST_HDR + 1];
      const unsigned hw = (unsigned)rec[ST_HDR + 2];
      const unsigned swaps = hw >> 24;
      const double hg = H * sp.Gamma;
      const double hgi = 1.0 / (H * sp.Gamma);
      double x[32], o[32];
#pragma unroll 1
      for (int s = 0; s < S; ++s) {
        const double* Jv = rec + ST_J + s * 276;
        double* Zs = ZT + s * N * SV;
        const int iters = (int)((hw >> (4 * s)) & 15u);
        // G = (H gamma) J_i Y_tlm + sum_j theta_ij Z_tlm_j   (:706-716)
#pragma unroll
        for (int i = 0; i < N; ++i) x[i] = YT[i * SV];
        cbm4_cell::j_vec_seq(Jv, x, o);
#pragma unroll
        for (int i = 0; i < N; ++i) { o[i] = hg * o[i]; x[i] = 0.0; }
        for (int j = 0; j < s; ++j) {
          const double th = sp.Theta[s][j];
          const double* Zj = ZT + j * N * SV;
#pragma unroll
          for (int i = 0; i < N; ++i) o[i] = o[i] + th * Zj[i * SV];
        }
#pragma unroll
        for (int i = 0; i < N; ++i) G[i * SV] = o[i];
        // fixed number of modified-Newton iterations on Z_tlm (x), :724-775
#pragma unroll 1
        for (int it = 0; it < iters; ++it) {
          asm volatile("" ::: "memory");   // the operands are re-read from shared memory (broadcast), not hoisted into spilled registers
          cbm4_cell::j_vec_seq(Jv, x, o);
#pragma unroll
          for (int i = 0; i < N; ++i) o[i] = hgi * ((hg * o[i] + G[i * SV]) - x[i]);
          asm volatile("" ::: "memory");
          cbm4_cell::lun_solve_seq(rec + ST_LU, rec + ST_DINV, swaps, o);
#pragma unroll
          for (int i = 0; i < N; ++i) x[i] = x[i] + o[i];
        }
#pragma unroll
        for (int i = 0; i < N; ++i) Zs[i * SV] = x[i];
      }
      // Y_tlm += sum_i d_i Z_tlm_i   (:851-862)
      for (int s = 0; s < S; ++s) {
        const double d = sp.D[s];
        if (d != 0.0) {
          const double* Zs = ZT + s * N * SV;
#pragma unroll 4
          for (int i = 0; i < N; ++i) YT[i * SV] = YT[i * SV] + d * Zs[i * SV];
        }
      }
    }
    if (active) {
#pragma unroll 4
      for (int i = 0; i < N; ++i) ycol[i] = YT[i * SV];
    }
  }
}


// ---- tensor-memory variant of the column sweep (the shipped one for the static-pattern path): the lane's seven vectors
// (Y_tlm, Z_tlm_1..5, G = 224 doubles) live in the lane's 256 doubles of tensor memory, shared memory only stages the records
// (14.3 KB per warp), one CTA of four sweep warps per SM.  Same operations in the same order as k_sdirk_tlm_cell.
constexpr uint32_t ST_TM_Y = 0, ST_TM_G = 64, ST_TM_Z = 128;          // 32-bit columns; Z_tlm_s at ST_TM_Z + 64 s
constexpr int ST_TM_WARPS = 4;
constexpr size_t SDIRK_TLM_TM_SMEM = sizeof(double) * (size_t)ST_TM_WARPS * ST_REC;

__global__ void __launch_bounds__(ST_TM_WARPS * 32, 1)
k_sdirk_tlm_cell_tm(SdirkParams sp, SdirkTlmArgs a) {
  extern __shared__ __align__(16) double cell_smem[];
  __shared__ uint32_t tmem_base_s;
  constexpr int N = 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* const rec = cell_smem + (size_t)warp * ST_REC;
  if (warp == 0) tmem::alloc512(&tmem_base_s);
  tmem::fence_before_sync();
  __syncthreads();
  tmem::fence_after_sync();
  const uint32_t tm = tmem_base_s + ((uint32_t)(warp * 32) << 16);
  const bool active = lane < a.ntlm;
  const int S = sp.S;
  for (;;) {
    unsigned long long slot = 0;
    if (lane == 0) slot = atomicAdd(a.queue, 1ull);
    slot = __shfl_sync(0xffffffffu, slot, 0);
    if (slot >= (unsigned long long)a.ncells) break;
    const long long r0 = a.rec_off[slot], r1 = a.rec_off[slot + 1];
    const int cell = a.cell[slot];
    double* const ycol = a.y_tlm + ((size_t)cell * a.ntlm + (active ? lane : 0)) * N;
    double x[32], o[32];
#pragma unroll
    for (int i = 0; i < N; ++i) x[i] = ycol[i];
    tmem::st32(tm + ST_TM_Y, x);
    tmem::wait_st();
    for (long long q = r0; q < r1; ++q) {
      __syncwarp();
      {
        const double2* src = reinterpret_cast<const double2*>(a.fat + (size_t)q * ST_REC);
        double2* dst = reinterpret_cast<double2*>(rec);
        for (int e = lane; e < ST_REC / 2; e += 32) dst[e] = __ldg(src + e);
      }
      __syncwarp();
      const double H = rec[ST_HDR + 1];
      const unsigned hw = (unsigned)rec[ST_HDR + 2];
      const unsigned swaps = hw >> 24;
      const double hg = H * sp.Gamma;
      const double hgi = 1.0 / (H * sp.Gamma);
#pragma unroll 1
      for (int s = 0; s < S; ++s) {
        const double* Jv = rec + ST_J + s * 276;
        const int iters = (int)((hw >> (4 * s)) & 15u);
        // G = (H gamma) J_i Y_tlm + sum_j theta_ij Z_tlm_j   (:706-716)
        tmem::ld32(tm + ST_TM_Y, x);
        cbm4_cell::j_vec_seq(Jv, x, o);
#pragma unroll
        for (int i = 0; i < N; ++i) { o[i] = hg * o[i]; x[i] = 0.0; }
        for (int j = 0; j < s; ++j) {
          const double th = sp.Theta[s][j];
          double t[16];
          tmem::ld16(tm + ST_TM_Z + 64 * j, t);
#pragma unroll
          for (int i = 0; i < 16; ++i) o[i] = o[i] + th * t[i];
          tmem::ld16(tm + ST_TM_Z + 64 * j + 32, t);
#pragma unroll
          for (int i = 0; i < 16; ++i) o[16 + i] = o[16 + i] + th * t[i];
        }
        tmem::st32(tm + ST_TM_G, o);
        tmem::wait_st();
        // fixed number of modified-Newton iterations on Z_tlm (x), :724-775
#pragma unroll 1
        for (int it = 0; it < iters; ++it) {
          asm volatile("" ::: "memory");
          cbm4_cell::j_vec_seq(Jv, x, o);
          {
            double t[16];
            tmem::ld16(tm + ST_TM_G, t);
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = hgi * ((hg * o[i] + t[i]) - x[i]);
            tmem::ld16(tm + ST_TM_G + 32, t);
#pragma unroll
            for (int i = 0; i < 16; ++i) o[16 + i] = hgi * ((hg * o[16 + i] + t[i]) - x[16 + i]);
          }
          asm volatile("" ::: "memory");
          cbm4_cell::lun_solve_seq(rec + ST_LU, rec + ST_DINV, swaps, o);
#pragma unroll
          for (int i = 0; i < N; ++i) x[i] = x[i] + o[i];
        }
        tmem::st32(tm + ST_TM_Z + 64 * s, x);
        tmem::wait_st();
      }
      // Y_tlm += sum_i d_i Z_tlm_i   (:851-862)
      tmem::ld32(tm + ST_TM_Y, x);
      for (int s = 0; s < S; ++s) {
        const double d = sp.D[s];
        if (d != 0.0) {
          double t[16];
          tmem::ld16(tm + ST_TM_Z + 64 * s, t);
#pragma unroll
          for (int i = 0; i < 16; ++i) x[i] = x[i] + d * t[i];
          tmem::ld16(tm + ST_TM_Z + 64 * s + 32, t);
#pragma unroll
          for (int i = 0; i < 16; ++i) x[16 + i] = x[16 + i] + d * t[i];
        }
      }
      tmem::st32(tm + ST_TM_Y, x);
      tmem::wait_st();
    }
    tmem::ld32(tm + ST_TM_Y, x);
    if (active) {
#pragma unroll
      for (int i = 0; i < N; ++i) ycol[i] = x[i];
    }
  }
  tmem::fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem::dealloc512(tmem_base_s);
}

}  // namespace fatode
