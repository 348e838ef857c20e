// General-pivoting pass of the SDIRK sensitivity families (TLM/SDIRK_TLM, ADJ/SDIRK_ADJ) for CBM-IV.
//
// The static-pattern kernels (sdirk_cell_tlm.cuh, sdirk_cell_adj.cuh) hand back every system whose forward factorisations leave
// the static pivot pattern (IERR -20 from the first pass: all systems at rtol >= 1e-3, none at rtol <= 1e-4 in the driver
// problem).  Those systems are integrated again from their initial state with netlib's general partial pivoting throughout:
//   forward   k_sdirk_fwd_cell<Cbm4DenseCell, .., MODE>  (dense DGETF2 / DGETRS('N') per thread, same tape)
//   prepare   k_sdirk_{tlm,adj}_prep_dense               (dense DGETF2 of the step / stage matrices per thread)
//   sweep     k_sdirk_{tlm,adj}_cell_dense               (lane = column; ONE dense factorisation per warp in shared memory,
//                                                          DGETRS('N') / DGETRS('T') on the lane's shared-memory vector)
// Same operation order as the reference (and as the static kernels wherever the pivot sequences coincide), so state, counters
// and Y_tlm / Lambda stay bit-identical to the oracle.  About 3x the arithmetic of the static path; a fallback.
#pragma once
#include "sdirk_cell_tlm.cuh"
#include "sdirk_cell_adj.cuh"
#include "model_cbm4_dense_cell.cuh"

namespace fatode {

constexpr int DN_LU = 32 * 32 + 32;          // dense factors + pivot rows (as doubles)
// TLM record: T, H, packed iteration counts, S | LU | J_1..J_5
constexpr int DT_HDR = 0, DT_LU = 4, DT_J = DT_LU + DN_LU, DT_REC = DT_J + 5 * 276;          // 2440 doubles
// ADJ record: T, H, S, pad | AP 5 x 32 | per stage: LU, J, pad
constexpr int DA_HDR = 0, DA_P = 4, DA_STAGE = DA_P + 5 * 32, DA_LU = 0, DA_J = DN_LU, DA_CHUNK = DN_LU + 276 + 4;   // 1336
constexpr int DA_REC = DA_STAGE + 5 * DA_CHUNK;                                               // 6844 doubles
static_assert(DT_REC % 4 == 0 && DA_REC % 4 == 0 && DA_CHUNK % 4 == 0 && DT_J % 4 == 0 && DA_STAGE % 4 == 0, "32-byte sections");
// distinct lane counts: each preparation kernel keeps its own instantiation of the generated routines (see sdirk_cell_adj.cuh)
constexpr int DT_PREP_LANES = 16, DA_PREP_LANES = 14;

template <int SV>
__device__ __forceinline__ void dense_store_lu(const double* a, double* out) {
#pragma unroll 2
  for (int e = 0; e < DN_LU; e += 4) st_global_256(out + e, a[e * SV], a[(e + 1) * SV], a[(e + 2) * SV], a[(e + 3) * SV]);
}

__global__ void __launch_bounds__(32, 1)
k_sdirk_tlm_prep_dense(SdirkParams sp, Cbm4Const mc, SdirkTlmPrepArgs a) {
  extern __shared__ __align__(16) double cell_smem[];
  constexpr int SV = DT_PREP_LANES, N = 32;
  if (threadIdx.x >= SV) return;
  double* const lu = cell_smem + threadIdx.x;
  double* const Tv = lu + DN_LU * SV;
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
    double* out = a.fat + (size_t)q * DT_REC;
    const double T = rec[0], H = rec[1];
    if (part < S) {
      const double* Z = rec + 4 + N + part * N;
#pragma unroll 4
      for (int i = 0; i < N; ++i) Tv[i * SV] = rec[4 + i] + Z[i];
      Cbm4Cell::photo(T + sp.C[part] * H, mc, PH);
      cbm4_cell::jac_values<SV>(Tv, mc.fix, PH, out + DT_J + part * 276);
    } else {
#pragma unroll 4
      for (int i = 0; i < N; ++i) Tv[i * SV] = rec[4 + i];
      Cbm4Cell::photo(T, mc, PH);
      cbm4_cell::build_E_dense<SV>(Tv, mc.fix, PH, 1.0 / (H * sp.Gamma), lu);
      dense32_lu_factor<SV>(lu);
      st_global_256(out + DT_HDR, T, H, rec[2], rec[3]);
      dense_store_lu<SV>(lu, out + DT_LU);
    }
  }
}

__global__ void __launch_bounds__(32, 1)
k_sdirk_adj_prep_dense(SdirkParams sp, Cbm4Const mc, SdirkAdjPrepArgs a) {
  extern __shared__ __align__(16) double cell_smem[];
  constexpr int SV = DA_PREP_LANES, N = 32;
  if (threadIdx.x >= SV) return;
  double* const lu = cell_smem + threadIdx.x;
  double* const Tv = lu + DN_LU * SV;
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
    double* out = a.fat + (size_t)q * DA_REC;
    const double T = rec[0], H = rec[1];
    const double* Z = rec + 4 + N + s * N;
#pragma unroll 4
    for (int i = 0; i < N; ++i) Tv[i * SV] = rec[4 + i] + Z[i];
    if (s == 0) st_global_256(out + DA_HDR, T, H, (double)S, 0.0);
    double* ch = out + DA_STAGE + s * DA_CHUNK;
    Cbm4Cell::photo(T + sp.C[s] * H, mc, PH);
    cbm4_cell::jac_values<SV>(Tv, mc.fix, PH, ch + DA_J);
    cbm4_cell::arp_values<SV>(Tv, mc.fix, out + DA_P + s * 32);
    cbm4_cell::build_E_dense<SV>(Tv, mc.fix, PH, 1.0 / (H * sp.Gamma), lu);
    dense32_lu_factor<SV>(lu);
    dense_store_lu<SV>(lu, ch + DA_LU);
    st_global_256(ch + DA_J + 276, 0.0, 0.0, 0.0, 0.0);
  }
}

constexpr int DT_VEC = 8;   // per-lane vectors: Y_tlm, Z_tlm(5), G, W (right-hand side of the dense solve)
constexpr size_t SDIRK_TLM_DENSE_SMEM = sizeof(double) * ((size_t)DT_VEC * 32 * 32 + DT_REC);

__global__ void __launch_bounds__(32, 2)
k_sdirk_tlm_cell_dense(SdirkParams sp, SdirkTlmArgs a) {
  extern __shared__ __align__(16) double cell_smem[];
  constexpr int N = 32, SV = 32;
  const int lane = threadIdx.x;
  double* const rec = cell_smem;
  double* const pv = cell_smem + DT_REC + lane;
  double* const YT = pv;
  double* const ZT = pv + N * SV;
  double* const G = pv + 6 * N * SV;
  double* const W = pv + 7 * N * SV;
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
        const double2* src = reinterpret_cast<const double2*>(a.fat + (size_t)q * DT_REC);
        double2* dst = reinterpret_cast<double2*>(rec);
        for (int e = lane; e < DT_REC / 2; e += 32) dst[e] = __ldg(src + e);
      }
      __syncwarp();
      if (!active) continue;
      const double H = rec[DT_HDR + 1];
      const unsigned hw = (unsigned)rec[DT_HDR + 2];
      const double hg = H * sp.Gamma;
      const double hgi = 1.0 / (H * sp.Gamma);
      double x[32], o[32];
#pragma unroll 1
      for (int s = 0; s < S; ++s) {
        const double* Jv = rec + DT_J + s * 276;
        double* Zs = ZT + s * N * SV;
        const int iters = (int)((hw >> (4 * s)) & 15u);
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
#pragma unroll 1
        for (int it = 0; it < iters; ++it) {
          asm volatile("" ::: "memory");
          cbm4_cell::j_vec_seq(Jv, x, o);
#pragma unroll
          for (int i = 0; i < N; ++i) W[i * SV] = hgi * ((hg * o[i] + G[i * SV]) - x[i]);
          dense32_solve_n_shared<SV>(rec + DT_LU, W);
#pragma unroll
          for (int i = 0; i < N; ++i) x[i] = x[i] + W[i * SV];
        }
#pragma unroll
        for (int i = 0; i < N; ++i) Zs[i * SV] = x[i];
      }
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

constexpr size_t SDIRK_ADJ_DENSE_SMEM = sizeof(double) * ((size_t)SA_VEC * 32 * 32 + DA_STAGE + DA_CHUNK);

__global__ void __launch_bounds__(32, 3)
k_sdirk_adj_cell_dense(SdirkParams sp, SdirkAdjCoef co, SdirkAdjArgs a) {
  extern __shared__ __align__(16) double cell_smem[];
  constexpr int N = 32, SV = 32;
  const int lane = threadIdx.x;
  double* const hdr = cell_smem;
  double* const ch = cell_smem + DA_STAGE;
  double* const pv = cell_smem + DA_STAGE + DA_CHUNK + lane;
  double* const LAM = pv;
  double* const MU = pv + 1 * N * SV;
  double* const U = pv + 2 * N * SV;
  const bool active = lane < a.nadj;
  const int S = sp.S;
  for (;;) {
    unsigned long long slot = 0;
    if (lane == 0) slot = atomicAdd(a.queue, 1ull);
    slot = __shfl_sync(0xffffffffu, slot, 0);
    if (slot >= (unsigned long long)a.ncells) break;
    const long long r0 = a.rec_off[slot], r1 = a.rec_off[slot + 1];
    const int cell = a.cell[slot];
#pragma unroll 4
    for (int i = 0; i < N; ++i) { LAM[i * SV] = (i == lane) ? 1.0 : 0.0; MU[i * SV] = 0.0; }
    for (long long q = r1 - 1; q >= r0; --q) {
      const double* rec = a.fat + (size_t)q * DA_REC;
      __syncwarp();
      {
        const double2* src = reinterpret_cast<const double2*>(rec);
        double2* dst = reinterpret_cast<double2*>(hdr);
        for (int e = lane; e < DA_STAGE / 2; e += 32) dst[e] = __ldg(src + e);
      }
      double H = 0.0, hgi = 0.0;
      double x[32], o[32];
#pragma unroll 1
      for (int s = S - 1; s >= 0; --s) {
        __syncwarp();
        {
          const double2* src = reinterpret_cast<const double2*>(rec + DA_STAGE + s * DA_CHUNK);
          double2* dst = reinterpret_cast<double2*>(ch);
          for (int e = lane; e < DA_CHUNK / 2; e += 32) dst[e] = __ldg(src + e);
        }
        __syncwarp();
        if (!active) continue;
        H = hdr[DA_HDR + 1];
        hgi = 1.0 / (H * sp.Gamma);
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
        cbm4_cell::jt_vec_seq(ch + DA_J, x, o);
        double* Us = U + s * N * SV;
#pragma unroll
        for (int i = 0; i < N; ++i) Us[i * SV] = hgi * (H * o[i]);
        dense32_solve_t_shared<SV>(ch + DA_LU, Us);
      }
      if (!active) continue;
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
          const double* AP = hdr + DA_P + s * 32;
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
}

}  // namespace fatode
