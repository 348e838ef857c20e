// Discrete Rosenbrock adjoint, backward sweep — producer/consumer warp pairs with TMEM-resident
// running sums (second-generation kernel; same mathematics as ros_warp_adj.cuh, which documents the
// regrouping against ros_DadjInt, ADJ/ROS_ADJ/ROS_ADJ_f90_Integrator.F90:1177-1441).
//
// One CTA per SM, 8 warps = 4 cells in flight:
//   warps 4..7  PRODUCERS (lane = species).  Everything that does not depend on the adjoint variables:
//               J(T,Y0) -> LU (same code as the forward sweep) -> row-major L\U, 1/diag(U), pivot
//               directory; Hessian(T); dJ/dt; per stage the values of J(T_i,Y_i), of
//               M'_i = Hess x K_i + h*gamma_i*dJ/dt and of a_p.  They only read the tape, so they run one
//               work item (= one stage of one step) ahead of their consumer, into double buffers.
//   warps 0..3  CONSUMERS (lane = adjoint column).  x = m_i*lambda + W_i, two triangular substitutions
//               with the 32-vector in registers, static sparse J^T x / M'^T x / S^T x, updates of the
//               running sums.  Their private per-lane arrays W_1..W_5, Lambda, Lambda-increment and Mu
//               (256 doubles per lane) live in TENSOR MEMORY: consumer warp q owns TMEM lanes 32q..32q+31
//               and moves 16 or 32 doubles per LDTM/STTM.  This frees 55 KB of shared memory per cell,
//               which is what allows 4 cells (8 warps) per SM instead of 3 single warps.
// The pair hands items over in lockstep through a 64-thread named barrier.
#pragma once
#include "fatode_common.cuh"
#include "ros_warp_fwd.cuh"
#include "ros_warp_adj.cuh"
#include "tmem.cuh"

namespace fatode {

constexpr int ADJ2_SLOTS = 4;           // cells in flight per CTA
constexpr int ADJ2_THREADS = 256;

template <class Model>
struct Adj2Smem {
  // per-slot layout in doubles
  static constexpr int OFF_MODEL = 0;                                   // producer: model scratch (HESS in its L2 area)
  static constexpr int OFF_HS = Model::WS_DOUBLES;                      // [NHESS] Hessian(T) of the current step
  static constexpr int OFF_DJ = OFF_HS + ((Model::NHESS + 1) / 2) * 2;  // [32]
  static constexpr int OFF_KB = OFF_DJ + 32;                            // [32]
  static constexpr int OFF_J0 = OFF_KB + 32;                            // [288] J(T,Y0)
  static constexpr int OFF_JD = OFF_J0 + 288;                           // [288] J(T+Delta,Y0)
  static constexpr int OFF_TILE = OFF_JD + 288;                         // [32][LDA]
  static constexpr int OFF_LU = OFF_TILE + 32 * LDA;                    // 2 x {LUs[32*34], DINV[32], WHO[16]}
  static constexpr int LU_STRIDE = 32 * 34 + 32 + 16;
  static constexpr int OFF_STG = OFF_LU + 2 * LU_STRIDE;                // 2 x {JV[280], MV[260], AP[32], HDR[4]}
  static constexpr int STG_JV = 0, STG_MV = 280, STG_AP = 540, STG_HDR = 572, STG_STRIDE = 576;
  static constexpr int OFF_CTRL = OFF_STG + 2 * STG_STRIDE;             // [8]: cell id, nsteps
  static constexpr int DOUBLES = OFF_CTRL + 8;
};

// TMEM columns (32-bit) of a consumer lane
constexpr uint32_t TM_W = 0;          // W[5][32]  -> 320 columns
constexpr uint32_t TM_LAM = 320;      // Lambda[32]
constexpr uint32_t TM_MU = 384;       // Mu[32] (29 used)
constexpr uint32_t TM_LACC = 448;     // Lambda increment of the current step [32]

__device__ __forceinline__ void pair_barrier(int slot) { asm volatile("bar.sync %0, 64;" ::"r"(slot + 1) : "memory"); }

// ---------------------------------------------------------------------------------------------------------------
// producer: prepare work item (step, is) of the current cell
template <class Model>
__device__ __forceinline__ void adj2_produce(const RosParams& rp, double* sm, const double* __restrict__ ent, int is,
                                             int step_parity, int stg, bool with_mu, int lane) {
  typedef Adj2Smem<Model> L;
  const int S = rp.S;
  double* ws = sm + L::OFF_MODEL;
  double* HS = sm + L::OFF_HS;
  double* DJ = sm + L::OFF_DJ;
  double* KB = sm + L::OFF_KB;
  double* J0 = sm + L::OFF_J0;
  double* JD = sm + L::OFF_JD;
  double* tile = sm + L::OFF_TILE;
  double* LUb = sm + L::OFF_LU + step_parity * L::LU_STRIDE;
  double* LUs = LUb;
  double* DINV = LUb + 32 * 34;
  int* WHO = reinterpret_cast<int*>(LUb + 32 * 34 + 32);
  double* STG = sm + L::OFF_STG + stg * L::STG_STRIDE;
  const double dDir = (rp.tend >= rp.tstart) ? 1.0 : -1.0;
  const double T = ent[0], H = ent[1];
  const double* Yst = ent + 4;
  const double* Kst = ent + 4 + S * 32;
  const double y0 = Yst[lane];
  if (is == S - 1) {   // first item of a step: factorisation and step-wide quantities
    Model::set_time(ws, T, lane);
    Model::set_y(ws, y0, lane);
    Model::jac(ws, J0, lane);
    const double ghinv = __ddiv_rn(1.0, __dmul_rn(__dmul_rn(dDir, H), rp.Gamma[0]));
    Model::template scatter_E<LDA>(J0, ghinv, tile, lane);
    int pos;
    warp_lu_factor(tile, WHO, pos, lane);
    for (int r = 0; r < 32; ++r) LUs[r * 34 + lane] = tile[lane * LDA + WHO[r]];
    __syncwarp();
    DINV[lane] = 1.0 / LUs[lane * 34 + lane];
    Model::hess(ws, HS, lane);
    if (!rp.autonomous) {
      const double Delta = sqrt(rp.roundoff) * fmax(1.0e-5, fabs(T));   // DeltaMin of ros_DadjInt :1212
      Model::set_time(ws, T + Delta, lane);
      Model::jac(ws, JD, lane);
      if (lane < Model::NJT) {
        const int e = Model::jac_time_entry(lane);
        DJ[lane] = (JD[e] - J0[e]) * (1.0 / Delta);
      }
    } else {
      DJ[lane] = 0.0;
    }
    __syncwarp();
  }
  const double Tau = T + rp.Alpha[is] * dDir * H;
  const double yi = Yst[is * 32 + lane];
  const double ki = Kst[is * 32 + lane];
  Model::set_time(ws, Tau, lane);
  Model::set_y(ws, yi, lane);
  Model::jac(ws, STG + L::STG_JV, lane);
  double ap = 0.0;
  if (with_mu) ap = Model::arp(ws, lane);
  KB[lane] = ki;
  Model::set_y(ws, y0, lane);
  if (with_mu) ap += Model::aj(ws, KB, lane);
  STG[L::STG_AP + lane] = ap;
  const double hg = rp.autonomous ? 0.0 : H * rp.Gamma[is];
  Model::build_M(ws, HS, KB, DJ, hg, STG + L::STG_MV, lane);
  {
    const bool ident = __all_sync(FATODE_FULL_MASK, WHO[lane] == lane);
    if (lane == 0) {
      STG[L::STG_HDR + 0] = H;
      STG[L::STG_HDR + 1] = ident ? 1.0 : 0.0;
    }
  }
  __syncwarp();
}


template <class Model, int HALF>
__device__ __forceinline__ void adj2_lambda_half(const double* __restrict__ MV, const double (&x)[32], const double (&v)[32],
                                                 bool first, bool last, uint32_t tm) {
  double acc[16];
  if (!first) tmem::ld16(tm + TM_LACC + 32 * HALF, acc);
  Model::mt_rows(MV, x, [&](auto rc, double h) {
    constexpr int r = decltype(rc)::value;
    if constexpr (r / 16 == HALF) {
      const double t = v[r] + h;
      acc[r - 16 * HALF] = first ? t : (acc[r - 16 * HALF] + t);
    }
  });
  if (last) {
    double lam[16];
    tmem::ld16(tm + TM_LAM + 32 * HALF, lam);
#pragma unroll
    for (int r = 0; r < 16; ++r) lam[r] += acc[r];
    tmem::st16(tm + TM_LAM + 32 * HALF, lam);
  } else {
    tmem::st16(tm + TM_LACC + 32 * HALF, acc);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// consumer: one work item.  tm = TMEM address of this warp's lane quarter, column 0.
template <class Model>
__device__ __forceinline__ void adj2_consume(const RosParams& rp, const double* sm, int is, int step_parity, int stg,
                                             bool with_mu, uint32_t tm) {
  typedef Adj2Smem<Model> L;
  const int S = rp.S;
  const double* LUb = sm + L::OFF_LU + step_parity * L::LU_STRIDE;
  const double* LUs = LUb;
  const double* DINV = LUb + 32 * 34;
  const int* WHO = reinterpret_cast<const int*>(LUb + 32 * 34 + 32);
  const double* STG = sm + L::OFF_STG + stg * L::STG_STRIDE;
  const double* JV = STG + L::STG_JV;
  const double* MV = STG + L::STG_MV;
  const double* AP = STG + L::STG_AP;
  const double dDir = (rp.tend >= rp.tstart) ? 1.0 : -1.0;
  const double H = STG[L::STG_HDR + 0];
  const bool ident = STG[L::STG_HDR + 1] != 0.0;
  const bool first = (is == S - 1);
  const double mi = rp.M[is];

  double x[32];
  {   // x = m_i * lambda + W_i      (:1285-1300; W_i already holds sum_j a_ji V_j + c_ji/h U_j)
    tmem::wait_st();
    tmem::ld32(tm + TM_LAM, x);
    if (!first) {
      double w[16];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        tmem::ld16(tm + TM_W + 64 * is + 32 * c, w);
#pragma unroll
        for (int r = 0; r < 16; ++r) x[16 * c + r] = fma(mi, x[16 * c + r], w[r]);
      }
    } else {
#pragma unroll
      for (int r = 0; r < 32; ++r) x[r] = mi * x[r];
    }
  }
  solve_transposed(LUs, DINV, x);
  if (!ident) {   // x = P^T z: natural index who[r] <- logical r; scratch = a dead W slab
    const uint32_t sc = tm + TM_W + 64 * (first ? (S > 1 ? S - 2 : 0) : is);
#pragma unroll
    for (int r = 0; r < 32; ++r) tmem::st1(sc + 2 * WHO[r], x[r]);
    tmem::wait_st();
    tmem::ld32(sc, x);
  }
  double v[32];
  Model::jt_vec(JV, x, v);                 // V_i = J(T_i,Y_i)^T U_i  (:1315)
  // running sums for the earlier stages j < is: W_j (+)= a_ij V_i + c_ij/h U_i   (:1292-1300)
  {
    const double inv_dH = 1.0 / (dDir * H);
    for (int j = 0; j < is; ++j) {
      const double ca = rp.A[is * (is - 1) / 2 + j];
      const double cc = rp.C[is * (is - 1) / 2 + j] * inv_dH;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        double w[16];
        if (!first) tmem::ld16(tm + TM_W + 64 * j + 32 * c, w);
#pragma unroll
        for (int r = 0; r < 16; ++r) {
          const double t = fma(ca, v[16 * c + r], cc * x[16 * c + r]);
          w[r] = first ? t : (w[r] + t);
        }
        tmem::st16(tm + TM_W + 64 * j + 32 * c, w);
      }
    }
  }
  // Lambda increment: LACC (+)= V_i + M'^T U_i     (:1328-1332 and the dJ/dt term :1402-1403); at the last stage of
  // the step Lambda += increment (:1324-1339 run after the stage loop).  Two halves of 16 rows bound the register use.
  adj2_lambda_half<Model, 0>(MV, x, v, first, is == 0, tm);
  adj2_lambda_half<Model, 1>(MV, x, v, first, is == 0, tm);
  if (with_mu) {   // Mu += a_p * (S^T U_i)_p     (:1352-1365)
    double mu[32];
    tmem::ld32(tm + TM_MU, mu);
    Model::stoich_cols(x, [&](auto pc, double s) {
      constexpr int p = decltype(pc)::value;
      mu[p] = fma(AP[p], s, mu[p]);
    });
    tmem::st32(tm + TM_MU, mu);
  }
}

}  // namespace fatode
