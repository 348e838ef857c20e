// Discrete Rosenbrock adjoint, backward sweep over the FAT tape written by k_ros_fwd_cell (ros_cell_fwd.cuh).
//
// Mathematics and regrouping against ros_DadjInt (ADJ/ROS_ADJ/ROS_ADJ_f90_Integrator.F90:1177-1441) are those of
// ros_warp_adj2.cuh: lane m of a warp owns adjoint column m of one cell (the NADJ columns never couple,
// :1285-1317) with its 32-vector in registers and its running sums W_1..W_5, Lambda, the Lambda increment and Mu
// in TENSOR MEMORY (256 doubles per lane).  What changed: everything that does not depend on the adjoint
// variables — the L\U factors of the step, J(T_i,Y_i), M'_i = Hess x K_i + h*gamma_i*dJ/dt and a_p per stage —
// now arrives precomputed on the tape, so there are no producer warps; each warp streams its cell's tape through
// two double-buffered shared-memory slots with 1-D bulk TMA copies (cp.async.bulk + mbarrier complete_tx) issued
// one work item (= one stage of one step) ahead, and the transposed solve runs on the static sparse pattern
// (335 FMA instead of 1056, cbm4_cell::lut_solve).
//
// One CTA per SM: warps 0..3 keep their private state in tensor memory (4 x 32 lanes x 512 columns); further warps
// (ADJ3_WARPS > 4) would keep it in shared memory — measured on B200: 4 warps 1464 ms, 5 warps 1659 ms, 6 warps
// 1617 ms per 32768 cells, so four it is.  Cells are pulled from an atomic queue.
#pragma once
#include "fatode_common.cuh"
#include "ros_cell_fwd.cuh"
#include "ros_warp_adj2.cuh"   // TMEM column map (TM_W, TM_LAM, TM_MU, TM_LACC) shared with the general kernel
#include "tmem.cuh"

namespace fatode {

constexpr int ADJ3_WARPS = 4;          // cells in flight per SM (4: all private state in tensor memory; 5 and 6 measured slower)
constexpr int ADJ3_WARP_DOUBLES = 2 * FT_STEP_COPY + 2 * FT_STAGE + 8;   // step blocks, stage blocks, 4 mbarriers (+pad)

namespace mbar {
__device__ __forceinline__ uint32_t saddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void init(void* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(saddr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void expect_tx(void* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(saddr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, void* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(saddr(dst)),
               "l"(src), "r"(bytes), "r"(saddr(bar))
               : "memory");
}
__device__ __forceinline__ void wait(void* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(saddr(bar)), "r"(parity)
        : "memory");
  } while (!ok);
}
}  // namespace mbar

// k_expand_tape: fills the per-stage blocks of the fat tape from what the forward sweep recorded (T, H, Ystage, K_1..K_S):
//   JV = J(T_i, Y_i) values (:1311), MV = M'_i = Hess(T,Y) x K_i + h*gamma_i*dJ/dt (:1330, :1379-1403),
//   AP = a_p = dF/dp reactant product at Y_i + (dJ/dp x K_i) at Y (:1348, :1363).
// None of it depends on the adjoint variables, every (step, stage) is independent: one thread per item, millions of
// threads per wave, so the straight-line model code runs at full occupancy instead of inside the latency-bound
// forward loop.
// once per call: Hessian(T) of the mechanism into a staging buffer (copied to c_cbm4_hess by the host)
__global__ void k_cbm4_hess(double fix, double* __restrict__ out) {
  double y[32], ph[cbm4_cell::NPHOTO];
  for (int i = 0; i < 32; ++i) y[i] = 0.0;
  for (int i = 0; i < cbm4_cell::NPHOTO; ++i) ph[i] = 0.0;
  cbm4_cell::hess_values<1>(y, fix, ph, out);
}

// which cells of a wave take part in the second sweep: successful ones; for the tangent-linear sweep also the cells whose
// state sweep ended with one of the reference's error codes (ros_TLM_Int has advanced Y_tlm up to that point)
__device__ __forceinline__ bool sweep_takes_cell(int ierr, bool tlm) { return ierr == 1 || (tlm && ierr < 0 && ierr > -100); }

constexpr int XP_ROW = 226;   // shared-memory row of one step's inputs (Y0 + K = 224 doubles, +2: conflict-free across steps)
template <class Cell>
__global__ void __launch_bounds__(FT_CHUNK * 6, 4)
k_expand_tape(RosParams rp, typename Cell::Const mc, unsigned nchunks, const int* __restrict__ chunk_owner,
              const int* __restrict__ chunk_seq, const int* __restrict__ ids, const int* __restrict__ ierr,
              const int* __restrict__ nacc, const int* __restrict__ last_chunk, const double* __restrict__ tape,
              double* __restrict__ fat, const long long* __restrict__ fat_off, const int* __restrict__ batch_of,
              int batch, int mode) {   // mode 0: no a_p, 1: a_p (adjoint with Mu), 2: tangent-linear sweep (hg = 0, dJ/dt in the AP slot)
  // one block per tape chunk (FT_CHUNK steps x S stages = one thread per item); the steps' inputs are staged through
  // shared memory with coalesced loads, the outputs go straight to the tape
  __shared__ __align__(16) double in_s[FT_CHUNK * XP_ROW];
  constexpr int N = Cell::N;
  const int S = rp.S;
  const long c = blockIdx.x;
  const int w = chunk_owner[c];
  if (batch_of[w] != batch || !sweep_takes_cell(ierr[ids ? ids[w] : w], mode == 2)) return;   // other batches / unfinished cells
  const long long fo = fat_off[w];            // first stage block of this cell in its batch's fat buffer
  const int cnt = (c == last_chunk[w]) ? min(((nacc[w] - 1) % FT_CHUNK) + 1, nacc[w]) : FT_CHUNK;
  constexpr int NE = FT_STEP;
  const double* chunk = tape + (size_t)c * FT_CHUNK * NE;
  const int row2 = (N + S * N) / 2;   // double2 per step
  for (int idx = threadIdx.x; idx < cnt * row2; idx += blockDim.x) {
    const int q = idx / row2, o = idx - q * row2;
    *reinterpret_cast<double2*>(in_s + q * XP_ROW + 2 * o) = *reinterpret_cast<const double2*>(chunk + (size_t)q * NE + FT_Y0 + 2 * o);
  }
  __syncthreads();
  const int q = threadIdx.x / S, is = threadIdx.x - q * S;
  if (q >= cnt) return;
  const double* e = chunk + (size_t)q * NE;
  const double T = e[0], H = e[1];
  const double dDir = (rp.tend >= rp.tstart) ? 1.0 : -1.0;
  const double* in = in_s + q * XP_ROW;      // [Y0(32)][K(S x 32)]
  double Y0[N], Yi[N], Ki[N], DJ[cbm4_cell::NJT], JT0[cbm4_cell::NJT], PHT[Cell::NPH], PHS[Cell::NPH];
#pragma unroll
  for (int i = 0; i < N; ++i) { Y0[i] = in[i]; Ki[i] = in[N + is * N + i]; }
  // Y_i as the forward sweep formed it (same operations, same order): refreshed only at stages with ros_NewF,
  // stale otherwise (:1015-1018)
  int sr = is;
  while (sr > 0 && !rp.NewF[sr]) --sr;
#pragma unroll
  for (int i = 0; i < N; ++i) Yi[i] = Y0[i];
  for (int j = 0; j < sr; ++j) {
    const double a = rp.A[sr * (sr - 1) / 2 + j];
    if (a != 0.0) {
      const double* kj = in + N + j * N;
#pragma unroll
      for (int i = 0; i < N; ++i) Yi[i] = __dadd_rn(Yi[i], __dmul_rn(a, kj[i]));
    }
  }
  if (!rp.autonomous) {
    Cell::photo(T, mc, PHT);   // dJ/dt by the reference's forward difference (:1379-1382), DeltaMin of ros_DadjInt :1212
    const double Delta = sqrt(rp.roundoff) * fmax(1.0e-5, fabs(T));
    Cell::photo(T + Delta, mc, PHS);
    cbm4_cell::jac_time<1>(Y0, mc.fix, PHT, JT0);
    cbm4_cell::jac_time<1>(Y0, mc.fix, PHS, DJ);
    const double rd = 1.0 / Delta;
#pragma unroll
    for (int n = 0; n < cbm4_cell::NJT; ++n) DJ[n] = (DJ[n] - JT0[n]) * rd;
  } else {
#pragma unroll
    for (int n = 0; n < cbm4_cell::NJT; ++n) DJ[n] = 0.0;
  }
  double* sb = fat + ((size_t)(fo + (long long)chunk_seq[c] * FT_CHUNK + q) * S + is) * FT_STAGE;
  Cell::photo(T + rp.Alpha[is] * dDir * H, mc, PHS);
  cbm4_cell::jac_values<1>(Yi, mc.fix, PHS, sb + FT_JV);
  cbm4_cell::build_M<1>(Ki, DJ, (rp.autonomous || mode == 2) ? 0.0 : H * rp.Gamma[is], sb + FT_MV);
  if (mode == 1) cbm4_cell::param_terms<1>(Yi, Y0, mc.fix, Ki, sb + FT_AP);
  else if (mode == 2) {   // the tangent-linear sweep applies dJ/dt to the stage vector itself (LSS_Mul_Jac2, TLM :620)
    static_assert(cbm4_cell::NJT <= 32, "dJ/dt entries fit the AP slot");
#pragma unroll
    for (int p = 0; p < 32; p += 4)
      st_global_256(sb + FT_AP + p, p < cbm4_cell::NJT ? DJ[p] : 0.0, p + 1 < cbm4_cell::NJT ? DJ[p + 1] : 0.0,
                    p + 2 < cbm4_cell::NJT ? DJ[p + 2] : 0.0, p + 3 < cbm4_cell::NJT ? DJ[p + 3] : 0.0);
  } else
    for (int p = 0; p < 32; p += 4) st_global_256(sb + FT_AP + p, 0.0, 0.0, 0.0, 0.0);
}

// ---- private per-lane state of a backward warp: 256 doubles (W_1..W_5, Lambda, Mu, Lambda increment), addressed in
// 32-bit "columns" like tensor memory (double d = columns 2d, 2d+1).  Warps 0..3 of a CTA keep it in TENSOR MEMORY
// (their 32-lane quarter, all 512 columns); that is all the tensor memory an SM has, so any further warp keeps the same
// layout in SHARED memory ([double][lane], conflict-free).
struct PrivTmem {
  uint32_t tm;
  __device__ __forceinline__ void ld16(uint32_t col, double (&d)[16]) const { tmem::ld16(tm + col, d); }
  __device__ __forceinline__ void ld32(uint32_t col, double (&d)[32]) const { tmem::ld32(tm + col, d); }
  __device__ __forceinline__ void st16(uint32_t col, const double (&d)[16]) const { tmem::st16(tm + col, d); }
  __device__ __forceinline__ void st32(uint32_t col, const double (&d)[32]) const { tmem::st32(tm + col, d); }
  __device__ __forceinline__ void wait_st() const { tmem::wait_st(); }
};
struct PrivSmem {
  double* b;   // this lane's column: double d at b[d * 32]
  __device__ __forceinline__ void ld16(uint32_t col, double (&d)[16]) const {
#pragma unroll
    for (int r = 0; r < 16; ++r) d[r] = b[((col >> 1) + r) * 32];
  }
  __device__ __forceinline__ void ld32(uint32_t col, double (&d)[32]) const {
#pragma unroll
    for (int r = 0; r < 32; ++r) d[r] = b[((col >> 1) + r) * 32];
  }
  __device__ __forceinline__ void st16(uint32_t col, const double (&d)[16]) const {
#pragma unroll
    for (int r = 0; r < 16; ++r) b[((col >> 1) + r) * 32] = d[r];
  }
  __device__ __forceinline__ void st32(uint32_t col, const double (&d)[32]) const {
#pragma unroll
    for (int r = 0; r < 32; ++r) b[((col >> 1) + r) * 32] = d[r];
  }
  __device__ __forceinline__ void wait_st() const {}
};
constexpr int ADJ3_PRIV_DOUBLES = 256 * 32;   // shared-memory private state of one warp

template <class Model, int HALF, class Priv>
__device__ __forceinline__ void adj3_lambda_half(const double* __restrict__ MV, const double (&x)[32], const double (&v)[32],
                                                 bool first, bool last, const Priv& pv) {
  double acc[16];
  if (!first) pv.ld16(TM_LACC + 32 * HALF, acc);
  Model::mt_rows(MV, x, [&](auto rc, double h) {
    constexpr int r = decltype(rc)::value;
    if constexpr (r / 16 == HALF) {
      const double t = v[r] + h;
      acc[r - 16 * HALF] = first ? t : (acc[r - 16 * HALF] + t);
    }
  });
  if (last) {
    double lam[16];
    pv.ld16(TM_LAM + 32 * HALF, lam);
#pragma unroll
    for (int r = 0; r < 16; ++r) lam[r] += acc[r];
    pv.st16(TM_LAM + 32 * HALF, lam);
  } else {
    pv.st16(TM_LACC + 32 * HALF, acc);
  }
}

template <class Model, int HALF, class Priv>
__device__ __forceinline__ void adj3_mu_half(const double* __restrict__ AP, const double (&x)[32], const Priv& pv) {
  double mu[16];
  pv.ld16(TM_MU + 32 * HALF, mu);
  Model::stoich_cols(x, [&](auto pc, double s) {
    constexpr int p = decltype(pc)::value;
    if constexpr (p / 16 == HALF) mu[p - 16 * HALF] = fma(AP[p], s, mu[p - 16 * HALF]);
  });
  pv.st16(TM_MU + 32 * HALF, mu);
}

// one work item: stage `is` of the step whose per-step block is `SB`, stage block `STG`
template <class Model, class Priv>
__device__ __forceinline__ void adj3_consume(const RosParams& rp, const double* __restrict__ SB, const double* __restrict__ STG,
                                             int is, bool with_mu, const Priv& pv) {
  const int S = rp.S;
  const double* lu = SB + FT_HDR;
  const double* dinv = SB + FT_HDR + FT_LU;
  const double* JV = STG + FT_JV;
  const double* MV = STG + FT_MV;
  const double* AP = STG + FT_AP;
  const double dDir = (rp.tend >= rp.tstart) ? 1.0 : -1.0;
  const double H = SB[1];
  const unsigned swaps = (unsigned)SB[2];
  const bool first = (is == S - 1);
  const double mi = rp.M[is];

  double x[32];
  {   // x = m_i * lambda + W_i      (:1285-1300; W_i already holds sum_j a_ji V_j + c_ji/h U_j)
    pv.wait_st();
    pv.ld32(TM_LAM, x);
    if (!first) {
      double w[16];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        pv.ld16(TM_W + 64 * is + 32 * c, w);
#pragma unroll
        for (int r = 0; r < 16; ++r) x[16 * c + r] = fma(mi, x[16 * c + r], w[r]);
      }
    } else {
#pragma unroll
      for (int r = 0; r < 32; ++r) x[r] = mi * x[r];
    }
  }
  cbm4_cell::lut_solve(lu, dinv, swaps, x);   // LSS_Solve(Transp, U_i(m)) (:1301-1304) on the forward sweep's factors
  double v[32];
  Model::jt_vec(JV, x, v);                    // V_i = J(T_i,Y_i)^T U_i  (:1315)
  // running sums for the earlier stages j < is: W_j (+)= a_ij V_i + c_ij/h U_i   (:1292-1300)
  {
    const double inv_dH = 1.0 / (dDir * H);
    for (int j = 0; j < is; ++j) {
      const double ca = rp.A[is * (is - 1) / 2 + j];
      const double cc = rp.C[is * (is - 1) / 2 + j] * inv_dH;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        double w[16];
        if (!first) pv.ld16(TM_W + 64 * j + 32 * c, w);
#pragma unroll
        for (int r = 0; r < 16; ++r) {
          const double t = fma(ca, v[16 * c + r], cc * x[16 * c + r]);
          w[r] = first ? t : (w[r] + t);
        }
        pv.st16(TM_W + 64 * j + 32 * c, w);
      }
    }
  }
  // Lambda increment: LACC (+)= V_i + M'^T U_i     (:1328-1332 and the dJ/dt term :1402-1403); at the last stage of
  // the step Lambda += increment (:1324-1339 run after the stage loop).  Two halves of 16 rows bound the register use.
  adj3_lambda_half<Model, 0>(MV, x, v, first, is == 0, pv);
  adj3_lambda_half<Model, 1>(MV, x, v, first, is == 0, pv);
  if (with_mu) {   // Mu += a_p * (S^T U_i)_p     (:1352-1365); two halves of 16 parameters bound the register use
    adj3_mu_half<Model, 0>(AP, x, pv);
    adj3_mu_half<Model, 1>(AP, x, pv);
  }
}

struct AdjArgs {
  long nwave;
  const int* ids; const int* nacc; const int* last_chunk; const double* tape; const int* chunk_prev; const double* fat;
  const long long* fat_off; const int* ierr; const double* y_final; double* lambda; double* mu; int* istatus;
  unsigned long long* queue; const int* order; int nadj; int with_mu;
};

// the cell loop of one backward warp
template <class Model, class Priv>
__device__ __forceinline__ void adj3_warp_loop(const RosParams& rp, const AdjArgs& a, double* sm, const Priv& pv, int lane) {
  double* SBUF = sm;                                   // [2][FT_STEP_COPY]
  double* GBUF = sm + 2 * FT_STEP_COPY;                // [2][FT_STAGE]
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(sm + 2 * FT_STEP_COPY + 2 * FT_STAGE);   // [0,1] step, [2,3] stage
  const int S = rp.S;
  const bool with_mu = a.with_mu != 0;
  uint32_t ph_step0 = 0u, ph_step1 = 0u, ph_stage0 = 0u, ph_stage1 = 0u;   // phase parity per barrier (warp-uniform)
  for (;;) {
    unsigned long long q = 0;
    if (lane == 0) q = atomicAdd(a.queue, 1ull);
    q = __shfl_sync(FATODE_FULL_MASK, q, 0);
    if (q >= (unsigned long long)a.nwave) break;
    if (a.order) q = (unsigned long long)a.order[q];   // longest trajectories first
    const long cell = a.ids ? a.ids[q] : (long)q;
    if (a.ierr[cell] != 1) continue;
    const int nsteps = a.nacc[q];
    const int n_items = nsteps * S;
    {   // ADJINIT (model routine): lane = adjoint column m
      double lam[32], m0[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) lam[i] = Model::adjinit_lambda(i, lane, a.y_final[cell * 32 + i]);
#pragma unroll
      for (int p = 0; p < 32; ++p) m0[p] = 0.0;
      pv.st32(TM_LAM, lam);
      pv.st32(TM_MU, m0);
      pv.wait_st();
    }
    // prefetch cursor: walks the chunk list backwards, one work item ahead of the computation
    int pf_chunk = a.last_chunk[q];
    int pf_step = nsteps - 1;                      // forward-order index of the step the cursor is in
    const double* fat_cell = a.fat + (size_t)a.fat_off[q] * S * FT_STAGE;
    auto issue = [&](int t) {                      // executed by lane 0
      const int si = t / S, is = S - 1 - (t - si * S);
      const double* ent = a.tape + ((size_t)pf_chunk * FT_CHUNK + (pf_step % FT_CHUNK)) * FT_STEP;
      if (is == S - 1) {
        mbar::expect_tx(bars + (si & 1), FT_STEP_COPY * 8);
        mbar::bulk_g2s(SBUF + (si & 1) * FT_STEP_COPY, ent, FT_STEP_COPY * 8, bars + (si & 1));
      }
      mbar::expect_tx(bars + 2 + (t & 1), FT_STAGE * 8);
      mbar::bulk_g2s(GBUF + (t & 1) * FT_STAGE, fat_cell + ((size_t)pf_step * S + is) * FT_STAGE, FT_STAGE * 8, bars + 2 + (t & 1));
    };
    auto advance = [&](int t) {                    // cursor moves from item t to item t + 1 (all lanes, uniform)
      if ((t + 1) % S == 0) {
        if (pf_step % FT_CHUNK == 0) pf_chunk = a.chunk_prev[pf_chunk];
        --pf_step;
      }
    };
    if (n_items > 0 && lane == 0) issue(0);
    for (int t = 0; t < n_items; ++t) {
      const int si = t / S, is = S - 1 - (t - si * S);
      __syncwarp();                                // every lane is done with the buffers item t + 1 will overwrite
      if (t + 1 < n_items) {
        advance(t);
        if (lane == 0) issue(t + 1);
      }
      if (is == S - 1) {
        if (si & 1) { mbar::wait(bars + 1, ph_step1); ph_step1 ^= 1u; }
        else { mbar::wait(bars + 0, ph_step0); ph_step0 ^= 1u; }
      }
      if (t & 1) { mbar::wait(bars + 3, ph_stage1); ph_stage1 ^= 1u; }
      else { mbar::wait(bars + 2, ph_stage0); ph_stage0 ^= 1u; }
      adj3_consume<Model, Priv>(rp, SBUF + (si & 1) * FT_STEP_COPY, GBUF + (t & 1) * FT_STAGE, is, with_mu, pv);
    }
    // results: column m of Lambda / Mu is this lane's private vector
    pv.wait_st();
    {
      double lam[32];
      pv.ld32(TM_LAM, lam);   // warp-collective for tensor memory: executed by the whole warp, outside any lane-dependent branch
      if (lane < a.nadj) {
        double* dst = a.lambda + ((long)cell * a.nadj + lane) * 32;
#pragma unroll
        for (int i = 0; i < 32; i += 2) *reinterpret_cast<double2*>(dst + i) = make_double2(lam[i], lam[i + 1]);
      }
    }
    if (with_mu) {
      double m0[32];
      pv.ld32(TM_MU, m0);
      if (lane < a.nadj) {
        double* dst = a.mu + ((long)cell * a.nadj + lane) * Model::NP;
#pragma unroll
        for (int p = 0; p < Model::NP; ++p) dst[p] = m0[p];
      }
    }
    if (lane == 0) {   // counters of the backward sweep (ADJ :1257-1267, :1303, :1312, :1381)
      int* is_ = a.istatus + cell * 20;
      is_[Nstp] += nsteps;
      is_[Njac] += nsteps * (2 + S + (rp.autonomous ? 0 : 1));
      is_[Ndec] += nsteps;
      is_[Nsol] += nsteps * S * a.nadj;
    }
    __syncwarp();
  }
}

// ids[w]: cell of wave slot w; nacc[w], last_chunk[w]: its tape; order[]: processing order of the slots.
// Only slots with ierr[cell] == 1 are processed.  Dynamic shared memory: ADJ3_WARPS stream buffers, then the private
// state of the warps beyond the four that tensor memory serves.
template <class Model>
__global__ void __launch_bounds__(ADJ3_WARPS * 32, 1)
k_ros_adj_cell(RosParams rp, AdjArgs a) {
  extern __shared__ __align__(16) double smem[];
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* sm = smem + (size_t)warp * ADJ3_WARP_DOUBLES;
  if (warp == 0) tmem::alloc512(&tmem_base_s);
  if (lane == 0) {
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(sm + 2 * FT_STEP_COPY + 2 * FT_STAGE);
    for (int b = 0; b < 4; ++b) mbar::init(bars + b, 1);
    mbar::fence_init();
  }
  tmem::fence_before_sync();
  __syncthreads();
  tmem::fence_after_sync();
  if (warp < 4) {
    PrivTmem pv{tmem_base_s + ((uint32_t)(warp * 32) << 16)};
    adj3_warp_loop<Model, PrivTmem>(rp, a, sm, pv, lane);
  } else if constexpr (ADJ3_WARPS > 4) {
    PrivSmem pv{smem + (size_t)ADJ3_WARPS * ADJ3_WARP_DOUBLES + (size_t)(warp - 4) * ADJ3_PRIV_DOUBLES + lane};
    adj3_warp_loop<Model, PrivSmem>(rp, a, sm, pv, lane);
  }
  tmem::fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem::dealloc512(tmem_base_s);
}


// =====================================================================================================================
// Tangent linear model, ROS_TLM with forward error control (ICNTRL(12) = 0): ros_TLM_Int
// (TLM/ROS_TLM/ROS_TLM_f90_Integrator.F90:462-707) split into the state sweep (k_ros_fwd_cell, identical step sequence
// because the TLM does not enter the error estimate) and this sweep over the tape in FORWARD order.
// Lane m of a warp owns column m of Y_tlm of one cell.  Private per-lane state in tensor memory: K_tlm[6][32]
// (columns 0..383), the stage function Fcn_tlm (384..447) and Y_tlm itself (448..511).  Per stage i:
//   Ynew = Y_tlm (+ sum_j a_ij K_tlm_j at ros_NewF stages);  Fcn_tlm = J(T_i,Y_i) Ynew                        (:588-604)
//   k = Fcn_tlm + sum_j c_ij/(dir h) K_tlm_j + h gamma_i (dJ/dt) Ynew + Hess_Vec(K_i, Y_tlm);  k <- E^-1 k     (:606-633)
// and after the stages Y_tlm += sum_j m_j K_tlm_j (:642-648).  The operands J_i, Hess x K_i, dJ/dt come from
// k_expand_tape (mode 2), the L\U factors from the forward sweep.
constexpr uint32_t TMT_K = 0, TMT_F = 384, TMT_Y = 448;

template <class Model>
__device__ __forceinline__ void tlm_consume(const RosParams& rp, const double* __restrict__ SB, const double* __restrict__ STG,
                                            int is, const PrivTmem& pv) {
  const int S = rp.S;
  const double* lu = SB + FT_HDR;
  const double* dinv = SB + FT_HDR + FT_LU;
  const double* JV = STG + FT_JV;
  const double* MV = STG + FT_MV;
  const double* DJ = STG + FT_AP;
  const double dDir = (rp.tend >= rp.tstart) ? 1.0 : -1.0;
  const double H = SB[1];
  const unsigned swaps = (unsigned)SB[2];
  pv.wait_st();
  double yn[32], k[32];
  pv.ld32(TMT_Y, yn);
  if (is == 0 || rp.NewF[is]) {
    for (int j = 0; j < is; ++j) {
      const double a = rp.A[is * (is - 1) / 2 + j];
      if (a != 0.0) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          double w[16];
          pv.ld16(TMT_K + 64 * j + 32 * c, w);
#pragma unroll
          for (int r = 0; r < 16; ++r) yn[16 * c + r] = fma(a, w[r], yn[16 * c + r]);
        }
      }
    }
    cbm4_cell::j_vec(JV, yn, k);          // Fcn_tlm = J(T_i, Y_i) Ynew_tlm
    pv.st32(TMT_F, k);
  } else {
    pv.ld32(TMT_F, k);                    // stage without a new function evaluation: previous Fcn_tlm, Ynew = Y_tlm
  }
  {
    const double inv_dH = 1.0 / (dDir * H);
    for (int j = 0; j < is; ++j) {
      const double hc = rp.C[is * (is - 1) / 2 + j] * inv_dH;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        double w[16];
        pv.ld16(TMT_K + 64 * j + 32 * c, w);
#pragma unroll
        for (int r = 0; r < 16; ++r) k[16 * c + r] = fma(hc, w[r], k[16 * c + r]);
      }
    }
  }
  if (!rp.autonomous && rp.Gamma[is] != 0.0) cbm4_cell::dj_vec_add(DJ, dDir * H * rp.Gamma[is], yn, k);
  pv.ld32(TMT_Y, yn);                     // Hess_Vec(T, Y, K_i, Y_tlm): the step's Y_tlm, not the stage vector
  cbm4_cell::m_vec_add(MV, yn, k);
  cbm4_cell::lun_solve(lu, dinv, swaps, k);
  pv.st32(TMT_K + 64 * is, k);
  if (is == S - 1) {                      // new solution (:642-648)
    pv.wait_st();
    for (int j = 0; j < S; ++j) {
      const double m = rp.M[j];
      if (m != 0.0) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          double w[16];
          pv.ld16(TMT_K + 64 * j + 32 * c, w);
#pragma unroll
          for (int r = 0; r < 16; ++r) yn[16 * c + r] = fma(m, w[r], yn[16 * c + r]);
        }
      }
    }
    pv.st32(TMT_Y, yn);
  }
}

struct TlmArgs {
  long nwave;
  const int* ids; const int* nacc; const int* first_chunk; const double* tape; const int* chunk_next; const double* fat;
  const long long* fat_off; const int* ierr; double* y_tlm; int* istatus; unsigned long long* queue; const int* order; int ntlm;
};

template <class Model>
__global__ void __launch_bounds__(ADJ3_WARPS * 32, 1)
k_ros_tlm_cell(RosParams rp, TlmArgs a) {
  extern __shared__ __align__(16) double smem[];
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* sm = smem + (size_t)warp * ADJ3_WARP_DOUBLES;
  double* SBUF = sm;
  double* GBUF = sm + 2 * FT_STEP_COPY;
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(sm + 2 * FT_STEP_COPY + 2 * FT_STAGE);
  if (warp == 0) tmem::alloc512(&tmem_base_s);
  if (lane == 0) {
    for (int b = 0; b < 4; ++b) mbar::init(bars + b, 1);
    mbar::fence_init();
  }
  tmem::fence_before_sync();
  __syncthreads();
  tmem::fence_after_sync();
  const PrivTmem pv{tmem_base_s + ((uint32_t)(warp * 32) << 16)};
  const int S = rp.S;
  uint32_t ph_step0 = 0u, ph_step1 = 0u, ph_stage0 = 0u, ph_stage1 = 0u;
  for (;;) {
    unsigned long long q = 0;
    if (lane == 0) q = atomicAdd(a.queue, 1ull);
    q = __shfl_sync(FATODE_FULL_MASK, q, 0);
    if (q >= (unsigned long long)a.nwave) break;
    if (a.order) q = (unsigned long long)a.order[q];
    const long cell = a.ids ? a.ids[q] : (long)q;
    if (!sweep_takes_cell(a.ierr[cell], true)) continue;
    const int nsteps = a.nacc[q];
    const int n_items = nsteps * S;
    double* ycol = a.y_tlm + ((long)cell * a.ntlm + (lane < a.ntlm ? lane : 0)) * 32;
    {   // Y_tlm(:, m) of the caller
      double y0[32];
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        const double2 v = *reinterpret_cast<const double2*>(ycol + i);
        y0[i] = lane < a.ntlm ? v.x : 0.0;
        y0[i + 1] = lane < a.ntlm ? v.y : 0.0;
      }
      pv.st32(TMT_Y, y0);
      pv.wait_st();
    }
    int pf_chunk = a.first_chunk[q];
    int pf_step = 0;
    const double* fat_cell = a.fat + (size_t)a.fat_off[q] * S * FT_STAGE;
    auto issue = [&](int t) {                      // executed by lane 0
      const int si = t / S, is = t - si * S;
      const double* ent = a.tape + ((size_t)pf_chunk * FT_CHUNK + (pf_step % FT_CHUNK)) * FT_STEP;
      if (is == 0) {
        mbar::expect_tx(bars + (si & 1), FT_STEP_COPY * 8);
        mbar::bulk_g2s(SBUF + (si & 1) * FT_STEP_COPY, ent, FT_STEP_COPY * 8, bars + (si & 1));
      }
      mbar::expect_tx(bars + 2 + (t & 1), FT_STAGE * 8);
      mbar::bulk_g2s(GBUF + (t & 1) * FT_STAGE, fat_cell + ((size_t)pf_step * S + is) * FT_STAGE, FT_STAGE * 8, bars + 2 + (t & 1));
    };
    auto advance = [&](int t) {                    // cursor moves from item t to item t + 1 (all lanes, uniform)
      if ((t + 1) % S == 0) {
        ++pf_step;
        if (pf_step % FT_CHUNK == 0) pf_chunk = a.chunk_next[pf_chunk];
      }
    };
    if (n_items > 0 && lane == 0) issue(0);
    for (int t = 0; t < n_items; ++t) {
      const int si = t / S, is = t - si * S;
      __syncwarp();
      if (t + 1 < n_items) {
        advance(t);
        if (lane == 0) issue(t + 1);
      }
      if (is == 0) {
        if (si & 1) { mbar::wait(bars + 1, ph_step1); ph_step1 ^= 1u; }
        else { mbar::wait(bars + 0, ph_step0); ph_step0 ^= 1u; }
      }
      if (t & 1) { mbar::wait(bars + 3, ph_stage1); ph_stage1 ^= 1u; }
      else { mbar::wait(bars + 2, ph_stage0); ph_stage0 ^= 1u; }
      tlm_consume<Model>(rp, SBUF + (si & 1) * FT_STEP_COPY, GBUF + (t & 1) * FT_STAGE, is, pv);
    }
    pv.wait_st();
    {
      double y1[32];
      pv.ld32(TMT_Y, y1);
      if (lane < a.ntlm) {
#pragma unroll
        for (int i = 0; i < 32; i += 2) *reinterpret_cast<double2*>(ycol + i) = make_double2(y1[i], y1[i + 1]);
      }
    }
    if (lane == 0) {   // counters of ros_TLM_Int on top of the state sweep's (:543-562, 600-601, 625-632)
      int* is_ = a.istatus + cell * 20;
      int n_newf = 0;
      for (int s = 1; s < S; ++s) n_newf += rp.NewF[s] ? 1 : 0;
      // the state sweep counted one Jacobian per step set-up and one solve per stage: LSS_Jac2 doubles the former when
      // non-autonomous, LSS_Jac1 adds one per new-function stage of every attempt, every TLM column adds a solve
      is_[Njac] = is_[Njac] * (rp.autonomous ? 1 : 2) + is_[Nstp] * n_newf;
      is_[Nsol] = is_[Nsol] * (1 + a.ntlm);
    }
    __syncwarp();
  }
  tmem::fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem::dealloc512(tmem_base_s);
}

}  // namespace fatode
