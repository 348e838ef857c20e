// Discrete Rosenbrock adjoint, backward sweep, TWO WARPS PER CELL (round 2; replaces k_ros_adj_cell on the default path).
//
// k_ros_adj_cell (ros_cell_adj.cuh) gives one warp to a cell (lane m = adjoint column m) and can keep only four cells
// per SM — the per-lane running sums fill the tensor memory — so each warp scheduler sees ONE warp: the FP64 pipe was
// 30 % busy, every dependent instruction waited out its full latency (profiles/README.md, r1i).  The arithmetic of a
// stage, ros_DadjInt (ADJ/ROS_ADJ/ROS_ADJ_f90_Integrator.F90:1277-1365), splits into two halves of about equal size that
// only meet through the 64 numbers U_i, V_i per column:
//
//   solver warp  (warp p,   p = 0..3):  U_i = m_i Lambda + sum_{j>i} (a_ji V_j + c_ji/h U_j)      (:1285-1300)
//                                       U_i <- (PLU)^-T U_i  on the forward sweep's factors      (:1301-1304)
//                                       V_i = J(T_i,Y_i)^T U_i                                   (:1315)
//   accumulator warp (warp p + 4):      Lambda increment += V_i + (Hess x K_i + h gamma_i dJ/dt)^T U_i   (:1328-1332, :1402)
//                                       Mu += a_p (S^T U_i)_p                                    (:1352-1365)
//                                       W_j += a_ij V_i + c_ij/h U_i   for the stages j < i - 1
//
// Both warps of a pair sit on the same scheduler (warp id mod 4) and on the same quarter of tensor memory, lane m of
// either warp is adjoint column m: the accumulator owns the running sums W_j, Lambda, the Lambda increment and Mu in the
// lane's 512 columns, the solver only reads Lambda and W_i there.  The accumulator runs one stage behind the solver:
// U_i, V_i travel through a double-buffered shared-memory block ([row][lane], conflict-free), the solver adds the
// contribution of the stage it has just finished (j = i + 1) from its own registers, so the sum W_i it reads needs the
// accumulator to be done with stage i + 2 only — two stages of slack.  Per step the solver waits once for the new Lambda.
// That puts two independent instruction streams on every scheduler (latency hiding the single-warp kernel could not have),
// drops the W_j read-modify-write of the newest stage (10 instead of 15 per step) and the last W vector.
//
// Synchronisation per pair (shared-memory mbarriers, 32 arrivals each; tensor-memory accesses bracketed by
// tcgen05.fence::before/after_thread_sync):
//   full[t&1]   solver -> accumulator   U_i, V_i of work item t are in XBUF[t&1]
//   bdone[t&1]  accumulator -> solver   item t consumed: XBUF[t&1] is free and every W_j holds the stages up to item t
//   lam         accumulator -> solver   Lambda of the finished step is in tensor memory
// Tape streaming as before (1-D bulk TMA + complete_tx mbarriers, one item ahead): the solver loads the step block (L\U,
// 1/U_kk, H, interchange flags) and J_i, the accumulator loads M'_i and a_p of the same stage block.
#pragma once
#include "ros_cell_adj.cuh"

namespace fatode {

constexpr int ADJP_PAIRS = 4;                            // cells in flight per SM (tensor memory: 4 x 32 lanes x 512 columns)
constexpr int ADJP_XB = 64 * 32 + 8;                     // U_i, V_i [64 rows][32 lanes] + 1/(dir h) (+pad)
constexpr int FT_MAP = FT_STAGE - FT_MV;                 // M'_i and a_p of a stage block: 292 doubles = 2336 B
constexpr int ADJP_PAIR_DOUBLES = 2 * FT_STEP_COPY + 2 * FT_MV + 2 * FT_MAP + 2 * ADJP_XB + 12;   // + 11 mbarriers
static_assert((FT_MV * 8) % 16 == 0 && (FT_MAP * 8) % 16 == 0 && (ADJP_XB * 8) % 16 == 0, "bulk copies are 16-byte granular");

namespace mbar {
__device__ __forceinline__ void arrive(void* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(saddr(bar)) : "memory");
}
}  // namespace mbar

__device__ __forceinline__ void pair_sync(int pair) {   // the 64 threads of one solver / accumulator pair
  asm volatile("bar.sync %0, 64;\n" ::"r"(pair + 1) : "memory");
}

struct AdjPairSmem {
  double* SBUF; double* JBUF; double* MBUF; double* XBUF;
  unsigned long long* astep; unsigned long long* ajv; unsigned long long* bmv; unsigned long long* full;
  unsigned long long* bdone; unsigned long long* lam;
  __device__ __forceinline__ explicit AdjPairSmem(double* sm) {
    SBUF = sm;
    JBUF = SBUF + 2 * FT_STEP_COPY;
    MBUF = JBUF + 2 * FT_MV;
    XBUF = MBUF + 2 * FT_MAP;
    astep = reinterpret_cast<unsigned long long*>(XBUF + 2 * ADJP_XB);
    ajv = astep + 2; bmv = astep + 4; full = astep + 6; bdone = astep + 8; lam = astep + 10;
  }
};

// ---- solver warp: work items of one cell ---------------------------------------------------------------------------
template <class Model>
__device__ __forceinline__ void adjp_solver(const RosParams& rp, const AdjArgs& a, const AdjPairSmem& s, uint32_t tm, int lane,
                                            long q, int nsteps, uint32_t& ph) {
  // ph: bit k = parity of the next completion to wait for on barrier k: 0,1 astep, 2,3 ajv, 4,5 bdone, 6 lam
  const int S = rp.S;
  const int n_items = nsteps * S;
  const double dDir = (rp.tend >= rp.tstart) ? 1.0 : -1.0;
  int pf_chunk = a.last_chunk[q];
  int pf_step = nsteps - 1;                      // forward-order index of the step the prefetch cursor is in
  const double* fat_cell = a.fat + (size_t)a.fat_off[q] * S * FT_STAGE;
  auto issue = [&](int t) {                      // executed by lane 0
    const int si = t / S, is = S - 1 - (t - si * S);
    if (is == S - 1) {
      const double* ent = a.tape + ((size_t)pf_chunk * FT_CHUNK + (pf_step % FT_CHUNK)) * FT_STEP;
      mbar::expect_tx(s.astep + (si & 1), FT_STEP_COPY * 8);
      mbar::bulk_g2s(s.SBUF + (si & 1) * FT_STEP_COPY, ent, FT_STEP_COPY * 8, s.astep + (si & 1));
    }
    mbar::expect_tx(s.ajv + (t & 1), FT_MV * 8);
    mbar::bulk_g2s(s.JBUF + (t & 1) * FT_MV, fat_cell + ((size_t)pf_step * S + is) * FT_STAGE + FT_JV, FT_MV * 8, s.ajv + (t & 1));
  };
  auto advance = [&](int t) {                    // cursor moves from item t to item t + 1 (all lanes, uniform)
    if ((t + 1) % S == 0) {
      if (pf_step % FT_CHUNK == 0) pf_chunk = a.chunk_prev[pf_chunk];
      --pf_step;
    }
  };
  auto wait_on = [&](void* bar, int k) { mbar::wait(bar, (ph >> k) & 1u); ph ^= 1u << k; };
  double x[32], v[32];
#pragma unroll
  for (int r = 0; r < 32; ++r) { x[r] = 0.0; v[r] = 0.0; }
  if (n_items > 0 && lane == 0) issue(0);
  for (int t = 0; t < n_items; ++t) {
    const int si = t / S, is = S - 1 - (t - si * S);
    __syncwarp();                                // every lane is done with the buffers item t + 1 will overwrite
    if (t + 1 < n_items) {
      advance(t);
      if (lane == 0) issue(t + 1);
    }
    if (t >= 2) wait_on(s.bdone + (t & 1), 4 + (t & 1));   // XBUF[t&1] free, W complete up to item t-2
    if (is == S - 1) {
      if (si >= 1) wait_on(s.lam, 6);                  // Lambda of the previous step
      wait_on(s.astep + (si & 1), si & 1);
    }
    wait_on(s.ajv + (t & 1), 2 + (t & 1));
    tmem::fence_after_sync();
    const double* SB = s.SBUF + (si & 1) * FT_STEP_COPY;
    const double inv_dH = 1.0 / (dDir * SB[1]);
    const unsigned swaps = (unsigned)SB[2];
    const double mi = rp.M[is];
    // U_i = m_i Lambda + W_i + a_{i+1,i} V_{i+1} + c_{i+1,i}/h U_{i+1}     (:1285-1300; the last two from registers)
    if (is == S - 1) {
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        double l[16];
        tmem::ld16(tm + TM_LAM + 32 * c, l);
#pragma unroll
        for (int r = 0; r < 16; ++r) x[16 * c + r] = mi * l[r];
      }
    } else {
      const double ca = rp.A[(is + 1) * is / 2 + is];
      const double cc = rp.C[(is + 1) * is / 2 + is] * inv_dH;
      if (is < S - 2) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          double l[16], w[16];
          tmem::ld16(tm + TM_LAM + 32 * c, l);
          tmem::ld16(tm + TM_W + 64 * is + 32 * c, w);
#pragma unroll
          for (int r = 0; r < 16; ++r) x[16 * c + r] = fma(ca, v[16 * c + r], fma(cc, x[16 * c + r], fma(mi, l[r], w[r])));
        }
      } else {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          double l[16];
          tmem::ld16(tm + TM_LAM + 32 * c, l);
#pragma unroll
          for (int r = 0; r < 16; ++r) x[16 * c + r] = fma(ca, v[16 * c + r], fma(cc, x[16 * c + r], mi * l[r]));
        }
      }
    }
    // LSS_Solve(Transp, U_i(m)) (:1301-1304); the dominant interchange outcome has its own (shorter, bit-identical) code
    if (swaps == cbm4_cell::LUT_DOM_SWAPS) cbm4_cell::lut_solve_dom(SB + FT_HDR, SB + FT_HDR + FT_LU, x);
    else cbm4_cell::lut_solve(SB + FT_HDR, SB + FT_HDR + FT_LU, swaps, x);
    Model::jt_vec1(s.JBUF + (t & 1) * FT_MV, x, v);                       // V_i = J(T_i,Y_i)^T U_i (:1315)
    double* XB = s.XBUF + (t & 1) * ADJP_XB;
#pragma unroll
    for (int r = 0; r < 32; ++r) { XB[r * 32 + lane] = x[r]; XB[(32 + r) * 32 + lane] = v[r]; }
    if (lane == 0) XB[64 * 32] = inv_dH;
    tmem::fence_before_sync();
    mbar::arrive(s.full + (t & 1));
  }
  // drain: every arrival of the accumulator is matched by one wait, so the phase bits stay aligned across cells
  for (int t = n_items; t < n_items + 2; ++t)
    if (t >= 2) wait_on(s.bdone + (t & 1), 4 + (t & 1));
  if (nsteps >= 1) wait_on(s.lam, 6);
}

// ---- accumulator warp ----------------------------------------------------------------------------------------------
// zero the Lambda increment and the running sums W_0..W_{S-3} of a lane (start of a step)
__device__ __forceinline__ void adjp_clear_sums(uint32_t tm, int S) {
  double z[32];
#pragma unroll
  for (int r = 0; r < 32; ++r) z[r] = 0.0;
  tmem::st32(tm + TM_LACC, z);
  for (int j = 0; j + 2 < S; ++j) tmem::st32(tm + TM_W + 64 * j, z);
}

// Lambda increment and the running sums of one work item (accumulator warp).  One instantiation of the generated sparse
// rows per kernel: the code of the two warps of a pair has to share the instruction caches of one scheduler.
template <class Model>
__device__ __forceinline__ void adjp_accumulate(const RosParams& rp, const double* __restrict__ MV, const double (&u)[32],
                                                const double* __restrict__ vs, double inv_dH, int is, uint32_t tm) {
  // vs: V_i of this lane in the exchange block, row r at vs[r * 32] (read where it is used: u and the 32 sums fill the registers)
  {   // Lambda increment: LACC += V_i + M'^T U_i; at the last stage of the step Lambda += increment   (:1324-1339, :1402)
    // No first-stage special case: LACC and the W_j are ZERO when a step begins (cleared below and at ADJINIT with
    // register-free stores of RZ), so every item runs the same straight-line code.
    double acc[32];
    tmem::ld32(tm + TM_LACC, acc);
    Model::mt_rows(MV, u, [&](auto rc, double h) {
      constexpr int r = decltype(rc)::value;
      acc[r] += vs[r * 32] + h;
    });
    if (is == 0) {
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        double lam[16];
        tmem::ld16(tm + TM_LAM + 32 * c, lam);
#pragma unroll
        for (int r = 0; r < 16; ++r) lam[r] += acc[16 * c + r];
        tmem::st16(tm + TM_LAM + 32 * c, lam);
      }
      adjp_clear_sums(tm, rp.S);
    } else {
      tmem::st32(tm + TM_LACC, acc);
    }
  }
  // running sums of the earlier stages except the next one (the solver adds that contribution itself):
  // W_j += a_ij V_i + c_ij/h U_i, j < i - 1     (:1292-1300)
  for (int j = 0; j + 1 < is; ++j) {
    const double ca = rp.A[is * (is - 1) / 2 + j];
    const double cc = rp.C[is * (is - 1) / 2 + j] * inv_dH;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      double w[16];
      tmem::ld16(tm + TM_W + 64 * j + 32 * c, w);
#pragma unroll
      for (int r = 0; r < 16; ++r) w[r] = fma(ca, vs[(16 * c + r) * 32], fma(cc, u[16 * c + r], w[r]));
      tmem::st16(tm + TM_W + 64 * j + 32 * c, w);
    }
  }
}

template <class Model>
__device__ __forceinline__ void adjp_mu(const double* __restrict__ AP, const double (&u)[32], uint32_t tm) {
  double mu[32];
  tmem::ld32(tm + TM_MU, mu);
  Model::stoich_cols(u, [&](auto pc, double sp) {
    constexpr int p = decltype(pc)::value;
    mu[p] = fma(AP[p], sp, mu[p]);
  });
  tmem::st32(tm + TM_MU, mu);
}

template <class Model>
__device__ __forceinline__ void adjp_accumulator(const RosParams& rp, const AdjArgs& a, const AdjPairSmem& s, uint32_t tm, int lane,
                                                 long q, long cell, int nsteps, uint32_t& ph) {
  // ph: bit k = parity of the next completion to wait for: 0,1 bmv, 2,3 full
  const int S = rp.S;
  const int n_items = nsteps * S;
  const bool with_mu = a.with_mu != 0;
  const double* fat_cell = a.fat + (size_t)a.fat_off[q] * S * FT_STAGE;
  auto issue = [&](int t) {                      // executed by lane 0
    const int si = t / S, is = S - 1 - (t - si * S);
    mbar::expect_tx(s.bmv + (t & 1), FT_MAP * 8);
    mbar::bulk_g2s(s.MBUF + (t & 1) * FT_MAP, fat_cell + ((size_t)(nsteps - 1 - si) * S + is) * FT_STAGE + FT_MV, FT_MAP * 8,
                   s.bmv + (t & 1));
  };
  if (n_items > 0 && lane == 0) issue(0);
  for (int t = 0; t < n_items; ++t) {
    const int si = t / S, is = S - 1 - (t - si * S);
    __syncwarp();                                // every lane is done with MBUF[(t+1)&1] (item t - 1)
    if (t + 1 < n_items && lane == 0) issue(t + 1);
    { const int k = t & 1; mbar::wait(s.bmv + k, (ph >> k) & 1u); ph ^= 1u << k; }
    { const int k = 2 + (t & 1); mbar::wait(s.full + (t & 1), (ph >> k) & 1u); ph ^= 1u << k; }
    tmem::fence_after_sync();
    const double* XB = s.XBUF + (t & 1) * ADJP_XB;
    const double* MV = s.MBUF + (t & 1) * FT_MAP;
    const double* AP = MV + (FT_AP - FT_MV);
    const double inv_dH = XB[64 * 32];
    double u[32];
#pragma unroll
    for (int r = 0; r < 32; ++r) u[r] = XB[r * 32 + lane];
    tmem::wait_st();                             // this warp's stores of the previous item are in place
    adjp_accumulate<Model>(rp, MV, u, XB + 32 * 32 + lane, inv_dH, is, tm);
    tmem::wait_st();
    tmem::fence_before_sync();
    if (is == 0) mbar::arrive(s.lam);            // Lambda of this step is complete
    mbar::arrive(s.bdone + (t & 1));             // XBUF[t&1] consumed, W_j up to date
    if (with_mu) {   // Mu += a_p * (S^T U_i)_p     (:1352-1365)
      adjp_mu<Model>(AP, u, tm);
    }
  }
  // results: column m of Lambda / Mu is this lane's private vector
  tmem::wait_st();
  {
    double lam[32];
    tmem::ld32(tm + TM_LAM, lam);   // warp-collective: executed by the whole warp, outside any lane-dependent branch
    if (lane < a.nadj) {
      double* dst = a.lambda + ((long)cell * a.nadj + lane) * 32;
#pragma unroll
      for (int i = 0; i < 32; i += 2) *reinterpret_cast<double2*>(dst + i) = make_double2(lam[i], lam[i + 1]);
    }
  }
  if (with_mu) {
    double m0[32];
    tmem::ld32(tm + TM_MU, m0);
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
}

// ids[w]: cell of wave slot w; nacc[w], last_chunk[w]: its tape; order[]: processing order of the slots.  Only slots with
// ierr[cell] == 1 are processed.  256 threads: warps 0..3 solvers, 4..7 accumulators; one CTA per SM.
template <class Model>
__global__ void __launch_bounds__(ADJP_PAIRS * 64, 1)
k_ros_adj_pair(RosParams rp, AdjArgs a) {
  extern __shared__ __align__(16) double smem[];
  __shared__ uint32_t tmem_base_s;
  __shared__ unsigned long long mail[ADJP_PAIRS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pair = warp & 3;
  const bool solver = warp < 4;
  const AdjPairSmem s(smem + (size_t)pair * ADJP_PAIR_DOUBLES);
  if (warp == 0) tmem::alloc512(&tmem_base_s);
  if (solver && lane == 0) {
    for (int b = 0; b < 6; ++b) mbar::init(s.astep + b, 1);     // transaction barriers: one arrive.expect_tx per use
    for (int b = 6; b < 11; ++b) mbar::init(s.astep + b, 32);   // full, bdone, lam: every lane arrives
    mbar::fence_init();
  }
  tmem::fence_before_sync();
  __syncthreads();
  tmem::fence_after_sync();
  const uint32_t tm = tmem_base_s + ((uint32_t)(pair * 32) << 16);
  uint32_t phase = 0u;   // mbarrier phase parities of this warp's waits, carried across cells
  for (;;) {
    if (solver && lane == 0) mail[pair] = atomicAdd(a.queue, 1ull);
    pair_sync(pair);                             // mailbox written; the accumulator is done with the previous cell
    unsigned long long q = *reinterpret_cast<volatile unsigned long long*>(mail + pair);
    if (q >= (unsigned long long)a.nwave) break;
    if (a.order) q = (unsigned long long)a.order[q];   // longest trajectories first
    const long cell = a.ids ? a.ids[q] : (long)q;
    const bool take = a.ierr[cell] == 1;
    const int nsteps = a.nacc[q];
    if (take && !solver) {   // ADJINIT (model routine): lane = adjoint column m
      double lam[32], m0[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) lam[i] = Model::adjinit_lambda(i, lane, a.y_final[cell * 32 + i]);
#pragma unroll
      for (int p = 0; p < 32; ++p) m0[p] = 0.0;
      tmem::st32(tm + TM_LAM, lam);
      tmem::st32(tm + TM_MU, m0);
      adjp_clear_sums(tm, rp.S);
      tmem::wait_st();
      tmem::fence_before_sync();
    }
    pair_sync(pair);                             // both warps have read the mailbox; ADJINIT is in tensor memory
    if (!take) continue;
    if (solver) adjp_solver<Model>(rp, a, s, tm, lane, (long)q, nsteps, phase);
    else adjp_accumulator<Model>(rp, a, s, tm, lane, (long)q, cell, nsteps, phase);
  }
  tmem::fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem::dealloc512(tmem_base_s);
}

}  // namespace fatode
