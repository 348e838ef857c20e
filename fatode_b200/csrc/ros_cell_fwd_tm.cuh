// Rosenbrock forward sweep, one cell per thread, with the stage vectors in TENSOR MEMORY: two warps (58 cells) per SM.
//
// Same algorithm, same arithmetic and the same tape as k_ros_fwd_cell (ros_cell_fwd.cuh; reference: ros_FwdInt
// ADJ/ROS_ADJ/ROS_ADJ_f90_Integrator.F90:877-1173, ros_Integrator FWD/ROS/ROS_f90_Integrator.F90:433-629,
// ros_TLM_Int's state part TLM/ROS_TLM/ROS_TLM_f90_Integrator.F90:462-707).  What changes is where a cell's data lives.
//
// k_ros_fwd_cell keeps 751 doubles (6 KB) per cell in shared memory, so one warp of 32 cells fills an SM and three of
// its four schedulers idle; the kernel is latency bound (one instruction per ~10 cycles, profiles/README.md).  Here the
// vectors that only the kernel's own vector code touches — the stage vectors K_1..K_6, Fcn0 = f(T,Y) and dF/dT, 256
// doubles per cell — live in the thread's tensor-memory lane (512 32-bit columns of TMEM lane 32*warp + lane, moved 16
// doubles per tcgen05.ld/st), and what the generated model / linear-algebra code addresses (Y, Ynew, Fcn, 1/U_kk and the
// 367 L\U values: 495 doubles = 3.9 KB) stays in shared memory.  29 cells per warp, two warps per SM (227 KB), each
// warp on its own scheduler.
//
// The right-hand side of a stage is assembled in REGISTERS, solved in registers (cbm4_cell::lu_solve_reg: the
// operations of DGETRS('N') in netlib's order, quotients by U_kk in Markstein's correctly rounded three-operation form)
// and stored to tensor memory once: per component the same operations in the same order as before, so state, step
// sizes, counters and tape are bit-identical to k_ros_fwd_cell and to the oracle.
//
// tcgen05.ld/st are warp-wide (.sync.aligned): every thread of the warp executes them, converged.  The per-lane
// control flow of k_ros_fwd_cell (lanes at different points of their step loops, lanes without a cell) therefore
// becomes predication: `go` guards everything that touches the lane's shared-memory slot or global memory, the tensor
// memory accesses sit between the guarded blocks behind a __syncwarp().  Lanes 29..31 of a warp have no shared-memory
// slot and never compute.
#pragma once
#include "ros_cell_fwd.cuh"
#include "tmem.cuh"

namespace fatode {

constexpr int FWD_TM_WARPS = 2, FWD_TM_LANES = 29;
constexpr uint32_t FTM_K = 0, FTM_F0 = 384, FTM_DF = 448;      // tensor-memory columns (two per double): K_j at 64 j

template <class Model, int LANES>
struct CellSmemTM {
  static constexpr int N_ = Model::N;
  static constexpr int OFF_Y = 0, OFF_YN = N_, OFF_FC = 2 * N_, OFF_DINV = 3 * N_, OFF_LU = 4 * N_, DOUBLES = OFF_LU + Model::NLU;
  static constexpr size_t WARP_BYTES = sizeof(double) * DOUBLES * LANES;
  static constexpr size_t BYTES = WARP_BYTES * FWD_TM_WARPS;
};

template <class Model, int MODE, int LANES>
__global__ void __launch_bounds__(FWD_TM_WARPS * 32, 1)
k_ros_fwd_cell_tm(RosParams rp, typename Model::Const mc, long nwave, const int* __restrict__ ids, const double* __restrict__ y_in,
                  double* __restrict__ y_out, const double* __restrict__ atol_g, const double* __restrict__ rtol_g,
                  int* __restrict__ istatus, double* __restrict__ rstatus, int* __restrict__ ierr_out, int* __restrict__ nacc_out,
                  CellTape tape, int with_mu, int debug_irregular_every, unsigned long long* __restrict__ queue) {
  extern __shared__ __align__(16) double cell_smem[];
  __shared__ uint32_t tmem_base_s;
  constexpr int N = Model::N;
  static_assert(N == 32 && LANES <= 32, "tensor-memory layout: 32-vectors");
  constexpr int SV = LANES;
  typedef CellSmemTM<Model, LANES> L;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) tmem::alloc512(&tmem_base_s);
  tmem::fence_before_sync();
  __syncthreads();
  tmem::fence_after_sync();
  const uint32_t tm = tmem_base_s + ((uint32_t)(warp * 32) << 16);
  const bool act = lane < LANES;
  const int S = rp.S;
  const bool adj = (rp.family == FAM_ADJ);
  const bool count_la = (rp.family != FAM_FWD);
  const double Roundoff = rp.roundoff;
  const double Tend = rp.tend;
  const int Direction = (Tend >= rp.tstart) ? 1 : -1;
  const double dDir = (double)Direction;
  double* const base = cell_smem + (size_t)warp * L::DOUBLES * LANES + (act ? lane : 0);
  double* const y = base + L::OFF_Y * SV;       // element i at y[i * SV]
  double* const Ynew = base + L::OFF_YN * SV;
  double* const Fcn = base + L::OFF_FC * SV;
  double* const dinv = base + L::OFF_DINV * SV;
  double* const lu = base + L::OFF_LU * SV;
  bool keep_fcn = false;                         // a later stage reuses the function value of stage 1 (NewF false)
  for (int is = 1; is < S; ++is) keep_fcn = keep_fcn || !rp.NewF[is];
  double PHT[Model::NPH], PHS[Model::NPH];
  unsigned swaps = 0;
  int nfun = 0, njac = 0, nstp = 0, nacc = 0, nrej = 0, ndec = 0, nsol = 0, nsng = 0;
  double T = 0.0, H = 0.0, hexit = 0.0, texit = 0.0, hnew_out = 0.0;
  bool RejectLastH = false, RejectMoreH = false, need_setup = true, forced_irregular = false, have = false, done = !act;
  int ierr = 1, cur_chunk = -1;
  long w = 0, cell = 0;

  for (;;) {
    if (!have && !done) {   // next cell for this lane
      w = (long)atomicAdd(queue, 1ull);
      if (w >= nwave) {
        done = true;
      } else {
        cell = ids ? ids[w] : w;
#pragma unroll 4
        for (int i = 0; i < N; ++i) y[i * SV] = y_in[cell * N + i];
        nfun = njac = nstp = nacc = nrej = ndec = nsol = nsng = 0;
        T = rp.tstart;
        hexit = 0.0; texit = 0.0; hnew_out = 0.0;
        H = fmin(fmax(fabs(rp.hmin), fabs(rp.hstart)), fabs(rp.hmax));
        if (fabs(H) <= __dmul_rn(10.0, Roundoff)) H = rp.deltamin_start;
        H = __dmul_rn(dDir, H);
        RejectLastH = false; RejectMoreH = false;
        ierr = 1;
        cur_chunk = -1;
        need_setup = true;
        forced_irregular = debug_irregular_every > 0 && (cell % debug_irregular_every) == 0;   // test hook
        have = true;
      }
    }
    if (__all_sync(FATODE_FULL_MASK, done)) break;
    bool go = have;        // this lane runs an attempt in this iteration
    bool fin = false;      // ... or its cell ends here
    bool setup_now = false;
    if (go) {
      if (forced_irregular && nstp >= 3) { ierr = IERR_IRREGULAR; fin = true; }
      else if (need_setup) {   // top of the reference's time loop (:905-938)
        if (!((Direction > 0 && __dadd_rn(__dsub_rn(T, Tend), Roundoff) <= 0.0) ||
              (Direction < 0 && __dadd_rn(__dsub_rn(Tend, T), Roundoff) <= 0.0))) fin = true;
        else if (nstp > rp.max_no_steps) { ierr = -6; fin = true; }
        else if (__dadd_rn(T, __dmul_rn(rp.tenth, H)) == T || H <= Roundoff) { ierr = -7; fin = true; }
        else {
          if (adj) hexit = H;
          H = fmin(H, fabs(__dsub_rn(Tend, T)));
          Model::photo(T, mc, PHT);
          Model::template fun<SV>(y, PHT, mc, Fcn);        // Fcn0, committed to tensor memory below
          nfun++;
          njac++;                                   // J(T,Y) itself is evaluated inside build_E (same operands)
          if (!rp.autonomous) {                     // ros_FunTimeDerivative
            const double Delta = __dmul_rn(sqrt(Roundoff), fmax(rp.deltamin_fd, fabs(T)));
            Model::photo(__dadd_rn(T, Delta), mc, PHS);
            Model::template fun<SV>(y, PHS, mc, Ynew);     // scratch: Ynew is free here
            nfun++;
            const double rd = __ddiv_rn(1.0, Delta);
#pragma unroll 8
            for (int i = 0; i < N; ++i) Ynew[i * SV] = __dmul_rn(rd, __dadd_rn(Ynew[i * SV], __dmul_rn(-1.0, Fcn[i * SV])));
          }
          need_setup = false;
          setup_now = true;
        }
      }
      if (fin) go = false;
    }
    __syncwarp();
    if (__any_sync(FATODE_FULL_MASK, setup_now)) {   // Fcn0 (and dF/dT) of the lanes that just set a step up -> tensor memory
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        double t[16];
        tmem::ld16(tm + FTM_F0 + 32 * h, t);
        if (setup_now) {
#pragma unroll
          for (int i = 0; i < 16; ++i) t[i] = Fcn[(16 * h + i) * SV];
        }
        __syncwarp();
        tmem::st16(tm + FTM_F0 + 32 * h, t);
        if (!rp.autonomous) {
          tmem::ld16(tm + FTM_DF + 32 * h, t);
          if (setup_now) {
#pragma unroll
            for (int i = 0; i < 16; ++i) t[i] = Ynew[(16 * h + i) * SV];
          }
          __syncwarp();
          tmem::st16(tm + FTM_DF + 32 * h, t);
        }
      }
      tmem::wait_st();
    }
    // ---- one attempt: ros_PrepareMatrix (:1894-1948), the stages, the error estimate (:940-1136)
    if (go) {
      int Nconsecutive = 0;
      int code;
      for (;;) {
        const double ghinv = __ddiv_rn(1.0, __dmul_rn(__dmul_rn(dDir, H), rp.Gamma[0]));
        Model::template build_E<SV>(y, PHT, mc, ghinv, lu);
        code = Model::template lu_factor<SV>(lu, dinv, swaps);
        if (count_la) ndec++;
        if (code != LU_SINGULAR) break;
        nsng++;
        if (++Nconsecutive > 5) break;
        H = __dmul_rn(H, 0.5);
      }
      if (code == LU_IRREGULAR) { ierr = IERR_IRREGULAR; fin = true; go = false; }
      else if (code == LU_SINGULAR) { ierr = -8; fin = true; go = false; }
    }
    const double dH = __dmul_rn(dDir, H);
#pragma unroll 1
    for (int is = 0; is < S; ++is) {
      double v[32];
      if (is > 0 && rp.NewF[is]) {
        // Ynew = Y + sum_j a_ij K_j  (DAXPY per j in the reference: same order per component; a == 0 skipped)
        if (go) {
#pragma unroll
          for (int i = 0; i < N; ++i) v[i] = y[i * SV];
        } else {
#pragma unroll
          for (int i = 0; i < N; ++i) v[i] = 0.0;
        }
        __syncwarp();
#pragma unroll 1
        for (int j = 0; j < is; ++j) {
          const double a = rp.A[is * (is - 1) / 2 + j];
          if (a != 0.0) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              double kk[16];
              tmem::ld16(tm + FTM_K + 64 * j + 32 * h, kk);
#pragma unroll
              for (int i = 0; i < 16; ++i) v[16 * h + i] = __dadd_rn(v[16 * h + i], __dmul_rn(a, kk[i]));
            }
          }
        }
        if (go) {
#pragma unroll
          for (int i = 0; i < N; ++i) Ynew[i * SV] = v[i];
          const double Tau = __dadd_rn(T, __dmul_rn(__dmul_rn(rp.Alpha[is], dDir), H));
          Model::photo(Tau, mc, PHS);
          Model::template fun<SV>(Ynew, PHS, mc, Fcn);
          nfun++;
        }
      }
      // right-hand side: K_is = Fcn + sum_j (c_ij/h) K_j + h*gamma_i*dFdT   (:1021-1036), same order per component
      __syncwarp();
      if (is == 0) {
        double t[16];
        tmem::ld16(tm + FTM_F0, t);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = t[i];
        tmem::ld16(tm + FTM_F0 + 32, t);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[16 + i] = t[i];
        if (keep_fcn && go) {
#pragma unroll
          for (int i = 0; i < N; ++i) Fcn[i * SV] = v[i];
        }
      } else if (go) {
#pragma unroll
        for (int i = 0; i < N; ++i) v[i] = Fcn[i * SV];
      } else {
#pragma unroll
        for (int i = 0; i < N; ++i) v[i] = 0.0;
      }
      __syncwarp();
#pragma unroll 1
      for (int j = 0; j < is; ++j) {
        const double c = rp.C[is * (is - 1) / 2 + j];
        if (c != 0.0) {                              // c/h == 0 exactly when c == 0 (h finite, non-zero)
          const double hc = __ddiv_rn(c, dH);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            double kk[16];
            tmem::ld16(tm + FTM_K + 64 * j + 32 * h, kk);
#pragma unroll
            for (int i = 0; i < 16; ++i) v[16 * h + i] = __dadd_rn(v[16 * h + i], __dmul_rn(hc, kk[i]));
          }
        }
      }
      if (!rp.autonomous && rp.Gamma[is] != 0.0) {
        const double HG = __dmul_rn(dH, rp.Gamma[is]);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          double kk[16];
          tmem::ld16(tm + FTM_DF + 32 * h, kk);
#pragma unroll
          for (int i = 0; i < 16; ++i) v[16 * h + i] = __dadd_rn(v[16 * h + i], __dmul_rn(HG, kk[i]));
        }
      }
      if (go) {
        Model::template lu_solve_reg<SV>(lu, dinv, swaps, v);
        if (count_la) nsol++;
      }
      __syncwarp();
      {
        double t[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) t[i] = v[i];
        tmem::st16(tm + FTM_K + 64 * is, t);
#pragma unroll
        for (int i = 0; i < 16; ++i) t[i] = v[16 + i];
        tmem::st16(tm + FTM_K + 64 * is + 32, t);
        tmem::wait_st();
      }
    }
    // ---- new solution, error estimate, ros_ErrorNorm (sequential sum i = 1..N)
    double acc = 0.0;
    __syncwarp();
#pragma unroll 1
    for (int h = 0; h < 2; ++h) {
      double yn[16], ye[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) { yn[i] = go ? y[(16 * h + i) * SV] : 0.0; ye[i] = 0.0; }
      __syncwarp();
#pragma unroll 1
      for (int j = 0; j < S; ++j) {
        const double cm = rp.M[j], ce = rp.E[j];
        if (cm != 0.0 || ce != 0.0) {
          double kk[16];
          tmem::ld16(tm + FTM_K + 64 * j + 32 * h, kk);
          if (cm != 0.0) {
#pragma unroll
            for (int i = 0; i < 16; ++i) yn[i] = __dadd_rn(yn[i], __dmul_rn(cm, kk[i]));
          }
          if (ce != 0.0) {
#pragma unroll
            for (int i = 0; i < 16; ++i) ye[i] = __dadd_rn(ye[i], __dmul_rn(ce, kk[i]));
          }
        }
      }
      if (go) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int c = 16 * h + i;
          const double yi = y[c * SV];
          Ynew[c * SV] = yn[i];
          const int ti = rp.vector_tol ? c : 0;
          const double Ymax = fmax(fabs(yi), fabs(yn[i]));
          const double Scale = __dadd_rn(atol_g[ti], __dmul_rn(rtol_g[ti], Ymax));
          const double q = __ddiv_rn(ye[i], Scale);
          acc = __dadd_rn(acc, __dmul_rn(q, q));
        }
      }
    }
    bool accepted = false;
    if (go) {
      const double Err = fmax(sqrt(__ddiv_rn(acc, (double)N)), 1.0e-10);
      const double Fac = fmin(rp.facmax, fmax(rp.facmin, __ddiv_rn(rp.facsafe, fpm::root_elo(Err, rp.ELO))));
      double Hnew = __dmul_rn(H, Fac);
      nstp++;
      if (Err <= 1.0 || H <= rp.hmin) {
        nacc++;
        accepted = true;
        if (MODE == 1) {
          // ---- ros_DPush, fat form: chunk bookkeeping, then T, H, L\U, 1/U_kk, Y(T) (K follows from tensor memory below)
          const int slot = (nacc - 1) % FT_CHUNK;
          if (slot == 0) {
            const unsigned c = atomicAdd(tape.pool_head, 1u);
            if (c >= tape.pool_chunks) { ierr = IERR_DEFERRED; fin = true; go = false; accepted = false; }
            else {
              tape.chunk_prev[c] = cur_chunk;
              tape.chunk_owner[c] = (int)w;
              tape.chunk_seq[c] = (nacc - 1) / FT_CHUNK;
              if (tape.chunk_next) {
                if (cur_chunk >= 0) tape.chunk_next[cur_chunk] = (int)c; else tape.first_chunk[w] = (int)c;
              }
              cur_chunk = (int)c;
            }
          }
          if (accepted) {
            double* e = tape.base + ((size_t)cur_chunk * FT_CHUNK + slot) * FT_STEP;
            st_global_256(e, T, H, (double)swaps, 0.0);
            double* el = e + FT_HDR;
#pragma unroll 4
            for (int q = 0; q + 3 < Model::NLU; q += 4) st_global_256(el + q, lu[q * SV], lu[(q + 1) * SV], lu[(q + 2) * SV], lu[(q + 3) * SV]);
            {
              constexpr int q = (Model::NLU / 4) * 4;   // tail (NLU = 367: three values + pad)
              st_global_256(el + q, q < Model::NLU ? lu[q * SV] : 0.0, q + 1 < Model::NLU ? lu[(q + 1) * SV] : 0.0,
                            q + 2 < Model::NLU ? lu[(q + 2) * SV] : 0.0, 0.0);
            }
#pragma unroll 4
            for (int q = 0; q < N; q += 4)
              st_global_256(e + FT_HDR + FT_LU + q, dinv[q * SV], dinv[(q + 1) * SV], dinv[(q + 2) * SV], dinv[(q + 3) * SV]);
#pragma unroll 4
            for (int q = 0; q < N; q += 4) st_global_256(e + FT_Y0 + q, y[q * SV], y[(q + 1) * SV], y[(q + 2) * SV], y[(q + 3) * SV]);
          }
        }
        if (accepted) {
#pragma unroll 8
          for (int i = 0; i < N; ++i) y[i * SV] = Ynew[i * SV];
          T = __dadd_rn(T, dH);
          Hnew = fmax(rp.hmin, fmin(Hnew, rp.hmax));
          if (RejectLastH) Hnew = fmin(Hnew, H);
          hexit = H; hnew_out = Hnew; texit = T;
          RejectLastH = false; RejectMoreH = false;
          H = Hnew;
          need_setup = true;
        }
      } else {
        if (RejectMoreH) Hnew = __dmul_rn(H, rp.facrej);
        RejectMoreH = RejectLastH;
        RejectLastH = true;
        H = Hnew;
        if (nacc >= 1) nrej++;
      }
    }
    __syncwarp();
    if (MODE == 1 && __any_sync(FATODE_FULL_MASK, accepted)) {   // the stage vectors of the accepted steps -> tape
      double* const ek = accepted ? tape.base + ((size_t)cur_chunk * FT_CHUNK + (nacc - 1) % FT_CHUNK) * FT_STEP + FT_K : nullptr;
#pragma unroll 1
      for (int j = 0; j < S; ++j) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          double kk[16];
          tmem::ld16(tm + FTM_K + 64 * j + 32 * h, kk);
          if (accepted) {
#pragma unroll
            for (int q = 0; q < 16; q += 4) st_global_256(ek + j * N + 16 * h + q, kk[q], kk[q + 1], kk[q + 2], kk[q + 3]);
          }
          __syncwarp();
        }
      }
    }
    if (fin) {   // ---- results of this cell
      ierr_out[cell] = ierr;
      if (ierr != IERR_IRREGULAR && ierr != IERR_DEFERRED) {   // otherwise untouched outputs: the cell is re-run elsewhere
#pragma unroll 4
        for (int i = 0; i < N; ++i) y_out[cell * N + i] = y[i * SV];
        int* is_ = istatus + cell * 20;
        for (int q = 0; q < 20; ++q) is_[q] = 0;
        is_[Nfun] = nfun; is_[Njac] = njac; is_[Nstp] = nstp; is_[Nacc] = nacc; is_[Nrej] = nrej;
        is_[Ndec] = ndec; is_[Nsol] = nsol; is_[Nsng] = nsng;
        double* rs = rstatus + cell * 20;
        for (int q = 0; q < 20; ++q) rs[q] = 0.0;
        rs[Ntexit] = texit; rs[Nhexit] = hexit; rs[Nhnew] = hnew_out;
        nacc_out[w] = nacc;
        if (MODE == 1) tape.last_chunk[w] = cur_chunk;
      }
      have = false;
    }
    __syncwarp();
  }
  tmem::fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem::dealloc512(tmem_base_s);
}

}  // namespace fatode
