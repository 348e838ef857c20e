// Rosenbrock forward sweep, ONE CELL PER THREAD (the fast path of INTEGRATE / INTEGRATE_ADJ's forward part).
//
// Replaces, for models that ship straight-line per-thread code (gen/<model>_cell_gen.h):
//   ros_FwdInt            ADJ/ROS_ADJ/ROS_ADJ_f90_Integrator.F90:877-1173   (ros_Integrator FWD/ROS :433-629)
//   ros_FunTimeDerivative :1870-1890,  ros_PrepareMatrix :1894-1948,  ros_ErrorNorm :1837-1866
//   lapack_decomp/solve   ADJ/ROS_ADJ/LS_Solver.F90:31-54  (netlib DGETF2 / DGETRS('N'))
//   ros_DPush             :735-755  (the trajectory tape)
// Arithmetic and operation order are the reference's (no FMA contraction in this translation unit), so the
// states, step sizes and ISTATUS counters are bit-identical to the warp-per-cell kernel and to the oracle.
//
// Mapping: 32 cells per warp, each thread runs the whole scalar algorithm on its own cell with the state,
// stage vectors and the sparse L\U factors in its own slots of shared memory (6 KB per cell, one warp per SM).  Compared with the warp-per-cell kernel this needs ~30x fewer warp instructions per cell.
// Divergence: the nested "step / until accepted" loops are flattened into ONE loop over attempts (a lane whose
// attempt was rejected skips the step set-up of its neighbours instead of making them wait), so a warp runs
// max-over-lanes(attempts) iterations (+-4 % for CBM-IV).
//
// The LU is DGETF2 restricted to the pivot interchanges observed along CBM-IV trajectories (see
// tools/gen_cbm4_cell.py).  A cell that asks for any other pivot, meets a singular/tiny pivot or needs the
// step-halving retry of ros_PrepareMatrix leaves this kernel with IERR_IRREGULAR and is integrated by the general
// warp-per-cell kernel (full pivoting) instead; the two paths give identical bits.
//
// Adjoint mode writes the trajectory tape: per accepted step T, H, the sparse L\U factors (reused by the backward
// sweep — the reference refactorises the same matrix, :1261-1267), Y(T) and the stage vectors K.  Tape space comes
// from a chunk pool (FT_CHUNK accepted steps per chunk, linked backwards), so no per-cell slab has to be guessed; a
// cell that finds the pool empty is deferred to the next wave (IERR_DEFERRED).  k_expand_tape (ros_cell_adj.cuh)
// later derives from (Y, K) the per-stage operands of the backward sweep that do not depend on the adjoint
// variables — J(T_i,Y_i), M'_i = Hess x K_i + h*gamma_i*dJ/dt, a_p — into the "fat" stage blocks.
#pragma once
#include "fatode_common.cuh"
#include "model_cbm4.cuh"
#include "portable_math.h"

namespace fatode {

// 32-byte store (one full DRAM sector) to 32-byte aligned global memory: STG.256
__device__ __forceinline__ void st_global_256(double* p, double a, double b, double c, double d) {
  asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}

__constant__ double c_cbm4_rc[81];     // RCONST with the time-independent entries (set per call by the host)
__constant__ double c_cbm4_hess[258];  // Hessian values: rate coefficients only for this mechanism (set per call)

}  // namespace fatode

#define CBM4_CELL_FN __device__ __noinline__
#define CBM4_CELL_INL __device__ __forceinline__
#define CBM4_CELL_RC(r) fatode::c_cbm4_rc[r]
#define CBM4_CELL_HESS(h) fatode::c_cbm4_hess[h]
#define CBM4_CELL_ST4(p, a, b, c, d) fatode::st_global_256(p, a, b, c, d)
#include "gen/cbm4_cell_gen.h"

namespace fatode {

constexpr int IERR_TAPE_OVERFLOW = -1000;   // internal (general path): slab too small, the wave is re-run in exact mode
constexpr int IERR_IRREGULAR = -1001;       // internal: the cell needs the general (full pivoting) kernel
constexpr int IERR_DEFERRED = -1002;        // internal: tape pool exhausted, the cell is re-run in the next wave

// ---- fat tape geometry (doubles) -----------------------------------------------------------------------------
constexpr int FT_HDR = 8;                   // T, H, swaps, (pad)
constexpr int FT_LU = 368;                  // NLU = 367 (+1 pad)
constexpr int FT_DINV = 32;
constexpr int FT_STEP_COPY = FT_HDR + FT_LU + FT_DINV;     // what the backward sweep loads per step: 408 doubles = 3264 B
constexpr int FT_Y0 = FT_STEP_COPY;                         // Y(T) [32]   } inputs of k_expand_tape
constexpr int FT_K = FT_Y0 + 32;                            // K [S][32]   }
constexpr int FT_STEP = FT_K + 6 * 32;                      // forward-written part of an entry: 632 doubles
constexpr int FT_JV = 0, FT_MV = 276, FT_AP = 536;          // inside a stage block (each part 32-byte aligned, zero padded)
constexpr int FT_STAGE = 568;                               // 4544 B (multiple of 32)
constexpr int FT_CHUNK = 16;                                // accepted steps per chunk
__host__ __device__ inline int fat_step_doubles(int S) { return S * FT_STAGE; }   // stage blocks of one accepted step
static_assert(cbm4_cell::NLU <= FT_LU && cbm4_cell::NJ == FT_MV && cbm4_cell::NJ + cbm4_cell::NM <= FT_AP && FT_AP % 4 == 0 &&
              FT_STAGE % 4 == 0 && FT_STEP % 4 == 0 && FT_MV % 4 == 0, "tape layout");

// ---- model policy of the per-thread kernels -----------------------------------------------------------------------
//   N, NLU (doubles of the factor storage), NPH (doubles of the time-dependent context), Const
//   photo(t, mc, PH)            time-dependent rate data at time t
//   fun<SV>(y, PH, mc, f)       f = FUN(t, y)
//   build_E<SV>(y, PH, mc, ghinv, lu)   E = ghinv*I - JAC(t, y) in the model's factor storage
//   lu_factor<SV>(lu, dinv, swaps) -> LU_OK | LU_SINGULAR (zero pivot: ros_PrepareMatrix halves H) | LU_IRREGULAR
//   lu_solve<SV>(lu, swaps, b)  DGETRS('N')
//   build_EZ / zlu_factor / zlu_solve: the same for the complex matrix -J + (a + i b) I (ZGETF2 / ZGETRS('N'), two planes)
// All arrays are thread-private with element stride SV.
enum { LU_OK = 0, LU_SINGULAR = 1, LU_IRREGULAR = 2 };

struct Cbm4Cell {
  static constexpr int N = 32, NP = 29, NLU = cbm4_cell::NLU, NPH = cbm4_cell::NPHOTO;
  static constexpr bool HAS_GENERAL_PATH = true;   // irregular cells are handed to the warp-per-cell kernels
  typedef Cbm4Const Const;
  template <int SV>
  __device__ static void build_E(const double* y, const double* PH, const Const& mc, double ghinv, double* lu) {
    cbm4_cell::build_E<SV>(y, mc.fix, PH, ghinv, lu);
  }
  template <int SV>
  __device__ static int lu_factor(double* lu, double* dinv, unsigned& swaps) {
    return cbm4_cell::lu_factor<SV>(lu, dinv, swaps) ? LU_IRREGULAR : LU_OK;   // singular matrices included
  }
  template <int SV>
  __device__ static void lu_solve(const double* lu, unsigned swaps, double* b) { cbm4_cell::lu_solve<SV>(lu, swaps, b); }
  template <int SV>   // the same solve with the right-hand side in registers (k_ros_fwd_cell_tm)
  __device__ __forceinline__ static void lu_solve_reg(const double* lu, const double* dinv, unsigned swaps, double (&b)[32]) {
    cbm4_cell::lu_solve_reg<SV>(lu, dinv, swaps, b);
  }
  // complex matrix -J + (a + i b) I of the implicit Runge-Kutta families in two planes (FWD/RK/LS_Solver.F90:22-32)
  template <int SV>
  __device__ static void build_EZ(const double* y, const double* PH, const Const& mc, double a, double b, double* zr, double* zi) {
    cbm4_cell::build_E<SV>(y, mc.fix, PH, a, zr);
    cbm4_cell::zimag_init<SV>(zi, b);
  }
  template <int SV>
  __device__ static int zlu_factor(double* zr, double* zi, unsigned& swaps) {
    return cbm4_cell::zlu_factor<SV>(zr, zi, swaps) ? LU_IRREGULAR : LU_OK;
  }
  template <int SV>
  __device__ static void zlu_solve(const double* zr, const double* zi, unsigned swaps, double* br, double* bi) {
    cbm4_cell::zlu_solve<SV>(zr, zi, swaps, br, bi);
  }
  __device__ static void photo(double t, const Const&, double* PH) {       // Update_SUN + photolysis part of Update_RCONST
    const double s = Cbm4Warp::sun(t);
#pragma unroll
    for (int k = 0; k < NPH; ++k) PH[k] = __dmul_rn(cbm4_dev::PHOTO_C[k], s);   // same order as cbm4_cell::PHOTO_IDX
  }
  template <int SV>
  __device__ static void fun(const double* y, const double* PH, const Const& mc, double* f) {   // dr.F90:83-98
    cbm4_cell::fun<SV>(y, mc.fix, PH, f);
    f[30 * SV] = __dadd_rn(f[30 * SV], mc.emis);
    f[25 * SV] = __dadd_rn(f[25 * SV], mc.emis);
  }
};

struct CellTape {
  double* base;                 // chunk pool of step entries (FT_STEP doubles per accepted step)
  int* chunk_prev;              // [pool_chunks] backward links
  int* chunk_owner;             // [pool_chunks] wave slot that owns the chunk
  int* chunk_seq;               // [pool_chunks] position of the chunk in its cell's trajectory (0, 1, ...)
  int* chunk_next;              // [pool_chunks] forward links (tangent-linear sweep); may be nullptr
  int* first_chunk;             // [cells of the wave]; may be nullptr
  unsigned int* pool_head;      // next free chunk
  unsigned int pool_chunks;
  int* last_chunk;              // [cells of the wave]
};

// Per-thread storage: the hot arrays live in SHARED memory, element e of a thread's array at base[e * LANES]
// (consecutive lanes -> consecutive doubles: conflict-free), 751 doubles = 6 KB per cell.  One block (= one warp, LANES
// of its 32 threads active) per SM: 188 KB with 32 cells; 26 cells (153 KB) leave room for a CTA of the backward
// kernel on the same SM, which is how the forward sweep of the next wave hides behind the backward sweep of this one.
// Thread-local memory only holds routine temporaries and spills.
constexpr int CELL_LANES_SOLO = 32, CELL_LANES_SHARED = 26;
template <class Model, int LANES>
struct CellSmem {
  static constexpr int N_ = Model::N;
  static constexpr int OFF_Y = 0, OFF_F0 = N_, OFF_DF = 2 * N_, OFF_YN = 3 * N_, OFF_FC = 4 * N_, OFF_DINV = 5 * N_, OFF_K = 6 * N_,
                       OFF_LU = 12 * N_, DOUBLES = OFF_LU + Model::NLU;
  static constexpr size_t BYTES = sizeof(double) * DOUBLES * LANES;
};

// MODE 0: forward only.  MODE 1: forward + fat tape (adjoint mode).
// Persistent kernel: every LANE pulls its next cell from an atomic queue as soon as its current one is done
// (slot w of the wave = cell ids[w], or cell w when ids == nullptr), so there is no max-over-lanes tail inside a
// warp.
template <class Model, int MODE, int LANES>
__global__ void __launch_bounds__(32, 1)
k_ros_fwd_cell(RosParams rp, typename Model::Const mc, long nwave, const int* __restrict__ ids, const double* __restrict__ y_in,
               double* __restrict__ y_out, const double* __restrict__ atol_g, const double* __restrict__ rtol_g,
               int* __restrict__ istatus, double* __restrict__ rstatus, int* __restrict__ ierr_out, int* __restrict__ nacc_out,
               CellTape tape, int with_mu, int debug_irregular_every, unsigned long long* __restrict__ queue) {
  extern __shared__ __align__(16) double cell_smem[];
  constexpr int N = Model::N;
  constexpr int SV = LANES;
  typedef CellSmem<Model, LANES> L;
  if (threadIdx.x >= LANES) return;
  const int S = rp.S;
  const bool adj = (rp.family == FAM_ADJ);
  const bool count_la = (rp.family != FAM_FWD);
  const double Roundoff = rp.roundoff;
  const double Tend = rp.tend;
  const int Direction = (Tend >= rp.tstart) ? 1 : -1;
  const double dDir = (double)Direction;
  double* const base = cell_smem + threadIdx.x;
  double* const y = base + L::OFF_Y * SV;       // element i at y[i * SV]
  double* const Fcn0 = base + L::OFF_F0 * SV;
  double* const dFdT = base + L::OFF_DF * SV;
  double* const Ynew = base + L::OFF_YN * SV;
  double* const Fcn = base + L::OFF_FC * SV;
  double* const dinv = base + L::OFF_DINV * SV;
  double* const K = base + L::OFF_K * SV;       // K[(is * N + i) * SV]
  double* const lu = base + L::OFF_LU * SV;
  double PHT[Model::NPH], PHS[Model::NPH];
  unsigned swaps = 0;
  int nfun = 0, njac = 0, nstp = 0, nacc = 0, nrej = 0, ndec = 0, nsol = 0, nsng = 0;
  double T = 0.0, H = 0.0, hexit = 0.0, texit = 0.0, hnew_out = 0.0;
  bool RejectLastH = false, RejectMoreH = false, need_setup = true, forced_irregular = false, have = false;
  int ierr = 1, cur_chunk = -1;
  long w = 0, cell = 0;

  for (;;) {
    if (!have) {   // next cell for this lane
      w = (long)atomicAdd(queue, 1ull);
      if (w >= nwave) break;
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
    bool fin = true;   // the cell ends in this iteration unless the attempt runs to its accept/reject decision
    do {
      if (forced_irregular && nstp >= 3) { ierr = IERR_IRREGULAR; break; }
      if (need_setup) {   // top of the reference's time loop (:905-938)
        if (!((Direction > 0 && __dadd_rn(__dsub_rn(T, Tend), Roundoff) <= 0.0) ||
              (Direction < 0 && __dadd_rn(__dsub_rn(Tend, T), Roundoff) <= 0.0)))
          break;
        if (nstp > rp.max_no_steps) { ierr = -6; break; }
        if (__dadd_rn(T, __dmul_rn(rp.tenth, H)) == T || H <= Roundoff) { ierr = -7; break; }
        if (adj) hexit = H;
        H = fmin(H, fabs(__dsub_rn(Tend, T)));
        Model::photo(T, mc, PHT);
        Model::template fun<SV>(y, PHT, mc, Fcn0);
        nfun++;
        njac++;                                   // J(T,Y) itself is evaluated inside build_E (same operands)
        if (!rp.autonomous) {                     // ros_FunTimeDerivative
          const double Delta = __dmul_rn(sqrt(Roundoff), fmax(rp.deltamin_fd, fabs(T)));
          Model::photo(__dadd_rn(T, Delta), mc, PHS);
          Model::template fun<SV>(y, PHS, mc, dFdT);
          nfun++;
          const double rd = __ddiv_rn(1.0, Delta);
#pragma unroll 8
          for (int i = 0; i < N; ++i) dFdT[i * SV] = __dmul_rn(rd, __dadd_rn(dFdT[i * SV], __dmul_rn(-1.0, Fcn0[i * SV])));
        }
        need_setup = false;
      }
      // ---- one attempt: ros_PrepareMatrix (:1894-1948), the stages, the error estimate (:940-1136)
      {
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
        if (code == LU_IRREGULAR) { ierr = IERR_IRREGULAR; break; }
        if (code == LU_SINGULAR) { ierr = -8; break; }
      }
      const double dH = __dmul_rn(dDir, H);
      for (int is = 0; is < S; ++is) {
        double* k = K + is * N * SV;
        if (is == 0) {
#pragma unroll 8
          for (int i = 0; i < N; ++i) Fcn[i * SV] = Fcn0[i * SV];
        } else if (rp.NewF[is]) {
          // Ynew = Y + sum_j a_ij K_j  (DAXPY per j in the reference: same order per component; a == 0 skipped)
          double ca[5];
#pragma unroll
          for (int j = 0; j < 5; ++j) ca[j] = (j < is) ? rp.A[is * (is - 1) / 2 + j] : 0.0;
#pragma unroll 4
          for (int i = 0; i < N; ++i) {
            double v = y[i * SV];
#pragma unroll
            for (int j = 0; j < 5; ++j)
              if (ca[j] != 0.0) v = __dadd_rn(v, __dmul_rn(ca[j], K[(j * N + i) * SV]));
            Ynew[i * SV] = v;
          }
          const double Tau = __dadd_rn(T, __dmul_rn(__dmul_rn(rp.Alpha[is], dDir), H));
          Model::photo(Tau, mc, PHS);
          Model::template fun<SV>(Ynew, PHS, mc, Fcn);
          nfun++;
        }
        {   // right-hand side: K_is = Fcn + sum_j (c_ij/h) K_j + h*gamma_i*dFdT   (:1021-1036), same order per component
          double hc[5];
#pragma unroll
          for (int j = 0; j < 5; ++j) hc[j] = (j < is) ? __ddiv_rn(rp.C[is * (is - 1) / 2 + j], dH) : 0.0;
          const bool tterm = !rp.autonomous && rp.Gamma[is] != 0.0;
          const double HG = __dmul_rn(dH, rp.Gamma[is]);
#pragma unroll 4
          for (int i = 0; i < N; ++i) {
            double v = Fcn[i * SV];
#pragma unroll
            for (int j = 0; j < 5; ++j)
              if (hc[j] != 0.0) v = __dadd_rn(v, __dmul_rn(hc[j], K[(j * N + i) * SV]));
            if (tterm) v = __dadd_rn(v, __dmul_rn(HG, dFdT[i * SV]));
            k[i * SV] = v;
          }
        }
        Model::template lu_solve<SV>(lu, swaps, k);
        if (count_la) nsol++;
      }
      // new solution, error estimate, ros_ErrorNorm (sequential sum i = 1..N)
      double acc = 0.0;
      {
        double cm[6], ce[6];
#pragma unroll
        for (int j = 0; j < 6; ++j) { cm[j] = (j < S) ? rp.M[j] : 0.0; ce[j] = (j < S) ? rp.E[j] : 0.0; }
#pragma unroll 2
        for (int i = 0; i < N; ++i) {
          const double yi = y[i * SV];
          double yn = yi, ye = 0.0;
#pragma unroll
          for (int j = 0; j < 6; ++j) {
            if (cm[j] != 0.0 || ce[j] != 0.0) {
              const double kj = K[(j * N + i) * SV];
              if (cm[j] != 0.0) yn = __dadd_rn(yn, __dmul_rn(cm[j], kj));
              if (ce[j] != 0.0) ye = __dadd_rn(ye, __dmul_rn(ce[j], kj));
            }
          }
          Ynew[i * SV] = yn;
          const int ti = rp.vector_tol ? i : 0;
          const double Ymax = fmax(fabs(yi), fabs(yn));
          const double Scale = __dadd_rn(atol_g[ti], __dmul_rn(rtol_g[ti], Ymax));
          const double q = __ddiv_rn(ye, Scale);
          acc = __dadd_rn(acc, __dmul_rn(q, q));
        }
      }
      const double Err = fmax(sqrt(__ddiv_rn(acc, (double)N)), 1.0e-10);
      const double Fac = fmin(rp.facmax, fmax(rp.facmin, __ddiv_rn(rp.facsafe, fpm::root_elo(Err, rp.ELO))));
      double Hnew = __dmul_rn(H, Fac);
      nstp++;
      if (Err <= 1.0 || H <= rp.hmin) {
        nacc++;
        if (MODE == 1) {
          // ---- ros_DPush, fat form
          const int slot = (nacc - 1) % FT_CHUNK;
          if (slot == 0) {
            const unsigned c = atomicAdd(tape.pool_head, 1u);
            if (c >= tape.pool_chunks) { ierr = IERR_DEFERRED; break; }
            tape.chunk_prev[c] = cur_chunk;
            tape.chunk_owner[c] = (int)w;
            tape.chunk_seq[c] = (nacc - 1) / FT_CHUNK;
            if (tape.chunk_next) {
              if (cur_chunk >= 0) tape.chunk_next[cur_chunk] = (int)c; else tape.first_chunk[w] = (int)c;
            }
            cur_chunk = (int)c;
          }
          double* e = tape.base + ((size_t)cur_chunk * FT_CHUNK + slot) * FT_STEP;
          st_global_256(e, T, H, (double)swaps, 0.0);
          double* el = e + FT_HDR;
          static_assert(MODE == 0 || Model::N == 32, "the tape layout is the 32-species one");
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
          // Y(T) and the stage vectors K: k_expand_tape derives the per-stage blocks from them
#pragma unroll 4
          for (int q = 0; q < N; q += 4) st_global_256(e + FT_Y0 + q, y[q * SV], y[(q + 1) * SV], y[(q + 2) * SV], y[(q + 3) * SV]);
          for (int j = 0; j < S; ++j) {
            const double* kj = K + j * N * SV;
#pragma unroll 4
            for (int q = 0; q < N; q += 4)
              st_global_256(e + FT_K + j * N + q, kj[q * SV], kj[(q + 1) * SV], kj[(q + 2) * SV], kj[(q + 3) * SV]);
          }
        }
#pragma unroll 8
        for (int i = 0; i < N; ++i) y[i * SV] = Ynew[i * SV];
        T = __dadd_rn(T, dH);
        Hnew = fmax(rp.hmin, fmin(Hnew, rp.hmax));
        if (RejectLastH) Hnew = fmin(Hnew, H);
        hexit = H; hnew_out = Hnew; texit = T;
        RejectLastH = false; RejectMoreH = false;
        H = Hnew;
        need_setup = true;
      } else {
        if (RejectMoreH) Hnew = __dmul_rn(H, rp.facrej);
        RejectMoreH = RejectLastH;
        RejectLastH = true;
        H = Hnew;
        if (nacc >= 1) nrej++;
      }
      fin = false;
    } while (0);
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
  }
}

}  // namespace fatode
